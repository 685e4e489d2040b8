// jpeg_gpuenc.cu -- the block-parallel JPEG entropy encoder on the device.  Every pass is a thin
// CUDA wrapper around a body from jpeg_gpuenc_core.h (the same bodies tests/emul/ runs serially on the CPU); the
// cross-block dependencies of jchuff.c / jcphuff.c (DC prediction, EOB runs, buffered correction bits, bit positions,
// 0xFF stuffing) are resolved with prefix scans (CUB DeviceScan -- plumbing, not one of the path's named kernels).
//
// Passes (one launch each for ANY number of images x scans; blockIdx.y = scan):
//   classify -> [max-scan: previous event] [sum-scan: trailing correction bits] -> groups -> histogram -> tables ->
//   lengths -> [sum-scan: bit offsets] -> scan sizes / buffer layout (on the device) -> zero -> emit -> ffcount -> [sum-scan] ->
//   layout -> scatter (byte stuffing) | D2H: stuffed scans + DHT payloads.  No host wait in between: see "host orchestration".
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <chrono>
#include <cstring>
#include "jpeg_gpuenc.h"
#include "stream_wait.h"
#include "launch_timer.h"

namespace b200 {

using namespace ge;

#define CU(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { err = std::string(#expr) + ": " + cudaGetErrorString(e_); return false; } } while (0)

// ---- kernels ------------------------------------------------------------------------------------------------------
__global__ void k_ge_groups(const Scan *__restrict__ scans, const uint32_t *__restrict__ meta, const int *__restrict__ evkey,
                            const int *__restrict__ prev, const uint32_t *__restrict__ tsum, uint32_t *__restrict__ gcount)
{
    const Scan s = scans[blockIdx.y];
    if (s.mode != MODE_AC_FIRST && s.mode != MODE_AC_REFINE) return;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > s.nblocks) return;
    int pg;
    if (b < s.nblocks) { if (!meta_event(meta[s.unit_base + b])) return; pg = prev[s.unit_base + b]; }
    else { const long long last = s.unit_base + s.nblocks - 1; pg = max(prev[last], evkey[last]); }
    const int pl = pg >= s.unit_base ? (int)(pg - s.unit_base) : -1;
    assign_groups(meta + s.unit_base, tsum + s.unit_base, s.nblocks, pl, b, gcount + s.unit_base);
}

// jchuff.c jpeg_gen_optimal_table with the two minimum searches spread over a warp (ties resolve to the LARGEST index,
// exactly like the sequential `<=` scans); the chain merges and the canonical code assignment stay on lane 0.
__global__ void k_ge_tables(const uint32_t *__restrict__ hist, Table *__restrict__ tabs, DhtOut *__restrict__ dht)
{
    __shared__ long long freq[257];
    __shared__ int codesize[257], others[257];
    __shared__ uint8_t bits[33];
    const int t = blockIdx.x, lane = threadIdx.x;
    const uint32_t *f = hist + (size_t)t * 256;
    for (int i = lane; i < 257; i += 32) { freq[i] = i < 256 ? (long long)f[i] : 1; codesize[i] = 0; others[i] = -1; }
    if (lane == 0) for (int i = 0; i < 33; i++) bits[i] = 0;
    __syncwarp();
    // argmin over the live entries (freq > 0), ties to the LARGEST index, entries above 10^9 never chosen (jchuff.c's `v = 1000000000L`
    // start value): each lane scans its 9 entries in ascending order, then three warp reductions (REDUX) pick the winner --
    // high word, low word, index -- instead of a five-step shuffle butterfly on (value, index) pairs.
    auto warp_argmin = [&](int exclude) {
        unsigned long long v = 1000000000ULL; int c = -1;
        for (int i = lane; i < 257; i += 32) { const unsigned long long fi = (unsigned long long)freq[i]; if (fi && fi <= v && i != exclude) { v = fi; c = i; } }
        const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
        const unsigned mhi = __reduce_min_sync(0xFFFFFFFFu, hi);
        const unsigned mlo = __reduce_min_sync(0xFFFFFFFFu, hi == mhi ? lo : 0xFFFFFFFFu);
        return __reduce_max_sync(0xFFFFFFFFu, (hi == mhi && lo == mlo) ? c : -1);
    };
    for (;;) {
        const int c1 = warp_argmin(-1);
        const int c2 = c1 < 0 ? -1 : warp_argmin(c1);
        if (c2 < 0) break;
        if (lane == 0) {
            int a = c1, b = c2;
            freq[a] += freq[b]; freq[b] = 0;
            codesize[a]++; while (others[a] >= 0) { a = others[a]; codesize[a]++; }
            others[a] = b;
            codesize[b]++; while (others[b] >= 0) { b = others[b]; codesize[b]++; }
        }
        __syncwarp();
    }
    if (lane == 0) {
        Table &T = tabs[t];
        for (int i = 0; i <= 256; i++) if (codesize[i]) bits[codesize[i] > 32 ? 32 : codesize[i]]++;
        for (int i = 32; i > 16; i--) while (bits[i] > 0) {
            int j = i - 2; while (bits[j] == 0) j--;
            bits[i] -= 2; bits[i - 1]++; bits[j + 1] += 2; bits[j]--;
        }
        int i = 16; while (i > 0 && bits[i] == 0) i--;
        if (i > 0) bits[i]--;
        // symbols sorted by (code length, symbol value): counting sort on the UNLIMITED lengths, as the IJG loop orders them
        int start[34];
        for (int l = 0; l < 34; l++) start[l] = 0;
        for (int s = 0; s <= 255; s++) if (codesize[s]) start[min(codesize[s], 32) + 1]++;
        for (int l = 1; l < 34; l++) start[l] += start[l - 1];
        const int p = start[33];
        for (int s = 0; s <= 255; s++) if (codesize[s]) T.vals[start[min(codesize[s], 32)]++] = (uint8_t)s;
        T.nvals = p;
        for (int k = 0; k < 17; k++) T.bits[k] = bits[k];
        for (int s = 0; s < 256; s++) { T.code[s] = 0; T.size[s] = 0; }
        uint32_t code = 0; int k = 0;
        for (int l = 1; l <= 16; l++) { for (int n = 0; n < bits[l]; n++, k++) { T.code[T.vals[k]] = code++; T.size[T.vals[k]] = (uint8_t)l; } code <<= 1; }
        DhtOut &D = dht[t];
        D.nvals = p;
        for (int k2 = 0; k2 < 17; k2++) D.bits[k2] = bits[k2];
        for (int k2 = 0; k2 < p; k2++) D.vals[k2] = T.vals[k2];
    }
}

// Sizes on the device: bits per scan from the bit-offset scan, then every scan's place in the group's bit buffer (word_base),
// its byte / 16-byte-group counts and the group index base -- what the host used to compute between two halves of the pipeline
// (one stream wait per megabatch less).  The buffers are sized from an ESTIMATE of the output (the input's size for a re-encode);
// if the real sizes do not fit, flags[0] is raised, every scan is given zero length so the back half does nothing, and the host
// re-runs the back half with exact sizes (flags[1..2] = words / groups needed).  One warp; scans in chunks of 32.
__global__ void k_ge_scanout(const Scan *__restrict__ scans, int nscans, const uint32_t *__restrict__ bitlen, const uint32_t *__restrict__ bitoff, uint32_t *__restrict__ total,
                             ScanOut *__restrict__ so, uint32_t words_cap, uint32_t groups_cap, uint32_t *__restrict__ flags)
{
    const int lane = threadIdx.x;
    uint32_t wbase = 0, gbase = 0;
    for (int c0 = 0; c0 < nscans; c0 += 32) {
        const int i = c0 + lane;
        uint32_t tb = 0;
        if (i < nscans) { const Scan s = scans[i]; const long long last = s.unit_base + s.nblocks - 1; tb = s.nblocks ? bitoff[last] + bitlen[last] - bitoff[s.unit_base] : 0; }
        const uint32_t nbytes = (tb + 7) / 8, ng = (nbytes + 15) / 16, nw = i < nscans ? (tb + 31) / 32 + 1 : 0;
        uint32_t wi = nw, gi = i < nscans ? ng : 0;                  // inclusive warp scans
        for (int d = 1; d < 32; d <<= 1) { const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, wi, d), b = __shfl_up_sync(0xFFFFFFFFu, gi, d); if (lane >= d) { wi += a; gi += b; } }
        if (i < nscans) { total[i] = tb; ScanOut o; o.total_bits = tb; o.nbytes = nbytes; o.ngroups = ng; o.group_base = gbase + gi - ng; o.word_base = wbase + wi - nw; o.pad_ = 0; so[i] = o; }
        wbase += __shfl_sync(0xFFFFFFFFu, wi, 31); gbase += __shfl_sync(0xFFFFFFFFu, gi, 31);
    }
    __syncwarp();
    const bool ovf = wbase > words_cap || gbase > groups_cap;
    if (lane == 0) { flags[0] = ovf ? 1u : 0u; flags[1] = wbase; flags[2] = gbase; flags[3] = 0; flags[4] = 0; }
    if (ovf) for (int i = lane; i < nscans; i += 32) { so[i].total_bits = 0; so[i].nbytes = 0; so[i].ngroups = 0; so[i].group_base = 0; so[i].word_base = 0; }
}

__global__ void k_ge_zero(const ScanOut *__restrict__ so, uint32_t *__restrict__ words)
{
    const ScanOut o = so[blockIdx.y];
    if (!o.total_bits) return;
    const uint32_t n = (o.total_bits + 31) / 32 + 1;
    uint32_t *w = words + o.word_base;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) w[i] = 0;
}

// ---- block-major passes: one thread per block; the block is read (and its threshold masks built) once per pass and
// serves every scan that visits it ------------------------------------------------------------------------------------
__device__ __forceinline__ int unit_of(const Scan &s, const BlockComp &bc, int row, int col)
{
    if (s.ns == 1) return (row < bc.rbh && col < bc.rbw) ? row * bc.rbw + col : -1;
    const int m = (row / bc.vs) * bc.mcux + col / bc.hs, q = bc.q_base + (row % bc.vs) * bc.hs + (col % bc.hs);
    return m * bc.blocks_per_mcu + q;
}
__device__ __forceinline__ BlockRef ref_of(const Scan &s, int u, const int16_t *blk)
{
    if (s.mode == MODE_SEQ || s.mode == MODE_DC_FIRST || s.ns > 1) return locate(s, u);   // needs the DC predecessor / the slot
    BlockRef r; r.blk = blk; r.prev = nullptr; r.slot = 0; return r;
}

// The classify pass builds each block's threshold masks once and leaves them in `masks` (24 bytes per block, indexed
// comp.mask_base + block); the histogram, length and emit passes read them back instead of re-deriving them from the 128-byte
// block (the mask construction was a quarter to a third of those passes' instructions).
__device__ __forceinline__ Masks3 load_masks(const Masks3 *__restrict__ masks, const BlockComp &bc, int i)
{
    const unsigned long long *q = reinterpret_cast<const unsigned long long *>(masks + bc.mask_base + i);   // 24-byte records
    Masks3 M;
    M.m[0] = __ldg(q); M.m[1] = __ldg(q + 1); M.m[2] = __ldg(q + 2);
    return M;
}

__global__ void k_geb_classify(const BlockComp *__restrict__ comps, const Scan *__restrict__ scans, uint32_t *__restrict__ meta, int *__restrict__ evkey, uint32_t *__restrict__ tail,
                               Masks3 *__restrict__ masks)
{
    const BlockComp bc = comps[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= bc.bw * bc.bh) return;
    const int row = i / bc.bw, col = i - row * bc.bw;
    const int16_t *blk = bc.coef + bc.comp_off + ((long long)row * bc.bw + col) * 64;
    const Masks3 M = make_masks3(blk);
    masks[bc.mask_base + i] = M;
    for (int j = 0; j < bc.nscan; j++) {
        const Scan &s = scans[bc.scan_idx[j]];
        const int u = unit_of(s, bc, row, col);
        if (u < 0) continue;
        const uint32_t m = classify_m(s, M);
        const long long g = s.unit_base + u;
        meta[g] = m; evkey[g] = meta_event(m) ? (int)g : -1; tail[g] = (uint32_t)meta_tail(m);
    }
}

// 512 blocks per CTA: the 16 KB shared histogram is zeroed and flushed once per CTA (a fifth of the pass at 128)
constexpr int HIST_THREADS = 512;
__global__ void __launch_bounds__(HIST_THREADS) k_geb_hist(const BlockComp *__restrict__ comps, const Scan *__restrict__ scans, const uint32_t *__restrict__ gcount, uint32_t *__restrict__ hist, const Masks3 *__restrict__ masks)
{
    __shared__ uint32_t h[4][1024];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) (&h[0][0])[i] = 0;
    __syncthreads();
    const BlockComp bc = comps[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < bc.bw * bc.bh) {
        const int row = i / bc.bw, col = i - row * bc.bw;
        const int16_t *blk = bc.coef + bc.comp_off + ((long long)row * bc.bw + col) * 64;
        const Masks3 M = load_masks(masks, bc, i);
        for (int j = 0; j < bc.nscan; j++) {
            const Scan &s = scans[bc.scan_idx[j]];
            const int u = unit_of(s, bc, row, col);
            if (u < 0) continue;
            uint32_t *hj = j < 4 ? h[j] : hist + (size_t)s.tab_base * 256;
            auto add = [&](int idx) { atomicAdd(&hj[idx], 1u); };
            HistSink<decltype(add)> sk(add);
            gen_block_m(s, ref_of(s, u, blk), M, gcount[s.unit_base + u], sk);
        }
    }
    __syncthreads();
    for (int j = 0; j < bc.nscan && j < 4; j++) {
        uint32_t *g = hist + (size_t)scans[bc.scan_idx[j]].tab_base * 256;
        for (int k = threadIdx.x; k < 1024; k += blockDim.x) if (h[j][k]) atomicAdd(&g[k], h[j][k]);
    }
}

__global__ void k_geb_len(const BlockComp *__restrict__ comps, const Scan *__restrict__ scans, const uint32_t *__restrict__ gcount, const Table *__restrict__ tabs, uint32_t *__restrict__ bitlen, const Masks3 *__restrict__ masks)
{
    const BlockComp bc = comps[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= bc.bw * bc.bh) return;
    const int row = i / bc.bw, col = i - row * bc.bw;
    const int16_t *blk = bc.coef + bc.comp_off + ((long long)row * bc.bw + col) * 64;
    const Masks3 M = load_masks(masks, bc, i);
    for (int j = 0; j < bc.nscan; j++) {
        const Scan &s = scans[bc.scan_idx[j]];
        const int u = unit_of(s, bc, row, col);
        if (u < 0) continue;
        LenSink sk; sk.tabs = tabs + s.tab_base;
        gen_block_m(s, ref_of(s, u, blk), M, gcount[s.unit_base + u], sk);
        bitlen[s.unit_base + u] = (uint32_t)sk.bits;
    }
}

__global__ void k_geb_emit(const BlockComp *__restrict__ comps, const Scan *__restrict__ scans, const uint32_t *__restrict__ gcount, const Table *__restrict__ tabs,
                           const uint32_t *__restrict__ bitoff, uint32_t *__restrict__ words, const Masks3 *__restrict__ masks, const ScanOut *__restrict__ so,
                           const uint32_t *__restrict__ flags)
{
    if (flags[0]) return;                   // the bit buffer is too small for this batch: the host re-runs this half with exact sizes
    const BlockComp bc = comps[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= bc.bw * bc.bh) return;
    const int row = i / bc.bw, col = i - row * bc.bw;
    const int16_t *blk = bc.coef + bc.comp_off + ((long long)row * bc.bw + col) * 64;
    const Masks3 M = load_masks(masks, bc, i);
    auto orw = [&](long long w, uint32_t v) { if (v) atomicOr(&words[w], v); };
    for (int j = 0; j < bc.nscan; j++) {
        const Scan &s = scans[bc.scan_idx[j]];
        const int u = unit_of(s, bc, row, col);
        if (u < 0) continue;
        EmitSink<decltype(orw)> sk(tabs + s.tab_base, orw, (long long)so[bc.scan_idx[j]].word_base, (unsigned long long)(bitoff[s.unit_base + u] - bitoff[s.unit_base]));
        gen_block_m(s, ref_of(s, u, blk), M, gcount[s.unit_base + u], sk);
        sk.finish();
    }
}

// byte i of a scan's unstuffed stream (big-endian within words), with flush_bits' padding ones in the last byte
__device__ __forceinline__ uint32_t scan_byte(const uint32_t *__restrict__ w, uint32_t i, uint32_t nbytes, uint32_t total_bits)
{
    uint32_t b = (w[i >> 2] >> (24 - 8 * (i & 3))) & 0xFF;
    if (i == nbytes - 1 && (total_bits & 7)) b |= (1u << (8 - (total_bits & 7))) - 1u;
    return b;
}

// 0xFF count per 16-byte group.  Grid (X, nscans + 1): row y < nscans strides over scan y's groups; the extra row zero-fills the
// group slots past the batch's last group, because the prefix sum that follows runs over the whole (capacity-sized) array.
__global__ void k_ge_ffcount(const ScanOut *__restrict__ so, int nscans, const uint32_t *__restrict__ words, uint32_t *__restrict__ ffcount, uint32_t groups_cap,
                             const uint32_t *__restrict__ flags)
{
    if ((int)blockIdx.y == nscans) {
        const uint32_t used = flags[0] ? 0u : flags[2];
        for (uint32_t g = used + blockIdx.x * blockDim.x + threadIdx.x; g < groups_cap; g += gridDim.x * blockDim.x) ffcount[g] = 0;
        return;
    }
    const ScanOut o = so[blockIdx.y];
    const uint32_t *w = words + o.word_base;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < o.ngroups; g += gridDim.x * blockDim.x) {
        uint32_t n = 0;
        for (uint32_t i = g * 16; i < g * 16 + 16 && i < o.nbytes; i++) n += scan_byte(w, i, o.nbytes, o.total_bits) == 0xFF;
        ffcount[o.group_base + g] = n;
    }
}

// per image: lay its scans out back to back in the image's output region; out_off / out_len per scan; an image that outgrows its
// region raises flags[0] (and flags[3] = the largest image size seen, so the host can size the retry)
__global__ void k_ge_layout(const ScanOut *__restrict__ so, int scans_per_image, int nimages, const uint32_t *__restrict__ ffcount, const uint32_t *__restrict__ ffoff,
                            uint32_t *__restrict__ out_off, uint32_t *__restrict__ out_len, uint32_t out_image_stride, uint32_t *__restrict__ flags)
{
    const int im = blockIdx.x * blockDim.x + threadIdx.x;
    if (im >= nimages) return;
    uint32_t off = 0;
    for (int k = 0; k < scans_per_image; k++) {
        const int si = im * scans_per_image + k;
        const ScanOut o = so[si];
        uint32_t ff = 0;
        if (o.ngroups) { const uint32_t lastg = o.group_base + o.ngroups - 1; ff = ffoff[lastg] + ffcount[lastg] - ffoff[o.group_base]; }
        out_off[si] = off; out_len[si] = o.nbytes + ff;
        off += o.nbytes + ff;
    }
    atomicMax(&flags[3], off);
    if (off > out_image_stride) atomicOr(&flags[4], 1u);
}

__global__ void k_ge_scatter(const ScanOut *__restrict__ so, const uint32_t *__restrict__ words, const uint32_t *__restrict__ ffoff,
                             const uint32_t *__restrict__ out_off, uint8_t *__restrict__ out, int scans_per_image, size_t out_image_stride, const uint32_t *__restrict__ flags)
{
    if (flags[4]) return;                   // some image does not fit its output region: nothing is written, the host retries
    const ScanOut o = so[blockIdx.y];
    const uint32_t *w = words + o.word_base;
    uint8_t *base = out + (size_t)(blockIdx.y / scans_per_image) * out_image_stride + out_off[blockIdx.y];
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < o.ngroups; g += gridDim.x * blockDim.x) {
        uint8_t *dst = base + (size_t)g * 16 + (ffoff[o.group_base + g] - ffoff[o.group_base]);
        for (uint32_t i = g * 16; i < g * 16 + 16 && i < o.nbytes; i++) {
            const uint32_t b = scan_byte(w, i, o.nbytes, o.total_bits);
            *dst++ = (uint8_t)b;
            if (b == 0xFF) *dst++ = 0;
        }
    }
}

// dummy blocks of partial MCUs (jccoefct.c / jctrans.c rule, jpeg_fill_dummy_blocks on the host): AC = 0, DC copied
__global__ void k_ge_fill_dummy(int16_t *__restrict__ coef, long long comp_off, int bw, int bh, int rbw, int rbh, int hs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= bw * bh) return;
    const int row = i / bw, col = i - row * bw;
    if (row < rbh && col < rbw) return;
    int r = row, c = col;
    while (r >= rbh || c >= rbw) { if (r >= rbh) { c = (c / hs) * hs + hs - 1; r--; } else c = rbw - 1; }
    int16_t *base = coef + comp_off;
    int16_t *dst = base + ((long long)row * bw + col) * 64;
    const int16_t dc = base[((long long)r * bw + c) * 64];
    for (int k = 1; k < 64; k++) dst[k] = 0;
    dst[0] = dc;
}

// ---- host orchestration ---------------------------------------------------------------------------------------------
// Three steps so that a megabatch costs the host one wait that overlaps device work plus the final one, and so that a caller with
// everything resident in HBM (bench.py's device-only figure) can enqueue the whole pass sequence without any wait:
//   prepare()  plan + buffers (sized from an estimate of the output) + H2D of the descriptors
//   enqueue()  every kernel; D2H of the sizes right after the bit-offset scan (event), D2H of DHT payloads / stuffed lengths at the end
//   finish()   wait for the sizes (the emit / stuffing kernels are still running), size and enqueue the D2H of the stuffed scans,
//              final wait; a batch that outgrew the estimate re-runs the back half with exact sizes
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

static thread_local unsigned long long *tl_enc_generation = nullptr;   // bumped on every reallocation (captured graphs hold the old pointers)
template <typename T> static bool grow(T *&p, size_t &cap, size_t need, bool host, std::string &err)
{
    if (need <= cap) return true;
    if (tl_enc_generation) ++*tl_enc_generation;
    if (p) { if (host) cudaFreeHost(p); else cudaFree(p); }
    p = nullptr; cap = 0;
    // several sizes depend on image CONTENT (bytes of entropy-coded output); round up to a power of two with headroom so
    // a slot stops reallocating after its first image of a given class (cudaFree / cudaHostAlloc stall every stream)
    size_t want = 1 << 16; while (want < need + need / 2) want <<= 1;
    void *q = nullptr;
    cudaError_t e = host ? cudaHostAlloc(&q, want, cudaHostAllocDefault) : cudaMalloc(&q, want);
    if (e != cudaSuccess) { err = std::string(host ? "cudaHostAlloc: " : "cudaMalloc: ") + cudaGetErrorString(e); return false; }
    p = (T *)q; cap = want; return true;
}

GpuEncoder::~GpuEncoder()
{
    cudaFree(d_scans); cudaFree(d_comps); cudaFree(d_meta); cudaFree(d_evkey); cudaFree(d_prev); cudaFree(d_tail); cudaFree(d_tsum); cudaFree(d_gcount);
    cudaFree(d_bitlen); cudaFree(d_bitoff); cudaFree(d_hist); cudaFree(d_tabs); cudaFree(d_dht); cudaFree(d_total); cudaFree(d_so);
    cudaFree(d_words); cudaFree(d_masks); cudaFree(d_ffcount); cudaFree(d_ffoff); cudaFree(d_outoff); cudaFree(d_outlen); cudaFree(d_out); cudaFree(d_temp);
    cudaFree(d_flags);
    cudaFreeHost(h_small); cudaFreeHost(h_out);
    if (ev_sizes) cudaEventDestroy((cudaEvent_t)ev_sizes);
}

bool GpuEncoder::size_back_buffers(size_t image_bytes, std::string &err)
{   // everything whose size follows the OUTPUT: bit buffer, 16-byte group arrays, stuffed bytes
    const int NS = (int)plan.scans.size();
    // capacities only grow (per image count): they are kernel arguments of the launch sequence, and a sequence whose arguments do
    // not change from megabatch to megabatch can be replayed as a CUDA graph
    if (nimg != cap_nimg) { cap_nimg = nimg; est_image_bytes = 0; }
    if (image_bytes <= est_image_bytes && words_cap) return true;
    est_image_bytes = image_bytes + image_bytes / 8;
    image_bytes = est_image_bytes;
    words_cap = (uint32_t)std::min<size_t>((size_t)nimg * (image_bytes / 4 + 1) + 2 * (size_t)NS + 64, 0xFFFFFF00u);
    groups_cap = (uint32_t)((size_t)words_cap / 4 + NS + 1);
    out_stride = align_up(image_bytes + image_bytes / 8 + 1024, 256);
    tl_enc_generation = &generation;
    size_t c;
    c = cap_words; if (!grow(d_words, c, (size_t)words_cap * 4 + 64, false, err)) return false; cap_words = c;
    c = cap_ff[0]; if (!grow(d_ffcount, c, (size_t)groups_cap * 4 + 4, false, err)) return false; cap_ff[0] = c;
    c = cap_ff[1]; if (!grow(d_ffoff, c, (size_t)groups_cap * 4 + 4, false, err)) return false; cap_ff[1] = c;
    c = cap_out; if (!grow(d_out, c, out_stride * nimg, false, err)) return false; cap_out = c;
    size_t tb3 = 0; cub::DeviceScan::ExclusiveSum((void *)nullptr, tb3, d_ffcount, d_ffoff, (int)groups_cap, (cudaStream_t)0);
    c = cap_temp; if (tb3 + 256 > c) { if (!grow(d_temp, c, tb3 + 256, false, err)) return false; cap_temp = c; }
    tl_enc_generation = nullptr;
    return true;
}

bool GpuEncoder::prepare(const JpegGeom &g, bool progressive, int16_t *const *d_coefs, int nimages, void *stream_, size_t out_bytes_hint, std::string &err)
{
    cudaStream_t st = (cudaStream_t)stream_;
    geom = g; prog = progressive;
    std::vector<const int16_t *> bases(d_coefs, d_coefs + nimages);
    coef_bases.assign(d_coefs, d_coefs + nimages);
    gpuenc_plan(g, progressive, bases.data(), nimages, plan);
    nimg = nimages;
    const int NS = (int)plan.scans.size();
    const long long U = plan.total_units;
    if (U >= (1ll << 31)) { err = "batch too large for the entropy encoder"; return false; }
    overflow = false;
    tl_enc_generation = &generation;
    for (auto &sc_ : plan.scans) if (!masks_cover(sc_.mode, sc_.Al)) { err = "scan script outside the device encoder's mask range"; overflow = true; return false; }
    if (!ev_sizes) { cudaEvent_t e; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming | (stream_wait_mode() == 0 ? 0 : cudaEventBlockingSync))); ev_sizes = e; }
    // ---- buffers whose size follows the INPUT
    size_t c;
    c = cap_scans; if (!grow(d_scans, c, NS * sizeof(Scan), false, err)) return false; cap_scans = c;
    const int NC = (int)plan.comps.size();
    c = cap_comps; if (!grow(d_comps, c, NC * sizeof(BlockComp), false, err)) return false; cap_comps = c;
    c = cap_u[0]; if (!grow(d_meta, c, U * 4, false, err)) return false; cap_u[0] = c;
    c = cap_u[1]; if (!grow(d_evkey, c, U * 4, false, err)) return false; cap_u[1] = c;
    c = cap_u[2]; if (!grow(d_prev, c, U * 4, false, err)) return false; cap_u[2] = c;
    c = cap_u[3]; if (!grow(d_tail, c, U * 4, false, err)) return false; cap_u[3] = c;
    c = cap_u[4]; if (!grow(d_tsum, c, U * 4, false, err)) return false; cap_u[4] = c;
    c = cap_u[5]; if (!grow(d_gcount, c, U * 4, false, err)) return false; cap_u[5] = c;
    c = cap_u[6]; if (!grow(d_bitlen, c, U * 4, false, err)) return false; cap_u[6] = c;
    c = cap_u[7]; if (!grow(d_bitoff, c, U * 4, false, err)) return false; cap_u[7] = c;
    c = cap_hist; if (!grow(d_hist, c, (size_t)NS * 4 * 256 * 4, false, err)) return false; cap_hist = c;
    c = cap_tabs; if (!grow(d_tabs, c, (size_t)NS * 4 * sizeof(Table), false, err)) return false; cap_tabs = c;
    c = cap_dht; if (!grow(d_dht, c, (size_t)NS * 4 * sizeof(DhtOut), false, err)) return false; cap_dht = c;
    c = cap_total; if (!grow(d_total, c, (size_t)NS * 4, false, err)) return false; cap_total = c;
    c = cap_so; if (!grow(d_so, c, (size_t)NS * sizeof(ScanOut), false, err)) return false; cap_so = c;
    c = cap_oo; if (!grow(d_outoff, c, (size_t)NS * 4, false, err)) return false; cap_oo = c;
    c = cap_ol; if (!grow(d_outlen, c, (size_t)NS * 4, false, err)) return false; cap_ol = c;
    c = cap_flags; if (!grow(d_flags, c, 64, false, err)) return false; cap_flags = c;
    c = cap_masks; if (!grow(d_masks, c, (size_t)plan.total_comp_blocks * sizeof(Masks3), false, err)) return false; cap_masks = c;
    o_scans = 0; o_total = o_scans + align_up((size_t)NS * sizeof(Scan), 256); o_outlen = o_total + align_up((size_t)NS * 4, 256);
    o_dht = o_outlen + align_up((size_t)NS * 4, 256); o_comps = o_dht + align_up((size_t)NS * 4 * sizeof(DhtOut), 256);
    o_flags = o_comps + align_up((size_t)NC * sizeof(BlockComp), 256);
    const size_t small_bytes = o_flags + 256;
    c = cap_small; if (!grow(h_small, c, small_bytes, true, err)) return false; cap_small = c;
    size_t tb1 = 0, tb2 = 0;
    cub::DeviceScan::ExclusiveScan((void *)nullptr, tb1, d_evkey, d_prev, cub::Max(), -1, (int)U, st);
    cub::DeviceScan::ExclusiveSum((void *)nullptr, tb2, d_tail, d_tsum, (int)U, st);
    c = cap_temp; if (!grow(d_temp, c, std::max(tb1, tb2) + 256, false, err)) return false; cap_temp = c;
    // ---- buffers whose size follows the OUTPUT: estimate now, exact on a retry.  A re-encode at lower quality does not grow, so
    // the caller's hint is the input's entropy-coded size; without a hint a third of the coefficient bytes (~ 1 byte / pixel).
    const size_t coef_bytes = (size_t)g.total_coefs * 2;
    if (coef_bytes != learned_for) { learned_for = coef_bytes; learned_image_bytes = 0; }
    size_t est = out_bytes_hint ? out_bytes_hint / nimages + out_bytes_hint / nimages / 4 : coef_bytes / 3;
    est = std::max(est, learned_image_bytes + learned_image_bytes / 8) + 8192;
    est = std::min(est, coef_bytes * 2 + (size_t)plan.scans_per_image * 64 + 8192);      // worst case: 128 B per block and scan... bounded by the retry anyway
    tl_enc_generation = nullptr;
    if (!size_back_buffers(est, err)) return false;
    memcpy(h_small + o_scans, plan.scans.data(), NS * sizeof(Scan));
    memcpy(h_small + o_comps, plan.comps.data(), NC * sizeof(BlockComp));
    (void)st;
    return true;
}

bool GpuEncoder::upload(void *stream_, std::string &err)
{   // H2D of the scan / component descriptors prepare() wrote into pinned memory
    cudaStream_t st = (cudaStream_t)stream_;
    const int NS = (int)plan.scans.size(), NC = (int)plan.comps.size();
    CU(cudaMemcpyAsync(d_scans, h_small + o_scans, NS * sizeof(Scan), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_comps, h_small + o_comps, NC * sizeof(BlockComp), cudaMemcpyHostToDevice, st));
    return true;
}

unsigned long long GpuEncoder::signature() const
{   // everything upload() + enqueue_front() + enqueue_sizes() + enqueue_back() bake into their driver calls
    unsigned long long h = 1469598103934665603ull;
    auto mix = [&](unsigned long long v) { h = (h ^ v) * 1099511628211ull; };
    mix((unsigned long long)nimg); mix((unsigned long long)plan.scans.size()); mix((unsigned long long)plan.comps.size()); mix((unsigned long long)plan.total_units);
    mix((unsigned long long)plan.max_comp_blocks); mix((unsigned long long)plan.scans_per_image); mix(words_cap); mix(groups_cap); mix(out_stride); mix(generation);
    mix((unsigned long long)geom.width); mix((unsigned long long)geom.height); mix(prog ? 1 : 0);
    for (int c = 0; c < geom.ncomp; c++) { mix((unsigned long long)geom.bw[c]); mix((unsigned long long)geom.bh[c]); mix((unsigned long long)geom.rbw[c]); mix((unsigned long long)geom.rbh[c]); mix((unsigned long long)geom.hs[c]); }
    for (auto pb : coef_bases) mix((unsigned long long)(uintptr_t)pb);
    int max_units = 0; for (auto &sc : plan.scans) max_units = std::max(max_units, sc.nblocks);
    mix((unsigned long long)max_units);
    return h;
}

// scan sizes / buffer layout on the device, and their way back to the host
bool GpuEncoder::enqueue_sizes(void *stream_, std::string &err)
{
    cudaStream_t st = (cudaStream_t)stream_;
    const int NS = (int)plan.scans.size();
    uint32_t *h_total = reinterpret_cast<uint32_t *>(h_small + o_total), *h_flags = reinterpret_cast<uint32_t *>(h_small + o_flags);
    k_ge_scanout<<<1, 32, 0, st>>>(d_scans, NS, d_bitlen, d_bitoff, d_total, d_so, words_cap, groups_cap, d_flags);
    LT_MARK("k_ge_scanout");
    CU(cudaMemcpyAsync(h_total, d_total, (size_t)NS * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_flags, d_flags, 12, cudaMemcpyDeviceToHost, st));
    LT_MARK("copy");
    return true;
}
bool GpuEncoder::mark_sizes(void *stream_, std::string &err)
{   // the event finish() waits for before it sizes the D2H of the scans (kept out of captured graphs: a plain stream operation)
    CU(cudaEventRecord((cudaEvent_t)ev_sizes, (cudaStream_t)stream_));
    return true;
}

bool GpuEncoder::enqueue_back(void *stream_, std::string &err)
{
    cudaStream_t st = (cudaStream_t)stream_;
    const int NS = (int)plan.scans.size(), NC = (int)plan.comps.size();
    uint32_t *h_flags = reinterpret_cast<uint32_t *>(h_small + o_flags);
    const dim3 gb(cdiv(plan.max_comp_blocks, 128), NC);
    k_ge_zero<<<dim3(64, NS), 256, 0, st>>>(d_so, d_words);
    LT_MARK("k_ge_zero");
    k_geb_emit<<<gb, 128, 0, st>>>(d_comps, d_scans, d_gcount, d_tabs, d_bitoff, d_words, d_masks, d_so, d_flags);
    LT_MARK("k_geb_emit");
    k_ge_ffcount<<<dim3(32, NS + 1), 128, 0, st>>>(d_so, NS, d_words, d_ffcount, groups_cap, d_flags);
    LT_MARK("k_ge_ffcount");
    size_t tb = cap_temp;
    cub::DeviceScan::ExclusiveSum(d_temp, tb, d_ffcount, d_ffoff, (int)groups_cap, st);
    LT_MARK("cub_scan");
    k_ge_layout<<<cdiv(nimg, 64), 64, 0, st>>>(d_so, plan.scans_per_image, nimg, d_ffcount, d_ffoff, d_outoff, d_outlen, (uint32_t)out_stride, d_flags);
    LT_MARK("k_ge_layout");
    k_ge_scatter<<<dim3(32, NS), 128, 0, st>>>(d_so, d_words, d_ffoff, d_outoff, d_out, plan.scans_per_image, out_stride, d_flags);
    LT_MARK("k_ge_scatter");
    CU(cudaMemcpyAsync(h_small + o_outlen, d_outlen, (size_t)NS * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_small + o_dht, d_dht, (size_t)NS * 4 * sizeof(DhtOut), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_flags + 4, d_flags, 32, cudaMemcpyDeviceToHost, st));         // second snapshot: image-level overflow (flags[3..4])
    CU(cudaGetLastError());
    launches += 7;
    return true;
}

bool GpuEncoder::enqueue(void *stream_, bool fill_dummy, std::string &err)
{
    return enqueue_front(stream_, fill_dummy, err) && enqueue_sizes(stream_, err) && mark_sizes(stream_, err) && enqueue_back(stream_, err);
}

bool GpuEncoder::enqueue_front(void *stream_, bool fill_dummy, std::string &err)
{
    cudaStream_t st = (cudaStream_t)stream_;
    const JpegGeom &g = geom;
    const int NS = (int)plan.scans.size(), NC = (int)plan.comps.size();
    const long long U = plan.total_units;
    int max_units = 0; for (auto &s : plan.scans) max_units = std::max(max_units, s.nblocks);
    launches = 0;
    const dim3 gb(cdiv(plan.max_comp_blocks, 128), NC);
    if (fill_dummy) {
        for (int im = 0; im < nimg; im++) for (int cc = 0; cc < g.ncomp; cc++) {
            if (g.rbw[cc] == g.bw[cc] && g.rbh[cc] == g.bh[cc]) continue;
            k_ge_fill_dummy<<<cdiv((long long)g.bw[cc] * g.bh[cc], 256), 256, 0, st>>>(coef_bases[im], g.comp_offset[cc], g.bw[cc], g.bh[cc], g.rbw[cc], g.rbh[cc], g.hs[cc]);
            LT_MARK("k_ge_fill_dummy");
            launches++;
        }
    }
    const dim3 gu1(cdiv(max_units + 1, 128), NS);
    k_geb_classify<<<gb, 128, 0, st>>>(d_comps, d_scans, d_meta, d_evkey, d_tail, d_masks);
    LT_MARK("k_geb_classify");
    size_t tb = cap_temp;
    cub::DeviceScan::ExclusiveScan(d_temp, tb, d_evkey, d_prev, cub::Max(), -1, (int)U, st);
    LT_MARK("cub_scan");
    tb = cap_temp;
    cub::DeviceScan::ExclusiveSum(d_temp, tb, d_tail, d_tsum, (int)U, st);
    LT_MARK("cub_scan");
    CU(cudaMemsetAsync(d_gcount, 0, U * 4, st));
    LT_MARK("memset");
    k_ge_groups<<<gu1, 128, 0, st>>>(d_scans, d_meta, d_evkey, d_prev, d_tsum, d_gcount);
    LT_MARK("k_ge_groups");
    CU(cudaMemsetAsync(d_hist, 0, (size_t)NS * 4 * 256 * 4, st));
    LT_MARK("memset");
    k_geb_hist<<<dim3(cdiv(plan.max_comp_blocks, HIST_THREADS), NC), HIST_THREADS, 0, st>>>(d_comps, d_scans, d_gcount, d_hist, d_masks);
    LT_MARK("k_geb_hist");
    k_ge_tables<<<NS * 4, 32, 0, st>>>(d_hist, d_tabs, d_dht);
    LT_MARK("k_ge_tables");
    k_geb_len<<<gb, 128, 0, st>>>(d_comps, d_scans, d_gcount, d_tabs, d_bitlen, d_masks);
    LT_MARK("k_geb_len");
    tb = cap_temp;
    cub::DeviceScan::ExclusiveSum(d_temp, tb, d_bitlen, d_bitoff, (int)U, st);
    LT_MARK("cub_scan");
    launches += 8;
    CU(cudaGetLastError());
    return true;
}

bool GpuEncoder::finish(void *stream_, bool fetch, std::string &err)
{
    cudaStream_t st = (cudaStream_t)stream_;
    const int NS = (int)plan.scans.size(), spi = plan.scans_per_image;
    uint32_t *h_total = reinterpret_cast<uint32_t *>(h_small + o_total), *h_flags = reinterpret_cast<uint32_t *>(h_small + o_flags);
    uint32_t *h_outlen = reinterpret_cast<uint32_t *>(h_small + o_outlen);
    const DhtOut *h_dht = reinterpret_cast<const DhtOut *>(h_small + o_dht);
    for (int attempt = 0;; attempt++) {
        // sizes are on the host while the emit / stuffing kernels still run
        const auto tw0 = std::chrono::steady_clock::now();
        if (fetch) { CU(event_wait((cudaEvent_t)ev_sizes)); } else { CU(stream_wait(st)); }
        wait_sizes_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tw0).count();
        size_t img_max = 0, img = 0;
        for (int si = 0; si < NS; si++) { img += (h_total[si] + 7) / 8; if ((si + 1) % spi == 0) { img_max = std::max(img_max, img); img = 0; } }
        if (h_flags[0]) {                                       // the estimate was too small for the bit buffer: exact sizes, back half again
            if (attempt >= 3) { err = "entropy encoder could not size its buffers"; return false; }
            CU(stream_wait(st));
            if (!size_back_buffers(img_max + img_max / 16 + 4096, err)) return false;
            retries++;
            if (!enqueue_sizes(st, err) || !mark_sizes(st, err) || !enqueue_back(st, err)) return false;
            continue;
        }
        if (fetch) {
            copy_bytes = std::min(out_stride, align_up(img_max + img_max / 8 + 256, 256));
            size_t c = cap_hout; if (!grow(h_out, c, copy_bytes * nimg, true, err)) return false; cap_hout = c;
            for (int im = 0; im < nimg; im++) CU(cudaMemcpyAsync(h_out + (size_t)im * copy_bytes, d_out + (size_t)im * out_stride, copy_bytes, cudaMemcpyDeviceToHost, st));
            CU(cudaGetLastError());
            const auto tw1 = std::chrono::steady_clock::now();
            CU(stream_wait(st));
            wait_final_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tw1).count();
        }
        if (h_flags[4 + 4]) {                                   // an image stuffed past its output region (more 0xFF bytes than one in eight)
            if (attempt >= 3) { err = "entropy encoder could not size its output"; return false; }
            if (!size_back_buffers((size_t)h_flags[4 + 3] + 4096, err)) return false;
            retries++;
            if (!enqueue_sizes(st, err) || !mark_sizes(st, err) || !enqueue_back(st, err)) return false;
            continue;
        }
        learned_image_bytes = std::max<size_t>(learned_image_bytes, h_flags[4 + 3]);
        if (fetch) {   // rare: an image stuffed more than the copied margin -> fetch everything at full stride
            bool shortc = false;
            for (int im = 0; im < nimg && !shortc; im++) { size_t tot = 0; for (int k = 0; k < spi; k++) tot += h_outlen[im * spi + k]; shortc = tot > copy_bytes; }
            if (shortc) {
                size_t c = cap_hout; if (!grow(h_out, c, out_stride * nimg, true, err)) return false; cap_hout = c;
                copy_bytes = out_stride;
                for (int j = 0; j < nimg; j++) CU(cudaMemcpyAsync(h_out + (size_t)j * copy_bytes, d_out + (size_t)j * out_stride, copy_bytes, cudaMemcpyDeviceToHost, st));
                CU(stream_wait(st));
            }
        }
        break;
    }
    // ---- describe the result
    results.assign((size_t)NS, EncodedScan());
    for (int si = 0; si < NS; si++) {
        EncodedScan &e = results[si];
        const int im = si / spi, k = si % spi;
        e.def = plan.defs[k];
        size_t off = 0; for (int j = 0; j < k; j++) off += h_outlen[im * spi + j];
        e.data = fetch ? h_out + (size_t)im * copy_bytes + off : nullptr; e.len = h_outlen[si];
        bool need[2][2]; jpeg_scan_tables_needed(geom, prog, e.def, need);
        for (int kind = 0; kind < 2; kind++) for (int t = 0; t < 2; t++) {
            e.has_tab[kind][t] = need[kind][t];
            const DhtOut &D = h_dht[(size_t)si * 4 + kind * 2 + t];
            memcpy(e.bits[kind][t], D.bits, 17); memcpy(e.vals[kind][t], D.vals, 256); e.nvals[kind][t] = D.nvals;
        }
    }
    return true;
}

bool GpuEncoder::encode(const JpegGeom &g, bool progressive, int16_t *const *d_coefs, int nimages, void *stream_, bool fill_dummy, std::string &err, size_t out_bytes_hint)
{
    return prepare(g, progressive, d_coefs, nimages, stream_, out_bytes_hint, err) && upload(stream_, err) && enqueue(stream_, fill_dummy, err) && finish(stream_, true, err);
}

} // namespace b200

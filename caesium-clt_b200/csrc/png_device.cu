// png_device.cu -- device orchestration of the lossless PNG path (libcaesium png::lossless -> oxipng::optimize_from_memory,
// /root/reference/src/compressor.rs:428,436-437): upload the decoded samples, apply the cheap lossless reductions
// (opaque alpha, grey RGB), then for every row-filter strategy of the optimisation preset run K6 (filter) + K7 (match,
// parse) and estimate the DEFLATE size from the token histogram; the winning strategy's tokens come back to the host,
// which Huffman-codes and frames them (png_host.cpp).
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <cstring>
#include "png_device.h"
#include "png_kernels.h"
#include "png_deflate.h"
#include <chrono>
#include <cstdlib>
#include "stream_wait.h"
#include "launch_timer.h"

namespace b200 {

#define CUP(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { err = std::string(#expr) + ": " + cudaGetErrorString(e_); return false; } } while (0)

static const int kChunk = 4096;
static const int kBlockTokens = 1 << 16;            // tokens per DEFLATE block (deflate_tokens' default)

template <typename T> static bool growp(T *&p, size_t &cap, size_t need, bool host, std::string &err)
{
    if (need <= cap) return true;
    if (p) { if (host) cudaFreeHost(p); else cudaFree(p); }
    p = nullptr; cap = 0;
    size_t want = 1 << 16; while (want < need + need / 4) want <<= 1;
    void *q = nullptr;
    cudaError_t e = host ? cudaHostAlloc(&q, want, cudaHostAllocDefault) : cudaMalloc(&q, want);
    if (e != cudaSuccess) { err = std::string(host ? "cudaHostAlloc: " : "cudaMalloc: ") + cudaGetErrorString(e); return false; }
    p = (T *)q; cap = want; return true;
}

PngDevice::~PngDevice()
{
    cudaFree(d_raw); cudaFree(d_raw2); cudaFree(d_filt); cudaFree(d_best); cudaFree(d_tok); cudaFree(d_out); cudaFree(d_counts); cudaFree(d_offsets);
    cudaFree(d_hist); cudaFree(d_sums); cudaFree(d_tlog); cudaFree(d_temp); cudaFreeHost(h_small); cudaFreeHost(h_tok); cudaFreeHost(h_raw);
    cudaFree(d_filt_all); cudaFree(d_fin); cudaFree(d_sums_in); cudaFree(d_sync); cudaFree(d_dfl); cudaFree(d_z); cudaFreeHost(h_z);
}

// oxipng presets: which row-filter strategies each optimisation level tries
std::vector<int> png_level_strategies(int level)
{
    switch (level) {
        case 0: return {PNGF_NONE};
        case 1: return {PNGF_NONE, PNGF_BIGRAMS};
        case 2: return {PNGF_NONE, PNGF_SUB, PNGF_ENTROPY, PNGF_BIGRAMS};
        case 3: case 4: return {PNGF_NONE, PNGF_BIGRAMS, PNGF_BIGENT, PNGF_BRUTE};
        case 5: return {PNGF_NONE, PNGF_BIGRAMS, PNGF_BIGENT, PNGF_BRUTE, PNGF_UP, PNGF_MINSUM};
        default: return {PNGF_NONE, PNGF_BIGRAMS, PNGF_BIGENT, PNGF_BRUTE, PNGF_UP, PNGF_MINSUM, PNGF_AVERAGE, PNGF_PAETH};
    }
}

// estimated DEFLATE payload bits of a token histogram under its own optimal (unlimited) code: sum f * (log2(total/f)) + extra
static double estimate_bits(const uint32_t *h)
{
    static const uint8_t lx[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint8_t dx[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    double tl = 0, td = 0, bits = 0;
    for (int i = 0; i < 286; i++) tl += h[i];
    for (int i = 0; i < 30; i++) td += h[286 + i];
    for (int i = 0; i < 286; i++) if (h[i]) bits += h[i] * (std::log2(tl / h[i]) + (i >= 257 ? lx[i - 257] : 0));
    for (int i = 0; i < 30; i++) if (h[286 + i]) bits += h[286 + i] * (std::log2(td / h[286 + i]) + dx[i]);
    return bits;
}

// One strategy: K6 into `filt` (skipped when the stream is already there), K7 over it.  Trials run the fixed-distance candidates only
// (enough to rank the strategies by their token statistics); the winner's stream then gets the hash candidates as well.
bool PngDevice::run_strategy(int strategy, int h, int rb, int bpp, void *stream_, std::string &err, uint8_t *filt, bool do_filter, bool with_hash)
{
    cudaStream_t st = (cudaStream_t)stream_;
    const size_t n = (size_t)h * (rb + 1);
    if (!filt) filt = d_filt;
    int rc = do_filter ? launch_png_filter(d_raw, filt, h, rb, bpp, strategy, d_tlog, st) : 0;
    if (!rc) rc = launch_png_match(filt, d_best, n, bpp, rb + 1, st);
    if (!rc && with_hash) rc = launch_png_hashmatch(filt, d_best, n, d_hist + 2048, st);
    if (rc) { err = std::string("png kernels: ") + cudaGetErrorString((cudaError_t)rc); return false; }
    CUP(cudaMemsetAsync(d_hist, 0, 316 * 4, st));
    rc = launch_png_parse(d_best, filt, n, kChunk, d_tok, d_counts, d_hist, st);
    if (rc) { err = std::string("png parse: ") + cudaGetErrorString((cudaError_t)rc); return false; }
    return true;
}

// number of tokens of the compacted stream = last offset + last count (both still on the device)
__global__ void k_png_ntok(const uint32_t *__restrict__ counts, const uint32_t *__restrict__ offsets, size_t nchunks, uint32_t *__restrict__ ntok)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) *ntok = offsets[nchunks - 1] + counts[nchunks - 1];
}

bool PngDevice::ensure_buffers(size_t nraw, size_t nmax, size_t rb, void *stream_, std::string &err)
{
    cudaStream_t st = (cudaStream_t)stream_;
    const size_t nchunks_max = (nmax + kChunk - 1) / kChunk;
    z_cap = nmax + nmax / 32 + ((nmax >> 16) + 2) * 512 + 4096;                 // the zlib payload: Huffman-coded literals cannot exceed ~8.1 bits each
    z_cap = (z_cap + 255) / 256 * 256;
    if (!growp(d_raw, cap_raw, nraw + 64, false, err) || !growp(d_raw2, cap_raw2, nraw + 64, false, err) ||
        !growp(d_filt, cap_filt, nmax + 64, false, err) || !growp(d_best, cap_best, nmax * 4, false, err) || !growp(d_tok, cap_tok, nmax * 4, false, err) ||
        !growp(d_out, cap_out, nmax * 4, false, err) || !growp(d_counts, cap_counts, nchunks_max * 4 + 4, false, err) || !growp(d_offsets, cap_offsets, nchunks_max * 4 + 4, false, err) ||
        !growp(d_hist, cap_hist, 316 * 4 * 16, false, err) || !growp(d_sums, cap_sums, ((nmax + 4095) / 4096) * 16 + 16, false, err) ||
        !growp(d_sums_in, cap_sums_in, ((nmax + 4095) / 4096) * 16 + 16, false, err) ||
        !growp(d_tlog, cap_tlog, (rb + 8) * 4, false, err) || !growp(h_small, cap_small, 1 << 16, true, err) ||
        !growp(d_sync, cap_sync, ((nraw / std::max<size_t>(rb, 1) + 31) / 32 + 16) * 4 + 2048 * 4 + 64, false, err) ||
        !growp(d_dfl, cap_dfl, png_deflate_scratch_bytes(nmax, kBlockTokens), false, err) || !growp(d_z, cap_z, z_cap + 64, false, err) ||
        !growp(h_z, cap_hz, z_cap + ((nmax + 4095) / 4096) * 32 + 256, true, err)) return false;
    size_t tb = 0; cub::DeviceScan::ExclusiveSum((void *)nullptr, tb, d_counts, d_offsets, (int)nchunks_max, st);
    if (!growp(d_temp, cap_temp, tb + 256, false, err)) return false;
    if (tlog_n < rb + 2) { std::vector<uint32_t> t(rb + 2); png_make_tlog(t.data(), rb + 1); CUP(cudaMemcpyAsync(d_tlog, t.data(), (rb + 2) * 4, cudaMemcpyHostToDevice, st)); CUP(stream_wait(st)); tlog_n = rb + 2; }
    return true;
}

static uint32_t combine_adler(const unsigned long long *sums, size_t n)
{   // Adler-32 of n bytes from per-4096-byte pieces (S = sum of bytes, T = sum of (len - k) * byte_k): a' = a + S, b' = b + len * a + T  (mod 65521)
    unsigned long long a = 1, b = 0;
    const size_t npieces = (n + 4095) / 4096;
    for (size_t p = 0; p < npieces; p++) {
        const unsigned long long len = std::min<size_t>(4096, n - p * 4096);
        b = (b + len * a + sums[2 * p + 1]) % 65521; a = (a + sums[2 * p]) % 65521;
    }
    return (uint32_t)((b << 16) | a);
}

uint8_t *PngDevice::input_buffer(size_t bytes, size_t &cap, std::string &err)
{
    if (!growp(h_raw, cap_hraw, bytes + 4096 + 64, true, err)) return nullptr;
    cap = cap_hraw;
    return h_raw;
}

// The lossless path proper: the inflated IDAT (filter byte + filtered bytes per row) sits in this object's pinned buffer
// (input_buffer()); everything from there to the finished zlib stream runs on the device -- un-filter (wavefront), Adler-32 check of
// the input, reductions, K6 / K7 per strategy, DEFLATE coding -- except a palette reduction, which (at most 256 colours) goes
// through the host and the raw-sample entry point below.
bool PngDevice::compress_filtered(PngInfo &info, size_t nfilt, uint32_t stored_adler, int level, void *stream_, std::vector<uint8_t> &zlib_stream, int *chosen, std::string &err)
{
    cudaStream_t st = (cudaStream_t)stream_;
    corrupt = false;
    const int h = (int)info.height; const size_t rb = info.row_bytes; const int bpp = info.bpp;
    const size_t nraw = (size_t)h * rb, nin = (size_t)h * (rb + 1);
    if (nfilt < nin) { err = "IDAT too short"; corrupt = true; return false; }
    const size_t nmax = nin + 64;
    if (!ensure_buffers(nraw, nmax, rb, st, err) || !growp(d_fin, cap_fin, nmax + 64, false, err)) return false;
    CUP(cudaMemcpyAsync(d_fin, h_raw, nin, cudaMemcpyHostToDevice, st)); LT_MARK("h2d");
    int rc = launch_png_adler(d_fin, nin, d_sums_in, st);
    uint32_t *d_un = d_sync, *d_flags = d_hist;
    uint32_t *d_set = reinterpret_cast<uint32_t *>(((uintptr_t)(d_sync + ((size_t)(h + 31) / 32 + 8)) + 7) & ~(uintptr_t)7);
    if (!rc) rc = launch_png_unfilter(d_fin, d_raw, h, (int)rb, bpp, d_un, st);
    if (rc) { err = std::string("png unfilter: ") + cudaGetErrorString((cudaError_t)rc); return false; }
    CUP(cudaMemsetAsync(d_flags, 0, 16, st));
    const bool eight = info.bit_depth == 8 && info.trns.empty();
    const bool probe_ag = eight && (info.color_type == 2 || info.color_type == 4 || info.color_type == 6);
    const bool probe_pal = png_palette_candidate(info);
    const size_t npix = (size_t)info.width * info.height;
    if (probe_ag && launch_png_probe(d_raw, npix, info.channels, d_flags, st)) { err = "png probe launch failed"; return false; }
    if (probe_pal && launch_png_colours(d_raw, npix, info.channels, d_set, d_flags, st)) { err = "png palette probe launch failed"; return false; }
    uint32_t *h_flags = reinterpret_cast<uint32_t *>(h_small);
    unsigned long long *h_sums_in = reinterpret_cast<unsigned long long *>(h_z);
    const size_t npieces_in = (nin + 4095) / 4096;
    CUP(cudaMemcpyAsync(h_flags, d_flags, 16, cudaMemcpyDeviceToHost, st));
    CUP(cudaMemcpyAsync(h_flags + 4, d_un, 8, cudaMemcpyDeviceToHost, st));
    CUP(cudaMemcpyAsync(h_sums_in, d_sums_in, npieces_in * 16, cudaMemcpyDeviceToHost, st));
    CUP(stream_wait(st)); LT_MARK("host_wait");
    if (h_flags[5]) { err = "bad filter type"; corrupt = true; return false; }
    if (combine_adler(h_sums_in, nin) != stored_adler) { err = "Adler-32 mismatch"; corrupt = true; return false; }
    if (probe_pal && h_flags[2] <= 256) {
        // few colours: oxipng's palette reduction (first-appearance order, tRNS layout, bit packing) runs on the host over the
        // reconstructed samples, and the indexed image takes the raw-sample entry point
        std::vector<uint8_t> raw(nraw);
        CUP(cudaMemcpy(raw.data(), d_raw, nraw, cudaMemcpyDeviceToHost));
        if (png_reduce_palette(info, raw)) return compress(info, raw, level, stream_, zlib_stream, chosen, err);
    }
    return reduce_and_code(info, probe_ag, h_flags, level, stream_, zlib_stream, chosen, err);
}

bool PngDevice::compress(PngInfo &info, const std::vector<uint8_t> &raw_in, int level, void *stream_, std::vector<uint8_t> &zlib_stream, int *chosen, std::string &err)
{
    cudaStream_t st = (cudaStream_t)stream_;
    corrupt = false;
    const int h = (int)info.height; const size_t rb = info.row_bytes;
    const size_t nraw = raw_in.size();
    const size_t nmax = (size_t)h * (rb + 1) + 64;
    size_t cap = 0;
    if (!ensure_buffers(nraw, nmax, rb, st, err) || !input_buffer(nraw, cap, err)) return false;
    memcpy(h_raw, raw_in.data(), nraw);
    CUP(cudaMemcpyAsync(d_raw, h_raw, nraw, cudaMemcpyHostToDevice, st));
    uint32_t *h_flags = reinterpret_cast<uint32_t *>(h_small);
    const bool probe_ag = info.bit_depth == 8 && info.trns.empty() && (info.color_type == 2 || info.color_type == 4 || info.color_type == 6);
    if (probe_ag) {
        uint32_t *d_flags = d_hist;
        CUP(cudaMemsetAsync(d_flags, 0, 16, st));
        if (launch_png_probe(d_raw, (size_t)info.width * info.height, info.channels, d_flags, st)) { err = "png probe launch failed"; return false; }
        CUP(cudaMemcpyAsync(h_flags, d_flags, 16, cudaMemcpyDeviceToHost, st));
        CUP(stream_wait(st)); LT_MARK("host_wait");
    }
    return reduce_and_code(info, probe_ag, h_flags, level, stream_, zlib_stream, chosen, err);
}

// d_raw holds the reconstructed samples; h_flags the probe results.  Lossless reductions (oxipng reduction::*: opaque alpha, grey
// RGB -- 8-bit samples without tRNS only), then every strategy of the preset, then the winner is DEFLATE-coded on the device.
bool PngDevice::reduce_and_code(PngInfo &info, bool probed, const uint32_t *h_flags, int level, void *stream_, std::vector<uint8_t> &zlib_stream, int *chosen, std::string &err)
{
    cudaStream_t st = (cudaStream_t)stream_;
    int h = (int)info.height; size_t rb = info.row_bytes; int bpp = info.bpp;
    if (probed) {
        const size_t npix = (size_t)info.width * info.height;
        const bool has_alpha = info.color_type == 4 || info.color_type == 6, is_rgb = info.color_type == 2 || info.color_type == 6;
        const bool drop_alpha = has_alpha && h_flags[0] == 0, to_grey = is_rgb && h_flags[1] == 0;
        if (drop_alpha || to_grey) {
            int mask = 0, ch = info.channels;
            const int ncolor = is_rgb ? 3 : 1;
            for (int c = 0; c < ncolor; c++) if (!to_grey || c == 0) mask |= 1 << c;
            if (has_alpha && !drop_alpha) mask |= 1 << (ch - 1);
            if (launch_png_repack(d_raw, d_raw2, npix, ch, mask, st)) { err = "png repack launch failed"; return false; }
            std::swap(d_raw, d_raw2); std::swap(cap_raw, cap_raw2);
            const bool grey = to_grey || !is_rgb, alpha = has_alpha && !drop_alpha;
            info.color_type = grey ? (alpha ? 4 : 0) : (alpha ? 6 : 2);
            info.channels = (grey ? 1 : 3) + (alpha ? 1 : 0);
            info.bits_per_pixel = 8 * info.channels; info.bpp = info.channels; info.row_bytes = (size_t)info.width * info.channels;
            rb = info.row_bytes; bpp = info.bpp;
        }
    }
    const size_t n = (size_t)h * (rb + 1);
    const size_t nchunks = (n + kChunk - 1) / kChunk;
    // ---- try every strategy of the preset; keep the one whose token histogram promises the smallest stream
    const std::vector<int> strategies = png_level_strategies(level);
    int best_s = strategies[0]; double best_bits = -1; size_t best_k = 0;
    const size_t fstride = (n + 64 + 255) / 256 * 256;
    if (strategies.size() > 1) {
        // every trial keeps its filtered stream (K6 is not repeated for the winner)
        if (!growp(d_filt_all, cap_filt_all, fstride * strategies.size() + 64, false, err)) return false;
        for (size_t k = 0; k < strategies.size(); k++) {
            if (!run_strategy(strategies[k], h, (int)rb, bpp, st, err, d_filt_all + k * fstride, true, false)) return false;
            CUP(cudaMemcpyAsync(h_small + 1024 + k * 316 * 4, d_hist, 316 * 4, cudaMemcpyDeviceToHost, st));
        }
        CUP(stream_wait(st)); LT_MARK("host_wait");
        for (size_t k = 0; k < strategies.size(); k++) {
            const double bits = estimate_bits(reinterpret_cast<const uint32_t *>(h_small + 1024 + k * 316 * 4));
            if (best_bits < 0 || bits < best_bits) { best_bits = bits; best_s = strategies[k]; best_k = k; }
        }
    }
    if (chosen) *chosen = best_s;
    // ---- the winner, for real: full match search (fixed + hash candidates), tokens compacted, Adler-32 pieces of the filtered
    //      stream, DEFLATE coding -- all on the device
    uint8_t *const wfilt = strategies.size() > 1 ? d_filt_all + best_k * fstride : d_filt;
    if (!run_strategy(best_s, h, (int)rb, bpp, st, err, wfilt, strategies.size() <= 1, true)) return false;
    size_t tb = cap_temp;
    cub::DeviceScan::ExclusiveSum(d_temp, tb, d_counts, d_offsets, (int)nchunks, st); LT_MARK("cub_scan");
    int rc = launch_png_compact(d_tok, d_counts, d_offsets, nchunks, kChunk, d_out, st);
    if (!rc) rc = launch_png_adler(wfilt, n, d_sums, st);
    if (rc) { err = std::string("png compact/adler: ") + cudaGetErrorString((cudaError_t)rc); return false; }
    unsigned long long *d_total = reinterpret_cast<unsigned long long *>(d_hist + 1024);          // [0] payload bits, [1] tokens
    uint32_t *d_ntok = d_hist + 1032;
    k_png_ntok<<<1, 32, 0, st>>>(d_counts, d_offsets, nchunks, d_ntok);
    const bool host_huffman = [] { const char *e = getenv("B200_PNG_HUFFMAN"); return e && !strcmp(e, "host"); }();
    unsigned long long *h_total = reinterpret_cast<unsigned long long *>(h_small + 64);
    const size_t npieces = (n + 4095) / 4096;
    unsigned long long *h_sums = reinterpret_cast<unsigned long long *>(h_z + z_cap + 64);
    if (!host_huffman) {
        rc = launch_png_deflate(d_out, d_ntok, n, kBlockTokens, d_dfl, reinterpret_cast<uint32_t *>(d_z), z_cap, d_total, st);
        if (rc) { err = std::string("png deflate: ") + cudaGetErrorString((cudaError_t)rc); return false; }
        CUP(cudaMemcpyAsync(h_total, d_total, 16, cudaMemcpyDeviceToHost, st));
        CUP(cudaMemcpyAsync(h_sums, d_sums, npieces * 16, cudaMemcpyDeviceToHost, st));
        CUP(stream_wait(st)); LT_MARK("host_wait");
        const size_t zbytes = (size_t)((h_total[0] + 7) / 8);
        if (zbytes + 8 <= z_cap) {
            CUP(cudaMemcpyAsync(h_z, d_z, zbytes, cudaMemcpyDeviceToHost, st)); LT_MARK("d2h");
            CUP(stream_wait(st)); LT_MARK("host_wait");
            const uint32_t adler = combine_adler(h_sums, n);
            zlib_stream.resize(zbytes + 4);
            memcpy(zlib_stream.data(), h_z, zbytes);
            zlib_stream[0] = 0x78; zlib_stream[1] = 0xDA;
            zlib_stream[zbytes] = adler >> 24; zlib_stream[zbytes + 1] = adler >> 16; zlib_stream[zbytes + 2] = adler >> 8; zlib_stream[zbytes + 3] = adler;
            last_deflate_ms = 0;
            return true;
        }
        // does not fit the device buffer (cannot happen for Huffman-coded literals; kept as a guard): the host codes the tokens
    } else {
        CUP(cudaMemcpyAsync(h_total + 1, d_ntok, 4, cudaMemcpyDeviceToHost, st));
        CUP(cudaMemcpyAsync(h_sums, d_sums, npieces * 16, cudaMemcpyDeviceToHost, st));
        CUP(stream_wait(st)); LT_MARK("host_wait");
        h_total[1] &= 0xFFFFFFFFull;
    }
    const size_t ntok = (size_t)h_total[1];
    if (!growp(h_tok, cap_htok, (n + 64) * 4 + 64, true, err)) return false;
    CUP(cudaMemcpyAsync(h_tok, d_out, ntok * 4, cudaMemcpyDeviceToHost, st));
    CUP(stream_wait(st)); LT_MARK("host_wait");
    const auto td = std::chrono::steady_clock::now();
    deflate_tokens(h_tok, ntok, combine_adler(h_sums, n), zlib_stream);
    last_deflate_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - td).count();
    return true;
}

// LZ77 tokens of a byte plane that sits on the host (the alpha plane of a WebP with transparency: bpp 1, stride = width): K7 with
// its pixel / row candidates and the hash chains, parallel parse, compaction; the tokens come back to the host (vp8l_alpha.cpp codes them).
bool PngDevice::plane_tokens(const uint8_t *plane, size_t n, int stride, void *stream_, std::vector<uint32_t> &tokens, std::string &err)
{
    cudaStream_t st = (cudaStream_t)stream_;
    if (!n || stride < 1) { err = "empty plane"; return false; }
    if (!ensure_buffers(n, n, (size_t)stride, st, err)) return false;
    if (!growp(h_raw, cap_hraw, n + 4096 + 64, true, err) || !growp(h_tok, cap_htok, (n + 64) * 4 + 64, true, err)) return false;
    memcpy(h_raw, plane, n);
    CUP(cudaMemcpyAsync(d_filt, h_raw, n, cudaMemcpyHostToDevice, st));
    const size_t nchunks = (n + kChunk - 1) / kChunk;
    int rc = launch_png_match(d_filt, d_best, n, 1, stride, st);
    if (!rc) rc = launch_png_hashmatch(d_filt, d_best, n, d_hist + 2048, st);
    if (rc) { err = std::string("png kernels: ") + cudaGetErrorString((cudaError_t)rc); return false; }
    CUP(cudaMemsetAsync(d_hist, 0, 316 * 4, st));
    rc = launch_png_parse(d_best, d_filt, n, kChunk, d_tok, d_counts, d_hist, st);
    if (rc) { err = std::string("png parse: ") + cudaGetErrorString((cudaError_t)rc); return false; }
    size_t tb = cap_temp;
    cub::DeviceScan::ExclusiveSum(d_temp, tb, d_counts, d_offsets, (int)nchunks, st);
    rc = launch_png_compact(d_tok, d_counts, d_offsets, nchunks, kChunk, d_out, st);
    if (rc) { err = std::string("png compact: ") + cudaGetErrorString((cudaError_t)rc); return false; }
    uint32_t *d_ntok = d_hist + 1032;
    k_png_ntok<<<1, 32, 0, st>>>(d_counts, d_offsets, nchunks, d_ntok);
    uint32_t *h_n = reinterpret_cast<uint32_t *>(h_small + 64);
    CUP(cudaMemcpyAsync(h_n, d_ntok, 4, cudaMemcpyDeviceToHost, st));
    CUP(stream_wait(st));
    const size_t ntok = *h_n;
    if (ntok > n) { err = "token count out of range"; return false; }
    CUP(cudaMemcpyAsync(h_tok, d_out, ntok * 4, cudaMemcpyDeviceToHost, st));
    CUP(stream_wait(st));
    tokens.assign(h_tok, h_tok + ntok);
    return true;
}

// ---- stage entry points (b200_png_filter / b200_png_lz77): plain allocate-run-free, used by the parity tests ----------------
bool png_stage_filter(const uint8_t *raw, int h, int rb, int bpp, int strategy, uint8_t *filtered, std::string &err)
{
    uint8_t *d_raw = nullptr, *d_filt = nullptr; uint32_t *d_tlog = nullptr;
    const size_t nraw = (size_t)h * rb, n = (size_t)h * (rb + 1);
    bool ok = false;
    do {
        if (cudaMalloc(&d_raw, nraw + 64) != cudaSuccess || cudaMalloc(&d_filt, n + 64) != cudaSuccess || cudaMalloc(&d_tlog, ((size_t)rb + 8) * 4) != cudaSuccess) { err = "cudaMalloc failed"; break; }
        std::vector<uint32_t> t(rb + 2); png_make_tlog(t.data(), rb + 1);
        cudaMemcpy(d_tlog, t.data(), t.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(d_raw, raw, nraw, cudaMemcpyHostToDevice);
        if (launch_png_filter(d_raw, d_filt, h, rb, bpp, strategy, d_tlog, nullptr)) { err = "png filter launch failed"; break; }
        cudaError_t e = cudaMemcpy(filtered, d_filt, n, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { err = std::string("png filter: ") + cudaGetErrorString(e); break; }
        ok = true;
    } while (0);
    cudaFree(d_raw); cudaFree(d_filt); cudaFree(d_tlog);
    return ok;
}

bool png_stage_lz77(const uint8_t *filtered, size_t n, int bpp, int stride, std::vector<uint32_t> &tokens, uint32_t *hist, std::string &err)
{
    uint8_t *d_filt = nullptr, *d_temp = nullptr; uint32_t *d_best = nullptr, *d_tok = nullptr, *d_out = nullptr, *d_counts = nullptr, *d_offsets = nullptr, *d_hist = nullptr;
    const size_t nchunks = (n + kChunk - 1) / kChunk;
    bool ok = false;
    do {
        size_t tb = 0; cub::DeviceScan::ExclusiveSum((void *)nullptr, tb, d_counts, d_offsets, (int)nchunks);
        if (cudaMalloc(&d_filt, n + 64) != cudaSuccess || cudaMalloc(&d_best, n * 4 + 64) != cudaSuccess || cudaMalloc(&d_tok, n * 4 + 64) != cudaSuccess ||
            cudaMalloc(&d_out, n * 4 + 64) != cudaSuccess || cudaMalloc(&d_counts, nchunks * 4 + 4) != cudaSuccess || cudaMalloc(&d_offsets, nchunks * 4 + 4) != cudaSuccess ||
            cudaMalloc(&d_hist, (320 + 544) * 4) != cudaSuccess || cudaMalloc(&d_temp, tb + 256) != cudaSuccess) { err = "cudaMalloc failed"; break; }
        cudaMemset(d_filt + n, 0, 64);
        cudaMemcpy(d_filt, filtered, n, cudaMemcpyHostToDevice);
        cudaMemset(d_hist, 0, 316 * 4);
        if (launch_png_match(d_filt, d_best, n, bpp, stride, nullptr) || launch_png_hashmatch(d_filt, d_best, n, d_hist + 320, nullptr) || launch_png_parse(d_best, d_filt, n, kChunk, d_tok, d_counts, d_hist, nullptr)) { err = "png lz77 launch failed"; break; }
        cub::DeviceScan::ExclusiveSum(d_temp, tb, d_counts, d_offsets, (int)nchunks);
        if (launch_png_compact(d_tok, d_counts, d_offsets, nchunks, kChunk, d_out, nullptr)) { err = "png compact launch failed"; break; }
        uint32_t last[2];
        cudaMemcpy(&last[0], d_offsets + (nchunks - 1), 4, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaMemcpy(&last[1], d_counts + (nchunks - 1), 4, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { err = std::string("png lz77: ") + cudaGetErrorString(e); break; }
        tokens.resize((size_t)last[0] + last[1]);
        cudaMemcpy(tokens.data(), d_out, tokens.size() * 4, cudaMemcpyDeviceToHost);
        e = cudaMemcpy(hist, d_hist, 316 * 4, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { err = std::string("png lz77: ") + cudaGetErrorString(e); break; }
        ok = true;
    } while (0);
    cudaFree(d_filt); cudaFree(d_best); cudaFree(d_tok); cudaFree(d_out); cudaFree(d_counts); cudaFree(d_offsets); cudaFree(d_hist); cudaFree(d_temp);
    return ok;
}

} // namespace b200

// jpeg_gpudec.cu -- device entropy DECODER for baseline single-scan JPEG files: CUDA wrappers around the bodies in
// jpeg_gpudec_core.h (self-synchronising parallel Huffman decoding) plus byte un-stuffing and the DC prefix sums.
// Pass order: unstuff (count, scan, scatter) -> round 0 -> rounds (in groups, one host check per group) -> block-count
// scan -> write -> DC gather / scan / scatter.  Coefficients land directly in the transform kernels' input buffer, so the
// host never sees them.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <cstring>
#include "jpeg_gpudec.h"
#include "jpeg_gpuenc_plan.h"

namespace b200 {

using namespace gd;

#define CUD(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { err = std::string(#expr) + ": " + cudaGetErrorString(e_); return FAILED; } } while (0)

// ---- un-stuffing: drop the 0x00 that follows every 0xFF -----------------------------------------------------------------
__global__ void k_gd_unstuff_count(const uint8_t *__restrict__ raw, uint32_t n, uint32_t *__restrict__ cnt)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g * 16 >= n) return;
    uint32_t c = 0;
    for (uint32_t j = g * 16; j < g * 16 + 16 && j < n; j++) c += (j > 0 && raw[j] == 0 && raw[j - 1] == 0xFF);
    cnt[g] = c;
}
__global__ void k_gd_unstuff_scatter(const uint8_t *__restrict__ raw, uint32_t n, const uint32_t *__restrict__ off, uint8_t *__restrict__ out)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g * 16 >= n) return;
    uint32_t o = g * 16 - off[g];
    for (uint32_t j = g * 16; j < g * 16 + 16 && j < n; j++) { if (j > 0 && raw[j] == 0 && raw[j - 1] == 0xFF) continue; out[o++] = raw[j]; }
}

// ---- synchronisation rounds ------------------------------------------------------------------------------------------------
__global__ void k_gd_round0(const uint8_t *__restrict__ stream, const Geometry *__restrict__ gp, const DecTable *__restrict__ tabs,
                            DecState *__restrict__ A, uint8_t *__restrict__ chg, uint32_t *__restrict__ nblk)
{
    const Geometry g = *gp;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.nsub) return;
    NullSink sk; DecState st; st.p = i * g.subseq_bits; st.k = 0; st.b = 0;
    A[i] = decode_subsequence(stream, g, tabs, i, st, sk);
    nblk[i] = sk.nblk; chg[i] = 1;
}

// thread i restarts from exit i-1 of the previous round; if that exit did not change last round, neither can ours
__global__ void k_gd_round(const uint8_t *__restrict__ stream, const Geometry *__restrict__ gp, const DecTable *__restrict__ tabs,
                           const DecState *__restrict__ A, DecState *__restrict__ B, const uint8_t *__restrict__ chg_in, uint8_t *__restrict__ chg_out,
                           uint32_t *__restrict__ nblk, uint32_t *__restrict__ any_changed)
{
    const Geometry g = *gp;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.nsub) return;
    if (i > 0 && !chg_in[i - 1]) { B[i] = A[i]; chg_out[i] = 0; return; }
    NullSink sk; DecState st;
    if (i == 0) { st.p = 0; st.k = 0; st.b = 0; } else st = A[i - 1];
    const DecState o = decode_subsequence(stream, g, tabs, i, st, sk);
    B[i] = o; nblk[i] = sk.nblk;
    const bool c = !same_state(o, A[i]);
    chg_out[i] = c ? 1 : 0;
    if (c) atomicOr(any_changed, 1u);
}

struct DevWriteSink {
    const ge::Scan *scan; uint32_t cur, total;
    __device__ __forceinline__ void coef(int k, int v) { if (cur < total) { const ge::BlockRef r = ge::locate(*scan, (int)cur); const_cast<int16_t *>(r.blk)[k] = (int16_t)v; } }
    __device__ __forceinline__ void block_done() { cur++; }
};

__global__ void k_gd_write(const uint8_t *__restrict__ stream, const Geometry *__restrict__ gp, const DecTable *__restrict__ tabs,
                           const DecState *__restrict__ A, const uint32_t *__restrict__ first, const ge::Scan *__restrict__ scan)
{
    const Geometry g = *gp;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.nsub) return;
    DevWriteSink sk{scan, first[i], g.total_blocks};
    DecState st;
    if (i == 0) { st.p = 0; st.k = 0; st.b = 0; } else st = A[i - 1];
    decode_subsequence(stream, g, tabs, i, st, sk);
}

// ---- DC: gather differences component-major, inclusive scan, subtract the component's base, scatter ------------------
__device__ __forceinline__ uint32_t dc_slot_index(const ge::Scan &s, uint32_t u, int *slot, uint32_t *comp_start)
{   // position of scan-order unit u inside the component-major difference array
    if (s.ns == 1) { *slot = 0; *comp_start = 0; return u; }
    const uint32_t m = u / s.blocks_per_mcu; int q = (int)(u - m * s.blocks_per_mcu), i = 0; uint32_t start = 0;
    const uint32_t mcus = (uint32_t)s.mcux * s.mcuy;
    while (q >= s.hs[i] * s.vs[i]) { q -= s.hs[i] * s.vs[i]; start += mcus * s.hs[i] * s.vs[i]; i++; }
    *slot = i; *comp_start = start;
    return start + m * s.hs[i] * s.vs[i] + q;
}
__global__ void k_gd_dc_gather(const ge::Scan *__restrict__ scan, uint32_t total, int32_t *__restrict__ d)
{
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= total) return;
    const ge::Scan s = *scan;
    int slot; uint32_t cs;
    d[dc_slot_index(s, u, &slot, &cs)] = ge::locate(s, (int)u).blk[0];
}
__global__ void k_gd_dc_scatter(const ge::Scan *__restrict__ scan, uint32_t total, const int32_t *__restrict__ sum)
{
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= total) return;
    const ge::Scan s = *scan;
    int slot; uint32_t cs;
    const uint32_t idx = dc_slot_index(s, u, &slot, &cs);
    const int32_t base = cs ? sum[cs - 1] : 0;
    const_cast<int16_t *>(ge::locate(s, (int)u).blk)[0] = (int16_t)(sum[idx] - base);
}

// ---- host ---------------------------------------------------------------------------------------------------------------------
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
template <typename T> static bool growd(T *&p, size_t &cap, size_t need, bool host, std::string &err)
{
    if (need <= cap) return true;
    if (p) { if (host) cudaFreeHost(p); else cudaFree(p); }
    p = nullptr; cap = 0;
    size_t want = align_up(need + need / 4, 1 << 12);
    void *q = nullptr;
    cudaError_t e = host ? cudaHostAlloc(&q, want, cudaHostAllocDefault) : cudaMalloc(&q, want);
    if (e != cudaSuccess) { err = std::string(host ? "cudaHostAlloc: " : "cudaMalloc: ") + cudaGetErrorString(e); return false; }
    p = (T *)q; cap = want; return true;
}

GpuDecoder::~GpuDecoder()
{
    cudaFreeHost(h_raw); cudaFree(d_raw); cudaFree(d_stream); cudaFree(d_cnt); cudaFree(d_off); cudaFree(d_A); cudaFree(d_B); cudaFree(d_chgA); cudaFree(d_chgB);
    cudaFree(d_nblk); cudaFree(d_first); cudaFree(d_dc); cudaFree(d_dcs); cudaFree(d_par); cudaFreeHost(h_par); cudaFree(d_temp);
}

GpuDecoder::Result GpuDecoder::decode(const JpegReader &rd, const JpegReader::DeviceScan &ds, int16_t *d_coefs, void *stream_, std::string &err)
{
    cudaStream_t st = (cudaStream_t)stream_;
    const JpegGeom &g = rd.geom();
    const size_t nraw = ds.ecs_end - ds.ecs_begin;
    if (nraw >= (1ull << 28)) { err = "entropy-coded segment too large for the device decoder"; return FAILED; }
    const uint32_t nstream = (uint32_t)(nraw - ds.stuffed);
    Geometry G{};
    int q = 0;
    for (int c = 0; c < g.ncomp; c++) for (int k = 0; k < (g.ncomp == 1 ? 1 : g.hs[c] * g.vs[c]); k++) { if (q >= 10) { err = "MCU too large"; return FAILED; } G.dc_tbl[q] = ds.td[c]; G.ac_tbl[q] = ds.ta[c]; q++; }
    G.blocks_per_mcu = q;
    G.total_blocks = g.ncomp == 1 ? (uint32_t)(g.rbw[0] * g.rbh[0]) : (uint32_t)(g.mcux * g.mcuy * q);
    G.nbits = nstream * 8; G.subseq_bits = SUBSEQ_BITS; G.nsub = (G.nbits + G.subseq_bits - 1) / G.subseq_bits;
    if (G.nsub == 0) { err = "empty scan"; return FAILED; }
    GpuEncPlan plan; const int16_t *base = d_coefs;
    gpuenc_plan(g, false, &base, 1, plan);
    const uint32_t ngroups = (uint32_t)((nraw + 15) / 16);
    // ---- buffers
    const size_t par_bytes = align_up(sizeof(Geometry), 256) + align_up(sizeof(ge::Scan), 256) + align_up(sizeof(DecTable) * 8, 256) + align_up(4 * (MAX_ROUNDS + 2), 256);
    if (!growd(h_raw, cap_hraw, nraw + 64, true, err) || !growd(d_raw, cap_raw, nraw + 64, false, err) || !growd(d_stream, cap_stream, (size_t)nstream + 64, false, err) ||
        !growd(d_cnt, cap_cnt, (size_t)ngroups * 4 + 4, false, err) || !growd(d_off, cap_off, (size_t)ngroups * 4 + 4, false, err) ||
        !growd(d_A, cap_A, (size_t)G.nsub * sizeof(DecState), false, err) || !growd(d_B, cap_B, (size_t)G.nsub * sizeof(DecState), false, err) ||
        !growd(d_chgA, cap_chgA, G.nsub, false, err) || !growd(d_chgB, cap_chgB, G.nsub, false, err) ||
        !growd(d_nblk, cap_nblk, (size_t)G.nsub * 4, false, err) || !growd(d_first, cap_first, (size_t)G.nsub * 4, false, err) ||
        !growd(d_dc, cap_dc, (size_t)G.total_blocks * 4, false, err) || !growd(d_dcs, cap_dcs, (size_t)G.total_blocks * 4, false, err) ||
        !growd(d_par, cap_par, par_bytes, false, err) || !growd(h_par, cap_hpar, par_bytes, true, err)) return FAILED;
    size_t t1 = 0, t2 = 0, t3 = 0;
    cub::DeviceScan::ExclusiveSum((void *)nullptr, t1, d_cnt, d_off, (int)ngroups, st);
    cub::DeviceScan::ExclusiveSum((void *)nullptr, t2, d_nblk, d_first, (int)G.nsub, st);
    cub::DeviceScan::InclusiveSum((void *)nullptr, t3, d_dc, d_dcs, (int)G.total_blocks, st);
    if (!growd(d_temp, cap_temp, std::max(t1, std::max(t2, t3)) + 256, false, err)) return FAILED;
    // ---- parameters
    uint8_t *hp = h_par;
    const size_t o_geo = 0, o_scan = align_up(sizeof(Geometry), 256), o_tab = o_scan + align_up(sizeof(ge::Scan), 256), o_flag = o_tab + align_up(sizeof(DecTable) * 8, 256);
    memcpy(hp + o_geo, &G, sizeof(G));
    memcpy(hp + o_scan, &plan.scans[0], sizeof(ge::Scan));
    DecTable *ht = reinterpret_cast<DecTable *>(hp + o_tab);
    memset(ht, 0, sizeof(DecTable) * 8);
    for (int id = 0; id < 4; id++) for (int kind = 0; kind < 2; kind++) if (rd.dht_present(kind, id)) build_dec_table(rd.dht_bits(kind, id), rd.dht_vals(kind, id), ht[kind * 4 + id]);
    memset(hp + o_flag, 0, 4 * (MAX_ROUNDS + 2));
    memcpy(h_raw, rd.data() + ds.ecs_begin, nraw);
    CUD(cudaMemcpyAsync(d_par, h_par, par_bytes, cudaMemcpyHostToDevice, st));
    CUD(cudaMemcpyAsync(d_raw, h_raw, nraw, cudaMemcpyHostToDevice, st));
    const Geometry *dG = reinterpret_cast<const Geometry *>(d_par + o_geo);
    const ge::Scan *dS = reinterpret_cast<const ge::Scan *>(d_par + o_scan);
    const DecTable *dT = reinterpret_cast<const DecTable *>(d_par + o_tab);
    uint32_t *dF = reinterpret_cast<uint32_t *>(d_par + o_flag);
    uint32_t *hF = reinterpret_cast<uint32_t *>(h_par + o_flag);
    // ---- unstuff
    k_gd_unstuff_count<<<cdiv(ngroups, 128), 128, 0, st>>>(d_raw, (uint32_t)nraw, d_cnt);
    size_t tb = cap_temp;
    cub::DeviceScan::ExclusiveSum(d_temp, tb, d_cnt, d_off, (int)ngroups, st);
    k_gd_unstuff_scatter<<<cdiv(ngroups, 128), 128, 0, st>>>(d_raw, (uint32_t)nraw, d_off, d_stream);
    CUD(cudaMemsetAsync(d_coefs, 0, (size_t)g.total_coefs * 2, st));
    // ---- rounds
    const int gs = cdiv(G.nsub, 64);
    k_gd_round0<<<gs, 64, 0, st>>>(d_stream, dG, dT, d_A, d_chgA, d_nblk);
    DecState *A = d_A, *B = d_B; uint8_t *cA = d_chgA, *cB = d_chgB;
    int rounds = 0; bool converged = false;
    while (!converged && rounds < MAX_ROUNDS) {
        const int first_round = rounds;
        for (int r = 0; r < ROUNDS_PER_GROUP && rounds < MAX_ROUNDS; r++, rounds++) {
            k_gd_round<<<gs, 64, 0, st>>>(d_stream, dG, dT, A, B, cA, cB, d_nblk, dF + rounds);
            std::swap(A, B); std::swap(cA, cB);
        }
        CUD(cudaMemcpyAsync(hF + first_round, dF + first_round, 4 * (rounds - first_round), cudaMemcpyDeviceToHost, st));
        CUD(cudaStreamSynchronize(st));
        for (int r = first_round; r < rounds; r++) if (hF[r] == 0) { converged = true; break; }
    }
    rounds_used = rounds;
    if (!converged) return NOT_CONVERGED;
    // ---- block counts -> first block of each subsequence -> write -> DC
    tb = cap_temp;
    cub::DeviceScan::ExclusiveSum(d_temp, tb, d_nblk, d_first, (int)G.nsub, st);
    k_gd_write<<<gs, 64, 0, st>>>(d_stream, dG, dT, A, d_first, dS);
    k_gd_dc_gather<<<cdiv(G.total_blocks, 128), 128, 0, st>>>(dS, G.total_blocks, d_dc);
    tb = cap_temp;
    cub::DeviceScan::InclusiveSum(d_temp, tb, d_dc, d_dcs, (int)G.total_blocks, st);
    k_gd_dc_scatter<<<cdiv(G.total_blocks, 128), 128, 0, st>>>(dS, G.total_blocks, d_dcs);
    CUD(cudaGetLastError());
    return OK;
}

} // namespace b200

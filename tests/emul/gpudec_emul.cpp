// tests/emul/gpudec_emul.cpp -- TEST INFRASTRUCTURE.  Serial CPU run of the device entropy DECODER's passes
// (caesium-clt_b200/csrc/jpeg_gpudec_core.h), in the order jpeg_gpudec.cu launches them.  Not part of the product.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../caesium-clt_b200/csrc/jpeg_gpudec_core.h"
#include "../../caesium-clt_b200/csrc/jpeg_gpuenc_plan.h"

using namespace b200;

namespace {
struct WriteSink {
    const ge::Scan *scan; int16_t *base; uint32_t cur, total;
    void coef(int k, int v) { if (cur < total) { ge::BlockRef r = ge::locate(*scan, (int)cur); const_cast<int16_t *>(r.blk)[k] = (int16_t)v; } }
    void block_done() { cur++; }
};
}

// returns 0 ok, 10 not eligible for the device decoder, 11 no convergence within max_rounds, other = parse error
extern "C" int emul_gpu_decode(const uint8_t *jpeg, size_t len, int subseq_bits, int max_rounds, int16_t *out, long long out_cap, int *rounds_used)
{
    std::string err;
    JpegReader rd(jpeg, len);
    if (!rd.read_header(err)) return 1;
    JpegReader::DeviceScan ds;
    if (!rd.device_decodable(ds)) return 10;
    const JpegGeom &g = rd.geom();
    if (out_cap < g.total_coefs) return 2;
    // pass: unstuff
    std::vector<uint8_t> stream;
    for (size_t i = ds.ecs_begin; i < ds.ecs_end; i++) { stream.push_back(jpeg[i]); if (jpeg[i] == 0xFF && i + 1 < ds.ecs_end && jpeg[i + 1] == 0) i++; }
    const size_t stream_bytes = stream.size();
    stream.resize((stream_bytes + 3) / 4 * 4 + 16, 0xFF);          // word alignment + 0xFF padding, as the device buffer has
    gd::Geometry G{};
    int q = 0;
    for (int c = 0; c < g.ncomp; c++) for (int k = 0; k < (g.ncomp == 1 ? 1 : g.hs[c] * g.vs[c]); k++) { G.dc_tbl[q] = ds.td[c]; G.ac_tbl[q] = ds.ta[c]; q++; }
    G.blocks_per_mcu = q;
    G.total_blocks = g.ncomp == 1 ? (uint32_t)(g.rbw[0] * g.rbh[0]) : (uint32_t)(g.mcux * g.mcuy * q);
    G.nbits = (uint32_t)stream_bytes * 8; G.subseq_bits = (uint32_t)subseq_bits; G.nsub = (G.nbits + G.subseq_bits - 1) / G.subseq_bits;
    std::vector<gd::DecTable> tabs(8);
    for (int id = 0; id < 4; id++) for (int kind = 0; kind < 2; kind++) if (rd.dht_present(kind, id)) gd::build_dec_table(rd.dht_bits(kind, id), rd.dht_vals(kind, id), tabs[kind * 4 + id]);
    memset(out, 0, (size_t)g.total_coefs * 2);
    GpuEncPlan plan; const int16_t *base = out;
    gpuenc_plan(g, false, &base, 1, plan);
    const ge::Scan &scan = plan.scans[0];
    // pass: round 0
    std::vector<gd::DecState> A(G.nsub), B(G.nsub);
    std::vector<uint32_t> nblk(G.nsub);
    for (uint32_t i = 0; i < G.nsub; i++) { gd::NullSink sk; gd::DecState st{i * G.subseq_bits, 0, 0}; A[i] = gd::decode_subsequence(stream.data(), G, tabs.data(), i, st, sk); nblk[i] = sk.nblk; }
    int rounds = 0; bool changed = true;
    while (changed && rounds < max_rounds) {
        changed = false; rounds++;
        for (uint32_t i = 0; i < G.nsub; i++) {
            gd::NullSink sk; gd::DecState st = i ? A[i - 1] : gd::DecState{0, 0, 0};
            B[i] = gd::decode_subsequence(stream.data(), G, tabs.data(), i, st, sk); nblk[i] = sk.nblk;
            if (!gd::same_state(B[i], A[i])) changed = true;
        }
        A.swap(B);
    }
    if (rounds_used) *rounds_used = rounds;
    if (changed) return 11;
    // pass: prefix sum + write
    std::vector<uint32_t> first(G.nsub); { uint32_t run = 0; for (uint32_t i = 0; i < G.nsub; i++) { first[i] = run; run += nblk[i]; } }
    for (uint32_t i = 0; i < G.nsub; i++) {
        WriteSink sk{&scan, out, first[i], G.total_blocks};
        gd::DecState st = i ? A[i - 1] : gd::DecState{0, 0, 0};
        gd::decode_subsequence(stream.data(), G, tabs.data(), i, st, sk);
    }
    // pass: DC prefix sums per component in scan order
    int pred[4] = {0, 0, 0, 0};
    for (uint32_t u = 0; u < G.total_blocks; u++) { ge::BlockRef r = ge::locate(scan, (int)u); int16_t *b = const_cast<int16_t *>(r.blk); pred[r.slot] += b[0]; b[0] = (int16_t)pred[r.slot]; }
    return 0;
}

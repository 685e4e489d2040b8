// tests/emul/gpuenc_emul.cpp -- TEST INFRASTRUCTURE.  Runs the bodies of the GPU entropy-encoder kernels
// (caesium-clt_b200/csrc/jpeg_gpuenc_core.h, shared __host__ __device__ code) in plain serial loops, pass by pass in the
// order jpeg_gpuenc.cu launches them, so the block-parallel formulation can be checked against the sequential writer on
// a box without a GPU.  Not linked into the product library.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../caesium-clt_b200/csrc/jpeg_gpuenc_plan.h"

using namespace b200;

extern "C" int emul_gpu_transcode(const uint8_t *jpeg, size_t len, int progressive, uint8_t **out, size_t *out_len)
{
    std::string err;
    JpegReader rd(jpeg, len);
    if (!rd.read_header(err)) return 1;
    const JpegGeom &g = rd.geom();
    std::vector<int16_t> coefs((size_t)g.total_coefs);
    if (!rd.decode(coefs.data(), err)) return 2;
    jpeg_fill_dummy_blocks(g, coefs.data());
    GpuEncPlan plan;
    const int16_t *base = coefs.data();
    gpuenc_plan(g, progressive != 0, &base, 1, plan);
    const long long U = plan.total_units;
    std::vector<uint32_t> meta(U), gcount(U, 0), bitlen(U);
    std::vector<long long> evkey(U), prev_ev(U);
    std::vector<uint32_t> tsum(U); std::vector<unsigned long long> bitoff(U);
    // pass: classify
    for (const ge::Scan &s : plan.scans) for (int u = 0; u < s.nblocks; u++) {
        ge::BlockRef b = ge::locate(s, u);
        uint32_t m = ge::classify(s, b.blk);
        meta[s.unit_base + u] = m;
        evkey[s.unit_base + u] = ge::meta_event(m) ? s.unit_base + u : -1;
    }
    // scans (global, as CUB would do them)
    { long long run = -1; for (long long i = 0; i < U; i++) { prev_ev[i] = run; if (evkey[i] > run) run = evkey[i]; } }
    { uint32_t run = 0; for (long long i = 0; i < U; i++) { tsum[i] = run; run += (uint32_t)ge::meta_tail(meta[i]); } }
    // pass: groups
    for (const ge::Scan &s : plan.scans) {
        if (s.mode != ge::MODE_AC_FIRST && s.mode != ge::MODE_AC_REFINE) continue;
        for (int b = 0; b <= s.nblocks; b++) {
            if (b < s.nblocks && !ge::meta_event(meta[s.unit_base + b])) continue;
            long long pg = b < s.nblocks ? prev_ev[s.unit_base + b] : -2;
            int prev;
            if (b == s.nblocks) { prev = -1; for (int j = s.nblocks - 1; j >= 0; j--) if (ge::meta_event(meta[s.unit_base + j])) { prev = j; break; } }
            else prev = pg >= s.unit_base ? (int)(pg - s.unit_base) : -1;
            ge::assign_groups(meta.data() + s.unit_base, tsum.data() + s.unit_base, s.nblocks, prev, b, gcount.data() + s.unit_base);
        }
    }
    // pass: histogram
    std::vector<uint32_t> hist(plan.scans.size() * 4 * 256, 0);
    for (const ge::Scan &s : plan.scans) for (int u = 0; u < s.nblocks; u++) {
        uint32_t *h = hist.data() + (size_t)s.tab_base * 256;
        auto add = [h](int idx) { h[idx]++; };
        ge::HistSink<decltype(add)> sk(add);
        ge::gen_block(s, ge::locate(s, u), gcount[s.unit_base + u], sk);
    }
    // pass: tables
    std::vector<ge::Table> tabs(plan.scans.size() * 4);
    std::vector<int> cs(257), oth(257); std::vector<long long> fr(257);
    for (size_t si = 0; si < plan.scans.size(); si++) {
        bool need[2][2]; jpeg_scan_tables_needed(g, progressive != 0, plan.defs[si % plan.scans_per_image], need);
        for (int kind = 0; kind < 2; kind++) for (int t = 0; t < 2; t++) if (need[kind][t])
            ge::build_table(hist.data() + ((size_t)plan.scans[si].tab_base + kind * 2 + t) * 256, tabs[plan.scans[si].tab_base + kind * 2 + t], cs.data(), oth.data(), fr.data());
    }
    // pass: lengths + offsets
    for (const ge::Scan &s : plan.scans) for (int u = 0; u < s.nblocks; u++) {
        ge::LenSink sk; sk.tabs = tabs.data() + s.tab_base;
        ge::gen_block(s, ge::locate(s, u), gcount[s.unit_base + u], sk);
        bitlen[s.unit_base + u] = (uint32_t)sk.bits;
    }
    { unsigned long long run = 0; for (long long i = 0; i < U; i++) { bitoff[i] = run; run += bitlen[i]; } }
    // pass: emit
    std::vector<uint32_t> words((size_t)plan.total_words, 0);
    std::vector<unsigned long long> total(plan.scans.size());
    for (size_t si = 0; si < plan.scans.size(); si++) {
        const ge::Scan &s = plan.scans[si];
        const unsigned long long base = bitoff[s.unit_base];
        total[si] = (s.nblocks ? bitoff[s.unit_base + s.nblocks - 1] + bitlen[s.unit_base + s.nblocks - 1] : base) - base;
        if ((long long)((total[si] + 31) / 32) > s.word_cap) return 3;
        uint32_t *w = words.data();
        auto orw = [w](long long i, uint32_t v) { w[i] |= v; };
        for (int u = 0; u < s.nblocks; u++) {
            ge::EmitSink<decltype(orw)> sk(tabs.data() + s.tab_base, orw, s.word_base, bitoff[s.unit_base + u] - base);
            ge::gen_block(s, ge::locate(s, u), gcount[s.unit_base + u], sk);
            sk.finish();
        }
    }
    // pass: pad + stuff, then assemble
    std::vector<std::vector<uint8_t>> data(plan.scans.size());
    std::vector<EncodedScan> enc(plan.scans.size());
    for (size_t si = 0; si < plan.scans.size(); si++) {
        const ge::Scan &s = plan.scans[si];
        const unsigned long long nb = total[si];
        const size_t nbytes = (size_t)((nb + 7) / 8);
        std::vector<uint8_t> &d = data[si];
        for (size_t i = 0; i < nbytes; i++) {
            uint32_t wv = words[(size_t)s.word_base + i / 4];
            uint8_t byte = (uint8_t)(wv >> (24 - 8 * (i & 3)));
            if (i == nbytes - 1 && (nb & 7)) byte |= (uint8_t)((1u << (8 - (nb & 7))) - 1);
            d.push_back(byte);
            if (byte == 0xFF) d.push_back(0);
        }
        EncodedScan &e = enc[si];
        e.def = plan.defs[si]; e.data = d.data(); e.len = d.size();
        bool need[2][2]; jpeg_scan_tables_needed(g, progressive != 0, e.def, need);
        for (int kind = 0; kind < 2; kind++) for (int t = 0; t < 2; t++) {
            e.has_tab[kind][t] = need[kind][t];
            const ge::Table &tb = tabs[s.tab_base + kind * 2 + t];
            memcpy(e.bits[kind][t], tb.bits, 17); memcpy(e.vals[kind][t], tb.vals, 256); e.nvals[kind][t] = tb.nvals;
        }
    }
    JpegWriteOptions wo; wo.progressive = progressive != 0; wo.copy_jfif = true;
    std::vector<uint8_t> file;
    if (!jpeg_assemble(g, wo, &rd.meta(), enc.data(), (int)enc.size(), file, err)) return 4;
    *out = (uint8_t *)malloc(file.size()); memcpy(*out, file.data(), file.size()); *out_len = file.size();
    return 0;
}

// SWAR threshold masks (what the kernels compute) against the per-coefficient definition, on random blocks that mix small values,
// zeros and the extremes; returns the number of blocks that differ.
extern "C" int emul_masks_check(int nblocks, unsigned seed)
{
    int bad = 0;
    unsigned x = seed * 2654435761u + 1u;
    auto rnd = [&x]() { x ^= x << 13; x ^= x >> 17; x ^= x << 5; return x; };
    static const int16_t special[] = {0, 1, -1, 2, -2, 3, -3, 4, -4, 5, -5, 7, -8, 255, -256, 32767, -32767, -32768, 16384, -16384};
    for (int n = 0; n < nblocks; n++) {
        alignas(16) int16_t blk[64];
        const unsigned density = rnd() % 5;
        for (int k = 0; k < 64; k++) {
            const unsigned r = rnd();
            if (density < 4 && (r & 7u) > density * 2u) blk[k] = 0;
            else if (r & 0x100u) blk[k] = special[(r >> 9) % (sizeof(special) / sizeof(special[0]))];
            else blk[k] = (int16_t)(r >> 16);
        }
        const ge::Masks3 a = ge::make_masks3(blk), b = ge::make_masks3_reference(blk);
        if (a.m[0] != b.m[0] || a.m[1] != b.m[1] || a.m[2] != b.m[2]) bad++;
    }
    return bad;
}

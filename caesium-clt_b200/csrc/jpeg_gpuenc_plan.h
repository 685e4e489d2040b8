// jpeg_gpuenc_plan.h -- host-side planning for the block-parallel entropy encoder: turns a geometry + scan script into
// the ge::Scan descriptors the kernels (or the CPU emulation in tests/emul/) iterate.  Plain C++, no CUDA.
#pragma once
#include <algorithm>
#include <vector>
#include "jpeg_gpuenc_core.h"
#include "jpeg_host.h"

namespace b200 {

// Block-major view used by the encoder passes: one descriptor per (image, component); a thread owns one block, reads it
// once, and serves every scan that visits the block (e.g. luma: DC scan + two AC bands + the refinement scan).
struct BlockComp {
    const int16_t *coef;                // image base
    long long comp_off;
    int bw, bh, rbw, rbh, hs, vs;
    int q_base, blocks_per_mcu, mcux;   // position of the component's first block inside an MCU (interleaved scans)
    int nscan, scan_idx[6];             // indices into GpuEncPlan::scans
    long long mask_base;                // first block of this component in the batch-wide per-block mask array
};

struct GpuEncPlan {
    std::vector<BlockComp> comps;       // image-major
    int max_comp_blocks = 0;
    std::vector<ge::Scan> scans;        // image-major: scans of image 0, then image 1, ...
    std::vector<ScanDef> defs;          // one script (shared by all images of the batch)
    int scans_per_image = 0;
    long long units_per_image = 0, words_per_image = 0;
    long long total_units = 0, total_words = 0, total_comp_blocks = 0;
};

// coef_base[i] = device (or host) pointer to image i's coefficient buffer (geometry g, zigzag)
inline void gpuenc_plan(const JpegGeom &g, bool progressive, const int16_t *const *coef_base, int nimages, GpuEncPlan &p)
{
    ScanDef sc[16];
    const int ns = jpeg_scan_script(g, progressive, sc);
    p.defs.assign(sc, sc + ns);
    p.scans_per_image = ns;
    p.scans.clear();
    long long unit = 0, word = 0;
    for (int im = 0; im < nimages; im++) {
        for (int si = 0; si < ns; si++) {
            const ScanDef &d = sc[si];
            ge::Scan s{};
            s.coef = coef_base[im];
            s.ns = d.ns; s.Ss = d.Ss; s.Se = d.Se; s.Al = d.Al;
            if (!progressive) s.mode = ge::MODE_SEQ;
            else if (d.Ss == 0) s.mode = ge::MODE_DC_FIRST;          // the script has no DC refinement scans
            else s.mode = d.Ah == 0 ? ge::MODE_AC_FIRST : ge::MODE_AC_REFINE;
            s.blocks_per_mcu = 0;
            for (int i = 0; i < d.ns; i++) {
                const int c = d.ci[i];
                s.comp[i] = c; s.hs[i] = g.hs[c]; s.vs[i] = g.vs[c]; s.bw[i] = g.bw[c]; s.comp_off[i] = g.comp_offset[c]; s.tbl[i] = c ? 1 : 0;
                s.blocks_per_mcu += g.hs[c] * g.vs[c];
            }
            s.mcux = g.mcux; s.mcuy = g.mcuy;
            s.rbw = g.rbw[d.ci[0]]; s.rbh = g.rbh[d.ci[0]];
            s.nblocks = d.ns > 1 ? g.mcux * g.mcuy * s.blocks_per_mcu : s.rbw * s.rbh;
            s.unit_base = unit; unit += s.nblocks;
            s.tab_base = (int)p.scans.size() * 4;
            s.word_base = word; s.word_cap = (long long)s.nblocks * 32 + 64; word += s.word_cap;   // 128 B per block: the size of its coefficients
            p.scans.push_back(s);
        }
        if (im == 0) { p.units_per_image = unit; p.words_per_image = word; }
    }
    p.total_units = unit; p.total_words = word;
    p.comps.clear(); p.max_comp_blocks = 0; p.total_comp_blocks = 0;
    for (int im = 0; im < nimages; im++) {
        int qb = 0;
        for (int c = 0; c < g.ncomp; c++) {
            BlockComp bc{};
            bc.coef = coef_base[im]; bc.comp_off = g.comp_offset[c];
            bc.bw = g.bw[c]; bc.bh = g.bh[c]; bc.rbw = g.rbw[c]; bc.rbh = g.rbh[c]; bc.hs = g.hs[c]; bc.vs = g.vs[c];
            bc.q_base = qb; bc.mcux = g.mcux;
            bc.blocks_per_mcu = 0; for (int cc = 0; cc < g.ncomp; cc++) bc.blocks_per_mcu += g.hs[cc] * g.vs[cc];
            qb += g.hs[c] * g.vs[c];
            bc.nscan = 0;
            for (int si = 0; si < ns; si++) for (int i = 0; i < sc[si].ns; i++) if (sc[si].ci[i] == c && bc.nscan < 6) bc.scan_idx[bc.nscan++] = im * ns + si;
            bc.mask_base = p.total_comp_blocks; p.total_comp_blocks += (long long)bc.bw * bc.bh;
            p.comps.push_back(bc);
            p.max_comp_blocks = std::max(p.max_comp_blocks, bc.bw * bc.bh);
        }
    }
}

} // namespace b200

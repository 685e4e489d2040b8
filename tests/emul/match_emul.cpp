// CPU emulation of k_png_match (csrc/png_kernels.cu) through the shared bodies of csrc/png_match_core.h: the CTA's windows, the
// per-candidate comparison bit arrays (one ballot per 32 bytes on the device, a loop here) and best_of(), against the definition --
// a byte-compare loop per position and candidate in the order ties are resolved.  Test infrastructure only.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../caesium-clt_b200/csrc/png_match_core.h"

using namespace b200::pm;

static void reference(const uint8_t *s, size_t n, int bpp, int stride, int chunk, uint32_t *best)
{
    int cand[NCAND]; candidates(bpp, stride, cand);
    for (size_t i = 0; i < n; i++) {
        const size_t chunk_end = (i / chunk + 1) * (size_t)chunk;
        const int maxlen = (int)std::min<size_t>(MATCH_MAX, std::min(n, chunk_end) - i);
        int bl = 0, bd = 0;
        if (maxlen >= 3)
            for (int c = 0; c < NCAND; c++) {
                const int d = cand[c];
                if (d < 1 || d > 32768 || (size_t)d > i) continue;
                int l = 0; while (l < maxlen && s[i + l] == s[i + l - d]) l++;
                if (l > bl) { bl = l; bd = d; }
                if (bl == maxlen) break;
            }
        best[i] = bl >= 3 ? ((uint32_t)bl << 16) | (uint32_t)bd : 0u;
    }
}

static void bitmask(const uint8_t *s, size_t n, int bpp, int stride, int chunk, uint32_t *best)
{
    int cand[NCAND]; candidates(bpp, stride, cand);
    std::vector<uint8_t> w0(WIN0), w1(WIN1), w2(WIN2);
    uint32_t eq[NCAND][MATCH_WORDS];
    for (long long i0 = 0; i0 < (long long)n; i0 += MATCH_T) {
        const long long b0 = i0 - NEAR_BACK, b1 = i0 - stride - ROW_SLACK, b2 = i0 - 2ll * stride;
        auto at = [&](long long p) -> uint8_t { return p >= 0 && p < (long long)n ? s[p] : (uint8_t)0; };
        for (int k = 0; k < WIN0; k++) { w0[k] = at(b0 + k); if (k < WIN1) w1[k] = at(b1 + k); if (k < WIN2) w2[k] = at(b2 + k); }
        for (int c = 0; c < NCAND; c++) {
            const int d = cand[c];
            const bool usable = d >= 1 && d <= 32768;
            const uint8_t *src = (window_of(c) == 1 ? w1.data() : window_of(c) == 2 ? w2.data() : w0.data()) + window_base(c, d, stride);
            for (int k = 0; k < MATCH_WORDS; k++) {
                uint32_t m = 0;                                     // the device: four bytes per lane (eq_nibble), eight lanes OR-reduced into a word
                for (int l = 0; l < 8; l++) {
                    uint32_t a, b; memcpy(&a, &w0[NEAR_BACK + 32 * k + 4 * l], 4); memcpy(&b, src + 32 * k + 4 * l, 4);
                    m |= (usable ? eq_nibble(a, b) : 0u) << (4 * l);
                }
                eq[c][k] = m;
            }
        }
        for (int t = 0; t < MATCH_T; t++) {
            const size_t i = (size_t)i0 + t;
            if (i >= n) break;
            const size_t chunk_end = (i / chunk + 1) * (size_t)chunk;
            const int maxlen = (int)std::min<size_t>(MATCH_MAX, std::min(n, chunk_end) - i);
            best[i] = best_of(eq, cand, t, (unsigned long long)i, maxlen);
        }
    }
}

// returns the number of positions where the two disagree (0 = the formulation is right); first_bad receives the first such position
extern "C" long long emul_match_compare(const uint8_t *s, unsigned long long n, int bpp, int stride, int chunk, long long *first_bad)
{
    std::vector<uint32_t> a(n), b(n);
    reference(s, n, bpp, stride, chunk, a.data());
    bitmask(s, n, bpp, stride, chunk, b.data());
    long long bad = 0; *first_bad = -1;
    for (size_t i = 0; i < n; i++) if (a[i] != b[i]) { if (!bad) *first_bad = (long long)i; bad++; }
    return bad;
}

// tests/emul/quant_emul.cpp -- TEST INFRASTRUCTURE.  The sign-free biased quantiser the transform kernels execute
// (caesium-clt_b200/csrc/jpeg_kernels.h: make_quant_dev / quant_biased / quant_pack) against mozjpeg's quantize() rule
// (jcdctmgr.c: round half away from zero, divisor quantval << 3), for every int16 input.  Not part of the product.
#include <cstdint>
#include "../../caesium-clt_b200/csrc/jpeg_kernels.h"

using namespace b200;

static int ref_quant(int x, int qv)
{
    const int d = qv << 3;
    int t = x < 0 ? -x : x;
    t = (t + (d >> 1)) / d;
    return x < 0 ? -t : t;
}

// one table whose 64 entries are qv[0..63]; every x in [-32768, 32767] through every entry, packed pairwise as the kernel does.
// returns the number of wrong int16 results
extern "C" long long emul_quant_check(const uint16_t *qv)
{
    QuantDev q;
    make_quant_dev(qv, &q);
    long long bad = 0;
    for (int j = 0; j < 32; j++) {
        const int ke = 2 * j, ko = 2 * j + 1;
        for (int x = -32768; x <= 32767; x++) {
            const int xo = -x - 1 + ((x * 7) & 3);          // a different value in the odd slot (stays inside int16)
            const int xoc = xo < -32768 ? -32768 : xo > 32767 ? 32767 : xo;
            const uint32_t qe = quant_biased(x, q.m[ke], q.c[ke], q.sh[ke]), qo = quant_biased(xoc, q.m[ko], q.c[ko], q.sh[ko]);
            const uint32_t w = quant_pack(qe, qo, q.kpair[j]);
            const int16_t ge = (int16_t)(w & 0xFFFF), go = (int16_t)(w >> 16);
            if (ge != (int16_t)ref_quant(x, qv[ke] ? qv[ke] : 1)) bad++;
            if (go != (int16_t)ref_quant(xoc, qv[ko] ? qv[ko] : 1)) bad++;
        }
    }
    return bad;
}
extern "C" int emul_quant_any_shift(const uint16_t *qv) { QuantDev q; make_quant_dev(qv, &q); return (int)q.any_shift; }

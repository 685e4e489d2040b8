// vp8_host.cpp -- see vp8_host.h.  Entropy coding of the WebP leg stays on the host (north_star: "VP8 bool-coder on the host").
#include "vp8_host.h"
#include <cmath>
#include <cstring>
#include "vp8_tables.h"

namespace b200 {

int vp8_qindex(int quality)
{   // libwebp: c = quality / 100; linearised, cube root, scaled to the 0..127 index range (segment alpha 0)
    const double c = quality / 100.0, lin = c < 0.75 ? c * (2.0 / 3.0) : 2.0 * c - 1.0;
    const int q = (int)(127.0 * (1.0 - std::pow(lin, 1.0 / 3.0)));
    return q < 0 ? 0 : q > 127 ? 127 : q;
}

void vp8_quant_factors(int q, int f[6])
{
    f[0] = VP8_DC_Q[q]; f[1] = VP8_AC_Q[q];
    f[2] = 2 * VP8_DC_Q[q]; f[3] = VP8_AC_Q[q] * 155 / 100; if (f[3] < 8) f[3] = 8;
    f[4] = VP8_DC_Q[q > 117 ? 117 : q]; f[5] = VP8_AC_Q[q];
}

namespace {

// RFC 6386 section 7.3 arithmetic encoder: 8-bit probabilities, range kept in [128, 255], carries propagated into the bytes
// already written.
class BoolWriter {
public:
    std::vector<uint8_t> bytes;
    void put(int bit, int prob)
    {
        const uint32_t split = 1 + (((range_ - 1) * (uint32_t)prob) >> 8);
        if (bit) { low_ += split; range_ -= split; } else range_ = split;
        int shift = __builtin_clz(range_) - 24;            // range < 128  <=>  shift > 0
        range_ <<= shift;
        while (shift-- > 0) {
            if (low_ & 0x80000000u) carry();
            low_ <<= 1;
            if (--count_ == 0) { bytes.push_back((uint8_t)(low_ >> 24)); low_ &= 0xFFFFFFu; count_ = 8; }
        }
    }
    void literal(int value, int nbits) { for (int i = nbits - 1; i >= 0; i--) put((value >> i) & 1, 128); }
    void finish()
    {
        int c = count_; uint32_t v = low_;
        if (v & (1u << (32 - c))) carry();
        v <<= c & 7;
        for (c >>= 3; c > 0; c--) v <<= 8;
        for (int i = 0; i < 4; i++, v <<= 8) bytes.push_back((uint8_t)(v >> 24));
    }
private:
    uint32_t range_ = 255, low_ = 0; int count_ = 24;
    void carry() { size_t i = bytes.size(); while (i > 0 && bytes[i - 1] == 0xFF) bytes[--i] = 0; if (i > 0) bytes[i - 1]++; }
};

const uint8_t kBands[17] = {0, 1, 2, 3, 6, 4, 5, 6, 6, 6, 6, 6, 6, 6, 6, 7, 0};
inline const uint8_t *probs(int type, int band, int ctx) { return VP8_COEF_PROBS + ((type * 8 + band) * 3 + ctx) * 11; }

// RFC 6386 13.2: one block's tokens; `lv` = 16 levels in zigzag order.  Returns the "has coded coefficients" context flag.
int put_block(BoolWriter &w, const int16_t *lv, int type, int first, int ctx)
{
    static const uint8_t kCat3[] = {173, 148, 140}, kCat4[] = {176, 155, 140, 135}, kCat5[] = {180, 157, 141, 134, 130},
                         kCat6[] = {254, 254, 243, 230, 196, 177, 153, 140, 133, 130, 129};
    int last = -1;
    for (int i = 15; i >= first; i--) if (lv[i]) { last = i; break; }
    const uint8_t *p = probs(type, kBands[first], ctx);
    w.put(last >= 0, p[0]);
    if (last < 0) return 0;
    for (int n = first; n < 16;) {
        const int c = lv[n++], v = c < 0 ? -c : c;
        w.put(v != 0, p[1]);
        if (!v) { p = probs(type, kBands[n], 0); continue; }          // a zero is never followed by an end-of-block check
        w.put(v > 1, p[2]);
        if (v == 1) p = probs(type, kBands[n], 1);
        else {
            w.put(v > 4, p[3]);
            if (v <= 4) { w.put(v != 2, p[4]); if (v != 2) w.put(v == 4, p[5]); }
            else {
                w.put(v > 10, p[6]);
                if (v <= 10) {
                    w.put(v > 6, p[7]);
                    if (v <= 6) w.put(v == 6, 159); else { w.put(v >= 9, 165); w.put(!(v & 1), 145); }
                } else {
                    const int cat = v < 19 ? 0 : v < 35 ? 1 : v < 67 ? 2 : 3;       // DCT_CAT3..6: bases 11, 19, 35, 67
                    static const uint8_t *const tabs[4] = {kCat3, kCat4, kCat5, kCat6};
                    static const int nbits[4] = {3, 4, 5, 11}, base[4] = {11, 19, 35, 67};
                    w.put(cat >> 1, p[8]); w.put(cat & 1, p[9 + (cat >> 1)]);
                    for (int i = nbits[cat] - 1, t = 0; i >= 0; i--, t++) w.put(((v - base[cat]) >> i) & 1, tabs[cat][t]);
                }
            }
            p = probs(type, kBands[n], 2);
        }
        w.put(c < 0, 128);
        if (n == 16) break;
        w.put(n <= last, p[0]);
        if (n > last) break;
    }
    return 1;
}

void put_le(std::vector<uint8_t> &o, uint32_t v, int nbytes) { for (int i = 0; i < nbytes; i++) o.push_back((uint8_t)(v >> (8 * i))); }

} // namespace

bool vp8_write_file(int width, int height, int qindex, const int16_t *levels, const uint8_t *modes, std::vector<uint8_t> &out)
{
    const int mbw = (width + 15) >> 4, mbh = (height + 15) >> 4, nmb = mbw * mbh;
    if (width < 1 || height < 1 || width > 16383 || height > 16383) return false;
    int nskip = 0;
    for (int i = 0; i < nmb; i++) nskip += modes[4 * i + 2];
    const bool use_skip = nskip > 0;
    int skip_p = (int)(((long long)(nmb - nskip) * 255) / nmb); if (skip_p < 1) skip_p = 1; if (skip_p > 255) skip_p = 255;
    // ---- first partition: frame header (RFC 6386 9.2-9.11, 19.2) and the per-macroblock modes (19.3)
    BoolWriter hd;
    hd.literal(0, 1);                   // color_space
    hd.literal(0, 1);                   // clamping_type
    hd.literal(0, 1);                   // segmentation_enabled
    hd.literal(0, 1);                   // filter_type
    hd.literal(0, 6);                   // loop_filter_level: none, so the decoder's output is exactly the reconstruction of K8
    hd.literal(0, 3);                   // sharpness_level
    hd.literal(0, 1);                   // loop_filter_adj_enable
    hd.literal(0, 2);                   // log2_nbr_of_dct_partitions
    hd.literal(qindex, 7);              // y_ac_qi
    hd.literal(0, 5);                   // five delta-present flags, all clear
    hd.literal(0, 1);                   // refresh_entropy_probs
    for (int i = 0; i < 4 * 8 * 3 * 11; i++) hd.put(0, VP8_COEF_UPDATE_PROBS[i]);      // default token probabilities kept
    hd.literal(use_skip, 1);            // mb_no_coeff_skip
    if (use_skip) hd.literal(skip_p, 8);
    for (int i = 0; i < nmb; i++) {
        const int ym = modes[4 * i], uvm = modes[4 * i + 1];
        if (use_skip) hd.put(modes[4 * i + 2], skip_p);
        hd.put(1, 145);                                                                 // a 16x16 mode, not B_PRED
        const bool tm_or_h = ym == 1 || ym == 3;
        hd.put(tm_or_h, 156);
        if (tm_or_h) hd.put(ym == 1, 128); else hd.put(ym == 2, 163);
        hd.put(uvm != 0, 142);
        if (uvm != 0) { hd.put(uvm != 2, 114); if (uvm != 2) hd.put(uvm != 3, 183); }
    }
    hd.finish();
    if (hd.bytes.size() >= (1u << 19)) return false;
    // ---- token partition (13): contexts are the "has coefficients" flags of the blocks above and to the left
    BoolWriter tk;
    std::vector<uint8_t> top((size_t)mbw * 9, 0);
    for (int my = 0; my < mbh; my++) {
        uint8_t left[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int mx = 0; mx < mbw; mx++) {
            const size_t mb = (size_t)my * mbw + mx;
            uint8_t *t = &top[(size_t)mx * 9];
            if (use_skip && modes[4 * mb + 2]) { memset(t, 0, 9); memset(left, 0, 9); continue; }
            const int16_t *lv = levels + mb * 400;
            t[8] = left[8] = (uint8_t)put_block(tk, lv, 1, 0, t[8] + left[8]);
            for (int b = 0; b < 16; b++) { const int x = b & 3, y = b >> 2; t[x] = left[y] = (uint8_t)put_block(tk, lv + 16 * (1 + b), 0, 1, t[x] + left[y]); }
            for (int c = 0; c < 2; c++)
                for (int b = 0; b < 4; b++) { const int x = 4 + 2 * c + (b & 1), y = 4 + 2 * c + (b >> 1); t[x] = left[y] = (uint8_t)put_block(tk, lv + 16 * (17 + 4 * c + b), 2, 0, t[x] + left[y]); }
        }
    }
    tk.finish();
    // ---- RIFF container (WebP simple lossy format)
    const size_t vp8_size = 10 + hd.bytes.size() + tk.bytes.size(), riff_payload = 4 + 8 + vp8_size + (vp8_size & 1);
    out.clear(); out.reserve(8 + riff_payload);
    out.insert(out.end(), {'R', 'I', 'F', 'F'}); put_le(out, (uint32_t)riff_payload, 4);
    out.insert(out.end(), {'W', 'E', 'B', 'P', 'V', 'P', '8', ' '}); put_le(out, (uint32_t)vp8_size, 4);
    put_le(out, (1u << 4) | ((uint32_t)hd.bytes.size() << 5), 3);      // key frame (bit 0 clear), version 0, show_frame, first partition size
    out.insert(out.end(), {0x9d, 0x01, 0x2a}); put_le(out, (uint32_t)width, 2); put_le(out, (uint32_t)height, 2);
    out.insert(out.end(), hd.bytes.begin(), hd.bytes.end());
    out.insert(out.end(), tk.bytes.begin(), tk.bytes.end());
    if (vp8_size & 1) out.push_back(0);
    return true;
}

} // namespace b200

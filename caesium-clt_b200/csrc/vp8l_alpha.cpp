// vp8l_alpha.cpp -- see vp8l_alpha.h.  WebP lossless bitstream (the "VP8L" specification: LSB-first bits, canonical prefix codes
// of at most 15 bits, five codes per group: green + length prefixes, red, blue, alpha, distance prefixes), restricted to what an
// alpha plane needs: no transforms, no colour cache, no meta prefix image; red / blue / alpha are single-symbol codes (zero bits).
#include "vp8l_alpha.h"
#include "dfl_core.h"
#include <cstring>

namespace b200 {
namespace {

struct BitsLsb {
    std::vector<uint8_t> &o; uint64_t acc = 0; int n = 0;
    explicit BitsLsb(std::vector<uint8_t> &out) : o(out) {}
    void put(uint32_t v, int nb)
    {
        if (!nb) return;
        acc |= (uint64_t)(v & (nb >= 32 ? 0xFFFFFFFFu : ((1u << nb) - 1u))) << n; n += nb;
        while (n >= 8) { o.push_back((uint8_t)acc); acc >>= 8; n -= 8; }
    }
    void flush() { if (n > 0) { o.push_back((uint8_t)acc); acc = 0; n = 0; } }
};

// the specification's prefix coding of lengths and distance codes: value >= 1 -> (prefix symbol, extra bit count, extra bits)
inline void prefix_of(uint32_t v, int &sym, int &nx, uint32_t &xv)
{
    const uint32_t d = v - 1;
    if (d < 4) { sym = (int)d; nx = 0; xv = 0; return; }
    const int hb = dfl::hibit(d);
    sym = 2 * hb + (int)((d >> (hb - 1)) & 1u); nx = hb - 1; xv = d & ((1u << nx) - 1u);
}

// distance codes 1..120 name a pixel of the neighbourhood: (dy << 4) | (8 - dx)  (table of the specification, section 5.2.2)
const uint8_t kCodeToPlane[120] = {
    0x18, 0x07, 0x17, 0x19, 0x28, 0x06, 0x27, 0x29, 0x16, 0x1a, 0x26, 0x2a, 0x38, 0x05, 0x37, 0x39, 0x15, 0x1b, 0x36, 0x3a,
    0x25, 0x2b, 0x48, 0x04, 0x47, 0x49, 0x14, 0x1c, 0x35, 0x3b, 0x46, 0x4a, 0x24, 0x2c, 0x58, 0x45, 0x4b, 0x34, 0x3c, 0x03,
    0x57, 0x59, 0x13, 0x1d, 0x56, 0x5a, 0x23, 0x2d, 0x44, 0x4c, 0x55, 0x5b, 0x33, 0x3d, 0x68, 0x02, 0x67, 0x69, 0x12, 0x1e,
    0x66, 0x6a, 0x22, 0x2e, 0x54, 0x5c, 0x43, 0x4d, 0x65, 0x6b, 0x32, 0x3e, 0x78, 0x01, 0x77, 0x79, 0x53, 0x5d, 0x11, 0x1f,
    0x64, 0x6c, 0x42, 0x4e, 0x76, 0x7a, 0x21, 0x2f, 0x75, 0x7b, 0x31, 0x3f, 0x63, 0x6d, 0x52, 0x5e, 0x00, 0x74, 0x7c, 0x41,
    0x4f, 0x10, 0x20, 0x62, 0x6e, 0x30, 0x73, 0x7d, 0x51, 0x5f, 0x40, 0x72, 0x7e, 0x61, 0x6f, 0x50, 0x71, 0x7f, 0x60, 0x70};

struct PlaneCodes {                     // pixel distance -> distance code for one image width
    int width; uint32_t near_dist[120];
    explicit PlaneCodes(int w) : width(w)
    {
        for (int i = 0; i < 120; i++) {
            const int dy = kCodeToPlane[i] >> 4, dx = 8 - (kCodeToPlane[i] & 15);
            const long long d = (long long)dy * w + dx;
            near_dist[i] = d >= 1 ? (uint32_t)d : 0u;       // (a decoder clamps codes that point forward to distance 1; never chosen here)
        }
    }
    uint32_t code_of(uint32_t dist) const
    {
        if (dist <= 7u * (uint32_t)width + 8u)
            for (int i = 0; i < 120; i++) if (near_dist[i] == dist) return (uint32_t)i + 1u;
        return dist + 120u;
    }
};

struct PrefixCode { std::vector<uint8_t> len; std::vector<uint16_t> code; int used = 0; };

void make_code(const std::vector<uint32_t> &freq, PrefixCode &pc, dfl::HuffScratch &S)
{
    const int n = (int)freq.size();
    if (n < 1 || n > dfl::MAXSYM) return;
    pc.len.assign(n, 0); pc.code.assign(n, 0); pc.used = 0;
    for (int i = 0; i < n; i++) pc.used += freq[i] != 0;
    dfl::huff_lengths(freq.data(), n, 15, pc.len.data(), S);
    dfl::canon_codes(pc.len.data(), n, pc.code.data());
}
inline void put_sym(BitsLsb &bw, const PrefixCode &pc, int s) { if (pc.used > 1) bw.put(pc.code[s], pc.len[s]); }   // a one-symbol code takes no bits

// one prefix code in the bitstream (section 6.2.1 / 6.2.2 of the specification)
void write_code(BitsLsb &bw, const PrefixCode &pc, dfl::HuffScratch &S)
{
    const int n = (int)pc.len.size();
    int s0 = -1, s1 = -1;
    for (int i = 0; i < n; i++) if (pc.len[i]) { if (s0 < 0) s0 = i; else if (s1 < 0) s1 = i; }
    if (pc.used == 0) { bw.put(1, 1); bw.put(0, 1); bw.put(0, 1); bw.put(0, 1); return; }                // simple code, one symbol: 0
    if (pc.used <= 2 && s0 < 256 && (pc.used == 1 || s1 < 256)) {
        bw.put(1, 1); bw.put((uint32_t)pc.used - 1u, 1);
        if (s0 < 2) { bw.put(0, 1); bw.put((uint32_t)s0, 1); } else { bw.put(1, 1); bw.put((uint32_t)s0, 8); }
        if (pc.used == 2) bw.put((uint32_t)s1, 8);
        return;
    }
    // normal code: the lengths, run-length coded with zero runs (17: 3..10, 18: 11..138), themselves prefix coded (limit 7)
    struct Tk { uint8_t sym, extra; };
    std::vector<Tk> tk;
    for (int i = 0; i < n;) {
        if (pc.len[i]) { tk.push_back({pc.len[i], 0}); i++; continue; }
        int run = 1; while (i + run < n && !pc.len[i + run]) run++;
        i += run;
        while (run >= 11) { const int r = run > 138 ? 138 : run; tk.push_back({18, (uint8_t)(r - 11)}); run -= r; }
        if (run >= 3) { tk.push_back({17, (uint8_t)(run - 3)}); run = 0; }
        while (run-- > 0) tk.push_back({0, 0});
    }
    std::vector<uint32_t> cf(19, 0);
    for (const Tk &t : tk) cf[t.sym]++;
    PrefixCode cl; cl.len.assign(19, 0); cl.code.assign(19, 0);
    for (int i = 0; i < 19; i++) cl.used += cf[i] != 0;
    dfl::huff_lengths(cf.data(), 19, 7, cl.len.data(), S);
    dfl::canon_codes(cl.len.data(), 19, cl.code.data());
    static const uint8_t order[19] = {17, 18, 0, 1, 2, 3, 4, 5, 16, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15};
    int ncodes = 19; while (ncodes > 4 && !cl.len[order[ncodes - 1]]) ncodes--;
    bw.put(0, 1);
    bw.put((uint32_t)ncodes - 4u, 4);
    for (int i = 0; i < ncodes; i++) bw.put(cl.len[order[i]], 3);
    bw.put(0, 1);                                   // every symbol's length follows (no max_symbol)
    for (const Tk &t : tk) {
        put_sym(bw, cl, t.sym);
        if (t.sym == 17) bw.put(t.extra, 3); else if (t.sym == 18) bw.put(t.extra, 7);
    }
}

} // namespace

bool vp8l_alpha_from_tokens(const uint32_t *tok, size_t ntok, int width, int height, std::vector<uint8_t> &alph, int filter)
{
    const size_t npix = (size_t)width * height;
    // ---- tokens -> operations: consecutive copies at the same distance continue each other (K7 cuts matches at 258 and at its
    //      parse-chunk ends; this format allows 4096 pixels per copy)
    struct Op { uint32_t len, dist; };              // len 0: literal `dist`
    std::vector<Op> ops; ops.reserve(ntok);
    size_t pos = 0;
    for (size_t i = 0; i < ntok; i++) {
        const uint32_t t = tok[i];
        if (!(t & 0x80000000u)) { ops.push_back({0u, t & 0xFFu}); pos++; continue; }
        const uint32_t len = ((t >> 16) & 0xFFu) + 3u, dist = (t & 0xFFFFu) + 1u;
        if (dist > pos) return false;
        if (!ops.empty() && ops.back().len && ops.back().dist == dist && ops.back().len + len <= 4096u) ops.back().len += len;
        else ops.push_back({len, dist});
        pos += len;
    }
    if (pos != npix) return false;
    // ---- statistics and codes
    const PlaneCodes planes(width);
    std::vector<uint32_t> gf(256 + 24, 0), df(40, 0);
    for (Op &o : ops) {
        if (!o.len) { gf[o.dist]++; continue; }
        int s, nx; uint32_t xv;
        prefix_of(o.len, s, nx, xv); gf[256 + s]++;
        prefix_of(planes.code_of(o.dist), s, nx, xv); df[s]++;
    }
    dfl::HuffScratch S;
    PrefixCode green, dist, zero;
    make_code(gf, green, S); make_code(df, dist, S);
    zero.len.assign(256, 0); zero.code.assign(256, 0); zero.used = 0;
    // ---- the chunk: header byte (no pre-processing, the caller's prediction filter, lossless compression) + image stream
    alph.clear(); alph.reserve(npix / 8 + 64);
    alph.push_back((uint8_t)(0x01 | ((filter & 3) << 2)));
    BitsLsb bw(alph);
    bw.put(0, 1);                       // no transform
    bw.put(0, 1);                       // no colour cache
    bw.put(0, 1);                       // one prefix-code group
    write_code(bw, green, S);
    write_code(bw, zero, S); write_code(bw, zero, S); write_code(bw, zero, S);      // red, blue, alpha: always 0
    write_code(bw, dist, S);
    for (const Op &o : ops) {
        if (!o.len) { put_sym(bw, green, (int)o.dist); continue; }
        int s, nx; uint32_t xv;
        prefix_of(o.len, s, nx, xv); put_sym(bw, green, 256 + s); bw.put(xv, nx);
        prefix_of(planes.code_of(o.dist), s, nx, xv); put_sym(bw, dist, s); bw.put(xv, nx);
    }
    bw.flush();
    return true;
}

// residual of one row under filter f (1 horizontal, 2 vertical, 3 gradient); prev = the row above (unfiltered), nullptr for row 0.
// The first row is always predicted from the left, the first column of later rows from above (the conventions the decoder undoes).
static void alpha_filter_row(int f, const uint8_t *row, const uint8_t *prev, int w, uint8_t *out)
{
    if (!prev) { out[0] = row[0]; for (int x = 1; x < w; x++) out[x] = (uint8_t)(row[x] - row[x - 1]); return; }
    out[0] = (uint8_t)(row[0] - prev[0]);
    if (f == 1) for (int x = 1; x < w; x++) out[x] = (uint8_t)(row[x] - row[x - 1]);
    else if (f == 2) for (int x = 1; x < w; x++) out[x] = (uint8_t)(row[x] - prev[x]);
    else for (int x = 1; x < w; x++) { const int g = (int)row[x - 1] + prev[x] - prev[x - 1]; out[x] = (uint8_t)(row[x] - (g < 0 ? 0 : g > 255 ? 255 : g)); }
}

int webp_alpha_choose_filter(const uint8_t *alpha, int width, int height, std::vector<uint8_t> &filtered)
{
    // order-0 cost of the residuals of every fourth row, per filter, in 1/1024 bit (integer log2 as the PNG leg's literal costs)
    uint32_t hist[4][256];
    memset(hist, 0, sizeof(hist));
    std::vector<uint8_t> tmp((size_t)width);
    size_t rows = 0;
    for (int y = 0; y < height; y += 4, rows++) {
        const uint8_t *row = alpha + (size_t)y * width, *prev = y ? row - width : nullptr;
        for (int x = 0; x < width; x++) hist[0][row[x]]++;
        for (int f = 1; f < 4; f++) { alpha_filter_row(f, row, prev, width, tmp.data()); for (int x = 0; x < width; x++) hist[f][tmp[x]]++; }
    }
    const unsigned long long total = (unsigned long long)rows * width;
    auto log2q = [](unsigned long long v) { int e = 63; while (!((v >> e) & 1ull)) e--; const unsigned long long fr = e >= 10 ? (v >> (e - 10)) & 1023ull : (v << (10 - e)) & 1023ull; return (unsigned long long)e * 1024ull + fr; };
    int best = 0; unsigned long long best_cost = ~0ull;
    for (int f = 0; f < 4; f++) {
        unsigned long long cost = 0;
        for (int v = 0; v < 256; v++) if (hist[f][v]) cost += hist[f][v] * (log2q(total) - log2q(hist[f][v]));
        if (cost < best_cost) { best_cost = cost; best = f; }            // ties: the simpler filter
    }
    if (best) {
        filtered.resize((size_t)width * height);
        for (int y = 0; y < height; y++) alpha_filter_row(best, alpha + (size_t)y * width, y ? alpha + (size_t)(y - 1) * width : nullptr, width, filtered.data() + (size_t)y * width);
    }
    return best;
}

bool webp_wrap_alpha(const std::vector<uint8_t> &f, const std::vector<uint8_t> &alph, int width, int height, std::vector<uint8_t> &out)
{
    if (f.size() < 20 || memcmp(f.data(), "RIFF", 4) || memcmp(f.data() + 8, "WEBPVP8 ", 8)) return false;
    const uint32_t vsz = (uint32_t)f[16] | ((uint32_t)f[17] << 8) | ((uint32_t)f[18] << 16) | ((uint32_t)f[19] << 24);
    if ((size_t)vsz + 20 > f.size()) return false;
    auto u32 = [&](uint32_t v) { for (int i = 0; i < 4; i++) out.push_back((uint8_t)(v >> (8 * i))); };
    auto u24 = [&](uint32_t v) { for (int i = 0; i < 3; i++) out.push_back((uint8_t)(v >> (8 * i))); };
    auto tag = [&](const char *t) { out.insert(out.end(), t, t + 4); };
    const size_t asz = alph.size();
    const size_t total = 4 + (8 + 10) + (8 + asz + (asz & 1)) + (8 + vsz + (vsz & 1));
    out.clear(); out.reserve(total + 8);
    tag("RIFF"); u32((uint32_t)total); tag("WEBP");
    tag("VP8X"); u32(10); out.push_back(0x10); u24(0); u24((uint32_t)width - 1); u24((uint32_t)height - 1);
    tag("ALPH"); u32((uint32_t)asz); out.insert(out.end(), alph.begin(), alph.end()); if (asz & 1) out.push_back(0);
    tag("VP8 "); u32(vsz); out.insert(out.end(), f.begin() + 20, f.begin() + 20 + vsz); if (vsz & 1) out.push_back(0);
    return true;
}

} // namespace b200

// jpeg_gpudec_core.h -- parallel decoding of a baseline (sequential Huffman, single interleaved or single-component
// scan, no restart markers) JPEG entropy-coded segment, written as __host__ __device__ code shared by the CUDA kernels
// (jpeg_gpudec.cu) and the serial CPU emulation in tests/emul/.  This is the decode half of SURVEY.md §8f rank 1; it
// replaces the host's jdhuff.c-style loop in front of caesium::compress_in_memory (/root/reference/src/compressor.rs:305).
//
// Method (self-synchronising Huffman decoding, Klein & Wiseman 2003; Weissenberger & Schmidt 2018): the unstuffed bit
// stream is cut into subsequences of SUBSEQ_BITS bits.  A decoder state is (p, k, b) = bit position of the next
// codeword, zigzag index inside the current block (0 = a DC symbol is next), index of the block inside its MCU.
//   round 0 : thread i decodes subsequence i from the guess (i * S, 0, 0) and records where and in which state it
//             leaves the subsequence (its "exit").
//   round r : thread i restarts from thread i-1's recorded exit and re-decodes; when nobody's exit changes the exits
//             are the true ones by induction from thread 0.  Huffman codes re-synchronise quickly on photographic
//             data, so a handful of rounds suffice; streams that do not converge within the round budget (degenerate
//             periodic content) are handed to the host decoder by the caller.
//   count   : blocks completed per subsequence -> exclusive prefix sum -> first block index of every subsequence.
//   write   : thread i decodes once more from its true state, writing coefficients (DC as the raw difference).
//   dc      : per-component prefix sum over the DC differences in scan order.
#pragma once
#include <cstdint>
#include "jpeg_gpuenc_core.h"      // ge::Scan / ge::locate: scan-order unit -> block address

namespace b200 {
namespace gd {

struct DecTable {               // one Huffman table, decode form (jdhuff.c jpeg_make_d_derived_tbl)
    uint16_t look[512];         // 9-bit lookahead: (len << 8) | symbol, 0 = code longer than 9 bits (or invalid)
    int32_t maxcode[18];        // maxcode[l] = largest code of length l (-1 if none); maxcode[17] = sentinel
    int32_t valoff[17];         // vals index = valoff[l] + code
    uint8_t vals[256];
};

struct DecState { uint32_t p; uint16_t k; uint16_t b; };   // 8 bytes

GE_HD bool same_state(const DecState &a, const DecState &b) { return a.p == b.p && a.k == b.k && a.b == b.b; }

// 32 bits of the unstuffed stream starting at bit position p (big-endian bit order).  The stream buffer is 4-byte
// aligned and followed by at least 8 bytes of 0xFF padding, so bits past the end read as 1s (like jdhuff.c's padding)
// and two aligned word loads always suffice.
GE_HD uint32_t load_be32(const uint8_t *__restrict__ s, uint32_t word)
{
    const uint32_t w = reinterpret_cast<const uint32_t *>(s)[word];
#if defined(__CUDA_ARCH__)
    return __byte_perm(w, 0, 0x0123);
#else
    return __builtin_bswap32(w);
#endif
}
GE_HD uint32_t peek32(const uint8_t *__restrict__ s, uint32_t /*nbits_total*/, uint32_t p)
{
    const uint32_t wi = p >> 5, sh = p & 31;
    const uint32_t w0 = load_be32(s, wi), w1 = load_be32(s, wi + 1);
    return sh ? (w0 << sh) | (w1 >> (32 - sh)) : w0;
}

// decode one Huffman symbol from the top bits of `bits` (32 valid bits); returns symbol, *len = code length (>= 1)
GE_HD int decode_symbol(const DecTable &t, uint32_t bits, int *len)
{
    const uint32_t e = t.look[bits >> 23];
    if (e) { *len = (int)(e >> 8); return (int)(e & 0xFF); }
    const int code16 = (int)(bits >> 16);
    for (int l = 10; l <= 16; l++) {
        const int c = code16 >> (16 - l);
        if (c <= t.maxcode[l]) { *len = l; return t.vals[(t.valoff[l] + c) & 0xFF]; }
    }
    *len = 16;                  // invalid code (only reachable while unsynchronised or on corrupt data): skip it
    return 0;
}

struct Geometry {               // what the decoder needs to know about the scan
    int blocks_per_mcu;
    int dc_tbl[10], ac_tbl[10]; // Huffman table ids (0..3) used by block q of the MCU
    uint32_t total_blocks;      // blocks in the scan
    uint32_t nbits;             // length of the unstuffed stream in bits
    uint32_t subseq_bits;
    uint32_t nsub;              // number of subsequences
};

// Decode from state `st` until the position leaves subsequence `i` (p >= (i+1)*S) or the stream ends.  Sink receives
// coef(block_ordinal_since_start, k, value) for every coefficient (DC as raw difference) and block_done().
template <class Sink>
GE_HD DecState decode_subsequence(const uint8_t *__restrict__ stream, const Geometry &g, const DecTable *__restrict__ tabs /*[0..3] DC ids, [4..7] AC ids*/,
                                  uint32_t i, DecState st, Sink &sk)
{
    const uint32_t end = (i + 1) * g.subseq_bits < g.nbits ? (i + 1) * g.subseq_bits : g.nbits;
    uint32_t p = st.p; int k = st.k, b = st.b;
    // 64-bit bit buffer in registers, left aligned at position p, refilled one aligned word at a time: the stream load is
    // off the critical path except once every 32 consumed bits (a symbol consumes at most 16 + 15 bits, so >= 32 buffered
    // bits always suffice)
    uint32_t nextw = (p >> 5) + 2;
    unsigned long long buf = (((unsigned long long)load_be32(stream, p >> 5) << 32) | load_be32(stream, (p >> 5) + 1)) << (p & 31);
    int have = 64 - (int)(p & 31);
    uint32_t pbuf = p;                              // position the buffer is aligned to
    while (p < end) {
        { const int used = (int)(p - pbuf); if (used) { buf <<= used; have -= used; pbuf = p; } }
        if (have < 32) { buf |= (unsigned long long)load_be32(stream, nextw++) << (32 - have); have += 32; }
        const uint32_t bits = (uint32_t)(buf >> 32);
        // One uniform body for DC and AC symbols (a DC symbol is "run 0, category s at index 0"), selects instead of
        // branches: the lanes of a warp sit at unrelated points of their blocks, so divergent paths would serialise.
        const bool dc = k == 0;
        int len;
        const int sym = decode_symbol(tabs[dc ? g.dc_tbl[b] : 4 + g.ac_tbl[b]], bits, &len);
        const int r = dc ? 0 : (sym >> 4), s = sym & 15;
        const uint32_t ext = s ? (bits << len) >> (32 - s) : 0u;
        const int v = s ? ((int)ext < (1 << (s - 1)) ? (int)ext - (1 << s) + 1 : (int)ext) : 0;
        const bool eob_or_zrl = !dc && s == 0;                                     // AC symbol without a value: ZRL or EOB
        int kw = k + r; if (kw > 63) kw = 63;                                       // corrupt / unsynchronised run: clamp
        if (!eob_or_zrl) sk.coef(kw, v);
        k = eob_or_zrl ? (r == 15 ? k + 16 : 64) : kw + 1;
        p += (uint32_t)(len + s);
        if (k >= 64) { k = 0; b++; if (b == g.blocks_per_mcu) b = 0; sk.block_done(); }
    }
    DecState o; o.p = p; o.k = (uint16_t)k; o.b = (uint16_t)b;
    return o;
}

struct NullSink { uint32_t nblk = 0; GE_HD void coef(int, int) {} GE_HD void block_done() { nblk++; } };

// jdhuff.c jpeg_make_d_derived_tbl from the DHT payload
inline void build_dec_table(const uint8_t bits[17], const uint8_t *vals, DecTable &t)
{
    for (int i = 0; i < 512; i++) t.look[i] = 0;
    for (int i = 0; i < 256; i++) t.vals[i] = vals[i];
    int code = 0, p = 0;
    for (int l = 1; l <= 16; l++) {
        t.valoff[l] = p - code;
        for (int i = 0; i < bits[l]; i++, p++, code++) {
            if (l <= 9) {
                const int first = code << (9 - l), cnt = 1 << (9 - l);
                if (first + cnt <= 512) for (int k = 0; k < cnt; k++) t.look[first + k] = (uint16_t)((l << 8) | vals[p]);
            }
        }
        t.maxcode[l] = bits[l] ? code - 1 : -1;
        code <<= 1;
    }
    t.maxcode[0] = -1; t.maxcode[17] = 0x7FFFFFFF; t.valoff[0] = 0;
}

} // namespace gd
} // namespace b200

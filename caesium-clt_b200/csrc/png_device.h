// png_device.h -- per-worker device state of the lossless PNG path (see png_device.cu).
#pragma once
#include <cstdint>
#include <cstddef>
#include <cmath>
#include <string>
#include <vector>
#include "png_host.h"

namespace b200 {

// strategies tried per oxipng optimisation level 0..6 (PngStrategy values)
std::vector<int> png_level_strategies(int level);

struct PngDevice {
    uint8_t *d_raw = nullptr, *d_raw2 = nullptr, *d_filt = nullptr, *d_temp = nullptr;
    uint32_t *d_best = nullptr, *d_tok = nullptr, *d_out = nullptr, *d_counts = nullptr, *d_offsets = nullptr, *d_hist = nullptr, *d_tlog = nullptr;
    unsigned long long *d_sums = nullptr;
    uint8_t *h_small = nullptr, *h_raw = nullptr;
    uint32_t *h_tok = nullptr;
    size_t cap_raw = 0, cap_raw2 = 0, cap_filt = 0, cap_best = 0, cap_tok = 0, cap_out = 0, cap_counts = 0, cap_offsets = 0, cap_hist = 0, cap_sums = 0,
           cap_tlog = 0, cap_temp = 0, cap_small = 0, cap_htok = 0, cap_hraw = 0;
    uint8_t *d_fin = nullptr, *d_dfl = nullptr, *d_z = nullptr, *h_z = nullptr; uint32_t *d_sync = nullptr; unsigned long long *d_sums_in = nullptr;
    size_t cap_fin = 0, cap_dfl = 0, cap_z = 0, cap_hz = 0, cap_sync = 0, cap_sums_in = 0, z_cap = 0;
    bool corrupt = false;                                // the last failure was the INPUT's fault (bad filter byte, Adler-32 mismatch)
    size_t tlog_n = 0;
    double last_deflate_ms = 0;                          // host Huffman/bit-packing time of the last compress() (tracing)
    ~PngDevice();

    // The lossless path: the caller inflates the IDAT stream into input_buffer() (pinned, `bytes` = height * (row_bytes + 1); the
    // buffer has 4096 bytes of slack for the inflate) and hands over its length and the stream's stored Adler-32; un-filtering,
    // checksum verification, reductions, K6 / K7 and the DEFLATE coding run on the device.
    uint8_t *input_buffer(size_t bytes, size_t &cap, std::string &err);
    bool compress_filtered(PngInfo &info, size_t nfilt, uint32_t stored_adler, int level, void *stream, std::vector<uint8_t> &zlib_stream, int *chosen_strategy, std::string &err);
    bool ensure_buffers(size_t nraw, size_t nmax, size_t rb, void *stream, std::string &err);
    bool reduce_and_code(PngInfo &info, bool probed, const uint32_t *h_flags, int level, void *stream, std::vector<uint8_t> &zlib_stream, int *chosen_strategy, std::string &err);
    // info/raw from png_decode; may rewrite info (colour-type reductions).  Produces the zlib stream of the re-filtered image.
    bool compress(PngInfo &info, const std::vector<uint8_t> &raw, int level, void *stream, std::vector<uint8_t> &zlib_stream, int *chosen_strategy, std::string &err);
    // filter + match + parse of d_raw with one strategy (results in d_filt / d_tok / d_counts / d_hist)
    bool run_strategy(int strategy, int h, int rb, int bpp, void *stream, std::string &err, uint8_t *filt = nullptr, bool do_filter = true, bool with_hash = true);
    // K7 over a byte plane on the host (bpp 1, stride = width) -> compacted LZ77 tokens on the host
    bool plane_tokens(const uint8_t *plane, size_t n, int stride, void *stream, std::vector<uint32_t> &tokens, std::string &err);
    uint8_t *d_filt_all = nullptr; size_t cap_filt_all = 0;          // the trials' filtered streams, one after another
};

// allocate-run-free stage helpers behind b200_png_filter / b200_png_lz77 (current device)
bool png_stage_filter(const uint8_t *raw, int h, int rb, int bpp, int strategy, uint8_t *filtered, std::string &err);
bool png_stage_lz77(const uint8_t *filtered, size_t n, int bpp, int stride, std::vector<uint32_t> &tokens, uint32_t *hist, std::string &err);

} // namespace b200

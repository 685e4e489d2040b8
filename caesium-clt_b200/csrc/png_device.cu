// png_device.cu -- device orchestration of the lossless PNG path (libcaesium png::lossless -> oxipng::optimize_from_memory,
// /root/reference/src/compressor.rs:428,436-437): upload the decoded samples, apply the cheap lossless reductions
// (opaque alpha, grey RGB), then for every row-filter strategy of the optimisation preset run K6 (filter) + K7 (match,
// parse) and estimate the DEFLATE size from the token histogram; the winning strategy's tokens come back to the host,
// which Huffman-codes and frames them (png_host.cpp).
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <cstring>
#include "png_device.h"
#include "png_kernels.h"
#include "stream_wait.h"

namespace b200 {

#define CUP(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { err = std::string(#expr) + ": " + cudaGetErrorString(e_); return false; } } while (0)

static const int kChunk = 4096;

template <typename T> static bool growp(T *&p, size_t &cap, size_t need, bool host, std::string &err)
{
    if (need <= cap) return true;
    if (p) { if (host) cudaFreeHost(p); else cudaFree(p); }
    p = nullptr; cap = 0;
    size_t want = 1 << 16; while (want < need + need / 4) want <<= 1;
    void *q = nullptr;
    cudaError_t e = host ? cudaHostAlloc(&q, want, cudaHostAllocDefault) : cudaMalloc(&q, want);
    if (e != cudaSuccess) { err = std::string(host ? "cudaHostAlloc: " : "cudaMalloc: ") + cudaGetErrorString(e); return false; }
    p = (T *)q; cap = want; return true;
}

PngDevice::~PngDevice()
{
    cudaFree(d_raw); cudaFree(d_raw2); cudaFree(d_filt); cudaFree(d_best); cudaFree(d_tok); cudaFree(d_out); cudaFree(d_counts); cudaFree(d_offsets);
    cudaFree(d_hist); cudaFree(d_sums); cudaFree(d_tlog); cudaFree(d_temp); cudaFreeHost(h_small); cudaFreeHost(h_tok); cudaFreeHost(h_raw);
}

// oxipng presets: which row-filter strategies each optimisation level tries
std::vector<int> png_level_strategies(int level)
{
    switch (level) {
        case 0: return {PNGF_NONE};
        case 1: return {PNGF_NONE, PNGF_BIGRAMS};
        case 2: return {PNGF_NONE, PNGF_SUB, PNGF_ENTROPY, PNGF_BIGRAMS};
        case 3: case 4: return {PNGF_NONE, PNGF_BIGRAMS, PNGF_BIGENT, PNGF_BRUTE};
        case 5: return {PNGF_NONE, PNGF_BIGRAMS, PNGF_BIGENT, PNGF_BRUTE, PNGF_UP, PNGF_MINSUM};
        default: return {PNGF_NONE, PNGF_BIGRAMS, PNGF_BIGENT, PNGF_BRUTE, PNGF_UP, PNGF_MINSUM, PNGF_AVERAGE, PNGF_PAETH};
    }
}

// estimated DEFLATE payload bits of a token histogram under its own optimal (unlimited) code: sum f * (log2(total/f)) + extra
static double estimate_bits(const uint32_t *h)
{
    static const uint8_t lx[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint8_t dx[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    double tl = 0, td = 0, bits = 0;
    for (int i = 0; i < 286; i++) tl += h[i];
    for (int i = 0; i < 30; i++) td += h[286 + i];
    for (int i = 0; i < 286; i++) if (h[i]) bits += h[i] * (std::log2(tl / h[i]) + (i >= 257 ? lx[i - 257] : 0));
    for (int i = 0; i < 30; i++) if (h[286 + i]) bits += h[286 + i] * (std::log2(td / h[286 + i]) + dx[i]);
    return bits;
}

bool PngDevice::run_strategy(int strategy, int h, int rb, int bpp, void *stream_, std::string &err)
{
    cudaStream_t st = (cudaStream_t)stream_;
    const size_t n = (size_t)h * (rb + 1);
    int rc = launch_png_filter(d_raw, d_filt, h, rb, bpp, strategy, d_tlog, st);
    if (!rc) rc = launch_png_match(d_filt, d_best, n, bpp, rb + 1, st);
    if (rc) { err = std::string("png kernels: ") + cudaGetErrorString((cudaError_t)rc); return false; }
    CUP(cudaMemsetAsync(d_hist, 0, 316 * 4, st));
    rc = launch_png_parse(d_best, d_filt, n, kChunk, d_tok, d_counts, d_hist, st);
    if (rc) { err = std::string("png parse: ") + cudaGetErrorString((cudaError_t)rc); return false; }
    return true;
}

bool PngDevice::compress(PngInfo &info, const std::vector<uint8_t> &raw_in, int level, void *stream_, std::vector<uint8_t> &zlib_stream, int *chosen, std::string &err)
{
    cudaStream_t st = (cudaStream_t)stream_;
    int h = (int)info.height; size_t rb = info.row_bytes; int bpp = info.bpp;
    const size_t nraw = raw_in.size();
    const size_t nmax = (size_t)h * (rb + 1) + 64;
    const size_t nchunks_max = (nmax + kChunk - 1) / kChunk;
    if (!growp(h_raw, cap_hraw, nraw + 64, true, err) || !growp(d_raw, cap_raw, nraw + 64, false, err) || !growp(d_raw2, cap_raw2, nraw + 64, false, err) ||
        !growp(d_filt, cap_filt, nmax + 64, false, err) || !growp(d_best, cap_best, nmax * 4, false, err) || !growp(d_tok, cap_tok, nmax * 4, false, err) ||
        !growp(d_out, cap_out, nmax * 4, false, err) || !growp(d_counts, cap_counts, nchunks_max * 4 + 4, false, err) || !growp(d_offsets, cap_offsets, nchunks_max * 4 + 4, false, err) ||
        !growp(d_hist, cap_hist, 316 * 4 * 16, false, err) || !growp(d_sums, cap_sums, ((nmax + 4095) / 4096) * 16 + 16, false, err) ||
        !growp(d_tlog, cap_tlog, (rb + 8) * 4, false, err) || !growp(h_small, cap_small, 1 << 16, true, err)) return false;
    size_t tb = 0; cub::DeviceScan::ExclusiveSum((void *)nullptr, tb, d_counts, d_offsets, (int)nchunks_max, st);
    if (!growp(d_temp, cap_temp, tb + 256, false, err)) return false;
    if (tlog_n < rb + 2) { std::vector<uint32_t> t(rb + 2); png_make_tlog(t.data(), rb + 1); CUP(cudaMemcpyAsync(d_tlog, t.data(), (rb + 2) * 4, cudaMemcpyHostToDevice, st)); CUP(stream_wait(st)); tlog_n = rb + 2; }
    memcpy(h_raw, raw_in.data(), nraw);
    CUP(cudaMemcpyAsync(d_raw, h_raw, nraw, cudaMemcpyHostToDevice, st));
    // ---- lossless reductions (oxipng reduction::*): 8-bit samples without tRNS only
    uint32_t *h_flags = reinterpret_cast<uint32_t *>(h_small);
    if (info.bit_depth == 8 && info.trns.empty() && (info.color_type == 2 || info.color_type == 4 || info.color_type == 6)) {
        uint32_t *d_flags = d_hist;
        CUP(cudaMemsetAsync(d_flags, 0, 8, st));
        const size_t npix = (size_t)info.width * info.height;
        int rc = launch_png_probe(d_raw, npix, info.channels, d_flags, st);
        if (rc) { err = "png probe launch failed"; return false; }
        CUP(cudaMemcpyAsync(h_flags, d_flags, 8, cudaMemcpyDeviceToHost, st));
        CUP(stream_wait(st));
        const bool has_alpha = info.color_type == 4 || info.color_type == 6, is_rgb = info.color_type == 2 || info.color_type == 6;
        const bool drop_alpha = has_alpha && h_flags[0] == 0, to_grey = is_rgb && h_flags[1] == 0;
        if (drop_alpha || to_grey) {
            int mask = 0, ch = info.channels;
            const int ncolor = is_rgb ? 3 : 1;
            for (int c = 0; c < ncolor; c++) if (!to_grey || c == 0) mask |= 1 << c;
            if (has_alpha && !drop_alpha) mask |= 1 << (ch - 1);
            rc = launch_png_repack(d_raw, d_raw2, npix, ch, mask, st);
            if (rc) { err = "png repack launch failed"; return false; }
            std::swap(d_raw, d_raw2); std::swap(cap_raw, cap_raw2);
            const bool grey = to_grey || !is_rgb, alpha = has_alpha && !drop_alpha;
            info.color_type = grey ? (alpha ? 4 : 0) : (alpha ? 6 : 2);
            info.channels = (grey ? 1 : 3) + (alpha ? 1 : 0);
            info.bits_per_pixel = 8 * info.channels; info.bpp = info.channels; info.row_bytes = (size_t)info.width * info.channels;
            rb = info.row_bytes; bpp = info.bpp;
        }
    }
    const size_t n = (size_t)h * (rb + 1);
    const size_t nchunks = (n + kChunk - 1) / kChunk;
    // ---- try every strategy of the preset; keep the one whose token histogram promises the smallest stream
    const std::vector<int> strategies = png_level_strategies(level);
    int best_s = strategies[0]; double best_bits = -1;
    if (strategies.size() > 1) {
        for (size_t k = 0; k < strategies.size(); k++) {
            if (!run_strategy(strategies[k], h, (int)rb, bpp, st, err)) return false;
            CUP(cudaMemcpyAsync(h_small + 1024 + k * 316 * 4, d_hist, 316 * 4, cudaMemcpyDeviceToHost, st));
        }
        CUP(stream_wait(st));
        for (size_t k = 0; k < strategies.size(); k++) {
            const double bits = estimate_bits(reinterpret_cast<const uint32_t *>(h_small + 1024 + k * 316 * 4));
            if (best_bits < 0 || bits < best_bits) { best_bits = bits; best_s = strategies[k]; }
        }
    }
    if (chosen) *chosen = best_s;
    // ---- the winner, for real: tokens compacted and brought back, Adler-32 pieces alongside
    if (!run_strategy(best_s, h, (int)rb, bpp, st, err)) return false;
    tb = cap_temp;
    cub::DeviceScan::ExclusiveSum(d_temp, tb, d_counts, d_offsets, (int)nchunks, st);
    int rc = launch_png_compact(d_tok, d_counts, d_offsets, nchunks, kChunk, d_out, st);
    if (!rc) rc = launch_png_adler(d_filt, n, d_sums, st);
    if (rc) { err = std::string("png compact/adler: ") + cudaGetErrorString((cudaError_t)rc); return false; }
    uint32_t *h_last = reinterpret_cast<uint32_t *>(h_small + 64);
    CUP(cudaMemcpyAsync(h_last, d_offsets + (nchunks - 1), 4, cudaMemcpyDeviceToHost, st));
    CUP(cudaMemcpyAsync(h_last + 1, d_counts + (nchunks - 1), 4, cudaMemcpyDeviceToHost, st));
    CUP(stream_wait(st));
    const size_t ntok = (size_t)h_last[0] + h_last[1];
    const size_t npieces = (n + 4095) / 4096;
    // sized for the worst case (one token per byte) so that images of one size never regrow it: cudaFreeHost / cudaHostAlloc
    // synchronise the whole device and stall every other worker
    if (!growp(h_tok, cap_htok, nmax * 4 + ((nmax + 4095) / 4096) * 16 + 64, true, err)) return false;
    CUP(cudaMemcpyAsync(h_tok, d_out, ntok * 4, cudaMemcpyDeviceToHost, st));
    unsigned long long *h_sums = reinterpret_cast<unsigned long long *>(h_tok + ((ntok * 4 + 15) / 16 * 16) / 4);
    CUP(cudaMemcpyAsync(h_sums, d_sums, npieces * 16, cudaMemcpyDeviceToHost, st));
    CUP(stream_wait(st));
    // Adler-32 of the filtered stream from the per-piece sums: a' = a + S, b' = b + len * a + T   (mod 65521)
    unsigned long long a = 1, b = 0;
    for (size_t p = 0; p < npieces; p++) {
        const unsigned long long len = std::min<size_t>(4096, n - p * 4096);
        b = (b + len * a + h_sums[2 * p + 1]) % 65521; a = (a + h_sums[2 * p]) % 65521;
    }
    const auto td = std::chrono::steady_clock::now();
    deflate_tokens(h_tok, ntok, (uint32_t)((b << 16) | a), zlib_stream);
    last_deflate_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - td).count();
    return true;
}

// ---- stage entry points (b200_png_filter / b200_png_lz77): plain allocate-run-free, used by the parity tests ----------------
bool png_stage_filter(const uint8_t *raw, int h, int rb, int bpp, int strategy, uint8_t *filtered, std::string &err)
{
    uint8_t *d_raw = nullptr, *d_filt = nullptr; uint32_t *d_tlog = nullptr;
    const size_t nraw = (size_t)h * rb, n = (size_t)h * (rb + 1);
    bool ok = false;
    do {
        if (cudaMalloc(&d_raw, nraw + 64) != cudaSuccess || cudaMalloc(&d_filt, n + 64) != cudaSuccess || cudaMalloc(&d_tlog, ((size_t)rb + 8) * 4) != cudaSuccess) { err = "cudaMalloc failed"; break; }
        std::vector<uint32_t> t(rb + 2); png_make_tlog(t.data(), rb + 1);
        cudaMemcpy(d_tlog, t.data(), t.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(d_raw, raw, nraw, cudaMemcpyHostToDevice);
        if (launch_png_filter(d_raw, d_filt, h, rb, bpp, strategy, d_tlog, nullptr)) { err = "png filter launch failed"; break; }
        cudaError_t e = cudaMemcpy(filtered, d_filt, n, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { err = std::string("png filter: ") + cudaGetErrorString(e); break; }
        ok = true;
    } while (0);
    cudaFree(d_raw); cudaFree(d_filt); cudaFree(d_tlog);
    return ok;
}

bool png_stage_lz77(const uint8_t *filtered, size_t n, int bpp, int stride, std::vector<uint32_t> &tokens, uint32_t *hist, std::string &err)
{
    uint8_t *d_filt = nullptr, *d_temp = nullptr; uint32_t *d_best = nullptr, *d_tok = nullptr, *d_out = nullptr, *d_counts = nullptr, *d_offsets = nullptr, *d_hist = nullptr;
    const size_t nchunks = (n + kChunk - 1) / kChunk;
    bool ok = false;
    do {
        size_t tb = 0; cub::DeviceScan::ExclusiveSum((void *)nullptr, tb, d_counts, d_offsets, (int)nchunks);
        if (cudaMalloc(&d_filt, n + 64) != cudaSuccess || cudaMalloc(&d_best, n * 4 + 64) != cudaSuccess || cudaMalloc(&d_tok, n * 4 + 64) != cudaSuccess ||
            cudaMalloc(&d_out, n * 4 + 64) != cudaSuccess || cudaMalloc(&d_counts, nchunks * 4 + 4) != cudaSuccess || cudaMalloc(&d_offsets, nchunks * 4 + 4) != cudaSuccess ||
            cudaMalloc(&d_hist, 316 * 4) != cudaSuccess || cudaMalloc(&d_temp, tb + 256) != cudaSuccess) { err = "cudaMalloc failed"; break; }
        cudaMemset(d_filt + n, 0, 64);
        cudaMemcpy(d_filt, filtered, n, cudaMemcpyHostToDevice);
        cudaMemset(d_hist, 0, 316 * 4);
        if (launch_png_match(d_filt, d_best, n, bpp, stride, nullptr) || launch_png_parse(d_best, d_filt, n, kChunk, d_tok, d_counts, d_hist, nullptr)) { err = "png lz77 launch failed"; break; }
        cub::DeviceScan::ExclusiveSum(d_temp, tb, d_counts, d_offsets, (int)nchunks);
        if (launch_png_compact(d_tok, d_counts, d_offsets, nchunks, kChunk, d_out, nullptr)) { err = "png compact launch failed"; break; }
        uint32_t last[2];
        cudaMemcpy(&last[0], d_offsets + (nchunks - 1), 4, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaMemcpy(&last[1], d_counts + (nchunks - 1), 4, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { err = std::string("png lz77: ") + cudaGetErrorString(e); break; }
        tokens.resize((size_t)last[0] + last[1]);
        cudaMemcpy(tokens.data(), d_out, tokens.size() * 4, cudaMemcpyDeviceToHost);
        e = cudaMemcpy(hist, d_hist, 316 * 4, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { err = std::string("png lz77: ") + cudaGetErrorString(e); break; }
        ok = true;
    } while (0);
    cudaFree(d_filt); cudaFree(d_best); cudaFree(d_tok); cudaFree(d_out); cudaFree(d_counts); cudaFree(d_offsets); cudaFree(d_hist); cudaFree(d_temp);
    return ok;
}

} // namespace b200

// webp_device.cu -- see webp_device.h.  caesium::convert_in_memory(.., WebP) (/root/reference/src/compressor.rs:288-292).
#include <cuda_runtime.h>
#include <cstring>
#include <chrono>
#include "webp_device.h"
#include "vp8_kernels.h"
#include "vp8_host.h"
#include "stream_wait.h"

namespace b200 {

#define CUW(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { err = std::string(#expr) + ": " + cudaGetErrorString(e_); return false; } } while (0)

template <typename T> static bool groww(T *&p, size_t &cap, size_t need, bool host, std::string &err)
{
    if (need <= cap) return true;
    if (p) { if (host) cudaFreeHost(p); else cudaFree(p); }
    p = nullptr; cap = 0;
    size_t want = 1 << 16; while (want < need) want <<= 1;
    void *q = nullptr;
    cudaError_t e = host ? cudaHostAlloc(&q, want, cudaHostAllocDefault) : cudaMalloc(&q, want);
    if (e != cudaSuccess) { err = std::string(host ? "cudaHostAlloc: " : "cudaMalloc: ") + cudaGetErrorString(e); return false; }
    p = (T *)q; cap = want; return true;
}

WebpDevice::~WebpDevice()
{
    cudaFree(d_planes); cudaFree(d_rgb); cudaFree(d_levels); cudaFree(d_modes); cudaFree(d_progress); cudaFreeHost(h_out); cudaFreeHost(h_rgb);
}

bool WebpDevice::encode_planes(const uint8_t *d_r, const uint8_t *d_g, const uint8_t *d_b, int w, int h, int quality, void *stream_,
                               std::vector<uint8_t> &out, std::string &err, int16_t *levels_out, uint8_t *modes_out)
{
    cudaStream_t st = (cudaStream_t)stream_;
    if (w < 1 || h < 1 || w > 16383 || h > 16383) { err = "WebP dimensions out of range"; return false; }
    if (quality < 0) quality = 0; if (quality > 100) quality = 100;
    Vp8Frame f; f.w = w; f.h = h; f.mbw = (w + 15) >> 4; f.mbh = (h + 15) >> 4;
    const size_t nmb = (size_t)f.mbw * f.mbh, ny = nmb * 256, nc = nmb * 64;
    const size_t lv_bytes = nmb * VP8_MB_COEFS * sizeof(int16_t), md_bytes = nmb * 4;
    if (!groww(d_planes, cap_planes, 2 * (ny + 2 * nc) + 256, false, err) || !groww(d_levels, cap_levels, lv_bytes, false, err) ||
        !groww(d_modes, cap_modes, md_bytes, false, err) || !groww(d_progress, cap_progress, sizeof(int) * (size_t)(f.mbh + 1), false, err) ||
        !groww(h_out, cap_hout, lv_bytes + md_bytes, true, err)) return false;
    uint8_t *Y = d_planes, *U = Y + ny, *V = U + nc;
    f.Y = Y; f.U = U; f.V = V; f.RY = V + nc; f.RU = f.RY + ny; f.RV = f.RU + nc;
    f.levels = d_levels; f.modes = d_modes; f.progress = d_progress;
    const int qi = vp8_qindex(quality);
    vp8_quant_factors(qi, f.q);
    int rc = launch_vp8_rgb_to_yuv(d_r, d_g, d_b, w, h, Y, U, V, st);
    if (!rc) rc = launch_vp8_encode(f, st);
    if (rc) { err = std::string("vp8 kernels: ") + cudaGetErrorString((cudaError_t)rc); return false; }
    CUW(cudaMemcpyAsync(h_out, d_levels, lv_bytes, cudaMemcpyDeviceToHost, st));
    CUW(cudaMemcpyAsync(h_out + lv_bytes, d_modes, md_bytes, cudaMemcpyDeviceToHost, st));
    const auto t0 = std::chrono::steady_clock::now();
    CUW(stream_wait(st));
    const auto t1 = std::chrono::steady_clock::now();
    struct Lap { WebpDevice *d; std::chrono::steady_clock::time_point a, b; ~Lap() { d->last_wait_ms = std::chrono::duration<double, std::milli>(b - a).count(); d->last_code_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - b).count(); } } lap{this, t0, t1};
    if (levels_out) memcpy(levels_out, h_out, lv_bytes);
    if (modes_out) memcpy(modes_out, h_out + lv_bytes, md_bytes);
    if (!vp8_write_file(w, h, qi, reinterpret_cast<const int16_t *>(h_out), h_out + lv_bytes, out)) { err = "VP8 frame cannot be framed (first partition too large)"; return false; }
    return true;
}

bool WebpDevice::encode_host_rgb(const uint8_t *rgb, int w, int h, int quality, void *stream_, std::vector<uint8_t> &out, std::string &err,
                                 int16_t *levels_out, uint8_t *modes_out)
{
    cudaStream_t st = (cudaStream_t)stream_;
    const size_t n = (size_t)w * h;
    if (!groww(h_rgb, cap_hrgb, 3 * n, true, err) || !groww(d_rgb, cap_rgb, 3 * n, false, err)) return false;
    memcpy(h_rgb, rgb, 3 * n);
    CUW(cudaMemcpyAsync(d_rgb, h_rgb, 3 * n, cudaMemcpyHostToDevice, st));
    return encode_planes(d_rgb, d_rgb + n, d_rgb + 2 * n, w, h, quality, st, out, err, levels_out, modes_out);
}

} // namespace b200

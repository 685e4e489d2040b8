// jpeg_kernels.h -- device work descriptors and launchers for the JPEG transform kernels
// (K1 dequant+IDCT, K2 chroma upsample, K4 downsample, K5 FDCT+quantise+zigzag and their fusions;
// SURVEY.md §8a row a6).  Plain C++ declarations so host translation units need no CUDA headers.
#pragma once
#include <cstdint>
#include <cstddef>

namespace b200 {

// Output quantiser of one component, in ZIGZAG order.  floor(t / d) for t < 2^19 is computed as
// umulhi(t, m) >> sh with m = ceil(2^s / d) << (32 - min(s,32)), s = 19 + ceil(log2 d): exact (jpeg_kernels.cu).
struct QuantDev {
    uint32_t m[64];
    uint32_t half_sh[64];   // (d >> 1) | (sh << 24), d = quantval << 3
};

// One unit of device work: one component of one image.
struct CompWork {
    const int16_t *cin;     // input coefficients  [bh_in*bw_in][64] zigzag, quantised
    int16_t *cout;          // output coefficients [bh_out*bw_out][64] zigzag
    uint8_t *plane;         // component-resolution samples, (bh_in*8) rows x pstride
    uint8_t *full;          // full-resolution plane W x H (generic path only), stride = fstride
    uint8_t *dplane;        // downsampled padded plane (rbh_out*8) x (rbw_out*8) (generic path only)
    const uint16_t *dq;     // 64 dequantisation multipliers, zigzag order
    const QuantDev *q;      // output quantiser
    int32_t bw_in, bh_in;   // allocated blocks of the input component
    int32_t rbw_in, rbh_in; // real blocks of the input component
    int32_t cw, ch;         // real sample dims of the input component
    int32_t bw_out, bh_out, rbw_out, rbh_out;
    int32_t W, H;           // image dims
    int32_t pstride, fstride;
    int32_t up_hx, up_vx;   // decoder upsampling ratio  (hmax/hs, vmax/vs of the input)
    int32_t dn_hx, dn_vx;   // encoder downsampling ratio (hmax/hs, vmax/vs of the output)
};

inline int work_tiles(int rbw, int rbh) { return ((rbw + 31) / 32) * rbh; }   // row-aligned tiles of 32 blocks
void make_quant_dev(const uint16_t qt_zigzag[64], QuantDev *out);

// Launchers.  `work` is a DEVICE array of n descriptors; max_tiles = max over the n items of
// work_tiles(real blocks across, down) for the grid the kernel iterates (sizes grid.x).  All asynchronous on `stream` (cudaStream_t).
// Return cudaError_t as int.
int launch_fused_same(const CompWork *work, int n, int max_tiles, void *stream);       // IDCT -> FDCT+quant, same geometry
int launch_idct_plane(const CompWork *work, int n, int max_tiles, void *stream);       // IDCT -> u8 plane
int launch_chroma420_refdct(const CompWork *work, int n, int max_tiles, void *stream); // h2v2 fancy up o h2v2 box down o FDCT+quant
int launch_upsample(const CompWork *work, int n, int max_w, int max_h, void *stream);   // plane -> full
int launch_downsample(const CompWork *work, int n, int max_w, int max_h, void *stream); // full -> dplane
int launch_fdct_plane(const CompWork *work, int n, int max_tiles, void *stream);       // dplane -> coefficients
int launch_memset_warm(void *p, size_t n, void *stream);

} // namespace b200

// jpeg_device.h -- device side of the JPEG path: per-GPU slot pools (stream + pinned staging + HBM buffers),
// work-list construction for the transform kernels, and the device-resident megabatch.
#pragma once
#include <cstdint>
#include <cstddef>
#include <string>
#include <vector>
#include "jpeg_host.h"
#include "jpeg_kernels.h"
#include "jpeg_gpudec.h"
#include "jpeg_gpuenc.h"

namespace b200 {

enum CompPath { PATH_FUSED = 0, PATH_C420 = 1, PATH_GENERIC = 2 };

// Per-image HBM footprint and per-component routing for an (input geometry, output geometry) pair.
struct ImagePlan {
    int path[4] = {0, 0, 0, 0};
    size_t in_bytes = 0, out_bytes = 0;
    size_t plane_off[4] = {0}, full_off[4] = {0}, dplane_off[4] = {0};
    size_t plane_bytes = 0, full_bytes = 0, dplane_bytes = 0;
    size_t scratch_bytes() const { return plane_bytes + full_bytes + dplane_bytes; }
};
bool plan_image(const JpegGeom &gin, const JpegGeom &gout, ImagePlan &plan, std::string &err);

// Work lists for one launch group (any number of images).
struct WorkLists {
    std::vector<CompWork> fused, idct, c420, up, down, fdct;
    int max_fused = 0, max_idct = 0, max_c420 = 0, max_fdct = 0, max_up_w = 0, max_up_h = 0, max_dn_w = 0, max_dn_h = 0;
    size_t total() const { return fused.size() + idct.size() + c420.size() + up.size() + down.size() + fdct.size(); }
    void clear() { fused.clear(); idct.clear(); c420.clear(); up.clear(); down.clear(); fdct.clear(); max_fused = max_idct = max_c420 = max_fdct = max_up_w = max_up_h = max_dn_w = max_dn_h = 0; }
};
// Append one image's work.  d_dq: device uint16[4][64] (per input component, zigzag); d_q: device QuantDev[4] (per output slot).
void append_image_work(const JpegGeom &gin, const JpegGeom &gout, const ImagePlan &plan,
                       const int16_t *d_in, int16_t *d_out, uint8_t *d_scratch,
                       const uint16_t *d_dq, const QuantDev *d_q, WorkLists &wl);
// Copy lists into `h_work` (contiguous, order fused|idct|c420|up|down|fdct); returns count.
size_t flatten_work(const WorkLists &wl, CompWork *h_work);
// Launch every non-empty list; d_work is the device copy of the flattened array.  which: 0 all, 1 fused, 2 idct, 3 c420(+generic tail).
int launch_work(const WorkLists &wl, const CompWork *d_work, void *stream, int which, int *launches);

// ---- device runtime ------------------------------------------------------------------------------------------
struct Slot {
    int dev = 0;
    void *stream = nullptr;
    int16_t *h_in = nullptr, *h_out = nullptr; size_t h_in_cap = 0, h_out_cap = 0;       // pinned
    int16_t *d_in = nullptr, *d_out = nullptr; size_t d_in_cap = 0, d_out_cap = 0;
    uint8_t *d_scratch = nullptr; size_t d_scratch_cap = 0;
    uint8_t *h_par = nullptr, *d_par = nullptr; size_t par_cap = 0, d_par_cap = 0;      // parameter block
    class GpuEncoder *enc = nullptr;                                                     // device entropy encoder (lazy)
    class GpuDecoder *dec = nullptr;                                                     // device entropy decoder (lazy)
    struct PngDevice *png = nullptr;                                                     // lossless PNG state (lazy, png_device.cu)
    struct WebpDevice *webp = nullptr;                                                   // WebP / VP8 state (lazy, webp_device.cu)
    // megabatch path: transform work lists of the current megabatch, and the captured launch sequence (two CUDA graphs, see
    // slot_run_group) with the signature it was captured for
    WorkLists group_wl; size_t group_par_bytes = 0, group_work_off = 0;
    void *graph_front = nullptr, *graph_back = nullptr; unsigned long long graph_sig = 0; bool graphs_broken = false;
    bool ensure(size_t in_bytes, size_t out_bytes, size_t scratch_bytes, size_t par_bytes, std::string &err);
    bool ensure_device(size_t in_bytes, size_t out_bytes, size_t scratch_bytes, size_t par_bytes, std::string &err);
};

int  runtime_init(int n_gpus, int only_device, std::string &err);   // returns device count (>0) or 0 with err
void runtime_shutdown();
int  runtime_device_count();
Slot *slot_acquire(int prefer_dev, std::string &err);               // blocks while all slots of the device are busy
void slot_release(Slot *s);
int  runtime_next_device();                                         // round-robin shard assignment
long long runtime_device_jobs(int dev_index);                       // slot acquisitions on that device so far
int  runtime_device_ordinal(int dev_index);                         // CUDA ordinal of the library's device number dev_index

// Run the transform for ONE image whose input coefficients already sit in s->h_in; result lands in s->h_out.
bool slot_transform(Slot *s, const JpegGeom &gin, const JpegGeom &gout, std::string &err, bool download = true, bool upload = true);
// D2H of the output coefficients left in HBM by a download=false transform (host-encoder fallback)
bool slot_download_coefs(Slot *s, size_t out_bytes, std::string &err);
// Entropy-decode a baseline single-scan file on the device into s->d_in (0 ok, 1 not converged -> host decode, 2 failed)
int slot_gpu_decode(Slot *s, const JpegReader &rd, const JpegReader::DeviceScan &ds, std::string &err);
// ---- megabatch (K same-shaped images per launch sequence; used by b200_compress_batch) ---------------------------------
struct GroupLayout { int K = 0; size_t in_stride = 0, out_stride = 0, scratch_stride = 0; };
bool slot_group_layout(Slot *s, const JpegGeom &gin, const JpegGeom &gout, int K, GroupLayout &L, std::string &err);   // sizes + ensure()
bool slot_decode_group(Slot *s, std::vector<GpuDecoder::Item> &items, std::string &err);      // items[k].d_coefs = d_in + k * in_stride
// decode -> (transform) -> encode of one megabatch enqueued back to back, one idle host wait at the end; results in s->enc->results,
// items[k].result says which images the device decoder settled
bool slot_run_group(Slot *s, std::vector<GpuDecoder::Item> &items, const JpegGeom *const *gins, const JpegGeom &gout, const GroupLayout &L, bool progressive,
                    bool lossless, std::string &err);
bool slot_transform_group(Slot *s, const JpegGeom *const *gins, const JpegGeom &gout, const GroupLayout &L, std::string &err);
// results in s->enc->results; from_input = encode the coefficients in d_in (lossless transcode) instead of d_out
bool slot_encode_group(Slot *s, const JpegGeom &gout, bool progressive, const GroupLayout &L, std::string &err, bool from_input = false);
// H2D of s->h_out into s->d_out (entry point that encodes caller-supplied coefficients on the device)
bool slot_upload_out_coefs(Slot *s, size_t bytes, std::string &err);
// Entropy-code the output coefficients sitting in s->d_out on the device; result in s->enc->results
bool slot_gpu_encode(Slot *s, const JpegGeom &gout, bool progressive, std::string &err, bool from_input = false);
// the same without fetching the stuffed scans (results carry lengths only); slot_gpu_fetch() brings them over afterwards
bool slot_gpu_encode_sizes(Slot *s, const JpegGeom &gout, bool progressive, std::string &err);
bool slot_gpu_fetch(Slot *s, std::string &err);
// Resize path (CSParameters.width/height): gout carries the TARGET dimensions; decode -> RGB -> Lanczos3 -> YCbCr -> encode.
// rgb_out != nullptr: stop after the resize and hand back the three device planes (R, G, B of the TARGET size, pitch = target
// width; a greyscale source returns its single plane three times) -- the front end of the format-conversion paths.
// host_rgb != nullptr: the source is not a JPEG -- planar samples [ncomp][H][W] (RGB, or one grey plane) replace the decode
// front end; gin then only carries width / height / ncomp with 1x1 sampling.
bool slot_transform_resized(Slot *s, const JpegGeom &gin, const JpegGeom &gout, std::string &err, bool download = true, bool upload = true,
                            uint8_t **rgb_out = nullptr, const uint8_t *host_rgb = nullptr);
// D2H of nplanes device planes of n bytes each (rgb_out of slot_transform_resized) into one host buffer, synchronised
bool slot_fetch_planes(Slot *s, uint8_t *const *d_planes, int nplanes, size_t n, uint8_t *host, std::string &err);
// Same front end, but stop after IDCT + upsample and copy planar full-res samples into `planes` (host).
bool slot_decode_planes(Slot *s, const JpegGeom &gin, uint8_t *planes, std::string &err);

// ---- device-resident megabatch ---------------------------------------------------------------------------------
struct JpegBatch {
    int dev = 0, n = 0;
    JpegGeom gin, gout; ImagePlan plan;
    int16_t *d_in = nullptr, *d_out = nullptr; uint8_t *d_scratch = nullptr;
    uint8_t *d_par = nullptr;
    WorkLists wl; const CompWork *d_work = nullptr;
    void *stream = nullptr; void *ev0 = nullptr, *ev1 = nullptr;
};
JpegBatch *batch_create(const JpegGeom &gin, const JpegGeom &gout, int n, std::string &err);
bool batch_upload(JpegBatch *b, int idx, const int16_t *coefs, std::string &err);
bool batch_run(JpegBatch *b, void *stream, int which, int *launches, std::string &err);
bool batch_download(JpegBatch *b, int idx, int16_t *coefs, std::string &err);
bool batch_time(JpegBatch *b, int which, int iters, float *ms, std::string &err);
void batch_destroy(JpegBatch *b);

} // namespace b200

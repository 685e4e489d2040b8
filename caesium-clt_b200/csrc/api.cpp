// api.cpp -- the C-ABI of include/b200_caesium.h: format dispatch, parameter mapping and error mapping that
// libcaesium's lib.rs performs behind caesium::{compress,convert,compress_to_size}_in_memory
// (call sites /root/reference/src/compressor.rs:287-306).  No CPU codec fallback exists anywhere below.
#include "../../include/b200_caesium.h"
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <exception>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "jpeg_device.h"
#include "jpeg_host.h"
#include "jpeg_gpuenc.h"
#include "resize_kernels.h"
#include "png_host.h"
#include "png_device.h"
#include "vp8_host.h"
#include "webp_device.h"
#include "vp8_decode.h"
#include "vp8l_alpha.h"
#include "jpeg_pipe.h"
#include "topology.h"
#include "launch_timer.h"

using namespace b200;

namespace {

b200_status ok_status() { b200_status s; s.code = B200_OK; s.message = nullptr; return s; }
b200_status make_status(int code, const std::string &msg)
{   // CaesiumError Display: "{message} [{code}]"
    b200_status s; s.code = code;
    std::string m = msg + " [" + std::to_string(code) + "]";
    s.message = (char *)malloc(m.size() + 1);
    if (s.message) memcpy(s.message, m.c_str(), m.size() + 1);
    return s;
}

// a header the reader refused: RGB-coded sources are "recognised but not on this path" (code 3), everything else is corrupt input
b200_status header_status(const std::string &err) { return make_status(err.compare(0, 9, "RGB-coded") == 0 ? B200_ERR_UNSUPPORTED : B200_ERR_CORRUPT_INPUT, err); }

std::once_flag g_once;
std::string g_init_err;
int g_forced_device = -1, g_forced_ngpus = 0;
std::atomic<int> g_entropy_mode{-1};     // -1 unset (env B200_ENTROPY); bit 0 = device entropy encoder, bit 1 = device entropy decoder (default 3)

bool ensure_runtime(std::string &err)
{
    std::call_once(g_once, [] { runtime_init(g_forced_ngpus, g_forced_device, g_init_err); });
    if (runtime_device_count() <= 0) { err = g_init_err.empty() ? "no CUDA device available; this build has no CPU fallback" : g_init_err; return false; }
    return true;
}

void layout_from_geom(const JpegGeom &g, b200_jpeg_layout *l)
{
    memset(l, 0, sizeof(*l));
    l->width = g.width; l->height = g.height; l->ncomp = g.ncomp; l->progressive = g.progressive;
    for (int c = 0; c < g.ncomp; c++) {
        l->hs[c] = g.hs[c]; l->vs[c] = g.vs[c]; l->bw[c] = g.bw[c]; l->bh[c] = g.bh[c]; l->rbw[c] = g.rbw[c]; l->rbh[c] = g.rbh[c];
        l->comp_offset[c] = g.comp_offset[c];
        memcpy(l->qt[c], g.qt[g.tq[c]], 128);
    }
    l->total_coefs = g.total_coefs;
}

bool geom_from_layout(const b200_jpeg_layout *l, JpegGeom &g, std::string &err)
{
    g = JpegGeom();
    if (!l || l->width <= 0 || l->height <= 0 || (l->ncomp != 1 && l->ncomp != 3)) { err = "invalid JPEG layout"; return false; }
    g.width = l->width; g.height = l->height; g.ncomp = l->ncomp; g.progressive = l->progressive != 0;
    int nslots = 0;
    for (int c = 0; c < l->ncomp; c++) {
        if (l->hs[c] < 1 || l->hs[c] > 4 || l->vs[c] < 1 || l->vs[c] > 4) { err = "invalid sampling factors"; return false; }
        g.cid[c] = c + 1; g.hs[c] = l->hs[c]; g.vs[c] = l->vs[c];
        int slot = -1;
        for (int t = 0; t < nslots; t++) if (!memcmp(g.qt[t], l->qt[c], 128)) slot = t;
        if (c == 1 && slot == 0 && nslots == 1) slot = -1;          // keep luma / chroma in separate slots like jpeg_set_defaults
        if (slot < 0) { slot = nslots++; memcpy(g.qt[slot], l->qt[c], 128); g.qt_present[slot] = true; }
        g.tq[c] = slot;
    }
    g.finalize();
    for (int c = 0; c < l->ncomp; c++) if (l->bw[c] != g.bw[c] || l->bh[c] != g.bh[c] || l->comp_offset[c] != g.comp_offset[c]) { err = "layout block counts do not match its dimensions"; return false; }
    return true;
}

int usable_cores()
{   // cgroup v2 quota if any (the GPU boxes expose 128 logical CPUs but cap the container), else hardware_concurrency
    unsigned hc = std::thread::hardware_concurrency(); if (!hc) hc = 1;
    std::ifstream f("/sys/fs/cgroup/cpu.max");
    std::string a; long long period = 0;
    if (f && (f >> a >> period) && a != "max" && period > 0) {
        long long q = atoll(a.c_str());
        if (q > 0) { unsigned n = (unsigned)((q + period - 1) / period); if (n >= 1 && n < hc) hc = n; }
    }
    return (int)hc;
}

// B200_TRACE=1: wall-clock per stage of the per-image call, summed over all images, printed by b200_shutdown()
std::atomic<long long> g_stage_ns[8];
std::atomic<long long> g_stage_n{0};
const bool g_trace = getenv("B200_TRACE") != nullptr;
struct StageTimer {
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(int i)
    {
        if (!g_trace) return;
        auto n = std::chrono::steady_clock::now();
        const long long ns = std::chrono::duration_cast<std::chrono::nanoseconds>(n - t).count();
        g_stage_ns[i] += ns; t = n;
        if (i == 5) g_stage_n++;
        static const bool verbose = getenv("B200_TRACE") && atoi(getenv("B200_TRACE")) >= 2;
        if (verbose) fprintf(stderr, "[b200 trace] stage %d: %.3f ms\n", i, ns / 1e6);
    }
};
void print_trace()
{
    if (!g_trace || !g_stage_n.load()) return;
    static const char *names[] = {"parse+scan-for-markers", "device entropy decode (incl. syncs)", "host entropy decode", "transform launch", "device entropy encode (incl. syncs)", "assemble file"};
    fprintf(stderr, "[b200 trace] %lld images; mean ms per image and stage:\n", g_stage_n.load());
    for (int i = 0; i < 6; i++) fprintf(stderr, "[b200 trace]   %-38s %8.3f\n", names[i], g_stage_ns[i].load() / 1e6 / (double)g_stage_n.load());
}

// ---- JPEG through the device ---------------------------------------------------------------------------------
b200_status jpeg_compress(const uint8_t *in, size_t in_len, const b200_params *p, int prefer_dev, std::vector<uint8_t> &out)
{
    std::string err;
    JpegReader rd(in, in_len);
    if (!rd.read_header(err)) return header_status(err);
    const JpegGeom &gin = rd.geom();
    JpegWriteOptions wo; wo.progressive = p->jpeg_progressive != 0; wo.keep_metadata = p->keep_metadata != 0; wo.preserve_icc = p->jpeg_preserve_icc != 0;
    if (p->jpeg_optimize) {
        // libcaesium jpeg::lossless: coefficient-domain transcode.  Baseline single-scan inputs are entropy-decoded and
        // re-encoded (optimal tables, progressive script) by the device coders; the coefficients never leave HBM.
        // Everything else (progressive input, no device) is transcoded on the calling thread -- there is no arithmetic
        // on this path, only entropy coding.
        std::string derr;
        JpegReader::DeviceScan ds;
        if (g_entropy_mode.load() < 0) {
            const char *e = getenv("B200_ENTROPY");
            g_entropy_mode.store(!e ? 3 : !strcmp(e, "host") ? 0 : !strcmp(e, "gpuenc") ? 1 : !strcmp(e, "gpudec") ? 2 : 3);
        }
        if (g_entropy_mode.load() == 3 && rd.device_decodable(ds) && ensure_runtime(derr)) {
            if (Slot *s = slot_acquire(prefer_dev < 0 ? runtime_next_device() : prefer_dev, derr)) {
                bool done = false;
                wo.copy_jfif = true;
                if (s->ensure((size_t)gin.total_coefs * 2, 0, 0, 1 << 14, derr) && slot_gpu_decode(s, rd, ds, derr) == 0 &&
                    slot_gpu_encode(s, gin, wo.progressive, derr, true))
                    done = jpeg_assemble(gin, wo, &rd.meta(), s->enc->results.data(), (int)s->enc->results.size(), out, derr);
                slot_release(s);
                if (done) return ok_status();
                out.clear();
            }
        }
        std::vector<int16_t> coefs((size_t)gin.total_coefs);
        if (!rd.decode(coefs.data(), err)) return make_status(B200_ERR_CORRUPT_INPUT, err);
        jpeg_fill_dummy_blocks(gin, coefs.data());
        wo.copy_jfif = true;
        if (!jpeg_write(gin, coefs.data(), wo, &rd.meta(), out, err)) return make_status(B200_ERR_INVALID_ARGUMENT, err);
        return ok_status();
    }
    if (!ensure_runtime(err)) return make_status(B200_ERR_NO_DEVICE, err);
    if (g_entropy_mode.load() < 0) {
        const char *e = getenv("B200_ENTROPY");
        g_entropy_mode.store(!e ? 3 : !strcmp(e, "host") ? 0 : !strcmp(e, "gpuenc") ? 1 : !strcmp(e, "gpudec") ? 2 : 3);
    }
    JpegGeom gout;
    if (!jpeg_output_geom(gin, (int)p->jpeg_quality, (int)p->jpeg_chroma_subsampling, gout, err)) return make_status(B200_ERR_INVALID_ARGUMENT, err);
    const bool resize = p->width || p->height;
    if (resize) {   // libcaesium resize::resize_image -> compute_dimensions
        uint32_t nw = 0, nh = 0;
        compute_resize_dimensions((uint32_t)gin.width, (uint32_t)gin.height, p->width, p->height, nw, nh);
        if (nw == 0 || nh == 0 || nw > 65535 || nh > 65535) return make_status(B200_ERR_INVALID_ARGUMENT, "invalid target dimensions");
        gout.width = (int)nw; gout.height = (int)nh; gout.finalize();
    }
    ImagePlan plan;
    if (!resize && !plan_image(gin, gout, plan, err)) return make_status(B200_ERR_UNSUPPORTED, err);
    if (resize) { plan.in_bytes = (size_t)gin.total_coefs * 2; plan.out_bytes = (size_t)gout.total_coefs * 2; }
    Slot *s = slot_acquire(prefer_dev < 0 ? runtime_next_device() : prefer_dev, err);
    if (!s) return make_status(B200_ERR_CUDA, err);
    b200_status st = ok_status();
    do {
        if (!s->ensure(plan.in_bytes, plan.out_bytes, plan.scratch_bytes(), 1 << 14, err)) { st = make_status(B200_ERR_OUT_OF_MEMORY, err); break; }
        const int mode = g_entropy_mode.load();
        const bool gpu_entropy = (mode & 1) != 0;
        // Entropy DECODE on the device for baseline single-scan files (jpeg_gpudec.cu): the scan's bytes go up, the
        // coefficients are born in HBM.  Progressive / multi-scan / restart-interval files, and the rare stream whose
        // parallel decode does not settle, are Huffman-decoded here on the calling thread instead.
        bool on_device = false;
        JpegReader::DeviceScan ds;
        StageTimer tm;
        if ((mode & 2) && rd.device_decodable(ds)) {
            tm.lap(0);
            const int r = slot_gpu_decode(s, rd, ds, err);
            tm.lap(1);
            if (r == 0) on_device = true;
            else if (r != 1) { st = make_status(B200_ERR_CUDA, err); break; }
        }
        if (!on_device && !rd.decode(s->h_in, err)) { st = make_status(B200_ERR_CORRUPT_INPUT, err); break; }
        tm.lap(2);
        if (!(resize ? slot_transform_resized(s, gin, gout, err, !gpu_entropy, !on_device) : slot_transform(s, gin, gout, err, !gpu_entropy, !on_device))) { st = make_status(B200_ERR_CUDA, err); break; }
        tm.lap(3);
        if (gpu_entropy) {
            // Huffman statistics, table construction, bit packing and 0xFF stuffing on the device (jpeg_gpuenc.cu); the
            // host only frames the scans.  A scan that outgrows its device buffer falls back to the host ENCODER
            // (still the same coefficients from the CUDA transform).
            if (slot_gpu_encode(s, gout, wo.progressive, err)) {
                tm.lap(4);
                if (!jpeg_assemble(gout, wo, &rd.meta(), s->enc->results.data(), (int)s->enc->results.size(), out, err)) st = make_status(B200_ERR_INVALID_ARGUMENT, err);
                tm.lap(5);
                break;
            }
            if (!s->enc || !s->enc->overflow) { st = make_status(B200_ERR_CUDA, err); break; }
            if (!slot_download_coefs(s, plan.out_bytes, err)) { st = make_status(B200_ERR_CUDA, err); break; }
        }
        jpeg_fill_dummy_blocks(gout, s->h_out);
        if (!jpeg_write(gout, s->h_out, wo, &rd.meta(), out, err)) { st = make_status(B200_ERR_INVALID_ARGUMENT, err); break; }
    } while (0);
    slot_release(s);
    return st;
}

b200_status give(std::vector<uint8_t> &v, uint8_t **out, size_t *out_len)
{
    *out = (uint8_t *)malloc(v.size() ? v.size() : 1);
    if (!*out) return make_status(B200_ERR_OUT_OF_MEMORY, "out of memory");
    memcpy(*out, v.data(), v.size()); *out_len = v.size();
    return ok_status();
}

// One megabatch: the images idx[] that are baseline single-scan JPEGs of one shape are decoded, transformed and encoded by
// ONE sequence of kernel launches on one slot.  done[k] = 1 for every image this function finished (successfully or with
// a final error in status[]); the caller runs the others through the per-image path.
void jpeg_compress_group(const uint8_t *const *in, const size_t *in_len, const std::vector<int> &idx, const b200_params *p, int dev,
                         uint8_t **out, size_t *out_len, b200_status *status, std::vector<char> &done)
{
    const int M = (int)idx.size();
    StageTimer tm;
    std::vector<std::unique_ptr<JpegReader>> rd((size_t)M);
    std::vector<JpegReader::DeviceScan> ds((size_t)M);
    std::vector<int> members;                       // positions k (into idx) that join the group
    std::string err;
    for (int k = 0; k < M; k++) {
        const int i = idx[k];
        if (b200_sniff_format(in[i], in_len[i]) != B200_FMT_JPEG) continue;
        rd[k].reset(new JpegReader(in[i], in_len[i]));
        if (!rd[k]->read_header(err) || !rd[k]->device_decodable(ds[k], true)) continue;      // the entropy-coded segment is walked on the device
        if (!members.empty()) {
            const JpegGeom &a = rd[members[0]]->geom(), &b = rd[k]->geom();
            bool same = a.width == b.width && a.height == b.height && a.ncomp == b.ncomp;
            for (int c = 0; same && c < a.ncomp; c++) same = a.hs[c] == b.hs[c] && a.vs[c] == b.vs[c];
            if (!same) continue;
        }
        members.push_back(k);
    }
    if (members.size() < 2) return;
    const JpegGeom &gin0 = rd[members[0]]->geom();
    const bool lossless = p->jpeg_optimize != 0;          // jpeg::lossless: decode -> encode, no transform, per-image tables kept
    JpegGeom gout;
    if (lossless) gout = gin0;
    else if (!jpeg_output_geom(gin0, (int)p->jpeg_quality, (int)p->jpeg_chroma_subsampling, gout, err)) return;
    Slot *s = slot_acquire(dev, err);
    if (!s) return;
    const int Kg = (int)members.size();
    bool ok = false;
    do {
        GroupLayout L;
        if (!slot_group_layout(s, gin0, gout, Kg, L, err)) break;
        std::vector<GpuDecoder::Item> items((size_t)Kg);
        std::vector<const JpegGeom *> gins((size_t)Kg);
        for (int m = 0; m < Kg; m++) {
            const int k = members[m];
            items[m].rd = rd[k].get(); items[m].ds = &ds[k]; items[m].result = GpuDecoder::FAILED;
            items[m].d_coefs = reinterpret_cast<int16_t *>(reinterpret_cast<uint8_t *>(s->d_in) + L.in_stride * m);
            gins[m] = &rd[k]->geom();
        }
        tm.lap(0);
        JpegWriteOptions wo; wo.progressive = p->jpeg_progressive != 0; wo.keep_metadata = p->keep_metadata != 0; wo.preserve_icc = p->jpeg_preserve_icc != 0;
        wo.copy_jfif = lossless;
        if (!slot_run_group(s, items, gins.data(), gout, L, wo.progressive, lossless, err)) break;
        tm.lap(4);
        const int spi = s->enc->plan.scans_per_image;
        for (int m = 0; m < Kg; m++) {
            const int k = members[m], i = idx[k];
            if (items[m].result != GpuDecoder::OK) continue;          // not converged: the per-image path decodes it on the host
            out[i] = nullptr; out_len[i] = 0;
            if (!jpeg_assemble_malloc(lossless ? rd[k]->geom() : gout, wo, &rd[k]->meta(), s->enc->results.data() + (size_t)m * spi, spi, &out[i], &out_len[i], err)) status[i] = make_status(B200_ERR_INVALID_ARGUMENT, err);
            else status[i] = ok_status();
            done[k] = 1;
        }
        tm.lap(5);
        ok = true;
    } while (0);
    (void)ok;
    slot_release(s);
}

// ---- PNG (lossless) through the device ---------------------------------------------------------------------------
// libcaesium png::compress: optimize == true -> png::lossless (oxipng, level = png.optimization_level); otherwise the lossy
// palette quantiser (imagequant), which is outside this path.  Resizing a PNG goes through the image crate's decoder and is
// likewise left to the reference.
b200_status png_compress(const uint8_t *in, size_t in_len, const b200_params *p, int prefer_dev, std::vector<uint8_t> &out)
{
    if (!p->png_optimize) return make_status(B200_ERR_UNSUPPORTED, "lossy PNG (imagequant) is outside the GPU path (route to caesium::compress_in_memory)");
    if (p->width || p->height) return make_status(B200_ERR_UNSUPPORTED, "PNG resize is outside the GPU path (route to caesium::compress_in_memory)");
    std::string err;
    PngInfo info; PngIdat idat;
    static const bool verbose = getenv("B200_TRACE") && atoi(getenv("B200_TRACE")) >= 2;
    const auto t0 = std::chrono::steady_clock::now();
    if (!png_parse_chunks(in, in_len, p->keep_metadata != 0, info, idat, err)) return make_status(err.find("interlace") != std::string::npos ? B200_ERR_UNSUPPORTED : B200_ERR_CORRUPT_INPUT, err);
    if (!ensure_runtime(err)) return make_status(B200_ERR_NO_DEVICE, err);
    Slot *s = slot_acquire(prefer_dev < 0 ? runtime_next_device() : prefer_dev, err);
    if (!s) return make_status(B200_ERR_CUDA, err);
    if (!s->png) s->png = new PngDevice();
    // the IDAT stream is inflated straight into the slot's pinned staging buffer; from there on everything is device work
    // (un-filter, checksum, reductions, filter trials, LZ77, DEFLATE coding) until the finished zlib stream comes back
    const size_t nin = (info.row_bytes + 1) * (size_t)info.height;
    size_t cap = 0, got = 0; uint32_t stored_adler = 0;
    uint8_t *buf = s->png->input_buffer(nin, cap, err);
    if (!buf) { slot_release(s); return make_status(B200_ERR_OUT_OF_MEMORY, err); }
    if (!zlib_inflate_to(idat.p, idat.n, buf, cap, nin, &got, &stored_adler, err)) { slot_release(s); return make_status(B200_ERR_CORRUPT_INPUT, err); }
    if (got < nin) { slot_release(s); return make_status(B200_ERR_CORRUPT_INPUT, "IDAT too short"); }
    const auto t1 = std::chrono::steady_clock::now();
    std::vector<uint8_t> z;
    int level = (int)p->png_optimization_level; if (level > 6) level = 6;
    const bool ok = s->png->compress_filtered(info, got, stored_adler, level, s->stream, z, nullptr, err);
    const bool corrupt = s->png->corrupt;
    const double deflate_ms = s->png->last_deflate_ms;
    slot_release(s);
    if (!ok) return make_status(corrupt ? B200_ERR_CORRUPT_INPUT : B200_ERR_CUDA, err);
    const auto t3 = std::chrono::steady_clock::now();
    png_write(info, z, out);
    if (verbose) {
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "[b200 trace] png %ux%u: parse + inflate %.1f ms, device (un-filter, filter trials, LZ77, DEFLATE coding; host Huffman %.1f) %.1f ms, container %.1f ms\n",
                info.width, info.height, ms(t0, t1), deflate_ms, ms(t1, t3), ms(t3, std::chrono::steady_clock::now()));
    }
    return ok_status();
}

// ---- conversion to WebP (lossy VP8) ----------------------------------------------------------------------------------
// libcaesium convert: decode -> (resize) -> webp::compress at parameters.webp.quality.  JPEG sources are decoded on the
// device (entropy decode, IDCT, upsample, YCbCr -> RGB, Lanczos3 when width/height are set) and never leave HBM before K8.
b200_status jpeg_to_webp(const uint8_t *in, size_t in_len, const b200_params *p, int prefer_dev, std::vector<uint8_t> &out)
{
    std::string err;
    JpegReader rd(in, in_len);
    if (!rd.read_header(err)) return header_status(err);
    const JpegGeom &gin = rd.geom();
    if (!ensure_runtime(err)) return make_status(B200_ERR_NO_DEVICE, err);
    uint32_t nw = (uint32_t)gin.width, nh = (uint32_t)gin.height;
    if (p->width || p->height) compute_resize_dimensions((uint32_t)gin.width, (uint32_t)gin.height, p->width, p->height, nw, nh);
    if (nw == 0 || nh == 0 || nw > 16383 || nh > 16383) return make_status(B200_ERR_INVALID_ARGUMENT, "invalid target dimensions for WebP");
    JpegGeom gout = gin; gout.width = (int)nw; gout.height = (int)nh; gout.finalize();
    if (g_entropy_mode.load() < 0) {
        const char *e = getenv("B200_ENTROPY");
        g_entropy_mode.store(!e ? 3 : !strcmp(e, "host") ? 0 : !strcmp(e, "gpuenc") ? 1 : !strcmp(e, "gpudec") ? 2 : 3);
    }
    Slot *s = slot_acquire(prefer_dev < 0 ? runtime_next_device() : prefer_dev, err);
    if (!s) return make_status(B200_ERR_CUDA, err);
    b200_status st = ok_status();
    static const bool verbose = getenv("B200_TRACE") && atoi(getenv("B200_TRACE")) >= 2;
    const auto t0 = std::chrono::steady_clock::now();
    auto t1 = t0, t2 = t0;
    do {
        if (!s->ensure((size_t)gin.total_coefs * 2, (size_t)gout.total_coefs * 2, 0, 1 << 14, err)) { st = make_status(B200_ERR_OUT_OF_MEMORY, err); break; }
        bool on_device = false;
        JpegReader::DeviceScan ds;
        if ((g_entropy_mode.load() & 2) && rd.device_decodable(ds)) {
            const int r = slot_gpu_decode(s, rd, ds, err);
            if (r == 0) on_device = true; else if (r != 1) { st = make_status(B200_ERR_CUDA, err); break; }
        }
        if (!on_device && !rd.decode(s->h_in, err)) { st = make_status(B200_ERR_CORRUPT_INPUT, err); break; }
        t1 = std::chrono::steady_clock::now();
        uint8_t *rgb[3] = {nullptr, nullptr, nullptr};
        if (!slot_transform_resized(s, gin, gout, err, false, !on_device, rgb)) { st = make_status(B200_ERR_CUDA, err); break; }
        t2 = std::chrono::steady_clock::now();
        if (!s->webp) s->webp = new WebpDevice();
        if (!s->webp->encode_planes(rgb[0], rgb[1], rgb[2], (int)nw, (int)nh, (int)p->webp_quality, s->stream, out, err)) { st = make_status(B200_ERR_CUDA, err); break; }
        if (verbose) {
            auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
            fprintf(stderr, "[b200 trace] jpeg %dx%d -> webp %ux%u: segment walk + entropy decode (device %d) %.1f ms, transform + resize launch %.1f ms, VP8 (wait for the device %.1f ms, boolean coder %.1f ms) %.1f ms\n",
                    gin.width, gin.height, nw, nh, (int)on_device, ms(t0, t1), ms(t1, t2), s->webp->last_wait_ms, s->webp->last_code_ms, ms(t2, std::chrono::steady_clock::now()));
        }
    } while (0);
    slot_release(s);
    return st;
}

// JPEG -> PNG (lossless PNG only: png.optimize): device decode (+ K3 resize) to RGB, samples back to the host as PNG rows, then
// the lossless PNG leg (K6 filter selection, K7 LZ77).  A greyscale JPEG becomes a greyscale PNG.
b200_status jpeg_to_png(const uint8_t *in, size_t in_len, const b200_params *p, int prefer_dev, std::vector<uint8_t> &out)
{
    if (!p->png_optimize) return make_status(B200_ERR_UNSUPPORTED, "lossy PNG (imagequant) is outside the GPU path (route to caesium::convert_in_memory)");
    std::string err;
    JpegReader rd(in, in_len);
    if (!rd.read_header(err)) return header_status(err);
    const JpegGeom &gin = rd.geom();
    if (!ensure_runtime(err)) return make_status(B200_ERR_NO_DEVICE, err);
    uint32_t nw = (uint32_t)gin.width, nh = (uint32_t)gin.height;
    if (p->width || p->height) compute_resize_dimensions((uint32_t)gin.width, (uint32_t)gin.height, p->width, p->height, nw, nh);
    if (nw == 0 || nh == 0 || nw > 65535 || nh > 65535) return make_status(B200_ERR_INVALID_ARGUMENT, "invalid target dimensions");
    JpegGeom gout = gin; gout.width = (int)nw; gout.height = (int)nh; gout.finalize();
    if (g_entropy_mode.load() < 0) {
        const char *e = getenv("B200_ENTROPY");
        g_entropy_mode.store(!e ? 3 : !strcmp(e, "host") ? 0 : !strcmp(e, "gpuenc") ? 1 : !strcmp(e, "gpudec") ? 2 : 3);
    }
    Slot *s = slot_acquire(prefer_dev < 0 ? runtime_next_device() : prefer_dev, err);
    if (!s) return make_status(B200_ERR_CUDA, err);
    b200_status st = ok_status();
    const int nc = gin.ncomp == 1 ? 1 : 3;
    const size_t n = (size_t)nw * nh;
    std::vector<uint8_t> planes((size_t)nc * n), raw;
    do {
        if (!s->ensure((size_t)gin.total_coefs * 2, (size_t)gout.total_coefs * 2, 0, 1 << 14, err)) { st = make_status(B200_ERR_OUT_OF_MEMORY, err); break; }
        bool on_device = false;
        JpegReader::DeviceScan ds;
        if ((g_entropy_mode.load() & 2) && rd.device_decodable(ds)) {
            const int r = slot_gpu_decode(s, rd, ds, err);
            if (r == 0) on_device = true; else if (r != 1) { st = make_status(B200_ERR_CUDA, err); break; }
        }
        if (!on_device && !rd.decode(s->h_in, err)) { st = make_status(B200_ERR_CORRUPT_INPUT, err); break; }
        uint8_t *rgb[3] = {nullptr, nullptr, nullptr};
        if (!slot_transform_resized(s, gin, gout, err, false, !on_device, rgb)) { st = make_status(B200_ERR_CUDA, err); break; }
        if (!slot_fetch_planes(s, rgb, nc, n, planes.data(), err)) { st = make_status(B200_ERR_CUDA, err); break; }
        raw.resize((size_t)nc * n);
        if (nc == 1) raw = planes;
        else for (size_t i = 0; i < n; i++) { raw[3 * i] = planes[i]; raw[3 * i + 1] = planes[n + i]; raw[3 * i + 2] = planes[2 * n + i]; }
        PngInfo info; info.width = nw; info.height = nh; info.bit_depth = 8; info.color_type = nc == 1 ? 0 : 2; info.channels = nc;
        info.bits_per_pixel = 8 * nc; info.bpp = nc; info.row_bytes = (size_t)nw * nc;
        if (!s->png) s->png = new PngDevice();
        std::vector<uint8_t> z;
        int level = (int)p->png_optimization_level; if (level > 6) level = 6;
        if (!s->png->compress(info, raw, level, s->stream, z, nullptr, err)) { st = make_status(B200_ERR_CUDA, err); break; }
        png_write(info, z, out);
    } while (0);
    slot_release(s);
    return st;
}

// Decoded PNG samples -> 8-bit planar samples on the host: palette looked up, sub-byte greys scaled, 16-bit -> high byte,
// alpha dropped (like the image crate's to_rgb8 / to_luma8).  allow_grey: grey colour types stay one plane (JPEG target).
void png_expand_planar(const PngInfo &info, const std::vector<uint8_t> &raw, bool allow_grey, std::vector<uint8_t> &planes, int &nc)
{
    const size_t w = info.width, h = info.height, n = w * h;
    const int bd = info.bit_depth, ct = info.color_type;
    nc = (allow_grey && (ct == 0 || ct == 4)) ? 1 : 3;
    planes.resize((size_t)nc * n);
    for (size_t y = 0; y < h; y++) {
        const uint8_t *row = raw.data() + y * info.row_bytes;
        for (size_t x = 0; x < w; x++) {
            uint8_t r, g, b;
            if (ct == 2 || ct == 6) { const size_t o = x * info.channels * (bd / 8); r = row[o]; g = row[o + bd / 8]; b = row[o + 2 * (bd / 8)]; }
            else {
                unsigned v;
                if (bd >= 8) v = row[x * info.channels * (bd / 8)];
                else v = (row[(x * bd) >> 3] >> (8 - bd - ((x * bd) & 7))) & ((1u << bd) - 1);
                if (ct == 3) { if (3 * v + 2 < info.plte.size()) { r = info.plte[3 * v]; g = info.plte[3 * v + 1]; b = info.plte[3 * v + 2]; } else r = g = b = 0; }
                else { if (bd < 8) v = v * 255 / ((1u << bd) - 1); r = g = b = (uint8_t)v; }
            }
            planes[y * w + x] = r;
            if (nc == 3) { planes[n + y * w + x] = g; planes[2 * n + y * w + x] = b; }
        }
    }
}

// Planar 8-bit samples on the host ([nc][H][W]: RGB or one grey plane) -> JPEG: they take the resize path's back end (K3 Lanczos3 when
// width/height are set, RGB -> YCbCr, K4 box downsample, K5 FDCT + quantise) and the device Huffman encoder.
b200_status planes_to_jpeg(const std::vector<uint8_t> &planes, uint32_t w, uint32_t h, int nc, const b200_params *p, int prefer_dev, std::vector<uint8_t> &out)
{
    std::string err;
    if (w > 65535 || h > 65535) return make_status(B200_ERR_INVALID_ARGUMENT, "image too large for JPEG");
    JpegGeom gin; gin.width = (int)w; gin.height = (int)h; gin.ncomp = nc;
    for (int c = 0; c < nc; c++) { gin.cid[c] = c + 1; gin.hs[c] = gin.vs[c] = 1; gin.tq[c] = 0; }
    gin.finalize();
    JpegGeom gout;
    if (!jpeg_output_geom(gin, (int)p->jpeg_quality, (int)p->jpeg_chroma_subsampling, gout, err)) return make_status(B200_ERR_INVALID_ARGUMENT, err);
    if (p->width || p->height) {
        uint32_t nw = 0, nh = 0;
        compute_resize_dimensions(w, h, p->width, p->height, nw, nh);
        if (nw == 0 || nh == 0 || nw > 65535 || nh > 65535) return make_status(B200_ERR_INVALID_ARGUMENT, "invalid target dimensions");
        gout.width = (int)nw; gout.height = (int)nh; gout.finalize();
    }
    if (!ensure_runtime(err)) return make_status(B200_ERR_NO_DEVICE, err);
    Slot *s = slot_acquire(prefer_dev < 0 ? runtime_next_device() : prefer_dev, err);
    if (!s) return make_status(B200_ERR_CUDA, err);
    b200_status st = ok_status();
    JpegWriteOptions wo; wo.progressive = p->jpeg_progressive != 0;
    do {
        if (!slot_transform_resized(s, gin, gout, err, false, false, nullptr, planes.data())) { st = make_status(B200_ERR_CUDA, err); break; }
        if (slot_gpu_encode(s, gout, wo.progressive, err)) {
            if (!jpeg_assemble(gout, wo, nullptr, s->enc->results.data(), (int)s->enc->results.size(), out, err)) st = make_status(B200_ERR_INVALID_ARGUMENT, err);
            break;
        }
        if (!s->enc || !s->enc->overflow) { st = make_status(B200_ERR_CUDA, err); break; }
        if (!slot_download_coefs(s, (size_t)gout.total_coefs * 2, err)) { st = make_status(B200_ERR_CUDA, err); break; }
        jpeg_fill_dummy_blocks(gout, s->h_out);
        if (!jpeg_write(gout, s->h_out, wo, nullptr, out, err)) st = make_status(B200_ERR_INVALID_ARGUMENT, err);
    } while (0);
    slot_release(s);
    return st;
}

b200_status png_to_jpeg(const uint8_t *in, size_t in_len, const b200_params *p, int prefer_dev, std::vector<uint8_t> &out)
{
    std::string err;
    PngInfo info; std::vector<uint8_t> raw;
    if (!png_decode(in, in_len, false, info, raw, err)) return make_status(err.find("interlace") != std::string::npos ? B200_ERR_UNSUPPORTED : B200_ERR_CORRUPT_INPUT, err);
    std::vector<uint8_t> planes; int nc = 3;
    png_expand_planar(info, raw, true, planes, nc);
    return planes_to_jpeg(planes, info.width, info.height, nc, p, prefer_dev, out);
}

// Planar RGB on the host -> lossy WebP (K3 resize when width / height are set, then K8)
// alpha (optional): one 8-bit plane of the source's size.  It is resized like the colour planes, its LZ77 tokens come from K7 on the
// device, and the file becomes VP8X + ALPH (VP8L-coded, lossless -- libwebp's default alpha_quality 100) + VP8.
b200_status rgb_to_webp(const std::vector<uint8_t> &rgb, uint32_t w, uint32_t h, const b200_params *p, int prefer_dev, std::vector<uint8_t> &out,
                        const std::vector<uint8_t> *alpha = nullptr)
{
    std::string err;
    uint32_t nw = w, nh = h;
    if (p->width || p->height) compute_resize_dimensions(w, h, p->width, p->height, nw, nh);
    if (nw == 0 || nh == 0 || nw > 16383 || nh > 16383 || w > 65535 || h > 65535) return make_status(B200_ERR_INVALID_ARGUMENT, "invalid dimensions for WebP");
    if (!ensure_runtime(err)) return make_status(B200_ERR_NO_DEVICE, err);
    Slot *s = slot_acquire(prefer_dev < 0 ? runtime_next_device() : prefer_dev, err);
    if (!s) return make_status(B200_ERR_CUDA, err);
    if (!s->webp) s->webp = new WebpDevice();
    bool ok;
    if (nw == w && nh == h) ok = s->webp->encode_host_rgb(rgb.data(), (int)nw, (int)nh, (int)p->webp_quality, s->stream, out, err);
    else {   // through the resize leg (K3 Lanczos3) first
        JpegGeom gin; gin.width = (int)w; gin.height = (int)h; gin.ncomp = 3;
        for (int c = 0; c < 3; c++) { gin.cid[c] = c + 1; gin.hs[c] = gin.vs[c] = 1; gin.tq[c] = 0; }
        gin.finalize();
        JpegGeom gout = gin; gout.width = (int)nw; gout.height = (int)nh; gout.finalize();
        uint8_t *planes[3] = {nullptr, nullptr, nullptr};
        ok = slot_transform_resized(s, gin, gout, err, false, false, planes, rgb.data()) &&
             s->webp->encode_planes(planes[0], planes[1], planes[2], (int)nw, (int)nh, (int)p->webp_quality, s->stream, out, err);
    }
    if (ok && alpha) {
        const size_t n = (size_t)nw * nh;
        std::vector<uint8_t> resized;
        const uint8_t *ap = alpha->data();
        if (nw != w || nh != h) {
            JpegGeom gin; gin.width = (int)w; gin.height = (int)h; gin.ncomp = 1; gin.cid[0] = 1; gin.hs[0] = gin.vs[0] = 1; gin.tq[0] = 0; gin.finalize();
            JpegGeom gout = gin; gout.width = (int)nw; gout.height = (int)nh; gout.finalize();
            uint8_t *planes[3] = {nullptr, nullptr, nullptr};
            resized.resize(n);
            ok = slot_transform_resized(s, gin, gout, err, false, false, planes, alpha->data()) && slot_fetch_planes(s, planes, 1, n, resized.data(), err);
            ap = resized.data();
        }
        bool opaque = true;
        for (size_t i = 0; ok && i < n; i++) if (ap[i] != 0xFF) { opaque = false; break; }
        if (ok && !opaque) {
            if (!s->png) s->png = new PngDevice();
            std::vector<uint32_t> tokens; std::vector<uint8_t> alph, wrapped, residual;
            const int filter = webp_alpha_choose_filter(ap, (int)nw, (int)nh, residual);
            ok = s->png->plane_tokens(filter ? residual.data() : ap, n, (int)nw, s->stream, tokens, err);
            if (ok && !(vp8l_alpha_from_tokens(tokens.data(), tokens.size(), (int)nw, (int)nh, alph, filter) && webp_wrap_alpha(out, alph, (int)nw, (int)nh, wrapped))) { ok = false; err = "alpha plane could not be coded"; }
            if (ok) out.swap(wrapped);
        }
    }
    slot_release(s);
    return ok ? ok_status() : make_status(B200_ERR_CUDA, err);
}

// The transparency of decoded PNG samples as one 8-bit plane (alpha channel: high byte of a 16-bit sample; tRNS: the palette's
// per-entry alpha or the colour key).  false: every pixel is opaque.
bool png_extract_alpha(const PngInfo &info, const std::vector<uint8_t> &raw, std::vector<uint8_t> &alpha)
{
    const size_t w = info.width, h = info.height;
    const int bd = info.bit_depth, ct = info.color_type;
    const bool channel = ct == 4 || ct == 6;
    if (!channel && info.trns.empty()) return false;
    alpha.assign(w * h, 0xFF);
    bool any = false;
    const size_t bps = bd >= 8 ? (size_t)bd / 8 : 1;
    for (size_t y = 0; y < h; y++) {
        const uint8_t *row = raw.data() + y * info.row_bytes;
        uint8_t *a = alpha.data() + y * w;
        for (size_t x = 0; x < w; x++) {
            uint8_t v = 0xFF;
            if (channel) v = row[(x * info.channels + info.channels - 1) * bps];
            else if (ct == 3) { const unsigned idx = (row[(x * bd) >> 3] >> (8 - bd - ((x * bd) & 7))) & ((1u << bd) - 1); if (idx < info.trns.size()) v = info.trns[idx]; }
            else if (ct == 0 && info.trns.size() >= 2) {
                const unsigned key = ((unsigned)info.trns[0] << 8) | info.trns[1];
                const unsigned sv = bd == 16 ? (((unsigned)row[2 * x] << 8) | row[2 * x + 1]) : bd == 8 ? row[x] : (row[(x * bd) >> 3] >> (8 - bd - ((x * bd) & 7))) & ((1u << bd) - 1);
                if (sv == key) v = 0;
            } else if (ct == 2 && info.trns.size() >= 6) {
                bool eq = true;
                for (int c = 0; c < 3 && eq; c++) {
                    const unsigned key = ((unsigned)info.trns[2 * c] << 8) | info.trns[2 * c + 1];
                    const unsigned sv = bd == 16 ? (((unsigned)row[(3 * x + c) * 2] << 8) | row[(3 * x + c) * 2 + 1]) : row[3 * x + c];
                    eq = sv == key;
                }
                if (eq) v = 0;
            }
            a[x] = v; any |= v != 0xFF;
        }
    }
    return any;
}

// WebP input (libcaesium webp::compress: decode, optional resize, re-encode at webp.quality): the VP8 bitstream is decoded on the
// calling thread (format plumbing, bit-exact with libwebp's decoder -- vp8_decode.cpp), the RGB goes through K3 / K8 like any other source.
b200_status webp_decode_status(const uint8_t *in, size_t in_len, WebpInfo &info, std::vector<uint8_t> &rgb, std::vector<uint8_t> *alpha = nullptr)
{
    std::string err;
    const int rc = webp_decode_rgb(in, in_len, info, rgb, err, alpha);
    if (rc == 1) return make_status(B200_ERR_UNSUPPORTED, err);
    if (rc) return make_status(B200_ERR_CORRUPT_INPUT, err);
    return ok_status();
}
b200_status webp_compress(const uint8_t *in, size_t in_len, const b200_params *p, int prefer_dev, std::vector<uint8_t> &out)
{
    if (p->webp_lossless) return make_status(B200_ERR_UNSUPPORTED, "lossless WebP (VP8L) is outside the GPU path (route to caesium::compress_in_memory)");
    WebpInfo info; std::vector<uint8_t> rgb, alpha;
    b200_status st = webp_decode_status(in, in_len, info, rgb, &alpha);
    if (st.code) return st;
    return rgb_to_webp(rgb, (uint32_t)info.width, (uint32_t)info.height, p, prefer_dev, out, alpha.empty() ? nullptr : &alpha);
}

// PNG source: samples are expanded to 8-bit RGB on the host (palette, grey, 16-bit -> high byte) and go through the same K8; an
// opaque alpha channel is discarded, real transparency becomes the file's alpha plane.
b200_status png_to_webp(const uint8_t *in, size_t in_len, const b200_params *p, int prefer_dev, std::vector<uint8_t> &out)
{
    std::string err;
    PngInfo info; std::vector<uint8_t> raw;
    if (!png_decode(in, in_len, false, info, raw, err)) return make_status(err.find("interlace") != std::string::npos ? B200_ERR_UNSUPPORTED : B200_ERR_CORRUPT_INPUT, err);
    uint32_t nw = info.width, nh = info.height;
    if (p->width || p->height) compute_resize_dimensions(info.width, info.height, p->width, p->height, nw, nh);
    if (nw == 0 || nh == 0 || nw > 16383 || nh > 16383 || info.width > 65535 || info.height > 65535) return make_status(B200_ERR_INVALID_ARGUMENT, "invalid dimensions for WebP");
    // transparency (alpha channel, tRNS) travels as the file's alpha plane: VP8X + ALPH next to the lossy frame
    std::vector<uint8_t> alpha;
    const bool has_alpha = png_extract_alpha(info, raw, alpha);
    std::vector<uint8_t> rgb; int nc = 3;
    png_expand_planar(info, raw, false, rgb, nc);
    return rgb_to_webp(rgb, info.width, info.height, p, prefer_dev, out, has_alpha ? &alpha : nullptr);
}

// Planar RGB on the host -> lossless PNG (K3 resize when asked, then the PNG leg's raw-sample entry point)
b200_status rgb_to_png(const std::vector<uint8_t> &rgb, uint32_t w, uint32_t h, const b200_params *p, int prefer_dev, std::vector<uint8_t> &out,
                       const std::vector<uint8_t> *alpha = nullptr)
{
    std::string err;
    uint32_t nw = w, nh = h;
    if (p->width || p->height) compute_resize_dimensions(w, h, p->width, p->height, nw, nh);
    if (nw == 0 || nh == 0 || nw > 65535 || nh > 65535 || w > 65535 || h > 65535) return make_status(B200_ERR_INVALID_ARGUMENT, "invalid target dimensions");
    if (!ensure_runtime(err)) return make_status(B200_ERR_NO_DEVICE, err);
    Slot *s = slot_acquire(prefer_dev < 0 ? runtime_next_device() : prefer_dev, err);
    if (!s) return make_status(B200_ERR_CUDA, err);
    b200_status st = ok_status();
    const size_t n = (size_t)nw * nh;
    std::vector<uint8_t> planes, raw(3 * n);
    const uint8_t *src = rgb.data();
    do {
        if (nw != w || nh != h) {
            JpegGeom gin; gin.width = (int)w; gin.height = (int)h; gin.ncomp = 3;
            for (int c = 0; c < 3; c++) { gin.cid[c] = c + 1; gin.hs[c] = gin.vs[c] = 1; gin.tq[c] = 0; }
            gin.finalize();
            JpegGeom gout = gin; gout.width = (int)nw; gout.height = (int)nh; gout.finalize();
            uint8_t *dp[3] = {nullptr, nullptr, nullptr};
            planes.resize(3 * n);
            if (!slot_transform_resized(s, gin, gout, err, false, false, dp, rgb.data()) || !slot_fetch_planes(s, dp, 3, n, planes.data(), err)) { st = make_status(B200_ERR_CUDA, err); break; }
            src = planes.data();
        }
        PngInfo info; info.width = nw; info.height = nh; info.bit_depth = 8;
        if (!alpha) {
            for (size_t i = 0; i < n; i++) { raw[3 * i] = src[i]; raw[3 * i + 1] = src[n + i]; raw[3 * i + 2] = src[2 * n + i]; }
            info.color_type = 2; info.channels = 3; info.bits_per_pixel = 24; info.bpp = 3; info.row_bytes = (size_t)nw * 3;
        } else {
            // transparency stays: RGBA samples (the alpha plane takes the same Lanczos3 as the colour planes)
            std::vector<uint8_t> ra;
            const uint8_t *ap = alpha->data();
            if (nw != w || nh != h) {
                JpegGeom gin; gin.width = (int)w; gin.height = (int)h; gin.ncomp = 1; gin.cid[0] = 1; gin.hs[0] = gin.vs[0] = 1; gin.tq[0] = 0; gin.finalize();
                JpegGeom gout = gin; gout.width = (int)nw; gout.height = (int)nh; gout.finalize();
                uint8_t *dp[3] = {nullptr, nullptr, nullptr};
                ra.resize(n);
                if (!slot_transform_resized(s, gin, gout, err, false, false, dp, alpha->data()) || !slot_fetch_planes(s, dp, 1, n, ra.data(), err)) { st = make_status(B200_ERR_CUDA, err); break; }
                ap = ra.data();
            }
            raw.resize(4 * n);
            for (size_t i = 0; i < n; i++) { raw[4 * i] = src[i]; raw[4 * i + 1] = src[n + i]; raw[4 * i + 2] = src[2 * n + i]; raw[4 * i + 3] = ap[i]; }
            info.color_type = 6; info.channels = 4; info.bits_per_pixel = 32; info.bpp = 4; info.row_bytes = (size_t)nw * 4;
        }
        png_reduce_palette(info, raw);
        if (!s->png) s->png = new PngDevice();
        std::vector<uint8_t> z;
        int level = (int)p->png_optimization_level; if (level > 6) level = 6;
        if (!s->png->compress(info, raw, level, s->stream, z, nullptr, err)) { st = make_status(B200_ERR_CUDA, err); break; }
        png_write(info, z, out);
    } while (0);
    slot_release(s);
    return st;
}

b200_status compress_dispatch(const uint8_t *in, size_t in_len, const b200_params *p, int prefer_dev, std::vector<uint8_t> &out)
{
    switch (b200_sniff_format(in, in_len)) {
        case B200_FMT_JPEG: return jpeg_compress(in, in_len, p, prefer_dev, out);
        case B200_FMT_PNG: return png_compress(in, in_len, p, prefer_dev, out);
        case B200_FMT_WEBP: return webp_compress(in, in_len, p, prefer_dev, out);
        case B200_FMT_GIF: return make_status(B200_ERR_UNSUPPORTED, "GIF is outside the GPU path (route to caesium::compress_in_memory)");
        case B200_FMT_TIFF: return make_status(B200_ERR_UNSUPPORTED, "TIFF is outside the GPU path (route to caesium::compress_in_memory)");
        default: return make_status(B200_ERR_UNKNOWN_FORMAT, "Unknown file type");
    }
}

// libcaesium compress_to_size: quality bisection in [1, 100] from 80, at most 10 tries, 2 % tolerance, the largest result under
// the limit wins.  `size_at(q, keep)` runs one try: it returns the output size at quality q and, when `keep` says so (the try
// is the best so far, or the smallest so far for return_smallest), leaves the file in `cur`.
template <class SizeAt>
static b200_status bisect_quality(SizeAt size_at, size_t max_output_size, bool return_smallest, uint32_t *quality_out, std::vector<uint8_t> &result)
{
    const size_t tolerance = max_output_size / 50;
    int lo = 1, hi = 100, q = 80;
    std::vector<uint8_t> best, smallest, cur; size_t best_size = 0, smallest_size = (size_t)-1;
    for (int tries = 0; tries < 10 && lo <= hi; tries++) {
        size_t sz = 0;
        auto want = [&](size_t size) { return (size <= max_output_size && size > best_size) || (return_smallest && size < smallest_size); };
        b200_status s = size_at(q, want, sz, cur);
        if (s.code) return s;
        if (sz < smallest_size) { smallest_size = sz; if (return_smallest) smallest = cur; }
        if (sz <= max_output_size) {
            if (sz > best_size) { best_size = sz; best.swap(cur); if (quality_out) *quality_out = (uint32_t)q; }
            if (max_output_size - sz <= tolerance) break;
            lo = q + 1;
        } else hi = q - 1;
        q = (lo + hi) / 2;
    }
    if (best_size) { result.swap(best); return ok_status(); }
    if (return_smallest && !smallest.empty()) { result.swap(smallest); return ok_status(); }
    return make_status(B200_ERR_TOO_LARGE, "Cannot compress to desired size");
}

// JPEG: the source is entropy-decoded ONCE, its coefficients stay in HBM, and every try re-runs only dequant/IDCT/resample/FDCT/
// quantise at the try's tables plus the device Huffman encoder; the encoder reports the scan lengths from the device and the
// stuffed bytes are fetched only for tries that become the answer ("decode once, re-quantise many").
static b200_status jpeg_to_size(const uint8_t *in, size_t in_len, b200_params *params, size_t max_output_size, bool return_smallest, std::vector<uint8_t> &result)
{
    std::string err;
    JpegReader rd(in, in_len);
    if (!rd.read_header(err)) return header_status(err);
    const JpegGeom &gin = rd.geom();
    if (!ensure_runtime(err)) return make_status(B200_ERR_NO_DEVICE, err);
    if (g_entropy_mode.load() < 0) { const char *e = getenv("B200_ENTROPY"); g_entropy_mode.store(!e ? 3 : !strcmp(e, "host") ? 0 : !strcmp(e, "gpuenc") ? 1 : !strcmp(e, "gpudec") ? 2 : 3); }
    if (params->width || params->height || (g_entropy_mode.load() & 1) == 0) {
        // resize, or the host-entropy mode: every try is a whole compress call (the resized planes are not kept between tries)
        auto size_at = [&](int q, auto want, size_t &sz, std::vector<uint8_t> &cur) {
            b200_params p = *params; p.jpeg_quality = (uint32_t)q; p.jpeg_optimize = 0;
            b200_status s = compress_dispatch(in, in_len, &p, -1, cur);
            sz = cur.size(); (void)want;
            return s;
        };
        return bisect_quality(size_at, max_output_size, return_smallest, &params->jpeg_quality, result);
    }
    Slot *s = slot_acquire(runtime_next_device(), err);
    if (!s) return make_status(B200_ERR_CUDA, err);
    b200_status st = ok_status();
    do {
        JpegGeom g0;
        if (!jpeg_output_geom(gin, 80, (int)params->jpeg_chroma_subsampling, g0, err)) { st = make_status(B200_ERR_INVALID_ARGUMENT, err); break; }
        ImagePlan plan;
        if (!plan_image(gin, g0, plan, err)) { st = make_status(B200_ERR_UNSUPPORTED, err); break; }
        if (!s->ensure(plan.in_bytes, plan.out_bytes, plan.scratch_bytes(), 1 << 14, err)) { st = make_status(B200_ERR_OUT_OF_MEMORY, err); break; }
        bool resident = false;                      // coefficients already in s->d_in?
        JpegReader::DeviceScan ds;
        if ((g_entropy_mode.load() & 2) && rd.device_decodable(ds)) {
            const int r = slot_gpu_decode(s, rd, ds, err);
            if (r == 0) resident = true; else if (r != 1) { st = make_status(B200_ERR_CUDA, err); break; }
        }
        if (!resident && !rd.decode(s->h_in, err)) { st = make_status(B200_ERR_CORRUPT_INPUT, err); break; }
        JpegWriteOptions wo; wo.progressive = params->jpeg_progressive != 0; wo.keep_metadata = params->keep_metadata != 0; wo.preserve_icc = params->jpeg_preserve_icc != 0;
        auto size_at = [&](int q, auto want, size_t &sz, std::vector<uint8_t> &cur) -> b200_status {
            JpegGeom gout; std::string e2;
            if (!jpeg_output_geom(gin, q, (int)params->jpeg_chroma_subsampling, gout, e2)) return make_status(B200_ERR_INVALID_ARGUMENT, e2);
            if (!slot_transform(s, gin, gout, e2, false, !resident)) return make_status(B200_ERR_CUDA, e2);
            resident = true;                        // the first try uploaded them if the host decoded
            if (!slot_gpu_encode_sizes(s, gout, wo.progressive, e2)) return make_status(B200_ERR_CUDA, e2);
            sz = jpeg_assembled_size(gout, wo, &rd.meta(), s->enc->results.data(), (int)s->enc->results.size());
            if (want(sz)) {
                if (!slot_gpu_fetch(s, e2)) return make_status(B200_ERR_CUDA, e2);
                if (!jpeg_assemble(gout, wo, &rd.meta(), s->enc->results.data(), (int)s->enc->results.size(), cur, e2)) return make_status(B200_ERR_INVALID_ARGUMENT, e2);
            }
            return ok_status();
        };
        st = bisect_quality(size_at, max_output_size, return_smallest, &params->jpeg_quality, result);
    } while (0);
    slot_release(s);
    return st;
}

} // namespace

extern "C" {

void b200_params_default(b200_params *p)
{   // CSParameters::new(): jpeg q80 auto-subsampling progressive, png q80 level 3, gif 80, webp 60... caesiumclt overwrites the qualities (compressor.rs:415-417)
    memset(p, 0, sizeof(*p));
    p->jpeg_quality = 80; p->jpeg_chroma_subsampling = B200_CS_AUTO; p->jpeg_progressive = 1; p->jpeg_preserve_icc = 1;
    p->png_quality = 80; p->png_optimization_level = 3; p->gif_quality = 80; p->webp_quality = 80;
}

int b200_init(int n_gpus) { g_forced_ngpus = n_gpus; std::string e; return ensure_runtime(e) ? B200_OK : B200_ERR_NO_DEVICE; }
int b200_init_device(int ordinal) { g_forced_device = ordinal; std::string e; return ensure_runtime(e) ? B200_OK : B200_ERR_NO_DEVICE; }
void b200_shutdown(void) { print_trace(); runtime_shutdown(); }
int b200_device_count(void) { return runtime_device_count(); }
long long b200_device_jobs(int index) { return runtime_device_jobs(index); }
int b200_device_numa_node(int index) { return index < 0 || index >= runtime_device_count() ? -1 : device_numa_node(runtime_device_ordinal(index)); }
const char *b200_version(void) { return "b200-caesium 0.1.0 (sm_100a)"; }
void b200_free(void *p) { free(p); }
int b200_set_entropy_mode(int mode) { if (mode < 0 || mode > 3) return B200_ERR_INVALID_ARGUMENT; g_entropy_mode.store(mode); return B200_OK; }

uint32_t b200_sniff_format(const uint8_t *d, size_t n)
{   // the magic numbers `infer` checks (scan_files.rs:30-40, compressor.rs:259-264)
    if (!d) return B200_FMT_UNKNOWN;
    if (n >= 3 && d[0] == 0xFF && d[1] == 0xD8 && d[2] == 0xFF) return B200_FMT_JPEG;
    if (n >= 8 && !memcmp(d, "\x89PNG\r\n\x1a\n", 8)) return B200_FMT_PNG;
    if (n >= 6 && (!memcmp(d, "GIF87a", 6) || !memcmp(d, "GIF89a", 6))) return B200_FMT_GIF;
    if (n >= 12 && !memcmp(d, "RIFF", 4) && !memcmp(d + 8, "WEBP", 4)) return B200_FMT_WEBP;
    if (n >= 4 && (!memcmp(d, "II*\0", 4) || !memcmp(d, "MM\0*", 4))) return B200_FMT_TIFF;
    return B200_FMT_UNKNOWN;
}

// ---- call coalescing (opt-in: B200_COALESCE=1) ----------------------------------------------------------------------------------
// The reference calls the codec one image at a time from every rayon worker (compressor.rs:81-83, :305); the GPU wants several
// same-shaped JPEGs per launch sequence.  With coalescing on, concurrent b200_compress_in_memory calls on JPEG inputs with equal
// parameters meet in a queue: the first caller to find no collector waits a few hundred microseconds for company (or until
// B200_COALESCE_TARGET calls have gathered), takes every matching call out of the queue and runs them as ONE b200_compress_batch
// on its own thread; the others sleep until their result is in.  Several such batches can be in flight.  Per call the semantics
// (result bytes, status, ownership) are those of the direct path.  Off by default until its effect is measured on a GPU box.
namespace {
struct CoReq { const uint8_t *in; size_t len; uint8_t *out = nullptr; size_t out_len = 0; b200_status st{0, nullptr}; bool done = false; b200_params params; };
struct Coalescer {
    std::mutex mu; std::condition_variable cv;
    std::vector<CoReq *> pending; bool collecting = false;
};
Coalescer g_co;
std::atomic<long> g_co_calls{0}, g_co_batches{0};
struct CoReport { ~CoReport() { if (getenv("B200_TRACE") && g_co_batches.load()) fprintf(stderr, "[b200 trace] coalescing: %ld calls in %ld batches\n", g_co_calls.load(), g_co_batches.load()); } } g_co_report;
int coalesce_mode()       // 0 off, 1 on
{
    static const int m = [] { const char *e = getenv("B200_COALESCE"); return e && atoi(e) > 0 ? 1 : 0; }();
    return m;
}
bool same_params_abi(const b200_params &a, const b200_params &b)
{
    return a.keep_metadata == b.keep_metadata && a.jpeg_quality == b.jpeg_quality && a.jpeg_chroma_subsampling == b.jpeg_chroma_subsampling &&
           a.jpeg_progressive == b.jpeg_progressive && a.jpeg_optimize == b.jpeg_optimize && a.jpeg_preserve_icc == b.jpeg_preserve_icc &&
           a.png_quality == b.png_quality && a.png_optimization_level == b.png_optimization_level && a.png_force_zopfli == b.png_force_zopfli &&
           a.png_optimize == b.png_optimize && a.gif_quality == b.gif_quality && a.webp_quality == b.webp_quality && a.webp_lossless == b.webp_lossless &&
           a.width == b.width && a.height == b.height;
}
b200_status coalesced_compress(const uint8_t *in, size_t in_len, const b200_params *params, uint8_t **out, size_t *out_len)
{
    static const int target = [] { const char *e = getenv("B200_COALESCE_TARGET"); const int v = e ? atoi(e) : 0; return v >= 2 && v <= 1024 ? v : 16; }();
    static const int window_us = [] { const char *e = getenv("B200_COALESCE_US"); const int v = e ? atoi(e) : 0; return v >= 1 && v <= 100000 ? v : 400; }();
    CoReq me; me.in = in; me.len = in_len; me.params = *params;
    std::unique_lock<std::mutex> lk(g_co.mu);
    g_co.pending.push_back(&me);
    g_co.cv.notify_all();                                     // a collector may be waiting for company
    while (!me.done) {
        if (g_co.collecting) { g_co.cv.wait(lk); continue; }
        bool queued = false; for (CoReq *r : g_co.pending) if (r == &me) { queued = true; break; }
        if (!queued) { g_co.cv.wait(lk); continue; }         // my call is inside somebody's batch
        // become the collector for calls with my parameters
        g_co.collecting = true;
        const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(window_us);
        for (;;) {
            int have = 0; for (CoReq *r : g_co.pending) if (same_params_abi(r->params, me.params)) have++;
            if (have >= target || g_co.cv.wait_until(lk, deadline) == std::cv_status::timeout) break;
        }
        std::vector<CoReq *> batch, left;
        for (CoReq *r : g_co.pending) (same_params_abi(r->params, me.params) && (int)batch.size() < 4 * target ? batch : left).push_back(r);
        g_co.pending.swap(left);
        g_co.collecting = false;
        g_co.cv.notify_all();                                 // leftovers / new arrivals elect their own collector
        lk.unlock();
        const int n = (int)batch.size();
        g_co_calls += n; g_co_batches++;
        bool ran = false;
        try {
            std::vector<const uint8_t *> ins((size_t)n); std::vector<size_t> lens((size_t)n); std::vector<uint8_t *> outs((size_t)n, nullptr); std::vector<size_t> ol((size_t)n, 0);
            std::vector<b200_status> sts((size_t)n, b200_status{0, nullptr});
            for (int i = 0; i < n; i++) { ins[(size_t)i] = batch[(size_t)i]->in; lens[(size_t)i] = batch[(size_t)i]->len; }
            const int rc = b200_compress_batch(ins.data(), lens.data(), n, &me.params, std::min(n, 16), outs.data(), ol.data(), sts.data());
            lk.lock();
            for (int i = 0; i < n; i++) {
                CoReq *r = batch[(size_t)i];
                if (rc < 0) r->st = make_status(B200_ERR_INVALID_ARGUMENT, "batch call failed");
                else { r->st = sts[(size_t)i]; r->out = outs[(size_t)i]; r->out_len = ol[(size_t)i]; }
                r->done = true;
            }
            ran = true;
        } catch (...) {}
        if (!ran) {                                           // nobody may be left waiting on a batch that died (allocation failure)
            if (!lk.owns_lock()) lk.lock();
            for (CoReq *r : batch) if (!r->done) { r->st = make_status(B200_ERR_OUT_OF_MEMORY, "out of memory in the coalesced batch"); r->done = true; }
        }
        g_co.cv.notify_all();
    }
    lk.unlock();
    if (me.st.code) { if (me.out) b200_free(me.out); return me.st; }
    *out = me.out; *out_len = me.out_len;
    return me.st;
}
} // namespace

b200_status b200_compress_in_memory(const uint8_t *in, size_t in_len, const b200_params *params, uint8_t **out, size_t *out_len)
{
    if (!in || !params || !out || !out_len) return make_status(B200_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr; *out_len = 0;
    if (coalesce_mode() && b200_sniff_format(in, in_len) == B200_FMT_JPEG) {
        try { return coalesced_compress(in, in_len, params, out, out_len); }
        catch (const std::exception &e) { return make_status(B200_ERR_OUT_OF_MEMORY, e.what()); } catch (...) { return make_status(B200_ERR_INVALID_ARGUMENT, "unexpected failure"); }
    }
    try {
        std::vector<uint8_t> v;
        b200_status s = compress_dispatch(in, in_len, params, -1, v);
        if (s.code) return s;
        return give(v, out, out_len);
    } catch (const std::exception &e) { return make_status(B200_ERR_OUT_OF_MEMORY, e.what()); } catch (...) { return make_status(B200_ERR_INVALID_ARGUMENT, "unexpected failure"); }
}

b200_status b200_convert_in_memory(const uint8_t *in, size_t in_len, const b200_params *params, uint32_t fmt, uint8_t **out, size_t *out_len)
{
    if (!in || !params || !out || !out_len) return make_status(B200_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr; *out_len = 0;
    uint32_t src = b200_sniff_format(in, in_len);
    if (src == B200_FMT_UNKNOWN) return make_status(B200_ERR_UNKNOWN_FORMAT, "Unknown file type");
    if (src == fmt) return make_status(B200_ERR_SAME_FORMAT, "Cannot convert to the same format");
    const bool to_webp = fmt == B200_FMT_WEBP, png_to_jpg = fmt == B200_FMT_JPEG && src == B200_FMT_PNG, jpg_to_png = fmt == B200_FMT_PNG && src == B200_FMT_JPEG;
    if (jpg_to_png) {
        try { std::vector<uint8_t> v; b200_status s = jpeg_to_png(in, in_len, params, -1, v); if (s.code) return s; return give(v, out, out_len); }
        catch (const std::exception &e) { return make_status(B200_ERR_OUT_OF_MEMORY, e.what()); }
    }
    if (src == B200_FMT_WEBP && (fmt == B200_FMT_JPEG || fmt == B200_FMT_PNG)) {
        // WebP source: decoded on the calling thread (vp8_decode.cpp), then the same back ends as a PNG source
        try {
            if (fmt == B200_FMT_JPEG && params->jpeg_optimize) return make_status(B200_ERR_UNSUPPORTED, "lossless conversion to JPEG is outside the GPU path (route to caesium::convert_in_memory)");
            if (fmt == B200_FMT_PNG && !params->png_optimize) return make_status(B200_ERR_UNSUPPORTED, "lossy PNG (imagequant) is outside the GPU path (route to caesium::convert_in_memory)");
            WebpInfo wi; std::vector<uint8_t> rgb, alpha, v;
            b200_status s = webp_decode_status(in, in_len, wi, rgb, &alpha);
            if (s.code) return s;
            // a JPEG has no alpha (the image crate's to_rgb8 drops it); a PNG keeps it as an RGBA image
            s = fmt == B200_FMT_JPEG ? planes_to_jpeg(rgb, (uint32_t)wi.width, (uint32_t)wi.height, 3, params, -1, v)
                                     : rgb_to_png(rgb, (uint32_t)wi.width, (uint32_t)wi.height, params, -1, v, alpha.empty() ? nullptr : &alpha);
            if (s.code) return s;
            return give(v, out, out_len);
        } catch (const std::exception &e) { return make_status(B200_ERR_OUT_OF_MEMORY, e.what()); }
    }
    if (!to_webp && !png_to_jpg) return make_status(B200_ERR_UNSUPPORTED, "this conversion is outside the GPU path (route to caesium::convert_in_memory)");
    if (to_webp && params->webp_lossless) return make_status(B200_ERR_UNSUPPORTED, "lossless WebP (VP8L) is outside the GPU path (route to caesium::convert_in_memory)");
    if (png_to_jpg && params->jpeg_optimize) return make_status(B200_ERR_UNSUPPORTED, "lossless conversion to JPEG is outside the GPU path (route to caesium::convert_in_memory)");
    try {
        std::vector<uint8_t> v;
        b200_status s = png_to_jpg ? png_to_jpeg(in, in_len, params, -1, v)
                      : src == B200_FMT_JPEG ? jpeg_to_webp(in, in_len, params, -1, v)
                      : src == B200_FMT_PNG ? png_to_webp(in, in_len, params, -1, v)
                      : make_status(B200_ERR_UNSUPPORTED, "conversion from this format is outside the GPU path (route to caesium::convert_in_memory)");
        if (s.code) return s;
        return give(v, out, out_len);
    } catch (const std::exception &e) { return make_status(B200_ERR_OUT_OF_MEMORY, e.what()); }
}

// WebP: the source is decoded once, its RGB uploaded once (resized once if asked); every try runs K8 at the try's quality and the
// host boolean coder.
static b200_status webp_to_size(const uint8_t *in, size_t in_len, b200_params *params, size_t max_output_size, bool return_smallest, std::vector<uint8_t> &result)
{
    if (params->webp_lossless) return make_status(B200_ERR_UNSUPPORTED, "lossless WebP (VP8L) is outside the GPU path (route to caesium::compress_to_size_in_memory)");
    std::string err;
    WebpInfo wi; std::vector<uint8_t> rgb, alpha;
    b200_status st = webp_decode_status(in, in_len, wi, rgb, &alpha);
    if (st.code) return st;
    if (!alpha.empty()) return make_status(B200_ERR_UNSUPPORTED, "compress_to_size on a WebP with an alpha plane is outside the GPU path (route to caesium::compress_to_size_in_memory)");
    uint32_t w = (uint32_t)wi.width, h = (uint32_t)wi.height, nw = w, nh = h;
    if (params->width || params->height) compute_resize_dimensions(w, h, params->width, params->height, nw, nh);
    if (nw == 0 || nh == 0 || nw > 16383 || nh > 16383) return make_status(B200_ERR_INVALID_ARGUMENT, "invalid dimensions for WebP");
    if (!ensure_runtime(err)) return make_status(B200_ERR_NO_DEVICE, err);
    Slot *s = slot_acquire(runtime_next_device(), err);
    if (!s) return make_status(B200_ERR_CUDA, err);
    if (!s->webp) s->webp = new WebpDevice();
    do {
        uint8_t *planes[3] = {nullptr, nullptr, nullptr};
        JpegGeom gin; gin.width = (int)w; gin.height = (int)h; gin.ncomp = 3;
        for (int c = 0; c < 3; c++) { gin.cid[c] = c + 1; gin.hs[c] = gin.vs[c] = 1; gin.tq[c] = 0; }
        gin.finalize();
        JpegGeom gout = gin; gout.width = (int)nw; gout.height = (int)nh; gout.finalize();
        // the (possibly resized) RGB planes stay in the slot's scratch memory for all tries
        if (!slot_transform_resized(s, gin, gout, err, false, false, planes, rgb.data())) { st = make_status(B200_ERR_CUDA, err); break; }
        auto size_at = [&](int q, auto want, size_t &sz, std::vector<uint8_t> &cur) -> b200_status {
            std::string e2; (void)want;
            if (!s->webp->encode_planes(planes[0], planes[1], planes[2], (int)nw, (int)nh, q, s->stream, cur, e2)) return make_status(B200_ERR_CUDA, e2);
            sz = cur.size();
            return ok_status();
        };
        st = bisect_quality(size_at, max_output_size, return_smallest, &params->webp_quality, result);
    } while (0);
    slot_release(s);
    return st;
}

b200_status b200_compress_to_size_in_memory(const uint8_t *in, size_t in_len, b200_params *params, size_t max_output_size, uint8_t return_smallest,
                                            uint8_t **out, size_t *out_len)
{
    if (!in || !params || !out || !out_len) return make_status(B200_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr; *out_len = 0;
    try {
        // input already small enough is returned unchanged
        if (in_len <= max_output_size) { std::vector<uint8_t> v(in, in + in_len); return give(v, out, out_len); }
        uint32_t fmt = b200_sniff_format(in, in_len);
        std::vector<uint8_t> result;
        b200_status s;
        if (fmt == B200_FMT_JPEG) s = jpeg_to_size(in, in_len, params, max_output_size, return_smallest != 0, result);
        else if (fmt == B200_FMT_WEBP) s = webp_to_size(in, in_len, params, max_output_size, return_smallest != 0, result);
        else if (fmt == B200_FMT_PNG) s = make_status(B200_ERR_UNSUPPORTED, "compress_to_size on a PNG bisects the lossy (imagequant) quality, which is outside the GPU path (route to caesium::compress_to_size_in_memory)");
        else s = make_status(fmt == B200_FMT_UNKNOWN ? B200_ERR_UNKNOWN_FORMAT : B200_ERR_UNSUPPORTED, "compress_to_size for this format is outside the GPU path (route to caesium::compress_to_size_in_memory)");
        if (s.code) return s;
        return give(result, out, out_len);
    } catch (const std::exception &e) { return make_status(B200_ERR_OUT_OF_MEMORY, e.what()); } catch (...) { return make_status(B200_ERR_INVALID_ARGUMENT, "unexpected failure"); }
}

int b200_compress_batch(const uint8_t *const *in, const size_t *in_len, int n, const b200_params *params, int n_threads,
                        uint8_t **out, size_t *out_len, b200_status *status)
{
    if (!in || !in_len || !params || !out || !out_len || !status || n < 0) return -1;
    if (n_threads <= 0) n_threads = usable_cores();
    if (n_threads > n) n_threads = n;
    std::atomic<int> failed{0};
    { std::string e; ensure_runtime(e); }
    const int ndev = std::max(1, runtime_device_count());
    if (g_entropy_mode.load() < 0) { const char *e = getenv("B200_ENTROPY"); g_entropy_mode.store(!e ? 3 : !strcmp(e, "host") ? 0 : !strcmp(e, "gpuenc") ? 1 : !strcmp(e, "gpudec") ? 2 : 3); }
    // Megabatches: with both entropy stages on the device, consecutive images are processed K at a time -- one launch
    // sequence (decode rounds, transform, encode passes) for the whole group instead of one per image.  Images that do not
    // fit the group path (other formats, progressive input, resize, odd one out in shape) go through the per-image path.
    int K = 8; { const char *e = getenv("B200_MEGABATCH"); if (e) K = std::max(1, std::min(64, atoi(e))); }
    const bool grouped = runtime_device_count() > 0 && g_entropy_mode.load() == 3 && K > 1 && ((!params->width && !params->height) || params->jpeg_optimize);
    auto one = [&](int i, int dev) {
        out[i] = nullptr; out_len[i] = 0;
        try {
            std::vector<uint8_t> v;
            status[i] = compress_dispatch(in[i], in_len[i], params, dev, v);
            if (!status[i].code) status[i] = give(v, &out[i], &out_len[i]);
        } catch (const std::exception &e) { status[i] = make_status(B200_ERR_OUT_OF_MEMORY, e.what()); } catch (...) { status[i] = make_status(B200_ERR_INVALID_ARGUMENT, "unexpected failure"); }
        if (status[i].code) failed++;
    };
    auto run_threads = [](int nt, const std::function<void()> &fn) {
        std::vector<std::thread> th;
        for (int t = 1; t < nt; t++) th.emplace_back(fn);
        fn();
        for (auto &t : th) t.join();
    };
    // phase 1: JPEGs, K at a time (one megabatch = one long launch sequence on one slot of one device).  Megabatches are sharded
    // over the devices by bytes -- longest-processing-time-first: largest megabatch to the least loaded device --
    // and every device gets its own workers, bound to the CPUs of the device's NUMA node; a worker whose device runs dry takes
    // work from the most loaded one.  Whatever a megabatch could not take is left for phase 2.
    std::vector<int> rest;
    if (grouped) {
        std::vector<int> jpegs;
        for (int i = 0; i < n; i++) (b200_sniff_format(in[i], in_len[i]) == B200_FMT_JPEG ? jpegs : rest).push_back(i);
        if (jpegs.size() < 2) { rest.insert(rest.end(), jpegs.begin(), jpegs.end()); jpegs.clear(); }
        const int nj = (int)jpegs.size();
        if (nj) {
            const int nchunks = (nj + K - 1) / K;
            std::vector<size_t> cbytes((size_t)nchunks, 0);
            for (int c = 0; c < nchunks; c++) for (int j = c * K; j < std::min(nj, c * K + K); j++) cbytes[(size_t)c] += in_len[jpegs[(size_t)j]];
            std::vector<int> order((size_t)nchunks); for (int c = 0; c < nchunks; c++) order[(size_t)c] = c;
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cbytes[(size_t)a] > cbytes[(size_t)b]; });
            std::vector<std::vector<int>> devq((size_t)ndev); std::vector<size_t> load((size_t)ndev, 0);
            for (int c : order) { int d = 0; for (int e = 1; e < ndev; e++) if (load[(size_t)e] < load[(size_t)d]) d = e; devq[(size_t)d].push_back(c); load[(size_t)d] += cbytes[(size_t)c]; }
            std::vector<std::atomic<int>> cursor((size_t)ndev); for (auto &c : cursor) c.store(0);
            std::mutex rest_mu;
            int max_workers = 16; { const char *e = getenv("B200_GROUP_WORKERS"); if (e) max_workers = std::max(1, std::min(32, atoi(e))); }
            std::atomic<int> worker_id{0};
            auto group_worker = [&]() {
                const int home = worker_id.fetch_add(1) % ndev;
                AffinityGuard pin(runtime_device_ordinal(home));
                for (;;) {
                    int dev = home, c = -1;
                    { const int k = cursor[(size_t)home].fetch_add(1); if (k < (int)devq[(size_t)home].size()) c = devq[(size_t)home][(size_t)k]; }
                    if (c < 0) {           // home queue empty: help the device with the most work left
                        int best = -1, left = 0;
                        for (int e = 0; e < ndev; e++) { const int l = (int)devq[(size_t)e].size() - cursor[(size_t)e].load(); if (l > left) { left = l; best = e; } }
                        if (best < 0) break;
                        const int k = cursor[(size_t)best].fetch_add(1);
                        if (k >= (int)devq[(size_t)best].size()) continue;
                        c = devq[(size_t)best][(size_t)k]; dev = best;
                    }
                    const int j0 = c * K, j1 = std::min(nj, j0 + K);
                    std::vector<int> idx(jpegs.begin() + j0, jpegs.begin() + j1);
                    std::vector<char> done(idx.size(), 0);
                    try { jpeg_compress_group(in, in_len, idx, params, dev, out, out_len, status, done); } catch (...) {}
                    for (size_t k = 0; k < idx.size(); k++) {
                        if (done[k]) { if (status[idx[k]].code) failed++; }
                        else { std::lock_guard<std::mutex> lk(rest_mu); rest.push_back(idx[k]); }
                    }
                }
            };
            run_threads(std::max(1, std::min(n_threads, std::min(max_workers * ndev, nchunks))), group_worker);
        }
    } else for (int i = 0; i < n; i++) rest.push_back(i);
    // phase 2: one image per call on every thread the caller allows (PNG, conversions' sources, progressive JPEGs, ...)
    if (!rest.empty()) {
        std::sort(rest.begin(), rest.end());
        std::atomic<int> nr{0};
        const int total = (int)rest.size();
        std::atomic<int> tid{0};
        run_threads(std::min(n_threads, total), [&]() {
            const int home = tid.fetch_add(1) % ndev;            // thread t serves device t % ndev from that device's NUMA node
            AffinityGuard pin(runtime_device_ordinal(home));
            for (;;) { const int r = nr.fetch_add(1); if (r >= total) break; one(rest[r], home); }
        });
    }
    return failed.load();
}

// ---- JPEG stage entry points ------------------------------------------------------------------------------------
b200_status b200_jpeg_decode_coefficients(const uint8_t *in, size_t in_len, b200_jpeg_layout *layout, int16_t **coefs)
{
    if (!in || !layout || !coefs) return make_status(B200_ERR_INVALID_ARGUMENT, "null argument");
    *coefs = nullptr;
    std::string err;
    JpegReader rd(in, in_len);
    if (!rd.read_header(err)) return header_status(err);
    int16_t *c = (int16_t *)malloc((size_t)rd.geom().total_coefs * 2 + 16);
    if (!c) return make_status(B200_ERR_OUT_OF_MEMORY, "out of memory");
    if (!rd.decode(c, err)) { free(c); return make_status(B200_ERR_CORRUPT_INPUT, err); }
    layout_from_geom(rd.geom(), layout);
    *coefs = c;
    return ok_status();
}

b200_status b200_jpeg_output_layout(const b200_jpeg_layout *in_layout, const b200_params *params, b200_jpeg_layout *out_layout)
{
    if (!in_layout || !params || !out_layout) return make_status(B200_ERR_INVALID_ARGUMENT, "null argument");
    std::string err; JpegGeom gin, gout;
    if (!geom_from_layout(in_layout, gin, err)) return make_status(B200_ERR_INVALID_ARGUMENT, err);
    if (!jpeg_output_geom(gin, (int)params->jpeg_quality, (int)params->jpeg_chroma_subsampling, gout, err)) return make_status(B200_ERR_INVALID_ARGUMENT, err);
    gout.progressive = params->jpeg_progressive != 0;
    layout_from_geom(gout, out_layout);
    return ok_status();
}

b200_status b200_jpeg_requantize(const b200_jpeg_layout *in_layout, const int16_t *in_coefs, const b200_jpeg_layout *out_layout, int16_t *out_coefs)
{
    if (!in_layout || !in_coefs || !out_layout || !out_coefs) return make_status(B200_ERR_INVALID_ARGUMENT, "null argument");
    std::string err; JpegGeom gin, gout;
    if (!geom_from_layout(in_layout, gin, err) || !geom_from_layout(out_layout, gout, err)) return make_status(B200_ERR_INVALID_ARGUMENT, err);
    if (!ensure_runtime(err)) return make_status(B200_ERR_NO_DEVICE, err);
    ImagePlan plan;
    if (!plan_image(gin, gout, plan, err)) return make_status(B200_ERR_UNSUPPORTED, err);
    Slot *s = slot_acquire(runtime_next_device(), err);
    if (!s) return make_status(B200_ERR_CUDA, err);
    b200_status st = ok_status();
    do {
        if (!s->ensure(plan.in_bytes, plan.out_bytes, plan.scratch_bytes(), 1 << 14, err)) { st = make_status(B200_ERR_OUT_OF_MEMORY, err); break; }
        memcpy(s->h_in, in_coefs, plan.in_bytes);
        memset(s->h_out, 0, plan.out_bytes);
        if (!slot_transform(s, gin, gout, err)) { st = make_status(B200_ERR_CUDA, err); break; }
        jpeg_fill_dummy_blocks(gout, s->h_out);
        memcpy(out_coefs, s->h_out, plan.out_bytes);
    } while (0);
    slot_release(s);
    return st;
}

b200_status b200_jpeg_encode_coefficients(const b200_jpeg_layout *layout, const int16_t *coefs, int progressive, uint8_t **out, size_t *out_len)
{
    if (!layout || !coefs || !out || !out_len) return make_status(B200_ERR_INVALID_ARGUMENT, "null argument");
    std::string err; JpegGeom g;
    if (!geom_from_layout(layout, g, err)) return make_status(B200_ERR_INVALID_ARGUMENT, err);
    JpegWriteOptions wo; wo.progressive = progressive != 0;
    std::vector<uint8_t> v;
    if (!jpeg_write(g, coefs, wo, nullptr, v, err)) return make_status(B200_ERR_INVALID_ARGUMENT, err);
    return give(v, out, out_len);
}

b200_status b200_jpeg_encode_coefficients_device(const b200_jpeg_layout *layout, const int16_t *coefs, int progressive, uint8_t **out, size_t *out_len)
{
    if (!layout || !coefs || !out || !out_len) return make_status(B200_ERR_INVALID_ARGUMENT, "null argument");
    std::string err; JpegGeom g;
    if (!geom_from_layout(layout, g, err)) return make_status(B200_ERR_INVALID_ARGUMENT, err);
    if (!ensure_runtime(err)) return make_status(B200_ERR_NO_DEVICE, err);
    Slot *s = slot_acquire(runtime_next_device(), err);
    if (!s) return make_status(B200_ERR_CUDA, err);
    b200_status st = ok_status();
    std::vector<uint8_t> v;
    do {
        const size_t bytes = (size_t)g.total_coefs * 2;
        if (!s->ensure(256, bytes, 256, 1 << 14, err)) { st = make_status(B200_ERR_OUT_OF_MEMORY, err); break; }
        memcpy(s->h_out, coefs, bytes);
        if (!slot_upload_out_coefs(s, bytes, err)) { st = make_status(B200_ERR_CUDA, err); break; }
        JpegWriteOptions wo; wo.progressive = progressive != 0;
        if (!slot_gpu_encode(s, g, wo.progressive, err)) { st = make_status(B200_ERR_CUDA, err); break; }
        if (!jpeg_assemble(g, wo, nullptr, s->enc->results.data(), (int)s->enc->results.size(), v, err)) { st = make_status(B200_ERR_INVALID_ARGUMENT, err); break; }
    } while (0);
    slot_release(s);
    if (st.code) return st;
    return give(v, out, out_len);
}

b200_status b200_jpeg_decode_planes(const b200_jpeg_layout *in_layout, const int16_t *in_coefs, uint8_t *planes)
{
    if (!in_layout || !in_coefs || !planes) return make_status(B200_ERR_INVALID_ARGUMENT, "null argument");
    std::string err; JpegGeom gin;
    if (!geom_from_layout(in_layout, gin, err)) return make_status(B200_ERR_INVALID_ARGUMENT, err);
    if (!ensure_runtime(err)) return make_status(B200_ERR_NO_DEVICE, err);
    Slot *s = slot_acquire(runtime_next_device(), err);
    if (!s) return make_status(B200_ERR_CUDA, err);
    b200_status st = ok_status();
    do {
        if (!s->ensure((size_t)gin.total_coefs * 2, 256, 256, 1 << 14, err)) { st = make_status(B200_ERR_OUT_OF_MEMORY, err); break; }
        memcpy(s->h_in, in_coefs, (size_t)gin.total_coefs * 2);
        if (!slot_decode_planes(s, gin, planes, err)) { st = make_status(B200_ERR_CUDA, err); break; }
    } while (0);
    slot_release(s);
    return st;
}

void b200_jpeg_quant_table(int quality, int which, uint16_t out[64]) { jpeg_quant_table(quality, which, out); }

// ---- megabatch --------------------------------------------------------------------------------------------------
struct b200_jpeg_batch { JpegBatch *b; };

b200_status b200_jpeg_batch_create(const b200_jpeg_layout *in_layout, const b200_jpeg_layout *out_layout, int n, b200_jpeg_batch **batch)
{
    if (!in_layout || !out_layout || !batch) return make_status(B200_ERR_INVALID_ARGUMENT, "null argument");
    *batch = nullptr;
    std::string err; JpegGeom gin, gout;
    if (!geom_from_layout(in_layout, gin, err) || !geom_from_layout(out_layout, gout, err)) return make_status(B200_ERR_INVALID_ARGUMENT, err);
    if (!ensure_runtime(err)) return make_status(B200_ERR_NO_DEVICE, err);
    JpegBatch *b = batch_create(gin, gout, n, err);
    if (!b) return make_status(B200_ERR_CUDA, err);
    *batch = new b200_jpeg_batch{b};
    return ok_status();
}
b200_status b200_jpeg_batch_upload(b200_jpeg_batch *b, int index, const int16_t *in_coefs)
{
    std::string err; if (!b || !in_coefs) return make_status(B200_ERR_INVALID_ARGUMENT, "null argument");
    return batch_upload(b->b, index, in_coefs, err) ? ok_status() : make_status(B200_ERR_CUDA, err);
}
b200_status b200_jpeg_batch_run(b200_jpeg_batch *b, void *cuda_stream, int *launches)
{
    std::string err; if (!b) return make_status(B200_ERR_INVALID_ARGUMENT, "null argument");
    return batch_run(b->b, cuda_stream, 0, launches, err) ? ok_status() : make_status(B200_ERR_CUDA, err);
}
b200_status b200_jpeg_batch_download(b200_jpeg_batch *b, int index, int16_t *out_coefs)
{
    std::string err; if (!b || !out_coefs) return make_status(B200_ERR_INVALID_ARGUMENT, "null argument");
    return batch_download(b->b, index, out_coefs, err) ? ok_status() : make_status(B200_ERR_CUDA, err);
}
b200_status b200_jpeg_batch_time(b200_jpeg_batch *b, int which, int iters, float *ms_per_run)
{
    std::string err; if (!b || !ms_per_run) return make_status(B200_ERR_INVALID_ARGUMENT, "null argument");
    return batch_time(b->b, which, iters, ms_per_run, err) ? ok_status() : make_status(B200_ERR_CUDA, err);
}
void b200_jpeg_batch_destroy(b200_jpeg_batch *b) { if (b) { batch_destroy(b->b); delete b; } }

// ---- device-resident full path ------------------------------------------------------------------------------------------
struct b200_jpeg_pipe { JpegPipe *p; };
b200_status b200_jpeg_pipe_create(const uint8_t *const *in, const size_t *in_len, int n, const b200_params *params, int group, b200_jpeg_pipe **pipe)
{
    if (!in || !in_len || !params || !pipe || n <= 0) return make_status(B200_ERR_INVALID_ARGUMENT, "invalid argument");
    *pipe = nullptr;
    std::string err;
    if (!ensure_runtime(err)) return make_status(B200_ERR_NO_DEVICE, err);
    try {
        JpegPipe *P = pipe_create(in, in_len, n, params, group > 0 ? group : 8, err);
        if (!P) return make_status(B200_ERR_INVALID_ARGUMENT, err);
        *pipe = new b200_jpeg_pipe{P};
    } catch (const std::exception &e) { return make_status(B200_ERR_OUT_OF_MEMORY, e.what()); }
    return ok_status();
}
b200_status b200_jpeg_pipe_run(b200_jpeg_pipe *p, void *cuda_stream, int which, int *launches)
{
    std::string err; if (!p) return make_status(B200_ERR_INVALID_ARGUMENT, "null argument");
    return pipe_run(p->p, cuda_stream, which, launches, err) ? ok_status() : make_status(B200_ERR_CUDA, err);
}
b200_status b200_jpeg_pipe_finish(b200_jpeg_pipe *p, size_t *out_sizes, int *not_settled, int *enc_retries)
{
    std::string err; if (!p) return make_status(B200_ERR_INVALID_ARGUMENT, "null argument");
    return pipe_finish(p->p, out_sizes, not_settled, enc_retries, err) ? ok_status() : make_status(B200_ERR_CUDA, err);
}
b200_status b200_jpeg_pipe_fetch(b200_jpeg_pipe *p, int index, uint8_t **out, size_t *out_len)
{
    std::string err; if (!p || !out || !out_len) return make_status(B200_ERR_INVALID_ARGUMENT, "null argument");
    std::vector<uint8_t> v;
    if (!pipe_fetch(p->p, index, v, err)) return make_status(B200_ERR_CUDA, err);
    return give(v, out, out_len);
}
b200_status b200_jpeg_pipe_kernel_times(b200_jpeg_pipe *p, int iters, char *text, size_t cap)
{
    std::string err; if (!p || !text || !cap || iters <= 0) return make_status(B200_ERR_INVALID_ARGUMENT, "invalid argument");
    std::map<std::string, std::pair<double, int>> t;
    if (!pipe_kernel_times(p->p, iters, t, err)) return make_status(B200_ERR_CUDA, err);
    std::string s;
    for (auto &kv : t) { char b[160]; snprintf(b, sizeof b, "%s %.6f %d\n", kv.first.c_str(), kv.second.first, kv.second.second); s += b; }
    if (s.size() + 1 > cap) return make_status(B200_ERR_INVALID_ARGUMENT, "text buffer too small");
    memcpy(text, s.c_str(), s.size() + 1);
    return ok_status();
}
void b200_jpeg_pipe_destroy(b200_jpeg_pipe *p) { if (p) { pipe_destroy(p->p); delete p; } }

// ---- PNG stage entry points ----------------------------------------------------------------------------------------
b200_status b200_png_decode(const uint8_t *in, size_t in_len, b200_png_info *info, uint8_t **raw)
{
    if (!in || !info || !raw) return make_status(B200_ERR_INVALID_ARGUMENT, "null argument");
    std::string err; PngInfo pi; std::vector<uint8_t> r;
    if (!png_decode(in, in_len, false, pi, r, err)) return make_status(err.find("interlace") != std::string::npos ? B200_ERR_UNSUPPORTED : B200_ERR_CORRUPT_INPUT, err);
    info->width = pi.width; info->height = pi.height; info->bit_depth = pi.bit_depth; info->color_type = pi.color_type; info->bpp = pi.bpp; info->row_bytes = pi.row_bytes;
    *raw = (uint8_t *)malloc(r.size() ? r.size() : 1);
    if (!*raw) return make_status(B200_ERR_OUT_OF_MEMORY, "malloc failed");
    memcpy(*raw, r.data(), r.size());
    return ok_status();
}
b200_status b200_png_decode_reduced(const uint8_t *in, size_t in_len, b200_png_info *info, uint8_t **raw, uint8_t *palette_rgba, int *npalette)
{
    if (!in || !info || !raw || !palette_rgba || !npalette) return make_status(B200_ERR_INVALID_ARGUMENT, "null argument");
    std::string err; PngInfo pi; std::vector<uint8_t> r;
    if (!png_decode(in, in_len, false, pi, r, err)) return make_status(err.find("interlace") != std::string::npos ? B200_ERR_UNSUPPORTED : B200_ERR_CORRUPT_INPUT, err);
    *npalette = 0;
    if (png_reduce_palette(pi, r)) {
        *npalette = (int)(pi.plte.size() / 3);
        for (int k = 0; k < *npalette; k++) {
            palette_rgba[4 * k] = pi.plte[3 * k]; palette_rgba[4 * k + 1] = pi.plte[3 * k + 1]; palette_rgba[4 * k + 2] = pi.plte[3 * k + 2];
            palette_rgba[4 * k + 3] = (size_t)k < pi.trns.size() ? pi.trns[k] : 255;
        }
    }
    info->width = pi.width; info->height = pi.height; info->bit_depth = pi.bit_depth; info->color_type = pi.color_type; info->bpp = pi.bpp; info->row_bytes = pi.row_bytes;
    *raw = (uint8_t *)malloc(r.size() ? r.size() : 1);
    if (!*raw) return make_status(B200_ERR_OUT_OF_MEMORY, "malloc failed");
    memcpy(*raw, r.data(), r.size());
    return ok_status();
}
b200_status b200_png_filter(const uint8_t *raw, int h, int row_bytes, int bpp, int strategy, uint8_t *filtered)
{
    if (!raw || !filtered || h <= 0 || row_bytes <= 0 || bpp < 1 || bpp > 8 || strategy < 0 || strategy > 9) return make_status(B200_ERR_INVALID_ARGUMENT, "invalid argument");
    std::string err;
    if (!ensure_runtime(err)) return make_status(B200_ERR_NO_DEVICE, err);
    Slot *s = slot_acquire(runtime_next_device(), err);              // makes the slot's device current for this thread
    if (!s) return make_status(B200_ERR_CUDA, err);
    const bool ok = png_stage_filter(raw, h, row_bytes, bpp, strategy, filtered, err);
    slot_release(s);
    return ok ? ok_status() : make_status(B200_ERR_CUDA, err);
}
b200_status b200_png_lz77(const uint8_t *filtered, size_t n, int bpp, int stride, uint32_t **tokens, size_t *ntokens, uint32_t *hist)
{
    if (!filtered || !n || !tokens || !ntokens || !hist || bpp < 1 || stride < 1) return make_status(B200_ERR_INVALID_ARGUMENT, "invalid argument");
    std::string err;
    if (!ensure_runtime(err)) return make_status(B200_ERR_NO_DEVICE, err);
    Slot *s = slot_acquire(runtime_next_device(), err);
    if (!s) return make_status(B200_ERR_CUDA, err);
    std::vector<uint32_t> t;
    const bool ok = png_stage_lz77(filtered, n, bpp, stride, t, hist, err);
    slot_release(s);
    if (!ok) return make_status(B200_ERR_CUDA, err);
    *tokens = (uint32_t *)malloc(t.size() * 4 + 4);
    if (!*tokens) return make_status(B200_ERR_OUT_OF_MEMORY, "malloc failed");
    memcpy(*tokens, t.data(), t.size() * 4); *ntokens = t.size();
    return ok_status();
}
b200_status b200_png_deflate_tokens(const uint32_t *tokens, size_t ntokens, uint32_t adler, uint8_t **out, size_t *out_len)
{
    if ((!tokens && ntokens) || !out || !out_len) return make_status(B200_ERR_INVALID_ARGUMENT, "null argument");
    std::vector<uint8_t> z; deflate_tokens(tokens, ntokens, adler, z);
    *out = (uint8_t *)malloc(z.size() + 1);
    if (!*out) return make_status(B200_ERR_OUT_OF_MEMORY, "malloc failed");
    memcpy(*out, z.data(), z.size()); *out_len = z.size();
    return ok_status();
}
int b200_webp_alpha_filter(const uint8_t *alpha, int width, int height, uint8_t *filtered)
{
    if (!alpha || !filtered || width < 1 || height < 1) return -1;
    std::vector<uint8_t> f;
    const int k = webp_alpha_choose_filter(alpha, width, height, f);
    memcpy(filtered, k ? f.data() : alpha, (size_t)width * height);
    return k;
}
b200_status b200_webp_alpha_chunk(const uint32_t *tokens, size_t ntokens, int width, int height, int filter, uint8_t **out, size_t *out_len)
{
    if (!tokens || !out || !out_len || width < 1 || height < 1 || width > 16383 || height > 16383 || filter < 0 || filter > 3) return make_status(B200_ERR_INVALID_ARGUMENT, "invalid argument");
    std::vector<uint8_t> a;
    if (!vp8l_alpha_from_tokens(tokens, ntokens, width, height, a, filter)) return make_status(B200_ERR_INVALID_ARGUMENT, "the tokens do not cover the plane");
    *out = (uint8_t *)malloc(a.size() + 1);
    if (!*out) return make_status(B200_ERR_OUT_OF_MEMORY, "malloc failed");
    memcpy(*out, a.data(), a.size()); *out_len = a.size();
    return ok_status();
}
b200_status b200_webp_wrap_alpha(const uint8_t *simple_file, size_t file_len, const uint8_t *alph, size_t alph_len, int width, int height, uint8_t **out, size_t *out_len)
{
    if (!simple_file || !alph || !out || !out_len || width < 1 || height < 1 || width > 16383 || height > 16383) return make_status(B200_ERR_INVALID_ARGUMENT, "invalid argument");
    std::vector<uint8_t> f(simple_file, simple_file + file_len), a(alph, alph + alph_len), o;
    if (!webp_wrap_alpha(f, a, width, height, o)) return make_status(B200_ERR_INVALID_ARGUMENT, "not a simple lossy WebP file");
    *out = (uint8_t *)malloc(o.size() + 1);
    if (!*out) return make_status(B200_ERR_OUT_OF_MEMORY, "malloc failed");
    memcpy(*out, o.data(), o.size()); *out_len = o.size();
    return ok_status();
}
// ---- WebP stage entry points -----------------------------------------------------------------------------------------
b200_status b200_webp_encode_rgb(const uint8_t *rgb, int w, int h, int quality, uint8_t **out, size_t *out_len, int16_t *levels, uint8_t *modes)
{
    if (!rgb || !out || !out_len || w < 1 || h < 1 || w > 16383 || h > 16383) return make_status(B200_ERR_INVALID_ARGUMENT, "invalid argument");
    *out = nullptr; *out_len = 0;
    std::string err;
    if (!ensure_runtime(err)) return make_status(B200_ERR_NO_DEVICE, err);
    Slot *s = slot_acquire(runtime_next_device(), err);
    if (!s) return make_status(B200_ERR_CUDA, err);
    if (!s->webp) s->webp = new WebpDevice();
    std::vector<uint8_t> v;
    const bool ok = s->webp->encode_host_rgb(rgb, w, h, quality, s->stream, v, err, levels, modes);
    slot_release(s);
    return ok ? give(v, out, out_len) : make_status(B200_ERR_CUDA, err);
}
b200_status b200_webp_decode(const uint8_t *in, size_t in_len, int *width, int *height, uint8_t **rgb)
{
    if (!in || !width || !height || !rgb) return make_status(B200_ERR_INVALID_ARGUMENT, "null argument");
    *rgb = nullptr;
    try {
        WebpInfo wi; std::vector<uint8_t> v, a;
        b200_status s = webp_decode_status(in, in_len, wi, v, &a);
        if (s.code) return s;
        *width = wi.width; *height = wi.height;
        size_t n = 0;
        return give(v, rgb, &n);
    } catch (const std::exception &e) { return make_status(B200_ERR_OUT_OF_MEMORY, e.what()); }
}
b200_status b200_webp_decode_rgba(const uint8_t *in, size_t in_len, int *width, int *height, uint8_t **rgb, uint8_t **alpha)
{
    if (!in || !width || !height || !rgb || !alpha) return make_status(B200_ERR_INVALID_ARGUMENT, "null argument");
    *rgb = nullptr; *alpha = nullptr;
    try {
        WebpInfo wi; std::vector<uint8_t> v, a;
        b200_status s = webp_decode_status(in, in_len, wi, v, &a);
        if (s.code) return s;
        *width = wi.width; *height = wi.height;
        size_t n = 0;
        if (!a.empty()) { s = give(a, alpha, &n); if (s.code) return s; }
        s = give(v, rgb, &n);
        if (s.code) { free(*alpha); *alpha = nullptr; }
        return s;
    } catch (const std::exception &e) { return make_status(B200_ERR_OUT_OF_MEMORY, e.what()); }
}
b200_status b200_webp_write_levels(int w, int h, int quality, const int16_t *levels, const uint8_t *modes, uint8_t **out, size_t *out_len)
{
    if (!levels || !modes || !out || !out_len) return make_status(B200_ERR_INVALID_ARGUMENT, "null argument");
    std::vector<uint8_t> v;
    if (!vp8_write_file(w, h, vp8_qindex(quality < 0 ? 0 : quality > 100 ? 100 : quality), levels, modes, v)) return make_status(B200_ERR_INVALID_ARGUMENT, "frame cannot be written as VP8");
    return give(v, out, out_len);
}
unsigned long long b200_webp_d2h_bytes(void) { return webp_d2h_bytes_total(); }
int b200_webp_qindex(int quality, int factors[6])
{
    const int q = vp8_qindex(quality < 0 ? 0 : quality > 100 ? 100 : quality);
    if (factors) vp8_quant_factors(q, factors);
    return q;
}

b200_status b200_png_device_times(const uint8_t *in, size_t in_len, int level, int iters, char *text, size_t cap)
{
    if (!in || !text || !cap || iters <= 0) return make_status(B200_ERR_INVALID_ARGUMENT, "invalid argument");
    std::string err;
    PngInfo info0; PngIdat idat;
    if (!png_parse_chunks(in, in_len, false, info0, idat, err)) return make_status(B200_ERR_CORRUPT_INPUT, err);
    if (!ensure_runtime(err)) return make_status(B200_ERR_NO_DEVICE, err);
    Slot *s = slot_acquire(runtime_next_device(), err);
    if (!s) return make_status(B200_ERR_CUDA, err);
    if (!s->png) s->png = new PngDevice();
    b200_status st = ok_status();
    std::map<std::string, std::pair<double, int>> acc;
    do {
        const size_t nin = (info0.row_bytes + 1) * (size_t)info0.height;
        size_t bcap = 0, got = 0; uint32_t adler = 0;
        uint8_t *buf = s->png->input_buffer(nin, bcap, err);
        if (!buf) { st = make_status(B200_ERR_OUT_OF_MEMORY, err); break; }
        if (!zlib_inflate_to(idat.p, idat.n, buf, bcap, nin, &got, &adler, err) || got < nin) { st = make_status(B200_ERR_CORRUPT_INPUT, err.empty() ? "IDAT too short" : err); break; }
        std::vector<uint8_t> z;
        for (int it = 0; it <= iters; it++) {          // iteration 0 warms buffers up and is not counted
            PngInfo info = info0;
            LaunchTimer lt; lt.begin((cudaStream_t)s->stream);
            tl_launch_timer = it ? &lt : nullptr;
            const bool ok = s->png->compress_filtered(info, got, adler, level < 0 ? 0 : level > 6 ? 6 : level, s->stream, z, nullptr, err);
            tl_launch_timer = nullptr;
            if (!ok) { st = make_status(B200_ERR_CUDA, err); break; }
            if (it) lt.collect(acc);
        }
    } while (0);
    slot_release(s);
    if (st.code) return st;
    std::string out;
    for (auto &kv : acc) { char b[160]; snprintf(b, sizeof b, "%s %.6f %d\n", kv.first.c_str(), kv.second.first / kv.second.second, kv.second.second / iters); out += b; }
    if (out.size() + 1 > cap) return make_status(B200_ERR_INVALID_ARGUMENT, "text buffer too small");
    memcpy(text, out.c_str(), out.size() + 1);
    return ok_status();
}

int b200_png_level_strategies(int level, int *out)
{
    const std::vector<int> v = png_level_strategies(level < 0 ? 0 : level > 6 ? 6 : level);
    if (out) for (size_t i = 0; i < v.size(); i++) out[i] = v[i];
    return (int)v.size();
}

} // extern "C"

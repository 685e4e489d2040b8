// jpeg_device.cu -- see jpeg_device.h.  Device runtime for the JPEG path of caesium::compress_in_memory
// (/root/reference/src/compressor.rs:305): one pool of "slots" per GPU so that the blocking, one-image-per-thread
// callers of the reference's rayon map (compressor.rs:81-83) each get a private stream, pinned staging buffers and
// HBM buffers; images are sharded round-robin over the initialised GPUs (no cross-GPU traffic on this path).
#include <cuda_runtime.h>
#include <malloc.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <vector>
#include <atomic>
#include "jpeg_device.h"
#include "resize_kernels.h"
#include "jpeg_gpuenc.h"
#include "jpeg_gpudec.h"
#include "png_device.h"
#include "webp_device.h"
#include "stream_wait.h"
#include "launch_timer.h"

namespace b200 {

thread_local LaunchTimer *tl_launch_timer = nullptr;

#define CU(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { err = std::string(#expr) + ": " + cudaGetErrorString(e_); return false; } } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

bool plan_image(const JpegGeom &gin, const JpegGeom &gout, ImagePlan &p, std::string &err)
{
    p = ImagePlan();
    if (gin.ncomp != gout.ncomp || gin.width != gout.width || gin.height != gout.height) { err = "geometry mismatch between decode and encode side"; return false; }
    p.in_bytes = (size_t)gin.total_coefs * 2; p.out_bytes = (size_t)gout.total_coefs * 2;
    size_t off_plane = 0, off_full = 0, off_d = 0;
    for (int c = 0; c < gin.ncomp; c++) {
        if (gin.hmax % gin.hs[c] || gin.vmax % gin.vs[c] || gout.hmax % gout.hs[c] || gout.vmax % gout.vs[c]) { err = "fractional sampling ratio unsupported"; return false; }
        int uhx = gin.hmax / gin.hs[c], uvx = gin.vmax / gin.vs[c], dhx = gout.hmax / gout.hs[c], dvx = gout.vmax / gout.vs[c];
        if (uhx == 1 && uvx == 1 && dhx == 1 && dvx == 1) p.path[c] = PATH_FUSED;
        else if (uhx == 2 && uvx == 2 && dhx == 2 && dvx == 2) p.path[c] = PATH_C420;
        else p.path[c] = PATH_GENERIC;
        if (p.path[c] != PATH_FUSED) { p.plane_off[c] = off_plane; off_plane += align_up((size_t)gin.bw[c] * 8 * gin.bh[c] * 8, 256); }
        if (p.path[c] == PATH_GENERIC) {
            p.full_off[c] = off_full; off_full += align_up((size_t)gin.width * gin.height, 256);
            p.dplane_off[c] = off_d; off_d += align_up((size_t)gout.rbw[c] * 8 * gout.rbh[c] * 8, 256);
        }
    }
    p.plane_bytes = off_plane; p.full_bytes = off_full; p.dplane_bytes = off_d;
    return true;
}

void append_image_work(const JpegGeom &gin, const JpegGeom &gout, const ImagePlan &plan,
                       const int16_t *d_in, int16_t *d_out, uint8_t *d_scratch,
                       const uint16_t *d_dq, const QuantDev *d_q, WorkLists &wl)
{
    for (int c = 0; c < gin.ncomp; c++) {
        CompWork w; memset(&w, 0, sizeof(w));
        w.cin = d_in + gin.comp_offset[c]; w.cout = d_out + gout.comp_offset[c];
        w.dq = d_dq + 64 * c; w.q = d_q + gout.tq[c];
        w.bw_in = gin.bw[c]; w.bh_in = gin.bh[c]; w.rbw_in = gin.rbw[c]; w.rbh_in = gin.rbh[c]; w.cw = gin.cw[c]; w.ch = gin.ch[c];
        w.bw_out = gout.bw[c]; w.bh_out = gout.bh[c]; w.rbw_out = gout.rbw[c]; w.rbh_out = gout.rbh[c];
        w.W = gin.width; w.H = gin.height;
        w.pstride = gin.bw[c] * 8; w.fstride = gin.width;
        w.up_hx = gin.hmax / gin.hs[c]; w.up_vx = gin.vmax / gin.vs[c];
        w.dn_hx = gout.hmax / gout.hs[c]; w.dn_vx = gout.vmax / gout.vs[c];
        if (plan.path[c] != PATH_FUSED) w.plane = d_scratch + plan.plane_off[c];
        if (plan.path[c] == PATH_GENERIC) {
            w.full = d_scratch + plan.plane_bytes + plan.full_off[c];
            w.dplane = d_scratch + plan.plane_bytes + plan.full_bytes + plan.dplane_off[c];
        }
        const int t_in = work_tiles(w.rbw_in, w.rbh_in), t_out = work_tiles(w.rbw_out, w.rbh_out);
        switch (plan.path[c]) {
            case PATH_FUSED: wl.fused.push_back(w); wl.max_fused = std::max(wl.max_fused, t_out); break;
            case PATH_C420:
                wl.idct.push_back(w); wl.max_idct = std::max(wl.max_idct, t_in);
                wl.c420.push_back(w); wl.max_c420 = std::max(wl.max_c420, t_out);
                break;
            default:
                wl.idct.push_back(w); wl.max_idct = std::max(wl.max_idct, t_in);
                wl.up.push_back(w); wl.max_up_w = std::max(wl.max_up_w, w.W); wl.max_up_h = std::max(wl.max_up_h, w.H);
                wl.down.push_back(w); wl.max_dn_w = std::max(wl.max_dn_w, w.rbw_out * 8); wl.max_dn_h = std::max(wl.max_dn_h, w.rbh_out * 8);
                wl.fdct.push_back(w); wl.max_fdct = std::max(wl.max_fdct, t_out);
        }
    }
}

size_t flatten_work(const WorkLists &wl, CompWork *h)
{
    size_t n = 0;
    for (const auto *v : {&wl.fused, &wl.idct, &wl.c420, &wl.up, &wl.down, &wl.fdct}) { if (!v->empty()) memcpy(h + n, v->data(), v->size() * sizeof(CompWork)); n += v->size(); }
    return n;
}

int launch_work(const WorkLists &wl, const CompWork *d, void *stream, int which, int *launches)
{
    int rc = 0, n = 0;
    const CompWork *p_fused = d, *p_idct = p_fused + wl.fused.size(), *p_c420 = p_idct + wl.idct.size();
    const CompWork *p_up = p_c420 + wl.c420.size(), *p_down = p_up + wl.up.size(), *p_fdct = p_down + wl.down.size();
    if ((which == 0 || which == 1) && !wl.fused.empty()) { rc = launch_fused_same(p_fused, (int)wl.fused.size(), wl.max_fused, stream); n++; LT_MARK("k_fused_same"); if (rc) return rc; }
    if ((which == 0 || which == 2) && !wl.idct.empty()) { rc = launch_idct_plane(p_idct, (int)wl.idct.size(), wl.max_idct, stream); n++; LT_MARK("k_idct_plane"); if (rc) return rc; }
    if ((which == 0 || which == 3) && !wl.c420.empty()) { rc = launch_chroma420_refdct(p_c420, (int)wl.c420.size(), wl.max_c420, stream); n++; LT_MARK("k_chroma420_refdct"); if (rc) return rc; }
    if (which == 0 || which == 3) {
        if (!wl.up.empty()) { rc = launch_upsample(p_up, (int)wl.up.size(), wl.max_up_w, wl.max_up_h, stream); n++; if (rc) return rc; }
        if (!wl.down.empty()) { rc = launch_downsample(p_down, (int)wl.down.size(), wl.max_dn_w, wl.max_dn_h, stream); n++; if (rc) return rc; }
        if (!wl.fdct.empty()) { rc = launch_fdct_plane(p_fdct, (int)wl.fdct.size(), wl.max_fdct, stream); n++; if (rc) return rc; }
    }
    if (launches) *launches = n;
    return 0;
}

// ================================================================================================================
// runtime: devices and slots
// ================================================================================================================
namespace {
struct DevicePool {
    int ordinal = 0;
    std::mutex mu; std::condition_variable cv;
    std::vector<Slot *> free_slots; int created = 0; int max_slots = 48;
    std::atomic<long long> jobs{0};             // slot acquisitions (a megabatch or a single image each)
};
std::mutex g_mu;
std::vector<DevicePool *> g_devs;
std::atomic<unsigned> g_rr{0};
bool g_inited = false;
}

int runtime_init(int n_gpus, int only_device, std::string &err)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_inited) return (int)g_devs.size();
    int count = 0;
    // How callers wait for their stream is decided in stream_wait.h (hybrid poll-then-sleep by default); B200_SYNC=block
    // additionally asks the driver for blocking synchronisation (no-op if a context already exists, e.g. torch's).
    // One stream per in-flight image, dozens in flight: with the default 8 hardware work queues, streams alias onto the
    // same queue and serialise behind each other.  32 is the maximum; only effective if set before the context exists.
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    const bool blocking = stream_wait_mode() == 1;
    // Every image hands a freshly malloc'ed file (~1 MB) to the caller, thousands per second from several threads.  With
    // glibc's defaults each of those is an mmap + page faults + munmap (or a heap that is trimmed back to the OS as soon as
    // the caller frees), all serialised on the process's mmap lock: measured, that alone cost 45 % of the end-to-end rate.
    // Keep freed memory in the allocator instead.  B200_MALLOPT=0 leaves the process's malloc settings untouched.
    {
        const char *mo = getenv("B200_MALLOPT");
        if (!(mo && !strcmp(mo, "0"))) { mallopt(M_MMAP_THRESHOLD, 32 << 20); mallopt(M_TRIM_THRESHOLD, 2000000000); mallopt(M_TOP_PAD, 64 << 20); }
    }
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) { err = std::string("no CUDA device available (") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") + "); this build has no CPU fallback"; return 0; }
    std::vector<int> ords;
    if (only_device >= 0) { if (only_device >= count) { err = "CUDA device ordinal out of range"; return 0; } ords.push_back(only_device); }
    else { int n = n_gpus <= 0 ? count : std::min(n_gpus, count); for (int i = 0; i < n; i++) ords.push_back(i); }
    for (int o : ords) {
        if (blocking && cudaSetDevice(o) == cudaSuccess) { cudaSetDeviceFlags(cudaDeviceScheduleBlockingSync); cudaGetLastError(); }
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, o) != cudaSuccess) { err = "cudaGetDeviceProperties failed"; return 0; }
        if (prop.major != 10) { err = "device " + std::to_string(o) + " is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + "; this library ships sm_100a kernels only"; for (auto *d : g_devs) delete d; g_devs.clear(); return 0; }
        auto *d = new DevicePool(); d->ordinal = o; g_devs.push_back(d);
    }
    g_inited = true;
    return (int)g_devs.size();
}

static void print_group_trace();
static void drop_graphs(Slot *s);
void runtime_shutdown()
{
    print_group_trace();
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto *d : g_devs) {
        cudaSetDevice(d->ordinal);
        for (Slot *s : d->free_slots) {
            if (s->stream) cudaStreamDestroy((cudaStream_t)s->stream);
            drop_graphs(s);
            cudaFreeHost(s->h_in); cudaFreeHost(s->h_out); cudaFree(s->d_in); cudaFree(s->d_out); cudaFree(s->d_scratch); cudaFreeHost(s->h_par); cudaFree(s->d_par); delete s->enc; delete s->dec; delete s->png; delete s->webp;
            delete s;
        }
        delete d;
    }
    g_devs.clear(); g_inited = false;
}

int runtime_device_count() { std::lock_guard<std::mutex> lk(g_mu); return g_inited ? (int)g_devs.size() : 0; }
long long runtime_device_jobs(int i) { return g_devs.empty() || i < 0 || i >= (int)g_devs.size() ? 0 : g_devs[(size_t)i]->jobs.load(); }
int runtime_device_ordinal(int i) { return g_devs.empty() ? 0 : g_devs[(size_t)i % g_devs.size()]->ordinal; }
int runtime_next_device() { size_t n = g_devs.size(); return n ? (int)(g_rr.fetch_add(1) % n) : 0; }

Slot *slot_acquire(int prefer, std::string &err)
{
    if (g_devs.empty()) { err = "library not initialised (no CUDA device)"; return nullptr; }
    DevicePool *d = g_devs[(size_t)prefer % g_devs.size()];
    d->jobs++;
    std::unique_lock<std::mutex> lk(d->mu);
    for (;;) {
        if (!d->free_slots.empty()) { Slot *s = d->free_slots.back(); d->free_slots.pop_back(); lk.unlock(); cudaSetDevice(d->ordinal); return s; }
        if (d->created < d->max_slots) {
            d->created++; lk.unlock();
            cudaSetDevice(d->ordinal);
            Slot *s = new Slot(); s->dev = (int)((size_t)prefer % g_devs.size());
            cudaStream_t st;
            if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { delete s; err = "cudaStreamCreate failed"; lk.lock(); d->created--; return nullptr; }
            s->stream = st;
            return s;
        }
        d->cv.wait(lk);
    }
}

void slot_release(Slot *s)
{
    if (!s) return;
    DevicePool *d = g_devs[(size_t)s->dev];
    { std::lock_guard<std::mutex> lk(d->mu); d->free_slots.push_back(s); }
    d->cv.notify_one();
}

template <typename T> static bool grow_host(T *&p, size_t &cap, size_t need, std::string &err)
{
    if (need <= cap) return true;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    size_t want = align_up(need + need / 8, 1 << 16);
    void *q = nullptr;
    cudaError_t e = cudaHostAlloc(&q, want, cudaHostAllocDefault);
    if (e != cudaSuccess) { err = std::string("cudaHostAlloc: ") + cudaGetErrorString(e); return false; }
    p = (T *)q; cap = want; return true;
}
template <typename T> static bool grow_dev(T *&p, size_t &cap, size_t need, std::string &err)
{
    if (need <= cap) return true;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = align_up(need + need / 8, 1 << 16);
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, want);
    if (e != cudaSuccess) { err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return false; }
    p = (T *)q; cap = want; return true;
}

bool Slot::ensure_device(size_t in_bytes, size_t out_bytes, size_t scratch_bytes, size_t par_bytes, std::string &err)
{   // megabatch path: coefficients never visit the host, so only the HBM side (and the small parameter block) grows
    return grow_dev(d_in, d_in_cap, in_bytes, err) && grow_dev(d_out, d_out_cap, out_bytes, err) &&
           grow_dev(d_scratch, d_scratch_cap, std::max<size_t>(scratch_bytes, 256), err) &&
           grow_host(h_par, par_cap, par_bytes, err) && grow_dev(d_par, d_par_cap, par_bytes, err);
}

bool Slot::ensure(size_t in_bytes, size_t out_bytes, size_t scratch_bytes, size_t par_bytes, std::string &err)
{
    return grow_host(h_in, h_in_cap, in_bytes, err) && grow_host(h_out, h_out_cap, out_bytes, err) &&
           grow_dev(d_in, d_in_cap, in_bytes, err) && grow_dev(d_out, d_out_cap, out_bytes, err) &&
           grow_dev(d_scratch, d_scratch_cap, std::max<size_t>(scratch_bytes, 256), err) &&
           grow_host(h_par, par_cap, par_bytes, err) && grow_dev(d_par, d_par_cap, par_bytes, err);
}

// parameter block layout: QuantDev q[4] | uint16 dq[4][64] | CompWork work[...]
static const size_t PAR_Q = 0, PAR_DQ = sizeof(QuantDev) * 4, PAR_WORK = PAR_DQ + sizeof(uint16_t) * 256;
static size_t par_bytes_for(size_t nwork) { return PAR_WORK + nwork * sizeof(CompWork); }

static void fill_tables(uint8_t *h_par, const JpegGeom &gin, const JpegGeom *gout)
{
    QuantDev *q = reinterpret_cast<QuantDev *>(h_par + PAR_Q);
    uint16_t *dq = reinterpret_cast<uint16_t *>(h_par + PAR_DQ);
    if (gout) for (int t = 0; t < 4; t++) if (gout->qt_present[t]) make_quant_dev(gout->qt[t], &q[t]);
    for (int c = 0; c < gin.ncomp; c++) memcpy(dq + 64 * c, gin.qt[gin.tq[c]], 128);
}

bool slot_transform(Slot *s, const JpegGeom &gin, const JpegGeom &gout, std::string &err, bool download, bool upload)
{
    ImagePlan plan;
    if (!plan_image(gin, gout, plan, err)) return false;
    cudaStream_t st = (cudaStream_t)s->stream;
    const size_t pbytes = par_bytes_for(4 * 6);
    if (!s->ensure(plan.in_bytes, plan.out_bytes, plan.scratch_bytes(), pbytes, err)) return false;
    fill_tables(s->h_par, gin, &gout);
    WorkLists wl;
    append_image_work(gin, gout, plan, s->d_in, s->d_out, s->d_scratch,
                      reinterpret_cast<const uint16_t *>(s->d_par + PAR_DQ), reinterpret_cast<const QuantDev *>(s->d_par + PAR_Q), wl);
    size_t nw = flatten_work(wl, reinterpret_cast<CompWork *>(s->h_par + PAR_WORK));
    CU(cudaMemcpyAsync(s->d_par, s->h_par, par_bytes_for(nw), cudaMemcpyHostToDevice, st));
    if (upload) CU(cudaMemcpyAsync(s->d_in, s->h_in, plan.in_bytes, cudaMemcpyHostToDevice, st));
    int rc = launch_work(wl, reinterpret_cast<const CompWork *>(s->d_par + PAR_WORK), st, 0, nullptr);
    if (rc) { err = std::string("kernel launch: ") + cudaGetErrorString((cudaError_t)rc); return false; }
    if (!download) return true;          // the coefficients stay in HBM for the device entropy encoder
    CU(cudaMemcpyAsync(s->h_out, s->d_out, plan.out_bytes, cudaMemcpyDeviceToHost, st));
    CU(stream_wait(st));
    return true;
}

// ---- megabatch: K same-shaped images per launch sequence (b200_compress_batch) -------------------------------------
// Buffers of image k live at d_in + k * in_stride etc.; the input coefficients are already in HBM (device decoder) and
// the output coefficients stay there (device encoder).  Parameter block: QuantDev q[4] | dq[K][4][64] | CompWork[].
bool slot_group_layout(Slot *s, const JpegGeom &gin, const JpegGeom &gout, int K, GroupLayout &L, std::string &err)
{
    ImagePlan plan;
    if (!plan_image(gin, gout, plan, err)) return false;
    L.K = K;
    L.in_stride = align_up(plan.in_bytes, 256); L.out_stride = align_up(plan.out_bytes, 256); L.scratch_stride = align_up(std::max<size_t>(plan.scratch_bytes(), 256), 256);
    const size_t par = align_up(sizeof(QuantDev) * 4, 256) + align_up(sizeof(uint16_t) * 256 * K, 256) + sizeof(CompWork) * (size_t)K * 4 * 6 + 256;
    return s->ensure_device(L.in_stride * K, L.out_stride * K, L.scratch_stride * K, par, err);
}

// host half: quantiser constants, per-image dequantisation tables and the work descriptors into the slot's pinned parameter block
bool slot_transform_group_prepare(Slot *s, const JpegGeom *const *gins, const JpegGeom &gout, const GroupLayout &L, std::string &err)
{
    const int K = L.K;
    const size_t o_q = 0, o_dq = align_up(sizeof(QuantDev) * 4, 256), o_work = o_dq + align_up(sizeof(uint16_t) * 256 * K, 256);
    QuantDev *q = reinterpret_cast<QuantDev *>(s->h_par + o_q);
    for (int t = 0; t < 4; t++) if (gout.qt_present[t]) make_quant_dev(gout.qt[t], &q[t]);
    WorkLists &wl = s->group_wl; wl.clear();
    for (int k = 0; k < K; k++) {
        const JpegGeom &gin = *gins[k];
        ImagePlan plan;
        if (!plan_image(gin, gout, plan, err)) return false;
        uint16_t *dq = reinterpret_cast<uint16_t *>(s->h_par + o_dq) + 256 * k;
        for (int c = 0; c < gin.ncomp; c++) memcpy(dq + 64 * c, gin.qt[gin.tq[c]], 128);
        append_image_work(gin, gout, plan, reinterpret_cast<const int16_t *>(reinterpret_cast<const uint8_t *>(s->d_in) + L.in_stride * k),
                          reinterpret_cast<int16_t *>(reinterpret_cast<uint8_t *>(s->d_out) + L.out_stride * k), s->d_scratch + L.scratch_stride * k,
                          reinterpret_cast<const uint16_t *>(s->d_par + o_dq) + 256 * k, reinterpret_cast<const QuantDev *>(s->d_par + o_q), wl);
    }
    const size_t nw = flatten_work(wl, reinterpret_cast<CompWork *>(s->h_par + o_work));
    s->group_par_bytes = o_work + nw * sizeof(CompWork); s->group_work_off = o_work;
    return true;
}
// stream half: parameter block up, the transform kernels
bool slot_transform_group_enqueue(Slot *s, std::string &err)
{
    cudaStream_t st = (cudaStream_t)s->stream;
    CU(cudaMemcpyAsync(s->d_par, s->h_par, s->group_par_bytes, cudaMemcpyHostToDevice, st));
    int rc = launch_work(s->group_wl, reinterpret_cast<const CompWork *>(s->d_par + s->group_work_off), st, 0, nullptr);
    if (rc) { err = std::string("kernel launch: ") + cudaGetErrorString((cudaError_t)rc); return false; }
    return true;
}
bool slot_transform_group(Slot *s, const JpegGeom *const *gins, const JpegGeom &gout, const GroupLayout &L, std::string &err)
{
    return slot_transform_group_prepare(s, gins, gout, L, err) && slot_transform_group_enqueue(s, err);
}

// (Measured and dropped: issuing the decode passes on a highest-priority stream so that their small latency-bound grids cut in
// front of other megabatches' 24k-CTA encoder grids LOWERED the batch rate by 20 % -- 4,430 -> 3,550 images/s with 8 group
// workers, 4,690 -> 4,030 with 16.  Everything of a megabatch stays on the slot's one stream.)
//
// One megabatch, front to back, with ONE host wait that leaves the GPU idle (the last): entropy decode (fixed number of rounds,
// flags read afterwards), transform, entropy encode (its sizes come back while the emit kernels run).  Round 1 waited three
// times per megabatch with an empty stream behind each wait.
// B200_TRACE: host wall-clock per megabatch, summed: staging (copy into pinned + tables), launching, the two waits
static std::atomic<long long> g_grp_ns[5];
static std::atomic<long long> g_grp_n{0};
static const bool g_grp_trace = getenv("B200_TRACE") != nullptr;
static void print_group_trace()
{
    const long long n = g_grp_n.load();
    if (!g_grp_trace || !n) return;
    static const char *names[] = {"stage inputs (memcpy to pinned, tables, H2D enqueue)", "enqueue decode + transform + encode (launch overhead)", "wait: sizes (GPU still busy)", "wait: final (after D2H enqueue)", "whole megabatch on the host"};
    fprintf(stderr, "[b200 trace] %lld megabatches; mean host ms per megabatch:\n", n);
    for (int i = 0; i < 5; i++) fprintf(stderr, "[b200 trace]   %-58s %8.3f\n", names[i], g_grp_ns[i].load() / 1e6 / (double)n);
}

// The launch sequence of a megabatch is ~70 driver calls (kernels, CUB scans, memsets, small copies).  Sixteen worker threads
// issuing them concurrently spend more time in the driver's process-wide lock than the kernels take to run (B200_TRACE showed
// 1.8 ms of pure enqueue time per megabatch, 24 us per call), and that -- not the GPU -- capped the C-ABI rate at 85 % of the
// device-resident rate.  The sequence is the same from megabatch to megabatch (sizes are high-water marks, buffers are reused), so
// it is captured once per slot as two CUDA graphs -- everything up to the scan sizes, and the bit-packing / stuffing half -- and
// replayed: three driver calls per megabatch.  A changed signature (other geometry, a grown buffer, a new high-water mark)
// re-captures.  B200_GRAPHS=0 issues the calls directly.
static bool graphs_enabled() { static const bool on = [] { const char *e = getenv("B200_GRAPHS"); return !(e && !strcmp(e, "0")); }(); return on; }
static void drop_graphs(Slot *s)
{
    if (s->graph_front) { cudaGraphExecDestroy((cudaGraphExec_t)s->graph_front); s->graph_front = nullptr; }
    if (s->graph_back) { cudaGraphExecDestroy((cudaGraphExec_t)s->graph_back); s->graph_back = nullptr; }
    s->graph_sig = 0;
}
template <class Fn> static bool capture_graph(cudaStream_t st, void *&exec_out, Fn fn, std::string &err)
{
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return false; }
    const bool ok = fn();
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(st, &g);
    if (!ok || e != cudaSuccess || !g) { if (g) cudaGraphDestroy(g); cudaGetLastError(); if (err.empty()) err = "graph capture failed"; return false; }
    cudaGraphExec_t ex = nullptr;
    const cudaError_t ei = cudaGraphInstantiate(&ex, g, 0);
    cudaGraphDestroy(g);
    if (ei != cudaSuccess) { cudaGetLastError(); err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ei); return false; }
    exec_out = ex;
    return true;
}

bool slot_run_group(Slot *s, std::vector<GpuDecoder::Item> &items, const JpegGeom *const *gins, const JpegGeom &gout, const GroupLayout &L, bool progressive,
                    bool lossless, std::string &err)
{
    if (!s->dec) s->dec = new GpuDecoder();
    if (!s->enc) s->enc = new GpuEncoder();
    cudaStream_t st = (cudaStream_t)s->stream;
    const auto t0 = std::chrono::steady_clock::now();
    // ---- host half of all three stages (pinned staging, descriptors, plans); nothing touches the stream yet
    if (!s->dec->prepare(items, st, err)) return false;
    if (!lossless && !slot_transform_group_prepare(s, gins, gout, L, err)) return false;
    std::vector<int16_t *> bases((size_t)L.K);
    for (int k = 0; k < L.K; k++) bases[k] = lossless ? reinterpret_cast<int16_t *>(reinterpret_cast<uint8_t *>(s->d_in) + L.in_stride * k)
                                                      : reinterpret_cast<int16_t *>(reinterpret_cast<uint8_t *>(s->d_out) + L.out_stride * k);
    // a re-encode at lower quality (or a transcode with optimal tables) does not grow: the inputs' entropy-coded size sizes the output buffers
    if (!s->enc->prepare(gout, progressive, bases.data(), L.K, st, s->dec->raw_bytes(), err)) return false;
    const auto t1 = std::chrono::steady_clock::now();
    auto front = [&]() { return s->dec->upload(st, err) && s->dec->enqueue(st, err) && (lossless || slot_transform_group_enqueue(s, err)) &&
                                s->enc->upload(st, err) && s->enc->enqueue_front(st, true, err) && s->enc->enqueue_sizes(st, err); };
    auto back = [&]() { return s->enc->enqueue_back(st, err); };
    bool launched = false;
    if (graphs_enabled() && !s->graphs_broken) {
        unsigned long long sig = s->dec->signature() * 1099511628211ull ^ s->enc->signature();
        sig = (sig ^ (unsigned long long)(uintptr_t)s->d_par ^ ((unsigned long long)s->group_par_bytes << 20) ^ (lossless ? 0x9e3779b97f4a7c15ull : 0)) * 1099511628211ull + (unsigned long long)L.K;
        if (!s->graph_front || s->graph_sig != sig) {
            drop_graphs(s);
            std::string gerr;
            if (capture_graph(st, s->graph_front, front, gerr) && capture_graph(st, s->graph_back, back, gerr)) s->graph_sig = sig;
            else { drop_graphs(s); s->graphs_broken = true; }      // capture not possible here: issue the calls directly from now on
        }
        if (s->graph_front && s->graph_back) {
            if (cudaGraphLaunch((cudaGraphExec_t)s->graph_front, st) != cudaSuccess || !s->enc->mark_sizes(st, err) ||
                cudaGraphLaunch((cudaGraphExec_t)s->graph_back, st) != cudaSuccess) { err = std::string("cudaGraphLaunch: ") + cudaGetErrorString(cudaGetLastError()); return false; }
            launched = true;
        }
    }
    if (!launched && !(front() && s->enc->mark_sizes(st, err) && back())) return false;
    const auto t2 = std::chrono::steady_clock::now();
    if (!s->enc->finish(st, true, err)) return false;
    s->dec->finish(items);
    if (g_grp_trace) {
        const auto t3 = std::chrono::steady_clock::now();
        auto ns = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return (long long)std::chrono::duration_cast<std::chrono::nanoseconds>(b - a).count(); };
        g_grp_ns[0] += ns(t0, t1); g_grp_ns[1] += ns(t1, t2); g_grp_ns[2] += (long long)(s->enc->wait_sizes_ms * 1e6); g_grp_ns[3] += (long long)(s->enc->wait_final_ms * 1e6);
        g_grp_ns[4] += ns(t0, t3); g_grp_n++;
    }
    return true;
}

bool slot_decode_group(Slot *s, std::vector<GpuDecoder::Item> &items, std::string &err)
{
    if (!s->dec) s->dec = new GpuDecoder();
    return s->dec->decode(items, s->stream, err);
}

bool slot_encode_group(Slot *s, const JpegGeom &gout, bool progressive, const GroupLayout &L, std::string &err, bool from_input)
{
    if (!s->enc) s->enc = new GpuEncoder();
    std::vector<int16_t *> bases((size_t)L.K);
    for (int k = 0; k < L.K; k++) bases[k] = from_input ? reinterpret_cast<int16_t *>(reinterpret_cast<uint8_t *>(s->d_in) + L.in_stride * k)
                                                        : reinterpret_cast<int16_t *>(reinterpret_cast<uint8_t *>(s->d_out) + L.out_stride * k);
    return s->enc->encode(gout, progressive, bases.data(), L.K, s->stream, true, err);
}

bool slot_download_coefs(Slot *s, size_t out_bytes, std::string &err)
{
    cudaStream_t st = (cudaStream_t)s->stream;
    CU(cudaMemcpyAsync(s->h_out, s->d_out, out_bytes, cudaMemcpyDeviceToHost, st));
    CU(stream_wait(st));
    return true;
}

int slot_gpu_decode(Slot *s, const JpegReader &rd, const JpegReader::DeviceScan &ds, std::string &err)
{   // 0 = coefficients are in s->d_in, 1 = not converged (decode on the host instead), 2 = failure
    if (!s->dec) s->dec = new GpuDecoder();
    std::vector<GpuDecoder::Item> items(1);
    items[0].rd = &rd; items[0].ds = &ds; items[0].d_coefs = s->d_in; items[0].result = GpuDecoder::FAILED;
    if (!s->dec->decode(items, s->stream, err)) return 2;
    return (int)items[0].result;
}

bool slot_upload_out_coefs(Slot *s, size_t bytes, std::string &err)
{
    cudaStream_t st = (cudaStream_t)s->stream;
    CU(cudaMemcpyAsync(s->d_out, s->h_out, bytes, cudaMemcpyHostToDevice, st));
    return true;
}

bool slot_gpu_encode(Slot *s, const JpegGeom &gout, bool progressive, std::string &err, bool from_input)
{
    if (!s->enc) s->enc = new GpuEncoder();
    int16_t *base = from_input ? s->d_in : s->d_out;
    return s->enc->encode(gout, progressive, &base, 1, s->stream, true, err);
}

bool slot_gpu_encode_sizes(Slot *s, const JpegGeom &gout, bool progressive, std::string &err)
{
    if (!s->enc) s->enc = new GpuEncoder();
    int16_t *base = s->d_out;
    return s->enc->prepare(gout, progressive, &base, 1, s->stream, 0, err) && s->enc->upload(s->stream, err) && s->enc->enqueue(s->stream, true, err) && s->enc->finish(s->stream, false, err);
}
bool slot_gpu_fetch(Slot *s, std::string &err) { return s->enc && s->enc->finish(s->stream, true, err); }

bool slot_fetch_planes(Slot *s, uint8_t *const *d_planes, int nplanes, size_t n, uint8_t *host, std::string &err)
{
    cudaStream_t st = (cudaStream_t)s->stream;
    for (int c = 0; c < nplanes; c++) CU(cudaMemcpyAsync(host + (size_t)c * n, d_planes[c], n, cudaMemcpyDeviceToHost, st));
    CU(stream_wait(st));
    return true;
}

bool slot_decode_planes(Slot *s, const JpegGeom &gin, uint8_t *planes, std::string &err)
{
    // route every component through idct (+ upsample) by planning against a 4:4:4 output of the same size
    JpegGeom gout = gin;
    for (int c = 0; c < gin.ncomp; c++) { gout.hs[c] = gout.vs[c] = 1; }
    gout.finalize();
    cudaStream_t st = (cudaStream_t)s->stream;
    ImagePlan plan; plan = ImagePlan();
    size_t off_plane = 0, off_full = 0;
    for (int c = 0; c < gin.ncomp; c++) {
        if (gin.hmax % gin.hs[c] || gin.vmax % gin.vs[c]) { err = "fractional sampling ratio unsupported"; return false; }
        plan.path[c] = PATH_GENERIC;
        plan.plane_off[c] = off_plane; off_plane += align_up((size_t)gin.bw[c] * 8 * gin.bh[c] * 8, 256);
        plan.full_off[c] = off_full; off_full += align_up((size_t)gin.width * gin.height, 256);
    }
    plan.plane_bytes = off_plane; plan.full_bytes = off_full; plan.dplane_bytes = 0;
    plan.in_bytes = (size_t)gin.total_coefs * 2; plan.out_bytes = 256;
    const size_t pbytes = par_bytes_for(4 * 6);
    if (!s->ensure(plan.in_bytes, std::max(plan.out_bytes, plan.full_bytes), plan.scratch_bytes(), pbytes, err)) return false;
    fill_tables(s->h_par, gin, nullptr);
    WorkLists wl;
    append_image_work(gin, gout, plan, s->d_in, s->d_out, s->d_scratch,
                      reinterpret_cast<const uint16_t *>(s->d_par + PAR_DQ), reinterpret_cast<const QuantDev *>(s->d_par + PAR_Q), wl);
    wl.down.clear(); wl.fdct.clear();
    size_t nw = flatten_work(wl, reinterpret_cast<CompWork *>(s->h_par + PAR_WORK));
    CU(cudaMemcpyAsync(s->d_par, s->h_par, par_bytes_for(nw), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(s->d_in, s->h_in, plan.in_bytes, cudaMemcpyHostToDevice, st));
    int rc = launch_work(wl, reinterpret_cast<const CompWork *>(s->d_par + PAR_WORK), st, 0, nullptr);
    if (rc) { err = std::string("kernel launch: ") + cudaGetErrorString((cudaError_t)rc); return false; }
    uint8_t *h = reinterpret_cast<uint8_t *>(s->h_out);
    for (int c = 0; c < gin.ncomp; c++)
        CU(cudaMemcpyAsync(h + (size_t)c * gin.width * gin.height, s->d_scratch + plan.plane_bytes + plan.full_off[c], (size_t)gin.width * gin.height, cudaMemcpyDeviceToHost, st));
    CU(stream_wait(st));
    memcpy(planes, h, (size_t)gin.ncomp * gin.width * gin.height);
    return true;
}

// ---- resize path: decode -> (YCbCr->RGB) -> Lanczos3 -> (RGB->YCbCr) -> encode side --------------------------------
bool slot_transform_resized(Slot *s, const JpegGeom &gin, const JpegGeom &gout, std::string &err, bool download, bool upload, uint8_t **rgb_out,
                            const uint8_t *host_rgb)
{
    const int W = gin.width, H = gin.height, NW = gout.width, NH = gout.height, nc = gin.ncomp;
    if (gout.ncomp != nc) { err = "component count mismatch"; return false; }
    cudaStream_t st = (cudaStream_t)s->stream;
    ResizeAxis av, ah;
    make_resize_axis(H, NH, av);
    make_resize_axis(W, NW, ah);
    // scratch layout
    size_t off = 0, plane_off[4], full_off[4], rz_off[4], dpl_off[4];
    for (int c = 0; c < nc; c++) {
        if (gin.hmax % gin.hs[c] || gin.vmax % gin.vs[c] || gout.hmax % gout.hs[c] || gout.vmax % gout.vs[c]) { err = "fractional sampling ratio unsupported"; return false; }
        plane_off[c] = off; off += align_up((size_t)gin.bw[c] * 8 * gin.bh[c] * 8, 256);
    }
    for (int c = 0; c < nc; c++) { full_off[c] = off; off += align_up((size_t)W * H, 256); }
    for (int c = 0; c < nc; c++) { rz_off[c] = off; off += align_up((size_t)NW * NH, 256); }
    for (int c = 0; c < nc; c++) { dpl_off[c] = off; off += align_up((size_t)gout.rbw[c] * 8 * gout.rbh[c] * 8, 256); }
    const size_t tmp_off = off; off += align_up((size_t)NH * W * sizeof(float), 256);
    // parameter block: tables | work | axis tables
    const size_t nwork = (size_t)nc * 4;
    size_t p_axis = align_up(par_bytes_for(nwork), 256);
    const size_t lv = p_axis, cv = lv + align_up(sizeof(int) * NH, 256), wv = cv + align_up(sizeof(int) * NH, 256);
    const size_t lh = wv + align_up(sizeof(float) * av.weights.size(), 256), chh = lh + align_up(sizeof(int) * NW, 256), wh = chh + align_up(sizeof(int) * NW, 256);
    const size_t pbytes = wh + align_up(sizeof(float) * ah.weights.size(), 256);
    const size_t in_bytes = (size_t)gin.total_coefs * 2, out_bytes = (size_t)gout.total_coefs * 2;
    if (!s->ensure(in_bytes, out_bytes, off, pbytes, err)) return false;
    fill_tables(s->h_par, gin, &gout);
    memcpy(s->h_par + lv, av.left.data(), sizeof(int) * NH); memcpy(s->h_par + cv, av.count.data(), sizeof(int) * NH);
    memcpy(s->h_par + wv, av.weights.data(), sizeof(float) * av.weights.size());
    memcpy(s->h_par + lh, ah.left.data(), sizeof(int) * NW); memcpy(s->h_par + chh, ah.count.data(), sizeof(int) * NW);
    memcpy(s->h_par + wh, ah.weights.data(), sizeof(float) * ah.weights.size());
    WorkLists wl;
    const uint16_t *d_dq = reinterpret_cast<const uint16_t *>(s->d_par + PAR_DQ);
    const QuantDev *d_q = reinterpret_cast<const QuantDev *>(s->d_par + PAR_Q);
    for (int c = 0; c < nc; c++) {
        CompWork d; memset(&d, 0, sizeof(d));   // decode side
        d.cin = s->d_in + gin.comp_offset[c]; d.dq = d_dq + 64 * c; d.q = d_q;
        d.bw_in = gin.bw[c]; d.bh_in = gin.bh[c]; d.rbw_in = gin.rbw[c]; d.rbh_in = gin.rbh[c]; d.cw = gin.cw[c]; d.ch = gin.ch[c];
        d.W = W; d.H = H; d.pstride = gin.bw[c] * 8; d.fstride = W;
        d.up_hx = gin.hmax / gin.hs[c]; d.up_vx = gin.vmax / gin.vs[c]; d.dn_hx = d.dn_vx = 1;
        d.plane = s->d_scratch + plane_off[c]; d.full = s->d_scratch + full_off[c];
        wl.idct.push_back(d); wl.max_idct = std::max(wl.max_idct, work_tiles(d.rbw_in, d.rbh_in));
        wl.up.push_back(d); wl.max_up_w = W; wl.max_up_h = H;
        CompWork e; memset(&e, 0, sizeof(e));   // encode side
        e.cout = s->d_out + gout.comp_offset[c]; e.dq = d_dq; e.q = d_q + gout.tq[c];
        e.bw_out = gout.bw[c]; e.bh_out = gout.bh[c]; e.rbw_out = gout.rbw[c]; e.rbh_out = gout.rbh[c];
        e.W = NW; e.H = NH; e.fstride = NW; e.full = s->d_scratch + rz_off[c]; e.dplane = s->d_scratch + dpl_off[c];
        e.dn_hx = gout.hmax / gout.hs[c]; e.dn_vx = gout.vmax / gout.vs[c]; e.up_hx = e.up_vx = 1;
        wl.down.push_back(e); wl.max_dn_w = std::max(wl.max_dn_w, e.rbw_out * 8); wl.max_dn_h = std::max(wl.max_dn_h, e.rbh_out * 8);
        wl.fdct.push_back(e); wl.max_fdct = std::max(wl.max_fdct, work_tiles(e.rbw_out, e.rbh_out));
    }
    size_t nw_ = flatten_work(wl, reinterpret_cast<CompWork *>(s->h_par + PAR_WORK));
    (void)nw_;
    CU(cudaMemcpyAsync(s->d_par, s->h_par, pbytes, cudaMemcpyHostToDevice, st));
    if (upload && !host_rgb) CU(cudaMemcpyAsync(s->d_in, s->h_in, in_bytes, cudaMemcpyHostToDevice, st));
    const CompWork *dw = reinterpret_cast<const CompWork *>(s->d_par + PAR_WORK);
    const CompWork *p_idct = dw, *p_up = p_idct + wl.idct.size(), *p_down = p_up + wl.up.size(), *p_fdct = p_down + wl.down.size();
    auto chk = [&](int rc, const char *what) { if (rc) { err = std::string(what) + ": " + cudaGetErrorString((cudaError_t)rc); return false; } return true; };
    uint8_t *full[3] = {s->d_scratch + full_off[0], nc == 3 ? s->d_scratch + full_off[1] : nullptr, nc == 3 ? s->d_scratch + full_off[2] : nullptr};
    uint8_t *rz[3] = {s->d_scratch + rz_off[0], nc == 3 ? s->d_scratch + rz_off[1] : nullptr, nc == 3 ? s->d_scratch + rz_off[2] : nullptr};
    if (host_rgb) {   // samples that never were a JPEG (PNG source): planar RGB (or one grey plane) straight into the full-resolution planes
        for (int c = 0; c < nc; c++) CU(cudaMemcpyAsync(full[c], host_rgb + (size_t)c * W * H, (size_t)W * H, cudaMemcpyHostToDevice, st));
    } else {
        if (!chk(launch_idct_plane(p_idct, nc, wl.max_idct, st), "idct")) return false;
        if (!chk(launch_upsample(p_up, nc, W, H, st), "upsample")) return false;
        if (nc == 3 && !chk(launch_ycc_to_rgb(full[0], full[1], full[2], (size_t)W * H, st), "ycc_to_rgb")) return false;
    }
    float *tmp = reinterpret_cast<float *>(s->d_scratch + tmp_off);
    for (int c = 0; c < nc; c++) {
        if (NW == W && NH == H) {   // imageops::resize copies when the dimensions are unchanged
            CU(cudaMemcpyAsync(rz[c], full[c], (size_t)W * H, cudaMemcpyDeviceToDevice, st));
            continue;
        }
        if (!chk(launch_resize_v(full[c], W, H, W, tmp, NH, reinterpret_cast<const int *>(s->d_par + lv), reinterpret_cast<const int *>(s->d_par + cv),
                                 reinterpret_cast<const float *>(s->d_par + wv), av.cap, st), "resize_v")) return false;
        if (!chk(launch_resize_h(tmp, W, rz[c], NW, NH, NW, reinterpret_cast<const int *>(s->d_par + lh), reinterpret_cast<const int *>(s->d_par + chh),
                                 reinterpret_cast<const float *>(s->d_par + wh), ah.cap, st), "resize_h")) return false;
    }
    if (rgb_out) { rgb_out[0] = rz[0]; rgb_out[1] = nc == 3 ? rz[1] : rz[0]; rgb_out[2] = nc == 3 ? rz[2] : rz[0]; return true; }
    if (nc == 3 && !chk(launch_rgb_to_ycc(rz[0], rz[1], rz[2], (size_t)NW * NH, st), "rgb_to_ycc")) return false;
    if (!chk(launch_downsample(p_down, nc, wl.max_dn_w, wl.max_dn_h, st), "downsample")) return false;
    if (!chk(launch_fdct_plane(p_fdct, nc, wl.max_fdct, st), "fdct")) return false;
    if (!download) return true;
    CU(cudaMemcpyAsync(s->h_out, s->d_out, out_bytes, cudaMemcpyDeviceToHost, st));
    CU(stream_wait(st));
    return true;
}

// ================================================================================================================
// megabatch
// ================================================================================================================
JpegBatch *batch_create(const JpegGeom &gin, const JpegGeom &gout, int n, std::string &err)
{
    if (g_devs.empty()) { err = "library not initialised (no CUDA device)"; return nullptr; }
    if (n <= 0) { err = "empty batch"; return nullptr; }
    auto *b = new JpegBatch();
    b->dev = 0; b->n = n; b->gin = gin; b->gout = gout;
    auto fail = [&](const std::string &m) -> JpegBatch * { err = m; batch_destroy(b); return nullptr; };
    if (!plan_image(gin, gout, b->plan, err)) { batch_destroy(b); return nullptr; }
    cudaSetDevice(g_devs[0]->ordinal);
    const size_t in_b = align_up(b->plan.in_bytes, 256), out_b = align_up(b->plan.out_bytes, 256), sc_b = align_up(std::max<size_t>(b->plan.scratch_bytes(), 256), 256);
    void *p = nullptr;
    if (cudaMalloc(&p, in_b * n) != cudaSuccess) return fail("cudaMalloc (batch input) failed"); b->d_in = (int16_t *)p;
    if (cudaMalloc(&p, out_b * n) != cudaSuccess) return fail("cudaMalloc (batch output) failed"); b->d_out = (int16_t *)p;
    if (cudaMalloc(&p, sc_b * n) != cudaSuccess) return fail("cudaMalloc (batch scratch) failed"); b->d_scratch = (uint8_t *)p;
    cudaStream_t st; if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return fail("cudaStreamCreate failed"); b->stream = st;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); b->ev0 = e0; b->ev1 = e1;
    size_t nwork_max = (size_t)n * 4 * 6;
    size_t pbytes = par_bytes_for(nwork_max);
    std::vector<uint8_t> hpar(pbytes);
    if (cudaMalloc(&p, pbytes) != cudaSuccess) return fail("cudaMalloc (batch params) failed"); b->d_par = (uint8_t *)p;
    fill_tables(hpar.data(), gin, &gout);
    for (int i = 0; i < n; i++)
        append_image_work(gin, gout, b->plan, (const int16_t *)((const uint8_t *)b->d_in + in_b * i), (int16_t *)((uint8_t *)b->d_out + out_b * i), b->d_scratch + sc_b * i,
                          reinterpret_cast<const uint16_t *>(b->d_par + PAR_DQ), reinterpret_cast<const QuantDev *>(b->d_par + PAR_Q), b->wl);
    size_t nw = flatten_work(b->wl, reinterpret_cast<CompWork *>(hpar.data() + PAR_WORK));
    if (cudaMemcpy(b->d_par, hpar.data(), par_bytes_for(nw), cudaMemcpyHostToDevice) != cudaSuccess) return fail("cudaMemcpy (batch params) failed");
    b->d_work = reinterpret_cast<const CompWork *>(b->d_par + PAR_WORK);
    cudaMemset(b->d_in, 0, in_b * n);
    return b;
}

bool batch_upload(JpegBatch *b, int idx, const int16_t *coefs, std::string &err)
{
    if (!b || idx < 0 || idx >= b->n) { err = "bad batch index"; return false; }
    cudaSetDevice(g_devs[0]->ordinal);
    const size_t in_b = align_up(b->plan.in_bytes, 256);
    CU(cudaMemcpy((uint8_t *)b->d_in + in_b * idx, coefs, b->plan.in_bytes, cudaMemcpyHostToDevice));
    return true;
}

bool batch_run(JpegBatch *b, void *stream, int which, int *launches, std::string &err)
{
    if (!b) { err = "null batch"; return false; }
    cudaSetDevice(g_devs[0]->ordinal);
    int rc = launch_work(b->wl, b->d_work, stream ? stream : b->stream, which, launches);
    if (rc) { err = std::string("kernel launch: ") + cudaGetErrorString((cudaError_t)rc); return false; }
    return true;
}

bool batch_download(JpegBatch *b, int idx, int16_t *coefs, std::string &err)
{
    if (!b || idx < 0 || idx >= b->n) { err = "bad batch index"; return false; }
    cudaSetDevice(g_devs[0]->ordinal);
    CU(stream_wait((cudaStream_t)b->stream));
    CU(cudaDeviceSynchronize());
    const size_t out_b = align_up(b->plan.out_bytes, 256);
    CU(cudaMemcpy(coefs, (uint8_t *)b->d_out + out_b * idx, b->plan.out_bytes, cudaMemcpyDeviceToHost));
    return true;
}

bool batch_time(JpegBatch *b, int which, int iters, float *ms, std::string &err)
{
    if (!b || iters <= 0) { err = "bad arguments"; return false; }
    cudaSetDevice(g_devs[0]->ordinal);
    cudaStream_t st = (cudaStream_t)b->stream;
    CU(stream_wait(st));
    CU(cudaEventRecord((cudaEvent_t)b->ev0, st));
    for (int i = 0; i < iters; i++) {
        int rc = launch_work(b->wl, b->d_work, st, which, nullptr);
        if (rc) { err = std::string("kernel launch: ") + cudaGetErrorString((cudaError_t)rc); return false; }
    }
    CU(cudaEventRecord((cudaEvent_t)b->ev1, st));
    CU(cudaEventSynchronize((cudaEvent_t)b->ev1));
    float t = 0; CU(cudaEventElapsedTime(&t, (cudaEvent_t)b->ev0, (cudaEvent_t)b->ev1));
    *ms = t / iters;
    return true;
}

void batch_destroy(JpegBatch *b)
{
    if (!b) return;
    if (!g_devs.empty()) cudaSetDevice(g_devs[0]->ordinal);
    cudaFree(b->d_in); cudaFree(b->d_out); cudaFree(b->d_scratch); cudaFree(b->d_par);
    if (b->stream) cudaStreamDestroy((cudaStream_t)b->stream);
    if (b->ev0) cudaEventDestroy((cudaEvent_t)b->ev0);
    if (b->ev1) cudaEventDestroy((cudaEvent_t)b->ev1);
    delete b;
}

} // namespace b200

// compressor.cpp -- see compressor.h.  Every function cites the reference lines it mirrors.
#include "compressor.h"
#include <algorithm>
#include <atomic>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <future>
#include <memory>
#include <thread>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>

namespace fs = std::filesystem;

namespace b200clt {

static bool read_file_to_vec(const std::string &path, std::vector<uint8_t> &buf)
{   // compressor.rs:600-605
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) return false;
    std::streamsize n = f.tellg();
    if (n < 0) return false;
    buf.resize((size_t)n);
    f.seekg(0);
    return n == 0 || (bool)f.read(reinterpret_cast<char *>(buf.data()), n);
}

static uint32_t map_supported_formats(OutputFormat f)
{   // compressor.rs:589-598
    switch (f) {
        case OutputFormat::Jpeg: return B200_FMT_JPEG;
        case OutputFormat::Png: return B200_FMT_PNG;
        case OutputFormat::Gif: return B200_FMT_GIF;
        case OutputFormat::Webp: return B200_FMT_WEBP;
        case OutputFormat::Tiff: return B200_FMT_TIFF;
        default: return B200_FMT_UNKNOWN;
    }
}

static unsigned be16(const uint8_t *p) { return (p[0] << 8) | p[1]; }
static unsigned be32(const uint8_t *p) { return ((unsigned)p[0] << 24) | (p[1] << 16) | (p[2] << 8) | p[3]; }
static unsigned le16(const uint8_t *p) { return p[0] | (p[1] << 8); }
static unsigned le24(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16); }

static int jpeg_exif_orientation(const std::vector<uint8_t> &b)
{   // kamadak-exif: Reader::read_from_container + get_field(Orientation, PRIMARY)  (compressor.rs:546-553)
    size_t i = 2;
    while (i + 4 <= b.size() && b[i] == 0xFF) {
        unsigned m = b[i + 1];
        if (m == 0xDA || m == 0xD9) break;
        size_t L = be16(&b[i + 2]);
        if (L < 2 || i + 2 + L > b.size()) break;
        if (m == 0xE1 && L >= 16 && !memcmp(&b[i + 4], "Exif\0\0", 6)) {
            const uint8_t *t = &b[i + 10]; size_t tl = L - 8;
            bool le = t[0] == 'I';
            auto u16 = [&](size_t o) -> unsigned { return o + 2 <= tl ? (le ? le16(t + o) : be16(t + o)) : 0; };
            auto u32 = [&](size_t o) -> unsigned { return o + 4 <= tl ? (le ? (le16(t + o) | (le16(t + o + 2) << 16)) : be32(t + o)) : 0; };
            size_t ifd = u32(4);
            unsigned n = ifd + 2 <= tl ? u16(ifd) : 0;
            for (unsigned k = 0; k < n; k++) { size_t e = ifd + 2 + 12 * (size_t)k; if (e + 12 > tl) break; if (u16(e) == 0x0112) return (int)u16(e + 8); }
            return 1;
        }
        i += 2 + L;
    }
    return 1;
}

bool get_real_resolution(const std::vector<uint8_t> &b, bool keep_metadata, size_t &w, size_t &h, std::string &err)
{   // imagesize::blob_size for the four discoverable formats, then the EXIF swap of compressor.rs:538-561
    uint32_t fmt = b200_sniff_format(b.data(), b.size());
    w = h = 0;
    if (fmt == B200_FMT_JPEG) {
        size_t i = 2;
        while (i + 4 <= b.size()) {
            if (b[i] != 0xFF) { i++; continue; }
            unsigned m = b[i + 1];
            if (m == 0xFF) { i++; continue; }
            if (m == 0x01 || (m >= 0xD0 && m <= 0xD8)) { i += 2; continue; }
            size_t L = be16(&b[i + 2]);
            if (m >= 0xC0 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) { if (i + 9 <= b.size()) { h = be16(&b[i + 5]); w = be16(&b[i + 7]); } break; }
            if (L < 2) break;
            i += 2 + L;
        }
    } else if (fmt == B200_FMT_PNG) {
        if (b.size() >= 24) { w = be32(&b[16]); h = be32(&b[20]); }
    } else if (fmt == B200_FMT_GIF) {
        if (b.size() >= 10) { w = le16(&b[6]); h = le16(&b[8]); }
    } else if (fmt == B200_FMT_WEBP) {
        if (b.size() >= 30) {
            if (!memcmp(&b[12], "VP8 ", 4)) { w = le16(&b[26]) & 0x3FFF; h = le16(&b[28]) & 0x3FFF; }
            else if (!memcmp(&b[12], "VP8L", 4)) { unsigned v = le16(&b[21]) | (le16(&b[23]) << 16); w = (v & 0x3FFF) + 1; h = ((v >> 14) & 0x3FFF) + 1; }
            else if (!memcmp(&b[12], "VP8X", 4)) { w = le24(&b[24]) + 1; h = le24(&b[27]) + 1; }
        }
    }
    if (!w || !h) { err = "Could not determine image size"; return false; }
    if (fmt == B200_FMT_JPEG && keep_metadata) {
        int o = jpeg_exif_orientation(b);
        if (o >= 5 && o <= 8) std::swap(w, h);
    }
    return true;
}

static bool is_resize_needed(const CompressionOptions &o) { return o.width || o.height || o.long_edge || o.short_edge; }   // :186-188

bool build_compression_parameters(const CompressionOptions &options, const std::vector<uint8_t> &buffer, b200_params &p, std::string &err)
{   // compressor.rs:411-446
    b200_params_default(&p);
    uint32_t quality = options.quality.value_or(80);
    p.jpeg_quality = p.png_quality = p.webp_quality = quality;
    p.gif_quality = options.lossless ? 100 : (quality == 0 ? 1 : quality);
    p.jpeg_preserve_icc = !options.strip_icc;
    p.jpeg_optimize = p.png_optimize = p.webp_lossless = options.lossless;
    p.keep_metadata = options.exif;
    p.jpeg_chroma_subsampling = options.jpeg_chroma_subsampling;
    p.jpeg_progressive = !options.jpeg_baseline;
    p.png_optimization_level = options.png_opt_level;
    p.png_force_zopfli = options.zopfli;
    if (is_resize_needed(options)) {   // build_resize_parameters, :503-536
        size_t width = 0, height = 0;
        if (!get_real_resolution(buffer, options.exif, width, height, err)) return false;
        if (options.width || options.height) { p.width = options.width.value_or(0); p.height = options.height.value_or(0); }
        else if (options.long_edge) { if (width > height) p.width = *options.long_edge; else p.height = *options.long_edge; }
        else if (options.short_edge) { if (width < height) p.width = *options.short_edge; else p.height = *options.short_edge; }
        if (options.no_upscale && (p.width >= width || p.height >= height)) p.width = p.height = 0;
    }
    return true;
}

static std::string status_message(b200_status &s)
{
    std::string m = s.message ? s.message : "";
    b200_free(s.message); s.message = nullptr;
    return m;
}

static bool run_codec(const std::vector<uint8_t> &buf, b200_params &params, const CompressionOptions &options, CompressionResult &result, std::vector<uint8_t> &out);

bool perform_image_compression(const std::string &input_file, const CompressionOptions &options, CompressionResult &result, std::vector<uint8_t> &out)
{   // compressor.rs:266-315
    std::vector<uint8_t> buf;
    if (!read_file_to_vec(input_file, buf)) { result.message = "Error reading input file"; return false; }
    b200_params params; std::string err;
    if (!build_compression_parameters(options, buf, params, err)) { result.message = "Error building compression parameters: " + err; return false; }
    return run_codec(buf, params, options, result, out);
}

static std::string absolute_path(const std::string &p) { std::error_code ec; auto a = fs::absolute(p, ec); return ec ? p : a.lexically_normal().string(); }

bool compute_output_full_path(const std::string &output_directory, const std::string &input_file_path, const std::string &base_directory,
                              bool keep_structure, const std::string &suffix, OutputFormat format, bool same_folder_as_input,
                              std::string &out_dir, std::string &out_name)
{   // compressor.rs:448-501
    fs::path in(input_file_path);
    std::string ext;
    switch (format) {
        case OutputFormat::Jpeg: ext = "jpg"; break;
        case OutputFormat::Png: ext = "png"; break;
        case OutputFormat::Webp: ext = "webp"; break;
        case OutputFormat::Tiff: ext = "tiff"; break;
        case OutputFormat::Gif: ext = "gif"; break;
        default: ext = in.extension().string(); if (!ext.empty() && ext[0] == '.') ext = ext.substr(1);
    }
    out_name = in.stem().string() + suffix;
    if (!ext.empty()) out_name += "." + ext;
    if (!keep_structure) { out_dir = output_directory; return true; }
    fs::path parent = in.parent_path();
    if (parent.empty()) parent = ".";
    std::error_code ec;
    if (!fs::exists(parent, ec)) return false;
    std::string aparent = absolute_path(parent.string());
    if (same_folder_as_input) { out_dir = aparent; return true; }
    if (!base_directory.empty()) {
        fs::path rel = fs::path(aparent).lexically_relative(fs::path(base_directory));
        std::string r = rel.string();
        if (r.empty() || r.rfind("..", 0) == 0) return false;          // strip_prefix failed
        out_dir = r == "." ? output_directory : (fs::path(output_directory) / rel).string();
    } else {
        std::string pre = aparent; pre.erase(std::remove(pre.begin(), pre.end(), ':'), pre.end());
        while (!pre.empty() && pre[0] == '/') pre.erase(0, 1);
        out_dir = (fs::path(output_directory) / pre).string();
    }
    return true;
}

static bool setup_output_path(const std::string &input_file, const CompressionOptions &options, CompressionResult &result, bool dry_run, std::string &full)
{   // compressor.rs:190-241
    std::string outdir;
    if (options.same_folder_as_input) { fs::path p = fs::path(input_file).parent_path(); outdir = p.empty() ? "." : p.string(); }
    else if (options.output_folder) outdir = *options.output_folder;
    else { result.message = "Error getting output directory"; return false; }
    std::string dir, name;
    if (!compute_output_full_path(outdir, input_file, options.base_path, options.keep_structure, options.suffix.value_or(""), options.format,
                                  options.same_folder_as_input || outdir == options.base_path, dir, name)) return false;
    full = (fs::path(dir) / name).string();
    if (dry_run) return true;
    std::error_code ec;
    if (!fs::exists(dir, ec) && !fs::create_directories(dir, ec) && ec) { result.message = "Error creating output directory"; return false; }
    return true;
}

static std::string bytesize_str(uint64_t b)
{   // bytesize::ByteSize Display (SI): "1.5 KB"
    static const char *u[] = {"B", "KB", "MB", "GB", "TB", "PB"};
    if (b < 1000) return std::to_string(b) + " B";
    double v = (double)b; int i = 0;
    while (v >= 1000.0 && i < 5) { v /= 1000.0; i++; }
    char s[64]; snprintf(s, sizeof s, "%.1f %s", v, u[i]); return s;
}

// perform_compression (compressor.rs:103-184) in two halves around the codec call, so that start_compression can hand the codec
// calls of many files to b200_compress_batch at once.  prepare_compression = everything up to the file read and the parameter
// mapping (the first lines of perform_image_compression, :266-285); finish_compression = the size policies and the write.
struct Prepared {
    CompressionResult r;
    struct stat st {};
    uint64_t original = 0;
    std::vector<uint8_t> buf;          // file bytes, read when the file reaches the codec
    b200_params params {};
    bool need_codec = false;           // false: r is final (skipped, dry run, or failed before the codec)
};

static void prepare_compression(const std::string &input_file, const CompressionOptions &options, bool dry_run, Prepared &pr)
{
    CompressionResult &r = pr.r; r = CompressionResult(); r.original_path = input_file; pr.need_codec = false;
    if (stat(input_file.c_str(), &pr.st) != 0) { r.message = "Error reading file metadata"; return; }
    const uint64_t original = (uint64_t)pr.st.st_size;
    if (original > MAX_FILE_SIZE) { r.message = "File exceeds 500Mb, skipping."; r.status = CompressionStatus::Skipped; return; }
    r.original_size = original; pr.original = original;
    std::string outp;
    if (!setup_output_path(input_file, options, r, dry_run, outp)) { r.message = "Error setting up output path"; return; }
    r.output_path = outp;
    std::error_code ec;
    if (options.overwrite_policy == OverwritePolicy::Never && fs::exists(outp, ec)) {   // :243-257
        r.status = CompressionStatus::Skipped; r.compressed_size = original; r.message = "File already exists, skipped due overwrite policy"; return;
    }
    if (dry_run) { r.status = CompressionStatus::Success; r.compressed_size = original; return; }
    if (!read_file_to_vec(input_file, pr.buf)) { r.message = "Error reading input file"; return; }
    std::string err;
    if (!build_compression_parameters(options, pr.buf, pr.params, err)) { r.message = "Error building compression parameters: " + err; return; }
    pr.need_codec = true;
}

static void finish_compression(Prepared &pr, const CompressionOptions &options, const uint8_t *img, size_t img_len)
{
    CompressionResult &r = pr.r;
    const uint64_t original = pr.original, outsz = img_len;
    const std::string &outp = r.output_path;
    std::error_code ec;
    if (options.min_savings && original != 0) {   // :317-362
        uint64_t actual = original > outsz ? original - outsz : 0;
        const MinSavingsThreshold &t = *options.min_savings;
        char m[160];
        if (t.is_percentage) {
            double sp = (double)actual / (double)original * 100.0;
            if (sp < t.percent) { snprintf(m, sizeof m, "Insufficient savings: %.2f%% < %.2f%%, skipped", sp, t.percent); r.status = CompressionStatus::Skipped; r.compressed_size = original; r.message = m; return; }
        } else if (actual < t.bytes) {
            r.status = CompressionStatus::Skipped; r.compressed_size = original;
            r.message = "Insufficient savings: " + bytesize_str(actual) + " < " + bytesize_str(t.bytes) + ", skipped"; return;
        }
    }
    if (options.overwrite_policy == OverwritePolicy::Bigger && fs::exists(outp, ec)) {   // :364-389
        auto existing = fs::file_size(outp, ec);
        if (ec) r.message = "Error reading existing file metadata";
        else if (existing <= outsz) { r.status = CompressionStatus::Skipped; r.compressed_size = original; r.message = "File already exists, skipped due overwrite policy"; return; }
    }
    {   // write_compressed_file, :391-409
        int fd = open(outp.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
        if (fd < 0) { r.message = "Error creating output file"; return; }
        size_t off = 0;
        while (off < img_len) { ssize_t n = write(fd, img + off, img_len - off); if (n <= 0) { close(fd); r.message = "Error writing output file"; return; } off += (size_t)n; }
        if (options.keep_dates) {   // preserve_file_times, :563-588
            struct timespec ts[2] = {pr.st.st_atim, pr.st.st_mtim};
            if (futimens(fd, ts) != 0) { close(fd); r.message = "Error preserving file times"; return; }
        }
        close(fd);
    }
    r.status = CompressionStatus::Success; r.compressed_size = outsz;
}

// the 4-way dispatch of perform_image_compression (:287-306) on bytes that are already in memory
static bool run_codec(const std::vector<uint8_t> &buf, b200_params &params, const CompressionOptions &options, CompressionResult &result, std::vector<uint8_t> &out)
{
    uint8_t *o = nullptr; size_t ol = 0;
    b200_status st;
    if (options.max_size && options.format != OutputFormat::Original) {
        uint8_t *c = nullptr; size_t cl = 0;
        st = b200_convert_in_memory(buf.data(), buf.size(), &params, map_supported_formats(options.format), &c, &cl);
        if (st.code) { b200_free(st.message); return false; }   // `.ok()?` at :294 swallows the error: empty message, status stays Error
        st = b200_compress_to_size_in_memory(c, cl, &params, *options.max_size, 1, &o, &ol);
        b200_free(c);
    } else if (options.max_size) {
        st = b200_compress_to_size_in_memory(buf.data(), buf.size(), &params, *options.max_size, 1, &o, &ol);
    } else if (options.format != OutputFormat::Original) {
        st = b200_convert_in_memory(buf.data(), buf.size(), &params, map_supported_formats(options.format), &o, &ol);
    } else {
        st = b200_compress_in_memory(buf.data(), buf.size(), &params, &o, &ol);
    }
    if (st.code) { result.message = "Error compressing file: " + status_message(st); return false; }
    out.assign(o, o + ol);
    b200_free(o);
    return true;
}

CompressionResult perform_compression(const std::string &input_file, const CompressionOptions &options, bool dry_run)
{   // compressor.rs:103-184
    Prepared pr;
    prepare_compression(input_file, options, dry_run, pr);
    if (!pr.need_codec) return pr.r;
    std::vector<uint8_t> img;
    if (!run_codec(pr.buf, pr.params, options, pr.r, img)) return pr.r;
    finish_compression(pr, options, img.data(), img.size());
    return pr.r;
}

static int usable_cores()
{
    unsigned hc = std::thread::hardware_concurrency(); if (!hc) hc = 1;
    std::ifstream f("/sys/fs/cgroup/cpu.max"); std::string a; long long period = 0;
    if (f && (f >> a >> period) && a != "max" && period > 0) { long long q = atoll(a.c_str()); if (q > 0) { unsigned n = (unsigned)((q + period - 1) / period); if (n >= 1 && n < hc) hc = n; } }
    return (int)hc;
}

static bool same_params(const b200_params &a, const b200_params &b)      // field by field: the struct has padding
{
    return a.keep_metadata == b.keep_metadata && a.jpeg_quality == b.jpeg_quality && a.jpeg_chroma_subsampling == b.jpeg_chroma_subsampling &&
           a.jpeg_progressive == b.jpeg_progressive && a.jpeg_optimize == b.jpeg_optimize && a.jpeg_preserve_icc == b.jpeg_preserve_icc &&
           a.png_quality == b.png_quality && a.png_optimization_level == b.png_optimization_level && a.png_force_zopfli == b.png_force_zopfli &&
           a.png_optimize == b.png_optimize && a.gif_quality == b.gif_quality && a.webp_quality == b.webp_quality && a.webp_lossless == b.webp_lossless &&
           a.width == b.width && a.height == b.height;
}

// Runs fn(0..n-1) on `threads` threads (the calling one included).
template <class Fn> static void parallel_for(size_t n, int threads, Fn fn)
{
    std::atomic<size_t> next{0};
    auto worker = [&] { for (;;) { const size_t i = next.fetch_add(1); if (i >= n) break; fn(i); } };
    const int nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(1, threads), n));
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(worker);
    worker();
    for (auto &t : th) t.join();
}

std::vector<CompressionResult> start_compression(const std::vector<std::string> &files, const CompressionOptions &options, bool dry_run, int threads)
{   // compressor.rs:74-101: par_iter().map(perform_compression).collect() -- results in input order
    std::vector<CompressionResult> results(files.size());
    int n = threads <= 0 ? usable_cores() : std::min(threads, usable_cores());   // main.rs:287-292 get_parallelism_count
    n = std::max(1, std::min<int>(n, (int)std::max<size_t>(1, files.size())));
    // --max-size and --format go through calls that have no batch form: plain per-file map, as the reference does it
    if (dry_run || options.max_size || options.format != OutputFormat::Original) {
        parallel_for(files.size(), n, [&](size_t i) { results[i] = perform_compression(files[i], options, dry_run); });
        return results;
    }
    // The plain compress call (:305) of a whole chunk of files goes to b200_compress_batch -- the batch form of this very map,
    // which packs same-shaped JPEGs into megabatches on the GPU.  Per file the policy code before and after the codec is the
    // same as in perform_compression; chunks bound the bytes held in memory (inputs + outputs of at most 256 files / 1 GiB).
    constexpr size_t CHUNK_FILES = 256; constexpr uint64_t CHUNK_BYTES = 512ull << 20;
    std::vector<std::pair<size_t, size_t>> chunks;
    for (size_t begin = 0; begin < files.size();) {
        size_t end = begin; uint64_t bytes = 0;
        while (end < files.size() && end - begin < CHUNK_FILES) {
            struct stat st; const uint64_t sz = stat(files[end].c_str(), &st) == 0 ? (uint64_t)st.st_size : 0;
            if (end > begin && bytes + sz > CHUNK_BYTES) break;
            bytes += sz; end++;
        }
        chunks.emplace_back(begin, end); begin = end;
    }
    auto prepare_chunk = [&](size_t c) {
        auto pr = std::make_unique<std::vector<Prepared>>(chunks[c].second - chunks[c].first);
        parallel_for(pr->size(), n, [&](size_t k) { prepare_compression(files[chunks[c].first + k], options, false, (*pr)[k]); });
        return pr;
    };
    // read-ahead: while chunk c is in the codec and being written out, chunk c + 1 is stat'ed, read and parameterised
    std::unique_ptr<std::vector<Prepared>> cur = chunks.empty() ? nullptr : prepare_chunk(0);
    for (size_t c = 0; c < chunks.size(); c++) {
        std::future<std::unique_ptr<std::vector<Prepared>>> ahead;
        if (c + 1 < chunks.size()) ahead = std::async(std::launch::async, prepare_chunk, c + 1);
        std::vector<Prepared> &pr = *cur;
        const size_t begin = chunks[c].first, m = pr.size();
        // files whose parameter blocks are identical share one batch call (resize requests can differ per file)
        std::vector<size_t> todo;
        for (size_t k = 0; k < m; k++) if (pr[k].need_codec) todo.push_back(k);
        std::vector<uint8_t *> outs(m, nullptr); std::vector<size_t> out_len(m, 0); std::vector<char> coded(m, 0);
        std::vector<char> taken(m, 0);
        for (size_t a = 0; a < todo.size(); a++) {
            if (taken[todo[a]]) continue;
            std::vector<size_t> grp;
            for (size_t b = a; b < todo.size(); b++) if (!taken[todo[b]] && same_params(pr[todo[a]].params, pr[todo[b]].params)) { grp.push_back(todo[b]); taken[todo[b]] = 1; }
            const int g = (int)grp.size();
            std::vector<const uint8_t *> ins((size_t)g); std::vector<size_t> lens((size_t)g); std::vector<uint8_t *> go((size_t)g, nullptr); std::vector<size_t> gl((size_t)g, 0);
            std::vector<b200_status> sts((size_t)g);
            for (int j = 0; j < g; j++) { ins[(size_t)j] = pr[grp[(size_t)j]].buf.data(); lens[(size_t)j] = pr[grp[(size_t)j]].buf.size(); sts[(size_t)j].code = 0; sts[(size_t)j].message = nullptr; }
            b200_compress_batch(ins.data(), lens.data(), g, &pr[grp[0]].params, n, go.data(), gl.data(), sts.data());
            for (int j = 0; j < g; j++) {
                const size_t k = grp[(size_t)j];
                if (sts[(size_t)j].code) { pr[k].r.message = "Error compressing file: " + status_message(sts[(size_t)j]); if (go[(size_t)j]) b200_free(go[(size_t)j]); }
                else { outs[k] = go[(size_t)j]; out_len[k] = gl[(size_t)j]; coded[k] = 1; }
            }
        }
        parallel_for(m, n, [&](size_t k) {
            if (coded[k]) { std::vector<uint8_t>().swap(pr[k].buf); finish_compression(pr[k], options, outs[k], out_len[k]); b200_free(outs[k]); }
            results[begin + k] = std::move(pr[k].r);
        });
        if (ahead.valid()) cur = ahead.get();
    }
    return results;
}

std::vector<std::string> scan_files(const std::vector<std::string> &args, bool recursive, std::string &base_path)
{   // scan_files.rs:50-92; the base path is the deepest common ancestor of the files' parent directories (:107-150)
    std::vector<std::string> files;
    auto valid = [](const fs::path &p) {
        uint8_t b[16]; std::ifstream f(p, std::ios::binary);
        if (!f.read(reinterpret_cast<char *>(b), 16)) return false;
        uint32_t t = b200_sniff_format(b, 16);
        return t == B200_FMT_JPEG || t == B200_FMT_PNG || t == B200_FMT_WEBP || t == B200_FMT_GIF;
    };
    std::error_code ec;
    for (const auto &a : args) {
        fs::path in(a);
        if (fs::is_directory(in, ec)) {
            std::vector<std::string> found;
            if (recursive) { for (auto it = fs::recursive_directory_iterator(in, fs::directory_options::skip_permission_denied, ec); it != fs::recursive_directory_iterator(); it.increment(ec)) if (fs::is_regular_file(it->symlink_status(ec)) && valid(it->path())) found.push_back(it->path().string()); }       // WalkDir follow_links(false)
            else { for (auto it = fs::directory_iterator(in, ec); it != fs::directory_iterator(); it.increment(ec)) if (fs::is_regular_file(it->symlink_status(ec)) && valid(it->path())) found.push_back(it->path().string()); }
            std::sort(found.begin(), found.end());
            files.insert(files.end(), found.begin(), found.end());
        } else if (fs::is_regular_file(in, ec) && valid(in)) files.push_back(a);
    }
    fs::path base; bool first = true;
    for (const auto &f : files) {
        fs::path parent = fs::path(absolute_path(f)).parent_path();
        if (first) { base = parent; first = false; continue; }
        fs::path common; auto bi = base.begin(); auto pi = parent.begin();
        for (; bi != base.end() && pi != parent.end() && *bi == *pi; ++bi, ++pi) common /= *bi;
        base = common;
    }
    base_path = base.string();
    return files;
}

bool parse_byte_size(const std::string &val, uint64_t &out)
{   // bytesize 2.x FromStr: number + optional unit (B, KB/MB/GB/TB = 1000^n, KiB/MiB/... = 1024^n, single letter K/M/G = 1000^n)
    size_t i = 0; std::string s = val;
    while (i < s.size() && (isdigit((unsigned char)s[i]) || s[i] == '.')) i++;
    if (i == 0) return false;
    double num = atof(s.substr(0, i).c_str());
    std::string u = s.substr(i);
    u.erase(std::remove_if(u.begin(), u.end(), [](char c) { return isspace((unsigned char)c); }), u.end());
    std::string ul; for (char c : u) ul += (char)tolower((unsigned char)c);
    double mul = 1;
    if (ul.empty() || ul == "b") mul = 1;
    else {
        static const char *pre = "kmgtp"; const char *p = strchr(pre, ul[0]);
        if (!p) return false;
        int e = (int)(p - pre) + 1;
        bool bin = ul.size() >= 2 && ul[1] == 'i';
        std::string rest = ul.substr(bin ? 2 : 1);
        if (!(rest.empty() || rest == "b")) return false;
        mul = std::pow(bin ? 1024.0 : 1000.0, e);
    }
    out = (uint64_t)(num * mul);
    return true;
}

bool parse_min_savings(const std::string &val, MinSavingsThreshold &out, std::string &err)
{   // options.rs:232-257
    std::string t = val;
    t.erase(0, t.find_first_not_of(" \t")); if (!t.empty()) t.erase(t.find_last_not_of(" \t") + 1);
    if (t.empty()) { err = "Value cannot be empty. Use percentage (e.g., '10%'), size with unit (e.g., '100KB', '1MB'), or plain number as bytes"; return false; }
    if (t.back() == '%') {
        char *end = nullptr; std::string n = t.substr(0, t.size() - 1);
        double p = strtod(n.c_str(), &end);
        if (end == n.c_str() || (*end && !isspace((unsigned char)*end))) { err = "Invalid percentage value: '" + n + "'"; return false; }
        if (p < 0.0 || p > 100.0) { err = "Percentage must be between 0 and 100, got " + n; return false; }
        out = {true, p, 0}; return true;
    }
    uint64_t b = 0;
    if (!parse_byte_size(t, b)) { err = "Invalid size format: '" + val + "'. Use percentage (e.g., '10%'), size with unit (e.g., '100KB', '1MB'), or plain number as bytes"; return false; }
    out = {false, 0.0, b}; return true;
}

} // namespace b200clt

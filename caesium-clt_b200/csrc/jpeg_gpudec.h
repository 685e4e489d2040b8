// jpeg_gpudec.h -- device-side JPEG entropy DECODER for baseline single-scan files (see jpeg_gpudec_core.h).
#pragma once
#include <cstdint>
#include <cstddef>
#include <string>
#include <vector>
#include "jpeg_gpudec_core.h"
#include "jpeg_host.h"

namespace b200 {

class GpuDecoder {
public:
    GpuDecoder() = default;
    ~GpuDecoder();
    GpuDecoder(const GpuDecoder &) = delete;
    GpuDecoder &operator=(const GpuDecoder &) = delete;
    enum Result { OK = 0, NOT_CONVERGED = 1, FAILED = 2 };
    // Decode the scan described by `ds` of the file behind `rd` straight into d_coefs (device, geometry rd.geom(), zigzag,
    // fully overwritten).  NOT_CONVERGED: the self-synchronisation did not settle within the round budget (degenerate
    // periodic streams) -- the caller decodes on the host instead.  Blocking on `stream` (one short host sync per group
    // of rounds).
    Result decode(const JpegReader &rd, const JpegReader::DeviceScan &ds, int16_t *d_coefs, void *stream, std::string &err);
    int rounds_used = 0;
    static constexpr int SUBSEQ_BITS = 1024, ROUNDS_PER_GROUP = 6, MAX_ROUNDS = 48;
private:
    uint8_t *h_raw = nullptr; size_t cap_hraw = 0;          // pinned staging of the entropy-coded segment
    uint8_t *d_raw = nullptr, *d_stream = nullptr; size_t cap_raw = 0, cap_stream = 0;
    uint32_t *d_cnt = nullptr, *d_off = nullptr; size_t cap_cnt = 0, cap_off = 0;
    gd::DecState *d_A = nullptr, *d_B = nullptr; size_t cap_A = 0, cap_B = 0;
    uint8_t *d_chgA = nullptr, *d_chgB = nullptr; size_t cap_chgA = 0, cap_chgB = 0;
    uint32_t *d_nblk = nullptr, *d_first = nullptr; size_t cap_nblk = 0, cap_first = 0;
    int32_t *d_dc = nullptr, *d_dcs = nullptr; size_t cap_dc = 0, cap_dcs = 0;
    uint8_t *d_par = nullptr, *h_par = nullptr; size_t cap_par = 0, cap_hpar = 0;   // Geometry | Scan | tables | round flags
    uint8_t *d_temp = nullptr; size_t cap_temp = 0;
};

} // namespace b200

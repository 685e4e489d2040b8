// host_copy.h -- copies between the caller's buffers and pinned staging memory with non-temporal stores.  Every image crosses host
// memory four times on its way through the path (file -> pinned, DMA out; DMA in, pinned -> the caller's malloc'ed result); at
// eight GPUs (26,000 images/s x ~3 MB) those copies, not the GPUs, are what the box runs out of: a plain memcpy reads the
// destination lines before overwriting them (write-allocate) and evicts the rest of the working set on the way.  Streaming stores
// skip both.  Destinations that a CPU reads again soon (small headers) should use memcpy.
#pragma once
#include <emmintrin.h>
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace b200 {

inline void stream_copy(void *dst_, const void *src_, size_t n)
{
    uint8_t *dst = static_cast<uint8_t *>(dst_); const uint8_t *src = static_cast<const uint8_t *>(src_);
    if (n < 4096) { memcpy(dst, src, n); return; }
    const size_t head = (16 - ((uintptr_t)dst & 15)) & 15;
    if (head) { memcpy(dst, src, head); dst += head; src += head; n -= head; }
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i)), b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 32)), d = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 48));
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), a); _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 32), c); _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 48), d);
    }
    _mm_sfence();
    if (i < n) memcpy(dst + i, src + i, n - i);
}

} // namespace b200

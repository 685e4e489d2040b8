// png_deflate.h -- device DEFLATE writer of the lossless PNG path (png_deflate.cu)
#pragma once
#include <cstddef>
#include <cstdint>

namespace b200 {
// scratch needed for a token stream of at most max_tokens tokens
size_t png_deflate_scratch_bytes(size_t max_tokens, int block_tokens);
// d_tok: tokens in HBM, *d_ntok their number (device); writes the zlib payload LSB-first into d_words starting at bit 16 (the two
// zlib header bytes are the host's) and d_total[0] = total bits incl. those 16, d_total[1] = *d_ntok.  When the payload would not
// fit words_cap_bytes nothing is written (d_total[0] still says how large it is).  Returns a cudaError_t.
int launch_png_deflate(const uint32_t *d_tok, const uint32_t *d_ntok, size_t max_tokens, int block_tokens, uint8_t *d_scratch, uint32_t *d_words, size_t words_cap_bytes,
                       unsigned long long *d_total, void *stream);
} // namespace b200

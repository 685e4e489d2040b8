// vp8l_decode.h -- host decoder of the WebP lossless bitstream (VP8L) and of the ALPH chunk (see vp8l_decode.cpp).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace b200 {

// image stream of implicit dimensions (no 5-byte header) -> ARGB pixels
bool vp8l_decode_stream(const uint8_t *data, size_t len, int width, int height, std::vector<uint32_t> &argb, std::string &err);
// payload of a 'VP8L' chunk (signature, dimensions, alpha hint, version, image stream)
bool vp8l_decode_file_chunk(const uint8_t *chunk, size_t len, int &width, int &height, bool &has_alpha, std::vector<uint32_t> &argb, std::string &err);
// payload of an 'ALPH' chunk -> the alpha plane (raw or VP8L-coded, horizontal / vertical / gradient filters undone)
bool webp_alpha_decode(const uint8_t *alph, size_t len, int width, int height, std::vector<uint8_t> &alpha, std::string &err);

} // namespace b200

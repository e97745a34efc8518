// Image input for image textures (implementation: image_in.cpp).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace rt2 {
// Decodes a PNG (8-bit, non-interlaced) or PPM (P3 / P6) file to 8-bit RGB, row 0 = top.  Returns false with `err` set.
bool DecodeImageFile(const std::string& path, int* width, int* height, std::vector<uint8_t>* rgb8, std::string* err);
}  // namespace rt2

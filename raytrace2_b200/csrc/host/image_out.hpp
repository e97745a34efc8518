// Image output: replaces util::WriteImage(vector<vec3>, w, h, path, png) (src/Util.cpp:39-79).
#pragma once
#include <cstdint>
#include <string>

namespace rt2 {
// sqrt gamma + clamp(255.999*c, 0, 255) -> u8, rows flipped so that dst row 0 is the TOP of the image
// (the reference hands bottom-up rows to stb with flip-on-write, Util.cpp:56-68).
void TonemapRGB8(const float* mean_rgb, int width, int height, uint8_t* dst);
// PNG (8-bit RGB, zlib deflate) or ASCII PPM "P3". Returns false with `err` set on I/O failure.
bool WriteImage(const float* mean_rgb, int width, int height, const std::string& out_path, bool png, std::string* err);
}  // namespace rt2

// See image_out.hpp.  The reference encodes PNG with the vendored stb_image_write; the byte stream differs (stb uses
// its own deflate), the decoded pixels are identical.
#include "image_out.hpp"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <vector>

namespace rt2 {

namespace {
// to_color lambda, Util.cpp:41-48: float sqrt, then double multiply and clamp, then int truncation
inline uint8_t ToByte(float c) {
  float g = std::sqrt(c);
  double v = std::clamp(static_cast<double>(g) * 255.999, 0.0, 255.0);
  int i = static_cast<int>(v);  // NaN (sqrt of a negative mean) -> INT_MIN -> unsigned char 0, as on x86-64
  return static_cast<uint8_t>(static_cast<unsigned char>(i));
}

void PutU32(std::vector<uint8_t>& v, uint32_t x) {
  v.push_back(static_cast<uint8_t>(x >> 24));
  v.push_back(static_cast<uint8_t>(x >> 16));
  v.push_back(static_cast<uint8_t>(x >> 8));
  v.push_back(static_cast<uint8_t>(x));
}

void PutChunk(std::vector<uint8_t>& out, const char* tag, const uint8_t* data, size_t n) {
  PutU32(out, static_cast<uint32_t>(n));
  size_t start = out.size();
  out.insert(out.end(), tag, tag + 4);
  if (n) out.insert(out.end(), data, data + n);
  uint32_t crc = static_cast<uint32_t>(crc32(0L, out.data() + start, static_cast<uInt>(n + 4)));
  PutU32(out, crc);
}
}  // namespace

void TonemapRGB8(const float* mean_rgb, int width, int height, uint8_t* dst) {
  for (int row = 0; row < height; row++) {
    int src_row = height - 1 - row;  // pixel row 0 is the bottom of the image (RayTracer.cpp:97-102, Camera.hpp:36-38)
    const float* s = mean_rgb + static_cast<size_t>(src_row) * width * 3;
    uint8_t* d = dst + static_cast<size_t>(row) * width * 3;
    for (int i = 0; i < width * 3; i++) d[i] = ToByte(s[i]);
  }
}

bool WriteImage(const float* mean_rgb, int width, int height, const std::string& out_path, bool png, std::string* err) {
  if (width <= 0 || height <= 0) {
    if (err) *err = "invalid image dims";
    return false;
  }
  std::vector<uint8_t> rgb(static_cast<size_t>(width) * height * 3);
  TonemapRGB8(mean_rgb, width, height, rgb.data());
  if (!png) {
    std::ofstream f(out_path);
    if (!f.is_open()) {
      if (err) *err = "cannot open " + out_path;
      return false;
    }
    f << "P3\n" << width << ' ' << height << "\n255\n";
    for (size_t i = 0; i < rgb.size(); i += 3)
      f << static_cast<int>(rgb[i]) << ' ' << static_cast<int>(rgb[i + 1]) << ' ' << static_cast<int>(rgb[i + 2]) << '\n';
    return true;
  }
  // filter type 0 on every scanline
  size_t stride = static_cast<size_t>(width) * 3;
  std::vector<uint8_t> raw((stride + 1) * height);
  for (int y = 0; y < height; y++) {
    raw[y * (stride + 1)] = 0;
    std::memcpy(&raw[y * (stride + 1) + 1], &rgb[y * stride], stride);
  }
  uLongf zlen = compressBound(static_cast<uLong>(raw.size()));
  std::vector<uint8_t> z(zlen);
  if (compress2(z.data(), &zlen, raw.data(), static_cast<uLong>(raw.size()), 6) != Z_OK) {
    if (err) *err = "zlib compress failed";
    return false;
  }
  std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  std::vector<uint8_t> ihdr;
  PutU32(ihdr, static_cast<uint32_t>(width));
  PutU32(ihdr, static_cast<uint32_t>(height));
  ihdr.push_back(8);  // bit depth
  ihdr.push_back(2);  // colour type RGB
  ihdr.push_back(0);
  ihdr.push_back(0);
  ihdr.push_back(0);
  PutChunk(out, "IHDR", ihdr.data(), ihdr.size());
  PutChunk(out, "IDAT", z.data(), zlen);
  PutChunk(out, "IEND", nullptr, 0);
  FILE* f = std::fopen(out_path.c_str(), "wb");
  if (!f) {
    if (err) *err = "cannot open " + out_path;
    return false;
  }
  size_t wr = std::fwrite(out.data(), 1, out.size(), f);
  std::fclose(f);
  if (wr != out.size()) {
    if (err) *err = "short write to " + out_path;
    return false;
  }
  return true;
}

}  // namespace rt2

// Image input for image textures (schema extension): binary / ASCII PPM and 8-bit PNG (grey, grey+alpha,
// RGB, RGBA, palette; non-interlaced), decoded to 8-bit RGB.  The reference vendors stb_image for this purpose but never
// calls it (dep/stb_image); this is an independent, minimal decoder on top of zlib's inflate.
#include "image_in.hpp"

#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace rt2 {
namespace {

bool ReadFile(const std::string& path, std::vector<uint8_t>* out) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::fseek(f, 0, SEEK_END);
  long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  if (n < 0) {
    std::fclose(f);
    return false;
  }
  out->resize(static_cast<size_t>(n));
  size_t got = n ? std::fread(out->data(), 1, static_cast<size_t>(n), f) : 0;
  std::fclose(f);
  return got == static_cast<size_t>(n);
}

// ---- PPM (P3 / P6), maxval <= 255 or 16-bit big endian ----
struct Cursor {
  const uint8_t* p;
  const uint8_t* end;
  void SkipSpaceAndComments() {
    while (p < end) {
      if (*p == '#') {
        while (p < end && *p != '\n') p++;
      } else if (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n') {
        p++;
      } else {
        break;
      }
    }
  }
  bool Int(long* v) {
    SkipSpaceAndComments();
    if (p >= end || *p < '0' || *p > '9') return false;
    long x = 0;
    while (p < end && *p >= '0' && *p <= '9') x = x * 10 + (*p++ - '0');
    *v = x;
    return true;
  }
};

bool DecodePPM(const std::vector<uint8_t>& d, int* w, int* h, std::vector<uint8_t>* rgb, std::string* err) {
  Cursor c{d.data() + 2, d.data() + d.size()};
  const bool binary = d[1] == '6';
  long W, H, maxv;
  if (!c.Int(&W) || !c.Int(&H) || !c.Int(&maxv) || W <= 0 || H <= 0 || maxv <= 0 || maxv > 65535 || W > 32768 || H > 32768) {
    *err = "malformed PPM header";
    return false;
  }
  rgb->resize(static_cast<size_t>(W) * H * 3);
  const size_t n = rgb->size();
  if (binary) {
    // exactly one whitespace byte follows maxval; a file that ends right after the header has none
    if (c.p >= c.end) {
      *err = "truncated PPM";
      return false;
    }
    c.p++;
    const size_t bytes = maxv > 255 ? 2 : 1;
    const long long remaining = static_cast<long long>(c.end - c.p);
    if (remaining < 0 || static_cast<unsigned long long>(remaining) < static_cast<unsigned long long>(n) * bytes) {
      *err = "truncated PPM";
      return false;
    }
    for (size_t i = 0; i < n; i++) {
      long v = bytes == 2 ? (c.p[2 * i] << 8 | c.p[2 * i + 1]) : c.p[i];
      if (v > maxv) v = maxv;  // like the P3 path
      (*rgb)[i] = static_cast<uint8_t>((v * 255 + maxv / 2) / maxv);
    }
  } else {
    for (size_t i = 0; i < n; i++) {
      long v;
      if (!c.Int(&v)) {
        *err = "truncated PPM";
        return false;
      }
      if (v > maxv) v = maxv;
      (*rgb)[i] = static_cast<uint8_t>((v * 255 + maxv / 2) / maxv);
    }
  }
  *w = static_cast<int>(W);
  *h = static_cast<int>(H);
  return true;
}

// ---- PNG ----
uint32_t Be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }

int Paeth(int a, int b, int c) {
  int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

bool DecodePNG(const std::vector<uint8_t>& d, int* w, int* h, std::vector<uint8_t>* rgb, std::string* err) {
  size_t pos = 8;
  uint32_t W = 0, H = 0;
  int depth = 0, ctype = -1, interlace = 0;
  std::vector<uint8_t> idat, palette;
  while (pos + 12 <= d.size()) {
    const uint32_t len = Be32(&d[pos]);
    const char* tag = reinterpret_cast<const char*>(&d[pos + 4]);
    if (pos + 12 + len > d.size()) break;
    const uint8_t* body = &d[pos + 8];
    if (!std::memcmp(tag, "IHDR", 4) && len >= 13) {
      W = Be32(body);
      H = Be32(body + 4);
      depth = body[8];
      ctype = body[9];
      interlace = body[12];
    } else if (!std::memcmp(tag, "PLTE", 4)) {
      palette.assign(body, body + len);
    } else if (!std::memcmp(tag, "IDAT", 4)) {
      idat.insert(idat.end(), body, body + len);
    } else if (!std::memcmp(tag, "IEND", 4)) {
      break;
    }
    pos += 12 + len;
  }
  int channels = 0;
  switch (ctype) {
    case 0: channels = 1; break;
    case 2: channels = 3; break;
    case 3: channels = 1; break;
    case 4: channels = 2; break;
    case 6: channels = 4; break;
    default: break;
  }
  if (W == 0 || H == 0 || W > 32768 || H > 32768 || channels == 0 || depth != 8 || interlace != 0) {
    *err = "unsupported PNG (need 8-bit, non-interlaced grey / RGB / RGBA / palette)";
    return false;
  }
  const size_t stride = static_cast<size_t>(W) * channels;
  std::vector<uint8_t> raw((stride + 1) * H);
  uLongf out_len = static_cast<uLongf>(raw.size());
  if (uncompress(raw.data(), &out_len, idat.data(), static_cast<uLong>(idat.size())) != Z_OK || out_len != raw.size()) {
    *err = "PNG inflate failed";
    return false;
  }
  // undo the scanline filters in place (rows of `cur` start after their filter byte)
  std::vector<uint8_t> prev(stride, 0), cur(stride);
  rgb->resize(static_cast<size_t>(W) * H * 3);
  for (uint32_t y = 0; y < H; y++) {
    const uint8_t* row = &raw[(stride + 1) * y];
    const int filter = row[0];
    for (size_t x = 0; x < stride; x++) {
      const int a = x >= static_cast<size_t>(channels) ? cur[x - channels] : 0;
      const int b = prev[x];
      const int c = x >= static_cast<size_t>(channels) ? prev[x - channels] : 0;
      int v = row[1 + x];
      switch (filter) {
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: v += Paeth(a, b, c); break;
        default: break;
      }
      cur[x] = static_cast<uint8_t>(v);
    }
    for (uint32_t x = 0; x < W; x++) {
      uint8_t* o = &(*rgb)[(static_cast<size_t>(y) * W + x) * 3];
      const uint8_t* s = &cur[static_cast<size_t>(x) * channels];
      if (ctype == 0 || ctype == 4) {
        o[0] = o[1] = o[2] = s[0];
      } else if (ctype == 3) {
        const size_t k = static_cast<size_t>(s[0]) * 3;
        if (k + 2 < palette.size()) {
          o[0] = palette[k], o[1] = palette[k + 1], o[2] = palette[k + 2];
        } else {
          o[0] = o[1] = o[2] = 0;
        }
      } else {
        o[0] = s[0], o[1] = s[1], o[2] = s[2];
      }
    }
    prev.swap(cur);
  }
  *w = static_cast<int>(W);
  *h = static_cast<int>(H);
  return true;
}

}  // namespace

bool DecodeImageFile(const std::string& path, int* width, int* height, std::vector<uint8_t>* rgb8, std::string* err) {
  std::vector<uint8_t> d;
  if (!ReadFile(path, &d)) {
    *err = "cannot read image file " + path;
    return false;
  }
  static const uint8_t kPngSig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  if (d.size() >= 8 && !std::memcmp(d.data(), kPngSig, 8)) return DecodePNG(d, width, height, rgb8, err);
  if (d.size() >= 2 && d[0] == 'P' && (d[1] == '3' || d[1] == '6')) return DecodePPM(d, width, height, rgb8, err);
  *err = "unknown image format (PNG and PPM are supported): " + path;
  return false;
}

}  // namespace rt2

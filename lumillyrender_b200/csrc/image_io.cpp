// image_io.cpp — output stage and HDR input of the host side.
//   PNG  : main.rs:160-164,171-173 + img.rs:52-63  — u8(clamp(x,0,1)^(1/gamma) * 255), truncating
//   HDR  : main.rs:155-158 + img.rs:40-50          — Radiance RGBE, linear RGB, `-Y H +X W`
//   IBL  : sky.rs:42-54                            — RGBE decode c * 2^(e-136), e = 0 -> 0
// The reference delegates the codecs to the `image` 0.18 crate; these are independent writers and a
// reader of the same standard formats (zlib supplies deflate + crc32).
#include <zlib.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <vector>

#include "host_scene.h"

namespace lr {

namespace {

void put_be32(std::vector<unsigned char>& v, uint32_t x) {
  v.push_back((unsigned char)(x >> 24)); v.push_back((unsigned char)(x >> 16)); v.push_back((unsigned char)(x >> 8)); v.push_back((unsigned char)x);
}
void png_chunk(std::vector<unsigned char>& out, const char* tag, const std::vector<unsigned char>& data) {
  put_be32(out, (uint32_t)data.size());
  const size_t start = out.size();
  out.insert(out.end(), tag, tag + 4);
  out.insert(out.end(), data.begin(), data.end());
  const uint32_t crc = (uint32_t)crc32(0L, out.data() + start, (uInt)(out.size() - start));
  put_be32(out, crc);
}

// main.rs:171-173.  Rust `as u8` saturates and maps NaN to 0.
unsigned char to_color(float x, float gamma) {
  const float c = std::fmin(std::fmax(x, 0.0f), 1.0f);      // f32::max/min ignore NaN: NaN.max(0.0) = 0.0
  const float v = std::pow(c, 1.0f / gamma) * 255.0f;
  if (!(v > 0.0f)) return 0;
  if (v >= 255.0f) return 255;
  return (unsigned char)v;
}

void float_to_rgbe(const float* rgb, unsigned char* out) {
  const float mx = std::fmax(rgb[0], std::fmax(rgb[1], rgb[2]));
  if (!(mx > 1e-32f)) { out[0] = out[1] = out[2] = out[3] = 0; return; }
  int e = 0;
  const float scale = std::frexp(mx, &e) * 256.0f / mx;
  for (int k = 0; k < 3; k++) {
    const float v = rgb[k] * scale;
    out[k] = (unsigned char)(v > 0.0f ? (v < 255.0f ? v : 255.0f) : 0.0f);
  }
  out[3] = (unsigned char)(e + 128);
}

// one channel of a scanline, new-style RLE (runs of >= 3 are worth encoding)
void rle_channel(const unsigned char* data, int n, int stride, std::vector<unsigned char>& out) {
  int i = 0;
  while (i < n) {
    int run = 1;
    while (i + run < n && run < 127 && data[(size_t)(i + run) * stride] == data[(size_t)i * stride]) run++;
    if (run >= 3) {
      out.push_back((unsigned char)(128 + run)); out.push_back(data[(size_t)i * stride]);
      i += run;
      continue;
    }
    const int start = i;
    int len = 0;
    while (i < n && len < 128) {
      int r = 1;
      while (i + r < n && r < 3 && data[(size_t)(i + r) * stride] == data[(size_t)i * stride]) r++;
      if (r >= 3) break;
      i++; len++;
    }
    out.push_back((unsigned char)len);
    for (int k = 0; k < len; k++) out.push_back(data[(size_t)(start + k) * stride]);
  }
}

}  // namespace

int load_hdr_file(const std::string& path, std::vector<float>& rgb, int& w, int& h) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return fail(LR_ERR_IO, "File `" + path + "` is not found.");
  std::string line;
  if (!std::getline(f, line) || (line.rfind("#?RADIANCE", 0) != 0 && line.rfind("#?RGBE", 0) != 0)) return fail(LR_ERR_PARSE, path + ": not a Radiance HDR file");
  bool format_ok = false;
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty()) break;
    if (line.rfind("FORMAT=", 0) == 0) {
      if (line != "FORMAT=32-bit_rle_rgbe") return fail(LR_ERR_UNSUPPORTED, path + ": unsupported " + line);
      format_ok = true;
    }
  }
  (void)format_ok;
  if (!std::getline(f, line)) return fail(LR_ERR_PARSE, path + ": missing resolution line");
  if (std::sscanf(line.c_str(), "-Y %d +X %d", &h, &w) != 2 || w <= 0 || h <= 0) return fail(LR_ERR_UNSUPPORTED, path + ": only `-Y H +X W` orientation is supported");
  rgb.assign((size_t)w * h * 3, 0.0f);
  std::vector<unsigned char> scan((size_t)w * 4);
  for (int y = 0; y < h; y++) {
    unsigned char head[4];
    if (!f.read((char*)head, 4)) return fail(LR_ERR_PARSE, path + ": truncated scanline");
    if (head[0] == 2 && head[1] == 2 && (head[2] & 0x80) == 0 && w >= 8 && w < 32768) {
      if (((int)head[2] << 8 | head[3]) != w) return fail(LR_ERR_PARSE, path + ": scanline width mismatch");
      for (int ch = 0; ch < 4; ch++) {
        int x = 0;
        while (x < w) {
          unsigned char cnt;
          if (!f.read((char*)&cnt, 1)) return fail(LR_ERR_PARSE, path + ": truncated RLE data");
          if (cnt > 128) {
            int run = cnt - 128;
            unsigned char val;
            if (!f.read((char*)&val, 1) || x + run > w) return fail(LR_ERR_PARSE, path + ": bad RLE run");
            while (run--) scan[(size_t)(x++) * 4 + ch] = val;
          } else {
            int len = cnt;
            if (len == 0 || x + len > w) return fail(LR_ERR_PARSE, path + ": bad RLE literal");
            while (len--) { unsigned char val; if (!f.read((char*)&val, 1)) return fail(LR_ERR_PARSE, path + ": truncated RLE data"); scan[(size_t)(x++) * 4 + ch] = val; }
          }
        }
      }
    } else {
      std::memcpy(scan.data(), head, 4);
      if (w > 1 && !f.read((char*)scan.data() + 4, (std::streamsize)(w - 1) * 4)) return fail(LR_ERR_PARSE, path + ": truncated flat scanline");
    }
    for (int x = 0; x < w; x++) {
      const unsigned char* p = &scan[(size_t)x * 4];
      float* o = &rgb[((size_t)y * w + x) * 3];
      if (p[3] == 0) { o[0] = o[1] = o[2] = 0.0f; continue; }
      const float s = std::ldexp(1.0f, (int)p[3] - 136);
      o[0] = (float)p[0] * s; o[1] = (float)p[1] * s; o[2] = (float)p[2] * s;
    }
  }
  return LR_OK;
}

}  // namespace lr

using namespace lr;

extern "C" {

static int lr_save_png_body(const char* path, const float* rgb, int32_t w, int32_t h, float gamma) {
  if (!path || !rgb || w <= 0 || h <= 0) return fail(LR_ERR_INVALID, "bad argument");
  std::vector<unsigned char> raw((size_t)h * (1 + (size_t)w * 3));
  for (int y = 0; y < h; y++) {
    unsigned char* row = &raw[(size_t)y * (1 + (size_t)w * 3)];
    row[0] = 0;   // filter: none
    for (int x = 0; x < w; x++)
      for (int k = 0; k < 3; k++) row[1 + 3 * x + k] = to_color(rgb[((size_t)y * w + x) * 3 + k], gamma);
  }
  uLongf zlen = compressBound((uLong)raw.size());
  std::vector<unsigned char> z(zlen);
  if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return fail(LR_ERR_IO, "zlib compress failed");
  z.resize(zlen);
  std::vector<unsigned char> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  std::vector<unsigned char> ihdr;
  put_be32(ihdr, (uint32_t)w); put_be32(ihdr, (uint32_t)h);
  ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);   // 8-bit RGB
  png_chunk(out, "IHDR", ihdr);
  png_chunk(out, "IDAT", z);
  png_chunk(out, "IEND", {});
  FILE* f = std::fopen(path, "wb");
  if (!f) return fail(LR_ERR_IO, std::string("cannot create `") + path + "`");
  const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
  std::fclose(f);
  return ok ? LR_OK : fail(LR_ERR_IO, std::string("short write to `") + path + "`");
}
int lr_save_png(const char* path, const float* rgb, int32_t w, int32_t h, float gamma) { LR_GUARDED(lr_save_png_body(path, rgb, w, h, gamma)); }

static int lr_save_hdr_body(const char* path, const float* rgb, int32_t w, int32_t h) {
  if (!path || !rgb || w <= 0 || h <= 0) return fail(LR_ERR_INVALID, "bad argument");
  FILE* f = std::fopen(path, "wb");
  if (!f) return fail(LR_ERR_IO, std::string("cannot create `") + path + "`");
  std::fprintf(f, "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n", h, w);
  std::vector<unsigned char> scan((size_t)w * 4), enc;
  bool ok = true;
  for (int y = 0; y < h && ok; y++) {
    for (int x = 0; x < w; x++) float_to_rgbe(&rgb[((size_t)y * w + x) * 3], &scan[(size_t)x * 4]);
    if (w < 8 || w >= 32768) { ok = std::fwrite(scan.data(), 1, scan.size(), f) == scan.size(); continue; }
    enc.clear();
    enc.push_back(2); enc.push_back(2); enc.push_back((unsigned char)(w >> 8)); enc.push_back((unsigned char)(w & 255));
    for (int ch = 0; ch < 4; ch++) rle_channel(scan.data() + ch, w, 4, enc);
    ok = std::fwrite(enc.data(), 1, enc.size(), f) == enc.size();
  }
  std::fclose(f);
  return ok ? LR_OK : fail(LR_ERR_IO, std::string("short write to `") + path + "`");
}
int lr_save_hdr(const char* path, const float* rgb, int32_t w, int32_t h) { LR_GUARDED(lr_save_hdr_body(path, rgb, w, h)); }

static int lr_load_hdr_body(const char* path, float** rgb, int32_t* w, int32_t* h) {
  if (!path || !rgb || !w || !h) return fail(LR_ERR_INVALID, "bad argument");
  std::vector<float> px;
  int ww = 0, hh = 0;
  if (int rc = load_hdr_file(path, px, ww, hh)) return rc;
  float* out = (float*)std::malloc(px.size() * sizeof(float));
  if (!out) return fail(LR_ERR_INVALID, "out of memory");
  std::memcpy(out, px.data(), px.size() * sizeof(float));
  *rgb = out; *w = ww; *h = hh;
  return LR_OK;
}
int lr_load_hdr(const char* path, float** rgb, int32_t* w, int32_t* h) { LR_GUARDED(lr_load_hdr_body(path, rgb, w, h)); }

void lr_free(void* p) { std::free(p); }

}  // extern "C"

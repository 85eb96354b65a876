// image_io.cpp — see image_io.hpp
#include "image_io.hpp"

#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstring>
#include <fstream>

#ifdef PANO_WITH_OPENCV
#include <opencv2/imgcodecs.hpp>
#endif
#ifdef PANO_WITH_ZLIB
#include <zlib.h>
#endif
#ifdef PANO_WITH_NVJPEG
#include <cuda_runtime.h>
#include <nvjpeg.h>
#endif

namespace pano_io {

namespace {

std::vector<uint8_t> slurp(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return {};
  return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

std::string ext_of(const std::string& p) {
  size_t d = p.find_last_of('.');
  std::string e = d == std::string::npos ? "" : p.substr(d + 1);
  std::transform(e.begin(), e.end(), e.begin(), [](unsigned char c) { return (char)std::tolower(c); });
  return e;
}

// ---------------------------------------------------------------- PNM
bool pnm_token(const std::vector<uint8_t>& d, size_t& pos, int& out) {
  while (pos < d.size()) {
    if (d[pos] == '#') { while (pos < d.size() && d[pos] != '\n') pos++; }
    else if (std::isspace(d[pos])) pos++;
    else break;
  }
  if (pos >= d.size() || !std::isdigit(d[pos])) return false;
  out = 0;
  while (pos < d.size() && std::isdigit(d[pos])) out = out * 10 + (d[pos++] - '0');
  return true;
}

Image read_pnm(const std::vector<uint8_t>& d) {
  Image im;
  if (d.size() < 3 || d[0] != 'P' || (d[1] != '6' && d[1] != '5')) return im;
  const int ch = d[1] == '6' ? 3 : 1;
  size_t pos = 2;
  int w, h, mx;
  if (!pnm_token(d, pos, w) || !pnm_token(d, pos, h) || !pnm_token(d, pos, mx) || mx != 255) return im;
  pos++;  // single whitespace after maxval
  if (w <= 0 || h <= 0 || d.size() < pos + (size_t)w * h * ch) return im;
  im.w = w; im.h = h;
  im.bgr.resize((size_t)w * h * 3);
  const uint8_t* s = d.data() + pos;
  for (size_t i = 0; i < (size_t)w * h; i++) {
    if (ch == 3) { im.bgr[3 * i] = s[3 * i + 2]; im.bgr[3 * i + 1] = s[3 * i + 1]; im.bgr[3 * i + 2] = s[3 * i]; }
    else im.bgr[3 * i] = im.bgr[3 * i + 1] = im.bgr[3 * i + 2] = s[i];
  }
  return im;
}

bool write_ppm(const std::string& path, const uint8_t* bgr, int w, int h, size_t stride) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  fprintf(f, "P6\n%d %d\n255\n", w, h);
  std::vector<uint8_t> row((size_t)w * 3);
  for (int y = 0; y < h; y++) {
    const uint8_t* s = bgr + y * stride;
    for (int x = 0; x < w; x++) { row[3 * x] = s[3 * x + 2]; row[3 * x + 1] = s[3 * x + 1]; row[3 * x + 2] = s[3 * x]; }
    fwrite(row.data(), 1, row.size(), f);
  }
  return fclose(f) == 0;
}

// ---------------------------------------------------------------- BMP
uint32_t rd32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }

Image read_bmp(const std::vector<uint8_t>& d) {
  Image im;
  if (d.size() < 54 || d[0] != 'B' || d[1] != 'M') return im;
  uint32_t off = rd32(&d[10]);
  int w = (int)rd32(&d[18]), hh = (int)rd32(&d[22]);
  int bpp = d[28] | (d[29] << 8);
  uint32_t comp = rd32(&d[30]);
  if ((bpp != 24 && bpp != 32) || (comp != 0 && comp != 3) || w <= 0 || hh == 0) return im;
  bool flip = hh > 0;
  int h = hh > 0 ? hh : -hh;
  size_t rs = ((size_t)w * (bpp / 8) + 3) & ~(size_t)3;
  if (d.size() < off + rs * h) return im;
  im.w = w; im.h = h;
  im.bgr.resize((size_t)w * h * 3);
  for (int y = 0; y < h; y++) {
    const uint8_t* s = d.data() + off + rs * (flip ? h - 1 - y : y);
    uint8_t* o = &im.bgr[(size_t)y * w * 3];
    for (int x = 0; x < w; x++) { o[3 * x] = s[x * (bpp / 8)]; o[3 * x + 1] = s[x * (bpp / 8) + 1]; o[3 * x + 2] = s[x * (bpp / 8) + 2]; }
  }
  return im;
}

bool write_bmp(const std::string& path, const uint8_t* bgr, int w, int h, size_t stride) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  size_t rs = ((size_t)w * 3 + 3) & ~(size_t)3;
  uint8_t hd[54] = {0};
  auto w32 = [&](int o, uint32_t v) { hd[o] = v; hd[o + 1] = v >> 8; hd[o + 2] = v >> 16; hd[o + 3] = v >> 24; };
  hd[0] = 'B'; hd[1] = 'M';
  w32(2, (uint32_t)(54 + rs * h)); w32(10, 54); w32(14, 40); w32(18, w); w32(22, h);
  hd[26] = 1; hd[28] = 24; w32(34, (uint32_t)(rs * h));
  fwrite(hd, 1, 54, f);
  std::vector<uint8_t> row(rs, 0);
  for (int y = h - 1; y >= 0; y--) {
    memcpy(row.data(), bgr + y * stride, (size_t)w * 3);
    fwrite(row.data(), 1, rs, f);
  }
  return fclose(f) == 0;
}

// ---------------------------------------------------------------- PNG (zlib)
#ifdef PANO_WITH_ZLIB
uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | (p[1] << 16) | (p[2] << 8) | p[3]; }

Image read_png(const std::vector<uint8_t>& d) {
  Image im;
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (d.size() < 33 || memcmp(d.data(), sig, 8)) return im;
  size_t pos = 8;
  int w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
  std::vector<uint8_t> idat, plte;
  while (pos + 12 <= d.size()) {
    uint32_t len = be32(&d[pos]);
    const uint8_t* type = &d[pos + 4];
    const uint8_t* body = &d[pos + 8];
    if (pos + 12 + len > d.size()) return im;
    if (!memcmp(type, "IHDR", 4)) { w = be32(body); h = be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12]; }
    else if (!memcmp(type, "PLTE", 4)) plte.assign(body, body + len);
    else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
    else if (!memcmp(type, "IEND", 4)) break;
    pos += 12 + len;
  }
  int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
  if (w <= 0 || h <= 0 || depth != 8 || ch == 0 || interlace != 0) return im;
  size_t rb = (size_t)w * ch;
  std::vector<uint8_t> raw((rb + 1) * h);
  uLongf rl = raw.size();
  if (uncompress(raw.data(), &rl, idat.data(), idat.size()) != Z_OK || rl != raw.size()) return im;
  std::vector<uint8_t> cur(rb), prev(rb, 0);
  im.w = w; im.h = h;
  im.bgr.resize((size_t)w * h * 3);
  for (int y = 0; y < h; y++) {
    const uint8_t* s = &raw[(rb + 1) * y];
    int ft = s[0];
    for (size_t i = 0; i < rb; i++) {
      int a = i >= (size_t)ch ? cur[i - ch] : 0, b = prev[i], c = i >= (size_t)ch ? prev[i - ch] : 0, x = s[1 + i];
      int v;
      switch (ft) {
        case 0: v = x; break;
        case 1: v = x + a; break;
        case 2: v = x + b; break;
        case 3: v = x + ((a + b) >> 1); break;
        default: { int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
                   v = x + ((pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c)); }
      }
      cur[i] = (uint8_t)v;
    }
    uint8_t* o = &im.bgr[(size_t)y * w * 3];
    for (int x = 0; x < w; x++) {
      uint8_t r, g, b;
      if (ctype == 2 || ctype == 6) { r = cur[x * ch]; g = cur[x * ch + 1]; b = cur[x * ch + 2]; }
      else if (ctype == 3) { size_t pi = (size_t)cur[x] * 3; if (pi + 2 >= plte.size()) { r = g = b = 0; } else { r = plte[pi]; g = plte[pi + 1]; b = plte[pi + 2]; } }
      else { r = g = b = cur[x * ch]; }
      o[3 * x] = b; o[3 * x + 1] = g; o[3 * x + 2] = r;
    }
    prev.swap(cur);
  }
  return im;
}

bool write_png(const std::string& path, const uint8_t* bgr, int w, int h, size_t stride) {
  std::vector<uint8_t> raw(((size_t)w * 3 + 1) * h);
  for (int y = 0; y < h; y++) {
    uint8_t* o = &raw[((size_t)w * 3 + 1) * y];
    o[0] = 0;
    const uint8_t* s = bgr + y * stride;
    for (int x = 0; x < w; x++) { o[1 + 3 * x] = s[3 * x + 2]; o[2 + 3 * x] = s[3 * x + 1]; o[3 + 3 * x] = s[3 * x]; }
  }
  uLongf cl = compressBound(raw.size());
  std::vector<uint8_t> comp(cl);
  if (compress2(comp.data(), &cl, raw.data(), raw.size(), 3) != Z_OK) return false;
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  fwrite(sig, 1, 8, f);
  auto chunk = [&](const char* type, const uint8_t* body, uint32_t len) {
    uint8_t l[4] = {(uint8_t)(len >> 24), (uint8_t)(len >> 16), (uint8_t)(len >> 8), (uint8_t)len};
    fwrite(l, 1, 4, f);
    fwrite(type, 1, 4, f);
    if (len) fwrite(body, 1, len, f);
    uLong c = crc32(0, (const Bytef*)type, 4);
    if (len) c = crc32(c, body, len);
    uint8_t cc[4] = {(uint8_t)(c >> 24), (uint8_t)(c >> 16), (uint8_t)(c >> 8), (uint8_t)c};
    fwrite(cc, 1, 4, f);
  };
  uint8_t ih[13] = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w,
                    (uint8_t)(h >> 24), (uint8_t)(h >> 16), (uint8_t)(h >> 8), (uint8_t)h, 8, 2, 0, 0, 0};
  chunk("IHDR", ih, 13);
  chunk("IDAT", comp.data(), (uint32_t)cl);
  chunk("IEND", nullptr, 0);
  return fclose(f) == 0;
}
#endif

// ---------------------------------------------------------------- JPEG (nvJPEG)
#ifdef PANO_WITH_NVJPEG
struct NvJpeg {
  nvjpegHandle_t h = nullptr;
  nvjpegJpegState_t st = nullptr;
  bool ok = false;
  NvJpeg() {
    if (nvjpegCreateSimple(&h) != NVJPEG_STATUS_SUCCESS) return;
    if (nvjpegJpegStateCreate(h, &st) != NVJPEG_STATUS_SUCCESS) return;
    ok = true;
  }
};
NvJpeg& nvj() { static NvJpeg n; return n; }

Image read_jpeg(const std::vector<uint8_t>& d) {
  Image im;
  NvJpeg& n = nvj();
  if (!n.ok) return im;
  int nc = 0, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
  nvjpegChromaSubsampling_t ss;
  if (nvjpegGetImageInfo(n.h, d.data(), d.size(), &nc, &ss, ws, hs) != NVJPEG_STATUS_SUCCESS) return im;
  int w = ws[0], h = hs[0];
  nvjpegImage_t out;
  memset(&out, 0, sizeof out);
  if (cudaMalloc((void**)&out.channel[0], (size_t)w * h * 3) != cudaSuccess) return im;
  out.pitch[0] = (size_t)w * 3;
  bool good = nvjpegDecode(n.h, n.st, d.data(), d.size(), NVJPEG_OUTPUT_BGRI, &out, 0) == NVJPEG_STATUS_SUCCESS &&
              cudaStreamSynchronize(0) == cudaSuccess;
  if (good) {
    im.w = w; im.h = h;
    im.bgr.resize((size_t)w * h * 3);
    good = cudaMemcpy(im.bgr.data(), out.channel[0], im.bgr.size(), cudaMemcpyDeviceToHost) == cudaSuccess;
  }
  cudaFree(out.channel[0]);
  if (!good) im = Image();
  return im;
}

bool write_jpeg_device(const std::string& path, const uint8_t* dev, int w, int h, size_t stride) {
  NvJpeg& n = nvj();
  if (!n.ok) return false;
  nvjpegEncoderState_t es = nullptr;
  nvjpegEncoderParams_t ep = nullptr;
  bool good = nvjpegEncoderStateCreate(n.h, &es, 0) == NVJPEG_STATUS_SUCCESS &&
              nvjpegEncoderParamsCreate(n.h, &ep, 0) == NVJPEG_STATUS_SUCCESS;
  std::vector<uint8_t> bs;
  if (good) {
    nvjpegEncoderParamsSetQuality(ep, 95, 0);                        // cv::imwrite default quality
    nvjpegEncoderParamsSetSamplingFactors(ep, NVJPEG_CSS_420, 0);    // and chroma subsampling
    nvjpegImage_t src;
    memset(&src, 0, sizeof src);
    src.channel[0] = const_cast<uint8_t*>(dev);
    src.pitch[0] = stride;
    size_t len = 0;
    good = nvjpegEncodeImage(n.h, es, ep, &src, NVJPEG_INPUT_BGRI, w, h, 0) == NVJPEG_STATUS_SUCCESS &&
           nvjpegEncodeRetrieveBitstream(n.h, es, nullptr, &len, 0) == NVJPEG_STATUS_SUCCESS;
    if (good) {
      bs.resize(len);
      good = nvjpegEncodeRetrieveBitstream(n.h, es, bs.data(), &len, 0) == NVJPEG_STATUS_SUCCESS &&
             cudaStreamSynchronize(0) == cudaSuccess;
      bs.resize(len);
    }
  }
  if (ep) nvjpegEncoderParamsDestroy(ep);
  if (es) nvjpegEncoderStateDestroy(es);
  if (!good) return false;
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  fwrite(bs.data(), 1, bs.size(), f);
  return fclose(f) == 0;
}
#endif

// JPEG (and anything else OpenCV reads) through the Python cv2 wheel when OpenCV C++ is absent: the pixels are then
// exactly what the reference's cv::imread produces (same libjpeg-turbo build family), which nvJPEG's decoder does not
// guarantee (tests/test_cli.py reports the per-pixel difference).  The helper converts to a binary PPM in a temporary
// file.  PANO_JPEG=nvjpeg skips it (fast path, decode on the GPU); PANO_PYTHON overrides the interpreter.
Image read_via_cv2_helper(const std::string& path) {
  const char* py = getenv("PANO_PYTHON");
  char tmpl[] = "/tmp/pano_imread_XXXXXX";
  int fd = mkstemp(tmpl);
  if (fd < 0) return Image();
  close(fd);
  std::string out = tmpl;
  std::string q;                       // shell-quote the path
  for (char ch : path) { if (ch == '\'') q += "'\\''"; else q += ch; }
  std::string cmd = std::string(py ? py : "python3") +
      " -c \"import sys,cv2; im=cv2.imread(sys.argv[1]); sys.exit(1) if im is None else None; "
      "open(sys.argv[2],'wb').write(b'P6\\n%d %d\\n255\\n' % (im.shape[1], im.shape[0]) + im[:, :, ::-1].tobytes())\" '" +
      q + "' '" + out + "' 2>/dev/null";
  Image im;
  if (system(cmd.c_str()) == 0) {
    std::vector<uint8_t> d = slurp(out);
    if (d.size() > 4 && d[0] == 'P' && d[1] == '6') im = read_pnm(d);
  }
  unlink(out.c_str());
  return im;
}

}  // namespace

Image read_image(const std::string& path) {
#ifdef PANO_WITH_OPENCV
  cv::Mat m = cv::imread(path);
  Image im;
  if (m.empty()) return im;
  im.w = m.cols; im.h = m.rows;
  im.bgr.resize((size_t)m.cols * m.rows * 3);
  for (int y = 0; y < m.rows; y++) memcpy(&im.bgr[(size_t)y * m.cols * 3], m.ptr(y), (size_t)m.cols * 3);
  return im;
#else
  std::vector<uint8_t> d = slurp(path);
  if (d.size() < 4) return Image();
  if (d[0] == 'P' && (d[1] == '6' || d[1] == '5')) return read_pnm(d);
  if (d[0] == 'B' && d[1] == 'M') return read_bmp(d);
#ifdef PANO_WITH_ZLIB
  if (d[0] == 0x89 && d[1] == 'P') return read_png(d);
#endif
  if (d[0] == 0xff && d[1] == 0xd8) {
    const char* mode = getenv("PANO_JPEG");
    const bool want_nvjpeg = mode && std::string(mode) == "nvjpeg";
    if (!want_nvjpeg) {
      Image im = read_via_cv2_helper(path);   // the reference's decoder (cv2.imread), pixel for pixel
      if (!im.empty()) return im;
    }
#ifdef PANO_WITH_NVJPEG
    return read_jpeg(d);
#endif
  }
  return Image();
#endif
}

bool write_image(const std::string& path, const uint8_t* bgr, int w, int h, size_t stride) {
#ifdef PANO_WITH_OPENCV
  cv::Mat m(h, w, CV_8UC3, const_cast<uint8_t*>(bgr), stride);
  return cv::imwrite(path, m);
#else
  std::string e = ext_of(path);
  if (e == "ppm" || e == "pnm") return write_ppm(path, bgr, w, h, stride);
  if (e == "bmp") return write_bmp(path, bgr, w, h, stride);
#ifdef PANO_WITH_ZLIB
  if (e == "png") return write_png(path, bgr, w, h, stride);
#endif
#ifdef PANO_WITH_NVJPEG
  if (e == "jpg" || e == "jpeg") {
    uint8_t* dev = nullptr;
    if (cudaMalloc((void**)&dev, stride * h) != cudaSuccess) return false;
    bool ok = cudaMemcpy(dev, bgr, stride * h, cudaMemcpyHostToDevice) == cudaSuccess && write_jpeg_device(path, dev, w, h, stride);
    cudaFree(dev);
    return ok;
  }
#endif
  fprintf(stderr, "write_image: no encoder for '.%s' in this build (have: ppm, bmp%s%s)\n", e.c_str(),
#ifdef PANO_WITH_ZLIB
          ", png",
#else
          "",
#endif
#ifdef PANO_WITH_NVJPEG
          ", jpg");
#else
          "");
#endif
  return false;
#endif
}

bool write_image_device(const std::string& path, const uint8_t* bgr_dev, int w, int h, size_t stride) {
#if defined(PANO_WITH_NVJPEG) && !defined(PANO_WITH_OPENCV)
  std::string e = ext_of(path);
  if (e == "jpg" || e == "jpeg") return write_jpeg_device(path, bgr_dev, w, h, stride);
#endif
#if defined(PANO_WITH_NVJPEG) || defined(PANO_HAVE_CUDA)
  std::vector<uint8_t> host((size_t)w * 3 * h);
  if (cudaMemcpy2D(host.data(), (size_t)w * 3, bgr_dev, stride, (size_t)w * 3, h, cudaMemcpyDeviceToHost) != cudaSuccess)
    return false;
  return write_image(path, host.data(), w, h, (size_t)w * 3);
#else
  (void)bgr_dev; (void)w; (void)h; (void)stride; (void)path;
  return false;
#endif
}

}  // namespace pano_io

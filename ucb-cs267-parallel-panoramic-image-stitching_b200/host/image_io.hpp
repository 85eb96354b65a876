// image_io.hpp — image file I/O for the command-line executables (the only place images are
// decoded/encoded; the engine itself works on raw BGR8 buffers).
//
// The reference uses cv::imread / cv::imwrite (ref: src/reader/reader.cpp:61,72;
// src/serial/main.cpp:445).  When the build finds OpenCV (-DPANO_WITH_OPENCV) exactly those
// calls are used.  Without OpenCV C++ (the situation on our build and GPU boxes) built-in codecs
// are used instead: PPM/PGM (P6/P5), 24/32-bit BMP, PNG (8-bit, non-interlaced, via zlib) and
// baseline JPEG through nvJPEG (CUDA toolkit).  nvJPEG's decoder is not bit-identical to
// libjpeg-turbo's, so parity runs hand the same decoded buffer to both sides.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace pano_io {

struct Image {          // 8-bit BGR, interleaved, tightly packed rows (like a continuous cv::Mat)
  int w = 0, h = 0;
  std::vector<uint8_t> bgr;
  bool empty() const { return w <= 0 || h <= 0 || bgr.empty(); }
  size_t stride() const { return (size_t)w * 3; }
};

// returns an empty image on failure (cv::imread semantics)
Image read_image(const std::string& path);
// format chosen by the file extension (cv::imwrite semantics); false on failure
bool write_image(const std::string& path, const uint8_t* bgr, int w, int h, size_t stride);
// same, from a device-resident canvas (used by gpu_stitching: JPEG is encoded by nvJPEG straight
// from device memory, other formats copy the canvas to the host first)
bool write_image_device(const std::string& path, const uint8_t* bgr_dev, int w, int h, size_t stride);

}  // namespace pano_io

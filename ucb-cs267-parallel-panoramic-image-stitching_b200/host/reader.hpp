// reader.hpp — command-line parsing + image loading shared by the three executables
// (gpu_stitching, serial_stitching, openmp_stitching).
//
// Interface mirror of the reference's reader library (ref: src/reader/reader.hpp:8-15, reader.cpp:14-82): the
// same entry point and result fields, so the executables' main() reads like the reference's, with decoded images
// held as pano_io::Image (tightly packed BGR8 + size) instead of cv::Mat - OpenCV C++ is not a dependency here.
//
//   <exe> [--dir directory] [--out output_file] [image1 image2 ...]
//
//   * no arguments, --dir / --out without a value, --dir that is not a directory  -> message on stderr, exit(-1)
//   * a file that cannot be decoded                                               -> warning on stderr, skipped
//   * --dir given                                                                 -> its regular files replace the
//     file names of the command line (directory_iterator order; PANO_SORT_DIR=1 sorts them)
//   * --out not given                                                             -> "result.jpg"
#pragma once
#include <string>
#include <vector>

#include "image_io.hpp"

struct ImageReaderResult {
  std::vector<pano_io::Image> images;   // in command-line (or directory) order, unreadable ones left out
  std::string outputFile;               // where the caller writes the panorama
};

// Never returns on a usage error (exit(-1), as the reference does).
ImageReaderResult readImagesFromArgs(int argc, char** argv);

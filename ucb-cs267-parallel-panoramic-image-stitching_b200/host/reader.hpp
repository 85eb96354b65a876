// reader.hpp — command-line parsing + image loading shared by the three executables.
// Mirrors the reference's reader library (ref: src/reader/reader.hpp:8-15, reader.cpp:14-82):
//   [--dir directory] [--out output_file] [image1 image2 ...]
#pragma once
#include <string>
#include <vector>

#include "image_io.hpp"

struct ImageReaderResult {
  std::vector<pano_io::Image> images;
  std::string outputFile;
};

ImageReaderResult readImagesFromArgs(int argc, char** argv);

// reader.cpp — see reader.hpp.  Behaviour kept from the reference (ref: src/reader/reader.cpp):
// default output "result.jpg"; usage errors exit(-1); unreadable files are warned about and
// skipped; with --dir every regular file of the directory is tried in directory_iterator order
// (unsorted, as the reference does — set PANO_SORT_DIR=1 for a sorted, reproducible order);
// anything that is not --dir/--out is taken as an image file name.
#include "reader.hpp"

#include <algorithm>
#include <cstdlib>
#include <filesystem>
#include <iostream>

namespace fs = std::filesystem;

ImageReaderResult readImagesFromArgs(int argc, char** argv) {
  ImageReaderResult result;
  result.outputFile = "result.jpg";
  std::vector<std::string> fileNames;
  std::string dirName;
  if (argc < 2) {
    std::cerr << "Usage: " << argv[0] << " [--dir directory_name] [--out output_file_name] [image1 image2 ...]" << std::endl;
    std::exit(-1);
  }
  for (int i = 1; i < argc; i++) {
    std::string arg(argv[i]);
    if (arg == "--dir") {
      if (i + 1 >= argc) { std::cerr << "Error: --dir requires a directory name" << std::endl; std::exit(-1); }
      dirName = argv[++i];
    } else if (arg == "--out") {
      if (i + 1 >= argc) { std::cerr << "Error: --out requires an output file name" << std::endl; std::exit(-1); }
      result.outputFile = argv[++i];
    } else {
      fileNames.push_back(arg);
    }
  }
  if (!dirName.empty()) {
    if (!fs::exists(dirName) || !fs::is_directory(dirName)) {
      std::cerr << "Error: " << dirName << " is not a valid directory." << std::endl;
      std::exit(-1);
    }
    fileNames.clear();
    for (const auto& entry : fs::directory_iterator(dirName))
      if (entry.is_regular_file()) fileNames.push_back(entry.path().string());
    const char* s = std::getenv("PANO_SORT_DIR");
    if (s && *s && *s != '0') std::sort(fileNames.begin(), fileNames.end());
  }
  for (const auto& fileName : fileNames) {
    pano_io::Image img = pano_io::read_image(fileName);
    if (img.empty()) {
      std::cerr << "Warning: Unable to open image file: " << fileName << std::endl;
      continue;
    }
    if (const char* dump = std::getenv("PANO_DUMP_DECODED")) {   // test hook: the pixels the pipeline will see
      const std::string out = std::string(dump) + "/decoded_" + std::to_string(result.images.size()) + ".ppm";
      pano_io::write_image(out, img.bgr.data(), img.w, img.h, img.stride());
    }
    result.images.push_back(std::move(img));
  }
  return result;
}

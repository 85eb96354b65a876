// reader.cpp — see reader.hpp.  The command-line contract of the reference's reader library is kept
// (ref: src/reader/reader.cpp:14-82): default output "result.jpg"; usage errors and a bad --dir exit(-1) with
// the reference's messages; unreadable files are warned about and skipped; --dir tries every regular file of
// the directory in directory_iterator order (unsorted like the reference; PANO_SORT_DIR=1 gives a sorted,
// reproducible order) and overrides file names given on the command line; everything that is not an option
// is an image file name.  Decoding goes through pano_io (image_io.cpp), not OpenCV.
#include "reader.hpp"

#include <algorithm>
#include <cstdlib>
#include <filesystem>
#include <iostream>

namespace {

[[noreturn]] void die(const std::string& message) {
  std::cerr << message << std::endl;
  std::exit(-1);
}

bool env_flag(const char* name) {
  const char* v = std::getenv(name);
  return v && *v && *v != '0';
}

struct CommandLine {
  std::string directory, output = "result.jpg";
  std::vector<std::string> files;
};

// options that take one value; the text is what the reference prints when the value is missing
struct ValueOption {
  const char* flag;
  const char* missing;
  std::string CommandLine::*target;
};
const ValueOption kOptions[] = {
    {"--dir", "Error: --dir requires a directory name", &CommandLine::directory},
    {"--out", "Error: --out requires an output file name", &CommandLine::output},
};

CommandLine parse(int argc, char** argv) {
  if (argc < 2)
    die(std::string("Usage: ") + argv[0] + " [--dir directory_name] [--out output_file_name] [image1 image2 ...]");
  CommandLine cl;
  for (int i = 1; i < argc; i++) {
    const std::string word(argv[i]);
    const ValueOption* opt = std::find_if(std::begin(kOptions), std::end(kOptions),
                                          [&](const ValueOption& o) { return word == o.flag; });
    if (opt == std::end(kOptions)) {
      cl.files.push_back(word);
      continue;
    }
    if (++i >= argc) die(opt->missing);
    cl.*(opt->target) = argv[i];
  }
  return cl;
}

std::vector<std::string> regular_files_of(const std::string& directory) {
  namespace fs = std::filesystem;
  std::error_code ec;
  if (!fs::is_directory(directory, ec)) die("Error: " + directory + " is not a valid directory.");
  std::vector<std::string> names;
  for (const fs::directory_entry& e : fs::directory_iterator(directory))
    if (e.is_regular_file()) names.push_back(e.path().string());
  if (env_flag("PANO_SORT_DIR")) std::sort(names.begin(), names.end());
  return names;
}

}  // namespace

ImageReaderResult readImagesFromArgs(int argc, char** argv) {
  const CommandLine cl = parse(argc, argv);
  const std::vector<std::string> names = cl.directory.empty() ? cl.files : regular_files_of(cl.directory);
  const char* dump_dir = std::getenv("PANO_DUMP_DECODED");   // test hook: the pixels the pipeline will see
  ImageReaderResult result;
  result.outputFile = cl.output;
  for (const std::string& name : names) {
    pano_io::Image img = pano_io::read_image(name);
    if (img.empty()) {
      std::cerr << "Warning: Unable to open image file: " << name << std::endl;
      continue;
    }
    if (dump_dir)
      pano_io::write_image(std::string(dump_dir) + "/decoded_" + std::to_string(result.images.size()) + ".ppm",
                           img.bgr.data(), img.w, img.h, img.stride());
    result.images.push_back(std::move(img));
  }
  return result;
}

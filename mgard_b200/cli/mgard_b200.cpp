// Command line front end with the interface of the reference's `mgard` executable
// (reference src/cli/executable.cpp:13-90, src/cli/cli_internal.cpp,
// include/cli/cli_internal.tpp):
//
//   mgard-b200 compress   --datatype float|double --shape 129x129x129
//                         --smoothness <s|inf> --tolerance <tau>
//                         --input <file> --output <file>
//   mgard-b200 decompress --input <file> --output <file>
//
// over the C ABI (mgb_cpu_compress / mgb_cpu_decompress).  Files written by either
// executable are read by the other.  One extension: --lossless zlib|zstd picks the
// payload kind the reference fixes at build time (default zlib, the kind every
// reference build can read).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "mgard_b200.h"

namespace {

int usage(const char *prog, int rc) {
  std::fprintf(rc ? stderr : stdout,
               "MGARD is a compressor for scientific data.\n\n"
               "Usage:\n"
               "  %s compress --datatype float|double --shape <n0xn1x...> --smoothness <s|inf>\n"
               "             --tolerance <tau> --input <filename> --output <filename> [--lossless zlib|zstd]\n"
               "  %s decompress --input <filename> --output <filename>\n"
               "  %s --version | --help\n",
               prog, prog, prog);
  return rc;
}

bool read_file(const std::string &name, std::vector<unsigned char> &buf) {
  std::ifstream in(name, std::ios_base::binary);
  if (!in) {
    std::cerr << "failed to open '" << name << "'" << std::endl;
    return false;
  }
  in.seekg(0, std::ios_base::end);
  const std::streamoff size = in.tellg();
  in.seekg(0, std::ios_base::beg);
  buf.resize((size_t)size);
  in.read(reinterpret_cast<char *>(buf.data()), size);
  if (!in) {
    std::cerr << "failed to read from '" << name << "'" << std::endl;
    return false;
  }
  return true;
}

bool write_file(const std::string &name, const void *data, size_t size) {
  std::ofstream out(name, std::ios_base::binary);
  if (!out) {
    std::cerr << "failed to open '" << name << "'" << std::endl;
    return false;
  }
  out.write(static_cast<const char *>(data), (std::streamsize)size);
  if (!out) {
    std::cerr << "failed to write to '" << name << "'" << std::endl;
    return false;
  }
  return true;
}

} // namespace

int main(int argc, char **argv) {
  if (argc < 2)
    return usage(argv[0], 0);
  const std::string sub = argv[1];
  if (sub == "--help" || sub == "-h")
    return usage(argv[0], 0);
  if (sub == "--version") {
    std::printf("%s (MGARD-CPU convention, file format 1.0.0)\n", mgb_version());
    return 0;
  }
  if (sub != "compress" && sub != "decompress") {
    std::cerr << "PARSE ERROR: Couldn't find match for argument '" << sub << "'" << std::endl;
    return usage(argv[0], 1);
  }
  std::map<std::string, std::string> opt;
  for (int i = 2; i < argc; i++) {
    const std::string key = argv[i];
    if (key == "--help" || key == "-h")
      return usage(argv[0], 0);
    if (key.rfind("--", 0) != 0 || i + 1 >= argc) {
      std::cerr << "PARSE ERROR: Couldn't find match for argument '" << key << "'" << std::endl;
      return 1;
    }
    opt[key.substr(2)] = argv[++i];
  }
  const std::vector<std::string> required =
      sub == "compress" ? std::vector<std::string>{"output", "input", "tolerance", "smoothness", "shape", "datatype"}
                        : std::vector<std::string>{"output", "input"};
  for (const std::string &r : required)
    if (!opt.count(r)) {
      std::cerr << "PARSE ERROR: Required argument not provided: --" << r << std::endl;
      return 1;
    }

  std::vector<unsigned char> in;
  if (!read_file(opt["input"], in))
    return 1;

  if (sub == "decompress") {
    void *out = nullptr;
    int ndim = 0, dtype = 0;
    uint64_t shape[MGB_MAX_DIMS];
    const int rc = mgb_cpu_decompress(in.data(), in.size(), &out, &ndim, shape, &dtype);
    if (rc) {
      std::cerr << "decompression failed (status " << rc << ")" << std::endl;
      return 1;
    }
    size_t bytes = dtype == MGB_F32 ? 4 : 8;
    for (int d = 0; d < ndim; d++)
      bytes *= shape[d];
    const bool ok = write_file(opt["output"], out, bytes);
    std::free(out);
    return ok ? 0 : 1;
  }

  const std::string datatype = opt["datatype"];
  if (datatype != "float" && datatype != "double") {
    std::cerr << "PARSE ERROR: --datatype must be float or double" << std::endl;
    return 1;
  }
  const int dtype = datatype == "float" ? MGB_F32 : MGB_F64;
  std::vector<uint64_t> shape;
  {
    std::istringstream stream(opt["shape"]); // 'x'-delimited list (src/cli/arguments.cpp:7-18)
    std::string token;
    while (std::getline(stream, token, 'x')) {
      uint64_t n = 0;
      std::istringstream(token) >> n;
      shape.push_back(n);
    }
  }
  if (shape.empty() || shape.size() > MGB_MAX_DIMS) {
    std::cerr << "unsupported dimension " << shape.size() << std::endl;
    return 1;
  }
  size_t expected = dtype == MGB_F32 ? 4 : 8;
  for (uint64_t n : shape)
    expected *= n;
  if (expected != in.size()) {
    std::cerr << "expected " << expected << " bytes, read " << in.size() << std::endl;
    return 1;
  }
  const std::string sstr = opt["smoothness"];
  const double s = (sstr == "inf" || sstr == "infinity") ? std::numeric_limits<double>::infinity()
                                                         : std::strtod(sstr.c_str(), nullptr);
  const double tolerance = std::strtod(opt["tolerance"].c_str(), nullptr);
  int compressor = 1;
  if (opt.count("lossless")) {
    if (opt["lossless"] == "zstd")
      compressor = 2;
    else if (opt["lossless"] != "zlib") {
      std::cerr << "PARSE ERROR: --lossless must be zlib or zstd" << std::endl;
      return 1;
    }
  }
  void *out = nullptr;
  size_t out_size = 0;
  const int rc = mgb_cpu_compress((int)shape.size(), dtype, shape.data(), nullptr, s, tolerance, compressor,
                                  in.data(), &out, &out_size);
  if (rc) {
    std::cerr << "compression failed (status " << rc << ")" << std::endl;
    return 1;
  }
  const bool ok = write_file(opt["output"], out, out_size);
  std::free(out);
  return ok ? 0 : 1;
}

// mgard-x-b200: command line front end with the options of the reference's
// `mgard-x` executable (reference src/mgard-x/Executables/mgard-x.cpp:25-413,
// doc/MGARD-X.md:103-127), on top of the B200 engine (include/mgard_b200/compress_x.hpp).
//
//   -z / --compress    -i <original> -o <compressed> -dt <s|d> -dim <D> <n_1> .. <n_D>
//                      -em <abs|rel> -e <tol> -s <smoothness|inf> [-l huffman] [-u <coords file>]
//                      [-dd max-dim [-dd-size <planes>]] [-d auto|cuda] [-v 0..3]
//   -x / --decompress  -i <compressed> -o <decompressed> [-d auto|cuda] [-v 0..3]
//
// Like the reference, compression mode decompresses again and prints the
// achieved error next to the requested bound.  Files are raw little-endian
// arrays (slowest dimension first); compressed files are the self-describing
// MGARD stream, interchangeable with the reference's.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <string>
#include <vector>

#include "mgard_b200/compress_x.hpp"

namespace {

[[noreturn]] void usage(const std::string &err) {
  if (!err.empty())
    std::cerr << "[ERR] " << err << "\n";
  std::printf(
      "Options\n"
      "\t -z / --compress: compress mode\n"
      "\t\t -i / --input <path to original data>\n"
      "\t\t -o / --output <path to compressed data>\n"
      "\t\t -dt / --data-type <s/single|d/double>\n"
      "\t\t -dim / --dimension <D> <n_1 (slowest)> ... <n_D (fastest)>\n"
      "\t\t -em / --error-bound-mode <abs|rel>\n"
      "\t\t -e / --error-bound <float>\n"
      "\t\t -s / --smoothness <float|inf>\n"
      "\t\t (optional) -u / --coordinates <path>: D coordinate arrays of the data type, concatenated\n"
      "\t\t (optional) -l / --lossless <huffman|huffman-zstd>\n"
      "\t\t (optional) -dd / --domain-decomposition <max-dim> [-dd-size <planes per sub-domain>]\n"
      "\t\t (optional) -d / --device <auto|cuda>\n"
      "\t\t (optional) -v / --verbose <0|1|2|3>\n"
      "\n"
      "\t -x / --decompress: decompress mode\n"
      "\t\t -i / --input <path to compressed data>\n"
      "\t\t -o / --output <path to decompressed data>\n"
      "\t\t (optional) -d / --device <auto|cuda>, -v / --verbose <0|1|2|3>\n");
  std::exit(err.empty() ? 0 : 2);
}

int find(int argc, char **argv, const char *a, const char *b) {
  for (int i = 1; i < argc; i++)
    if (!std::strcmp(argv[i], a) || !std::strcmp(argv[i], b))
      return i;
  return -1;
}
bool has(int argc, char **argv, const char *a, const char *b) { return find(argc, argv, a, b) >= 0; }
std::string arg(int argc, char **argv, const char *what, const char *a, const char *b) {
  int i = find(argc, argv, a, b);
  if (i < 0 || i + 1 >= argc)
    usage(std::string("missing option ") + a + " (" + what + ")");
  return argv[i + 1];
}
double to_double(const std::string &s, const char *what) {
  if (s == "inf" || s == "infinity" || s == "Inf")
    return std::numeric_limits<double>::infinity();
  try {
    return std::stod(s);
  } catch (...) {
    usage(std::string("illegal value for ") + what + ": " + s);
  }
}

std::vector<unsigned char> read_file(const std::string &path) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f)
    usage("cannot open " + path);
  std::streamsize n = f.tellg();
  f.seekg(0);
  std::vector<unsigned char> buf((size_t)n);
  if (n && !f.read((char *)buf.data(), n))
    usage("cannot read " + path);
  return buf;
}
void write_file(const std::string &path, const void *p, size_t n) {
  std::ofstream f(path, std::ios::binary);
  if (!f || !f.write((const char *)p, (std::streamsize)n))
    usage("cannot write " + path);
}

// error measures of the reference's CLI (mgard-x.cpp:95-140, ErrorCalculator.h:36-97)
template <typename T>
double achieved_error(const T *a, const T *b, size_t n, bool linf, bool rel) {
  if (linf) {
    double e = 0, m = 0;
    for (size_t i = 0; i < n; i++) {
      e = std::max(e, std::fabs((double)a[i] - (double)b[i]));
      m = std::max(m, std::fabs((double)a[i]));
    }
    return rel ? e / m : e;
  }
  double e = 0, m = 0;
  for (size_t i = 0; i < n; i++) {
    double d = (double)a[i] - (double)b[i];
    e += d * d;
    m += (double)a[i] * (double)a[i];
  }
  e = std::sqrt(e / n);
  m = std::sqrt(m / n);
  return rel ? e / m : e;
}

template <typename T>
int do_compress(int argc, char **argv, mgard_x::data_type dtype, int verbose) {
  const std::string in = arg(argc, argv, "original data", "-i", "--input");
  const std::string out = arg(argc, argv, "compressed data", "-o", "--output");
  int di = find(argc, argv, "-dim", "--dimension");
  if (di < 0 || di + 1 >= argc)
    usage("missing option -dim");
  const int D = std::atoi(argv[di + 1]);
  if (D < 1 || D > 5 || di + 1 + D >= argc)
    usage("-dim needs <D> followed by D sizes (1 <= D <= 5)");
  std::vector<mgard_x::SIZE> shape;
  size_t n = 1;
  for (int d = 0; d < D; d++) {
    shape.push_back((mgard_x::SIZE)std::strtoull(argv[di + 2 + d], nullptr, 10));
    n *= shape.back();
  }
  const std::string em = arg(argc, argv, "error bound mode", "-em", "--error-bound-mode");
  if (em != "abs" && em != "rel")
    usage("illegal error bound mode: " + em);
  const auto mode = em == "rel" ? mgard_x::error_bound_type::REL : mgard_x::error_bound_type::ABS;
  const double tol = to_double(arg(argc, argv, "error bound", "-e", "--error-bound"), "-e");
  const double s = to_double(arg(argc, argv, "smoothness", "-s", "--smoothness"), "-s");
  mgard_x::Config config;
  if (has(argc, argv, "-l", "--lossless")) {
    const std::string l = arg(argc, argv, "lossless", "-l", "--lossless");
    if (l == "huffman-zstd")
      config.lossless = mgard_x::lossless_type::Huffman_Zstd;
    else if (l != "huffman")
      usage("-l huffman | huffman-zstd (huffman-lz4 needs nvcomp and is not built)");
  }
  if (has(argc, argv, "-dd", "--domain-decomposition")) {
    if (arg(argc, argv, "domain decomposition", "-dd", "--domain-decomposition") != "max-dim")
      usage("only -dd max-dim is available in this build");
    if (has(argc, argv, "-dd-size", "--domain-decomposition-size"))
      config.domain_decomposition_size = (mgard_x::SIZE)std::strtoull(
          arg(argc, argv, "size", "-dd-size", "--domain-decomposition-size").c_str(), nullptr, 10);
  }
  std::vector<unsigned char> file = read_file(in);
  if (file.size() != n * sizeof(T))
    std::cerr << "[WARN] input file size mismatch " << file.size() << " vs. " << n * sizeof(T) << "!\n";
  // like the reference: a short file is repeated until the array is full
  std::vector<T> u(n);
  if (file.size() < sizeof(T))
    usage("input file holds no data");
  const size_t have = file.size() / sizeof(T);
  for (size_t done = 0; done < n;) {
    size_t c = std::min(have, n - done);
    std::memcpy(u.data() + done, file.data(), c * sizeof(T));
    done += c;
  }
  std::vector<std::vector<T>> coord_store;
  std::vector<const mgard_x::Byte *> coords;
  if (has(argc, argv, "-u", "--coordinates")) {
    std::vector<unsigned char> cf = read_file(arg(argc, argv, "coordinates", "-u", "--coordinates"));
    size_t need = 0;
    for (auto v : shape)
      need += v;
    if (cf.size() != need * sizeof(T))
      usage("coordinate file must hold sum(n_d) values of the data type");
    size_t off = 0;
    for (int d = 0; d < D; d++) {
      coord_store.emplace_back((const T *)cf.data() + off, (const T *)cf.data() + off + shape[d]);
      off += shape[d];
    }
    for (auto &c : coord_store)
      coords.push_back((const mgard_x::Byte *)c.data());
  }
  mgard_x::pin_memory(u.data(), n * sizeof(T), config); // as the reference's CLI does
  void *compressed = nullptr;
  size_t compressed_size = 0;
  auto t0 = std::chrono::steady_clock::now();
  mgard_x::compress_status_type st =
      coords.empty()
          ? mgard_x::compress((mgard_x::DIM)D, dtype, shape, tol, s, mode, u.data(), compressed,
                              compressed_size, config, false)
          : mgard_x::compress((mgard_x::DIM)D, dtype, shape, tol, s, mode, u.data(), compressed,
                              compressed_size, coords, config, false);
  auto t1 = std::chrono::steady_clock::now();
  if (st != mgard_x::compress_status_type::Success) {
    std::cerr << "[ERR] Compression failed (status " << (int)st << ")\n";
    return 1;
  }
  write_file(out, compressed, compressed_size);
  std::cout << "[INFO] Compression ratio: " << (double)(n * sizeof(T)) / compressed_size << "\n";
  void *back = nullptr;
  auto t2 = std::chrono::steady_clock::now();
  st = mgard_x::decompress(compressed, compressed_size, back, config, false);
  auto t3 = std::chrono::steady_clock::now();
  if (st != mgard_x::compress_status_type::Success) {
    std::cerr << "[ERR] Decompression failed (status " << (int)st << ")\n";
    return 1;
  }
  const bool linf = std::isinf(s) && s > 0, rel = mode == mgard_x::error_bound_type::REL;
  const double err = achieved_error<T>(u.data(), (const T *)back, n, linf, rel);
  std::cout << std::scientific << "[INFO] " << (rel ? "Relative " : "Absolute ") << (linf ? "L_inf" : "L_2")
            << " error: " << err << " (" << (err < tol ? "Satisfied" : "Not Satisfied") << ")\n";
  if (verbose >= 2) {
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    std::cout << std::fixed << "[TIME] compress " << ms(t0, t1) << " ms, decompress " << ms(t2, t3)
              << " ms (host buffers, transfers included)\n";
  }
  mgard_x::unpin_memory(u.data(), config);
  std::free(compressed);
  std::free(back);
  mgard_x::release_cache(config);
  return err < tol ? 0 : 3;
}

int do_decompress(int argc, char **argv) {
  const std::string in = arg(argc, argv, "compressed data", "-i", "--input");
  const std::string out = arg(argc, argv, "decompressed data", "-o", "--output");
  std::vector<unsigned char> file = read_file(in);
  mgard_x::Config config;
  std::vector<mgard_x::SIZE> shape;
  mgard_x::data_type dtype;
  void *back = nullptr;
  auto st = mgard_x::decompress(file.data(), file.size(), back, shape, dtype, config, false);
  if (st != mgard_x::compress_status_type::Success) {
    std::cerr << "[ERR] Decompression failed (status " << (int)st << ")\n";
    return 1;
  }
  size_t n = 1;
  for (auto v : shape)
    n *= v;
  write_file(out, back, n * (dtype == mgard_x::data_type::Double ? 8 : 4));
  std::free(back);
  mgard_x::release_cache(config);
  return 0;
}

} // namespace

int main(int argc, char **argv) {
  int verbose = 0;
  if (has(argc, argv, "-v", "--verbose"))
    verbose = std::atoi(arg(argc, argv, "verbose", "-v", "--verbose").c_str());
  if (has(argc, argv, "-d", "--device")) {
    const std::string d = arg(argc, argv, "device", "-d", "--device");
    if (d != "auto" && d != "cuda")
      usage("this build has one backend: -d auto|cuda");
  }
  if (has(argc, argv, "-z", "--compress")) {
    const std::string dt = arg(argc, argv, "data type", "-dt", "--data-type");
    if (dt == "s" || dt == "single")
      return do_compress<float>(argc, argv, mgard_x::data_type::Float, verbose);
    if (dt == "d" || dt == "double")
      return do_compress<double>(argc, argv, mgard_x::data_type::Double, verbose);
    usage("illegal data type: " + dt);
  }
  if (has(argc, argv, "-x", "--decompress"))
    return do_decompress(argc, argv);
  usage("");
}

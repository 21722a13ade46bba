// Reads like the reference's README example for mgard::compress / decompress
// (reference README.md "Basic usage", tests/src/test_compress.cpp:75-118), built
// against include/mgard_b200/compress.hpp instead of <compress.hpp>.
#include <cmath>
#include <cstdio>
#include <limits>
#include <sstream>
#include <vector>

#include "mgard_b200/compress.hpp"

int main() {
  const std::array<std::size_t, 3> shape = {33, 20, 17};
  std::array<std::vector<double>, 3> coords;
  for (std::size_t d = 0; d < 3; ++d) {
    coords[d].resize(shape[d]);
    for (std::size_t i = 0; i < shape[d]; ++i)
      coords[d][i] = std::pow((double)i / (shape[d] - 1), 1.25);
  }
  const mgard::TensorMeshHierarchy<3, double> hierarchy(shape, coords);
  const std::size_t ndof = hierarchy.ndof();
  std::vector<double> u(ndof);
  for (std::size_t i = 0; i < ndof; ++i)
    u[i] = std::sin(0.01 * i) + 0.3 * std::cos(0.37 * i);
  const double s = std::numeric_limits<double>::infinity(), tolerance = 1e-3;
  const mgard::CompressedDataset<3, double> compressed = mgard::compress(hierarchy, u.data(), s, tolerance);
  std::printf("L %zu ndof %zu payload %zu bytes\n", hierarchy.L, ndof, compressed.size());
  const mgard::DecompressedDataset<3, double> decompressed = mgard::decompress(compressed);
  double err = 0;
  for (std::size_t i = 0; i < ndof; ++i)
    err = std::fmax(err, std::fabs(decompressed.data()[i] - u[i]));
  std::printf("max error %.3e (tolerance %.1e)\n", err, tolerance);
  std::ostringstream os;
  compressed.write(os);
  const std::string blob = os.str();
  const mgard::MemoryBuffer<const unsigned char> raw = mgard::decompress(blob.data(), blob.size());
  const bool same = raw.size == ndof * sizeof(double) &&
                    std::memcmp(raw.data.get(), decompressed.data(), raw.size) == 0;
  std::printf("self-describing decompress %s\n", same ? "identical" : "DIFFERS");
  bool threw = false;
  try {
    mgard::TensorMeshHierarchy<2, float> bad({1, 1});
  } catch (const std::domain_error &) {
    threw = true;
  }
  std::printf("degenerate shape %s\n", threw ? "rejected" : "ACCEPTED");
  return (err <= tolerance && same && threw) ? 0 : 1;
}

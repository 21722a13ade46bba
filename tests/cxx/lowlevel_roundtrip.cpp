// The low-level example of the reference's documentation (doc/MGARD-X.md:205-262), compiled
// against include/mgard_b200/compress_x_lowlevel.hpp instead of <compress_x_lowlevel.hpp>.
#include <cmath>
#include <cstdio>
#include <vector>

#include "mgard_b200/compress_x_lowlevel.hpp"

int main() {
  const mgard_x::SIZE n1 = 33, n2 = 40, n3 = 65;
  std::vector<mgard_x::SIZE> shape{n1, n2, n3};
  std::vector<float> u(n1 * n2 * n3);
  for (size_t i = 0; i < n1; i++)
    for (size_t j = 0; j < n2; j++)
      for (size_t k = 0; k < n3; k++)
        u[(i * n2 + j) * n3 + k] = std::sin(3.0f * (float)i / n1) * std::cos(2.0f * (float)j / n2) + 0.5f * (float)k / n3;
  float umax = 0;
  for (float x : u)
    umax = std::fmax(umax, std::fabs(x));
  mgard_x::Config config;
  mgard_x::Hierarchy<3, float, mgard_x::CUDA> hierarchy(shape, config);
  mgard_x::Compressor<3, float, mgard_x::CUDA> compressor(hierarchy, config);
  mgard_x::Array<3, float, mgard_x::CUDA> in_array(shape);
  in_array.load(u.data());
  mgard_x::Array<1, unsigned char, mgard_x::CUDA> compressed_array;
  const float tol = 1e-3f, s = INFINITY;
  float norm = 0;
  compressor.Compress(in_array, mgard_x::error_bound_type::REL, tol, s, norm, compressed_array, 0);
  mgard_x::DeviceRuntime<mgard_x::CUDA>::SyncQueue(0);
  std::printf("l_target %llu compressed %llu bytes norm %g\n", (unsigned long long)hierarchy.l_target(),
              (unsigned long long)compressed_array.shape(0), norm);
  mgard_x::Array<3, float, mgard_x::CUDA> out_array;
  compressor.Decompress(compressed_array, mgard_x::error_bound_type::REL, tol, s, norm, out_array, 0);
  mgard_x::DeviceRuntime<mgard_x::CUDA>::SyncQueue(0);
  float *back = out_array.hostCopy();
  float err = 0;
  for (size_t i = 0; i < u.size(); i++)
    err = std::fmax(err, std::fabs(back[i] - u[i]));
  std::printf("max error %g bound %g\n", err, tol * umax);
  if (norm != umax || err > tol * umax || compressed_array.shape(0) >= u.size() * 4)
    return 1;
  // the input array was not altered
  float *again = in_array.hostCopy();
  for (size_t i = 0; i < u.size(); i++)
    if (again[i] != u[i])
      return 2;
  std::printf("lowlevel ok\n");
  return 0;
}

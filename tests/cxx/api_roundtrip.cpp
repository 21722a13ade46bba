#include "mgard_b200/compress_x.hpp"
#include <cmath>
#include <cstdio>
#include <cstdlib>
int main() {
  std::vector<mgard_x::SIZE> shape = {33, 34, 35};
  size_t n = 33 * 34 * 35;
  std::vector<float> u(n);
  for (size_t i = 0; i < n; i++) u[i] = std::sin(0.01 * i);
  void *out = nullptr; size_t sz = 0;
  auto st = mgard_x::compress(3, mgard_x::data_type::Float, shape, 1e-3, INFINITY,
                              mgard_x::error_bound_type::REL, u.data(), out, sz, false);
  printf("compress status %d size %zu\n", (int)st, sz);
  if (st != mgard_x::compress_status_type::Success) return (int)st == 5 ? 0 : 1;
  void *back = nullptr; std::vector<mgard_x::SIZE> shp; mgard_x::data_type dt;
  st = mgard_x::decompress(out, sz, back, shp, dt, false);
  double err = 0; for (size_t i = 0; i < n; i++) err = std::fmax(err, std::fabs(((float*)back)[i] - u[i]));
  printf("decompress status %d dims %zu err %g\n", (int)st, shp.size(), err);
  free(out); free(back);
  return err <= 1e-3 ? 0 : 1;
}

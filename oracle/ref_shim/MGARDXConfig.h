/* Stand-in for the file the reference's CMake would generate from
 * include/MGARDXConfig.h.in: SERIAL backend only. Test infrastructure. */
#ifndef MGARD_X_CONFIG_H
#define MGARD_X_CONFIG_H
#define MGARD_ENABLE_SERIAL 1
#define MGARD_ENABLE_OPENMP 0
#define MGARD_ENABLE_CUDA 0
#define MGARD_ENABLE_HIP 0
#define MGARD_ENABLE_SYCL 0
#define MGARD_ENABLE_LEGACY_CUDA 0
#define MGARD_ENABLE_AUTO_TUNING 0
#define MGARD_ENABLE_EXTERNAL_COMPRESSOR 0
#endif

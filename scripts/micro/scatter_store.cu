// Microbenchmark (developer tool): one thread per 80 KB chunk writes its chunk front to back.
//   mode 0: 32-byte stores (st.v8.f32), one per iteration
//   mode 1: four 32-byte stores per iteration (a full 128-byte line)
//   mode 2: as 0, chunk of thread = interleaved over warps (lanes far apart)
//   mode 3: as 0 with a dependent ALU chain of `work` instructions between stores
//   mode 4: as 0, the lanes of a warp take every S-th chunk (S = work): a warp spans 32 * S chunks
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a scatter_store.cu -o scatter_store
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void k(float *out, long long nchunk, int chunk, int mode, int work) {
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long c = gid;
  if (mode == 2) {
    long long nwarp = (long long)gridDim.x * (blockDim.x >> 5);
    c = (long long)(threadIdx.x & 31) * nwarp + (long long)(threadIdx.x >> 5) * gridDim.x + blockIdx.x;
  }
  if (mode == 4) {
    long long wg = gid >> 5, S = work;
    c = (wg / S) * (32 * S) + (long long)(threadIdx.x & 31) * S + (wg % S);
    work = 0;
  }
  if (c >= nchunk) return;
  float *dst = out + c * chunk;
  float v = (float)c;
  unsigned x = (unsigned)c * 2654435761u;
  int step = mode == 1 ? 32 : 8;
  for (int i = 0; i < chunk; i += step) {
    for (int w = 0; w < work; w++) x = x * 1664525u + 1013904223u;   // dependent chain
    v += (float)(x >> 31);
    for (int j = 0; j < step; j += 8)
      asm volatile("st.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(dst + i + j), "f"(v) : "memory");
  }
}
int main(int argc, char **argv) {
  long long nchunk = 52686; int chunk = 20480;
  float *out; cudaMalloc(&out, nchunk * chunk * 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int mode = 0; mode < 4; mode++) for (int threads : {128, 256, 384}) for (int work : {0, 50, 200}) {
    if (mode != 3 && work) continue;
    if (mode == 3 && !work) continue;
    int blocks = (int)((nchunk + threads - 1) / threads);
    k<<<blocks, threads>>>(out, nchunk, chunk, mode, work); cudaDeviceSynchronize();
    cudaEventRecord(a); k<<<blocks, threads>>>(out, nchunk, chunk, mode, work); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("mode %d threads %d work %d: %.3f ms  %.0f GB/s\n", mode, threads, work, ms, nchunk * chunk * 4.0 / ms / 1e6);
  }
  for (int S : {1, 2, 4, 8, 16, 32, 64}) {
    int threads = 384, blocks = (int)((nchunk + threads - 1) / threads) + 8;
    k<<<blocks, threads>>>(out, nchunk, chunk, 4, S); cudaDeviceSynchronize();
    cudaEventRecord(a); k<<<blocks, threads>>>(out, nchunk, chunk, 4, S); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("mode 4 S %d: %.3f ms  %.0f GB/s\n", S, ms, nchunk * chunk * 4.0 / ms / 1e6);
  }
  return 0;
}

// L2 -> SM read bandwidth of a B200 (microbenchmark): every CTA streams a buffer that fits L2 with 16-byte loads.
//   mode 0: CTA b starts at its own offset (all SMs read different lines at any moment)
//   mode 1: groups of 4 CTAs read the same lines at the same time (the kernel-spectrum pattern of the fused kernel)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void rd(const float4* __restrict__ p, size_t n4, int iters, int mode, float* out) {
  float4 acc = make_float4(0, 0, 0, 0);
  const size_t per = n4 / gridDim.x;
  const int b = mode == 1 ? (blockIdx.x / 4) * 4 : blockIdx.x;
  for (int it = 0; it < iters; ++it) {
    size_t base = ((size_t)b * per + (size_t)it * per * 7) % n4;
    for (size_t i = threadIdx.x; i < per; i += blockDim.x * 4) {
      size_t j0 = (base + i) % n4, j1 = (base + i + blockDim.x) % n4, j2 = (base + i + 2 * blockDim.x) % n4, j3 = (base + i + 3 * blockDim.x) % n4;
      float4 a = __ldg(p + j0), c = __ldg(p + j1), d = __ldg(p + j2), e = __ldg(p + j3);
      acc.x += a.x + c.x + d.x + e.x; acc.y += a.y + c.y + d.y + e.y; acc.z += a.z + c.z; acc.w += a.w + e.w;
    }
  }
  if (acc.x + acc.y + acc.z + acc.w == 12345.f) out[0] = acc.x;
}
int main(int argc, char** argv) {
  for (size_t mb : {16, 48, 96, 512}) {
    size_t bytes = mb << 20, n4 = bytes / 16;
    float4* p; float* out;
    cudaMalloc(&p, bytes); cudaMalloc(&out, 4); cudaMemset(p, 0, bytes);
    for (int mode = 0; mode < 2; ++mode)
      for (int cps : {2, 4, 8}) {
        int grid = 148 * cps, iters = 24;
        rd<<<grid, 256>>>(p, n4, 2, mode, out);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        rd<<<grid, 256>>>(p, n4, iters, mode, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double tb = (double)(n4 / grid) * grid * 16.0 * iters / (ms * 1e-3) / 1e12;
        printf("buffer %4zu MiB mode %d ctas/sm %d: %.2f TB/s (%.3f ms) %s\n", mb, mode, cps, tb, ms, cudaGetErrorString(cudaGetLastError()));
      }
    cudaFree(p); cudaFree(out);
  }
  return 0;
}

// cuda_shim.h — TEST INFRASTRUCTURE ONLY. A minimal host stand-in for the CUDA execution model so that the
// generic kernels of fft_conv_pytorch_b200/csrc/fc_kernels.cuh can be run on CPU threads by the `-m "not gpu"`
// tests (there is no GPU in the authoring container). One OS thread per CUDA thread of a block, blocks run
// one after another, __syncthreads() is a real barrier. Never linked into the shipped library.
#pragma once
#include <cmath>
using std::fmaf;
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __launch_bounds__(...)

struct float2 {
  float x, y;
};
struct float4 {
  float x, y, z, w;
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

extern thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
extern char* fc_emul_smem;
void __syncthreads();
void __syncwarp();  // barrier over the 32 host threads standing in for one warp
void fc_emul_launch(dim3 grid, dim3 block, size_t smem, std::function<void()> body);

#define FC_DYN_SMEM(name) float2* name = reinterpret_cast<float2*>(fc_emul_smem)

static inline unsigned __float_as_uint(float f) {
  unsigned u;
  std::memcpy(&u, &f, 4);
  return u;
}
static inline float __uint_as_float(unsigned u) {
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
template <class T>
static inline T __ldg(const T* p) {
  return *p;
}
static inline void sincospif(float x, float* s, float* c) {
  *s = (float)std::sin(M_PI * (double)x);
  *c = (float)std::cos(M_PI * (double)x);
}
static inline void sincospi(double x, double* s, double* c) {
  *s = std::sin(M_PI * x);
  *c = std::cos(M_PI * x);
}

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2 };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) {
  std::memset(d, v, n);
  return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
  std::memcpy(d, s, n);
  return cudaSuccess;
}

// Micro benchmark of the streaming K1 / K4 (csrc/fc_stream.cuh) against the register-path kernels at the BASELINE c2 geometry:
// bitwise comparison + CUDA-event timing with the L2 flushed between launches. Development tool (quick rebuilds, ablations):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DFC_PACKED_F32X2 [-DFC_STREAM_ABL=n] -I fft_conv_pytorch_b200/csrc \
//        scripts/micro/stream_bench.cu -o scripts/micro/stream_bench
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "fc_kernels.cuh"
#include "fc_fused.cuh"
#include "fc_tc.cuh"
#include "fc_stream.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                           const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool encode(CUtensorMap* tm, void* base, int64_t rows, int64_t bins, int64_t items, int64_t bs, int64_t is) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)(2 * rows), (cuuint64_t)bins, (cuuint64_t)items};
  const cuuint64_t strides[2] = {(cuuint64_t)bs * 8, (cuuint64_t)is * 8};
  const cuuint32_t box[3] = {32, 256, 1}, es[3] = {1, 1, 1};
  return ((enc_fn)p)(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__global__ void read_sweep(const float4* p, size_t n, float* sink) {
  float acc = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) acc += p[i].x;
  if (acc == 123.456f) *sink = acc;
}
static int g_flush_mode = 0;  // 0: memset (leaves L2 full of dirty lines), 1: memset + read sweep of a second buffer (clean lines)
static void* g_flush2 = nullptr;

template <typename F>
static float time_it(F launch, void* flush, size_t flush_bytes, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  std::vector<float> ms;
  for (int i = 0; i < reps + 3; ++i) {
    cudaMemsetAsync(flush, i, flush_bytes);
    if (g_flush_mode == 1) read_sweep<<<1184, 256>>>((const float4*)g_flush2, flush_bytes / 16, (float*)flush);
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float m;
    cudaEventElapsedTime(&m, e0, e1);
    if (i >= 3) ms.push_back(m);
  }
  std::sort(ms.begin(), ms.end());
  return ms[ms.size() / 2] * 1e3f;
}

int main(int argc, char** argv) {
  const int occ = argc > 1 ? atoi(argv[1]) : 3;
  const int whole = argc > 2 ? atoi(argv[2]) : 1;
  g_flush_mode = argc > 3 ? atoi(argv[3]) : 0;
  const int B = 64, R = 512, Nx = 512, M = 256, Rout = 448, Lout = 448;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float *x, *y1, *y2, *bias;
  float2 *s1, *s2, *tw, *ys;
  void* flush;
  const size_t flush_bytes = 256u << 20;
  const size_t xs = (size_t)B * R * Nx, ss = (size_t)B * 257 * R, yss = (size_t)B * 257 * Rout, os = (size_t)B * Rout * Lout;
  CK(cudaMalloc(&x, xs * 4)); CK(cudaMalloc(&s1, ss * 8)); CK(cudaMalloc(&s2, ss * 8)); CK(cudaMalloc(&tw, 512 * 8));
  CK(cudaMalloc(&ys, yss * 8)); CK(cudaMalloc(&y1, os * 4)); CK(cudaMalloc(&y2, os * 4)); CK(cudaMalloc(&bias, 8 * 4)); CK(cudaMalloc(&flush, flush_bytes)); CK(cudaMalloc(&g_flush2, flush_bytes)); CK(cudaMemset(g_flush2, 0, flush_bytes));
  {
    std::vector<float> h(std::max(xs, yss * 2));
    unsigned st = 12345;
    for (auto& v : h) { st = st * 1664525u + 1013904223u; v = ((st >> 8) & 0xffff) / 32768.f - 1.f; }
    CK(cudaMemcpy(x, h.data(), xs * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ys, h.data(), yss * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(bias, h.data(), 32, cudaMemcpyHostToDevice));
  }
  fc_twiddle_kernel<<<2, 256>>>(tw, 512, 0, 1.0);
  CK(cudaMemset(s1, 0xff, ss * 8)); CK(cudaMemset(s2, 0xff, ss * 8)); CK(cudaMemset(y1, 0xff, os * 4)); CK(cudaMemset(y2, 0xff, os * 4));

  fc_pass p;
  memset(&p, 0, sizeof(p));
  p.kind = FC_R2C; p.N = 512; p.M = M; p.T = 16; p.tw_len = 512; p.seg_n = 1; p.n_outer = B; p.R = R; p.tiles_per_outer = R / 16; p.n_tiles = (int64_t)B * R / 16;
  p.in_rs = Nx; p.in_es = 1; p.o_c2 = 1; p.o_q = 1; p.o_sA = (int64_t)R * Nx; p.out_os = 257 * R; p.out_rs = 1; p.out_es = R;
  p.imap.L = Nx; p.imap.pad = 0; p.imap.up = 1; p.imap.sub = 1; p.imap.ext = Nx; p.scale = 1.f; p.row_og = 1;
  fc_pass p4;
  memset(&p4, 0, sizeof(p4));
  p4.kind = FC_C2R; p4.N = 512; p4.M = M; p4.T = 16; p4.tw_len = 512; p4.seg_n = 1; p4.n_outer = B; p4.R = Rout; p4.tiles_per_outer = Rout / 16; p4.n_tiles = (int64_t)B * Rout / 16;
  p4.in_os = 257 * Rout; p4.in_rs = 1; p4.in_es = Rout; p4.out_os = (int64_t)Rout * Lout; p4.out_rs = Lout; p4.out_es = 1;
  p4.omap.Lout = Lout; p4.omap.os = 1; p4.omap.ob = 0; p4.omap.og = 1; p4.omap.lim = 512; p4.row_og = 1; p4.row_Lout = Rout; p4.cout = 8; p4.has_bias = 1; p4.scale = 1.f;

  const int smem_old = 16 * 257 * 8;
  auto k1 = fc_fast_r2c_kernel<256, 2, 8, 4>;
  auto k4 = fc_fast_c2r_kernel<256, 2, 8, 4>;
  CK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_old));
  CK(cudaFuncSetAttribute(k4, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_old));
#ifndef NSLOTS
#define NSLOTS 2
#endif
  constexpr int NS = NSLOTS;
  constexpr int kSm = fc_stream::smem_bytes(NS);
  CK(cudaFuncSetAttribute(fc_stream_r2c_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSm));
  CK(cudaFuncSetAttribute(fc_stream_c2r_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSm));
  fc_fast_r2c_args a1{p, x, s1, tw};
  fc_fast_c2r_args a4{p4, ys, y1, tw, bias};
  fc_stream_r2c_args b1;
  b1.p = p; b1.x = x; b1.out = s2; b1.tw = tw; b1.whole_tiles = whole;
  if (!encode(&b1.tmap, s2, R, 257, B, p.out_es, p.out_os)) { printf("tensor map encode failed\n"); return 1; }
  fc_stream_c2r_args b4;
  b4.p = p4; b4.in = ys; b4.out = y2; b4.tw = tw; b4.bias = bias;
  if (!encode(&b4.tmap, ys, Rout, 257, B, p4.in_es, p4.in_os)) { printf("tensor map encode failed\n"); return 1; }

  const int reps = 21;
  float t_k1 = time_it([&] { k1<<<sms * 4, 256, smem_old>>>(a1); }, flush, flush_bytes, reps);
  float t_k1s = time_it([&] { fc_stream_r2c_kernel<NS><<<sms * occ, 256, kSm>>>(b1); }, flush, flush_bytes, reps);
  float t_k4 = time_it([&] { k4<<<sms * 4, 256, smem_old>>>(a4); }, flush, flush_bytes, reps);
  float t_k4s = time_it([&] { fc_stream_c2r_kernel<NS><<<sms * occ, 256, kSm>>>(b4); }, flush, flush_bytes, reps);
  CK(cudaDeviceSynchronize());
  std::vector<float> h1(ss * 2), h2(ss * 2), g1(os), g2(os);
  CK(cudaMemcpy(h1.data(), s1, ss * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h2.data(), s2, ss * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(g1.data(), y1, os * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(g2.data(), y2, os * 4, cudaMemcpyDeviceToHost));
  size_t d1 = 0, d4 = 0;
  for (size_t i = 0; i < h1.size(); ++i) d1 += memcmp(&h1[i], &h2[i], 4) != 0;
  for (size_t i = 0; i < g1.size(); ++i) d4 += memcmp(&g1[i], &g2[i], 4) != 0;
  printf("NS=%d ABL=%d occ=%d whole=%d flush=%d K1 %.1f us  K1stream %.1f us (%zu words differ)   K4 %.1f us  K4stream %.1f us (%zu words differ)\n", NS, FC_STREAM_ABL, occ, whole, g_flush_mode, t_k1, t_k1s, d1,
         t_k4, t_k4s, d4);
  return 0;
}

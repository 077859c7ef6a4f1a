// fc_api.cu — the C ABI of include/fftconv_b200.h: plan management and kernel launches.
// Compiled by nvcc for sm_100a into libfftconv_b200.so. (tests/cpu_emul compiles the same file with
// -DFC_CPU_EMUL as host C++ to exercise the kernel logic without a GPU; that build is test-only.)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "fc_kernels.cuh"
#include "fc_fused.cuh"
#include "fc_pair.cuh"
#include "fc_column.cuh"
#include "fc_plane.cuh"
#include "fc_line.cuh"
#include "fc_tc.cuh"
#include "fc_stream.cuh"
#include "fc_plan.h"
#include "fc_tune.h"

#ifdef FC_CPU_EMUL
#define FC_LAUNCH(kfn, grid, block, smem, stream, arg) fc_emul_launch(grid, block, smem, [=]() { kfn(arg); })
#else
// Every kernel launched here starts with fc_grid_dep_sync() (fc_kernels.cuh), so it may be launched as a programmatic
// dependent of the kernel before it in the stream: its CTAs become resident while that kernel drains, and the launch
// latency, CTA start-up and (inside a captured graph) the kernel-boundary drain overlap with the predecessor's tail.
// (Tuning builds: PDL=0 restores plain stream-ordered launches for A/B timing.)
inline bool fc_pdl_enabled() {
  static const bool on = fc_tune_int("PDL", 1) != 0;
  return on;
}
template <typename K, typename A>
inline void fc_launch_pdl(K kfn, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, const A& arg) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = fc_pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kfn, arg);
}
#define FC_LAUNCH(kfn, grid, block, smem, stream, arg) fc_launch_pdl(kfn, grid, block, smem, stream, arg)
#endif

namespace {

thread_local std::string g_err;

// Optional per-launch event recorder (fc_conv_profiled).
#ifndef FC_CPU_EMUL
struct Recorder {
  std::vector<cudaEvent_t> ev;
  cudaStream_t st;
};
thread_local Recorder* g_rec = nullptr;
void rec_mark() {
  if (!g_rec) return;
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, g_rec->st);
  g_rec->ev.push_back(e);
}
#else
void rec_mark() {}
#endif

int set_err(int code, const std::string& m) {
  g_err = m;
  return code;
}

int check_cuda(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err((int)e, std::string(what) + ": " + cudaGetErrorString(e));
  return FC_OK;
}

const int kMaxSmem = 200 * 1024;
const char* const kSegMsg =
    "this plan splits the first axis into overlap-save segments and runs through fc_conv only; create the plan with FC_FLAG_NO_SEGMENT "
    "for the stage calls";
thread_local int g_num_sms = 148;  // SM count of the calling thread's current device (set by init_once)

// every instantiation of the fused kernel: (N, CI, NB, warps, plain, CTAs per SM)
#define FC_FUSED_ALL(X) \
  X(256, 8, 2, 8, true, 2) X(256, 8, 2, 8, false, 2) X(256, 8, 1, 8, true, 2) X(256, 8, 1, 8, false, 2) \
  X(512, 8, 2, 8, true, 2) X(512, 8, 2, 8, false, 2) X(512, 8, 1, 8, true, 2) X(512, 8, 1, 8, false, 2) \
  X(512, 8, 2, 8, true, 3) X(512, 8, 1, 8, true, 3) X(512, 8, 1, 8, true, 4) \
  X(512, 8, 2, 4, true, 3) X(512, 8, 1, 4, true, 3) X(256, 8, 2, 4, true, 3) \
  X(1024, 8, 1, 8, true, 2) X(1024, 8, 1, 8, false, 2) \
  X(256, 16, 2, 8, true, 2) X(256, 16, 2, 8, false, 2) X(256, 16, 1, 8, true, 2) X(256, 16, 1, 8, false, 2) \
  X(512, 16, 1, 8, true, 2) X(512, 16, 1, 8, false, 2) X(1024, 16, 1, 8, true, 1) X(1024, 16, 1, 8, false, 1)

// Instantiated variants of the transposing kernels K1 / K4: X(M, lines per warp, warps, CTAs per SM).
#define FC_FAST_ALL(X) \
  X(32, 2, 8, 4) X(64, 2, 8, 4) X(128, 2, 8, 4) \
  X(256, 2, 8, 3) X(256, 2, 8, 4) X(256, 1, 16, 2) X(512, 2, 8, 1) X(512, 2, 8, 2) X(512, 2, 8, 3) X(512, 1, 16, 1) X(512, 1, 16, 2) X(1024, 1, 16, 1)

// ... and of the contiguous complex pass K2 / K3: X(N, lines per warp, warps, CTAs per SM).
#define FC_FAST_C2C_ALL(X) X(32, 2, 8, 4) X(64, 2, 8, 4) X(128, 2, 8, 4) X(256, 2, 8, 3) X(512, 2, 8, 2) X(1024, 1, 8, 2) X(2048, 1, 8, 1)

// ... and of the two-axis plane kernels: X(NY, NZ).
#define FC_PLANE_ALL(X) X(32, 32) X(32, 64) X(64, 32) X(64, 64) X(32, 128) X(64, 128) X(128, 32) X(128, 64) X(128, 128)

// ... of the pair pipeline (fc_pair.cuh). K1p / K4p: X(M, pair lines per warp group, warps, CTAs per SM);
// KBp: X(N, channels per group, pair items per CTA, warps, plain, CTAs per SM).
#define FC_PAIR_ROW_ALL(X) X(128, 1, 8, 4) X(128, 2, 8, 2) X(256, 1, 8, 4) X(256, 2, 8, 2) X(512, 1, 8, 2) X(1024, 1, 8, 1)
#define FC_PAIR_FUSED_ALL(X) \
  X(256, 8, 2, 8, true, 2, 32) X(256, 8, 2, 8, false, 2, 32) X(256, 8, 1, 8, true, 2, 32) X(256, 8, 1, 8, false, 2, 32) \
  X(512, 8, 1, 8, true, 2, 32) X(512, 8, 1, 8, false, 2, 32) X(512, 8, 1, 16, true, 2, 64) X(512, 8, 1, 16, false, 2, 64) \
  X(512, 8, 2, 16, true, 1, 32) X(512, 8, 2, 16, false, 1, 32) X(1024, 8, 1, 8, true, 1, 32) X(1024, 8, 1, 8, false, 1, 32) \
  X(1024, 8, 1, 16, true, 1, 64) X(1024, 8, 1, 16, false, 1, 64) \
  X(256, 16, 1, 8, true, 2, 32) X(256, 16, 1, 8, false, 2, 32) X(512, 16, 1, 8, true, 1, 32) X(512, 16, 1, 8, false, 1, 32) \
  X(512, 16, 1, 16, true, 1, 64) X(512, 16, 1, 16, false, 1, 64)

// y-stage variants of K1p / K4p (16-line tiles): X(M, pair lines per warp group, warps, CTAs per SM, radix)
#define FC_PAIR_ROW_YS_ALL(X) X(128, 1, 8, 4, 8) X(128, 1, 8, 4, 4) X(256, 1, 16, 2, 8) X(256, 1, 16, 2, 4)
// fc_pair_fused64_kernel: X(sub-transform length, channels per group, CTAs per SM)
#define FC_PAIR_FUSED64_ALL(X) X(64, 8, 4) X(64, 8, 3) X(64, 8, 2) X(128, 8, 2) X(128, 8, 1)

void fused_set_attr() {
#ifndef FC_CPU_EMUL
#define FC_FUSED_ATTR(NN, CC, NBB, WW, PL, OC) \
  cudaFuncSetAttribute(fc_fused_axis_kernel<NN, CC, NBB, WW, PL, OC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  FC_FUSED_ALL(FC_FUSED_ATTR)
#undef FC_FUSED_ATTR
#endif
}

// Per-device one-time setup: the opt-in to > 48 KB of dynamic shared memory is a property of the function *in the
// current device's context*, and the SM count differs between devices; a process that drives several GPUs runs this
// once per device.
void init_once() {
#ifdef FC_CPU_EMUL
  return;
#else
  static std::mutex m;
  static int sms_of[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  std::lock_guard<std::mutex> lock(m);
  if (sms_of[dev] > 0) {
    g_num_sms = sms_of[dev];
    return;
  }
  cudaFuncSetAttribute(fc_pass_kernel<FC_R2C>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  cudaFuncSetAttribute(fc_pass_kernel<FC_C2C_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  cudaFuncSetAttribute(fc_pass_kernel<FC_C2C_INV>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  cudaFuncSetAttribute(fc_pass_kernel<FC_C2R>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
#define FC_FAST_ATTR(MM, NLL, NWW, OC)                                                                                   \
  cudaFuncSetAttribute(fc_fast_r2c_kernel<MM, NLL, NWW, OC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem); \
  cudaFuncSetAttribute(fc_fast_c2r_kernel<MM, NLL, NWW, OC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  FC_FAST_ALL(FC_FAST_ATTR)
#undef FC_FAST_ATTR
#define FC_FAST_C2C_ATTR(NN, NLL, NWW, OC) \
  cudaFuncSetAttribute(fc_fast_c2c_kernel<NN, NLL, NWW, OC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  FC_FAST_C2C_ALL(FC_FAST_C2C_ATTR)
#undef FC_FAST_C2C_ATTR
#define FC_PLANE_ATTR(NY, NZ)                                                                                  \
  cudaFuncSetAttribute(fc_plane_fwd_kernel<NY, NZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem); \
  cudaFuncSetAttribute(fc_plane_inv_kernel<NY, NZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  FC_PLANE_ALL(FC_PLANE_ATTR)
#undef FC_PLANE_ATTR
#define FC_PAIR_ROW_ATTR(MM, NLL, NWW, OC)                                                                               \
  cudaFuncSetAttribute(fc_pair_r2c_kernel<MM, NLL, NWW, OC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem); \
  cudaFuncSetAttribute(fc_pair_c2r_kernel<MM, NLL, NWW, OC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  FC_PAIR_ROW_ALL(FC_PAIR_ROW_ATTR)
#undef FC_PAIR_ROW_ATTR
#define FC_PAIR_FUSED_ATTR(NN, CC, NPP, WW, PL, OC, TP) \
  cudaFuncSetAttribute(fc_pair_fused_kernel<NN, CC, NPP, WW, PL, OC, TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  FC_PAIR_FUSED_ALL(FC_PAIR_FUSED_ATTR)
#undef FC_PAIR_FUSED_ATTR
#define FC_PAIR_ROW_YS_ATTR(MM, NLL, NWW, OC, YY)                                                                          \
  cudaFuncSetAttribute(fc_pair_r2c_kernel<MM, NLL, NWW, OC, YY>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem); \
  cudaFuncSetAttribute(fc_pair_c2r_kernel<MM, NLL, NWW, OC, YY>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  FC_PAIR_ROW_YS_ALL(FC_PAIR_ROW_YS_ATTR)
#undef FC_PAIR_ROW_YS_ATTR
#define FC_PAIR_FUSED64_ATTR(SS, CC, OC) \
  cudaFuncSetAttribute(fc_pair_fused64_kernel<SS, CC, OC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  FC_PAIR_FUSED64_ALL(FC_PAIR_FUSED64_ATTR)
#undef FC_PAIR_FUSED64_ATTR
  fused_set_attr();
  cudaFuncSetAttribute(fc_tc_gemm_kernel<1, 3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
  cudaFuncSetAttribute(fc_tc_gemm_kernel<2, 3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
  cudaFuncSetAttribute(fc_tc_gemm_kernel<2, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
#ifndef FC_CPU_EMUL
  cudaFuncSetAttribute(fc_contract_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  cudaFuncSetAttribute(fc_line_r2c_kernel<512, 8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  cudaFuncSetAttribute(fc_line_c2r_kernel<512, 8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
#endif
  cudaFuncSetAttribute(fc_tc_c2c_fwd_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  cudaFuncSetAttribute(fc_tc_c2c_inv_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  cudaFuncSetAttribute(fc_tc_c2c_fwd_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  cudaFuncSetAttribute(fc_tc_c2c_inv_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  cudaFuncSetAttribute(fc_stream_r2c_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, fc_stream::smem_bytes(2));
  cudaFuncSetAttribute(fc_stream_c2r_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, fc_stream::smem_bytes(2));
  cudaGetLastError();
  sms_of[dev] = sms;
  g_num_sms = sms;
#endif
}

int launch_pass(const fc_plan* pl, const fc_pass& p, const void* in, void* out, const float2* tw, const float* bias, cudaStream_t st) {
  fc_pass_args a;
  a.p = p;
  a.in = in;
  a.out = out;
  a.tw = tw;
  a.bias = bias;
  if (p.kind == FC_C2R) a.p.has_bias = bias ? 1 : 0;
  const size_t smem = (size_t)2 * p.T * p.pitch * sizeof(float2);
  if (smem > (size_t)kMaxSmem) return set_err(FC_EUNSUPPORTED, "pass tile does not fit shared memory");
  // Persistent CTAs: enough to fill every SM at the occupancy the tile allows, never more than there are tiles.
  int64_t per_sm = (int64_t)(220 * 1024) / (int64_t)(smem + 1024);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  int64_t grid = (int64_t)g_num_sms * per_sm;
  if (grid > p.n_tiles) grid = p.n_tiles;
  if (grid < 1) return FC_OK;
  // big tiles leave room for one CTA per SM only: give that CTA more warps to hide latency
  unsigned threads = (unsigned)pl->threads;
  if (pl->prob.threads == 0) {
    if (smem > 100 * 1024) threads = 1024;
  }
  dim3 g((unsigned)grid), b(threads);
  switch (p.kind) {
    case FC_R2C: {
      auto k = fc_pass_kernel<FC_R2C>;
      FC_LAUNCH(k, g, b, smem, st, a);
    } break;
    case FC_C2C_FWD: {
      auto k = fc_pass_kernel<FC_C2C_FWD>;
      FC_LAUNCH(k, g, b, smem, st, a);
    } break;
    case FC_C2C_INV: {
      auto k = fc_pass_kernel<FC_C2C_INV>;
      FC_LAUNCH(k, g, b, smem, st, a);
    } break;
    default: {
      auto k = fc_pass_kernel<FC_C2R>;
      FC_LAUNCH(k, g, b, smem, st, a);
    } break;
  }
  rec_mark();
  return check_cuda("axis pass launch");
}

int launch_contract(const float2* X, const float2* K, float2* Y, int64_t bins, int batch, int cin, int cout, int groups, cudaStream_t st) {
  fc_contract_args a;
  a.X = X;
  a.K = K;
  a.Y = Y;
  a.bins = bins;
  a.batch = batch;
  a.cin = cin;
  a.cout = cout;
  a.groups = groups;
  const int Og = cout / groups;
  static const char* ctile = fc_tune_str("CTILE");  // timing experiments: "tb,to,bx,by"
  if (ctile) {
    int tb = 8, to = 8, bx = 128, by = 1;
    std::sscanf(ctile, "%d,%d,%d,%d", &tb, &to, &bx, &by);
    a.btiles = (batch + tb - 1) / tb;
    a.otiles = (Og + to - 1) / to;
    dim3 b(bx, by), g((unsigned)((bins + bx - 1) / bx), (unsigned)(a.btiles * ((a.otiles + by - 1) / by)), (unsigned)groups);
    if (tb == 16 && to == 4) {
      auto k = fc_contract_kernel<16, 4>;
      FC_LAUNCH(k, g, b, 0, st, a);
    } else if (tb == 8 && to == 8) {
      auto k = fc_contract_kernel<8, 8>;
      FC_LAUNCH(k, g, b, 0, st, a);
    } else if (tb == 16 && to == 2) {
      auto k = fc_contract_kernel<16, 2>;
      FC_LAUNCH(k, g, b, 0, st, a);
    } else if (tb == 4 && to == 8) {
      auto k = fc_contract_kernel<4, 8>;
      FC_LAUNCH(k, g, b, 0, st, a);
    } else {
      auto k = fc_contract_kernel<4, 4>;
      FC_LAUNCH(k, g, b, 0, st, a);
    }
  } else if (batch >= 8 && Og >= 12 && cin / groups >= 48 && bins >= 4096 && !fc_tune_int("NO_CTILED", 0)) {
    // wide channel groups on long spectra: 16 x 16 tiles fed through shared memory (fc_contract_tiled_kernel). Measured against
    // the register tiles below (profiles/r2b_contraction_paths.txt): 64 channels, 33792 bins: 1278 -> 973 us; 96 channels: 2757 ->
    // 2101 us; 64 channels, 8320 bins: 346 -> 285 us; but 32 channels, 33024 bins: 186 -> 251 us and 64 channels, 257 bins
    // (batch 256): 143 -> 162 us, hence the bounds.
    a.btiles = (batch + FC_CT_TB - 1) / FC_CT_TB;
    a.otiles = (Og + FC_CT_TO - 1) / FC_CT_TO;
    dim3 b(256), g((unsigned)((bins + 31) / 32), (unsigned)(a.btiles * a.otiles), (unsigned)groups);
    const size_t smem = (size_t)FC_CT_KC * (FC_CT_TB + FC_CT_TO) * 32 * sizeof(float2);
    auto k = fc_contract_tiled_kernel;
    FC_LAUNCH(k, g, b, smem, st, a);
  } else if (batch >= 5 && Og >= 32) {
    // wide channels (BASELINE c4): 4 output tiles per CTA read the same signal spectrum (L1 hits for 3 of them)
    a.btiles = (batch + 7) / 8;
    a.otiles = (Og + 7) / 8;
    const int oy = 4;
    dim3 b(32, oy), g((unsigned)((bins + 31) / 32), (unsigned)(a.btiles * ((a.otiles + oy - 1) / oy)), (unsigned)groups);
    auto k = fc_contract_kernel<8, 8>;
    FC_LAUNCH(k, g, b, 0, st, a);
  } else if (batch >= 3 && batch <= 4 && Og >= 8) {
    // small batch, 8+ output channels per group (BASELINE c3, c5): the kernel spectrum is the dominant stream; a 4 x 8
    // tile reads it once per thread and two output tiles per CTA share the signal spectrum through L1
    a.btiles = 1;
    a.otiles = (Og + 7) / 8;
    const int oy = a.otiles >= 2 ? 2 : 1, bx = oy == 2 ? 32 : 64;  // (64-thread CTAs: c3 35.5 -> 34.8 us against 128)
    dim3 b(bx, oy), g((unsigned)((bins + bx - 1) / bx), (unsigned)((a.otiles + oy - 1) / oy), (unsigned)groups);
    auto k = fc_contract_kernel<4, 8>;
    FC_LAUNCH(k, g, b, 0, st, a);
  } else {
    const int threads = 128;
    const unsigned gx = (unsigned)((bins + threads - 1) / threads);
    if (batch >= 5 && batch <= 64 && Og >= 5) {
      // medium batch: 4 x 8 tiles (half the accumulators of 8 x 8, twice the resident warps; the kernel spectrum is
      // re-read per batch tile from L1 / L2): (32,8,128,128) k15 18.5 -> 16.4 us
      a.btiles = (batch + 3) / 4;
      a.otiles = (Og + 7) / 8;
      dim3 g((unsigned)((bins + 63) / 64), (unsigned)(a.btiles * a.otiles), (unsigned)groups), b(64);
      auto k = fc_contract_kernel<4, 8>;
      FC_LAUNCH(k, g, b, 0, st, a);
    } else if (batch >= 5 && Og >= 5) {
      a.btiles = (batch + 7) / 8;
      a.otiles = (Og + 7) / 8;
      dim3 g(gx, (unsigned)(a.btiles * a.otiles), (unsigned)groups), b(threads);
      auto k = fc_contract_kernel<8, 8>;
      FC_LAUNCH(k, g, b, 0, st, a);
    } else if (batch >= 3 || Og >= 3) {
      a.btiles = (batch + 3) / 4;
      a.otiles = (Og + 3) / 4;
      dim3 g(gx, (unsigned)(a.btiles * a.otiles), (unsigned)groups), b(threads);
      auto k = fc_contract_kernel<4, 4>;
      FC_LAUNCH(k, g, b, 0, st, a);
    } else {
      a.btiles = (batch + 1) / 2;
      a.otiles = (Og + 1) / 2;
      dim3 g(gx, (unsigned)(a.btiles * a.otiles), (unsigned)groups), b(threads);
      auto k = fc_contract_kernel<2, 2>;
      FC_LAUNCH(k, g, b, 0, st, a);
    }
  }
  rec_mark();
  return check_cuda("contraction launch");
}

// Variant of the transposing kernels for a line length: (lines per warp, warps, CTAs per SM). FFTCONV_B200_FAST
// ("nl,nw,occ") overrides it for timing experiments; unknown combinations fall back to the default.
struct fast_cfg {
  int nl, nw, occ;
};
fast_cfg fast_config(int M) {
  fast_cfg c = M <= 256 ? fast_cfg{2, 8, 4} : M == 512 ? fast_cfg{2, 8, 2} : fast_cfg{1, 16, 1};
  static const char* env = fc_tune_str("FAST");
  if (env) {
    fast_cfg e = c;
    std::sscanf(env, "%d,%d,%d", &e.nl, &e.nw, &e.occ);
    bool known = false;
#define FC_FAST_KNOWN(MM, NLL, NWW, OC) known = known || (M == MM && e.nl == NLL && e.nw == NWW && e.occ == OC);
    FC_FAST_ALL(FC_FAST_KNOWN)
#undef FC_FAST_KNOWN
    if (known) c = e;
  }
  return c;
}

int64_t fast_grid(const fc_pass& p, const fast_cfg& c, size_t smem) {
  int64_t per_sm = (int64_t)(224 * 1024) / (int64_t)(smem + 1024);
  if (per_sm > c.occ) per_sm = c.occ;
  if (per_sm < 1) per_sm = 1;
  int64_t grid = (int64_t)g_num_sms * per_sm;
  return grid > p.n_tiles ? p.n_tiles : grid;
}

#ifndef FC_CPU_EMUL
// ---- streaming K1 / K4 (fc_stream.cuh): tensor-map encoder and eligibility
typedef CUresult (*fc_tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                      CUtensorMapFloatOOBfill);
fc_tmap_encode_fn tmap_encoder() {
  static const fc_tmap_encode_fn fn = []() -> fc_tmap_encode_fn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return (fc_tmap_encode_fn)p;
  }();
  return fn;
}

// Tensor map of a transposed half spectrum [item][bin][row] (float2) seen as floats: box = 16 rows x 256 bins of one item,
// 128-byte swizzle (the shared-memory layout fc_stream_* read and write).
bool encode_spectrum_tmap(CUtensorMap* tm, const void* base, int64_t rows, int64_t bins, int64_t items, int64_t bin_stride, int64_t item_stride) {
  fc_tmap_encode_fn enc = tmap_encoder();
  if (!enc) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)(2 * rows), (cuuint64_t)bins, (cuuint64_t)items};
  const cuuint64_t strides[2] = {(cuuint64_t)bin_stride * 8, (cuuint64_t)item_stride * 8};
  const cuuint32_t box[3] = {32, 256, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

bool stream_r2c_ok(const fc_pass& p, const void* in, const void* out) {
  const fc_imap& im = p.imap;
  return p.M == 256 && p.seg_n == 1 && p.ystage == 0 && p.twiddle == 0 && p.o_c2 == 1 && p.o_q == 1 && im.up == 1 && im.sub == 1 &&
         im.pad >= 0 && (im.pad & 3) == 0 && im.L > 0 && (im.L & 3) == 0 && im.pad + im.L <= 2 * p.M && p.in_es == 1 && (p.in_rs & 3) == 0 &&
         (p.o_sA & 3) == 0 && aligned16(in) && p.out_rs == 1 && (p.out_es & 1) == 0 && (p.out_os & 1) == 0 && (p.R & 1) == 0 && aligned16(out) &&
         p.n_tiles < (1ll << 30) && p.n_outer < (1ll << 31);
}

bool stream_c2r_ok(const fc_pass& p, const void* in, const void* out) {
  const fc_omap& om = p.omap;
  return p.M == 256 && p.seg_n == 1 && p.ystage == 0 && p.twiddle == 0 && om.os == 1 && om.ob == 0 && om.og == 1 && !(om.Lout & 1) &&
         !(p.out_rs & 1) && !(p.out_os & 1) && p.row_og == 1 && p.in_rs == 1 && (p.in_es & 1) == 0 && (p.in_os & 1) == 0 && (p.R & 1) == 0 &&
         aligned16(in) && (reinterpret_cast<uintptr_t>(out) & 7) == 0 && p.n_tiles < (1ll << 30) && p.n_outer < (1ll << 31);
}
#endif

int launch_fast_r2c(const fc_pass& p, const void* in, void* out, const float2* tw, cudaStream_t st, bool allow_stream) {
#ifndef FC_CPU_EMUL
  if (allow_stream && stream_r2c_ok(p, in, out)) {
    fc_stream_r2c_args sa;
    if (encode_spectrum_tmap(&sa.tmap, out, p.R, p.M + 1, p.n_outer, p.out_es, p.out_os)) {
      sa.p = p;
      sa.x = (const float*)in;
      sa.out = (float2*)out;
      sa.tw = tw;
      sa.whole_tiles = (p.imap.pad == 0 && p.imap.L == 2 * p.M && p.in_rs == 2 * p.M && p.R % 16 == 0) ? 1 : 0;
      int64_t grid = (int64_t)g_num_sms * 3;
      if (grid > p.n_tiles) grid = p.n_tiles;
      if (grid < 1) return FC_OK;
      FC_LAUNCH(fc_stream_r2c_kernel<2>, dim3((unsigned)grid), dim3(256), (size_t)fc_stream::smem_bytes(2), st, sa);
      rec_mark();
      return check_cuda("streaming r2c launch");
    }
  }
#endif
  fc_fast_r2c_args a;
  a.p = p;
  a.x = (const float*)in;
  a.out = (float2*)out;
  a.tw = tw;
  const fast_cfg c = fast_config(p.M);
  const size_t smem = (size_t)p.T * (p.M + 1) * sizeof(float2);
  const int64_t grid = fast_grid(p, c, smem);
  if (grid < 1) return FC_OK;
  dim3 g((unsigned)grid), b(c.nw * 32);
  bool done = false;
#define FC_FAST_LAUNCH(MM, NLL, NWW, OC)                                  \
  if (!done && p.M == MM && c.nl == NLL && c.nw == NWW && c.occ == OC) { \
    auto k = fc_fast_r2c_kernel<MM, NLL, NWW, OC>;                       \
    FC_LAUNCH(k, g, b, smem, st, a);                                     \
    done = true;                                                         \
  }
  FC_FAST_ALL(FC_FAST_LAUNCH)
#undef FC_FAST_LAUNCH
  if (!done) return set_err(FC_EUNSUPPORTED, "no transposing R2C kernel for this line length");
  rec_mark();
  return check_cuda("fast r2c launch");
}

int launch_fast_c2r(const fc_pass& p, const void* in, void* out, const float2* tw, const float* bias, cudaStream_t st, bool allow_stream) {
#ifndef FC_CPU_EMUL
  if (allow_stream && stream_c2r_ok(p, in, out)) {
    fc_stream_c2r_args sa;
    if (encode_spectrum_tmap(&sa.tmap, in, p.R, p.M + 1, p.n_outer, p.in_es, p.in_os)) {
      sa.p = p;
      sa.p.has_bias = bias ? 1 : 0;
      sa.in = (const float2*)in;
      sa.out = (float*)out;
      sa.tw = tw;
      sa.bias = bias;
      int64_t grid = (int64_t)g_num_sms * 3;
      if (grid > p.n_tiles) grid = p.n_tiles;
      if (grid < 1) return FC_OK;
      FC_LAUNCH(fc_stream_c2r_kernel<2>, dim3((unsigned)grid), dim3(256), (size_t)fc_stream::smem_bytes(2), st, sa);
      rec_mark();
      return check_cuda("streaming c2r launch");
    }
  }
#endif
  fc_fast_c2r_args a;
  a.p = p;
  a.p.has_bias = bias ? 1 : 0;
  a.in = (const float2*)in;
  a.out = (float*)out;
  a.tw = tw;
  a.bias = bias;
  const fast_cfg c = fast_config(p.M);
  const size_t smem = (size_t)p.T * (p.M + 1) * sizeof(float2);
  const int64_t grid = fast_grid(p, c, smem);
  if (grid < 1) return FC_OK;
  dim3 g((unsigned)grid), b(c.nw * 32);
  bool done = false;
#define FC_FAST_LAUNCH(MM, NLL, NWW, OC)                                  \
  if (!done && p.M == MM && c.nl == NLL && c.nw == NWW && c.occ == OC) { \
    auto k = fc_fast_c2r_kernel<MM, NLL, NWW, OC>;                       \
    FC_LAUNCH(k, g, b, smem, st, a);                                     \
    done = true;                                                         \
  }
  FC_FAST_ALL(FC_FAST_LAUNCH)
#undef FC_FAST_LAUNCH
  if (!done) return set_err(FC_EUNSUPPORTED, "no transposing C2R kernel for this line length");
  rec_mark();
  return check_cuda("fast c2r launch");
}

int launch_column(const fc_pass& p, const void* in, void* out, const float2* tw, const float* bias, cudaStream_t st) {
  fc_col_args a;
  a.p = p;
  a.in = in;
  a.out = out;
  a.tw = tw;
  a.bias = bias;
  if (p.kind == FC_C2R) a.p.has_bias = bias ? 1 : 0;
  const int64_t cols = p.n_outer * p.R;
  if (cols < 1) return FC_OK;
  const int bt = cols >= (int64_t)g_num_sms * 256 ? 128 : 32;  // few columns: one warp per CTA spreads them over the SMs
  dim3 g((unsigned)((cols + bt - 1) / bt)), b(bt);
  if (p.kind == FC_R2C) {
    if (p.imap.mode == FC_PAD_CONSTANT && p.imap.up == 1 && p.imap.sub == 1) {
      auto k = fc_col_r2c_kernel<true>;
      FC_LAUNCH(k, g, b, 0, st, a);
    } else {
      auto k = fc_col_r2c_kernel<false>;
      FC_LAUNCH(k, g, b, 0, st, a);
    }
  } else {
    auto k = fc_col_c2r_kernel;
    FC_LAUNCH(k, g, b, 0, st, a);
  }
  rec_mark();
  return check_cuda("column pass launch");
}

int launch_plane(bool inverse, const fc_plane_desc& d, const void* in, void* out, const float2* tw, int tw_len, cudaStream_t st) {
  fc_plane_args a;
  std::memset(&a, 0, sizeof(a));
  a.in = (const float2*)in;
  a.out = (float2*)out;
  a.tw = tw;
  a.tw_len = tw_len;
  a.nkx = d.nkx;
  a.n_units = d.n_outer * d.nkx;
  a.in_os = d.in_os;
  a.out_os = d.out_os;
  a.imy = d.imy;
  a.imz = d.imz;
  a.conj_out = d.conj_out;
  a.scale = d.scale;
  a.omy = d.omy;
  a.omz = d.omz;
  if (a.n_units < 1) return FC_OK;
  // rows of the first pass live in the plane (pitch + 1); one exchange line (pitch + 2) per line of the second pass
  const int n1 = inverse ? d.nz : d.ny, n2 = inverse ? d.ny : d.nz;  // first-pass / second-pass line lengths
  const int rows = inverse ? d.ny : d.nz;
  const size_t smem = ((size_t)rows * (n1 + 1) + (size_t)FC_PLANE_WARPS * (256 / n2) * 2 * (n2 + 2)) * sizeof(float2);
  int64_t grid = (int64_t)g_num_sms * 3;
  if (grid > a.n_units) grid = a.n_units;
  dim3 g((unsigned)grid), b(FC_PLANE_WARPS * 32);
  bool done = false;
#define FC_PLANE_LAUNCH(NY, NZ)                     \
  if (!done && d.ny == NY && d.nz == NZ) {          \
    if (inverse) {                                  \
      auto k = fc_plane_inv_kernel<NY, NZ>;         \
      FC_LAUNCH(k, g, b, smem, st, a);              \
    } else {                                        \
      auto k = fc_plane_fwd_kernel<NY, NZ>;         \
      FC_LAUNCH(k, g, b, smem, st, a);              \
    }                                               \
    done = true;                                    \
  }
  FC_PLANE_ALL(FC_PLANE_LAUNCH)
#undef FC_PLANE_LAUNCH
  if (!done) return set_err(FC_EUNSUPPORTED, "no plane kernel for these extents");
  rec_mark();
  return check_cuda("plane pass launch");
}

// The real pass of a one-pass 1-d program on the warp engine (fc_line.cuh): 8 warps x 2 lines per CTA.
int launch_line(const fc_pass& p, const void* in, void* out, const float2* tw, const float* bias, cudaStream_t st) {
  fc_line_args a;
  a.p = p;
  a.p.has_bias = bias ? 1 : 0;
  a.in = in;
  a.out = out;
  a.tw = tw;
  a.bias = bias;
  if (p.n_outer < 1) return FC_OK;
  const size_t smem = (size_t)8 * 2 * p.M * sizeof(float2);
  const int occ = p.M == 256 ? 4 : 3;
  int64_t grid = (p.n_outer + 15) / 16;
  if (grid > (int64_t)g_num_sms * occ) grid = (int64_t)g_num_sms * occ;
  dim3 g((unsigned)grid), b(256);
  if (p.kind == FC_R2C && p.M == 256) { auto k = fc_line_r2c_kernel<256, 8, 4>; FC_LAUNCH(k, g, b, smem, st, a); }
  else if (p.kind == FC_R2C && p.M == 512) { auto k = fc_line_r2c_kernel<512, 8, 3>; FC_LAUNCH(k, g, b, smem, st, a); }
  else if (p.kind == FC_C2R && p.M == 256) { auto k = fc_line_c2r_kernel<256, 8, 4>; FC_LAUNCH(k, g, b, smem, st, a); }
  else if (p.kind == FC_C2R && p.M == 512) { auto k = fc_line_c2r_kernel<512, 8, 3>; FC_LAUNCH(k, g, b, smem, st, a); }
  else return set_err(FC_EUNSUPPORTED, "no line kernel for this pass");
  rec_mark();
  return check_cuda("line pass launch");
}

int launch_fast_c2c(const fc_pass& p, const void* in, void* out, const float2* tw, cudaStream_t st) {
  fc_fast_c2c_args a;
  a.p = p;
  a.in = (const float2*)in;
  a.out = (float2*)out;
  a.tw = tw;
  const int64_t n_lines = p.n_outer * p.R;
  if (n_lines < 1) return FC_OK;
  bool done = false;
#define FC_FAST_C2C_LAUNCH(NN, NLL, NWW, OC)                                                 \
  if (!done && p.N == NN) {                                                                  \
    const int gpw = NN >= 256 ? 1 : 256 / NN; /* line groups per warp */                     \
    const size_t smem = (size_t)NLL * NWW * gpw * (NN >= 256 ? NN : NN + 2) * sizeof(float2); \
    const int64_t ctas = (n_lines + NLL * NWW * gpw - 1) / (NLL * NWW * gpw);                \
    int64_t grid = (int64_t)g_num_sms * OC;                                                  \
    if (grid > ctas) grid = ctas;                                                            \
    dim3 g((unsigned)grid), b(NWW * 32);                                                     \
    auto k = fc_fast_c2c_kernel<NN, NLL, NWW, OC>;                                           \
    FC_LAUNCH(k, g, b, smem, st, a);                                                         \
    done = true;                                                                             \
  }
  FC_FAST_C2C_ALL(FC_FAST_C2C_LAUNCH)
#undef FC_FAST_C2C_LAUNCH
  if (!done) return set_err(FC_EUNSUPPORTED, "no contiguous complex pass kernel for this line length");
  rec_mark();
  return check_cuda("fast c2c launch");
}

int launch_fused(const fc_plan* pl, const fc_fused_desc& f, const void* in, const float2* kspec, void* out, const float2* tw, cudaStream_t st,
                 float* d_y, const float* d_bias) {
  const fc_problem& P = pl->prob;
  fc_fused_args a;
  a.xin = (const float2*)in;
  a.kspec = kspec;
  a.yout = (float2*)out;
  a.tw = tw;
  a.tw_len = pl->tw_len;
  a.B = P.batch;
  a.Cin = P.cin;
  a.Cout = P.cout;
  a.G = P.groups;
  a.Ig = P.cin / P.groups;
  a.Og = P.cout / P.groups;
  a.n_in = f.n_in;
  a.n_out = f.n_out;
  a.n_seg = f.n_seg > 1 ? f.n_seg : 1;
  a.seg_V = f.n_seg > 1 ? f.seg_V : f.N;
  a.seg_off = f.n_seg > 1 ? f.seg_off : 0;
  a.n_items = P.batch * a.n_seg;
  a.nbs = (a.n_items + f.nb - 1) / f.nb;
  a.R = f.R;
  a.Rk = f.Rk > 0 ? f.Rk : f.R;
  a.n_units = (int64_t)P.groups * f.R * a.nbs;
  a.imap = f.imap;
  a.omap = f.omap;
  a.fill_y = nullptr;
  a.fill_bias = nullptr;
  a.fill_og = a.fill_ob = a.fill_Lrow = a.fill_Lcol = a.fill_cout = a.fill_rpu = 0;
  a.fill_img_stride = a.fill_row_stride = a.fill_total = 0;
  if (f.fill_rows && d_y && a.n_units > 0) {
    const fc_pass& c2r = pl->prog.back().pass;  // the last kernel: geometry of the output rows
    a.fill_y = d_y;
    a.fill_bias = d_bias;
    a.fill_og = c2r.row_og;
    a.fill_ob = c2r.row_ob;
    a.fill_Lrow = c2r.row_Lout;
    a.fill_Lcol = c2r.omap.Lout;
    a.fill_cout = P.cout;
    a.fill_img_stride = c2r.out_os;
    a.fill_row_stride = c2r.out_rs;
    a.fill_total = (int64_t)P.batch * P.cout * c2r.row_Lout;
    a.fill_rpu = (int32_t)((a.fill_total + a.n_units - 1) / a.n_units);
  }
  const size_t smem = (size_t)f.nb * f.ci * f.N * sizeof(float2);
  {  // distance (in units) to the CTA of the next wave on the same SM: what this CTA prefetches into L2
    int64_t per_sm = (int64_t)(228 * 1024) / (int64_t)(smem + 1024 + 256);
    const int64_t reg_lim = f.occ;  // launch bounds
    if (per_sm > reg_lim) per_sm = reg_lim;
    if (per_sm < 1) per_sm = 1;
    a.prefetch_dist = (int)(g_num_sms * per_sm);
  }
  int64_t grid = a.n_units;
  int64_t cap = (int64_t)g_num_sms * 16;
  {  // A/B timing knob: "pgrid=K" in FFTCONV_B200_TUNE caps the grid at K CTAs per SM (persistent CTAs looping over units)
    static const int pgrid = []() {
      const char* t = fc_tune_str("TUNE");
      const char* q = t ? std::strstr(t, "pgrid=") : nullptr;
      return q ? std::atoi(q + 6) : 0;
    }();
    if (pgrid > 0) cap = (int64_t)g_num_sms * pgrid;
  }
  if (grid > cap) grid = cap;
  dim3 g((unsigned)grid), b((unsigned)f.warps * 32);
  bool ok = false;
#define FC_FUSED_CASE(NN, CC, NBB, WW, PL, OC)                                   \
  if (!ok && f.N == NN && f.ci == CC && f.nb == NBB && f.warps == WW && (f.plain != 0) == PL && f.occ == OC) { \
    auto k = fc_fused_axis_kernel<NN, CC, NBB, WW, PL, OC>;                   \
    FC_LAUNCH(k, g, b, smem, st, a);                                     \
    ok = true;                                                           \
  }
  FC_FUSED_ALL(FC_FUSED_CASE)
#undef FC_FUSED_CASE
  if (!ok) return set_err(FC_EUNSUPPORTED, "no fused kernel instantiation for this shape");
  rec_mark();
  return check_cuda("fused axis launch");
}

// ---- pair pipeline (fc_pair.cuh)
// Variant of K1p / K4p for a line length: (pair lines per warp group, CTAs per SM); 8 warps per CTA. The plan's tiling
// (fc_pair_tile_lines) is recomputed here for the variant actually launched.
struct pair_row_cfg {
  int nlp, occ;
};
pair_row_cfg pair_row_config(int M) {
  pair_row_cfg c = M <= 256 ? pair_row_cfg{1, 4} : M == 512 ? pair_row_cfg{1, 2} : pair_row_cfg{1, 1};
  static const char* env = fc_tune_str("PAIRROW");  // tuning builds: "nlp,occ"
  if (env) {
    pair_row_cfg e = c;
    std::sscanf(env, "%d,%d", &e.nlp, &e.occ);
    bool known = false;
#define FC_PAIR_KNOWN(MM, NLL, NWW, OC) known = known || (M == MM && e.nlp == NLL && e.occ == OC);
    FC_PAIR_ROW_ALL(FC_PAIR_KNOWN)
#undef FC_PAIR_KNOWN
    if (known) c = e;
  }
  return c;
}
void pair_retile(fc_pass& p, int nlp) {
  const int T = nlp * 8 * (p.M >= 256 ? 1 : 256 / p.M);
  p.T = T;
  p.tiles_per_outer = (p.R + T - 1) / T * p.seg_n;
  p.n_tiles = p.tiles_per_outer * p.n_outer;
}

int launch_pair_r2c(const fc_plan* pl, const fc_pass& p, const void* in, void* out, const float2* tw, cudaStream_t st) {
  fc_pair_r2c_args a;
  a.p = p;
  a.x = (const float*)in;
  a.out = (fc_c2*)out;
  a.tw = tw;
  a.B = pl->prob.batch;
  a.C = pl->prob.cin;
  {
    static const int abl1 = fc_tune_int("ABL1", 0);
    a.abl = abl1;
  }
  if (p.ystage > 0) {  // y-stage variant: the plan fixed the 16-line tiling
    if (p.n_tiles < 1) return FC_OK;
    const size_t smem = (size_t)16 * (p.M + (p.ystage == 8 ? 4 : 2)) * sizeof(fc_c2);  // line pitch of the y-stage tiles (fc_pair.cuh)
    static const int ys_occ = fc_tune_int("YSOCC", 0);
    bool done = false;
#define FC_PAIR_YS_LAUNCH(MM, NLL, NWW, OC, YY)                                  \
  if (!done && p.M == MM && p.ystage == YY) {                                    \
    int64_t per_sm = (int64_t)(224 * 1024) / (int64_t)(smem + 1024);            \
    if (per_sm > OC) per_sm = OC;                                                \
    if (ys_occ > 0 && per_sm > ys_occ) per_sm = ys_occ;                          \
    if (per_sm < 1) per_sm = 1;                                                  \
    int64_t grid = (int64_t)g_num_sms * per_sm;                                  \
    if (grid > p.n_tiles) grid = p.n_tiles;                                      \
    dim3 g((unsigned)grid), b(NWW * 32);                                         \
    auto k = fc_pair_r2c_kernel<MM, NLL, NWW, OC, YY>;                            \
    FC_LAUNCH(k, g, b, smem, st, a);                                             \
    done = true;                                                                 \
  }
    FC_PAIR_ROW_YS_ALL(FC_PAIR_YS_LAUNCH)
#undef FC_PAIR_YS_LAUNCH
    if (!done) return set_err(FC_EUNSUPPORTED, "no y-stage pair row kernel for this line length");
    rec_mark();
    return check_cuda("pair row (y stage) launch");
  }
  const pair_row_cfg c = pair_row_config(p.M);
  pair_retile(a.p, c.nlp);
  if (a.p.n_tiles < 1) return FC_OK;
  const size_t smem = (size_t)a.p.T * (p.M + 1) * sizeof(fc_c2);
  bool done = false;
#define FC_PAIR_LAUNCH(MM, NLL, NWW, OC)                                         \
  if (!done && p.M == MM && c.nlp == NLL && c.occ == OC) {                       \
    int64_t per_sm = (int64_t)(224 * 1024) / (int64_t)(smem + 1024);            \
    if (per_sm > OC) per_sm = OC;                                                \
    if (per_sm < 1) per_sm = 1;                                                  \
    int64_t grid = (int64_t)g_num_sms * per_sm;                                  \
    if (grid > a.p.n_tiles) grid = a.p.n_tiles;                                  \
    dim3 g((unsigned)grid), b(NWW * 32);                                         \
    auto k = fc_pair_r2c_kernel<MM, NLL, NWW, OC>;                               \
    FC_LAUNCH(k, g, b, smem, st, a);                                             \
    done = true;                                                                 \
  }
  FC_PAIR_ROW_ALL(FC_PAIR_LAUNCH)
#undef FC_PAIR_LAUNCH
  if (!done) return set_err(FC_EUNSUPPORTED, "no pair R2C kernel for this line length");
  rec_mark();
  return check_cuda("pair r2c launch");
}

int launch_pair_c2r(const fc_plan* pl, const fc_pass& p, const void* in, void* out, const float2* tw, const float* bias, cudaStream_t st) {
  fc_pair_c2r_args a;
  a.p = p;
  a.p.has_bias = bias ? 1 : 0;
  a.in = (const fc_c2*)in;
  a.out = (float*)out;
  a.tw = tw;
  a.bias = bias;
  a.B = pl->prob.batch;
  a.C = pl->prob.cout;
  if (p.ystage > 0) {  // y-stage variant: the plan fixed the 16-line tiling
    if (p.n_tiles < 1) return FC_OK;
    const size_t smem = (size_t)16 * (p.M + (p.ystage == 8 ? 4 : 2)) * sizeof(fc_c2);  // line pitch of the y-stage tiles (fc_pair.cuh)
    static const int ys_occ = fc_tune_int("YSOCC", 0);
    bool done = false;
#define FC_PAIR_YS_LAUNCH(MM, NLL, NWW, OC, YY)                                  \
  if (!done && p.M == MM && p.ystage == YY) {                                    \
    int64_t per_sm = (int64_t)(224 * 1024) / (int64_t)(smem + 1024);            \
    if (per_sm > OC) per_sm = OC;                                                \
    if (ys_occ > 0 && per_sm > ys_occ) per_sm = ys_occ;                          \
    if (per_sm < 1) per_sm = 1;                                                  \
    int64_t grid = (int64_t)g_num_sms * per_sm;                                  \
    if (grid > p.n_tiles) grid = p.n_tiles;                                      \
    dim3 g((unsigned)grid), b(NWW * 32);                                         \
    auto k = fc_pair_c2r_kernel<MM, NLL, NWW, OC, YY>;                            \
    FC_LAUNCH(k, g, b, smem, st, a);                                             \
    done = true;                                                                 \
  }
    FC_PAIR_ROW_YS_ALL(FC_PAIR_YS_LAUNCH)
#undef FC_PAIR_YS_LAUNCH
    if (!done) return set_err(FC_EUNSUPPORTED, "no y-stage pair row kernel for this line length");
    rec_mark();
    return check_cuda("pair row (y stage) launch");
  }
  const pair_row_cfg c = pair_row_config(p.M);
  pair_retile(a.p, c.nlp);
  if (a.p.n_tiles < 1) return FC_OK;
  const size_t smem = (size_t)a.p.T * (p.M + 1) * sizeof(fc_c2);
  bool done = false;
#define FC_PAIR_LAUNCH(MM, NLL, NWW, OC)                                         \
  if (!done && p.M == MM && c.nlp == NLL && c.occ == OC) {                       \
    int64_t per_sm = (int64_t)(224 * 1024) / (int64_t)(smem + 1024);            \
    if (per_sm > OC) per_sm = OC;                                                \
    if (per_sm < 1) per_sm = 1;                                                  \
    int64_t grid = (int64_t)g_num_sms * per_sm;                                  \
    if (grid > a.p.n_tiles) grid = a.p.n_tiles;                                  \
    dim3 g((unsigned)grid), b(NWW * 32);                                         \
    auto k = fc_pair_c2r_kernel<MM, NLL, NWW, OC>;                               \
    FC_LAUNCH(k, g, b, smem, st, a);                                             \
    done = true;                                                                 \
  }
  FC_PAIR_ROW_ALL(FC_PAIR_LAUNCH)
#undef FC_PAIR_LAUNCH
  if (!done) return set_err(FC_EUNSUPPORTED, "no pair C2R kernel for this line length");
  rec_mark();
  return check_cuda("pair c2r launch");
}

int launch_pair_fused(const fc_plan* pl, const fc_fused_desc& f, const void* in, const float2* kspec, void* out, const float2* tw, cudaStream_t st) {
  const fc_problem& P = pl->prob;
  fc_pair_fused_args a;
  a.xin = (const fc_c2*)in;
  a.kspec = kspec;
  a.yout = (fc_c2*)out;
  a.tw = tw;
  a.tw_len = pl->tw_len;
  a.BP = (P.batch + 1) / 2;
  a.Cin = P.cin;
  a.Cout = P.cout;
  a.G = P.groups;
  a.n_in = f.n_in;
  a.n_out = f.n_out;
  a.n_seg = f.n_seg > 1 ? f.n_seg : 1;
  a.seg_V = f.n_seg > 1 ? f.seg_V : f.N;
  a.seg_off = f.n_seg > 1 ? f.seg_off : 0;
  a.n_items = a.BP * a.n_seg;
  a.nbs = (a.n_items + f.nb - 1) / f.nb;
  a.R = f.R;
  a.Rk = f.Rk > 0 ? f.Rk : f.R;
  a.n_units = (int64_t)P.groups * f.R * a.nbs;
  a.imap = f.imap;
  a.omap = f.omap;
  {
    static const int kpf = fc_tune_int("KPF", 0);
    a.k_pf = kpf;
    static const int abl = fc_tune_int("ABL", 0);
    a.abl = abl;
    static const int ksh = fc_tune_int("KSHARE", 0);
    a.k_share = ksh;
    static const int dns = fc_tune_int("DESYNC", 0);
    a.desync_ns = dns;
    a.desync_mod = g_num_sms;
  }
  const size_t smem = (size_t)f.nb * f.ci * f.N * sizeof(fc_c2);
  {  // distance (in units) to the CTA of the next wave on the same SM: what this CTA prefetches into L2
    int64_t per_sm = (int64_t)(228 * 1024) / (int64_t)(smem + 1024 + 256);
    if (per_sm > f.occ) per_sm = f.occ;
    if (per_sm < 1) per_sm = 1;
    a.prefetch_dist = (int)(g_num_sms * per_sm);
    a.desync_grp = (int)per_sm;
  }
  int64_t grid = a.n_units;
  const int64_t cap = (int64_t)g_num_sms * 16;
  if (grid > cap) grid = cap;
  if (grid < 1) return FC_OK;
  dim3 g((unsigned)grid), b((unsigned)f.warps * 32);
  bool ok = false;
#define FC_PAIR_FUSED_CASE(NN, CC, NPP, WW, PL, OC, TP)                                                                  \
  if (!ok && f.N == NN && f.ci == CC && f.nb == NPP && f.warps == WW && (f.plain != 0) == PL && f.occ == OC) { \
    auto k = fc_pair_fused_kernel<NN, CC, NPP, WW, PL, OC, TP>;                                                          \
    FC_LAUNCH(k, g, b, smem, st, a);                                                                                      \
    ok = true;                                                                                                            \
  }
  FC_PAIR_FUSED_ALL(FC_PAIR_FUSED_CASE)
#undef FC_PAIR_FUSED_CASE
  if (!ok) return set_err(FC_EUNSUPPORTED, "no pair fused kernel instantiation for this shape");
  rec_mark();
  return check_cuda("pair fused launch");
}

int launch_pair_fused64(const fc_plan* pl, const fc_fused_desc& f, const void* in, const float2* kspec, void* out, const float2* tw, cudaStream_t st) {
  const fc_problem& P = pl->prob;
  fc_pair_fused64_args a;
  a.xin = (const fc_c2*)in;
  a.kspec = kspec;
  a.yout = (fc_c2*)out;
  a.tw = tw;
  a.tw_len = pl->tw_len;
  a.BP = (P.batch + 1) / 2;
  a.Cin = P.cin;
  a.Cout = P.cout;
  a.G = P.groups;
  a.YS = f.ystage;
  const int S = f.ystage_S, npi = 256 / S;
  a.nbs = (a.BP + npi - 1) / npi;
  a.R = f.R;
  a.Rk = f.Rk > 0 ? f.Rk : f.R;
  a.n_units = (int64_t)P.groups * f.R * f.ystage * a.nbs;
  const size_t smem = (size_t)256 * f.ci * sizeof(fc_c2) + (size_t)f.ci * f.ci * S * 8 + 16;
  static const int occ_t = fc_tune_int("KB64OCC", 0);
  const int occ = occ_t > 0 ? occ_t : (S == 64 ? 3 : 2);
  int64_t grid = a.n_units;
  const int64_t cap = (int64_t)g_num_sms * 32;
  if (grid > cap) grid = cap;
  if (grid < 1) return FC_OK;
  dim3 g((unsigned)grid), b(256);
  bool ok = false;
#define FC_PAIR_FUSED64_CASE(SS, CC, OC)                  \
  if (!ok && S == SS && f.ci == CC && occ == OC) {        \
    auto k = fc_pair_fused64_kernel<SS, CC, OC>;          \
    FC_LAUNCH(k, g, b, smem, st, a);                      \
    ok = true;                                            \
  }
  FC_PAIR_FUSED64_ALL(FC_PAIR_FUSED64_CASE)
#undef FC_PAIR_FUSED64_CASE
  if (!ok) return set_err(FC_EUNSUPPORTED, "no 64-point fused kernel instantiation for this shape");
  rec_mark();
  return check_cuda("pair fused64 launch");
}

// ---- tensor-core contraction (fc_tc.cuh)
int tc_padded_batch(int batch) { return batch <= 8 ? 8 : (batch + 7) / 8 * 8; }
bool tc_supported(int batch, int cin, int cout, int groups) {
  const int I = cin / groups, O = cout / groups;
  return I >= 32 && (2 * I) % 32 == 0 && O % 64 == 0 && tc_padded_batch(batch) <= FC_TC_MAX_BATCH;
}

int launch_tc_relayout(int mode, const void* in, void* out, int64_t bins, int batch, int cin, int cout, int groups, cudaStream_t st) {
  fc_tc_relayout_args a;
  a.in = (const float2*)in;
  a.out = (float*)out;
  a.bins = bins;
  a.I = cin / groups;
  a.C = cin;
  a.O = cout;
  a.B = batch;
  a.Bp = tc_padded_batch(batch);
  a.Og = cout / groups;
  a.mode = mode;
  a.rows = mode == 0 ? cout * a.I : (mode == 1 ? batch * cin : cout * a.Bp);
  dim3 g((unsigned)((bins + 31) / 32), (unsigned)((a.rows + 31) / 32)), b(256);
  auto k = fc_tc_relayout_kernel;
  FC_LAUNCH(k, g, b, 0, st, a);
  rec_mark();
  return check_cuda("tc relayout launch");
}

int launch_tc_gemm(const float* A, const float* Bt, float* D, int64_t bins, int batch, int cin, int cout, int groups, cudaStream_t st) {
#ifdef FC_CPU_EMUL
  (void)A; (void)Bt; (void)D; (void)bins; (void)batch; (void)cin; (void)cout; (void)groups; (void)st;
  return set_err(FC_EUNSUPPORTED, "tensor-core contraction is not available in the host emulation");
#else
  fc_tc_args a;
  a.A = A;
  a.Bt = Bt;
  a.D = D;
  a.n_items = bins * groups;
  a.O = cout / groups;
  a.I = cin / groups;
  a.B = tc_padded_batch(batch);
  const int N = 2 * a.B;
  // two 128-row tiles per pass (the Bt chunk is fetched once for both) while three stages and two accumulator sets fit:
  // 3 * (2*MT*16 KB + 2*N*128 B) <= 223 KB and 2*MT*N <= 512 TMEM columns -> N <= 32 for MT = 2, N <= 160 for MT = 1
  // ... and with N >= 96 two tiles per pass on two stages and one accumulator set (fc_tc.cuh, variants)
  static const int wide2 = fc_tune_int("TC_WIDE2", 1);
  a.tile_rows = a.O % 128 == 0 ? 128 : 64;
  const bool two = a.O % 256 == 0 && (N <= 32 || (N >= 96 && wide2));
  const int MT = two ? 2 : 1, stages = (two && N > 32) ? 2 : 3;
  if (N > 2 * FC_TC_MAX_BATCH) return set_err(FC_EUNSUPPORTED, "tensor-core contraction: batch chunk too wide");
  const size_t stage = (size_t)2 * MT * 128 * 128 + 2 * (size_t)N * 128;
  const size_t smem = stages * stage + 1024;
  int64_t grid = a.n_items < g_num_sms ? a.n_items : g_num_sms;
  dim3 g((unsigned)grid), b(FC_TC_THREADS);
  if (MT == 2 && stages == 3) {
    auto k = fc_tc_gemm_kernel<2, 3, 2>;
    k<<<g, b, smem, st>>>(a);
  } else if (MT == 2) {
    auto k = fc_tc_gemm_kernel<2, 2, 1>;
    k<<<g, b, smem, st>>>(a);
  } else {
    auto k = fc_tc_gemm_kernel<1, 3, 2>;
    k<<<g, b, smem, st>>>(a);
  }
  rec_mark();
  return check_cuda("tc gemm launch");
#endif
}

// Last forward / first inverse pass writing / reading the GEMM operands directly (fc_tc_c2c_*_kernel).
int launch_tc_c2c(bool inverse, const fc_plan* pl, const fc_pass& p, const void* lines, void* tc_buf, const float2* tw, cudaStream_t st) {
#ifdef FC_CPU_EMUL
  (void)inverse; (void)pl; (void)p; (void)lines; (void)tc_buf; (void)tw; (void)st;
  return set_err(FC_EUNSUPPORTED, "tensor-core contraction is not available in the host emulation");
#else
  const fc_problem& P = pl->prob;
  fc_tc_c2c_args a;
  a.tw = tw;
  a.tw_len = p.tw_len;
  a.B = P.batch;
  a.Bp = tc_padded_batch(P.batch);
  a.G = P.groups;
  a.I = P.cin / P.groups;
  a.R = (int)p.R;
  if (!inverse) {
    a.in = (const float2*)lines;
    a.out = (float2*)tc_buf;
    a.C = P.cin;
    a.os = p.in_os;
    a.rs = p.in_rs;
    a.n_tiles = (int64_t)P.batch * p.R * (P.cin / 32);
  } else {
    a.in = (const float2*)tc_buf;
    a.out = (float2*)const_cast<void*>(lines);
    a.C = P.cout;
    a.os = p.out_os;
    a.rs = p.out_rs;
    a.n_tiles = (int64_t)((P.batch + 31) / 32) * P.cout * p.R;
  }
  const size_t smem = (size_t)32 * (p.N + 1) * sizeof(float2);
  int64_t grid = (int64_t)g_num_sms * (p.N == 256 ? 3 : 1);
  if (grid > a.n_tiles) grid = a.n_tiles;
  if (grid < 1) return FC_OK;
  dim3 g((unsigned)grid), b(256);
  if (p.N == 256) {
    if (inverse) { auto k = fc_tc_c2c_inv_kernel<256>; FC_LAUNCH(k, g, b, smem, st, a); }
    else { auto k = fc_tc_c2c_fwd_kernel<256>; FC_LAUNCH(k, g, b, smem, st, a); }
  } else if (p.N == 512) {
    if (inverse) { auto k = fc_tc_c2c_inv_kernel<512>; FC_LAUNCH(k, g, b, smem, st, a); }
    else { auto k = fc_tc_c2c_fwd_kernel<512>; FC_LAUNCH(k, g, b, smem, st, a); }
  } else {
    return set_err(FC_EUNSUPPORTED, "no fused transform + operand layout kernel for this line length");
  }
  rec_mark();
  return check_cuda("tc c2c launch");
#endif
}

// Resolve a buffer id of a step to a pointer.
struct Bufs {
  const void* user_in;
  void* spec;
  void* sA;
  void* sB;
  void* user_out;
};
void* buf_ptr(const Bufs& b, int id) {
  switch (id) {
    case FC_BUF_USER_IN: return const_cast<void*>(b.user_in);
    case FC_BUF_SPEC: return b.spec;
    case FC_BUF_SA: return b.sA;
    case FC_BUF_SB: return b.sB;
    default: return b.user_out;
  }
}

int run_steps(const fc_plan* pl, const std::vector<fc_step>& steps, const Bufs& b, const float2* tw, const float* bias, cudaStream_t st) {
  for (const fc_step& s : steps) {
    int rc = launch_pass(pl, s.pass, buf_ptr(b, s.src), buf_ptr(b, s.dst), tw, bias, st);
    if (rc) return rc;
  }
  return FC_OK;
}

}  // namespace

extern "C" {

const char* fc_last_error(void) { return g_err.c_str(); }
const char* fc_version(void) { return "fftconv_b200 0.1 (sm_100a)"; }

int fc_plan_create(fc_plan** out, const fc_problem* problem) {
  if (!out || !problem) return set_err(FC_ENULL, "fc_plan_create: NULL argument");
  *out = nullptr;
  fc_plan* pl = new fc_plan();
  std::string msg;
  int rc = fc_plan_build(pl, problem, &msg);
  if (rc) {
    delete pl;
    return set_err(rc, msg);
  }
  *out = pl;
  return FC_OK;
}

void fc_plan_destroy(fc_plan* plan) { delete plan; }

int fc_plan_get_info(const fc_plan* plan, fc_plan_info* info) {
  if (!plan || !info) return set_err(FC_ENULL, "fc_plan_get_info: NULL argument");
  *info = plan->info;
  return FC_OK;
}

int fc_plan_describe(const fc_plan* plan, char* buf, size_t buflen) {
  if (!plan || !buf || !buflen) return set_err(FC_ENULL, "fc_plan_describe: NULL argument");
  std::string s = fc_plan_to_string(plan);
  size_t n = s.size() < buflen - 1 ? s.size() : buflen - 1;
  std::memcpy(buf, s.data(), n);
  buf[n] = 0;
  return (int)n;
}

int fc_plan_init_const(const fc_plan* plan, void* d_const, void* stream) {
  if (!plan || !d_const) return set_err(FC_ENULL, "fc_plan_init_const: NULL argument");
  init_once();
  dim3 g((unsigned)((plan->tw_len + 255) / 256)), b(256);
  if (g.x > 64) g.x = 64;
  float2* tw = reinterpret_cast<float2*>(d_const);
  const int len = plan->tw_len;
  const int len2 = plan->structure == FC_S_1D_SPLIT ? plan->N2 : 0;
  const double big = plan->structure == FC_S_1D_SPLIT ? (double)plan->N1 * (double)plan->N2 : 1.0;
#ifdef FC_CPU_EMUL
  fc_emul_launch(g, b, 0, [=]() { fc_twiddle_kernel(tw, len, len2, big); });
#else
  fc_twiddle_kernel<<<g, b, 0, (cudaStream_t)stream>>>(tw, len, len2, big);
#endif
  return check_cuda("twiddle table launch");
}

int fc_signal_spectrum(const fc_plan* plan, const void* d_const, const float* d_x, float* d_xspec, void* d_ws, void* stream) {
  if (!plan || !d_const || !d_x || !d_xspec || !d_ws) return set_err(FC_ENULL, "fc_signal_spectrum: NULL argument");
  if (plan->info.segments > 1) return set_err(FC_EUNSUPPORTED, kSegMsg);
  init_once();
  Bufs b{d_x, d_xspec, (char*)d_ws + plan->off_sA, (char*)d_ws + plan->off_sB, nullptr};
  return run_steps(plan, plan->sig_fwd, b, (const float2*)d_const, nullptr, (cudaStream_t)stream);
}

int fc_kernel_spectrum(const fc_plan* plan, const void* d_const, const float* d_w, float* d_kspec, void* d_ws, void* stream) {
  if (!plan || !d_const || !d_w || !d_kspec || !d_ws) return set_err(FC_ENULL, "fc_kernel_spectrum: NULL argument");
  init_once();
  if (plan->use_tc) {
    // pass-order spectrum into a temporary, then once into the bin-outermost planar layout of the tensor-core GEMM
    const int64_t tmp_off = (plan->info.kspec_workspace_bytes - ((plan->info.kspec_bytes + 255) / 256 * 256)) / 256 * 256;
    float* tmp = (float*)((char*)d_ws + tmp_off);
    Bufs b{d_w, tmp, (char*)d_ws + plan->off_sA, (char*)d_ws + plan->off_sB, nullptr};
    int rc = run_steps(plan, plan->ker_fwd, b, (const float2*)d_const, nullptr, (cudaStream_t)stream);
    if (rc) return rc;
    const fc_contract_desc& c = plan->contract;
    return launch_tc_relayout(0, tmp, d_kspec, c.bins, 1, c.cin, c.cout, c.groups, (cudaStream_t)stream);
  }
  Bufs b{d_w, d_kspec, (char*)d_ws + plan->off_sA, (char*)d_ws + plan->off_sB, nullptr};
  return run_steps(plan, plan->ker_fwd, b, (const float2*)d_const, nullptr, (cudaStream_t)stream);
}

int fc_contract(const fc_plan* plan, const float* d_xspec, const float* d_kspec, float* d_yspec, void* stream) {
  if (!plan || !d_xspec || !d_kspec || !d_yspec) return set_err(FC_ENULL, "fc_contract: NULL argument");
  if (plan->info.segments > 1) return set_err(FC_EUNSUPPORTED, kSegMsg);
  if (plan->info.fused)
    return set_err(FC_EUNSUPPORTED,
                   "fc_contract: this plan keeps the kernel spectrum in the bin-major layout of the fused axis kernel; create the plan with "
                   "FC_FLAG_NO_FUSED_MID for the stage calls, or use fc_conv");
  init_once();
  if (plan->use_tc) return set_err(FC_EUNSUPPORTED, "fc_contract: this plan keeps the kernel spectrum in the tensor-core layout; use fc_conv");
  const fc_contract_desc& c = plan->contract;
  return launch_contract((const float2*)d_xspec, (const float2*)d_kspec, (float2*)d_yspec, c.bins, c.batch, c.cin, c.cout, c.groups,
                         (cudaStream_t)stream);
}

int fc_inverse(const fc_plan* plan, const void* d_const, const float* d_yspec, const float* d_bias, float* d_y, void* d_ws, void* stream) {
  if (!plan || !d_const || !d_yspec || !d_y || !d_ws) return set_err(FC_ENULL, "fc_inverse: NULL argument");
  if (plan->info.segments > 1) return set_err(FC_EUNSUPPORTED, kSegMsg);
  init_once();
  Bufs b{nullptr, const_cast<float*>(d_yspec), (char*)d_ws + plan->off_sA, (char*)d_ws + plan->off_sB, d_y};
  return run_steps(plan, plan->inv, b, (const float2*)d_const, d_bias, (cudaStream_t)stream);
}

int fc_conv(const fc_plan* plan, const void* d_const, const float* d_x, const float* d_kspec, const float* d_bias, float* d_y, void* d_ws,
            void* stream) {
  if (!plan || !d_const || !d_x || !d_kspec || !d_y || !d_ws) return set_err(FC_ENULL, "fc_conv: NULL argument");
  init_once();
  cudaStream_t st = (cudaStream_t)stream;
  const float2* tw = (const float2*)d_const;
  char* ws = (char*)d_ws;
  void* xspec = ws + plan->off_xspec;
  void* yspec = ws + plan->off_yspec;
  for (const fc_launch& L : plan->prog) {
    Bufs b{d_x, L.spec_is_y ? yspec : xspec, ws + plan->off_sA, ws + plan->off_sB, d_y};
    int rc;
    switch (L.type) {
      case FC_L_PASS:
        rc = launch_pass(plan, L.pass, buf_ptr(b, L.src), buf_ptr(b, L.dst), tw, d_bias, st);
        break;
      case FC_L_FAST_R2C:
        rc = launch_fast_r2c(L.pass, buf_ptr(b, L.src), buf_ptr(b, L.dst), tw, st, (plan->prob.flags & FC_FLAG_STREAM_R2C) != 0);
        break;
      case FC_L_FAST_C2R:
        rc = launch_fast_c2r(L.pass, buf_ptr(b, L.src), buf_ptr(b, L.dst), tw, d_bias, st, !(plan->prob.flags & FC_FLAG_NO_STREAM));
        break;
      case FC_L_FAST_C2C:
        rc = launch_fast_c2c(L.pass, buf_ptr(b, L.src), buf_ptr(b, L.dst), tw, st);
        break;
      case FC_L_PAIR_R2C:
        rc = launch_pair_r2c(plan, L.pass, buf_ptr(b, L.src), buf_ptr(b, L.dst), tw, st);
        break;
      case FC_L_PAIR_C2R:
        rc = launch_pair_c2r(plan, L.pass, buf_ptr(b, L.src), buf_ptr(b, L.dst), tw, d_bias, st);
        break;
      case FC_L_PAIR_FUSED:
        rc = launch_pair_fused(plan, L.fused, buf_ptr(b, L.src), (const float2*)d_kspec, buf_ptr(b, L.dst), tw, st);
        break;
      case FC_L_PAIR_FUSED64:
        rc = launch_pair_fused64(plan, L.fused, buf_ptr(b, L.src), (const float2*)d_kspec, buf_ptr(b, L.dst), tw, st);
        break;
      case FC_L_PLANE_FWD:
      case FC_L_PLANE_INV:
        rc = launch_plane(L.type == FC_L_PLANE_INV, L.plane, buf_ptr(b, L.src), buf_ptr(b, L.dst), tw, plan->tw_len, st);
        break;
      case FC_L_COL_R2C:
      case FC_L_COL_C2R:
        rc = launch_column(L.pass, buf_ptr(b, L.src), buf_ptr(b, L.dst), tw, d_bias, st);
        break;
      case FC_L_LINE_R2C:
      case FC_L_LINE_C2R:
        rc = launch_line(L.pass, buf_ptr(b, L.src), buf_ptr(b, L.dst), tw, d_bias, st);
        break;
      case FC_L_CONTRACT: {
        const fc_contract_desc& c = plan->contract;
        rc = launch_contract((const float2*)xspec, (const float2*)d_kspec, (float2*)yspec, c.bins, c.batch, c.cin, c.cout, c.groups, st);
      } break;
      case FC_L_TC_X: {  // batches [b0, b0 + nb) of the chunk: signal spectrum -> GEMM operand blobs
        const fc_contract_desc& c = plan->contract;
        const int b0 = (int)L.pass.in_os, nb = (int)L.pass.n_outer;
        const int bp = tc_padded_batch(nb);
        rc = FC_OK;
        if (bp != nb) {
          const size_t xtc_bytes = (size_t)c.bins * c.groups * 2 * bp * 2 * (c.cin / c.groups) * 4;
          cudaError_t e = cudaMemsetAsync(ws + plan->off_xtc, 0, xtc_bytes, st);
          if (e != cudaSuccess) rc = set_err((int)e, "tc operand memset failed");
        }
        if (!rc)
          rc = launch_tc_relayout(1, (const float2*)xspec + (int64_t)b0 * c.cin * c.bins, ws + plan->off_xtc, c.bins, nb, c.cin, c.cout, c.groups, st);
      } break;
      case FC_L_TC_FWD: {  // (single batch chunk) last forward pass -> Bt blobs
        const fc_contract_desc& c = plan->contract;
        const int bp = tc_padded_batch(c.batch);
        rc = FC_OK;
        if (bp != c.batch) {
          const size_t xtc_bytes = (size_t)c.bins * c.groups * 2 * bp * 2 * (c.cin / c.groups) * 4;
          cudaError_t e = cudaMemsetAsync(ws + plan->off_xtc, 0, xtc_bytes, st);
          if (e != cudaSuccess) rc = set_err((int)e, "tc operand memset failed");
        }
        if (!rc) rc = launch_tc_c2c(false, plan, L.pass, buf_ptr(b, L.src), ws + plan->off_xtc, tw, st);
      } break;
      case FC_L_TC_INV:  // product D -> lines of the first inverse pass
        rc = launch_tc_c2c(true, plan, L.pass, buf_ptr(b, L.dst), ws + plan->off_ytc, tw, st);
        break;
      case FC_L_TC_GEMM: {
        const fc_contract_desc& c = plan->contract;
        rc = launch_tc_gemm(d_kspec, (const float*)(ws + plan->off_xtc), (float*)(ws + plan->off_ytc), c.bins, (int)L.pass.n_outer, c.cin, c.cout, c.groups, st);
      } break;
      case FC_L_TC_Y: {
        const fc_contract_desc& c = plan->contract;
        const int b0 = (int)L.pass.in_os, nb = (int)L.pass.n_outer;
        rc = launch_tc_relayout(2, ws + plan->off_ytc, (float2*)yspec + (int64_t)b0 * c.cout * c.bins, c.bins, nb, c.cin, c.cout, c.groups, st);
      } break;
      default:
        rc = launch_fused(plan, L.fused, buf_ptr(b, L.src), (const float2*)d_kspec, buf_ptr(b, L.dst), tw, st, d_y, d_bias);
        break;
    }
    if (rc) return rc;
  }
  return FC_OK;
}

int fc_conv_host(const fc_plan* plan, const void* d_const, const float* h_x, float* d_x_stage, const float* d_kspec, const float* d_bias,
                 float* d_y_stage, float* h_y, void* d_ws, void* stream) {
  if (!plan || !h_x || !d_x_stage || !d_y_stage || !h_y) return set_err(FC_ENULL, "fc_conv_host: NULL argument");
  const fc_problem& P = plan->user_prob;  // the caller's tensors (a batch-segmented plan runs a windowed problem)
  int64_t in_elems = (int64_t)P.batch * P.cin;
  for (int i = 0; i < P.ndim; ++i) in_elems *= P.in_size[i];
  cudaError_t e = cudaMemcpyAsync(d_x_stage, h_x, (size_t)in_elems * sizeof(float), cudaMemcpyHostToDevice, (cudaStream_t)stream);
  if (e != cudaSuccess) return set_err((int)e, std::string("H2D copy: ") + cudaGetErrorString(e));
  int rc = fc_conv(plan, d_const, d_x_stage, d_kspec, d_bias, d_y_stage, d_ws, stream);
  if (rc) return rc;
  e = cudaMemcpyAsync(h_y, d_y_stage, (size_t)plan->info.out_elems * sizeof(float), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  if (e != cudaSuccess) return set_err((int)e, std::string("D2H copy: ") + cudaGetErrorString(e));
  return FC_OK;
}

int fc_conv_profiled(const fc_plan* plan, const void* d_const, const float* d_x, const float* d_kspec, const float* d_bias, float* d_y,
                     void* d_ws, void* stream, float* ms_out, int max_n, int* n_out) {
  if (!ms_out || !n_out) return set_err(FC_ENULL, "fc_conv_profiled: NULL argument");
  *n_out = 0;
#ifdef FC_CPU_EMUL
  (void)max_n;
  return fc_conv(plan, d_const, d_x, d_kspec, d_bias, d_y, d_ws, stream);
#else
  Recorder rec;
  rec.st = (cudaStream_t)stream;
  g_rec = &rec;
  rec_mark();
  int rc = fc_conv(plan, d_const, d_x, d_kspec, d_bias, d_y, d_ws, stream);
  g_rec = nullptr;
  cudaError_t e = cudaStreamSynchronize(rec.st);
  if (!rc && e != cudaSuccess) rc = set_err((int)e, std::string("fc_conv_profiled: ") + cudaGetErrorString(e));
  int n = 0;
  for (size_t i = 1; i < rec.ev.size(); ++i) {
    float ms = 0.f;
    if (!rc) cudaEventElapsedTime(&ms, rec.ev[i - 1], rec.ev[i]);
    if (n < max_n) ms_out[n++] = ms;
  }
  for (cudaEvent_t ev : rec.ev) cudaEventDestroy(ev);
  *n_out = n;
  return rc;
#endif
}

int fc_plan_launch_info(const fc_plan* plan, int i, char* name, size_t namelen, int64_t* algo_bytes) {
  if (!plan || !name || !namelen || !algo_bytes) return set_err(FC_ENULL, "fc_plan_launch_info: NULL argument");
  if (i < 0 || i >= (int)plan->prog.size()) return set_err(FC_EINVAL, "fc_plan_launch_info: index out of range");
  const fc_launch& L = plan->prog[i];
  size_t n = L.name.size() < namelen - 1 ? L.name.size() : namelen - 1;
  std::memcpy(name, L.name.data(), n);
  name[n] = 0;
  *algo_bytes = L.bytes;
  return FC_OK;
}

int fc_tc_supported(int64_t batch, int64_t cin, int64_t cout, int64_t groups) {
  return tc_supported((int)batch, (int)cin, (int)cout, (int)groups) ? 1 : 0;
}

int64_t fc_tc_scratch_bytes(int64_t batch, int64_t cin, int64_t cout, int64_t groups, int64_t bins) {
  const int64_t bp = tc_padded_batch((int)batch), I = cin / groups;
  return (bins * groups * 2 * bp * 2 * I * 4 + 255) / 256 * 256 + bins * cout * 2 * bp * 4 + 512;
}

int fc_tc_prepare_kernel(const float* d_kspec, float* d_kspec_tc, int64_t cin, int64_t cout, int64_t groups, int64_t bins, void* stream) {
  if (!d_kspec || !d_kspec_tc) return set_err(FC_ENULL, "fc_tc_prepare_kernel: NULL argument");
  init_once();
  return launch_tc_relayout(0, d_kspec, d_kspec_tc, bins, 1, (int)cin, (int)cout, (int)groups, (cudaStream_t)stream);
}

int fc_tc_complex_matmul(const float* d_a, const float* d_b_tc, float* d_y, void* d_scratch, int64_t batch, int64_t cin, int64_t cout,
                         int64_t groups, int64_t bins, void* stream) {
  if (!d_a || !d_b_tc || !d_y || !d_scratch) return set_err(FC_ENULL, "fc_tc_complex_matmul: NULL argument");
  if (!tc_supported((int)batch, (int)cin, (int)cout, (int)groups)) return set_err(FC_EUNSUPPORTED, "fc_tc_complex_matmul: shape not supported");
  init_once();
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t bp = tc_padded_batch((int)batch), I = cin / groups;
  float* xtc = (float*)d_scratch;
  const int64_t xtc_bytes = (bins * groups * 2 * bp * 2 * I * 4 + 255) / 256 * 256;
  float* ytc = (float*)((char*)d_scratch + xtc_bytes);
  if (bp != batch) {
    cudaError_t e = cudaMemsetAsync(xtc, 0, (size_t)(bins * groups * 2 * bp * 2 * I * 4), st);
    if (e != cudaSuccess) return set_err((int)e, "fc_tc_complex_matmul: memset failed");
  }
  int rc = launch_tc_relayout(1, d_a, xtc, bins, (int)batch, (int)cin, (int)cout, (int)groups, st);
  if (rc) return rc;
  rc = launch_tc_gemm(d_b_tc, xtc, ytc, bins, (int)batch, (int)cin, (int)cout, (int)groups, st);
  if (rc) return rc;
  return launch_tc_relayout(2, ytc, d_y, bins, (int)batch, (int)cin, (int)cout, (int)groups, st);
}

int fc_complex_matmul(const float* d_a, const float* d_b, float* d_y, int64_t batch, int64_t cin, int64_t cout, int64_t groups, int64_t bins,
                      void* stream) {
  if (!d_a || !d_b || !d_y) return set_err(FC_ENULL, "fc_complex_matmul: NULL argument");
  if (batch < 1 || cin < 1 || cout < 1 || groups < 1 || bins < 1 || cin % groups || cout % groups)
    return set_err(FC_EINVAL, "fc_complex_matmul: bad shape");
  init_once();
  return launch_contract((const float2*)d_a, (const float2*)d_b, (float2*)d_y, bins, (int)batch, (int)cin, (int)cout, (int)groups,
                         (cudaStream_t)stream);
}

}  // extern "C"

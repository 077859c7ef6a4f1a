// fc_column.cuh — thread-per-column kernels for the strided pass of the four-step 1-d transform (N = 64 * N2).
//
// The signal line of N samples is viewed as a 64 x N2 matrix (n = n1*N2 + n2). The strided pass transforms the 64
// samples of a column; adjacent columns are adjacent in memory on both sides of the pass, so one thread per column
// makes every request of a warp a contiguous 128 / 256-byte run with no transposition at all, and a 64-point real
// transform (32 complex points after packing) fits the register file: no shared memory, no barrier, all 64 loads of
// a thread in flight at once. The contiguous pass over n2 (N2 = 256 .. 2048) runs on fc_fast_c2c_kernel.
//   fc_col_r2c_kernel  gather map (pad / zero-stuffing) -> packed real FFT -> four-step twiddle W_N^(k1*n2) -> [k1][n2]
//   fc_col_c2r_kernel  [k1][n2] -> conj twiddle -> Hermitian inverse -> crop / stride / lattice map + bias
#pragma once
#include "fc_kernels.cuh"

#define FC_COL_N 64 /* real points per column */

// exp(-2*pi*i*j/64): compile-time constants once the loops below are unrolled
FC_DEV float2 fc_w64(int j) {
  constexpr float kCos[64] = {
      1.f, 0.995184727f, 0.98078528f, 0.956940336f, 0.923879533f, 0.881921264f, 0.831469612f, 0.773010453f,
      0.707106781f, 0.634393284f, 0.555570233f, 0.471396737f, 0.382683432f, 0.290284677f, 0.195090322f, 0.0980171403f,
      0.f, -0.0980171403f, -0.195090322f, -0.290284677f, -0.382683432f, -0.471396737f, -0.555570233f, -0.634393284f,
      -0.707106781f, -0.773010453f, -0.831469612f, -0.881921264f, -0.923879533f, -0.956940336f, -0.98078528f, -0.995184727f,
      -1.f, -0.995184727f, -0.98078528f, -0.956940336f, -0.923879533f, -0.881921264f, -0.831469612f, -0.773010453f,
      -0.707106781f, -0.634393284f, -0.555570233f, -0.471396737f, -0.382683432f, -0.290284677f, -0.195090322f, -0.0980171403f,
      0.f, 0.0980171403f, 0.195090322f, 0.290284677f, 0.382683432f, 0.471396737f, 0.555570233f, 0.634393284f,
      0.707106781f, 0.773010453f, 0.831469612f, 0.881921264f, 0.923879533f, 0.956940336f, 0.98078528f, 0.995184727f};
  constexpr float kSin[64] = {
      0.f, 0.0980171403f, 0.195090322f, 0.290284677f, 0.382683432f, 0.471396737f, 0.555570233f, 0.634393284f,
      0.707106781f, 0.773010453f, 0.831469612f, 0.881921264f, 0.923879533f, 0.956940336f, 0.98078528f, 0.995184727f,
      1.f, 0.995184727f, 0.98078528f, 0.956940336f, 0.923879533f, 0.881921264f, 0.831469612f, 0.773010453f,
      0.707106781f, 0.634393284f, 0.555570233f, 0.471396737f, 0.382683432f, 0.290284677f, 0.195090322f, 0.0980171403f,
      0.f, -0.0980171403f, -0.195090322f, -0.290284677f, -0.382683432f, -0.471396737f, -0.555570233f, -0.634393284f,
      -0.707106781f, -0.773010453f, -0.831469612f, -0.881921264f, -0.923879533f, -0.956940336f, -0.98078528f, -0.995184727f,
      -1.f, -0.995184727f, -0.98078528f, -0.956940336f, -0.923879533f, -0.881921264f, -0.831469612f, -0.773010453f,
      -0.707106781f, -0.634393284f, -0.555570233f, -0.471396737f, -0.382683432f, -0.290284677f, -0.195090322f, -0.0980171403f};
  return make_float2(kCos[j & 63], -kSin[j & 63]);
}

// Forward 32-point DFT in registers, natural order in and out: n = 4*n1 + n2, k = k1 + 8*k2.
FC_DEV void fc_fft32(float2 (&z)[32]) {
  float2 A[4][8];
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) {
    float2 a[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) a[r] = z[4 * r + n2];
    fc_butterfly<8>(a);
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) A[n2][k1] = (n2 * k1 == 0) ? a[k1] : fc_mul(a[k1], fc_w64(2 * n2 * k1));
  }
#pragma unroll
  for (int k1 = 0; k1 < 8; ++k1) {
    float2 b[4];
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) b[n2] = A[n2][k1];
    fc_butterfly<4>(b);
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) z[k1 + 8 * k2] = b[k2];
  }
}

// Four-step twiddles W_N^(k*r), k = 0..32, of column r: exact table values at k = 1 and k = 8j, products in between
// (at most two roundings away from a table entry). w[k] for k = 8j + i is anchor[j] * low[i].
struct fc_col_twiddles {
  float2 low[8];     // W^(i*r), i < 8
  float2 anchor[5];  // W^(8j*r), j <= 4
  FC_DEV void init(int64_t r, const fc_pass& p, const float2* tw) {
    low[0] = make_float2(1.f, 0.f);
    const float2 w1 = fc_big_twiddle(1, r, p, tw);
    low[1] = w1;
    low[2] = fc_mul(w1, w1);
    low[3] = fc_mul(low[2], w1);
    low[4] = fc_mul(low[2], low[2]);
    low[5] = fc_mul(low[4], w1);
    low[6] = fc_mul(low[4], low[2]);
    low[7] = fc_mul(low[4], low[3]);
    anchor[0] = make_float2(1.f, 0.f);
#pragma unroll
    for (int j = 1; j <= 4; ++j) anchor[j] = fc_big_twiddle(8 * j, r, p, tw);
  }
  FC_DEV float2 at(int k) const { return (k & 7) == 0 ? anchor[k >> 3] : (k < 8 ? low[k] : fc_mul(anchor[k >> 3], low[k & 7])); }
};

struct fc_col_args {
  fc_pass p;
  const void* in;
  void* out;
  const float2* tw;
  const float* bias;  // C2R only, may be null
};

// PLAIN: constant padding mode without zero-stuffing / subsampling (the gather is a shift by the padding).
template <bool PLAIN>
__global__ void __launch_bounds__(128, 4) fc_col_r2c_kernel(fc_col_args a) {
  fc_grid_dep_sync();
  constexpr int M = FC_COL_N / 2;
  const fc_pass& p = a.p;
  const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= p.n_outer * p.R) return;
  const int64_t o = id / p.R, r = id - o * p.R;
  const int64_t o1 = o / p.o_c2, o2 = o - o1 * p.o_c2;
  const float* x = reinterpret_cast<const float*>(a.in) + (o1 / p.o_q) * p.o_sA + (o1 % p.o_q) * p.o_sB + o2 * p.o_sC;
  fc_imap im = p.imap;
  if (p.bseg_n > 1) im.pad -= (int)(o1 % p.bseg_n) * p.bseg_V;  // batch segment (o_c2 = bseg_c, o_q = bseg_n): its window of the line
  const int hi = (im.ext - im.pad < im.L) ? im.ext - im.pad : im.L;  // PLAIN: source index s = u - pad is live for 0 <= s < hi
  float2 z[M];
#pragma unroll
  for (int m = 0; m < M; ++m) {
    float v[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int u = (2 * m + h) * p.pos_n + (int)r;  // dense position (pos_r == 1)
      int s;
      if (PLAIN) {
        s = u - im.pad;
        if (s >= hi) s = -1;
      } else {
        s = fc_imap_src(im, u);
      }
      v[h] = s >= 0 ? __ldg(x + s) : 0.f;
    }
    z[m] = make_float2(v[0], v[1]);
  }
  fc_fft32(z);
  fc_col_twiddles w;
  w.init(r, p, a.tw);
  float2* y = reinterpret_cast<float2*>(a.out) + o * p.out_os + r;
  // untangle the packed transform (same algebra as the generic R2C pass), four-step twiddle, store bins 0..M
#pragma unroll
  for (int k = 0; k <= M; ++k) {
    float2 X;
    if (k == 0 || k == M) {
      X = make_float2(k == 0 ? z[0].x + z[0].y : z[0].x - z[0].y, 0.f);
    } else {
      const float2 zk = z[k], zc = fc_conj(z[M - k]);
      const float2 e = fc_scale(fc_add(zk, zc), 0.5f);
      const float2 od = fc_scale(fc_mul_mi(fc_sub(zk, zc)), 0.5f);
      X = fc_add(e, fc_mul(fc_w64(k), od));
    }
    if (k > 0) X = fc_mul(X, w.at(k));
    X = fc_scale(X, p.scale);
    if (p.conj_out) X = fc_conj(X);
    y[(int64_t)k * p.out_es] = X;
  }
}

__global__ void __launch_bounds__(128) fc_col_c2r_kernel(fc_col_args a) {
  fc_grid_dep_sync();
  constexpr int M = FC_COL_N / 2;
  const fc_pass& p = a.p;
  const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= p.n_outer * p.R) return;
  const int64_t o = id / p.R, r = id - o * p.R;
  const float2* yin = reinterpret_cast<const float2*>(a.in) + o * p.in_os + r;
  float2 Y[M + 1];
#pragma unroll
  for (int k = 0; k <= M; ++k) Y[k] = __ldg(yin + (int64_t)k * p.in_es);
  fc_col_twiddles w;
  w.init(r, p, a.tw);
#pragma unroll
  for (int k = 1; k <= M; ++k) Y[k] = fc_mul(Y[k], fc_conj(w.at(k)));
  // Hermitian half spectrum -> conj(Z) of the packed transform; a forward FFT then yields conj(z)
  float2 z[M];
#pragma unroll
  for (int k = 0; k < M; ++k) {
    const float2 yk = Y[k], ym = fc_conj(Y[M - k]);
    const float2 s = fc_add(yk, ym);
    const float2 d = fc_mul(fc_sub(yk, ym), fc_conj(fc_w64(k)));
    z[k] = make_float2(s.x - d.y, -(s.y + d.x));
  }
  fc_fft32(z);
  const fc_omap om = p.omap;
  const float b = (p.has_bias && a.bias) ? __ldg(a.bias + (o % (p.cout > 0 ? p.cout : 1))) : 0.f;
  float* y = reinterpret_cast<float*>(a.out) + o * p.out_os;
  int lout = om.Lout;
  if (p.bseg_n > 1) {  // batch segment: its run of the user's output line
    const int64_t ob = o / p.bseg_c;
    const int sg = (int)(ob % p.bseg_n);
    y = reinterpret_cast<float*>(a.out) + ((ob / p.bseg_n) * p.bseg_c + (o - ob * p.bseg_c)) * (int64_t)p.bseg_Lout + (int64_t)sg * p.bseg_Vo;
    const int left = p.bseg_Lout - sg * p.bseg_Vo;
    if (left < lout) lout = left;
  }
  const bool plain = om.og == 1 && om.os == 1;
#pragma unroll
  for (int n = 0; n < FC_COL_N; ++n) {
    const float val = (n & 1) ? -z[n >> 1].y : z[n >> 1].x;
    const int u = n * p.pos_n + (int)r;  // dense position
    if (plain) {
      const int j = u - om.ob;
      if (j >= 0 && j < lout) y[j] = ((u < om.lim) ? val : 0.f) + b;
    } else {
      for (int e = 0; e < om.og; ++e) {  // dense sample u owns the outputs j with (j*os + ob) / og == u
        const int t = u * om.og + e - om.ob;
        if (t < 0) continue;
        int j = t;
        if (om.os != 1) {
          if (t % om.os) continue;
          j = t / om.os;
        }
        if (j >= lout) continue;
        y[j] = ((e == 0 && u < om.lim) ? val : 0.f) + b;
      }
    }
  }
}

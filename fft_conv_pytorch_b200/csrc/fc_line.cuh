// fc_line.cuh — the real passes of short 1-d programs (lines of 512 or 1024 real points: one pass, nothing to transpose)
// on the warp engine of fc_fused.cuh.
//
// The generic block-level pass moves such lines at ~1 TB/s (profiles/r2b_shape_probe.txt). Here a warp owns NL = 2 lines
// from load to store, as in fc_fast_c2c_kernel: coalesced loads straight into the "lane + 32q" register layout, the packed
// half-length transform with a warp-private exchange line, untangle (or pre-twist) against the partner bins and coalesced
// stores; no block barrier, no staging tile.
//   fc_line_r2c_kernel  real line (zero-padding gather, batch-segment windows) -> half spectrum, scale / conjugate on store
//   fc_line_c2r_kernel  half spectrum -> real line, plain crop (+ batch-segment runs) and bias on store
// Replaces (reference functional.py) :60-62 F.pad, :70 rfftn, :75 irfftn, :76-87 crop + bias for 1-d signals.
#pragma once
#include "fc_fused.cuh"

struct fc_line_args {
  fc_pass p;  // FC_R2C or FC_C2R with R == 1, contiguous lines on both sides
  const void* in;
  void* out;
  const float2* tw;
  const float* bias;  // C2R only, may be null
};

template <int M, int NW, int OCC>
__global__ void __launch_bounds__(NW * 32, OCC) fc_line_r2c_kernel(fc_line_args a) {
  fc_grid_dep_sync();
  constexpr int E = M / 32, NL = 2;
  const fc_pass& p = a.p;
  FC_DYN_SMEM(smem);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float2* line0 = smem + (size_t)(w * NL) * M;
  fc_wofs ofs;
  ofs.init(lane);
  const fc_imap im = p.imap;
  const int tstep = p.tw_len / (2 * M);
  const float* x = reinterpret_cast<const float*>(a.in);
  float2* y = reinterpret_cast<float2*>(a.out);
  const int64_t n_lines = p.n_outer;
  for (int64_t g0 = ((int64_t)blockIdx.x * NW + w) * NL; g0 < n_lines; g0 += (int64_t)gridDim.x * NW * NL) {  // warp-uniform
    float2 v[NL][E];
    bool ok[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const int64_t o = g0 + l;
      ok[l] = o < n_lines;
      const int64_t oo = ok[l] ? o : 0;
      const int64_t o1 = oo / p.o_c2, o2 = oo - o1 * p.o_c2;
      const float* src = x + (o1 / p.o_q) * p.o_sA + (o1 % p.o_q) * p.o_sB + o2 * p.o_sC;
      const int pad = im.pad - (p.bseg_n > 1 ? (int)(o1 % p.bseg_n) * p.bseg_V : 0);  // batch segment: its window of the line
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const int u = 2 * (lane + 32 * q), s = u - pad;  // dense positions u, u + 1 <- source s, s + 1 (zero outside [0, L) and beyond ext)
        float re = 0.f, imv = 0.f;
        if (ok[l]) {
          if (u < im.ext && s >= 0 && s < im.L) re = __ldg(src + s);
          if (u + 1 < im.ext && s + 1 >= 0 && s + 1 < im.L) imv = __ldg(src + s + 1);
        }
        v[l][q] = make_float2(re, imv);
      }
    }
    fc_wfft<M, NL, M>(v, line0, ofs, a.tw, p.tw_len, lane);
    fc_wwrite<M, NL, M>(v, line0, ofs);
    FC_SYNCWARP();
    float nyq[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const float2 z0 = line0[l * M + fc_swz2(0)];
      nyq[l] = z0.x - z0.y;
    }
#pragma unroll
    for (int q = 0; q < E; ++q) {
      const int k = lane + 32 * q;
      const float2 wk = __ldg(a.tw + k * tstep);
      const int km = fc_swz2((M - k) & (M - 1));
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        const float2 zk = v[l][q];
        const float2 zc = fc_conj(line0[l * M + km]);
        const float2 e = fc_scale(fc_add(zk, zc), 0.5f);
        const float2 od = fc_scale(fc_mul_mi(fc_sub(zk, zc)), 0.5f);
        v[l][q] = fc_add(e, fc_mul(wk, od));
      }
    }
    FC_SYNCWARP();  // every partner has been read before the next pair of lines reuses the exchange buffers
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      if (!ok[l]) continue;
      float2* dst = y + (g0 + l) * p.out_os;
#pragma unroll
      for (int q = 0; q < E; ++q) {
        float2 val = fc_scale(v[l][q], p.scale);
        if (p.conj_out) val = fc_conj(val);
        dst[lane + 32 * q] = val;
      }
      if (lane == 0) dst[M] = make_float2(nyq[l] * p.scale, 0.f);
    }
  }
}

template <int M, int NW, int OCC>
__global__ void __launch_bounds__(NW * 32, OCC) fc_line_c2r_kernel(fc_line_args a) {
  fc_grid_dep_sync();
  constexpr int E = M / 32, NL = 2;
  const fc_pass& p = a.p;
  FC_DYN_SMEM(smem);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float2* line0 = smem + (size_t)(w * NL) * M;
  fc_wofs ofs;
  ofs.init(lane);
  const fc_omap om = p.omap;
  const int tstep = p.tw_len / (2 * M);
  const float2* xin = reinterpret_cast<const float2*>(a.in);
  float* y = reinterpret_cast<float*>(a.out);
  const int64_t n_lines = p.n_outer;
  for (int64_t g0 = ((int64_t)blockIdx.x * NW + w) * NL; g0 < n_lines; g0 += (int64_t)gridDim.x * NW * NL) {  // warp-uniform
    float2 v[NL][E];
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const int64_t o = g0 + l < n_lines ? g0 + l : 0;
      const float2* src = xin + o * p.in_os;
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const int k = lane + 32 * q;
        const float2 wk = fc_conj(__ldg(a.tw + k * tstep));
        const float2 yk = __ldg(src + k);
        const float2 ym = fc_conj(__ldg(src + M - k));  // k = 0: the Nyquist bin
        const float2 s = fc_add(yk, ym);
        const float2 d = fc_mul(fc_sub(yk, ym), wk);
        v[l][q] = make_float2(s.x - d.y, -(s.y + d.x));  // conj(Z[k]), Z = s + i*d
      }
    }
    fc_wfft<M, NL, M>(v, line0, ofs, a.tw, p.tw_len, lane);
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const int64_t o = g0 + l;
      if (o >= n_lines) continue;
      const float b = (p.has_bias && a.bias) ? __ldg(a.bias + (int)(o % (p.cout > 0 ? p.cout : 1))) : 0.f;
      float* dst = y + o * p.out_os;
      int lout = om.Lout;
      if (p.bseg_n > 1) {  // batch segment: its run of the caller's output line
        const int64_t ob = o / p.bseg_c;
        const int sg = (int)(ob % p.bseg_n);
        dst = y + ((ob / p.bseg_n) * p.bseg_c + (o - ob * p.bseg_c)) * (int64_t)p.bseg_Lout + (int64_t)sg * p.bseg_Vo;
        const int left = p.bseg_Lout - sg * p.bseg_Vo;
        if (left < lout) lout = left;
      }
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const int u = 2 * (lane + 32 * q);  // dense positions u, u + 1 (plain crop: output j = u - ob)
        const int j = u - om.ob;
        const float r0 = (u < om.lim ? v[l][q].x : 0.f) + b, r1 = (u + 1 < om.lim ? -v[l][q].y : 0.f) + b;
        if (j >= 0 && j < lout) dst[j] = r0;
        if (j + 1 >= 0 && j + 1 < lout) dst[j + 1] = r1;
      }
    }
  }
}

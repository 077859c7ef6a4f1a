// fc_fused.cuh — specialised sm_100a kernels for the hot shapes (checked against the generic fc_kernels.cuh).
//
// All three use the same register-resident warp FFT: one warp owns one line of M complex points, each lane keeps
// E = M/32 points in registers in the "lane + 32q" layout (which is both the coalesced HBM layout and the input
// layout of every Stockham stage), and consecutive radix-8/4 stages exchange data through a warp-private,
// XOR-swizzled shared-memory line with __syncwarp only — no block-wide barrier inside a transform.
//
//   fc_fast_r2c_kernel   K1: coalesced real rows -> R2C -> half spectrum stored transposed ([bin][row], 128-byte
//                        segments), so that the next axis is contiguous
//   fc_fused_axis_kernel KB: for one bin of the other axes: forward transform of the last remaining axis for every
//                        input channel, the per-bin grouped channel contraction with the cached kernel spectrum
//                        (complex_matmul, reference functional.py:11-16) and the inverse transform of that axis
//                        for every output channel, without the two spectrum round trips through HBM
//   fc_fast_c2r_kernel   K4: transposed load -> C2R -> crop / stride + bias -> coalesced real rows
#pragma once
#include "fc_kernels.cuh"

#define FC_SYNCWARP() __syncwarp()

// Swizzle of a warp's exchange line. For the three access patterns of the stages (16 consecutive points;
// stride-8 writes of the first stage; "8 consecutive, jump 64" writes of the second stage) the 16 lanes of a half
// warp touch 16 distinct 8-byte bank pairs.
FC_DEV int fc_swz2(int p) {
  const int h = p >> 4;
  return p ^ ((h & 7) | (((h >> 2) & 1) << 3));
}

// One radix-R stage on registers. v[t + NBF*r] = input r of butterfly t (= element lane + 32*(t + NBF*r)).
template <int M, int R, int Ns>
FC_DEV void fc_wstage(float2 (&v)[M / 32], const float2* tw, int tw_len, int lane) {
  constexpr int E = M / 32, NBF = E / R;
#pragma unroll
  for (int t = 0; t < NBF; ++t) {
    float2 a[R];
#pragma unroll
    for (int r = 0; r < R; ++r) a[r] = v[t + NBF * r];
    if (Ns > 1) {
      const int k = (lane + 32 * t) & (Ns - 1);
      const float2 w1 = __ldg(tw + k * (tw_len / (Ns * R)));
      float2 w = w1;
#pragma unroll
      for (int r = 1; r < R; ++r) {
        a[r] = fc_mul(a[r], w);
        if (r + 1 < R) w = fc_mul(w, w1);
      }
    }
    fc_butterfly<R>(a);
#pragma unroll
    for (int r = 0; r < R; ++r) v[t + NBF * r] = a[r];
  }
}

// Per-lane swizzled offsets of the exchange patterns, computed once per kernel (the XOR swizzle makes them
// non-affine in the lane id, so evaluating fc_swz2 at every access would cost more than the butterflies):
//   rd[j]  lane + 32q layout, q % 4 == j : element offset = rd[j] + 32*q                  (all stages read this way)
//   w1[r]  outputs of the (R=8, Ns=1) stage: 8*lane + r              (+ 256 per extra butterfly t)
//   w2[r]  outputs of the (R=8, Ns=8) stage: 64*(lane>>3) + (lane&7) + 8r  (+ 256 per extra butterfly t)
// fc_swz2(p + 256t) == fc_swz2(p) + 256t and fc_swz2(lane + 32(q+4)) == fc_swz2(lane + 32q) + 128 because the
// swizzle only looks at bits 4..6 of p.
struct fc_wofs {
  int rd[4];
  int w1[8];
  int w2[8];
  FC_DEV void init(int lane) {
#pragma unroll
    for (int j = 0; j < 4; ++j) rd[j] = fc_swz2(lane + 32 * j) - 32 * j;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      w1[r] = fc_swz2(8 * lane + r);
      w2[r] = fc_swz2(64 * (lane >> 3) + (lane & 7) + 8 * r);
    }
  }
};

template <int M>
FC_DEV void fc_wread(float2 (&v)[M / 32], const float2* line, const fc_wofs& o) {
#pragma unroll
  for (int q = 0; q < M / 32; ++q) v[q] = line[o.rd[q & 3] + 32 * q];
}
template <int M>
FC_DEV void fc_wwrite(const float2 (&v)[M / 32], float2* line, const fc_wofs& o) {
#pragma unroll
  for (int q = 0; q < M / 32; ++q) line[o.rd[q & 3] + 32 * q] = v[q];
}

// Exchange after the (8, Ns=1) or (8, Ns=8) stage: outputs go to their Stockham positions in the warp's line,
// then every lane reads the lane + 32q layout back.
template <int M, int Ns>
FC_DEV void fc_wxchg8(float2 (&v)[M / 32], float2* line, const fc_wofs& o) {
  constexpr int E = M / 32, NBF = E / 8;
#pragma unroll
  for (int t = 0; t < NBF; ++t)
#pragma unroll
    for (int r = 0; r < 8; ++r) line[(Ns == 1 ? o.w1[r] : o.w2[r]) + 256 * t] = v[t + NBF * r];
  FC_SYNCWARP();
  fc_wread<M>(v, line, o);
  FC_SYNCWARP();
}

// Generic exchange (any stage), used by the long transforms only.
template <int M, int R, int Ns>
FC_DEV void fc_wxchg(float2 (&v)[M / 32], float2* line, int lane) {
  constexpr int E = M / 32, NBF = E / R;
#pragma unroll
  for (int t = 0; t < NBF; ++t) {
    const int j = lane + 32 * t;
    const int k = j & (Ns - 1);
    const int j0 = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) line[fc_swz2(j0 + r * Ns)] = v[t + NBF * r];
  }
  FC_SYNCWARP();
#pragma unroll
  for (int q = 0; q < E; ++q) v[q] = line[fc_swz2(lane + 32 * q)];
  FC_SYNCWARP();
}

// Forward, unnormalised FFT of M points held by one warp. `line` = M float2 of warp-private shared memory.
template <int M>
struct fc_wfft;

template <>
struct fc_wfft<256> {
  static FC_DEV void run(float2 (&v)[8], float2* line, const fc_wofs& o, const float2* tw, int tw_len, int lane) {
    fc_wstage<256, 8, 1>(v, tw, tw_len, lane);
    fc_wxchg8<256, 1>(v, line, o);
    fc_wstage<256, 8, 8>(v, tw, tw_len, lane);
    fc_wxchg8<256, 8>(v, line, o);
    fc_wstage<256, 4, 64>(v, tw, tw_len, lane);
  }
};

template <>
struct fc_wfft<512> {
  static FC_DEV void run(float2 (&v)[16], float2* line, const fc_wofs& o, const float2* tw, int tw_len, int lane) {
    fc_wstage<512, 8, 1>(v, tw, tw_len, lane);
    fc_wxchg8<512, 1>(v, line, o);
    fc_wstage<512, 8, 8>(v, tw, tw_len, lane);
    fc_wxchg8<512, 8>(v, line, o);
    fc_wstage<512, 8, 64>(v, tw, tw_len, lane);
  }
};

template <>
struct fc_wfft<1024> {
  static FC_DEV void run(float2 (&v)[32], float2* line, const fc_wofs& o, const float2* tw, int tw_len, int lane) {
    fc_wstage<1024, 8, 1>(v, tw, tw_len, lane);
    fc_wxchg8<1024, 1>(v, line, o);
    fc_wstage<1024, 8, 8>(v, tw, tw_len, lane);
    fc_wxchg8<1024, 8>(v, line, o);
    fc_wstage<1024, 8, 64>(v, tw, tw_len, lane);
    fc_wxchg<1024, 8, 64>(v, line, lane);
    fc_wstage<1024, 2, 512>(v, tw, tw_len, lane);
  }
};

#define FC_FAST_WARPS 8
#define FC_FAST_TR 16 /* lines per tile of the transposing kernels */

// ------------------------------------------------------------------------------------------------ K1
struct fc_fast_r2c_args {
  fc_pass p;
  const float* x;
  float2* out;
  const float2* tw;
};

// Shared memory: FC_FAST_WARPS lines of M float2 + a (M+1) x (TR+1) transposition tile.
template <int M>
__global__ void __launch_bounds__(FC_FAST_WARPS * 32, (M <= 256) ? 3 : 2) fc_fast_r2c_kernel(fc_fast_r2c_args a) {
  constexpr int E = M / 32, TR = FC_FAST_TR, TP = TR + 1;
  const fc_pass& p = a.p;
  FC_DYN_SMEM(smem);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  float2* line = smem + w * M;
  float2* tile = smem + FC_FAST_WARPS * M;
  const int L = p.imap.L;
  const int tstep = p.tw_len / (2 * M);
  fc_wofs ofs;
  ofs.init(lane);
  for (int64_t t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
    const int64_t o = t / p.tiles_per_outer;
    const int64_t r0 = (t - o * p.tiles_per_outer) * TR;
    const int64_t o1 = o / p.o_c2, o2 = o - o1 * p.o_c2;
    const int64_t base = (o1 / p.o_q) * p.o_sA + (o1 % p.o_q) * p.o_sB + o2 * p.o_sC;
#pragma unroll 1
    for (int it = 0; it < TR / FC_FAST_WARPS; ++it) {
      const int l = w + FC_FAST_WARPS * it;
      const int64_t r = r0 + l;
      const bool valid = r < p.R;
      const float* row = a.x + base + (valid ? r : 0) * p.in_rs;
      float2 v[E];
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const int i0 = 2 * (lane + 32 * q);
        float2 val = make_float2(0.f, 0.f);
        if (valid) {
          if (i0 + 1 < L)
            val = __ldg(reinterpret_cast<const float2*>(row + i0));
          else if (i0 < L)
            val.x = __ldg(row + i0);
        }
        v[q] = val;
      }
      fc_wfft<M>::run(v, line, ofs, a.tw, p.tw_len, lane);
      fc_wwrite<M>(v, line, ofs);
      FC_SYNCWARP();
      // untangle the packed real transform (same algebra as the generic R2C pass)
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const int k = lane + 32 * q;
        const float2 zk = v[q];
        const float2 zc = fc_conj(line[fc_swz2((M - k) & (M - 1))]);
        const float2 e = fc_scale(fc_add(zk, zc), 0.5f);
        const float2 od = fc_scale(fc_mul_mi(fc_sub(zk, zc)), 0.5f);
        tile[k * TP + l] = fc_add(e, fc_mul(__ldg(a.tw + k * tstep), od));
      }
      if (lane == 0) {  // Nyquist bin k = M: E[0] - O[0]
        const float2 z0 = line[fc_swz2(0)];
        tile[M * TP + l] = make_float2(z0.x - z0.y, 0.f);
      }
      FC_SYNCWARP();
    }
    __syncthreads();
    for (int idx = tid; idx < (M + 1) * TR; idx += FC_FAST_WARPS * 32) {
      const int l = idx & (TR - 1), k = idx >> 4;
      if (r0 + l < p.R) a.out[o * p.out_os + (int64_t)k * p.out_es + r0 + l] = tile[k * TP + l];
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ K4
struct fc_fast_c2r_args {
  fc_pass p;
  const float2* in;
  float* out;
  const float2* tw;
  const float* bias;
};

template <int M>
__global__ void __launch_bounds__(FC_FAST_WARPS * 32, (M <= 256) ? 3 : 2) fc_fast_c2r_kernel(fc_fast_c2r_args a) {
  constexpr int E = M / 32, TR = FC_FAST_TR, TP = TR + 1;
  const fc_pass& p = a.p;
  FC_DYN_SMEM(smem);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  float2* line = smem + w * M;
  float2* tile = smem + FC_FAST_WARPS * M;
  const int tstep = p.tw_len / (2 * M);
  const fc_omap om = p.omap;
  fc_wofs ofs;
  ofs.init(lane);
  for (int64_t t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
    const int64_t o = t / p.tiles_per_outer;
    const int64_t r0 = (t - o * p.tiles_per_outer) * TR;
    for (int idx = tid; idx < (M + 1) * TR; idx += FC_FAST_WARPS * 32) {
      const int l = idx & (TR - 1), k = idx >> 4;
      tile[k * TP + l] = (r0 + l < p.R) ? __ldg(a.in + o * p.in_os + (int64_t)k * p.in_es + r0 + l) : make_float2(0.f, 0.f);
    }
    __syncthreads();
    const float b = p.has_bias ? __ldg(a.bias + (int)(o % p.cout)) : 0.f;
#pragma unroll 1
    for (int it = 0; it < TR / FC_FAST_WARPS; ++it) {
      const int l = w + FC_FAST_WARPS * it;
      const int64_t r = r0 + l;
      float2 v[E];
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const int k = lane + 32 * q;
        const float2 yk = tile[k * TP + l];
        const float2 ym = fc_conj(tile[(M - k) * TP + l]);
        const float2 s = fc_add(yk, ym);
        const float2 d = fc_mul(fc_sub(yk, ym), fc_conj(__ldg(a.tw + k * tstep)));
        v[q] = make_float2(s.x - d.y, -(s.y + d.x));  // conj(Z[k]), Z = s + i*d
      }
      fc_wfft<M>::run(v, line, ofs, a.tw, p.tw_len, lane);
      if (r < p.R) {
        float* yrow = a.out + o * p.out_os + r * p.out_rs;
        if (om.os == 1 && om.ob == 0 && om.og == 1 && !(om.Lout & 1) && !(p.out_rs & 1) && !(p.out_os & 1)) {
#pragma unroll
          for (int q = 0; q < E; ++q) {
            const int n0 = 2 * (lane + 32 * q);
            if (n0 < om.Lout) *reinterpret_cast<float2*>(yrow + n0) = make_float2(v[q].x + b, -v[q].y + b);
          }
        } else {
          // general crop / stride / lattice map: stage the real row in the warp's line and scatter from there
          float* rl = reinterpret_cast<float*>(line);
#pragma unroll
          for (int q = 0; q < E; ++q) line[lane + 32 * q] = make_float2(v[q].x, -v[q].y);
          FC_SYNCWARP();
          for (int n = lane; n < 2 * M; n += 32) {
            const float val = rl[n];
            for (int e = 0; e < om.og; ++e) {
              const int tt = n * om.og + e - om.ob;
              if (tt < 0 || (tt % om.os)) continue;
              const int j = tt / om.os;
              if (j >= om.Lout) continue;
              yrow[j] = ((e == 0 && n < om.lim) ? val : 0.f) + b;
            }
          }
        }
      }
      FC_SYNCWARP();
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ KB
struct fc_fused_args {
  const float2* xin;    // [(b*Cin + c)][R][n_in]   output of the previous forward pass
  const float2* kspec;  // [(o*Ig + i)][R][N]       cached kernel spectrum
  float2* yout;         // [(b*Cout + o)][R][n_out] input of the next inverse pass
  const float2* tw;
  int32_t tw_len;
  int32_t B, Cin, Cout, G, Ig, Og;
  int32_t n_in, n_out, nbs;
  int64_t R;
  int64_t n_units;
  fc_imap imap;
  fc_omap omap;
};

// N: transform length of the fused axis. CI: bound on channels per group (in and out). NB: batches per CTA.
// Shared memory: NB*CI lines of N float2 (each line doubles as its warp's exchange buffer).
template <int N, int CI, int NB>
__global__ void __launch_bounds__(FC_FAST_WARPS * 32, (N * NB <= 1024) ? 2 : 1) fc_fused_axis_kernel(fc_fused_args a) {
  constexpr int E = N / 32;
  FC_DYN_SMEM(xy);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int Ig = a.Ig, Og = a.Og;
  fc_wofs ofs;
  ofs.init(lane);
  const fc_omap om = a.omap;
  const bool plain_in = a.imap.mode == FC_PAD_CONSTANT && a.imap.pad == 0 && a.imap.up == 1 && a.imap.sub == 1;
  const int in_lim = a.imap.ext < a.imap.L ? a.imap.ext : a.imap.L;
  const bool plain_out = om.og == 1 && om.os == 1 && om.ob == 0;
  const int out_lim = om.Lout < om.lim ? om.Lout : om.lim;
  for (int64_t unit = blockIdx.x; unit < a.n_units; unit += gridDim.x) {
    const int bs = (int)(unit % a.nbs);
    const int64_t gr = unit / a.nbs;
    const int64_t r = gr % a.R;
    const int g = (int)(gr / a.R);
    const int b0 = bs * NB;
    // ---- phase 1: forward transform of every (batch, input channel) line of this bin
    const int n_it1 = (NB * Ig + FC_FAST_WARPS - 1) / FC_FAST_WARPS;
#pragma unroll 1
    for (int it = 0; it < n_it1; ++it) {
      const int ln = w + FC_FAST_WARPS * it;
      if (ln >= NB * Ig) continue;  // warp-uniform
      const int bl = ln / Ig;
      const int i = ln - bl * Ig;
      const bool active = b0 + bl < a.B;
      float2* line = xy + (size_t)(bl * CI + i) * N;
      const float2* src = a.xin + (((int64_t)(b0 + (active ? bl : 0)) * a.Cin + g * Ig + i) * a.R + r) * a.n_in;
      float2 v[E];
      if (plain_in) {
#pragma unroll
        for (int q = 0; q < E; ++q) {
          const int n = lane + 32 * q;
          v[q] = (active && n < in_lim) ? __ldg(src + n) : make_float2(0.f, 0.f);
        }
      } else {
#pragma unroll
        for (int q = 0; q < E; ++q) {
          const int s = fc_imap_src(a.imap, lane + 32 * q);
          v[q] = (active && s >= 0) ? __ldg(src + s) : make_float2(0.f, 0.f);
        }
      }
      fc_wfft<N>::run(v, line, ofs, a.tw, a.tw_len, lane);  // the line itself is the warp's exchange buffer
#pragma unroll
      for (int q = 0; q < E; ++q) line[lane + 32 * q] = v[q];
    }
    __syncthreads();
    // ---- phase 2: per-bin contraction over the input channels of the group, in place (X -> Y).
    // One bin per thread and iteration; the kernel-spectrum loads of output channel o+1 are in flight while
    // output channel o is accumulated.
    for (int u = tid; u < N; u += FC_FAST_WARPS * 32) {
      float2 xr[NB][CI];
#pragma unroll
      for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int i = 0; i < CI; ++i) xr[b][i] = xy[(size_t)(b * CI + i) * N + u];
      const int64_t kstride = a.R * N;  // between input channels
      const float2* kl = a.kspec + (((int64_t)(g * Og) * Ig) * a.R + r) * N + u;
      float2 kv[CI], kn[CI];
#pragma unroll
      for (int i = 0; i < CI; ++i) kv[i] = (i < Ig) ? __ldg(kl + (int64_t)i * kstride) : make_float2(0.f, 0.f);
      for (int o = 0; o < Og; ++o) {
        const float2* kl1 = kl + (int64_t)((o + 1 < Og) ? (o + 1) : o) * Ig * kstride;
#pragma unroll
        for (int i = 0; i < CI; ++i) kn[i] = (i < Ig) ? __ldg(kl1 + (int64_t)i * kstride) : make_float2(0.f, 0.f);
        float2 acc[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[b] = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < CI; ++i) {
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            acc[b].x += xr[b][i].x * kv[i].x - xr[b][i].y * kv[i].y;
            acc[b].y += xr[b][i].x * kv[i].y + xr[b][i].y * kv[i].x;
          }
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) xy[(size_t)(b * CI + o) * N + u] = acc[b];
#pragma unroll
        for (int i = 0; i < CI; ++i) kv[i] = kn[i];
      }
    }
    __syncthreads();
    // ---- phase 3: inverse transform of every (batch, output channel) line, crop / stride on store
    const int n_it3 = (NB * Og + FC_FAST_WARPS - 1) / FC_FAST_WARPS;
#pragma unroll 1
    for (int it = 0; it < n_it3; ++it) {
      const int ln = w + FC_FAST_WARPS * it;
      if (ln >= NB * Og) continue;  // warp-uniform
      const int bl = ln / Og;
      const int o = ln - bl * Og;
      const bool active = b0 + bl < a.B;
      float2* line = xy + (size_t)(bl * CI + o) * N;
      float2 v[E];
#pragma unroll
      for (int q = 0; q < E; ++q) v[q] = fc_conj(line[lane + 32 * q]);
      FC_SYNCWARP();  // the line becomes the exchange buffer: every lane must have read its inputs
      fc_wfft<N>::run(v, line, ofs, a.tw, a.tw_len, lane);
      float2* dst = a.yout + (((int64_t)(b0 + (active ? bl : 0)) * a.Cout + g * Og + o) * a.R + r) * a.n_out;
      if (plain_out) {
        if (active) {
#pragma unroll
          for (int q = 0; q < E; ++q) {
            const int n = lane + 32 * q;
            if (n < om.Lout) dst[n] = (n < out_lim) ? fc_conj(v[q]) : make_float2(0.f, 0.f);
          }
        }
      } else {
        // general crop / stride / lattice map: stage the line in shared memory and scatter from there
#pragma unroll
        for (int q = 0; q < E; ++q) line[lane + 32 * q] = fc_conj(v[q]);
        FC_SYNCWARP();
        if (active) {
          for (int n = lane; n < N; n += 32) {
            const float2 val = line[n];
            for (int e = 0; e < om.og; ++e) {
              const int tt = n * om.og + e - om.ob;
              if (tt < 0 || (tt % om.os)) continue;
              const int j = tt / om.os;
              if (j >= om.Lout) continue;
              dst[j] = (e == 0 && n < om.lim) ? val : make_float2(0.f, 0.f);
            }
          }
        }
      }
    }
    __syncthreads();
  }
}

// fc_pair.cuh — the "pair" pipeline of the fused 2-d program: K1p -> KBp -> K4p on packed fp32 pairs.
//
// Blackwell issues two fp32 operations per lane with one FADD2 / FMUL2 / FFMA2 (PTX add/mul/fma.rn.f32x2) when both
// operands sit in aligned 64-bit register pairs, and FFMA2 / FMUL2 take a *scalar* second operand that is broadcast
// to both halves. The kernels here are laid out around that: the two batch items 2p and 2p + 1 of one channel form a
// "pair line", every complex value is held as (re of item 0, re of item 1), (im of item 0, im of item 1), and
// everything the two items share — twiddle factors, the kernel spectrum — enters as the broadcast scalar operand.
// A radix-8 butterfly of two lines is then 52 packed instructions instead of 2 x 32, a twiddle multiply 4 instead of
// 2 x 4, and the per-bin channel contraction 2 FFMA2 per complex multiply-accumulate instead of 4 FFMA, with no
// register shuffling anywhere: spectra live in HBM and in shared memory as 16-byte slots {re0, re1, im0, im1}, moved
// with 128-bit loads and stores.
//
//   fc_pair_r2c_kernel   K1p: rows of the two images of a pair -> packed half spectra, stored transposed
//                        ([pair image][bin][row] slots) so that the next axis is contiguous
//   fc_pair_fused_kernel KBp: per bin of the other axis: forward transform of every input-channel pair line,
//                        per-bin grouped contraction with the cached kernel spectrum (complex_matmul, reference
//                        functional.py:11-16), inverse transform of every output-channel pair line
//   fc_pair_c2r_kernel   K4p: transposed load -> C2R -> crop / stride / lattice + bias -> rows of the two images
//
// Replaces (reference functional.py) :60-62 / :126-139 (padding, zero-stuffing: gather maps), :70 / :157 rfftn of the
// signal, :73 / :160 complex_matmul, :75 / :162 irfftn, :76-87 / :163-174 crop, stride, bias.
#pragma once
#include <cstring>
#include <type_traits>

#include "fc_fused.cuh"

#ifdef FC_TUNING
#define FC_ABL(a, bit) (((a).abl & (bit)) != 0)
#else
#define FC_ABL(a, bit) false
#endif

// ------------------------------------------------------------------------------------------------ packed pairs
#ifdef FC_CPU_EMUL
struct fc_p2 {
  float x, y;
};
FC_DEV fc_p2 p2_make(float lo, float hi) { return fc_p2{lo, hi}; }
FC_DEV float p2_lo(fc_p2 a) { return a.x; }
FC_DEV float p2_hi(fc_p2 a) { return a.y; }
FC_DEV fc_p2 p2_add(fc_p2 a, fc_p2 b) { return fc_p2{a.x + b.x, a.y + b.y}; }
FC_DEV fc_p2 p2_sub(fc_p2 a, fc_p2 b) { return fc_p2{a.x - b.x, a.y - b.y}; }
FC_DEV fc_p2 p2_mul(fc_p2 a, fc_p2 b) { return fc_p2{a.x * b.x, a.y * b.y}; }
FC_DEV fc_p2 p2_fma(fc_p2 a, fc_p2 b, fc_p2 c) { return fc_p2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
#else
typedef unsigned long long fc_p2;
FC_DEV fc_p2 p2_make(float lo, float hi) {
  fc_p2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
FC_DEV float p2_lo(fc_p2 a) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a));
  return lo;
}
FC_DEV float p2_hi(fc_p2 a) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a));
  return hi;
}
FC_DEV fc_p2 p2_add(fc_p2 a, fc_p2 b) {
  fc_p2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
FC_DEV fc_p2 p2_sub(fc_p2 a, fc_p2 b) {
  fc_p2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
FC_DEV fc_p2 p2_mul(fc_p2 a, fc_p2 b) {
  fc_p2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
FC_DEV fc_p2 p2_fma(fc_p2 a, fc_p2 b, fc_p2 c) {
  fc_p2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
#endif
// Scalar operand shared by both halves: ptxas folds the duplicate into the broadcast (".F32") operand form of
// FMUL2 / FFMA2, negation included, so neither costs an instruction.
FC_DEV fc_p2 p2_dup(float s) { return p2_make(s, s); }
FC_DEV fc_p2 p2_muls(fc_p2 a, float s) { return p2_mul(a, p2_dup(s)); }
FC_DEV fc_p2 p2_fmas(fc_p2 a, float s, fc_p2 c) { return p2_fma(a, p2_dup(s), c); }

// Two complex numbers (one per item of the pair): a 16-byte slot {re0, re1, im0, im1}.
struct alignas(16) fc_c2 {
  fc_p2 re, im;
};
FC_DEV fc_c2 c2_make(fc_p2 re, fc_p2 im) {
  fc_c2 r;
  r.re = re;
  r.im = im;
  return r;
}
FC_DEV fc_c2 c2_zero() { return c2_make(p2_make(0.f, 0.f), p2_make(0.f, 0.f)); }
FC_DEV fc_c2 c2_add(fc_c2 a, fc_c2 b) { return c2_make(p2_add(a.re, b.re), p2_add(a.im, b.im)); }
FC_DEV fc_c2 c2_sub(fc_c2 a, fc_c2 b) { return c2_make(p2_sub(a.re, b.re), p2_sub(a.im, b.im)); }
// Swapping the parts turns the forward transform into the (unnormalised) inverse: ifft(z) = swap(fft(swap(z))). With
// separate registers for the parts it is a renaming, not an instruction.
FC_DEV fc_c2 c2_swap(fc_c2 a) { return c2_make(a.im, a.re); }
// a + b*(-i), a - b*(-i)
FC_DEV fc_c2 c2_add_mi(fc_c2 a, fc_c2 b) { return c2_make(p2_add(a.re, b.im), p2_sub(a.im, b.re)); }
FC_DEV fc_c2 c2_sub_mi(fc_c2 a, fc_c2 b) { return c2_make(p2_sub(a.re, b.im), p2_add(a.im, b.re)); }
// a * w, w shared by the two items
FC_DEV fc_c2 c2_muls(fc_c2 a, float2 w) {
  return c2_make(p2_fmas(a.re, w.x, p2_muls(a.im, -w.y)), p2_fmas(a.re, w.y, p2_muls(a.im, w.x)));
}

#ifdef FC_CPU_EMUL
FC_DEV fc_c2 c2_ldg(const fc_c2* p) { return *p; }
FC_DEV fc_c2 c2_ld_stream(const fc_c2* p) { return *p; }
FC_DEV void c2_st_stream(fc_c2* p, fc_c2 v) { *p = v; }
#else
FC_DEV fc_c2 c2_ldg(const fc_c2* p) {
  const ulonglong2 t = __ldg(reinterpret_cast<const ulonglong2*>(p));
  return c2_make(t.x, t.y);
}
FC_DEV fc_c2 c2_ld_stream(const fc_c2* p) {
  const ulonglong2 t = __ldcs(reinterpret_cast<const ulonglong2*>(p));
  return c2_make(t.x, t.y);
}
FC_DEV void c2_st_stream(fc_c2* p, fc_c2 v) { __stcs(reinterpret_cast<ulonglong2*>(p), make_ulonglong2(v.re, v.im)); }
#endif

// Two adjacent slots (32 bytes = one sector, 32-byte aligned) with one 256-bit access (sm_100: LDG/STG.E.256).
#ifdef FC_CPU_EMUL
FC_DEV void c2x2_st(fc_c2* p, fc_c2 a, fc_c2 b) {
  p[0] = a;
  p[1] = b;
}
FC_DEV void c2x2_ld(const fc_c2* p, fc_c2& a, fc_c2& b) {
  a = p[0];
  b = p[1];
}
#else
FC_DEV void c2x2_st(fc_c2* p, fc_c2 a, fc_c2 b) {
  asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(a.re), "l"(a.im), "l"(b.re), "l"(b.im) : "memory");
}
FC_DEV void c2x2_ld(const fc_c2* p, fc_c2& a, fc_c2& b) {
  asm volatile("ld.global.nc.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(a.re), "=l"(a.im), "=l"(b.re), "=l"(b.im) : "l"(p));
}
#endif

#ifdef FC_CPU_EMUL
FC_DEV void fc_prefetch_l1(const void*) {}
#else
// Pull one 128-byte line into L1 (CCTL.E.PF1): no register is tied up and nothing waits for it.
FC_DEV void fc_prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#endif

// XOR swizzle of the slot index inside a pair line. 16-byte accesses are served a quarter warp at a time, so an access
// is conflict-free when its 8 lanes touch 8 different slots mod 8. That holds for every pattern of the Stockham
// stages below: 8 consecutive slots (the lane + G*q reads, the writes of the stages with Ns >= 8) and the stride-8
// writes of the first stage (8j + r -> 8j + (r ^ (j & 7))).
FC_DEV int fc_swz16(int p) { return p ^ ((p >> 3) & 7); }

// ------------------------------------------------------------------------------------------------ butterflies
template <int R>
FC_DEV void fc_pbutterfly(fc_c2* v);

template <>
FC_DEV void fc_pbutterfly<2>(fc_c2* v) {
  const fc_c2 a = v[0], b = v[1];
  v[0] = c2_add(a, b);
  v[1] = c2_sub(a, b);
}

template <>
FC_DEV void fc_pbutterfly<4>(fc_c2* v) {
  const fc_c2 b0 = c2_add(v[0], v[2]), b2 = c2_sub(v[0], v[2]);
  const fc_c2 b1 = c2_add(v[1], v[3]), d = c2_sub(v[1], v[3]);
  v[0] = c2_add(b0, b1);
  v[2] = c2_sub(b0, b1);
  v[1] = c2_add_mi(b2, d);
  v[3] = c2_sub_mi(b2, d);
}

// Radix 8 (forward, sign -1). The two 45-degree rotations are left unscaled and their factor 1/sqrt2 is applied by the
// FFMA2 of the last level: 52 packed instructions for two lines.
template <>
FC_DEV void fc_pbutterfly<8>(fc_c2* v) {
  const float h = 0.70710678118654752440f;
  const fc_c2 a0 = c2_add(v[0], v[4]), a4 = c2_sub(v[0], v[4]);
  const fc_c2 a1 = c2_add(v[1], v[5]), a5 = c2_sub(v[1], v[5]);
  const fc_c2 a2 = c2_add(v[2], v[6]), a6 = c2_sub(v[2], v[6]);
  const fc_c2 a3 = c2_add(v[3], v[7]), a7 = c2_sub(v[3], v[7]);
  const fc_c2 a5u = c2_make(p2_add(a5.re, a5.im), p2_sub(a5.im, a5.re));  // a5 * (1 - i)    [= sqrt2 * a5 * W8]
  const fc_c2 a7n = c2_make(p2_sub(a7.re, a7.im), p2_add(a7.re, a7.im));  // a7 * (1 + i)    [= -sqrt2 * a7 * W8^3]
  const fc_c2 b0 = c2_add(a0, a2), b2 = c2_sub(a0, a2);
  const fc_c2 b1 = c2_add(a1, a3), d13 = c2_sub(a1, a3);
  const fc_c2 b4 = c2_add_mi(a4, a6), b6 = c2_sub_mi(a4, a6);
  const fc_c2 b5u = c2_sub(a5u, a7n);  // sqrt2 * (a5 W8 + a7 W8^3)
  const fc_c2 s = c2_add(a5u, a7n);    // sqrt2 * (a5 W8 - a7 W8^3); b7 = s * (-i) / sqrt2
  v[0] = c2_add(b0, b1);
  v[4] = c2_sub(b0, b1);
  v[2] = c2_add_mi(b2, d13);
  v[6] = c2_sub_mi(b2, d13);
  v[1] = c2_make(p2_fmas(b5u.re, h, b4.re), p2_fmas(b5u.im, h, b4.im));
  v[5] = c2_make(p2_fmas(b5u.re, -h, b4.re), p2_fmas(b5u.im, -h, b4.im));
  v[3] = c2_make(p2_fmas(s.im, h, b6.re), p2_fmas(s.re, -h, b6.im));
  v[7] = c2_make(p2_fmas(s.im, -h, b6.re), p2_fmas(s.re, h, b6.im));
}

// ------------------------------------------------------------------------------------------------ the pair engine
// A group of G lanes (G = 32 for M >= 256, else M/8) owns NLP pair lines of M points; lane gl of the group holds the
// points gl + G*q, q < E = M/G, of each. v[l][t + NBF*r] = input r of butterfly t of pair line l.
template <int M, int G, int NLP, int R, int Ns>
FC_DEV void fc_pstage(fc_c2 (&v)[NLP][M / G], const float2* tw, int tw_len, int gl) {
  constexpr int E = M / G, NBF = E / R;
  float2 w[R];
  if (Ns > 1 && Ns <= G) fc_twiddle_powers<R>(__ldg(tw + (gl & (Ns - 1)) * (tw_len / (Ns * R))), w);  // same for every t
#pragma unroll
  for (int t = 0; t < NBF; ++t) {
    if (Ns > G) fc_twiddle_powers<R>(__ldg(tw + ((gl + G * t) & (Ns - 1)) * (tw_len / (Ns * R))), w);
#pragma unroll
    for (int l = 0; l < NLP; ++l) {
      fc_c2 a[R];
#pragma unroll
      for (int r = 0; r < R; ++r) a[r] = v[l][t + NBF * r];
      if (Ns > 1) {
#pragma unroll
        for (int r = 1; r < R; ++r) a[r] = c2_muls(a[r], w[r]);
      }
      fc_pbutterfly<R>(a);
#pragma unroll
      for (int r = 0; r < R; ++r) v[l][t + NBF * r] = a[r];
    }
  }
}

// Synchronise the G threads that own a pair line: one warp (G <= 32: the whole warp runs the same code) or the G/32
// adjacent warps of named barrier `bar`.
template <int G>
FC_DEV void fc_group_sync(int bar) {
  if constexpr (G <= 32) {
    FC_SYNCWARP();
  } else {
    fc_named_bar_sync(bar, G);
  }
}

// Exchange after a stage: outputs go to their Stockham positions in the group's lines (swizzled slots), then every
// lane reads the gl + G*q layout back. Pair line l of the group lives at line0 + l*LS.
template <int M, int G, int NLP, int LS, int R, int Ns>
FC_DEV void fc_pxchg(fc_c2 (&v)[NLP][M / G], fc_c2* line0, int gl, int bar) {
  constexpr int E = M / G, NBF = E / R;
#pragma unroll
  for (int t = 0; t < NBF; ++t) {
    const int j = gl + G * t;
    const int k = j & (Ns - 1);
    const int j0 = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int s = fc_swz16(j0 + r * Ns);
#pragma unroll
      for (int l = 0; l < NLP; ++l) line0[l * LS + s] = v[l][t + NBF * r];
    }
  }
  fc_group_sync<G>(bar);
#pragma unroll
  for (int q = 0; q < E; ++q) {
    const int s = fc_swz16(gl + G * q);
#pragma unroll
    for (int l = 0; l < NLP; ++l) v[l][q] = line0[l * LS + s];
  }
  fc_group_sync<G>(bar);
}

// Forward, unnormalised FFT of NLP pair lines of M points; on return v[l][q] is bin gl + G*q. Every lane of the warp
// must call it (warp-wide __syncwarp); the lines are clobbered. G = 64: two adjacent warps share a line (8 points per
// lane for M = 512: half the registers and half the dependent work per thread) and meet at named barrier `bar`.
template <int M, int G, int NLP, int LS>
FC_DEV void fc_pfft(fc_c2 (&v)[NLP][M / G], fc_c2* line0, const float2* tw, int tw_len, int gl, int bar = 0) {
  static_assert(M == 64 || M == 128 || M == 256 || M == 512 || M == 1024 || M == 2048, "unsupported pair FFT length");
  static_assert(G == (M >= 256 ? 32 : M / 8) || (G == 64 && M >= 512), "lanes per pair line");
  fc_pstage<M, G, NLP, 8, 1>(v, tw, tw_len, gl);
  fc_pxchg<M, G, NLP, LS, 8, 1>(v, line0, gl, bar);
  fc_pstage<M, G, NLP, 8, 8>(v, tw, tw_len, gl);
  if constexpr (M == 64) return;  // 8 x 8: a single exchange
  fc_pxchg<M, G, NLP, LS, 8, 8>(v, line0, gl, bar);
  if constexpr (M == 128) {
    fc_pstage<M, G, NLP, 2, 64>(v, tw, tw_len, gl);
  } else if constexpr (M == 256) {
    fc_pstage<M, G, NLP, 4, 64>(v, tw, tw_len, gl);
  } else if constexpr (M == 512) {
    fc_pstage<M, G, NLP, 8, 64>(v, tw, tw_len, gl);
  } else {
    fc_pstage<M, G, NLP, 8, 64>(v, tw, tw_len, gl);
    fc_pxchg<M, G, NLP, LS, 8, 64>(v, line0, gl, bar);
    fc_pstage<M, G, NLP, (M == 2048 ? 4 : 2), 512>(v, tw, tw_len, gl);
  }
}

// Pair image o = p*C + c of a (B, C, ...) tensor: the images (2p, c) and (2p + 1, c); the second one is missing for
// the last pair of an odd batch.
struct fc_pair_img {
  int64_t i0;  // image index of item 0: (2p)*C + c; item 1 is C images further
  bool has1;
};
FC_DEV fc_pair_img fc_pair_image(int o, int B, int C) {
  const int p = o / C, c = o - p * C;
  fc_pair_img r;
  r.i0 = (int64_t)(2 * p) * C + c;
  r.has1 = 2 * p + 1 < B;
  return r;
}

// ------------------------------------------------------------------------------------------------ K1p
struct fc_pair_r2c_args {
  fc_pass p;  // the pass of the fast R2C kernel, retiled over pair images: n_outer = ceil(B/2)*C, T pair lines per tile
  const float* x;
  fc_c2* out;  // [pair image][bin (+ segment*(M+1))][row] slots
  const float2* tw;
  int32_t B, C;
  int32_t abl;  // FC_TUNING builds only: 1 skip the transform, 2 skip the row loads, 4 skip the spectrum stores
};

// Shared memory: TR pair lines of pitch M + 1 slots (odd: the transposed sweep of the store phase is conflict-free).
// A line is its group's exchange buffer during the transform, then holds the untangled half spectrum (bin k at the
// swizzled slot of k, the Nyquist bin in the extra slot M).
// YS > 0 ("y stage"): the first radix-YS stage of the transform of the other (y) axis runs in this kernel's transposed
// store, where every value passes through a thread anyway. With N_y = 64*YS and y = 64*n1 + n2, ky = k1 + YS*k2:
//   X[ky] = sum_n2 W64^(n2 k2) * { W_Ny^(n2 k1) * sum_n1 W_YS^(n1 k1) x[64 n1 + n2] },
// so a tile holds the TR = 16 pair lines {n2 + 64 n1 : n1 < YS} of 16/YS adjacent n2 (tile line l = n1*(16/YS) + dn2), a
// thread takes one bin kx, runs the YS-point butterfly over n1 and the twiddle, and stores
// out[pair image][kx][k1][n2]: the fused kernel (fc_pair_fused64_kernel) is left with independent 64-point
// transforms, one per k1, whose 64 bins only need 1/YS of the kernel spectrum of the line.
template <int M, int NLP, int NW, int OCC, int YS = 0>
__global__ void __launch_bounds__(NW * 32, OCC) fc_pair_r2c_kernel(fc_pair_r2c_args a) {
  fc_grid_dep_sync();
  constexpr int G = M >= 256 ? 32 : M / 8;
  constexpr int E = M / G, GPW = 32 / G;
  constexpr int TR = NLP * NW * GPW, KS = NW * 32 / TR;  // TR pair lines per tile; KS bins per store sweep
  static_assert(TR >= 8 && (TR & (TR - 1)) == 0 && TR <= NW * 32, "a tile is a power-of-two number of pair lines");
  static_assert(YS == 0 || (TR == 16 && (YS == 4 || YS == 8)), "y stage: 16-line tiles, radix 4 or 8");
  constexpr int NQ = YS > 0 ? TR / YS : 1;  // adjacent n2 of a y-stage tile
  // line pitch: odd for the transposed sweep; y stage: the NQ lines (n1, dn) a quarter warp reads next to each other must
  // start 8/NQ slots apart mod 8
  constexpr int LP = YS == 0 ? M + 1 : (NQ == 2 ? M + 4 : M + 2);
  const fc_pass& p = a.p;
  FC_DYN_SMEM(smem_raw);
  fc_c2* smem = reinterpret_cast<fc_c2*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int gl = lane % G, gid = lane / G;
  const int lrow = (w * GPW + gid) * NLP;  // first tile line of this lane's group
  fc_c2* line0 = smem + lrow * LP;
  const int L = p.imap.L;
  const int tstep = p.tw_len / (2 * M);
  const int tpo = (int)p.tiles_per_outer, n_tiles = (int)p.n_tiles, R = (int)p.R;
  const int tps = tpo / p.seg_n;  // overlap-save segments: the tiles of an outer item run segment-major
  auto load_rows = [&](int t, fc_c2 (&v)[NLP][E]) {
    const int o = t / tpo;
    const int rem = t - o * tpo;
    const int sg = p.seg_n > 1 ? rem / tps : 0;
    const int r0 = (rem - sg * tps) * TR;
    const int ub = sg * p.seg_V - p.seg_off - p.imap.pad;  // source index of the segment's first dense position (even)
    const fc_pair_img pi = fc_pair_image(o, a.B, a.C);
    const float* img0 = a.x + pi.i0 * p.o_sA;
    const float* img1 = img0 + (int64_t)a.C * p.o_sA;
#pragma unroll
    for (int l = 0; l < NLP; ++l) {
      // y stage: tile t2 of an image holds the rows n2 + 64*n1 with n2 = t2*NQ + (line % NQ), n1 = line / NQ
      const int r = YS > 0 ? (r0 / TR) * NQ + ((lrow + l) % NQ) + p.ystage_S * ((lrow + l) / NQ) : r0 + lrow + l;
      const bool valid = r < R;
      const int64_t roff = (int64_t)(valid ? r : 0) * p.in_rs + ub;
      const float2* row0 = reinterpret_cast<const float2*>(img0 + roff) + gl;
      const float2* row1 = reinterpret_cast<const float2*>(img1 + roff) + gl;
#pragma unroll
      for (int q = 0; q < E; ++q) {  // L, the zero padding and ub are even (host check): the pair (2m, 2m + 1) is in or out together
        const bool in = valid && (unsigned)(ub + 2 * (gl + G * q)) < (unsigned)L && !FC_ABL(a, 2);
        const float2 x0 = in ? __ldg(row0 + G * q) : make_float2(0.f, 0.f);
        const float2 x1 = (in && pi.has1) ? __ldg(row1 + G * q) : make_float2(0.f, 0.f);
        v[l][q] = c2_make(p2_make(x0.x, x1.x), p2_make(x0.y, x1.y));
      }
    }
  };
  fc_c2 v[NLP][E];
  if ((int)blockIdx.x < n_tiles) load_rows(blockIdx.x, v);
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int o = t / tpo;
    const int rem = t - o * tpo;
    const int sg = p.seg_n > 1 ? rem / tps : 0;
    const int r0 = (rem - sg * tps) * TR;
    const int tn = t + gridDim.x;
    if (!FC_ABL(a, 1)) fc_pfft<M, G, NLP, LP>(v, line0, a.tw, p.tw_len, gl);
#pragma unroll
    for (int q = 0; q < E; ++q) {
      const int s = fc_swz16(gl + G * q);
#pragma unroll
      for (int l = 0; l < NLP; ++l) line0[l * LP + s] = v[l][q];
    }
    FC_SYNCWARP();
    // untangle the packed real transforms: X[k] = (z[k] + conj z[M-k])/2 + w_k * (-i)(z[k] - conj z[M-k])/2
    fc_p2 nyq[NLP];
#pragma unroll
    for (int l = 0; l < NLP; ++l) {
      const fc_c2 z0 = line0[l * LP];  // fc_swz16(0) == 0
      nyq[l] = p2_sub(z0.re, z0.im);   // bin k = M: E[0] - O[0]
    }
#pragma unroll
    for (int q = 0; q < E; ++q) {
      const int k = gl + G * q;
      const float2 wk = __ldg(a.tw + k * tstep);
      const float hx = 0.5f * wk.x, hy = 0.5f * wk.y;
      const int km = fc_swz16((M - k) & (M - 1));
#pragma unroll
      for (int l = 0; l < NLP; ++l) {
        const fc_c2 zk = v[l][q];
        const fc_c2 zm = line0[l * LP + km];
        const fc_p2 s_re = p2_add(zk.re, zm.re), s_im = p2_sub(zk.im, zm.im);
        const fc_p2 d_re = p2_sub(zk.re, zm.re), d_im = p2_add(zk.im, zm.im);
        v[l][q] = c2_make(p2_fmas(s_re, 0.5f, p2_fmas(d_im, hx, p2_muls(d_re, hy))),
                          p2_fmas(s_im, 0.5f, p2_fmas(d_re, -hx, p2_muls(d_im, hy))));
      }
    }
    FC_SYNCWARP();  // every partner has been read: the lines can take the spectrum
#pragma unroll
    for (int q = 0; q < E; ++q) {
      const int s = fc_swz16(gl + G * q);
#pragma unroll
      for (int l = 0; l < NLP; ++l) line0[l * LP + s] = v[l][q];
    }
    if (gl == 0) {
#pragma unroll
      for (int l = 0; l < NLP; ++l) line0[l * LP + M] = c2_make(nyq[l], p2_make(0.f, 0.f));
    }
    __syncthreads();
    if (tn < n_tiles) {
      load_rows(tn, v);  // in flight during the store below
      const int t2 = tn + gridDim.x;  // and pull the tile after that one into L2 (the segments of a row re-read it anyway)
      if (YS == 0 && t2 < n_tiles && p.seg_n == 1) {
        const int on = t2 / tpo;
        const int rn = (t2 - on * tpo) * TR;
        const int rows = (R - rn < TR) ? R - rn : TR;
        const fc_pair_img pn = fc_pair_image(on, a.B, a.C);
        const int span = rows * (int)p.in_rs;  // floats: the TR rows of an image are one contiguous run
        const float* nxt = a.x + pn.i0 * p.o_sA + (int64_t)rn * p.in_rs;
        for (int e = tid * 32; e < span; e += NW * 32 * 32) {
          fc_prefetch_l2(nxt + e);
          if (pn.has1) fc_prefetch_l2(nxt + (int64_t)a.C * p.o_sA + e);
        }
      }
    }
    if constexpr (YS > 0) {
      // y stage: thread <-> bin kx; per n2 of the tile: YS slots (n1) -> butterfly -> twiddle W_Ny^(n2 k1) -> out[kx][k1][n2]
      const int n2a = (r0 / TR) * NQ;
      const int S = p.ystage_S;              // sub-transform length: y = S*n1 + n2
      const int ts = p.tw_len / p.ystage_N;  // table step of W_Ny
      // task = (bin, n2): the NQ lanes of a bin are adjacent, so their 16-byte stores fill whole sectors
      for (int task = tid; task < (M + 1) * NQ && !FC_ABL(a, 4); task += NW * 32) {
        const int k = task / NQ, dn = task % NQ;
        const int sk = k < M ? fc_swz16(k) : M;
        fc_c2* dst = a.out + (int64_t)o * p.out_os + (int64_t)(k + sg * (M + 1)) * p.out_es + n2a + dn;
        const float2 w1 = __ldg(a.tw + (n2a + dn) * ts);  // W_Ny^n2; the twiddles W_Ny^(n2 k1) are built up from it
        const fc_c2* col = smem + dn * LP + sk;            // input n1 at col[n1 * NQ * LP]
        if constexpr (YS == 8) {
          // split radix, 4 outputs at a time (the rows of the next tile are in flight in the registers the transform
          // released: this phase has to live in what is left): even k1 from u[n1] + u[n1+4], odd k1 from
          // (u[n1] - u[n1+4]) W8^n1, each half a radix-4 butterfly
          const float h = 0.70710678118654752440f;
          const float2 w2 = fc_mul(w1, w1);
          {
            fc_c2 e[4];
#pragma unroll
            for (int n1 = 0; n1 < 4; ++n1) e[n1] = c2_add(col[n1 * NQ * LP], col[(n1 + 4) * NQ * LP]);
            fc_pbutterfly<4>(e);
            float2 wk = w2;
            dst[0] = e[0];
#pragma unroll
            for (int m = 1; m < 4; ++m) {
              dst[2 * m * S] = c2_muls(e[m], wk);
              wk = fc_mul(wk, w2);
            }
          }
          {
            const fc_c2 d0 = c2_sub(col[0], col[4 * NQ * LP]), d1 = c2_sub(col[1 * NQ * LP], col[5 * NQ * LP]);
            const fc_c2 d2 = c2_sub(col[2 * NQ * LP], col[6 * NQ * LP]), d3 = c2_sub(col[3 * NQ * LP], col[7 * NQ * LP]);
            const fc_c2 r1 = c2_make(p2_add(d1.re, d1.im), p2_sub(d1.im, d1.re));   // sqrt2 * d1 * W8
            const fc_c2 r3n = c2_make(p2_sub(d3.re, d3.im), p2_add(d3.re, d3.im));  // -sqrt2 * d3 * W8^3
            const fc_c2 t0 = c2_add_mi(d0, d2), t2 = c2_sub_mi(d0, d2);
            const fc_c2 sm = c2_sub(r1, r3n), q = c2_add(r1, r3n);
            fc_c2 od[4];
            od[0] = c2_make(p2_fmas(sm.re, h, t0.re), p2_fmas(sm.im, h, t0.im));
            od[2] = c2_make(p2_fmas(sm.re, -h, t0.re), p2_fmas(sm.im, -h, t0.im));
            od[1] = c2_make(p2_fmas(q.im, h, t2.re), p2_fmas(q.re, -h, t2.im));
            od[3] = c2_make(p2_fmas(q.im, -h, t2.re), p2_fmas(q.re, h, t2.im));
            float2 wk = w1;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              dst[(2 * m + 1) * S] = c2_muls(od[m], wk);
              wk = fc_mul(wk, w2);
            }
          }
        } else {
          fc_c2 u[YS];
#pragma unroll
          for (int n1 = 0; n1 < YS; ++n1) u[n1] = col[n1 * NQ * LP];
          fc_pbutterfly<YS>(u);
          float2 wk = w1;
          dst[0] = u[0];
#pragma unroll
          for (int k1 = 1; k1 < YS; ++k1) {
            dst[k1 * S] = c2_muls(u[k1], wk);
            wk = fc_mul(wk, w1);
          }
        }
      }
    } else {  // transposed store: thread (l = tid % TR, k = tid / TR + KS*it) writes TR consecutive rows of one bin (TR*16 bytes)
      const int l = tid & (TR - 1);
      if (r0 + l < R && !FC_ABL(a, 4)) {
        const int k0 = tid / TR;
        fc_c2* dst = a.out + (int64_t)o * p.out_os + r0 + l + (int64_t)(k0 + sg * (M + 1)) * p.out_es;
        const fc_c2* src = smem + l * LP;
        const int64_t dstep = KS * p.out_es;
#pragma unroll 8
        for (int it = 0; it < M / KS; ++it) dst[it * dstep] = src[fc_swz16(k0 + KS * it)];
        if (k0 == 0) dst[(M / KS) * dstep] = src[M];  // the Nyquist bin
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ K4p
struct fc_pair_c2r_args {
  fc_pass p;  // the pass of the fast C2R kernel, retiled over pair images: n_outer = ceil(B/2)*Cout
  const fc_c2* in;  // [pair image][bin (+ segment*(M+1))][row] slots
  float* out;
  const float2* tw;
  const float* bias;
  int32_t B, C;  // batch, output channels
};

// YS > 0 ("y stage", see K1p): the last radix-YS stage of the inverse transform of the other axis runs in the transposed
// load: in[pair image][kx][k1][n2] holds the 64-point inverse transforms of the fused kernel, a thread takes one bin kx
// and one n2, multiplies by W_Ny^(-n2 k1), runs the inverse YS-point butterfly over k1 and puts row 64*n1 + n2 into tile
// line n1*(16/YS) + dn2.
template <int M, int NLP, int NW, int OCC, int YS = 0>
__global__ void __launch_bounds__(NW * 32, OCC) fc_pair_c2r_kernel(fc_pair_c2r_args a) {
  fc_grid_dep_sync();
  constexpr int G = M >= 256 ? 32 : M / 8;
  constexpr int E = M / G, GPW = 32 / G;
  constexpr int TR = NLP * NW * GPW, KS = NW * 32 / TR;
  static_assert(TR >= 8 && (TR & (TR - 1)) == 0 && TR <= NW * 32, "a tile is a power-of-two number of pair lines");
  static_assert(YS == 0 || (TR == 16 && (YS == 4 || YS == 8)), "y stage: 16-line tiles, radix 4 or 8");
  constexpr int NQ = YS > 0 ? TR / YS : 1;  // adjacent n2 of a y-stage tile
  constexpr int LP = YS == 0 ? M + 1 : (NQ == 2 ? M + 4 : M + 2);  // (see K1p)
  const fc_pass& p = a.p;
  FC_DYN_SMEM(smem_raw);
  fc_c2* smem = reinterpret_cast<fc_c2*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int gl = lane % G, gid = lane / G;
  const int lrow = (w * GPW + gid) * NLP;
  fc_c2* line0 = smem + lrow * LP;
  const int tstep = p.tw_len / (2 * M);
  const fc_omap om = p.omap;
  const bool plain_out = om.os == 1 && om.ob == 0 && om.og == 1 && !(om.Lout & 1) && !(p.out_rs & 1) && !(p.out_os & 1) && p.row_og == 1 &&
                         p.seg_n == 1;
  const int tpo = (int)p.tiles_per_outer, n_tiles = (int)p.n_tiles, R = (int)p.R;
  const int tps = tpo / p.seg_n;
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int o = t / tpo;
    const int rem = t - o * tpo;
    const int sg = p.seg_n > 1 ? rem / tps : 0;
    const int r0 = (rem - sg * tps) * TR;
    if constexpr (YS > 0) {
      const int n2a = (r0 / TR) * NQ;
      const int S = p.ystage_S;
      const int ts = p.tw_len / p.ystage_N;
      for (int task = tid; task < (M + 1) * NQ; task += NW * 32) {
        const int k = task / NQ, dn = task % NQ;
        const fc_c2* src = a.in + (int64_t)o * p.in_os + (int64_t)(k + sg * (M + 1)) * p.in_es + n2a + dn;
        fc_c2 u[YS];
#pragma unroll
        for (int k1 = 0; k1 < YS; ++k1) u[k1] = c2_ldg(src + k1 * S);
        // inverse butterfly = forward butterfly between two swaps of the parts; twiddle W_Ny^(-n2 k1)
        float2 wy[YS];
        fc_twiddle_powers<YS>(__ldg(a.tw + (n2a + dn) * ts), wy);
        u[0] = c2_swap(u[0]);
#pragma unroll
        for (int k1 = 1; k1 < YS; ++k1) u[k1] = c2_swap(c2_muls(u[k1], make_float2(wy[k1].x, -wy[k1].y)));
        fc_pbutterfly<YS>(u);
#pragma unroll
        for (int n1 = 0; n1 < YS; ++n1) smem[(n1 * NQ + dn) * LP + k] = c2_swap(u[n1]);
      }
    } else {  // transposed load: thread (l = tid % TR, k = tid / TR + KS*it) reads TR consecutive rows of one bin
      const int l = tid & (TR - 1);
      const bool ok = r0 + l < R;
      const int k0 = tid / TR;
      const fc_c2* src = a.in + (int64_t)o * p.in_os + r0 + (ok ? l : 0) + (int64_t)(k0 + sg * (M + 1)) * p.in_es;
      fc_c2* dst = smem + l * LP + k0;
      const int64_t sstep = KS * p.in_es;
      constexpr int NIT = (M + KS) / KS;  // ceil((M + 1) / KS)
      constexpr int CH = NIT < 8 ? NIT : 8;  // loads of a thread in flight before the first shared-memory store
#pragma unroll 1
      for (int i0 = 0; i0 < NIT; i0 += CH) {
        fc_c2 tbuf[CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          const int it = i0 + i;
          tbuf[i] = (ok && it < NIT && k0 + it * KS <= M) ? c2_ldg(src + it * sstep) : c2_zero();
        }
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          const int it = i0 + i;
          if (it < NIT && k0 + it * KS <= M) dst[it * KS] = tbuf[i];
        }
      }
    }
    {  // L2 prefetch of the next tile of this CTA: (M+1) segments of TR slots
      const int tn = t + gridDim.x;
      if (YS > 0 && tn < n_tiles) {  // y stage: the 32-byte pieces [kx][k1][n2 of the tile] this thread reads next
        const int on = tn / tpo;
        const int remn = tn - on * tpo;
        const int sn = remn / tps;
        const int n2n = (remn - sn * tps) * NQ;
        for (int task = tid * NQ; task < (M + 1) * NQ; task += NW * 32 * NQ) {
          const fc_c2* nxt = a.in + (int64_t)on * p.in_os + (int64_t)(task / NQ + sn * (M + 1)) * p.in_es + n2n;
#pragma unroll
          for (int k1 = 0; k1 < YS; ++k1) fc_prefetch_l2(nxt + k1 * p.ystage_S);
        }
      }
      if (YS == 0 && tn < n_tiles) {
        const int on = tn / tpo;
        const int remn = tn - on * tpo;
        const int sn = p.seg_n > 1 ? remn / tps : 0;
        const int rn = (remn - sn * tps) * TR;
        const fc_c2* nxt = a.in + (int64_t)on * p.in_os + rn + (int64_t)sn * (M + 1) * p.in_es;
        constexpr int PL = TR / 8;  // 128-byte lines per segment
        for (int k = tid; k < (M + 1) * PL; k += NW * 32) fc_prefetch_l2(nxt + (int64_t)(k / PL) * p.in_es + 8 * (k % PL));
      }
    }
    __syncthreads();
    const fc_pair_img pi = fc_pair_image(o, a.B, a.C);
    const float b = p.has_bias ? __ldg(a.bias + (o % a.C)) : 0.f;
    // y stage: the lines of the rows the crop drops (64*n1 + n2 >= R: one n1 in eight at BASELINE c2) skip the transform
    const bool line_live = !(YS > 0 && G == 32 && NLP == 1) || ((r0 / TR) * NQ + (lrow % NQ) + p.ystage_S * (lrow / NQ)) < R;
    if (line_live) {
    // Hermitian pre-twist: Z[k] = (Y[k] + conj Y[M-k]) + i * conj(w_k) * (Y[k] - conj Y[M-k]); the transform runs on
    // swap(Z) and the real samples come out as x[2m] = im, x[2m+1] = re of the result
    fc_c2 v[NLP][E];
#pragma unroll
    for (int q = 0; q < E; ++q) {
      const int k = gl + G * q;
      const float2 wk = __ldg(a.tw + k * tstep);
#pragma unroll
      for (int l = 0; l < NLP; ++l) {
        const fc_c2 yk = line0[l * LP + k];
        const fc_c2 ym = line0[l * LP + M - k];
        const fc_p2 s_re = p2_add(yk.re, ym.re), s_im = p2_sub(yk.im, ym.im);
        const fc_p2 t_re = p2_sub(yk.re, ym.re), t_im = p2_add(yk.im, ym.im);
        // d = t * conj(w_k): d.re = t.re*wx + t.im*wy, d.im = t.im*wx - t.re*wy;  Z = (s.re - d.im, s.im + d.re)
        const fc_p2 z_re = p2_fmas(t_re, wk.y, p2_fmas(t_im, -wk.x, s_re));
        const fc_p2 z_im = p2_fmas(t_im, wk.y, p2_fmas(t_re, wk.x, s_im));
        v[l][q] = c2_make(z_im, z_re);
      }
    }
    FC_SYNCWARP();  // the lines are exchange buffers from here on
    fc_pfft<M, G, NLP, LP>(v, line0, a.tw, p.tw_len, gl);
    if (plain_out) {
#pragma unroll
      for (int l = 0; l < NLP; ++l) {
        const int64_t r = YS > 0 ? (r0 / TR) * NQ + ((lrow + l) % NQ) + p.ystage_S * ((lrow + l) / NQ) : r0 + lrow + l;
        if (r < p.R) {
          float* y0 = a.out + pi.i0 * p.out_os + r * p.out_rs;
          float* y1 = y0 + (int64_t)a.C * p.out_os;
#pragma unroll
          for (int q = 0; q < E; ++q) {
            const int n0 = 2 * (gl + G * q);
            if (n0 < om.Lout) {
              const fc_c2 f = v[l][q];
              fc_st_stream(reinterpret_cast<float2*>(y0 + n0), make_float2(p2_lo(f.im) + b, p2_lo(f.re) + b));
              if (pi.has1) fc_st_stream(reinterpret_cast<float2*>(y1 + n0), make_float2(p2_hi(f.im) + b, p2_hi(f.re) + b));
            }
          }
        }
      }
    } else {
      // general crop / stride / lattice map: stage the real rows of the two items in the group's lines (item h at floats
      // [2M*h, 2M*h + 2M) of the pair line) and scatter from there
      FC_SYNCWARP();
#pragma unroll
      for (int l = 0; l < NLP; ++l) {
        float2* f2 = reinterpret_cast<float2*>(line0 + l * LP);
#pragma unroll
        for (int q = 0; q < E; ++q) {
          const fc_c2 f = v[l][q];
          f2[gl + G * q] = make_float2(p2_lo(f.im), p2_lo(f.re));
          f2[M + gl + G * q] = make_float2(p2_hi(f.im), p2_hi(f.re));
        }
      }
      FC_SYNCWARP();
      // this (row, segment) line owns the dense samples n in [n_lo, n_hi), i.e. the outputs j with
      // n(j) = (j*os + ob) / og in that range: a contiguous run of j because n(j) is monotone
      const int n_lo = sg * p.seg_V, n_hi = n_lo + p.seg_V;
      const int c_lo = n_lo * om.og - om.ob, c_hi = n_hi * om.og - om.ob;
      const int j_lo = c_lo > 0 ? (c_lo + om.os - 1) / om.os : 0;
      int j_hi = c_hi > 0 ? (c_hi + om.os - 1) / om.os : 0;
      if (j_hi > om.Lout) j_hi = om.Lout;
      for (int lh = 0; lh < 2 * NLP; ++lh) {
        const int l = lh >> 1, h = lh & 1;
        const int64_t r = YS > 0 ? (r0 / TR) * NQ + ((lrow + l) % NQ) + p.ystage_S * ((lrow + l) / NQ) : r0 + lrow + l;
        if (r >= p.R || (h && !pi.has1)) continue;
        const float* rl = reinterpret_cast<const float*>(line0 + l * LP) + 2 * M * h + (p.seg_off - n_lo);
        float* img = a.out + (pi.i0 + (int64_t)h * a.C) * p.out_os;
        for (int er = 0; er < p.row_og; ++er) {  // output rows owned by this dense line (one unless row lattice)
          const int64_t jr = r * p.row_og + er - p.row_ob;
          if (jr < 0 || jr >= p.row_Lout) continue;
          float* yrow = img + jr * p.out_rs;
          if (er != 0) {  // a row between the lattice rows: bias only
            fc_fill_run<G>(yrow, j_lo, j_hi, b, gl);
            continue;
          }
          if (om.og == 2 && om.os == 1 && om.ob >= 0) {
            // lattice of 2 (BASELINE c5: stride 2, dilation 2): of every aligned output pair exactly one is a dense
            // sample, the other is bias only -> one shared-memory read and one 8-byte store per two outputs
            const int mis = (int)((reinterpret_cast<uintptr_t>(yrow + j_lo) >> 2) & 1);
            const int ja = j_lo + mis < j_hi ? j_lo + mis : j_hi;
            const int npairs = (j_hi - ja) >> 1;
            const int par = (ja + om.ob) & 1;        // 0: the first output of a pair is the dense sample, 1: the second
            const int nb = (ja + om.ob + par) >> 1;  // dense index of the sample of pair 0
            float2* y2 = reinterpret_cast<float2*>(yrow + ja);
#pragma unroll 4
            for (int m = gl; m < npairs; m += G) {
              const int n = nb + m;
              const float val = (n < om.lim ? rl[n] : 0.f) + b;
              y2[m] = par ? make_float2(b, val) : make_float2(val, b);
            }
            if (gl < 2) {  // the unaligned first and the odd last output of the run
              const int j = gl == 0 ? j_lo : ja + 2 * npairs;
              if (gl == 0 ? (mis && j_lo < j_hi) : j < j_hi) {
                const int tt = j + om.ob, n = tt >> 1;
                yrow[j] = ((tt & 1) == 0 && n < om.lim ? rl[n] : 0.f) + b;
              }
            }
            continue;
          }
          // output-driven (coalesced stores): output j takes dense sample n = (j*os + ob) / og when the remainder
          // is 0 and n < lim, else it is bias only
#pragma unroll 8
          for (int j = j_lo + gl; j < j_hi; j += G) {
            const int tt = j * om.os + om.ob;
            const int n = om.og == 1 ? tt : om.og == 2 ? (tt >> 1) : tt / om.og;
            const bool live = tt == n * om.og && n < om.lim;
            yrow[j] = (live ? rl[n] : 0.f) + b;
          }
        }
      }
    }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ KBp
struct fc_pair_fused_args {
  const fc_c2* xin;     // [(p*Cin + c)][R][n_in] slots     output of K1p
  const float2* kspec;  // bin-major kernel spectrum [group][line][o][i][N] (fc_fused.cuh)
  fc_c2* yout;          // [(p*Cout + o)][R][n_out] slots   input of K4p
  const float2* tw;
  int32_t tw_len;
  int32_t BP, Cin, Cout, G;  // batch pairs = ceil(B/2)
  int32_t n_in, n_out, nbs;
  int32_t n_items;  // BP * n_seg: (batch pair, segment) items per (group, bin); NP of them per unit
  int32_t n_seg, seg_V, seg_off;
  int32_t prefetch_dist;
  int32_t desync_ns, desync_mod, desync_grp;  // start-up delay of CTA group g = blockIdx / desync_mod of the first wave (g < desync_grp CTAs per SM): g * desync_ns
  int32_t k_share;  // FC_TUNING builds only: item-major contraction threads (one item per thread, kernel spectrum shared through L1)
  int32_t abl;   // FC_TUNING builds only: phase ablation bits for timing experiments (results are wrong by construction)
  int32_t k_pf;  // output channels of kernel-spectrum lines the contraction keeps in flight towards L1 ahead of its loads (0: off)
  int64_t R, Rk, n_units;
  fc_imap imap;
  fc_omap omap;
};

// Per-bin contraction over the CI input channels of the group, in place (X -> Y) in the CTA's pair lines. A thread
// owns one bin of IPT pair items: the signal values (CI slots -> 4*CI registers per item) are loaded once, every kernel
// value is the broadcast operand of 4 FFMA2 per item (two complex multiply-accumulates), and the four partial sums per
// output channel are independent chains. The kernel values of two input channels of a bin are adjacent in the
// spectrum ([o][i/2][n][i%2], fc_pass::out_il): one 16-byte load per channel pair, so the loads the register file
// can keep in flight carry twice the bytes (this phase is bound by the latency of these loads). The loop over output
// channels is unrolled: every load is the thread's base pointer plus a compile-time offset and ptxas hoists the loads
// of the following channels above the FFMA2s of this one. IPT = 2 reads the kernel spectrum once for two pair items.
#ifndef FC_KPRE
#define FC_KPRE 0  // kernel-spectrum loads (16 bytes each) a contraction thread issues before the barrier that ends phase 1
#endif

// The first FC_KPRE kernel-spectrum values of the thread's first bin: requested before the phase barrier, so that their
// latency runs while the CTA waits for its slowest transform (the registers of the transform are free by then).
template <int N, int CI, int NP, int IPT, int W>
FC_DEV void fc_pair_contract_preload(float4 (&kpre)[FC_KPRE > 0 ? FC_KPRE : 1], const fc_pair_fused_args& a, int g, int rk, int tid) {
  if (FC_KPRE > 0 && tid < N * (NP / IPT)) {
    const float4* kp = reinterpret_cast<const float4*>(a.kspec + ((int64_t)g * a.Rk + rk) * ((int64_t)CI * CI * N)) + (tid & (N - 1));
#pragma unroll
    for (int j = 0; j < FC_KPRE; ++j) kpre[j] = FC_ABL(a, 2) ? make_float4(1.f, 0.5f, 0.25f, 2.f) : __ldg(kp + j * N);
  }
}

// ITEM_MAJOR (IPT = 1, NP > 1): the threads split into NP equal groups, one per pair item, that walk the same bins at the
// same time, so that the kernel-spectrum lines one group pulls from L2 are L1 hits for the others.
template <int N, int CI, int NP, int IPT, int W, bool ITEM_MAJOR = false>
FC_DEV void fc_pair_contract(fc_c2* xy, const fc_pair_fused_args& a, int g, int rk, int tid, const float4 (&kpre)[FC_KPRE > 0 ? FC_KPRE : 1]) {
  static_assert(NP % IPT == 0 && CI % 2 == 0, "items per thread divide the items of a CTA; channel pairs");
  static_assert(FC_KPRE <= CI * CI / 2, "preloaded values are the first ones of a bin");
  auto body = [&](int idx, auto first) {
    constexpr bool FIRST = decltype(first)::value;
    constexpr int TPI = W * 32 / NP;  // ITEM_MAJOR: threads per item
    const int n = ITEM_MAJOR ? (idx % TPI) + (idx / (W * 32)) * TPI : idx & (N - 1);
    const int pg = ITEM_MAJOR ? (idx % (W * 32)) / TPI : (idx / N) * IPT;
    fc_c2* xb = xy + (size_t)(pg * CI) * N + n;  // line (pg + t, c) at xb + ((t*CI + c)*N
    fc_c2 x[IPT][CI];
#pragma unroll
    for (int t = 0; t < IPT; ++t)
#pragma unroll
      for (int i = 0; i < CI; ++i) x[t][i] = xb[(size_t)(t * CI + i) * N];
    const float4* kp = reinterpret_cast<const float4*>(a.kspec + ((int64_t)g * a.Rk + rk) * ((int64_t)CI * CI * N)) + n;
#pragma unroll
    for (int o = 0; o < CI; ++o) {
      if (a.k_pf > 0) {
        // software prefetch: the lines of output channel o + k_pf (of this thread's next bin once the channels run out)
        const int on = o + a.k_pf;
        const float4* pp = on < CI ? kp + on * (CI / 2) * N : kp + (on - CI) * (CI / 2) * N + W * 32;
        if (on < CI || idx + W * 32 < N * (NP / IPT)) {
#pragma unroll
          for (int i2 = 0; i2 < CI / 2; ++i2) fc_prefetch_l1(pp + i2 * N);
        }
      }
      fc_p2 rr[IPT], ii[IPT], ri[IPT], ir[IPT];  // sum x.re*k.re, x.im*k.im, x.re*k.im, x.im*k.re
#pragma unroll
      for (int t = 0; t < IPT; ++t) rr[t] = ii[t] = ri[t] = ir[t] = p2_make(0.f, 0.f);
#pragma unroll
      for (int i2 = 0; i2 < CI / 2; ++i2) {
        constexpr int dummy = 0;
        (void)dummy;
        const int j = o * (CI / 2) + i2;
        float4 k;
        if (FIRST && j < FC_KPRE)
          k = kpre[j < FC_KPRE ? j : 0];
        else
          k = FC_ABL(a, 2) ? make_float4(1.f, 0.5f, 0.25f, 2.f) : __ldg(kp + j * N);
#pragma unroll
        for (int t = 0; t < IPT; ++t) {
          rr[t] = p2_fmas(x[t][2 * i2].re, k.x, rr[t]);
          ii[t] = p2_fmas(x[t][2 * i2].im, k.y, ii[t]);
          ri[t] = p2_fmas(x[t][2 * i2].re, k.y, ri[t]);
          ir[t] = p2_fmas(x[t][2 * i2].im, k.x, ir[t]);
          rr[t] = p2_fmas(x[t][2 * i2 + 1].re, k.z, rr[t]);
          ii[t] = p2_fmas(x[t][2 * i2 + 1].im, k.w, ii[t]);
          ri[t] = p2_fmas(x[t][2 * i2 + 1].re, k.w, ri[t]);
          ir[t] = p2_fmas(x[t][2 * i2 + 1].im, k.z, ir[t]);
        }
      }
#pragma unroll
      for (int t = 0; t < IPT; ++t) xb[(size_t)(t * CI + o) * N] = c2_make(p2_sub(rr[t], ii[t]), p2_add(ri[t], ir[t]));
    }
  };
  if (tid < N * (NP / IPT)) body(tid, std::true_type());
  for (int idx = tid + W * 32; idx < N * (NP / IPT); idx += W * 32) body(idx, std::false_type());
}

// N: transform length of the fused axis. CI: channels per group (in and out; full groups only). NP: pair items per CTA
// (NP = 2: a contraction thread takes both, reading the kernel spectrum once for four batch items). W: warps. PLAIN: identity gather map with all N points stored, a plain crop on store, a single segment.
// Shared memory: NP*CI pair lines of N slots; a line is its warp's exchange buffer during the transforms and carries the
// spectrum in natural order between the phases.
// TPL: threads per pair line in the transforms: 32 (a warp per line) or 64 (two warps per line, 8 points per lane at
// N = 512: twice the resident warps at half the registers).
template <int N, int CI, int NP, int W, bool PLAIN, int OCC, int TPL = 32>
__global__ void __launch_bounds__(W * 32, OCC) fc_pair_fused_kernel(fc_pair_fused_args a) {
  fc_grid_dep_sync();
  constexpr int E = N / TPL, WPL = TPL / 32, NG = W / WPL;  // points per lane; warps per line; line groups of the CTA
  static_assert(TPL == 32 || TPL == 64, "a warp or two per pair line");
  FC_DYN_SMEM(smem_raw);
  fc_c2* xy = reinterpret_cast<fc_c2*>(smem_raw);
  const int tid = threadIdx.x, w = tid >> 5;
  const int lane = (w % WPL) * 32 + (tid & 31);  // position among the TPL threads of a line
  const int grp = w / WPL, bar = 2 + grp;         // line group of this warp and its named barrier (TPL = 64)
  const fc_omap om = a.omap;
  const int out_lim = om.Lout < om.lim ? om.Lout : om.lim;
  // zero padding without zero-stuffing / subsampling: dense position u holds source u - pad for u in [u_lo, u_hi)
  const bool simple_in = a.imap.mode == FC_PAD_CONSTANT && a.imap.up == 1 && a.imap.sub == 1;
  const int u_lo = a.imap.pad > 0 ? a.imap.pad : 0;
  const int u_hi = a.imap.ext < a.imap.L + a.imap.pad ? a.imap.ext : a.imap.L + a.imap.pad;
  const int n_units = (int)a.n_units, R = (int)a.R, Rk = (int)a.Rk, nsx = R / Rk;  // (host: n_units < 2^31)
#ifndef FC_CPU_EMUL
  // The CTAs that share an SM start together and, their units being equal, would run their three phases in lockstep:
  // load latency, transforms and contraction of both at the same moments. Delaying every second CTA of the first wave
  // by a fraction of a unit puts one CTA's transforms under the other's kernel-spectrum loads for the rest of the grid.
  if (a.desync_ns > 0) {
    const int grp = blockIdx.x / a.desync_mod;
    if (grp > 0 && grp < a.desync_grp) {
      const unsigned total = (unsigned)a.desync_ns * (unsigned)grp;
      for (unsigned waited = 0; waited < total; waited += 1000) __nanosleep(1000);
    }
  }
#endif
  for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const int bs = unit % a.nbs;
    const int gr = unit / a.nbs;
    // lines of the other axis: when that axis is segmented (Rk < R) the segments sharing kernel line rk run back to
    // back, so the kernel-spectrum slice of (g, rk) is read from HBM once and from L2 afterwards
    int r, rk, g;
    if (nsx == 1) {
      g = gr / R;
      r = rk = gr - g * R;
    } else {
      const int t = gr / nsx;
      g = t / Rk;
      rk = t - g * Rk;
      r = (gr - t * nsx) * Rk + rk;
    }
    const int it0 = bs * NP;
    if (a.k_pf > 0) {  // the kernel-spectrum lines the contraction reads first travel to L1 while the transforms run
      const float4* kp = reinterpret_cast<const float4*>(a.kspec + ((int64_t)g * a.Rk + rk) * ((int64_t)CI * CI * N)) + (tid & (N - 1));
      for (int o = 0; o < a.k_pf; ++o)
#pragma unroll
        for (int i2 = 0; i2 < CI / 2; ++i2) fc_prefetch_l1(kp + (o * (CI / 2) + i2) * N);
    }
    // ---- phase 1: forward transform of every (pair item, input channel) line of this bin
#pragma unroll 1
    for (int tk = grp; tk < CI * NP; tk += NG) {  // uniform over the threads of a line
      const int pg = tk / CI, i = tk - pg * CI;
      fc_c2* line0 = xy + (size_t)(pg * CI + i) * N;
      fc_c2 v[1][E];
      const int item = it0 + pg;
      const bool active = item < a.n_items;
      const int it = active ? item : it0;
      const int bp = PLAIN ? it : it / a.n_seg, sg = PLAIN ? 0 : it - bp * a.n_seg;
      const fc_c2* src = a.xin + (((int64_t)bp * a.Cin + g * CI + i) * R + r) * a.n_in;
      if (PLAIN) {
#pragma unroll
        for (int q = 0; q < E; ++q) v[0][q] = (active && !FC_ABL(a, 8)) ? c2_ld_stream(src + lane + TPL * q) : c2_zero();
      } else if (simple_in) {
        const int ub = sg * a.seg_V - a.seg_off + lane;
        src -= a.imap.pad;
#pragma unroll
        for (int q = 0; q < E; ++q) {
          const int u = ub + TPL * q;
          v[0][q] = (active && u >= u_lo && u < u_hi) ? c2_ldg(src + u) : c2_zero();
        }
      } else {
        const int ub = sg * a.seg_V - a.seg_off + lane;
#pragma unroll
        for (int q = 0; q < E; ++q) {
          const int s = fc_imap_src(a.imap, ub + TPL * q);
          v[0][q] = (active && s >= 0) ? c2_ldg(src + s) : c2_zero();
        }
      }
      if (!FC_ABL(a, 1)) fc_pfft<N, TPL, 1, N>(v, line0, a.tw, a.tw_len, lane, bar);  // the line itself is the exchange buffer
#pragma unroll
      for (int q = 0; q < E; ++q) line0[lane + TPL * q] = v[0][q];
    }
    float4 kpre[FC_KPRE > 0 ? FC_KPRE : 1];
    fc_pair_contract_preload<N, CI, NP, NP, W>(kpre, a, g, rk, tid);
    fc_named_bar_sync(1, W * 32);
    // ---- L2 prefetch for the unit that runs `prefetch_dist` units later (the next wave on this SM)
    if (a.prefetch_dist > 0) {
      const int un = unit + a.prefetch_dist;
      if (un < n_units) {
        const int bsn = un % a.nbs;
        const int grn = un / a.nbs;
        int rn, rkn, gn;
        if (nsx == 1) {
          gn = grn / R;
          rn = rkn = grn - gn * R;
        } else {
          const int t = grn / nsx;
          gn = t / Rk;
          rkn = t - gn * Rk;
          rn = (grn - t * nsx) * Rk + rkn;
        }
        if (a.n_seg == 1) {  // (segments of one batch pair share their input line: nothing to pull ahead)
          const int per_line = (a.n_in * 16 + 127) / 128;  // 128-byte lines per input pair line
          for (int idx = tid; idx < NP * CI * per_line; idx += W * 32) {
            const int ln = idx / per_line, seg = idx - ln * per_line;
            const int pl = ln / CI, i = ln - pl * CI;
            if (bsn * NP + pl < a.BP)
              fc_prefetch_l2(a.xin + (((int64_t)(bsn * NP + pl) * a.Cin + gn * CI + i) * a.R + rn) * a.n_in + seg * 8);
          }
        }
        if (bsn == 0) {  // and, once per bin, its slice of the kernel spectrum (one contiguous block)
          const float2* ks = a.kspec + ((int64_t)gn * a.Rk + rkn) * ((int64_t)CI * CI * N);
          for (int idx = tid; idx < CI * CI * (N / 16); idx += W * 32) fc_prefetch_l2(ks + idx * 16);
        }
      }
    }
    // ---- phase 2: per-bin contraction over the input channels of the group, in place (X -> Y)
    if (!FC_ABL(a, 32)) {
#ifdef FC_TUNING
      if (NP > 1 && a.k_share)
        fc_pair_contract<N, CI, NP, 1, W, true>(xy, a, g, rk, tid, kpre);
      else
#endif
        fc_pair_contract<N, CI, NP, NP, W>(xy, a, g, rk, tid, kpre);
    }
    fc_named_bar_sync(1, W * 32);
    // ---- phase 3: inverse transform of every (pair item, output channel) line, crop / stride on store
#pragma unroll 1
    for (int tk = grp; tk < CI * NP; tk += NG) {  // uniform over the threads of a line
      const int pg = tk / CI, o = tk - pg * CI;
      const int item = it0 + pg;
      fc_c2* line0 = xy + (size_t)(pg * CI + o) * N;
      fc_c2 v[1][E];
#pragma unroll
      for (int q = 0; q < E; ++q) v[0][q] = c2_swap(line0[lane + TPL * q]);
      fc_group_sync<TPL>(bar);  // the line becomes the exchange buffer: every lane must have read its inputs
      if (!FC_ABL(a, 4)) fc_pfft<N, TPL, 1, N>(v, line0, a.tw, a.tw_len, lane, bar);
      if (PLAIN) {
        if (item < a.n_items && !FC_ABL(a, 16)) {  // PLAIN: items are batch pairs
          fc_c2* dst = a.yout + (((int64_t)item * a.Cout + g * CI + o) * R + r) * a.n_out;
#pragma unroll
          for (int q = 0; q < E; ++q) {
            const int n = lane + TPL * q;
            if (n < out_lim) dst[n] = c2_swap(v[0][q]);  // PLAIN: Lout <= lim
          }
        }
      } else {
        // general crop / stride / lattice map: stage the line in shared memory, then output-driven coalesced stores
#pragma unroll
        for (int q = 0; q < E; ++q) line0[lane + TPL * q] = c2_swap(v[0][q]);
        fc_group_sync<TPL>(bar);
        if (item < a.n_items) {
          const int bp = item / a.n_seg, sg = item - bp * a.n_seg;
          fc_c2* dst = a.yout + (((int64_t)bp * a.Cout + g * CI + o) * R + r) * a.n_out;
          // this item owns the dense outputs n in [n_lo, n_hi), i.e. the outputs j with n(j) = (j*os + ob) / og in that
          // range: a contiguous run of j because n(j) is monotone
          const int n_lo = sg * a.seg_V, n_hi = n_lo + a.seg_V;
          const int c_lo = n_lo * om.og - om.ob, c_hi = n_hi * om.og - om.ob;
          int j_lo = c_lo > 0 ? (c_lo + om.os - 1) / om.os : 0;
          int j_hi = c_hi > 0 ? (c_hi + om.os - 1) / om.os : 0;
          if (j_hi > om.Lout) j_hi = om.Lout;
          const fc_c2* ln = line0 + (a.seg_off - n_lo);
#pragma unroll 4
          for (int j = j_lo + lane; j < j_hi; j += TPL) {
            const int tt = j * om.os + om.ob;
            const int n = om.og == 1 ? tt : om.og == 2 ? (tt >> 1) : tt / om.og;
            dst[j] = (tt == n * om.og && n < om.lim) ? ln[n] : c2_zero();
          }
        }
      }
    }
    fc_named_bar_sync(1, W * 32);
  }
}

// ------------------------------------------------------------------------------------------------ bulk async copy
// One contiguous block global -> shared through the TMA engine (cp.async.bulk, SASS UBLKCP), completion counted in
// bytes on an mbarrier. The host-thread emulation copies synchronously (the phase barrier that follows orders it).
#ifdef FC_CPU_EMUL
FC_DEV void fc_mbar_init(uint64_t*, uint32_t) {}
FC_DEV void fc_mbar_wait(uint64_t*, uint32_t) {}
FC_DEV void fc_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t*) { std::memcpy(dst, src, bytes); }
#else
FC_DEV uint32_t fc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
FC_DEV void fc_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fc_smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
FC_DEV void fc_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(fc_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// expect `bytes` on `bar`, then copy them (issued by one thread)
FC_DEV void fc_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fc_smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(fc_smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(fc_smem_u32(bar))
               : "memory");
}
#endif

// ------------------------------------------------------------------------------------------------ KB64
// The fused kernel of the y-stage program (K1p / K4p with YS > 0). With ky = k1 + YS*k2 the transform of the fused axis
// has been cut into YS independent 64-point transforms per line, and sub-problem k1 only touches the bins {k1 + YS*k2}:
// 1/YS of the kernel spectrum of the line. A unit is therefore (group, bin kx of the other axis, k1) for NPI pair items
// at once (NPI = 4: eight batch items): 4*CI pair lines of 64 slots = 32 KB of shared memory for CI = 8, one contiguous
// 32 KB chunk [o][i/2][k2][i%2] of kernel spectrum that is read from L2 once for all eight batch items (the four items'
// threads walk it in step, three of them hit L1), a single exchange per transform instead of two, and five to six CTAs
// per SM whose phases interleave. The one-line-per-unit kernel above re-read the 262 KB slice of a bin for every pair
// item: 270 MB from L2 at BASELINE c2, where L2 -> SM bandwidth (measured 9 TB/s, profiles/r2_l2_bandwidth.txt) is
// the binding resource.
struct fc_pair_fused64_args {
  const fc_c2* xin;     // [(p*Cin + c)][R][YS][S] slots     output of K1p (y stage)
  const float2* kspec;  // [group][Rk][YS][o][i/2][S][i%2]     (fc_pass::out_split / out_il)
  fc_c2* yout;          // [(p*Cout + o)][R][YS][S] slots    input of K4p (y stage)
  const float2* tw;
  int32_t tw_len;
  int32_t BP, Cin, Cout, G, YS;
  int32_t nbs;      // blocks of NPI pair items per (group, bin, k1)
  int64_t R, Rk, n_units;  // lines of the other axis; kernel-spectrum lines (line r uses r % Rk: row segments share them)
};

// S: length of the sub-transforms (64: 8 lanes per line, one exchange; 128: 16 lanes per line, two exchanges), NPI = 256/S.
// Shared memory: [NPI*CI pair lines][S] slots, then the kernel-spectrum chunk [CI*CI/2][S] float4 (32 KB for S = 64, CI = 8),
// then the mbarrier. The chunk is requested by one thread at the top of the unit (one cp.async.bulk) and lands while
// the forward transforms run: the contraction reads it from shared memory and never waits on L2.
template <int S, int CI, int OCC>
__global__ void __launch_bounds__(256, OCC) fc_pair_fused64_kernel(fc_pair_fused64_args a) {
  fc_grid_dep_sync();
  constexpr int W = 8, NPI = W * 32 / S, NLINES = NPI * CI;  // pair items and pair lines of a unit
  constexpr int G = S / 8, LPW = 32 / G;                      // lanes per line (8 points per lane), lines per warp pass
  constexpr uint32_t KBYTES = CI * CI * S * 8;
  static_assert((S == 64 || S == 128) && NLINES % (W * LPW) == 0, "whole warp passes; one (bin, item) per contraction thread");
  FC_DYN_SMEM(smem_raw);
  fc_c2* xy = reinterpret_cast<fc_c2*>(smem_raw);  // [item][channel][64]
  float4* ks = reinterpret_cast<float4*>(xy + NLINES * S);
  uint64_t* bar = reinterpret_cast<uint64_t*>(ks + CI * CI / 2 * S);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int gl = lane % G, gid = lane / G;
  const int n_units = (int)a.n_units, R = (int)a.R, Rk = (int)a.Rk, YS = a.YS;
  if (tid == 0) fc_mbar_init(bar, 1);
  __syncthreads();
  uint32_t parity = 0;
  for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const int bs = unit % a.nbs;
    int t = unit / a.nbs;
    const int k1 = t % YS;
    t /= YS;
    const int r = t % R, g = t / R;
    const int it0 = bs * NPI;
    if (tid == 0) fc_bulk_g2s(ks, a.kspec + ((((int64_t)g * Rk + r % Rk) * YS + k1) * ((int64_t)CI * CI * S)), KBYTES, bar);
    // ---- phase 1: 64-point forward transforms of the (item, input channel) lines
#pragma unroll 1
    for (int ln = w * LPW + gid; ln < NLINES; ln += W * LPW) {  // uniform over the lanes of a group; all groups run every pass
      const int pg = ln / CI, i = ln - pg * CI;
      const bool active = it0 + pg < a.BP;
      const fc_c2* src = a.xin + ((((int64_t)(active ? it0 + pg : it0) * a.Cin + g * CI + i) * R + r) * YS + k1) * S;
      fc_c2* line0 = xy + (size_t)ln * S;
      fc_c2 v[1][8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[0][q] = active ? c2_ld_stream(src + gl + G * q) : c2_zero();
      fc_pfft<S, G, 1, S>(v, line0, a.tw, a.tw_len, gl);
#pragma unroll
      for (int q = 0; q < 8; ++q) line0[gl + G * q] = v[0][q];
    }
    fc_named_bar_sync(1, W * 32);
    fc_mbar_wait(bar, parity);  // the kernel-spectrum chunk has landed
    parity ^= 1;
    // ---- phase 2: contraction of bin k2 = tid % S of item tid / S over the input channels, in place
    {
      const int k2 = tid % S, pg = tid / S;
      fc_c2* xb = xy + (size_t)(pg * CI) * S + k2;
      fc_c2 x[CI];
#pragma unroll
      for (int i = 0; i < CI; ++i) x[i] = xb[i * S];
      const float4* kp = ks + k2;
#pragma unroll
      for (int o = 0; o < CI; ++o) {
        fc_p2 rr = p2_make(0.f, 0.f), ii = rr, ri = rr, ir = rr;
#pragma unroll
        for (int i2 = 0; i2 < CI / 2; ++i2) {
          const float4 k = kp[(o * (CI / 2) + i2) * S];
          rr = p2_fmas(x[2 * i2].re, k.x, rr);
          ii = p2_fmas(x[2 * i2].im, k.y, ii);
          ri = p2_fmas(x[2 * i2].re, k.y, ri);
          ir = p2_fmas(x[2 * i2].im, k.x, ir);
          rr = p2_fmas(x[2 * i2 + 1].re, k.z, rr);
          ii = p2_fmas(x[2 * i2 + 1].im, k.w, ii);
          ri = p2_fmas(x[2 * i2 + 1].re, k.w, ri);
          ir = p2_fmas(x[2 * i2 + 1].im, k.z, ir);
        }
        xb[o * S] = c2_make(p2_sub(rr, ii), p2_add(ri, ir));
      }
    }
    fc_named_bar_sync(1, W * 32);  // (also: every thread is done with the chunk before the next unit's copy overwrites it)
    // ---- phase 3: 64-point inverse transforms of the (item, output channel) lines
#pragma unroll 1
    for (int ln = w * LPW + gid; ln < NLINES; ln += W * LPW) {
      const int pg = ln / CI, o = ln - pg * CI;
      fc_c2* line0 = xy + (size_t)ln * S;
      fc_c2 v[1][8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[0][q] = c2_swap(line0[gl + G * q]);
      FC_SYNCWARP();
      fc_pfft<S, G, 1, S>(v, line0, a.tw, a.tw_len, gl);
      if (it0 + pg < a.BP) {
        fc_c2* dst = a.yout + ((((int64_t)(it0 + pg) * a.Cout + g * CI + o) * R + r) * YS + k1) * S;
#pragma unroll
        for (int q = 0; q < 8; ++q) dst[gl + G * q] = c2_swap(v[0][q]);
      }
    }
    fc_named_bar_sync(1, W * 32);
  }
}

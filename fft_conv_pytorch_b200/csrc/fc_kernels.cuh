// fc_kernels.cuh — generic axis-pass and contraction kernels (any power-of-two extent, any channel count).
//
// These are the "always correct" kernels of the pipeline; the fused / specialised kernels for the
// benchmark shapes live in fc_fused.cuh and are checked against these. The code is plain CUDA C++ with no
// warp intrinsics so that tests/cpu_emul can run the very same source on host threads (test infrastructure
// only; the shipped library contains device code only).
#pragma once
#include "fc_types.h"

#ifdef FC_CPU_EMUL
#include "cuda_shim.h"
#else
#include <cuda_runtime.h>
#define FC_DYN_SMEM(name)                      \
  extern __shared__ float4 fc_dyn_smem_raw[];  \
  float2* name = reinterpret_cast<float2*>(fc_dyn_smem_raw)
#endif

#define FC_DEV __device__ __forceinline__

// Programmatic dependent launch: every kernel of the library is launched with the programmatic-stream-serialization
// attribute (fc_api.cu: FC_LAUNCH), so its CTAs may become resident while the previous kernel of the stream is still
// draining. The first statement of every kernel lets its own successor do the same and then waits until the
// predecessor grids have completed and their writes are visible; nothing before it may touch global memory.
#ifdef FC_CPU_EMUL
FC_DEV void fc_grid_dep_sync() {}
#else
FC_DEV void fc_grid_dep_sync() { asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory"); }
#endif

// Named barrier over a subset of a CTA's threads (the compute warps of the fused kernels); tests/cpu_emul supplies a
// host stand-in with the same semantics.
#ifdef FC_CPU_EMUL
void fc_emul_named_barrier(int id, int count);  // tests/cpu_emul/cuda_shim.cpp
FC_DEV void fc_named_bar_sync(int id, int threads) { fc_emul_named_barrier(id, threads); }
#else
FC_DEV void fc_named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
#endif

// Streaming accesses: data that is read or written exactly once by the pipeline (the real input, a spectrum on its way
// out of L2, the final output) is moved with the evict-first policy so that it does not push the spectra the next kernel
// is about to read out of the 126 MB L2.
#if defined(FC_CPU_EMUL) || defined(FC_NO_STREAM_HINTS)
template <typename T>
FC_DEV T fc_ld_stream(const T* p) { return __ldg(p); }
template <typename T>
FC_DEV void fc_st_stream(T* p, T v) { *p = v; }
#else
template <typename T>
FC_DEV T fc_ld_stream(const T* p) { return __ldcs(p); }
template <typename T>
FC_DEV void fc_st_stream(T* p, T v) { __stcs(p, v); }
#endif

// ------------------------------------------------------------------------------------------------ complex helpers
FC_DEV float2 fc_c(float x, float y) { return make_float2(x, y); }
#if defined(FC_CPU_EMUL) || !defined(FC_PACKED_F32X2)
FC_DEV float2 fc_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
FC_DEV float2 fc_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#else
// Blackwell packed fp32: one FADD2 adds both components of a complex number (halves the issue slots of the
// add/sub-dominated butterflies). Same IEEE result as two scalar adds.
FC_DEV float2 fc_add(float2 a, float2 b) {
  float2 r;
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
FC_DEV float2 fc_sub(float2 a, float2 b) {
  float2 r;
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
#endif
FC_DEV float2 fc_mul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
FC_DEV float2 fc_conj(float2 a) { return make_float2(a.x, -a.y); }
FC_DEV float2 fc_mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)
FC_DEV float2 fc_scale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }

// XOR swizzle of the element index inside a line (only bits 4..6 of the index are looked at, only the low 4 bits
// change). For the access patterns of a Stockham stage — 16 consecutive points; the stride-8 / stride-4 writes of the
// first stage; the "8 consecutive, jump 64" writes of the second stage — the 16 lanes of a half warp touch 16
// distinct 8-byte bank pairs.
FC_DEV int fc_swz2(int p) {
  const int h = p >> 4;
  return p ^ ((h & 7) | (((h >> 2) & 1) << 3));
}
FC_DEV int fc_swz(int i) { return fc_swz2(i); }

// ------------------------------------------------------------------------------------------------ maps
// Dense position u -> source index, or -1 for a structural zero. (fc_types.h: fc_imap)
FC_DEV int fc_imap_src(const fc_imap& m, int u) {
  if (u < 0 || u >= m.ext) return -1;
  int w = u;
  if (m.up > 1) {
    if (u % m.up) return -1;
    w = u / m.up;
  }
  int v = w * m.sub - m.pad;
  if (v < 0 || v >= m.L) {
    if (m.mode == FC_PAD_CONSTANT) return -1;
    if (m.mode == FC_PAD_REFLECT)
      v = (v < 0) ? -v : 2 * (m.L - 1) - v;
    else if (m.mode == FC_PAD_REPLICATE)
      v = (v < 0) ? 0 : m.L - 1;
    else
      v = (v < 0) ? v + m.L : v - m.L;  // circular
  }
  return v;
}

// ------------------------------------------------------------------------------------------------ radix butterflies
// Forward DFT (sign -1) of R points held in registers; results in natural order.
template <int R>
FC_DEV void fc_butterfly(float2* v);

template <>
FC_DEV void fc_butterfly<2>(float2* v) {
  float2 a = v[0], b = v[1];
  v[0] = fc_add(a, b);
  v[1] = fc_sub(a, b);
}

template <>
FC_DEV void fc_butterfly<4>(float2* v) {
  float2 b0 = fc_add(v[0], v[2]), b2 = fc_sub(v[0], v[2]);
  float2 b1 = fc_add(v[1], v[3]), b3 = fc_mul_mi(fc_sub(v[1], v[3]));
  v[0] = fc_add(b0, b1);
  v[2] = fc_sub(b0, b1);
  v[1] = fc_add(b2, b3);
  v[3] = fc_sub(b2, b3);
}

template <>
FC_DEV void fc_butterfly<8>(float2* v) {
  const float h = 0.70710678118654752440f;
  float2 a0 = fc_add(v[0], v[4]), a4 = fc_sub(v[0], v[4]);
  float2 a1 = fc_add(v[1], v[5]), a5 = fc_sub(v[1], v[5]);
  float2 a2 = fc_add(v[2], v[6]), a6 = fc_sub(v[2], v[6]);
  float2 a3 = fc_add(v[3], v[7]), a7 = fc_sub(v[3], v[7]);
  a5 = make_float2(h * (a5.x + a5.y), h * (a5.y - a5.x));   // * (1 - i)/sqrt2
  a6 = fc_mul_mi(a6);                                       // * (-i)
  a7 = make_float2(h * (a7.y - a7.x), -h * (a7.x + a7.y));  // * (-1 - i)/sqrt2
  float2 b0 = fc_add(a0, a2), b2 = fc_sub(a0, a2);
  float2 b1 = fc_add(a1, a3), b3 = fc_mul_mi(fc_sub(a1, a3));
  float2 b4 = fc_add(a4, a6), b6 = fc_sub(a4, a6);
  float2 b5 = fc_add(a5, a7), b7 = fc_mul_mi(fc_sub(a5, a7));
  v[0] = fc_add(b0, b1);
  v[4] = fc_sub(b0, b1);
  v[2] = fc_add(b2, b3);
  v[6] = fc_sub(b2, b3);
  v[1] = fc_add(b4, b5);
  v[5] = fc_sub(b4, b5);
  v[3] = fc_add(b6, b7);
  v[7] = fc_sub(b6, b7);
}

// One Stockham stage of radix R over T lines of M points: in -> out (both swizzled, same pitch).
// Ns = product of the radices already applied. tw is the table exp(-2*pi*i*j/tw_len).
template <int R>
FC_DEV void fc_fft_stage(const float2* in, float2* out, int M, int Ns, int T, int pitch, const float2* tw, int tw_len) {
  const int per = M / R;  // butterflies per line (a power of two)
  const int lper = 31 - __clz(per);
  const int total = T * per;
  const int tw_step = tw_len / (Ns * R);
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int line = idx >> lper;
    const int j = idx - line * per;
    const int k = j & (Ns - 1);
    const float2* src = in + line * pitch;
    float2* dst = out + line * pitch;
    float2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = src[fc_swz(j + r * per)];
    if (Ns > 1) {
      const float2 w1 = __ldg(tw + k * tw_step);
      float2 w = w1;
#pragma unroll
      for (int r = 1; r < R; ++r) {
        v[r] = fc_mul(v[r], w);
        if (r + 1 < R) w = fc_mul(w, w1);
      }
    }
    fc_butterfly<R>(v);
    const int j0 = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) dst[fc_swz(j0 + r * Ns)] = v[r];
  }
}

// Forward complex FFT of T lines of M points (M power of two), unnormalised, Stockham autosort with
// radix 8/4/2 stages ping-ponging between a and b. Data must be visible (barrier) on entry; a barrier
// follows every stage. Returns the buffer holding the result (swizzled, natural order).
FC_DEV float2* fc_fft_forward(float2* a, float2* b, int M, int T, int pitch, const float2* tw, int tw_len) {
  int Ns = 1;
  float2* in = a;
  float2* out = b;
  while (Ns < M) {
    const int rem = M / Ns;
    int R;
    if (rem >= 8 && rem != 16)
      R = 8;
    else if (rem >= 4)
      R = 4;
    else
      R = 2;
    if (R == 8)
      fc_fft_stage<8>(in, out, M, Ns, T, pitch, tw, tw_len);
    else if (R == 4)
      fc_fft_stage<4>(in, out, M, Ns, T, pitch, tw, tw_len);
    else
      fc_fft_stage<2>(in, out, M, Ns, T, pitch, tw, tw_len);
    __syncthreads();
    Ns *= R;
    float2* t = in;
    in = out;
    out = t;
  }
  return in;
}

// Four-step twiddle W_N^(k*r), N = N1*N2 (powers of two), from two small tables: with e = (k*r) mod N = a*N2 + b,
// W_N^e = W_N1^a * W_N^b. tw holds exp(-2*pi*i*j/tw_len) (tw_len >= N1) followed by exp(-2*pi*i*b/N), b < N2.
FC_DEV float2 fc_big_twiddle(int64_t k, int64_t r, const fc_pass& p, const float2* tw) {
  const int64_t e = (k * r) & (p.twN - 1);
  const int l2 = 31 - __clz(p.tw2_len);
  const int a = (int)(e >> l2), b = (int)(e & (p.tw2_len - 1));
  const int n1 = (int)(p.twN >> l2);
  return fc_mul(__ldg(tw + a * (p.tw_len / n1)), __ldg(tw + p.tw_len + b));
}

// ------------------------------------------------------------------------------------------------ the axis pass
struct fc_pass_args {
  fc_pass p;
  const void* in;
  void* out;
  const float2* tw;   // twiddle table, p.tw_len entries
  const float* bias;  // C2R only, may be null
};

// Per-line bookkeeping kept in shared memory for the current tile.
struct fc_line_info {
  int64_t in_base;
  int64_t out_base;
  int64_t r;  // line index inside the outer item (twiddle / composite position)
  int32_t valid;
  int32_t bias_idx;
  int32_t shift;  // batch segments (fc_pass::bseg_*), R2C: dense offset of this item's segment, added to the gather position
  int32_t lout;   // C2R: outputs of this line (omap.Lout, or what is left of the user's line for the last batch segment)
};

#define FC_MAX_T 128
#define FC_LD_BATCH 8 /* global loads a thread of the generic pass keeps in flight */

template <int KIND>
__global__ void fc_pass_kernel(fc_pass_args a) {
  fc_grid_dep_sync();
  const fc_pass& p = a.p;
  FC_DYN_SMEM(smem);
  float2* bufA = smem;
  float2* bufB = smem + (size_t)p.T * p.pitch;
  __shared__ fc_line_info lines[FC_MAX_T];

  const int T = p.T, M = p.M, N = p.N, pitch = p.pitch;
  const int tid = threadIdx.x, nth = blockDim.x;
  const int lN = 31 - __clz(N), lM = 31 - __clz(M);
  // idx / (M + 1) for idx < 2^20 through a multiply-high: ceil(2^32 / W) is exact there because W <= 4097
  const unsigned rW = (unsigned)((0x100000000ull + (unsigned)M) / (unsigned)(M + 1));
  (void)lN;
  (void)lM;
  (void)rW;

  for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    // ---- line bookkeeping
    for (int lt = tid; lt < T; lt += nth) {
      int64_t o, r;
      int valid;
      if (p.flat) {
        const int64_t g = tile * T + lt;
        valid = g < p.n_outer * p.R;
        o = valid ? g / p.R : 0;
        r = valid ? g - o * p.R : 0;
      } else {
        o = tile / p.tiles_per_outer;
        r = (tile - o * p.tiles_per_outer) * T + lt;
        valid = r < p.R;
        if (!valid) r = 0;
      }
      fc_line_info li;
      if (KIND == FC_R2C) {
        const int64_t o1 = o / p.o_c2, o2 = o - o1 * p.o_c2;
        li.in_base = (o1 / p.o_q) * p.o_sA + (o1 % p.o_q) * p.o_sB + o2 * p.o_sC + r * p.in_rs;
      } else {
        li.in_base = o * p.in_os + r * p.in_rs;
      }
      // (bin-major kernel spectrum of the fused plans: outer items split into groups of out_oq, see fc_types.h)
      if (p.out_oq > 0) {
        const int64_t q = o % p.out_oq;
        li.out_base = (o / p.out_oq) * p.out_osA + (p.out_il > 1 ? (q / p.out_il) * p.out_il * p.out_os + q % p.out_il : q * p.out_os) + r * p.out_rs;
      } else {
        li.out_base = o * p.out_os + r * p.out_rs;
      }
      li.r = r;
      li.valid = valid;
      li.bias_idx = (int)(o % (p.cout > 0 ? p.cout : 1));
      li.shift = 0;
      li.lout = p.omap.Lout;
      if ((KIND == FC_R2C || KIND == FC_C2R) && p.bseg_n > 1) {
        const int64_t ob = o / p.bseg_c;                  // (batch, segment) item
        const int sg = (int)(ob % p.bseg_n);
        if (KIND == FC_R2C) {
          li.shift = sg * p.bseg_V;
        } else {
          li.out_base = ((ob / p.bseg_n) * p.bseg_c + (o - ob * p.bseg_c)) * (int64_t)p.bseg_Lout + (int64_t)sg * p.bseg_Vo + r * p.out_rs;
          const int left = p.bseg_Lout - sg * p.bseg_Vo;
          li.lout = left < p.omap.Lout ? left : p.omap.Lout;
        }
      }
      lines[lt] = li;
    }
    __syncthreads();

    // ---- load tile into bufA. Each thread first issues FC_LD_BATCH independent global loads, then stores them to
    // shared memory: the pass is a load -> transform -> store loop per CTA, so the loads in flight per thread are
    // what hides the HBM latency.
    if (KIND == FC_R2C) {
      const float* x = reinterpret_cast<const float*>(a.in);
      float* dst = reinterpret_cast<float*>(bufA);
      const int total = T * N;
      for (int base = tid; base < total; base += nth * FC_LD_BATCH) {
        float val[FC_LD_BATCH];
        int off[FC_LD_BATCH];
#pragma unroll
        for (int u = 0; u < FC_LD_BATCH; ++u) {
          const int idx = base + u * nth;
          val[u] = 0.f;
          off[u] = -1;
          if (idx < total) {
            int l, n;
            if (p.in_rfast) {
              l = idx & (T - 1);
              n = idx >> p.log2T;
            } else {
              l = idx >> lN;
              n = idx & (N - 1);
            }
            const fc_line_info li = lines[l];
            off[u] = 2 * (l * pitch + fc_swz(n >> 1)) + (n & 1);
            if (li.valid) {
              const int pos = n * p.pos_n + (int)li.r * p.pos_r;  // dense position < 2^25 (plan limit)
              int s = fc_imap_src(p.imap, pos);
              if (p.bseg_n > 1) {  // batch segment (plain zero-padding gather): the segment's window of the channel's line
                s = pos - p.imap.pad + li.shift;
                if (pos >= p.imap.ext || s < 0 || s >= p.imap.L) s = -1;
              }
              if (s >= 0) val[u] = __ldg(x + li.in_base + (int64_t)s * p.in_es);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < FC_LD_BATCH; ++u)
          if (off[u] >= 0) dst[off[u]] = val[u];
      }
    } else {
      // complex kinds: C2C_FWD (gather map), C2C_INV (conjugate), C2R (bins 0..M, plain layout, four-step twiddle)
      const float2* x = reinterpret_cast<const float2*>(a.in);
      const int W = M + 1;
      const int total = (KIND == FC_C2R) ? T * W : T * N;
      for (int base = tid; base < total; base += nth * FC_LD_BATCH) {
        float2 val[FC_LD_BATCH];
        int off[FC_LD_BATCH];
#pragma unroll
        for (int u = 0; u < FC_LD_BATCH; ++u) {
          const int idx = base + u * nth;
          val[u] = make_float2(0.f, 0.f);
          off[u] = -1;
          if (idx < total) {
            int l, n;
            if (p.in_rfast) {
              l = idx & (T - 1);
              n = idx >> p.log2T;
            } else if (KIND == FC_C2R) {
              l = (int)__umulhi((unsigned)idx, rW);
              n = idx - l * W;
            } else {
              l = idx >> lN;
              n = idx & (N - 1);
            }
            const fc_line_info li = lines[l];
            off[u] = l * pitch + (KIND == FC_C2R ? n : fc_swz(n));
            if (li.valid) {
              if (KIND == FC_C2C_FWD) {
                const int s = fc_imap_src(p.imap, n);
                if (s >= 0) val[u] = __ldg(x + li.in_base + (int64_t)s * p.in_es);
              } else if (KIND == FC_C2C_INV) {
                val[u] = fc_conj(__ldg(x + li.in_base + (int64_t)n * p.in_es));
              } else {
                val[u] = __ldg(x + li.in_base + (int64_t)n * p.in_es);
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < FC_LD_BATCH; ++u) {
          if (off[u] < 0) continue;
          if (KIND == FC_C2R && p.twiddle) {  // four-step twiddle, applied once the loads have been issued
            const int idx = base + u * nth;
            int l, n;
            if (p.in_rfast) {
              l = idx & (T - 1);
              n = idx >> p.log2T;
            } else {
              l = (int)__umulhi((unsigned)idx, rW);
              n = idx - l * W;
            }
            val[u] = fc_mul(val[u], fc_conj(fc_big_twiddle(n, lines[l].r, p, a.tw)));
          }
          bufA[off[u]] = val[u];
        }
      }
    }
    __syncthreads();

    // ---- transform
    float2* res;
    if (KIND == FC_C2R) {
      // Hermitian half spectrum Y[0..M] (plain, bufA) -> conj(Z) (swizzled, bufB); then forward FFT gives conj(z).
      const int total = T * M;
      const int tstep = p.tw_len / N;
      for (int idx = tid; idx < total; idx += nth) {
        const int l = idx >> lM;
        const int k = idx & (M - 1);
        const float2 yk = bufA[l * pitch + k];
        const float2 ym = fc_conj(bufA[l * pitch + (M - k)]);
        const float2 s = fc_add(yk, ym);
        const float2 d = fc_mul(fc_sub(yk, ym), fc_conj(__ldg(a.tw + k * tstep)));
        // Z = s + i*d ; store conj(Z)
        bufB[l * pitch + fc_swz(k)] = make_float2(s.x - d.y, -(s.y + d.x));
      }
      __syncthreads();
      res = fc_fft_forward(bufB, bufA, M, T, pitch, a.tw, p.tw_len);
    } else {
      res = fc_fft_forward(bufA, bufB, M, T, pitch, a.tw, p.tw_len);
    }

    // ---- store
    if (KIND == FC_R2C) {
      float2* other = (res == bufA) ? bufB : bufA;
      // untangle the packed real transform: Z (swizzled, res) -> X[0..M] (plain, other)
      const int W = M + 1;
      const int total = T * W;
      const int tstep = p.tw_len / N;
      for (int idx = tid; idx < total; idx += nth) {
        const int l = (int)__umulhi((unsigned)idx, rW);
        const int k = idx - l * W;
        const float2 zk = res[l * pitch + fc_swz(k & (M - 1))];
        const float2 zc = fc_conj(res[l * pitch + fc_swz((M - k) & (M - 1))]);
        const float2 e = fc_scale(fc_add(zk, zc), 0.5f);
        const float2 o = fc_scale(fc_mul_mi(fc_sub(zk, zc)), 0.5f);
        other[l * pitch + k] = fc_add(e, fc_mul(__ldg(a.tw + k * tstep), o));
      }
      __syncthreads();
      float2* y = reinterpret_cast<float2*>(a.out);
      for (int idx = tid; idx < total; idx += nth) {
        int l, k;
        if (p.out_rfast) {
          l = idx & (T - 1);
          k = idx >> p.log2T;
        } else {
          l = (int)__umulhi((unsigned)idx, rW);
          k = idx - l * W;
        }
        const fc_line_info li = lines[l];
        if (!li.valid) continue;
        float2 val = other[l * pitch + k];
        if (p.twiddle) {
          val = fc_mul(val, fc_big_twiddle(k, li.r, p, a.tw));
        }
        val = fc_scale(val, p.scale);
        if (p.conj_out) val = fc_conj(val);
        y[li.out_base + (int64_t)k * p.out_es] = val;
      }
    } else if (KIND == FC_C2C_FWD) {
      float2* y = reinterpret_cast<float2*>(a.out);
      const int total = T * N;
      for (int idx = tid; idx < total; idx += nth) {
        int l, k;
        if (p.out_rfast) {
          l = idx & (T - 1);
          k = idx >> p.log2T;
        } else {
          l = idx >> lN;
          k = idx & (N - 1);
        }
        const fc_line_info li = lines[l];
        if (!li.valid) continue;
        float2 val = fc_scale(res[l * pitch + fc_swz(k)], p.scale);
        if (p.conj_out) val = fc_conj(val);
        const int64_t off = p.out_split > 1 ? (int64_t)(k & (p.out_split - 1)) * p.out_split_stride + (int64_t)(k / (int)p.out_split) * p.out_es
                                            : (int64_t)k * p.out_es;
        y[li.out_base + off] = val;
      }
    } else if (KIND == FC_C2C_INV) {
      float2* y = reinterpret_cast<float2*>(a.out);
      const fc_omap om = p.omap;
      const int total = T * N;
      for (int idx = tid; idx < total; idx += nth) {
        int l, n;
        if (p.out_rfast) {
          l = idx & (T - 1);
          n = idx >> p.log2T;
        } else {
          l = idx >> lN;
          n = idx & (N - 1);
        }
        const fc_line_info li = lines[l];
        if (!li.valid) continue;
        const float2 val = fc_conj(res[l * pitch + fc_swz(n)]);
        // dense index n owns the outputs j with (j*os + ob) / og == n
        if (om.og == 1 && om.os == 1) {  // plain crop: no division on the common path
          const int j = n - om.ob;
          if (j >= 0 && j < om.Lout) y[li.out_base + (int64_t)j * p.out_es] = (n < om.lim) ? val : make_float2(0.f, 0.f);
        } else {
          for (int e = 0; e < om.og; ++e) {
            const int t = n * om.og + e - om.ob;
            if (t < 0) continue;
            int j = t;
            if (om.os != 1) {
              if (t % om.os) continue;
              j = t / om.os;
            }
            if (j >= om.Lout) continue;
            const bool live = (e == 0) && (n < om.lim);
            y[li.out_base + (int64_t)j * p.out_es] = live ? val : make_float2(0.f, 0.f);
          }
        }
      }
    } else {  // FC_C2R
      float* y = reinterpret_cast<float*>(a.out);
      const fc_omap om = p.omap;
      const int total = T * N;
      for (int idx = tid; idx < total; idx += nth) {
        int l, n;
        if (p.out_rfast) {
          l = idx & (T - 1);
          n = idx >> p.log2T;
        } else {
          l = idx >> lN;
          n = idx & (N - 1);
        }
        const fc_line_info li = lines[l];
        if (!li.valid) continue;
        const float2 z = res[l * pitch + fc_swz(n >> 1)];
        const float val = (n & 1) ? -z.y : z.x;
        const float b = p.has_bias ? __ldg(a.bias + li.bias_idx) : 0.f;
        const int u = n * p.pos_n + (int)li.r * p.pos_r;  // dense position < 2^25 (plan limit)
        if (p.row_og == 1 && om.og == 1 && om.os == 1) {  // plain crop: no division on the common path
          const int j = u - om.ob;
          if (j >= 0 && j < li.lout) y[li.out_base + (int64_t)j * p.out_es] = ((u < om.lim) ? val : 0.f) + b;
        } else {
          for (int er = 0; er < p.row_og; ++er) {  // output rows owned by this dense line (one unless row lattice)
            const int jr = (int)li.r * p.row_og + er - p.row_ob;
            if (p.row_og > 1 && (jr < 0 || jr >= p.row_Lout)) continue;
            float* yrow = y + li.out_base + (p.row_og > 1 ? (int64_t)(jr - (int)li.r) * p.out_rs : 0);
            for (int e = 0; e < om.og; ++e) {
              const int t = u * om.og + e - om.ob;
              if (t < 0) continue;
              int j = t;
              if (om.os != 1) {  // strided forward convolution: only every os-th dense sample is an output
                if (t % om.os) continue;
                j = t / om.os;
              }
              if (j >= li.lout) continue;
              const bool live = (e == 0) && (er == 0) && (u < om.lim);
              yrow[(int64_t)j * p.out_es] = (live ? val : 0.f) + b;
            }
          }
        }
      }
    }
    __syncthreads();  // smem (lines, buffers) is reused by the next tile
  }
}

// ------------------------------------------------------------------------------------------------ contraction
// Y[b, g*Og + o, f] = sum_i X[b, g*Ig + i, f] * K[g*Og + o, i, f]      (reference complex_matmul, functional.py:11-16)
// One thread per frequency bin (coalesced 8-byte accesses), a TB x TO register tile of outputs per thread; wide
// channel counts put several output tiles into one CTA so that they share the signal spectrum through L1.
struct fc_contract_args {
  const float2* X;
  const float2* K;
  float2* Y;
  int64_t bins;
  int32_t batch, cin, cout, groups;
  int32_t btiles, otiles;
};

// Contraction for channel groups too wide for register tiles alone (17 ... channels per group off the tensor-core path):
// a CTA computes a 16-item x 16-output tile for 32 adjacent bins. The input channels go through shared memory eight at a
// time (Xs / Ks: [8][16][32] complex, bins innermost, so both the coalesced global loads and the LDS of a warp are
// contiguous 256-byte rows), and thread (bin lane, sub-tile) accumulates 4 items x 8 outputs in registers: 12 LDS.64 feed
// 32 complex multiply-accumulates, and every signal / kernel value crosses L2 once per 16 outputs / 16 items instead of
// once per 8 (the register-tile kernel below is L2-bound there: 0.8 - 1.3 TB/s at 64 channels, profiles/r2b_shape_probe.txt).
#define FC_CT_TB 16
#define FC_CT_TO 16
#define FC_CT_KC 8
__global__ void __launch_bounds__(256, 3) fc_contract_tiled_kernel(fc_contract_args a) {
  fc_grid_dep_sync();
  FC_DYN_SMEM(smem);
  float2* Xs = smem;                                // [KC][TB][32]
  float2* Ks = smem + FC_CT_KC * FC_CT_TB * 32;     // [KC][TO][32]
  const int tid = threadIdx.x, lane = tid & 31, sub = tid >> 5;
  const int bsub = (sub & 3) * 4, osub = (sub >> 2) * 8;  // this thread's 4 items and 8 outputs inside the tile
  const int Ig = a.cin / a.groups, Og = a.cout / a.groups;
  const int g = blockIdx.z;
  const int bt = blockIdx.y / a.otiles, ot = blockIdx.y - bt * a.otiles;
  const int b0 = bt * FC_CT_TB, o0 = ot * FC_CT_TO;
  const int64_t f0 = (int64_t)blockIdx.x * 32;
  const bool f_ok = f0 + lane < a.bins;
  float2 acc[4][8];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[j][o] = make_float2(0.f, 0.f);
  for (int i0 = 0; i0 < Ig; i0 += FC_CT_KC) {
    // stage the next eight input channels: row r = (kk, j) of Xs / Ks is 32 bins of one (item, channel) / (output, channel)
    for (int r = sub; r < FC_CT_KC * FC_CT_TB; r += 8) {
      const int kk = r / FC_CT_TB, j = r - kk * FC_CT_TB;
      const int i = i0 + kk, b = b0 + j, o = o0 + j;
      const bool iv = i < Ig && f_ok;
      Xs[r * 32 + lane] = (iv && b < a.batch) ? __ldg(a.X + ((int64_t)b * a.cin + g * Ig + i) * a.bins + f0 + lane) : make_float2(0.f, 0.f);
      Ks[r * 32 + lane] = (iv && o < Og) ? __ldg(a.K + ((int64_t)(g * Og + o) * Ig + i) * a.bins + f0 + lane) : make_float2(0.f, 0.f);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < FC_CT_KC; ++kk) {
      float2 xv[4], kv[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) xv[j] = Xs[(kk * FC_CT_TB + bsub + j) * 32 + lane];
#pragma unroll
      for (int o = 0; o < 8; ++o) kv[o] = Ks[(kk * FC_CT_TO + osub + o) * 32 + lane];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          acc[j][o].x = fmaf(xv[j].x, kv[o].x, acc[j][o].x);
          acc[j][o].y = fmaf(xv[j].x, kv[o].y, acc[j][o].y);
          acc[j][o].x = fmaf(-xv[j].y, kv[o].y, acc[j][o].x);
          acc[j][o].y = fmaf(xv[j].y, kv[o].x, acc[j][o].y);
        }
    }
    __syncthreads();
  }
  if (!f_ok) return;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int b = b0 + bsub + j;
    if (b >= a.batch) continue;
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const int oc = o0 + osub + o;
      if (oc < Og) a.Y[((int64_t)b * a.cout + g * Og + oc) * a.bins + f0 + lane] = acc[j][o];
    }
  }
}

template <int TB, int TO>
__global__ void fc_contract_kernel(fc_contract_args a) {
  fc_grid_dep_sync();
  // blockDim = (bins, output-channel tiles): the tiles of one CTA read the same signal spectrum, which then comes
  // from L1 for all but the first of them
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= a.bins) return;
  const int Ig = a.cin / a.groups, Og = a.cout / a.groups;
  const int g = blockIdx.z;
  const int oblocks = (a.otiles + blockDim.y - 1) / blockDim.y;
  const int bt = blockIdx.y / oblocks;
  const int ot = (blockIdx.y - bt * oblocks) * blockDim.y + threadIdx.y;
  if (ot >= a.otiles) return;
  const int b0 = bt * TB, o0 = ot * TO;
  float2 acc[TB][TO];
#pragma unroll
  for (int i = 0; i < TB; ++i)
#pragma unroll
    for (int j = 0; j < TO; ++j) acc[i][j] = make_float2(0.f, 0.f);
  const float2* xp[TB];
  const float2* kp[TO];
#pragma unroll
  for (int i = 0; i < TB; ++i) {
    const int b = (b0 + i < a.batch) ? b0 + i : a.batch - 1;
    xp[i] = a.X + ((int64_t)b * a.cin + (int64_t)g * Ig) * a.bins + f;
  }
#pragma unroll
  for (int j = 0; j < TO; ++j) {
    const int o = (o0 + j < Og) ? o0 + j : Og - 1;
    kp[j] = a.K + ((int64_t)(g * Og + o) * Ig) * a.bins + f;
  }
  for (int c = 0; c < Ig; ++c) {
    float2 xv[TB], kv[TO];
#pragma unroll
    for (int i = 0; i < TB; ++i) xv[i] = __ldg(xp[i] + (int64_t)c * a.bins);
#pragma unroll
    for (int j = 0; j < TO; ++j) kv[j] = __ldg(kp[j] + (int64_t)c * a.bins);
#pragma unroll
    for (int i = 0; i < TB; ++i)
#pragma unroll
      for (int j = 0; j < TO; ++j) {
        acc[i][j].x = fmaf(xv[i].x, kv[j].x, acc[i][j].x);
        acc[i][j].y = fmaf(xv[i].x, kv[j].y, acc[i][j].y);
        acc[i][j].x = fmaf(-xv[i].y, kv[j].y, acc[i][j].x);
        acc[i][j].y = fmaf(xv[i].y, kv[j].x, acc[i][j].y);
      }
  }
#pragma unroll
  for (int i = 0; i < TB; ++i) {
    if (b0 + i >= a.batch) continue;
#pragma unroll
    for (int j = 0; j < TO; ++j) {
      if (o0 + j >= Og) continue;
      a.Y[((int64_t)(b0 + i) * a.cout + (int64_t)g * Og + o0 + j) * a.bins + f] = acc[i][j];
    }
  }
}

// Twiddle tables, evaluated in double precision: tw[j] = exp(-2*pi*i*j/len) for j < len, then (four-step plans)
// tw[len + b] = exp(-2*pi*i*b/big) for b < len2.
__global__ void fc_twiddle_kernel(float2* tw, int len, int len2, double big) {
  fc_grid_dep_sync();
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < len + len2; j += gridDim.x * blockDim.x) {
    double s, c;
    if (j < len)
      sincospi(-2.0 * (double)j / (double)len, &s, &c);
    else
      sincospi(-2.0 * (double)(j - len) / big, &s, &c);
    tw[j] = make_float2((float)c, (float)s);
  }
}

// fc_tc.cuh — tensor-core path of the per-bin channel contraction (complex_matmul, reference functional.py:11-16) for
// wide channel counts (BASELINE c4: 256 x 256 channels, 16 batches): a batched complex GEMM on the 5th-generation
// tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM) with a 3xTF32 split to keep fp32 accuracy.
//
// Per frequency bin f and group, with O = Cout/g, I = Cin/g:
//     Y[b, o] = sum_i X[b, i] * K[o, i]                      (complex)
// is the real GEMM  D[o, n] = sum_k A[o, k] * Bt[n, k],  M = O, N = 2B, Kdim = 2I:
//     A [o, :]      = [ Kr[o, 0..I) | Ki[o, 0..I) ]
//     Bt[2b,   :]   = [ Xr[b, 0..I) | -Xi[b, 0..I) ]   -> D[o, 2b]   = Re Y[b, o]
//     Bt[2b+1, :]   = [ Xi[b, 0..I) |  Xr[b, 0..I) ]   -> D[o, 2b+1] = Im Y[b, o]
// The spectra are produced bin-innermost by the FFT passes; three relayout kernels put them into the bin-outermost,
// channel-innermost form above (the kernel spectrum once, when it is cached).
//
// 3xTF32: the tensor core reads fp32 containers and uses the top 19 bits. With a = a_hi + a_lo (a_hi = a with the low
// 13 mantissa bits cleared, exactly what the hardware sees; a_lo = a - a_hi exactly representable),
// a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo; the dropped a_lo*b_lo term is 2^-22 relative.
#pragma once
#include "fc_kernels.cuh"

// ------------------------------------------------------------------------------------------------ relayouts
// in: [rows][bins] complex (bins contiguous)  ->  out per bin, see `mode`. One CTA moves a 32 x 32 tile through
// shared memory so that both sides are coalesced.
//   mode 0 (kernel spectrum): rows = (o, i);  out[f][o][0][i] = re, out[f][o][1][i] = im          (A rows)
//   mode 1 (signal spectrum): rows = (b, i);  out[f][2b][i] = re, out[f][2b][I+i] = -im,
//                                             out[f][2b+1][i] = im, out[f][2b+1][I+i] = re        (Bt rows)
//   mode 2 (product, inverse direction): in: D[f][o][2B] (float pairs = complex Y[f][o][b]) -> out[(b*O + o)][f]
struct fc_tc_relayout_args {
  const float2* in;
  float* out;
  int64_t bins;
  int32_t I;     // input channels per group
  int32_t C;     // mode 1: all input channels (G*I); mode 0 / 2: unused
  int32_t rows;  // mode 0: Cout*I, mode 1: B*Cin, mode 2: Cout*Bp (o-major, b inner)
  int32_t O;     // mode 0 / 2: all output channels (G*O_g)
  int32_t B;     // real batch
  int32_t Bp;    // padded batch (N = 2*Bp columns per bin); rows b >= B of Bt must be zero-filled by the caller
  int32_t mode;
};

__global__ void fc_tc_relayout_kernel(fc_tc_relayout_args a) {
  __shared__ float2 tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads
  const int64_t f0 = (int64_t)blockIdx.x * 32;
  const int r0 = blockIdx.y * 32;
  if (a.mode != 2) {
    // read rows r0.. (bins contiguous), write per bin with the row index contiguous
    for (int j = ty; j < 32; j += 8) {
      const int r = r0 + j;
      const int64_t f = f0 + tx;
      tile[j][tx] = (r < a.rows && f < a.bins) ? __ldg(a.in + (int64_t)r * a.bins + f) : make_float2(0.f, 0.f);
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
      const int64_t f = f0 + j;
      const int r = r0 + tx;
      if (r >= a.rows || f >= a.bins) continue;
      const float2 v = tile[tx][j];
      if (a.mode == 0) {
        const int o = r / a.I, i = r - o * a.I;  // o over all groups
        float* dst = a.out + ((int64_t)f * a.O + o) * (2 * a.I);
        dst[i] = v.x;
        dst[a.I + i] = v.y;
      } else {
        const int b = r / a.C, c = r - b * a.C;
        const int g = c / a.I, i = c - g * a.I;
        const int G = a.C / a.I;
        float* dst = a.out + (((int64_t)f * G + g) * (2 * a.Bp) + 2 * b) * (2 * a.I);
        dst[i] = v.x;
        dst[a.I + i] = -v.y;
        dst[2 * a.I + i] = v.y;
        dst[3 * a.I + i] = v.x;
      }
    }
  } else {
    // in: [f][o][bp] complex, rows index (o, bp) contiguous per bin -> out[(b*O + o)][f] for b < B
    const float2* in = a.in;
    for (int j = ty; j < 32; j += 8) {
      const int64_t f = f0 + j;
      const int r = r0 + tx;  // r = o*Bp + b
      tile[j][tx] = (r < a.rows && f < a.bins) ? __ldg(in + (int64_t)f * a.rows + r) : make_float2(0.f, 0.f);
    }
    __syncthreads();
    float2* out = reinterpret_cast<float2*>(a.out);
    for (int j = ty; j < 32; j += 8) {
      const int r = r0 + j;
      const int64_t f = f0 + tx;
      if (r >= a.rows || f >= a.bins) continue;
      const int o = r / a.Bp, b = r - o * a.Bp;
      if (b < a.B) out[((int64_t)b * a.O + o) * a.bins + f] = tile[tx][j];
    }
  }
}

#ifndef FC_CPU_EMUL
// ------------------------------------------------------------------------------------------------ tcgen05 helpers
namespace fc_tc {

FC_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

FC_DEV void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
FC_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
FC_DEV void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
FC_DEV void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
FC_DEV void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
FC_DEV void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// TMEM allocation by one full warp; the base address lands in *slot (shared memory).
FC_DEV void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
FC_DEV void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// Shared-memory matrix descriptor: K-major operand, 128-byte rows, SWIZZLE_128B, 8-row groups 1024 bytes apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version [46,48) = 1, layout [61,64) = 2).
FC_DEV uint64_t smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;            // leading byte offset: unused for swizzled K-major operands
  d |= (uint64_t)(1024 >> 4) << 32;  // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;            // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B tf32, both K-major, N>>3 at [17,23), M>>4 at [24,29).
FC_DEV constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

FC_DEV void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once every MMA issued so far by this thread has completed (implies fence::before_thread_sync).
FC_DEV void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 columns of fp32 from TMEM: thread t of the warp gets row (lane base + t), columns c0..c0+31.
FC_DEV void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

}  // namespace fc_tc

// ------------------------------------------------------------------------------------------------ the GEMM kernel
struct fc_tc_args {
  const float* A;   // [bins][G][O][2I]   kernel spectrum, relayout mode 0
  const float* Bt;  // [bins][G][2B][2I]  signal spectrum, relayout mode 1
  float* D;         // [bins][G][O][2B]   product (complex Y[f][g][o][b])
  int64_t n_items;  // bins * G
  int32_t O, I, B;  // per group; O % 128 == 0, (2I) % 32 == 0, B = padded batch: 2B in {16, 32, 48, 64}
};

// One CTA (256 threads) per SM, persistent over (bin, group) items. K loop in chunks of 32 fp32 (one 128-byte swizzle
// row); 3 shared-memory stages, each holding A (raw + low part) for up to 256 rows and Bt (raw + low part). All 8
// warps load (registers -> swizzled shared memory, computing the low parts on the way); thread 0 issues the MMAs;
// tcgen05.commit frees a stage and publishes the accumulator; warps 0-3 / 4-7 drain the two 128-row accumulators.
#define FC_TC_STAGES 3
#define FC_TC_MAXN 64
template <int MT /* 128-row tiles of O: 1 or 2 */>
__global__ void __launch_bounds__(256, 1) fc_tc_gemm_kernel(fc_tc_args a) {
  using namespace fc_tc;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = 2 * a.B, KD = 2 * a.I, n_chunks = KD / 32;
  constexpr int A_BYTES = MT * 128 * 128;  // one operand copy of A per stage: MT*128 rows x 128 bytes
  const int B_BYTES = N * 128;
  const int stage_bytes = 2 * A_BYTES + 2 * ((B_BYTES + 1023) & ~1023);
  unsigned char* sbase = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar_free[FC_TC_STAGES];  // MMA done with the stage
  __shared__ __align__(8) uint64_t bar_acc;                 // accumulator of the current item complete
  __shared__ uint32_t tmem_slot;
  if (tid == 0) {
    for (int s = 0; s < FC_TC_STAGES; ++s) mbar_init(&bar_free[s], 1);
    mbar_init(&bar_acc, 1);
    fence_barrier_init();
  }
  const uint32_t tmem_cols = (MT * N <= 32) ? 32 : (MT * N <= 64 ? 64 : (MT * N <= 128 ? 128 : (MT * N <= 256 ? 256 : 512)));
  if (warp == 0) tmem_alloc(&tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t idesc = idesc_tf32(128, N);

  // loader geometry: a chunk of A is MT*128 rows x 8 sixteen-byte pieces; thread t owns piece (t & 7) of rows t>>3 + 32*j
  const int piece = tid & 7;
  uint32_t free_phase[FC_TC_STAGES] = {0, 0, 0};  // parity to wait for on each stage (number of completed uses & 1)
  uint32_t stage_used[FC_TC_STAGES] = {0, 0, 0};
  uint32_t acc_phase = 0;
  int64_t chunk_seq = 0;  // running chunk counter of this CTA -> stage = chunk_seq % STAGES

  for (int64_t item = blockIdx.x; item < a.n_items; item += gridDim.x) {
    const float* Ag = a.A + item * (int64_t)a.O * KD;
    const float* Bg = a.Bt + item * (int64_t)N * KD;
    for (int mt0 = 0; mt0 < a.O; mt0 += MT * 128) {  // O > MT*128: several passes over the K loop
      // global -> registers, one chunk ahead of the chunk being written to shared memory
      float4 ra[MT * 4], rb[FC_TC_MAXN / 32], na[MT * 4], nb[FC_TC_MAXN / 32];
      auto load_chunk = [&](int c, float4 (&xa)[MT * 4], float4 (&xb)[FC_TC_MAXN / 32]) {
#pragma unroll
        for (int j = 0; j < MT * 4; ++j) {
          const int row = (tid >> 3) + 32 * j;
          xa[j] = __ldg(reinterpret_cast<const float4*>(Ag + (int64_t)(mt0 + row) * KD + c * 32) + piece);
        }
#pragma unroll
        for (int j = 0; j < FC_TC_MAXN / 32; ++j) {
          const int row = (tid >> 3) + 32 * j;
          xb[j] = (row < N) ? __ldg(reinterpret_cast<const float4*>(Bg + (int64_t)row * KD + c * 32) + piece) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      load_chunk(0, ra, rb);
      for (int c = 0; c < n_chunks; ++c, ++chunk_seq) {
        const int s = (int)(chunk_seq % FC_TC_STAGES);
        unsigned char* st = sbase + (size_t)s * stage_bytes;
        unsigned char* a_hi = st;
        unsigned char* a_lo = st + A_BYTES;
        unsigned char* b_hi = st + 2 * A_BYTES;
        unsigned char* b_lo = b_hi + ((B_BYTES + 1023) & ~1023);
        if (c + 1 < n_chunks) load_chunk(c + 1, na, nb);  // in flight while chunk c is staged and multiplied
        // wait until the MMAs that last read this stage are done
        if (stage_used[s]) {
          mbar_wait(&bar_free[s], free_phase[s]);
          free_phase[s] ^= 1;
        }
        stage_used[s] = 1;
        tc_fence_after();
        // registers -> swizzled shared memory: raw value (the tensor core ignores the low 13 bits) and exact low part
#pragma unroll
        for (int j = 0; j < MT * 4; ++j) {
          const int row = (tid >> 3) + 32 * j;
          const uint32_t off = (uint32_t)(row >> 3) * 1024 + (uint32_t)(row & 7) * 128 + (uint32_t)((piece ^ (row & 7)) << 4);
          const float4 v = ra[j];
          float4 lo;
          lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
          lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
          lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
          lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
          *reinterpret_cast<float4*>(a_hi + off) = v;
          *reinterpret_cast<float4*>(a_lo + off) = lo;
        }
#pragma unroll
        for (int j = 0; j < FC_TC_MAXN / 32; ++j) {
          const int row = (tid >> 3) + 32 * j;
          if (row < N) {
            const uint32_t off = (uint32_t)(row >> 3) * 1024 + (uint32_t)(row & 7) * 128 + (uint32_t)((piece ^ (row & 7)) << 4);
            const float4 v = rb[j];
            float4 lo;
            lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
            lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
            lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
            lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
            *reinterpret_cast<float4*>(b_hi + off) = v;
            *reinterpret_cast<float4*>(b_lo + off) = lo;
          }
        }
        fence_proxy_async();  // make the generic-proxy writes visible to the tensor core (async proxy)
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
          tc_fence_after();
          const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint32_t d = tmem_base + (uint32_t)(mt * N);  // accumulator of this 128-row tile: N columns
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // 4 x (K = 8 tf32 = 32 bytes) per 128-byte row
              const uint64_t dah = smem_desc_sw128(ah + mt * 16384 + k * 32);
              const uint64_t dal = smem_desc_sw128(al + mt * 16384 + k * 32);
              const uint64_t dbh = smem_desc_sw128(bh + k * 32);
              const uint64_t dbl = smem_desc_sw128(bl + k * 32);
              umma_tf32(d, dah, dbh, idesc, (c | k) ? 1u : 0u);
              umma_tf32(d, dal, dbh, idesc, 1u);
              umma_tf32(d, dah, dbl, idesc, 1u);
            }
          }
          umma_commit(&bar_free[s]);
          if (c == n_chunks - 1) umma_commit(&bar_acc);
        }
#pragma unroll
        for (int j = 0; j < MT * 4; ++j) ra[j] = na[j];
#pragma unroll
        for (int j = 0; j < FC_TC_MAXN / 32; ++j) rb[j] = nb[j];
      }
      // ---- epilogue of this (item, row block): TMEM -> registers -> global
      mbar_wait(&bar_acc, acc_phase);
      acc_phase ^= 1;
      tc_fence_after();
      {
        const int mt = warp >> 2;  // warps 0-3: tile 0, warps 4-7: tile 1
        if (mt < MT) {
          const int row = (warp & 3) * 32 + lane;  // TMEM lane = accumulator row
          float* drow = a.D + (item * (int64_t)a.O + mt0 + mt * 128 + row) * N;
          for (int c0 = 0; c0 < N; c0 += 32) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(mt * N + c0), v);
#pragma unroll
            for (int q = 0; q < 32; q += 4)
              if (c0 + q < N) *reinterpret_cast<float4*>(drow + c0 + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
          }
        }
      }
      tc_fence_before();
      __syncthreads();  // the accumulator may be overwritten by the next item's first MMA
      tc_fence_after();
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}
#endif  // !FC_CPU_EMUL

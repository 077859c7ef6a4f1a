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
#include "fc_fused.cuh"

// ------------------------------------------------------------------------------------------------ relayouts
// The GEMM operands live in HBM as the exact shared-memory images of the K-chunks the tensor core consumes
// ("blobs"): a chunk is 32 consecutive fp32 of the K dimension (one 128-byte row per matrix row), rows in groups of 8
// (1024 bytes), the eight 16-byte pieces of a row XOR-swizzled with the row index (UMMA SWIZZLE_128B, K-major). A
// blob is contiguous, so one bulk async copy (cp.async.bulk) moves it into its pipeline stage at full DRAM efficiency.
//   A blobs : [bin][group][tile][chunk] x (128 rows x 32 fp32)             kernel spectrum, rows = output channels (128 per tile;
//             64 per tile when Cout/g is a multiple of 64 only: the GEMM then zero-fills the upper half of its 128-row operand)
//   Bt blobs: [bin][group][chunk] x (N rows x 32 fp32)                     signal spectrum, rows = (batch, re/im); the
//             raw values (the tensor core ignores the low 13 mantissa bits = the "hi" operand); like A's, the "lo" operand
//             (value - truncated value) is derived in shared memory by the GEMM kernel, so it never crosses HBM
// One CTA of the relayout kernels moves a 32 x 32 (rows x bins) tile through shared memory so that the bin-innermost
// side is coalesced; mode 2 brings the product D[bin][o][2*Bp] back to the pass-order layout [(b*Cout + o)][bin].
FC_DEV int64_t fc_tc_swz_off(int r, int j) {  // float offset of (row r, fp32 column j < 32) inside a blob
  return (int64_t)(r >> 3) * 256 + (r & 7) * 32 + ((((j >> 2) ^ (r & 7)) << 2) | (j & 3));
}
FC_DEV float fc_tc_lo(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

struct fc_tc_relayout_args {
  const float2* in;
  float* out;
  int64_t bins;
  int32_t I;     // input channels per group
  int32_t C;     // all input channels (G*I)
  int32_t rows;  // mode 0: Cout*I, mode 1: B*Cin, mode 2: Cout*Bp (o-major, b inner)
  int32_t O;     // all output channels (G*Og)
  int32_t Og;    // output channels per group
  int32_t B;     // real batch
  int32_t Bp;    // padded batch (N = 2*Bp rows of Bt per bin); rows b >= B must be zero-filled by the caller
  int32_t mode;
};

__global__ void fc_tc_relayout_kernel(fc_tc_relayout_args a) {
  fc_grid_dep_sync();
  __shared__ float2 tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads
  const int64_t f0 = (int64_t)blockIdx.x * 32;
  const int r0 = blockIdx.y * 32;
  const int KD = 2 * a.I, n_chunks = KD / 32;
  if (a.mode != 2) {
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
        const int g = o / a.Og, og = o - g * a.Og;
        const int G = a.O / a.Og, prow = a.Og % 128 == 0 ? 128 : 64, passes = a.Og / prow;  // one blob per tile of 128 (or 64) output rows
        const int pass = og / prow, row = og - pass * prow;
        const int64_t blob = (((f * G + g) * passes + pass) * n_chunks);
        // K column i holds re, column I + i holds im
        a.out[(blob + (i >> 5)) * (prow * 32) + fc_tc_swz_off(row, i & 31)] = v.x;
        a.out[(blob + ((a.I + i) >> 5)) * (prow * 32) + fc_tc_swz_off(row, (a.I + i) & 31)] = v.y;
      } else {
        const int b = r / a.C, c = r - b * a.C;
        const int g = c / a.I, i = c - g * a.I;
        const int G = a.C / a.I, N = 2 * a.Bp;
        const int64_t blob = (f * G + g) * n_chunks;
        // rows 2b (-> Re Y) and 2b+1 (-> Im Y); columns i and I + i
        const float vals[4] = {v.x, -v.y, v.y, v.x};
        const int rows[4] = {2 * b, 2 * b, 2 * b + 1, 2 * b + 1};
        const int cols[4] = {i, a.I + i, i, a.I + i};
#pragma unroll
        for (int e = 0; e < 4; ++e) a.out[(blob + (cols[e] >> 5)) * (N * 32) + fc_tc_swz_off(rows[e], cols[e] & 31)] = vals[e];
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

// ------------------------------------------------------------------------------------------------ transforms fused with the relayouts
// The relayout kernels above cost a full round trip of the spectra through HBM on either side of the GEMM (BASELINE c4:
// 1.85 of 6.2 ms). When the neighbouring passes are contiguous complex transforms (the second pass of the four-step 1-d
// layout) they do the relayout themselves:
//   fc_tc_c2c_fwd_kernel  last forward pass: a tile is the lines of 32 consecutive input channels of one (batch item, line
//                         index r); after the transforms the tile is read transposed (one bin, 32 channels = one 128-byte
//                         K-chunk row) and goes straight into the Bt blobs: four 128-byte runs per bin (re / -im / im / re).
//   fc_tc_c2c_inv_kernel  first inverse pass: a tile is the lines of 32 consecutive batch items of one (output channel,
//                         line index r), gathered from the GEMM's product D[bin][o][Bp] in 256-byte runs (one bin, 32 batch
//                         items), transformed and stored as contiguous lines.
// Both reuse the warp engine of fc_fused.cuh (two lines per warp in registers, warp-private exchange lines inside the tile).
struct fc_tc_c2c_args {
  const float2* in;  // fwd: lines [(b*C + c)][R][N]; inv: D [R*N bins][O][Bp]
  float2* out;       // fwd: Bt blobs (as float*); inv: lines [(b*O + o)] at out_os, line r at out_rs
  const float2* tw;
  int32_t tw_len;
  int32_t B, Bp;     // batch items of this GEMM chunk and their padded count (N = 2*Bp rows of Bt)
  int32_t C, I, G;   // channels of the tile axis (fwd: input channels, C = G*I; inv: C = all output channels)
  int32_t R;         // lines per (batch, channel) item
  int64_t os, rs;    // item and line stride (complex elements) of the line-ordered side
  int64_t n_tiles;
};

#ifndef FC_CPU_EMUL
template <int N>
__global__ void __launch_bounds__(256, N == 256 ? 3 : 1) fc_tc_c2c_fwd_kernel(fc_tc_c2c_args a) {
  fc_grid_dep_sync();
  constexpr int E = N / 32, LP = N + 1, NL = 2;
  FC_DYN_SMEM(smem);  // 32 lines of pitch N + 1 (odd: the transposed read of the store phase is conflict-free)
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  fc_wofs ofs;
  ofs.init(lane);
  const int cblks = a.C / 32, n_chunks = 2 * a.I / 32, Np = 2 * a.Bp;
  float* xtc = reinterpret_cast<float*>(a.out);
  for (int64_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    const int cblk = (int)(t % cblks);
    const int64_t t1 = t / cblks;
    const int r = (int)(t1 % a.R), b = (int)(t1 / a.R);
#pragma unroll
    for (int round = 0; round < 2; ++round) {
      const int l0 = round * 16 + NL * w;
      float2* line0 = smem + l0 * LP;
      float2 v[NL][E];
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        const float2* src = a.in + ((int64_t)b * a.C + cblk * 32 + l0 + l) * a.os + (int64_t)r * a.rs;
#pragma unroll
        for (int q = 0; q < E; ++q) v[l][q] = fc_ld_stream(src + lane + 32 * q);
      }
      fc_wfft<N, NL, LP>(v, line0, ofs, a.tw, a.tw_len, lane);
      __syncwarp();
#pragma unroll
      for (int l = 0; l < NL; ++l)
#pragma unroll
        for (int q = 0; q < E; ++q) line0[l * LP + lane + 32 * q] = v[l][q];
    }
    __syncthreads();
    {  // lane = channel of the block; K columns: channel i of the group -> column i (re block) and I + i (im block)
      const int c0 = cblk * 32, g = c0 / a.I, i0 = c0 - g * a.I;
      const int ch_re = i0 >> 5, ch_im = (a.I + i0) >> 5;
      const int row0 = 2 * b, row1 = 2 * b + 1;
      const int off0 = (row0 >> 3) * 256 + (row0 & 7) * 32 + ((((lane >> 2) ^ (row0 & 7)) << 2) | (lane & 3));
      const int off1 = (row1 >> 3) * 256 + (row1 & 7) * 32 + ((((lane >> 2) ^ (row1 & 7)) << 2) | (lane & 3));
      const int64_t blob = (int64_t)Np * 32;  // floats per Bt blob
      for (int j = w; j < N; j += 8) {
        const float2 z = smem[lane * LP + j];
        const int64_t f = (int64_t)r * N + j;
        float* bre = xtc + ((f * a.G + g) * n_chunks + ch_re) * blob;
        float* bim = xtc + ((f * a.G + g) * n_chunks + ch_im) * blob;
        bre[off0] = z.x;   // row 2b   = [ Xr | -Xi ]
        bim[off0] = -z.y;
        bre[off1] = z.y;   // row 2b+1 = [ Xi |  Xr ]
        bim[off1] = z.x;
      }
    }
    __syncthreads();
  }
}

template <int N>
__global__ void __launch_bounds__(256, N == 256 ? 3 : 1) fc_tc_c2c_inv_kernel(fc_tc_c2c_args a) {
  fc_grid_dep_sync();
  constexpr int E = N / 32, LP = N + 1, NL = 2;
  FC_DYN_SMEM(smem);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  fc_wofs ofs;
  ofs.init(lane);
  const int bblks = (a.B + 31) / 32;
  for (int64_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    const int bblk = (int)(t % bblks);
    const int64_t t1 = t / bblks;
    const int o = (int)(t1 % a.C), r = (int)(t1 / a.C);
    const int b = bblk * 32 + lane;
    {  // gather: bin j of the 32 batch items = one 256-byte run of D
      const float2* d = a.in + (((int64_t)r * N) * a.C + o) * a.Bp + b;
      const int64_t dstep = (int64_t)a.C * a.Bp;
      for (int j = w; j < N; j += 8) {
        const float2 z = b < a.B ? fc_ld_stream(d + j * dstep) : make_float2(0.f, 0.f);
        smem[lane * LP + j] = fc_conj(z);
      }
    }
    __syncthreads();
#pragma unroll
    for (int round = 0; round < 2; ++round) {
      const int l0 = round * 16 + NL * w;
      float2* line0 = smem + l0 * LP;
      float2 v[NL][E];
#pragma unroll
      for (int l = 0; l < NL; ++l)
#pragma unroll
        for (int q = 0; q < E; ++q) v[l][q] = line0[l * LP + lane + 32 * q];
      __syncwarp();
      fc_wfft<N, NL, LP>(v, line0, ofs, a.tw, a.tw_len, lane);
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        const int bb = bblk * 32 + l0 + l;
        if (bb >= a.B) continue;
        float2* dst = a.out + ((int64_t)bb * a.C + o) * a.os + (int64_t)r * a.rs;
#pragma unroll
        for (int q = 0; q < E; ++q) dst[lane + 32 * q] = fc_conj(v[l][q]);
      }
    }
    __syncthreads();
  }
}
#endif

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
FC_DEV void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
FC_DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-d bulk async copy global -> shared (TMA engine, no tensor map); completion is counted in bytes on `bar`.
FC_DEV void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
FC_DEV void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
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
  const float* A;   // A blobs  [item][tile][chunk][128 x 32]
  const float* Bt;  // Bt blobs [item][chunk][N x 32]
  float* D;         // [item][O][N]   product (complex Y[f][g][o][b])
  int64_t n_items;  // bins * G
  int32_t O, I, B;  // per group; O % 64 == 0, (2I) % 32 == 0, B = padded batch (multiple of 8): N = 2B <= 160 (MT = 1) or <= 32 (MT = 2)
  int32_t tile_rows;  // output rows per A tile: 128, or 64 (O % 128 != 0; MT = 1): the MMA still runs M = 128 on a zero upper half
};

// One CTA per SM, persistent over (bin, group) items; warp-specialised:
//   producer (warp 8, one lane): for every chunk, waits until the stage is free and issues the bulk async copies
//     (MT A blobs, one Bt blob) that complete on the stage's "full" mbarrier;
//   splitters (warps 0-7): wait "full", derive the low parts of A and Bt in shared memory, fence to the async proxy and
//     arrive on the stage's "ready" mbarrier; after the last chunk of an item they drain the accumulator
//     (warps 0-3 / 4-7: the two 128-row tiles) from TMEM straight to global memory and release it ("acc_free");
//   MMA issuer (warp 9, one lane): waits "ready", issues the 3 x 4 x MT tcgen05.mma of the chunk into one of two
//     accumulator sets in TMEM, commits them to the stage's "free" mbarrier and, on the last chunk, to "acc_full".
// Variants: <MT, STAGES, NBUF> = tiles of 128 output rows per pass (the Bt chunk is fetched once per pass), operand stages
// in shared memory, accumulator sets in TMEM:
//   <2, 3, 2>  N <= 32            (BASELINE c4 unsegmented, 16 batch items)
//   <1, 3, 2>  N <= 160           (one tile per pass: Bt crosses L2 -> SM once per 128 output rows)
//   <2, 2, 1>  N <= 160, O % 256  (two tiles per pass on 2 x 104 KB stages and ONE 320-column accumulator set: the issuer
//                                  waits for the drain of a pass before the next one starts)
#define FC_TC_THREADS 320
template <int MT /* 128-row tiles of O per pass: 1 or 2 */, int FC_TC_STAGES = 3, int NBUF = 2>
__global__ void __launch_bounds__(FC_TC_THREADS, 1) fc_tc_gemm_kernel(fc_tc_args a) {
  using namespace fc_tc;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = 2 * a.B, KD = 2 * a.I, n_chunks = KD / 32;
  constexpr int A_BYTES = MT * 128 * 128;  // one copy of the A chunk: MT*128 rows x 128 bytes
  const int B_BYTES = N * 128;             // one copy of the Bt chunk
  const int stage_bytes = 2 * A_BYTES + 2 * B_BYTES;
  unsigned char* sbase = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar_full[FC_TC_STAGES];   // bulk copies of the stage have landed      (tx bytes)
  __shared__ __align__(8) uint64_t bar_ready[FC_TC_STAGES];  // low parts written, stage ready for the MMAs (256 arrivals)
  __shared__ __align__(8) uint64_t bar_free[FC_TC_STAGES];   // MMAs reading the stage are done             (commit)
  __shared__ __align__(8) uint64_t bar_acc_full[2];          // accumulator set complete                    (commit)
  __shared__ __align__(8) uint64_t bar_acc_free[2];          // accumulator set drained                     (256 arrivals)
  __shared__ uint32_t tmem_slot;
  if (tid == 0) {
    for (int s = 0; s < FC_TC_STAGES; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_ready[s], 256);
      mbar_init(&bar_free[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_acc_full[s], 1);
      mbar_init(&bar_acc_free[s], 256);
    }
    fence_barrier_init();
  }
  const uint32_t acc_cols = (uint32_t)(MT * N);  // columns of one accumulator set
  const uint32_t all_cols = NBUF * acc_cols;
  const uint32_t tmem_cols = (all_cols <= 32) ? 32 : (all_cols <= 64 ? 64 : (all_cols <= 128 ? 128 : (all_cols <= 256 ? 256 : 512)));
  if (warp == 0) tmem_alloc(&tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t idesc = idesc_tf32(128, N);

  const int TR = a.tile_rows;  // 128, or 64 with MT == 1
  const uint32_t a_load = (uint32_t)TR * 128u;  // bytes of one A tile in HBM
  if (TR < 128) {  // rows TR..127 of every stage's A operand (raw and low part) stay zero: they are never written again
    for (int s = 0; s < FC_TC_STAGES; ++s)
      for (int e = tid; e < (128 - TR) * 8 * 2; e += FC_TC_THREADS) {
        unsigned char* half = sbase + (size_t)s * stage_bytes + (e >= (128 - TR) * 8 ? A_BYTES : 0) + (size_t)TR * 128;
        *reinterpret_cast<float4*>(half + ((e % ((128 - TR) * 8)) << 4)) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    fence_proxy_async();
    __syncthreads();
  }
  const int passes = a.O / (MT * TR);
  const int64_t my_items = (a.n_items > (int64_t)blockIdx.x) ? (a.n_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t n_acc = my_items * passes;       // accumulator uses (item, pass) of this CTA
  const int64_t total = n_acc * n_chunks;        // flat chunk sequence q = acc_use * n_chunks + c

  if (warp == 8) {
    // ---------------- producer
    if (lane == 0) {
      for (int64_t q = 0; q < total; ++q) {
        const int s = (int)(q % FC_TC_STAGES);
        const int64_t use = q / FC_TC_STAGES;
        if (use > 0) mbar_wait(&bar_free[s], (uint32_t)((use - 1) & 1));
        const int c = (int)(q % n_chunks);
        const int64_t ip = q / n_chunks;
        const int pass = (int)(ip % passes);
        const int64_t item = blockIdx.x + (ip / passes) * gridDim.x;
        unsigned char* st = sbase + (size_t)s * stage_bytes;
        mbar_expect_tx(&bar_full[s], (uint32_t)(MT * a_load + B_BYTES));
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)  // the tiles of this pass are n_chunks blobs apart
          bulk_g2s(st + mt * 16384, a.A + ((((item * passes + pass) * MT + mt) * n_chunks + c) * (int64_t)(a_load / 4)), a_load, &bar_full[s]);
        bulk_g2s(st + 2 * A_BYTES, a.Bt + ((item * n_chunks + c) * (int64_t)(B_BYTES / 4)), B_BYTES, &bar_full[s]);
      }
    }
  } else if (warp == 9) {
    // ---------------- MMA issuer
    if (lane == 0) {
      for (int64_t q = 0; q < total; ++q) {
        const int c = (int)(q % n_chunks);
        const int64_t ip = q / n_chunks;  // accumulator use
        const int buf = (int)(ip % NBUF);
        const int s = (int)(q % FC_TC_STAGES);
        if (c == 0 && ip >= NBUF) mbar_wait(&bar_acc_free[buf], (uint32_t)(((ip / NBUF) - 1) & 1));  // set drained by the epilogue
        mbar_wait(&bar_ready[s], (uint32_t)((q / FC_TC_STAGES) & 1));
        tc_fence_after();
        unsigned char* st = sbase + (size_t)s * stage_bytes;
        const uint32_t ah = smem_u32(st), al = ah + A_BYTES, bh = ah + 2 * A_BYTES, bl = bh + B_BYTES;
        const uint64_t dah0 = smem_desc_sw128(ah), dal0 = smem_desc_sw128(al), dbh0 = smem_desc_sw128(bh), dbl0 = smem_desc_sw128(bl);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const uint32_t d = tmem_base + (uint32_t)buf * acc_cols + (uint32_t)(mt * N);
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 4 x (K = 8 tf32 = 32 bytes) per 128-byte row; the start address field is in 16-byte units
            const uint64_t adv_a = (uint64_t)((mt * 16384 + k * 32) >> 4), adv_b = (uint64_t)((k * 32) >> 4);
            umma_tf32(d, dah0 + adv_a, dbh0 + adv_b, idesc, (c | k) ? 1u : 0u);
            umma_tf32(d, dal0 + adv_a, dbh0 + adv_b, idesc, 1u);
            umma_tf32(d, dah0 + adv_a, dbl0 + adv_b, idesc, 1u);
          }
        }
        umma_commit(&bar_free[s]);
        if (c == n_chunks - 1) umma_commit(&bar_acc_full[buf]);
      }
    }
  } else {
    // ---------------- splitters + epilogue (256 threads)
    const int piece = tid & 7;
    for (int64_t q = 0; q < total; ++q) {
      const int c = (int)(q % n_chunks);
      const int s = (int)(q % FC_TC_STAGES);
      unsigned char* st = sbase + (size_t)s * stage_bytes;
      unsigned char* a_hi = st;
      unsigned char* a_lo = st + A_BYTES;
      mbar_wait(&bar_full[s], (uint32_t)((q / FC_TC_STAGES) & 1));
      // low part of A: thread t owns 16-byte piece (t & 7) of rows t>>3 + 32*j (the blob is already swizzled; the
      // low part keeps the same positions)
#pragma unroll
      for (int j = 0; j < MT * 4; ++j) {
        if (TR < 128 && j >= TR / 32) break;  // (the zero upper half keeps its zero low part)
        const int row = (tid >> 3) + 32 * j;
        const uint32_t off = (uint32_t)row * 128 + (uint32_t)(piece << 4);
        const float4 v = *reinterpret_cast<const float4*>(a_hi + off);
        *reinterpret_cast<float4*>(a_lo + off) = make_float4(fc_tc_lo(v.x), fc_tc_lo(v.y), fc_tc_lo(v.z), fc_tc_lo(v.w));
      }
      {  // low part of Bt: N rows x eight 16-byte pieces, same positions
        unsigned char* b_hi = st + 2 * A_BYTES;
        unsigned char* b_lo = b_hi + B_BYTES;
        for (int pc = tid; pc < N * 8; pc += 256) {
          const float4 v = *reinterpret_cast<const float4*>(b_hi + (pc << 4));
          *reinterpret_cast<float4*>(b_lo + (pc << 4)) = make_float4(fc_tc_lo(v.x), fc_tc_lo(v.y), fc_tc_lo(v.z), fc_tc_lo(v.w));
        }
      }
      fence_proxy_async();  // generic-proxy writes of a_lo / b_lo -> visible to the tensor core (async proxy)
      mbar_arrive(&bar_ready[s]);
      if (c == n_chunks - 1) {
        // ---- epilogue of this (item, pass): TMEM -> registers -> global
        const int64_t ip = q / n_chunks;
        const int buf = (int)(ip % NBUF);
        const int pass = (int)(ip % passes);
        const int64_t item = blockIdx.x + (ip / passes) * gridDim.x;
        mbar_wait(&bar_acc_full[buf], (uint32_t)((ip / NBUF) & 1));
        tc_fence_after();
        const int mt = warp >> 2;  // warps 0-3: tile 0, warps 4-7: tile 1
        const int row = (warp & 3) * 32 + lane;  // TMEM lane = accumulator row
        if (mt < MT && row < TR) {
          float* drow = a.D + (item * (int64_t)a.O + (int64_t)pass * MT * TR + mt * TR + row) * N;
          for (int c0 = 0; c0 < N; c0 += 32) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)buf * acc_cols + (uint32_t)(mt * N + c0), v);
#pragma unroll
            for (int qq = 0; qq < 32; qq += 4)
              if (c0 + qq < N) *reinterpret_cast<float4*>(drow + c0 + qq) = make_float4(v[qq], v[qq + 1], v[qq + 2], v[qq + 3]);
          }
        }
        tc_fence_before();
        mbar_arrive(&bar_acc_free[buf]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}
#endif  // !FC_CPU_EMUL

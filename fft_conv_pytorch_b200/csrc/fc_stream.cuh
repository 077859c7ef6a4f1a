// fc_stream.cuh — K1 / K4 for 512-point real rows as streaming kernels on the bulk-copy (TMA) engine of sm_100a.
//
// The register-path K1 / K4 of fc_fused.cuh are phase-additive (profiles/r2_k1_phase_ablation.txt: row loads, transform
// and transposed store cost 6 us each on top of an 18 us skeleton at BASELINE c2) because every byte moves through the
// warps' load / store instructions and the tile loop is barrier-phased. Here the warps only transform:
//
//   fc_stream_r2c_kernel  K1: a 16-row tile (32 KB, contiguous in HBM) arrives by cp.async.bulk into a two-deep ring and
//                         signals an mbarrier; the warps transform their two rows with the rows' own ring slot as exchange
//                         buffer, write the half spectrum into a [bin][16 rows] tile in the 128-byte-swizzle layout
//                         (one conflict-free STS.128 per bin and row pair), and ONE thread hands that tile to the tensor-map
//                         store (cp.async.bulk.tensor, box 16 rows x 256 bins: the transposition is done by the copy engine
//                         while the warps are already in the next tile).
//   fc_stream_c2r_kernel  K4: the [bin][16 rows] tile is fetched by one tensor-map load per tile (same swizzle), the warps
//                         pick up their bins and Hermitian partners with LDS.128, run the inverse transform and store the
//                         cropped real rows straight from registers (coalesced 256-byte runs).
//
// Scope: M = 256 (512-point rows), no overlap-save segments, identity / zero-padding gather with 16-byte aligned rows on
// the way in, plain crop on the way out; everything else stays on fc_fast_r2c_kernel / fc_fast_c2r_kernel (fc_api.cu
// decides). Same arithmetic, in the same order, as those kernels: results are bit-identical.
#pragma once
#ifndef FC_CPU_EMUL
#include <cuda.h>  // CUtensorMap (the encoder is fetched through cudaGetDriverEntryPoint in fc_api.cu: no libcuda link)

#include "fc_fused.cuh"
#include "fc_tc.cuh"

namespace fc_stream {

using fc_tc::smem_u32;

// 3-d tensor-map copies between a [x = 32 floats][y = 256 bins][z = 1 item] box and shared memory.
FC_DEV void tma_store_3d(const CUtensorMap* tm, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tm), "r"(smem_u32(src)), "r"(c0),
               "r"(c1), "r"(c2)
               : "memory");
}
FC_DEV void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
FC_DEV void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
FC_DEV void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
FC_DEV void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
FC_DEV void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
FC_DEV void prefetch_tmap(const CUtensorMap* tm) { asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory"); }

constexpr int kTR = 16;          // rows per tile
constexpr int kTileBytes = 32768;  // 16 rows x 256 float2 = 256 bins x 16 rows x 8 bytes
constexpr int kNyqBytes = kTR * 8;
// dynamic shared memory: two 32 KB ring slots (1024-byte aligned: the swizzle is a function of the address), the Nyquist
// bins, the mbarriers. A slot is, in turn, the landing buffer of a bulk copy, the exchange buffer of the warps' transforms
// and (K1) the source of the tensor-map store: 66 KB per CTA, three CTAs (24 warps) per SM.
constexpr int smem_bytes(int ns) { return 1024 + ns * (kTileBytes + kNyqBytes) + 64; }

}  // namespace fc_stream

struct fc_stream_r2c_args {
  fc_pass p;
  const float* x;
  float2* out;
  const float2* tw;
  int32_t whole_tiles;  // 1: every tile is one contiguous run of 16 full rows (one bulk copy); 0: one copy of L floats per row
  CUtensorMap tmap;     // out as [item][bin][2*row] floats, box 32 x 256 x 1, 128-byte swizzle
};

#ifndef FC_STREAM_ABL
#define FC_STREAM_ABL 0  // timing experiments (scripts/micro/stream_bench.cu): 1 = no transform, 2 = no store, 4 = no load wait
#endif

// grid: persistent, <= 3 CTAs per SM; 256 threads; warp w owns tile rows 2w, 2w + 1.
// Life of ring slot s = it % 2 for the CTA's it-th tile:
//   bulk copy lands (mbarrier full[s]) -> every warp reads its two rows into registers and uses them as exchange buffer
//   -> barrier -> the warps write the half spectrum into the slot in the [bin][row] swizzle layout -> barrier -> thread 0
//   issues the tensor-map store -> (top of the next iteration) thread 0 waits until the store has read the slot and
//   requests tile it + 2 into it: that copy has the whole of tile it + 1 to land.
template <int NS>
__global__ void __launch_bounds__(256, NS == 2 ? 3 : 2) fc_stream_r2c_kernel(const __grid_constant__ fc_stream_r2c_args a) {
  using namespace fc_stream;
  fc_grid_dep_sync();
  constexpr int M = 256, NL = 2, E = 8;
  const fc_pass& p = a.p;
  extern __shared__ float4 fc_dyn_smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fc_dyn_smem_raw) + 1023) & ~(uintptr_t)1023);
  float4* otn = reinterpret_cast<float4*>(base + NS * kTileBytes);  // Nyquist bin of the 16 rows
  uint64_t* full = reinterpret_cast<uint64_t*>(base + NS * (kTileBytes + kNyqBytes));
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int lrow = NL * w;
  const int L = p.imap.L, pad = p.imap.pad;
  const int tstep = p.tw_len / (2 * M);
  const int tpo = (int)p.tiles_per_outer, n_tiles = (int)p.n_tiles, R = (int)p.R;
  const int G = gridDim.x;
  fc_wofs ofs;
  ofs.init(lane);
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) fc_tc::mbar_init(&full[s], 1);
    fc_tc::fence_barrier_init();
    prefetch_tmap(&a.tmap);
  }
  __syncthreads();
  // warp 0 requests tile t into ring slot s
  auto request = [&](int t, int s) {
    const int o = t / tpo;
    const int r0 = (t - o * tpo) * kTR;
    const float* img = a.x + (int64_t)o * p.o_sA + (int64_t)r0 * p.in_rs;
    float2* dst = reinterpret_cast<float2*>(base + s * kTileBytes);
    if (a.whole_tiles) {
      if (lane == 0) {
        fc_tc::mbar_expect_tx(&full[s], kTileBytes);
        fc_tc::bulk_g2s(dst, img, kTileBytes, &full[s]);
      }
    } else {
      const int rows = R - r0 < kTR ? R - r0 : kTR;
      if (lane == 0) fc_tc::mbar_expect_tx(&full[s], (uint32_t)(rows * L * 4));
      __syncwarp();
      if (lane < rows) fc_tc::bulk_g2s(reinterpret_cast<float*>(dst + lane * M) + pad, img + (int64_t)lane * p.in_rs, (uint32_t)(L * 4), &full[s]);
    }
  };
  if (w == 0) {
    if ((int)blockIdx.x < n_tiles) request(blockIdx.x, 0);
    if ((int)blockIdx.x + G < n_tiles) request(blockIdx.x + G, 1);
  }
  // swizzled position of (bin lane + 32q, row pair w): 128 bytes per bin, the 16-byte chunk index XORed with bin % 8
  const uint32_t ot_off = (uint32_t)lane * 128u + (uint32_t)((w ^ (lane & 7)) << 4);
  int it = 0, s = 0, ph = 0;  // slot and mbarrier phase of tile it: s = it % NS, ph = (it / NS) & 1
  for (int t = blockIdx.x; t < n_tiles; t += G, ++it) {
    const int o = t / tpo;
    const int r0 = (t - o * tpo) * kTR;
    uint8_t* slot = base + s * kTileBytes;
    float2* line0 = reinterpret_cast<float2*>(slot) + lrow * M;
    if (NS == 2 && w == 0 && it >= 1) {  // the other slot: its store (tile it - 1) must have read it before tile it + 1 may land there
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
      if (t + G < n_tiles) request(t + G, s ^ 1);
    }
    if (!(FC_STREAM_ABL & 4)) fc_tc::mbar_wait(&full[s], ph);
    float2 v[NL][E];
#pragma unroll
    for (int l = 0; l < NL; ++l)
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const int u = 2 * (lane + 32 * q);  // dense positions u, u + 1 <- source u - pad (pad and L are even)
        v[l][q] = (unsigned)(u - pad) < (unsigned)L ? line0[l * M + lane + 32 * q] : make_float2(0.f, 0.f);
      }
    float nyq[NL] = {0.f, 0.f};
    if (!(FC_STREAM_ABL & 1)) {
      __syncwarp();  // the rows become the warp's exchange buffers
      fc_wfft<M, NL, M>(v, line0, ofs, a.tw, p.tw_len, lane);
      fc_wwrite<M, NL, M>(v, line0, ofs);
      __syncwarp();
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
    }
    __syncthreads();  // every warp is done with the slot as row / exchange buffer: it takes the spectrum tile
    if (NS == 3 && w == 0) {  // three slots: the one tile it - 1 was stored from (a whole transform ago) takes tile it + 2
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
      if (t + 2 * G < n_tiles) request(t + 2 * G, s == 0 ? 2 : s - 1);
    }
#pragma unroll
    for (int q = 0; q < E; ++q)
      *reinterpret_cast<float4*>(slot + ot_off + q * 4096) = make_float4(v[0][q].x, v[0][q].y, v[1][q].x, v[1][q].y);
    if (lane == 0) otn[s * (kNyqBytes / 16) + w] = make_float4(nyq[0], 0.f, nyq[1], 0.f);
    fc_tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0 && !(FC_STREAM_ABL & 2)) {
      tma_store_3d(&a.tmap, slot, 2 * r0, 0, o);
      const int rows = R - r0 < kTR ? R - r0 : kTR;
      bulk_s2g(a.out + (int64_t)o * p.out_os + r0 + (int64_t)M * p.out_es, otn + s * (kNyqBytes / 16), (uint32_t)(rows * 8));
      bulk_commit();
    }
    if (++s == NS) s = 0, ph ^= 1;
  }
  if (tid == 0) bulk_wait0();
}

struct fc_stream_c2r_args {
  fc_pass p;
  const float2* in;
  float* out;
  const float2* tw;
  const float* bias;
  CUtensorMap tmap;  // in as [item][bin][2*row] floats, box 32 x 256 x 1, 128-byte swizzle
};

// Life of ring slot s = it % 2: the tensor-map load of the [bin][row] tile lands (full[s]) -> every warp picks up its bins and
// their Hermitian partners -> barrier -> the slot is the exchange buffer of the warps' inverse transforms; a warp that is
// through arrives on done[s] -> (top of the next iteration) thread 0 waits for the eight arrivals and requests tile it + 2.
template <int NS>
__global__ void __launch_bounds__(256, NS == 2 ? 3 : 2) fc_stream_c2r_kernel(const __grid_constant__ fc_stream_c2r_args a) {
  using namespace fc_stream;
  fc_grid_dep_sync();
  constexpr int M = 256, NL = 2, E = 8;
  const fc_pass& p = a.p;
  extern __shared__ float4 fc_dyn_smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fc_dyn_smem_raw) + 1023) & ~(uintptr_t)1023);
  float4* nq = reinterpret_cast<float4*>(base + NS * kTileBytes);  // ring: Nyquist bin of the 16 rows
  uint64_t* full = reinterpret_cast<uint64_t*>(base + NS * (kTileBytes + kNyqBytes));
  uint64_t* done = full + NS;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int lrow = NL * w;
  const int tstep = p.tw_len / (2 * M);
  const int Lout = p.omap.Lout;
  const int tpo = (int)p.tiles_per_outer, n_tiles = (int)p.n_tiles, R = (int)p.R;
  const int G = gridDim.x;
  fc_wofs ofs;
  ofs.init(lane);
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      fc_tc::mbar_init(&full[s], 1);
      fc_tc::mbar_init(&done[s], 8);
    }
    fc_tc::fence_barrier_init();
    prefetch_tmap(&a.tmap);
  }
  __syncthreads();
  auto request = [&](int t, int s) {  // one thread
    const int o = t / tpo;
    const int r0 = (t - o * tpo) * kTR;
    const int rows = R - r0 < kTR ? R - r0 : kTR;
    fc_tc::mbar_expect_tx(&full[s], (uint32_t)(kTileBytes + rows * 8));
    tma_load_3d(base + s * kTileBytes, &a.tmap, 2 * r0, 0, o, &full[s]);
    fc_tc::bulk_g2s(nq + s * (kNyqBytes / 16), a.in + (int64_t)o * p.in_os + r0 + (int64_t)M * p.in_es, (uint32_t)(rows * 8), &full[s]);
  };
  if (tid == 0) {
    for (int s = 0; s < NS; ++s)
      if ((int)blockIdx.x + s * G < n_tiles) request(blockIdx.x + s * G, s);
  }
  // swizzled positions of (bin k, row pair w) for k = lane + 32q and of the partner bin M - k = (M - lane) - 32q
  const uint32_t off_k = (uint32_t)lane * 128u + (uint32_t)((w ^ (lane & 7)) << 4);
  const uint32_t off_m = (uint32_t)(M - lane) * 128u + (uint32_t)((w ^ ((M - lane) & 7)) << 4);  // lane 0, q = 0: the Nyquist buffer instead
  int it = 0, s = 0, ph = 0;  // slot and mbarrier phase of tile it: s = it % NS, ph = (it / NS) & 1
  for (int t = blockIdx.x; t < n_tiles; t += G, ++it) {
    const int o = t / tpo;
    const int r0 = (t - o * tpo) * kTR;
    uint8_t* tile = base + s * kTileBytes;
    float2* line0 = reinterpret_cast<float2*>(tile) + lrow * M;
    if (tid == 0 && it >= 1 && t + (NS - 1) * G < n_tiles) {  // the slot of tile it - 1 is free once every warp is through that tile
      const int sp = s == 0 ? NS - 1 : s - 1;
      fc_tc::mbar_wait(&done[sp], s == 0 ? ph ^ 1 : ph);
      request(t + (NS - 1) * G, sp);
    }
    if (!(FC_STREAM_ABL & 4)) fc_tc::mbar_wait(&full[s], ph);
    const float b = p.has_bias ? __ldg(a.bias + (o % p.cout)) : 0.f;
    float2 v[NL][E];
#pragma unroll
    for (int q = 0; q < E; ++q) {
      const int k = lane + 32 * q;
      const float4 yk = *reinterpret_cast<const float4*>(tile + off_k + q * 4096);
      const float4* pm = (q == 0 && lane == 0) ? nq + s * (kNyqBytes / 16) + w : reinterpret_cast<const float4*>(tile + off_m - q * 4096);
      const float4 ym = *pm;
      const float2 wk = fc_conj(__ldg(a.tw + k * tstep));
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        const float2 zk = l == 0 ? make_float2(yk.x, yk.y) : make_float2(yk.z, yk.w);
        const float2 zm = fc_conj(l == 0 ? make_float2(ym.x, ym.y) : make_float2(ym.z, ym.w));
        const float2 sm = fc_add(zk, zm);
        const float2 d = fc_mul(fc_sub(zk, zm), wk);
        v[l][q] = make_float2(sm.x - d.y, -(sm.y + d.x));
      }
    }
    __syncthreads();  // the tile is in registers: the slot becomes the exchange buffer of the warps
    if (!(FC_STREAM_ABL & 1)) fc_wfft<M, NL, M>(v, line0, ofs, a.tw, p.tw_len, lane);
    fc_tc::fence_proxy_async();  // generic accesses to the slot before the async-proxy write of the next request
    __syncwarp();
    if (lane == 0) fc_tc::mbar_arrive(&done[s]);
    if (!(FC_STREAM_ABL & 2)) {
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        const int r = r0 + lrow + l;
        if (r < R) {
          float* yrow = a.out + (int64_t)o * p.out_os + (int64_t)r * p.out_rs;
#pragma unroll
          for (int q = 0; q < E; ++q) {
            const int n0 = 2 * (lane + 32 * q);
            if (n0 < Lout) fc_st_stream(reinterpret_cast<float2*>(yrow + n0), make_float2(v[l][q].x + b, -v[l][q].y + b));
          }
        }
      }
    }
    if (++s == NS) s = 0, ph ^= 1;
  }
}
#endif  // !FC_CPU_EMUL

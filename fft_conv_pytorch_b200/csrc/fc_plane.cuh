// fc_plane.cuh — two axis passes of a 3-d transform in one kernel: a CTA holds the (z, y) plane of one (item, kx) in
// shared memory (at most 128 x 128 points = 129 KB), transforms its rows and then its columns with the group FFT engine
// (fc_fused.cuh: 8 points per lane, 32/G lines per warp) and writes the result in the layout the next step expects.
// The spectrum between the two passes never reaches HBM, and neither side needs a transposing tile:
//   fc_plane_fwd_kernel  [kx][z][y] (rows contiguous) -> y transform -> z transform -> [ky][kx][kz] (512-byte runs)
//   fc_plane_inv_kernel  [ky][kx][kz] -> z inverse -> y inverse -> crop -> [kx][jz][jy] (rows contiguous)
// They replace the C2C passes sig[1]+sig[2] / inv[0]+inv[1] of the 3-d program when both extents are 32, 64 or 128 and the
// maps are plain (constant padding, no zero-stuffing / subsampling; unit-stride crop).
#pragma once
#include "fc_fused.cuh"

struct fc_plane_args {
  const float2* in;
  float2* out;
  const float2* tw;
  int32_t tw_len;
  int32_t nkx;          // bins kept along x (planes per item)
  int64_t n_units;      // items * nkx
  int64_t in_os, out_os;  // item strides (complex elements)
  // forward: stored source extents and gather maps of the two axes (source index = dense index - pad)
  fc_imap imy, imz;
  int32_t conj_out;
  float scale;
  // inverse: crop maps of the two axes (output j takes dense index j + ob, zero at or beyond lim)
  fc_omap omy, omz;
};

#define FC_PLANE_WARPS 8

// Exchange line (pitch N + 2 float2 = 4 banks per line) of group gid inside its warp's block of 32/G * NL lines: the
// two groups of a half warp (G = 8) get lines 4 apart, i.e. 16 banks, so that their 8-element windows are disjoint
// (adjacent line pairs, 8 banks apart, overlap). G = 4 and G = 16 are conflict-free with adjacent pairs.
template <int G, int NL>
FC_DEV int fc_plane_scratch_line(int gid) {
  if constexpr (G == 8 && NL == 2) return (gid & 1) * 4 + (gid >> 1) * 2;
  else return gid * NL;
}
// Plane line (pitch N + 1) of group gid of warp w in a pass whose NW * 32/G * NL lines are exactly the plane's: the
// bank-disjoint assignment of K1 / K4 (fc_group_row); otherwise the plain one (z0 = first line of this warp pass).
template <int G, int NL, int LINES>
FC_DEV int fc_plane_line(int w, int gid, int z0) {
  if constexpr (LINES == FC_PLANE_WARPS * (32 / G) * NL) return fc_group_row<G, NL>(w, gid);
  else return z0 + gid * NL;
}

// Shared memory: the plane, NZ rows of pitch NY + 1 (the rows double as the exchange buffers of the row pass), then
// one exchange line of pitch NZ + 2 per line of the column pass.
template <int NY, int NZ>
__global__ void __launch_bounds__(FC_PLANE_WARPS * 32, 3) fc_plane_fwd_kernel(fc_plane_args a) {
  fc_grid_dep_sync();
  constexpr int GY = NY / 8, GZ = NZ / 8, NL = 2, NW = FC_PLANE_WARPS;
  constexpr int GPWY = 32 / GY, GPWZ = 32 / GZ;
  constexpr int PY = NY + 1, PZ = NZ + 2;
  FC_DYN_SMEM(smem);
  float2* plane = smem;
  float2* scratch = smem + NZ * PY;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int gly = lane % GY, gidy = lane / GY;
  const int glz = lane % GZ, gidz = lane / GZ;
  const fc_imap imy = a.imy, imz = a.imz;
  const int Ly = imy.L, Lz = imz.L;
  for (int64_t unit = blockIdx.x; unit < a.n_units; unit += gridDim.x) {
    const int64_t o = unit / a.nkx;
    const int kx = (int)(unit - o * a.nkx);
    const float2* src = a.in + o * a.in_os + (int64_t)kx * Lz * Ly;
    // ---- rows: transform along y (warp-uniform loop; NZ is a multiple of the rows a warp takes)
    for (int z0 = w * GPWY * NL; z0 < NZ; z0 += NW * GPWY * NL) {
      const int zr0 = fc_plane_line<GY, NL, NZ>(w, gidy, z0);
      float2 v[NL][8];
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        const int z = zr0 + l, zs = z - imz.pad;
        const bool zok = z < imz.ext && zs >= 0 && zs < Lz;
        const float2* row = src + (int64_t)(zok ? zs : 0) * Ly;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int y = gly + GY * q, ys = y - imy.pad;
          v[l][q] = (zok && y < imy.ext && ys >= 0 && ys < Ly) ? __ldg(row + ys) : make_float2(0.f, 0.f);
        }
      }
      float2* line0 = plane + zr0 * PY;
      fc_gfft<NY, GY, NL, PY>(v, line0, a.tw, a.tw_len, gly);
#pragma unroll
      for (int l = 0; l < NL; ++l)
#pragma unroll
        for (int q = 0; q < 8; ++q) line0[l * PY + gly + GY * q] = v[l][q];
    }
    __syncthreads();
    // ---- columns: transform along z, store [ky][kx][kz]
    for (int c0 = w * GPWZ * NL; c0 < NY; c0 += NW * GPWZ * NL) {
      const int ky0 = fc_plane_line<GZ, NL, NY>(w, gidz, c0);
      float2 v[NL][8];
#pragma unroll
      for (int l = 0; l < NL; ++l)
#pragma unroll
        for (int q = 0; q < 8; ++q) v[l][q] = plane[(glz + GZ * q) * PY + ky0 + l];
      float2* line0 = scratch + (size_t)(w * GPWZ * NL + fc_plane_scratch_line<GZ, NL>(gidz)) * PZ;
      fc_gfft<NZ, GZ, NL, PZ>(v, line0, a.tw, a.tw_len, glz);
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        float2* dst = a.out + o * a.out_os + (int64_t)(ky0 + l) * a.nkx * NZ + (int64_t)kx * NZ;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float2 val = fc_scale(v[l][q], a.scale);
          if (a.conj_out) val = fc_conj(val);
          dst[glz + GZ * q] = val;
        }
      }
    }
    __syncthreads();  // the plane is reused by the next unit
  }
}

// Shared memory: the plane, NY rows (ky) of pitch NZ + 1, then one exchange line of pitch NY + 2 per column-pass line.
template <int NY, int NZ>
__global__ void __launch_bounds__(FC_PLANE_WARPS * 32, 3) fc_plane_inv_kernel(fc_plane_args a) {
  fc_grid_dep_sync();
  constexpr int GY = NY / 8, GZ = NZ / 8, NL = 2, NW = FC_PLANE_WARPS;
  constexpr int GPWY = 32 / GY, GPWZ = 32 / GZ;
  constexpr int PZ = NZ + 1, PY = NY + 2;
  FC_DYN_SMEM(smem);
  float2* plane = smem;
  float2* scratch = smem + NY * PZ;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int gly = lane % GY, gidy = lane / GY;
  const int glz = lane % GZ, gidz = lane / GZ;
  const fc_omap omy = a.omy, omz = a.omz;
  const int Oy = omy.Lout, Oz = omz.Lout;
  for (int64_t unit = blockIdx.x; unit < a.n_units; unit += gridDim.x) {
    const int64_t o = unit / a.nkx;
    const int kx = (int)(unit - o * a.nkx);
    const float2* src = a.in + o * a.in_os + (int64_t)kx * NZ;
    // ---- rows (one per ky, kz contiguous): inverse transform along z (forward engine on conjugated data)
    for (int r0 = w * GPWZ * NL; r0 < NY; r0 += NW * GPWZ * NL) {
      const int ky0 = fc_plane_line<GZ, NL, NY>(w, gidz, r0);
      float2 v[NL][8];
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        const float2* row = src + (int64_t)(ky0 + l) * a.nkx * NZ;
#pragma unroll
        for (int q = 0; q < 8; ++q) v[l][q] = fc_conj(__ldg(row + glz + GZ * q));
      }
      float2* line0 = plane + ky0 * PZ;
      fc_gfft<NZ, GZ, NL, PZ>(v, line0, a.tw, a.tw_len, glz);
#pragma unroll
      for (int l = 0; l < NL; ++l)
#pragma unroll
        for (int q = 0; q < 8; ++q) line0[l * PZ + glz + GZ * q] = v[l][q];  // conj(result along z)
    }
    __syncthreads();
    // ---- columns (one per kept z): inverse transform along y, crop, store [kx][jz][jy]
    float2* dst0 = a.out + o * a.out_os + (int64_t)kx * Oz * Oy;
    for (int c0 = w * GPWY * NL; c0 < NZ; c0 += NW * GPWY * NL) {  // warp-uniform; columns past Oz idle along
      if (fc_plane_line<GY, NL, NZ>(w, 0, c0) >= Oz) break;  // (group 0 owns the warp's first column)
      const int jz0 = fc_plane_line<GY, NL, NZ>(w, gidy, c0);
      float2 v[NL][8];
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        const int n = jz0 + l + omz.ob;  // dense z index of output row jz
        const bool ok = jz0 + l < Oz && n < NZ && n < omz.lim;
#pragma unroll
        for (int q = 0; q < 8; ++q) v[l][q] = ok ? plane[(gly + GY * q) * PZ + n] : make_float2(0.f, 0.f);
      }
      float2* line0 = scratch + (size_t)(w * GPWY * NL + fc_plane_scratch_line<GY, NL>(gidy)) * PY;
      fc_gfft<NY, GY, NL, PY>(v, line0, a.tw, a.tw_len, gly);  // conj(conj(.)) = the inverse along both axes
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        if (jz0 + l >= Oz) continue;
        float2* dst = dst0 + (int64_t)(jz0 + l) * Oy;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int n = gly + GY * q, j = n - omy.ob;
          if (j >= 0 && j < Oy) dst[j] = (n < omy.lim) ? fc_conj(v[l][q]) : make_float2(0.f, 0.f);
        }
      }
    }
    __syncthreads();
  }
}

// fc_plan.h — host-side plan object: shape algebra + the program of axis passes for one convolution call.
#pragma once
#include <string>
#include <vector>

#include "fc_types.h"

enum { FC_BUF_USER_IN = 0, FC_BUF_SPEC = 1, FC_BUF_SA = 2, FC_BUF_SB = 3, FC_BUF_USER_OUT = 4 };

// Layout family of the transform program.
enum {
  FC_S_1D = 1,       // one real pass along the only axis
  FC_S_1D_SPLIT = 2, // four-step: N = N1*N2, real pass over the strided index, complex pass over the contiguous one
  FC_S_2D = 3,
  FC_S_3D = 4
};

// Lines per tile of the transposing fast kernels K1 / K4 (must match the instantiations in fc_api.cu).
inline int fc_fast_tile_lines(int M) { return M >= 256 ? 16 : 4096 / M; }

// The pair pipeline (fc_pair.cuh): pair lines per warp group and per tile of K1p / K4p (must match the instantiations
// in fc_api.cu: 8 warps per CTA).
inline int fc_pair_nlp(int M) { return M <= 256 ? 2 : 1; }
inline int fc_pair_tile_lines(int M) { return fc_pair_nlp(M) * 8 * (M >= 256 ? 1 : 256 / M); }

// Tensor-core contraction (fc_tc.cuh): batches run through the GEMM in chunks of up to FC_TC_MAX_BATCH (N = 2 * 80 = 160
// accumulator columns per 128-row tile: two sets in TMEM, three operand stages in shared memory), split evenly and padded to
// a multiple of 8.
#define FC_TC_MAX_BATCH 80
inline int fc_tc_chunk(int batch) {
  const int n = (batch + FC_TC_MAX_BATCH - 1) / FC_TC_MAX_BATCH;
  return (batch + n - 1) / n;
}

struct fc_step {
  fc_pass pass;
  int src;  // FC_BUF_*
  int dst;
};

// One kernel launch of fc_conv.
enum { FC_L_PASS = 0, FC_L_FAST_R2C = 1, FC_L_FAST_C2R = 2, FC_L_CONTRACT = 3, FC_L_FUSED = 4, FC_L_TC_X = 5, FC_L_TC_GEMM = 6, FC_L_TC_Y = 7, FC_L_FAST_C2C = 8, FC_L_COL_R2C = 9, FC_L_COL_C2R = 10, FC_L_PLANE_FWD = 11, FC_L_PLANE_INV = 12,
       FC_L_PAIR_R2C = 13, FC_L_PAIR_FUSED = 14, FC_L_PAIR_C2R = 15, FC_L_PAIR_FUSED64 = 16,
       FC_L_TC_FWD = 17,    // last forward pass writing the GEMM's Bt blobs itself (fc_tc_c2c_fwd_kernel)
       FC_L_TC_INV = 18,    // first inverse pass gathering from the GEMM's product (fc_tc_c2c_inv_kernel)
       FC_L_LINE_R2C = 19, FC_L_LINE_C2R = 20 };  // the real passes of a one-pass 1-d program on the warp engine (fc_line.cuh)

// Geometry of the fused "last forward axis -> contraction -> first inverse axis" kernel (fc_fused.cuh).
struct fc_fused_desc {
  int32_t N;      // transform length of the fused axis
  int32_t n_in;   // stored input elements per line
  int32_t n_out;  // stored output elements per line
  int32_t nb;     // batches per CTA
  int32_t ci;     // channel bound of the instantiation
  int32_t warps;  // warps per CTA
  int32_t occ;    // CTAs per SM the instantiation is compiled for (register cap)
  int32_t plain;  // identity gather map and plain crop: use the instantiation without the general map code
  // overlap-save segments of the fused axis (n_seg == 1: none): segment s transforms the dense positions
  // [s*seg_V - seg_off, s*seg_V - seg_off + N) and owns the dense outputs [s*seg_V, (s+1)*seg_V), which sit at local
  // index seg_off .. seg_off + seg_V - 1 of its circular result
  int32_t n_seg, seg_V, seg_off;
  int32_t fill_rows; // 1: the fused kernel also writes the bias-only output rows of a row lattice (see fc_fused_args::fill_*)
  int32_t ystage_S; // ... and the length of the sub-transforms the fused kernel is left with (64 or 128); N = ystage * ystage_S
  int32_t ystage; // pair program only: radix of the stage of this axis' transform that runs in K1p / K4p (0: none); the fused
                  // kernel is then fc_pair_fused64_kernel on independent 64-point sub-problems
  int64_t R;      // lines (bins of the other axes) per (batch, channel)
  int64_t Rk;     // ... per kernel-spectrum channel pair: line r multiplies kernel line r % Rk (Rk < R when the other
                  // axis is segmented: all its segments share one kernel spectrum)
  fc_imap imap;
  fc_omap omap;
};

// Geometry of the two-axis plane kernels of the 3-d program (fc_plane.cuh).
struct fc_plane_desc {
  int32_t ny, nz, nkx;
  int32_t conj_out;
  float scale;
  int64_t n_outer, in_os, out_os;
  fc_imap imy, imz;
  fc_omap omy, omz;
};

struct fc_launch {
  int type;
  fc_pass pass;  // FC_L_PASS / FC_L_FAST_*
  fc_fused_desc fused;
  fc_plane_desc plane;
  int src, dst;  // FC_BUF_* (FC_BUF_SPEC as src of an inverse step = product spectrum; as dst of a forward step = signal spectrum)
  int spec_is_y; // which spectrum buffer FC_BUF_SPEC means for this launch: 0 = xspec, 1 = yspec
  std::string name;
  int64_t bytes;  // algorithmic bytes: compulsory reads + writes of this launch
};

struct fc_axis {
  int L, K, stride, pad, dil, opad;
  int g;        // polyphase reduction factor gcd(stride, dilation)
  int N;        // transform extent
  int Nk;       // stored bins on this axis
  int Lout;     // output extent
  // overlap-save segmentation (first axis of a 2-d problem only; seg_n == 1: none). N is then the segment transform
  // length, seg_V = N - (dense kernel extent - 1) the outputs a segment owns, seg_off the local index of the first one
  // (0 for the correlation of fft_conv, dense kernel extent - 1 for the true convolution of fft_conv_transpose).
  int seg_n, seg_V, seg_off;
  int N_full;   // transform extent without segmentation
  fc_imap imap_sig, imap_ker;
  fc_omap omap;
};

struct fc_plan {
  fc_problem prob;       // the problem the program runs: the user's, or (batch segments) the windowed one with batch B * bseg_n
  fc_problem user_prob;  // what the caller asked for (sizes of the caller's tensors)
  int bseg_n;            // 1-d batch segments per line (fc_plan.cpp; 1: none)
  fc_plan_info info;
  int structure;
  int nd;
  int threads;
  int tw_len;
  int N1, N2;  // FC_S_1D_SPLIT factors
  fc_axis ax[FC_MAX_ND];
  std::vector<fc_step> sig_fwd, ker_fwd, inv;
  fc_contract_desc contract;
  std::vector<fc_launch> prog;  // what fc_conv launches, in order
  int64_t off_xspec, off_yspec, off_sA, off_sB;
  int64_t off_xtc, off_ytc;  // tensor-core contraction operands (use_tc)
  int use_tc;
  int pair;  // 1: fc_conv runs the packed batch-pair kernels (fc_pair.cuh); the spectrum buffers hold ceil(B/2) pair images per channel
  int64_t scratch_bytes;
};

// Returns FC_OK or a negative code; msg receives the reason.
int fc_plan_build(fc_plan* plan, const fc_problem* prob, std::string* msg);
// Choose fast / fused kernels where they apply and lay out the launch program of fc_conv.
void fc_plan_build_program(fc_plan* plan);
std::string fc_plan_to_string(const fc_plan* plan);

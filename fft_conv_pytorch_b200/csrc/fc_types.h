// fc_types.h — POD descriptors shared by the host planner (fc_plan.cpp) and the device kernels (fc_kernels.cuh).
//
// A convolution call is executed as a short program of "axis passes" over spectra kept in HBM:
//   R2C      real lines  -> half spectrum along one axis   (fused pad / zero-stuffing gather on load)
//   C2C_FWD  complex lines, forward transform of one more axis (gather map on load)
//   [contraction over channels, per frequency bin]
//   C2C_INV  inverse transform of one axis (crop / stride / lattice map on store)
//   C2R      half spectrum -> real lines (crop / stride / lattice map + bias on store)
// Each pass moves a tile of T lines through shared memory; on each side of a pass either the line
// elements are contiguous in HBM ("n-fast") or the T lines of a tile are adjacent ("r-fast", a transposing
// access), so every HBM access is made of >= T*8-byte contiguous segments.
#pragma once
#include <stdint.h>

#include "../../include/fftconv_b200.h"

enum { FC_R2C = 0, FC_C2C_FWD = 1, FC_C2C_INV = 2, FC_C2R = 3 };

// Gather map of a forward pass: dense transform position u in [0, N) -> source index, or "zero".
// Covers F.pad (reference functional.py:60-62; all four modes), the zero-stuffing of the transposed conv
// (functional.py:126-139, `up`), kernel dilation (functional.py:49-57, `up`) and the polyphase subsample `sub`.
struct fc_imap {
  int32_t L;     // source extent
  int32_t mode;  // FC_PAD_*
  int32_t pad;   // left padding
  int32_t up;    // zero-stuffing factor (>=1): only u % up == 0 carries data
  int32_t sub;   // subsample factor (>=1): dense position w reads padded position w*sub
  int32_t ext;   // dense extent; u >= ext is zero
};

// Scatter map of an inverse pass: output index j in [0, Lout) takes dense index q/og, q = j*os + ob, when
// q % og == 0 and q/og < lim, else 0 (+ bias in the last pass). Covers the crop+stride slice of
// functional.py:76-82 and the crop of functional.py:163-169.
struct fc_omap {
  int32_t Lout;
  int32_t os;
  int32_t ob;
  int32_t og;
  int32_t lim;
};

struct fc_pass {
  int32_t kind;       // FC_R2C .. FC_C2R
  int32_t N;          // transform length on this axis (real length for R2C / C2R)
  int32_t M;          // complex FFT length run in shared memory (N/2 for R2C / C2R, else N)
  int32_t T;          // lines per tile (power of two)
  int32_t log2T;
  int32_t pitch;      // shared-memory line pitch in complex elements
  int32_t n_in;       // stored input elements per line
  int32_t n_out;      // stored output elements per line
  int32_t flat;       // 1: tiles run over the flattened (outer, line) index (both sides n-fast)
  int32_t in_rfast;   // 1: input side is r-fast (lines of a tile adjacent in HBM)
  int32_t out_rfast;  // 1: output side is r-fast
  int32_t twiddle;    // 1: four-step twiddle W_twN^(k*r) on R2C store / conj on C2R load
  int32_t conj_out;   // forward passes: conjugate on store (kernel spectrum of the correlation)
  int32_t pos_n;      // dense position on the mapped axis: u = n*pos_n + r*pos_r
  int32_t pos_r;
  int32_t tw_len;     // length of the twiddle table (power of two >= every N of the plan)
  int32_t tw2_len;    // four-step plans: N2; a second table exp(-2*pi*i*b/(N1*N2)), b < N2, follows the first one
  int32_t cout;       // C2R: bias index = outer % cout
  int32_t has_bias;
  // C2R: lattice on the line (row) index: dense line r owns the output rows j with (j + row_ob) / row_og == r; only
  // the row with (j + row_ob) % row_og == 0 carries data, the others are bias only (polyphase-reduced transposed conv)
  int32_t row_og, row_ob, row_Lout;
  int32_t row_fill_skip;  // C2R (fast kernel): the bias-only rows between the lattice rows are written by the fused kernel
  // overlap-save segments along the line of an R2C / C2R pass (fast kernels K1 / K4 only; seg_n == 1: none). A line of
  // the pass is then a (row, segment) pair: segment s reads the dense positions s*seg_V - seg_off + [0, N) of its row and
  // its half spectrum is stored at bin offset s*(N/2+1); on the way back it owns the dense outputs s*seg_V + [0, seg_V),
  // found at local index seg_off + [0, seg_V). seg_V and seg_off are even.
  int32_t seg_n, seg_V, seg_off;
  // 1-d overlap-save with the segments as extra batch items (fc_plan.cpp, "batch segments"; bseg_n <= 1: none): outer item
  // o = (b*bseg_n + s)*bseg_c + c of the first R2C / last C2R pass. R2C: segment s reads its channel's line shifted by
  // s*bseg_V dense positions (imap.pad - s*bseg_V takes the place of imap.pad; the base of the line comes from o_c2 / o_q /
  // o_sA / o_sC as usual). C2R: it owns the outputs [s*bseg_Vo, (s+1)*bseg_Vo) of line (b*bseg_c + c) of the user's
  // output, whose lines are bseg_Lout long (out_os is ignored).
  int32_t bseg_n, bseg_c, bseg_V, bseg_Vo, bseg_Lout;
  float scale;        // forward passes: multiply on store (1/prod(N) folded into the kernel spectrum)
  int64_t twN;        // four-step twiddle modulus
  int64_t n_outer;
  int64_t R;          // lines per outer item
  int64_t tiles_per_outer;
  int64_t n_tiles;
  // element (outer o, line r, index n) lives at o*os + r*rs + n*es (complex elements, or floats on the real side)
  int64_t in_os, in_rs, in_es;
  int64_t out_os, out_rs, out_es;
  // generic pass only: when out_oq > 0 the output base of outer item o is (o / out_oq)*out_osA + (o % out_oq)*out_os
  // (the bin-major kernel spectrum of the fused plans: the out_oq = Og*Ig channel pairs of a group are adjacent lines)
  int64_t out_oq, out_osA;
  // ... and when out_il > 1 the out_oq lines of a group are interleaved out_il at a time: line q = o % out_oq starts at
  // (q / out_il) * out_il * out_os + q % out_il and its elements are out_es = out_il apart (the pair kernels read the
  // kernel values of two input channels of a bin with one 16-byte load)
  int64_t out_il;
  // ... and when out_split > 1 (a power of two) element k of a forward C2C line goes to
  // (k % out_split) * out_split_stride + (k / out_split) * out_es: the bins k1 + out_split*k2 of one residue k1 form a
  // contiguous chunk per (group, line) — the kernel-spectrum layout of the 64-bin sub-problems of fc_pair_fused64_kernel
  int64_t out_split, out_split_stride;
  // K1p / K4p: radix of the stage of the *other* axis' transform that runs in their transposed store / load (0: none;
  // fc_pair.cuh "y stage"): a tile is then the pair lines {n2 + 64*n1} of 16/ystage adjacent n2
  int32_t ystage, ystage_N, ystage_S;  // radix, transform length N = ystage * S, sub-transform length S (64 or 128)
  // R2C input base: base(o) = ((o/o_c2)/o_q)*o_sA + ((o/o_c2)%o_q)*o_sB + (o%o_c2)*o_sC   (replaces in_os)
  int64_t o_c2, o_q, o_sA, o_sB, o_sC;
  fc_imap imap;
  fc_omap omap;
};

struct fc_contract_desc {
  int64_t bins;  // F
  int32_t batch, cin, cout, groups;
};

// fc_plan.cpp — shape algebra of fft_conv / fft_conv_transpose and the axis-pass program.
//
// Mirrors (does not copy) the reference's argument handling:
//   forward     functional.py:44-66 (tuples, dilation, padding, transform extent) and :76-82 (crop + stride)
//   transposed  functional.py:103-154 (kernel regroup, dilation, zero-stuffing, extents) and :163-169 (crop)
// Differences, all result-preserving:
//   * the transform extent per axis is the next power of two >= the no-wrap minimum instead of "rounded to even";
//   * the transposed path runs a true (un-flipped, un-conjugated) circular convolution of the zero-stuffed
//     signal with the dilated kernel, which equals the reference's flip + left-pad + correlation;
//   * stride/dilation lattices with a common factor g are reduced by g (polyphase): only the populated
//     lattice is transformed, the remaining outputs are bias only.
#include "fc_plan.h"
#include "fc_tune.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>

namespace {

int ilog2(int64_t v) {
  int l = 0;
  while ((int64_t(1) << l) < v) ++l;
  return l;
}
int64_t next_pow2(int64_t v) { return int64_t(1) << ilog2(v); }
int gcd_i(int a, int b) {
  while (b) {
    int t = a % b;
    a = b;
    b = t;
  }
  return a;
}
int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

const int kMaxComplexLine = 4096;  // longest complex FFT done inside one CTA
const int kMaxRealLine = 8192;

fc_imap identity_imap(int n) {
  fc_imap m;
  m.L = n;
  m.mode = FC_PAD_CONSTANT;
  m.pad = 0;
  m.up = 1;
  m.sub = 1;
  m.ext = n;
  return m;
}
fc_omap identity_omap(int n) {
  fc_omap m;
  m.Lout = n;
  m.os = 1;
  m.ob = 0;
  m.og = 1;
  m.lim = n;
  return m;
}

// Fill the tile geometry of a pass once kind / N / R / n_outer / flags are set.
void finish_pass(fc_pass& p) {
  p.M = (p.kind == FC_R2C || p.kind == FC_C2R) ? p.N / 2 : p.N;
  if (p.M < 1) p.M = 1;
  // tile budget in complex elements: small tiles keep 4-6 CTAs per SM in flight (the pass is a load -> transform ->
  // store loop per CTA, so resident CTAs are what hides the HBM latency); a transposing side wants >= 16 lines
  const bool rfast = p.in_rfast || p.out_rfast;
  int budget = rfast ? 2048 : 4096;
  if (rfast && 16 * p.M > budget) budget = 16 * p.M;
  if (budget > 8192) budget = 8192;
  static const int env_flat = fc_tune_int("TILE_FLAT", 0);  // experiments
  if (!rfast && env_flat > 0) budget = env_flat;
  static const int env_rfast = fc_tune_int("TILE_RFAST", 0);  // experiments
  if (rfast && env_rfast > 0) budget = env_rfast;
  if (p.M > budget) budget = p.M;
  int T = 1;
  while (T * 2 * p.M <= budget && T * 2 <= 64) T *= 2;
  const int64_t lines_avail = p.flat ? p.n_outer * p.R : p.R;
  while (T > 1 && T / 2 >= lines_avail) T /= 2;
  p.T = T;
  p.log2T = ilog2(T);
  p.pitch = p.M + 1;
  if (p.flat) {
    p.tiles_per_outer = 0;
    p.n_tiles = (p.n_outer * p.R + T - 1) / T;
  } else {
    p.tiles_per_outer = (p.R + T - 1) / T;
    p.n_tiles = p.tiles_per_outer * p.n_outer;
  }
}

fc_pass blank_pass(int kind, int N, int tw_len) {
  fc_pass p;
  std::memset(&p, 0, sizeof(p));
  p.kind = kind;
  p.N = N;
  p.tw_len = tw_len;
  p.tw2_len = 1;
  p.scale = 1.f;
  p.pos_n = 1;
  p.pos_r = 0;
  p.o_c2 = 1;
  p.o_q = 1;
  p.twN = 1;
  p.row_og = 1;
  p.seg_n = 1;
  p.seg_V = N;
  p.imap = identity_imap(N);
  p.omap = identity_omap(N);
  return p;
}

// Description of the real tensor feeding a forward program (signal or weight).
struct SrcDesc {
  int64_t n_outer;                   // B*Cin or Cout*(Cin/g)
  int64_t o_c2, o_q, o_sA, o_sB, o_sC;
  fc_imap imap[FC_MAX_ND];
  int L[FC_MAX_ND];
  bool conj;
  float scale;
  bool segmented_x;  // the signal of a plan with overlap-save segments on the last axis (the kernel has one segment)
};

void build_forward(const fc_plan& pl, const SrcDesc& s, std::vector<fc_step>& out) {
  const int tw = pl.tw_len;
  auto set_src = [&](fc_pass& p) {
    p.o_c2 = s.o_c2;
    p.o_q = s.o_q;
    p.o_sA = s.o_sA;
    p.o_sB = s.o_sB;
    p.o_sC = s.o_sC;
  };
  if (pl.structure == FC_S_1D) {
    const fc_axis& a = pl.ax[0];
    fc_pass p = blank_pass(FC_R2C, a.N, tw);
    p.n_outer = s.n_outer;
    p.R = 1;
    p.flat = 1;
    set_src(p);
    p.in_rs = 0;
    p.in_es = 1;
    p.imap = s.imap[0];
    p.n_in = s.L[0];
    p.n_out = a.Nk;
    p.out_os = a.Nk;
    p.out_rs = 0;
    p.out_es = 1;
    p.conj_out = s.conj;
    p.scale = s.scale;
    finish_pass(p);
    out.push_back({p, FC_BUF_USER_IN, FC_BUF_SPEC});
  } else if (pl.structure == FC_S_1D_SPLIT) {
    const int N1 = pl.N1, N2 = pl.N2, Nk1 = N1 / 2 + 1;
    fc_pass p = blank_pass(FC_R2C, N1, tw);
    p.n_outer = s.n_outer;
    p.R = N2;
    set_src(p);
    p.in_rs = 0;
    p.in_es = 1;
    p.in_rfast = 1;
    p.pos_n = N2;
    p.pos_r = 1;
    p.imap = s.imap[0];
    p.n_in = N1;
    p.n_out = Nk1;
    p.out_os = (int64_t)Nk1 * N2;
    p.out_es = N2;
    p.out_rs = 1;
    p.out_rfast = 1;
    p.twiddle = 1;
    p.twN = (int64_t)N1 * N2;
    p.tw2_len = N2;
    finish_pass(p);
    out.push_back({p, FC_BUF_USER_IN, FC_BUF_SA});
    fc_pass q = blank_pass(FC_C2C_FWD, N2, tw);
    q.n_outer = s.n_outer;
    q.R = Nk1;
    q.flat = 1;
    q.in_os = (int64_t)Nk1 * N2;
    q.in_rs = N2;
    q.in_es = 1;
    q.n_in = N2;
    q.n_out = N2;
    q.out_os = (int64_t)Nk1 * N2;
    q.out_rs = N2;
    q.out_es = 1;
    q.conj_out = s.conj;
    q.scale = s.scale;
    finish_pass(q);
    out.push_back({q, FC_BUF_SA, FC_BUF_SPEC});
  } else if (pl.structure == FC_S_2D) {
    const fc_axis &ay = pl.ax[0], &ax = pl.ax[1];
    const int Ly = s.L[0], Lx = s.L[1];
    const int64_t nkx = (int64_t)ax.Nk * (s.segmented_x ? ax.seg_n : 1);  // bins of all segments of a row
    fc_pass p = blank_pass(FC_R2C, ax.N, tw);
    p.n_outer = s.n_outer;
    p.R = Ly;
    set_src(p);
    p.in_rs = Lx;
    p.in_es = 1;
    p.imap = s.imap[1];
    p.n_in = Lx;
    p.n_out = ax.Nk;
    p.out_os = nkx * Ly;  // [o][kx][y]
    p.out_es = Ly;
    p.out_rs = 1;
    p.out_rfast = 1;
    if (s.segmented_x) {
      p.seg_n = ax.seg_n;
      p.seg_V = ax.seg_V;
      p.seg_off = ax.seg_off;
    }
    finish_pass(p);
    out.push_back({p, FC_BUF_USER_IN, FC_BUF_SA});
    fc_pass q = blank_pass(FC_C2C_FWD, ay.N, tw);
    q.n_outer = s.n_outer;
    q.R = nkx;
    q.flat = 1;
    q.in_os = nkx * Ly;
    q.in_rs = Ly;
    q.in_es = 1;
    q.imap = s.imap[0];
    q.n_in = Ly;
    q.n_out = ay.N;
    q.out_os = nkx * ay.N;  // [o][kx][ky]
    q.out_rs = ay.N;
    q.out_es = 1;
    q.conj_out = s.conj;
    q.scale = s.scale;
    finish_pass(q);
    out.push_back({q, FC_BUF_SA, FC_BUF_SPEC});
  } else {  // FC_S_3D
    const fc_axis &az = pl.ax[0], &ay = pl.ax[1], &ax = pl.ax[2];
    const int Lz = s.L[0], Ly = s.L[1], Lx = s.L[2];
    fc_pass p = blank_pass(FC_R2C, ax.N, tw);
    p.n_outer = s.n_outer;
    p.R = (int64_t)Lz * Ly;
    set_src(p);
    p.in_rs = Lx;
    p.in_es = 1;
    p.imap = s.imap[2];
    p.n_in = Lx;
    p.n_out = ax.Nk;
    p.out_os = (int64_t)ax.Nk * Lz * Ly;  // [o][kx][z][y]
    p.out_es = (int64_t)Lz * Ly;
    p.out_rs = 1;
    p.out_rfast = 1;
    finish_pass(p);
    out.push_back({p, FC_BUF_USER_IN, FC_BUF_SA});
    fc_pass q = blank_pass(FC_C2C_FWD, ay.N, tw);
    q.n_outer = s.n_outer;
    q.R = (int64_t)ax.Nk * Lz;  // lines (kx, z)
    q.in_os = (int64_t)ax.Nk * Lz * Ly;
    q.in_rs = Ly;
    q.in_es = 1;
    q.imap = s.imap[1];
    q.n_in = Ly;
    q.n_out = ay.N;
    q.out_os = (int64_t)ay.N * ax.Nk * Lz;  // [o][ky][kx][z]
    q.out_es = (int64_t)ax.Nk * Lz;
    q.out_rs = 1;
    q.out_rfast = 1;
    finish_pass(q);
    out.push_back({q, FC_BUF_SA, FC_BUF_SB});
    fc_pass r = blank_pass(FC_C2C_FWD, az.N, tw);
    r.n_outer = s.n_outer;
    r.R = (int64_t)ay.N * ax.Nk;  // lines (ky, kx)
    r.flat = 1;
    r.in_os = (int64_t)ay.N * ax.Nk * Lz;
    r.in_rs = Lz;
    r.in_es = 1;
    r.imap = s.imap[0];
    r.n_in = Lz;
    r.n_out = az.N;
    r.out_os = (int64_t)ay.N * ax.Nk * az.N;  // [o][ky][kx][kz]
    r.out_rs = az.N;
    r.out_es = 1;
    r.conj_out = s.conj;
    r.scale = s.scale;
    finish_pass(r);
    out.push_back({r, FC_BUF_SB, FC_BUF_SPEC});
  }
}

void build_inverse(const fc_plan& pl, std::vector<fc_step>& out) {
  const int tw = pl.tw_len;
  const int64_t n_outer = (int64_t)pl.prob.batch * pl.prob.cout;
  const int has_bias = 1;  // resolved at launch time (null bias pointer => 0)
  if (pl.structure == FC_S_1D) {
    const fc_axis& a = pl.ax[0];
    fc_pass p = blank_pass(FC_C2R, a.N, tw);
    p.n_outer = n_outer;
    p.R = 1;
    p.flat = 1;
    p.in_os = a.Nk;
    p.in_rs = 0;
    p.in_es = 1;
    p.n_in = a.Nk;
    p.n_out = a.Lout;
    p.out_os = a.Lout;
    p.out_rs = 0;
    p.out_es = 1;
    p.omap = a.omap;
    p.cout = pl.prob.cout;
    p.has_bias = has_bias;
    if (!p.row_Lout) p.row_Lout = (int32_t)p.R;
    finish_pass(p);
    out.push_back({p, FC_BUF_SPEC, FC_BUF_USER_OUT});
  } else if (pl.structure == FC_S_1D_SPLIT) {
    const fc_axis& a = pl.ax[0];
    const int N1 = pl.N1, N2 = pl.N2, Nk1 = N1 / 2 + 1;
    fc_pass q = blank_pass(FC_C2C_INV, N2, tw);
    q.n_outer = n_outer;
    q.R = Nk1;
    q.flat = 1;
    q.in_os = (int64_t)Nk1 * N2;
    q.in_rs = N2;
    q.in_es = 1;
    q.n_in = N2;
    q.n_out = N2;
    q.out_os = (int64_t)Nk1 * N2;
    q.out_rs = N2;
    q.out_es = 1;
    finish_pass(q);
    out.push_back({q, FC_BUF_SPEC, FC_BUF_SA});
    fc_pass p = blank_pass(FC_C2R, N1, tw);
    p.n_outer = n_outer;
    p.R = N2;
    p.in_os = (int64_t)Nk1 * N2;
    p.in_es = N2;
    p.in_rs = 1;
    p.in_rfast = 1;
    p.n_in = Nk1;
    p.twiddle = 1;
    p.twN = (int64_t)N1 * N2;
    p.tw2_len = N2;
    p.pos_n = N2;
    p.pos_r = 1;
    p.out_os = a.Lout;
    p.out_rs = 0;
    p.out_es = 1;
    p.out_rfast = 1;
    p.n_out = a.Lout;
    p.omap = a.omap;
    p.cout = pl.prob.cout;
    p.has_bias = has_bias;
    if (!p.row_Lout) p.row_Lout = (int32_t)p.R;
    finish_pass(p);
    out.push_back({p, FC_BUF_SA, FC_BUF_USER_OUT});
  } else if (pl.structure == FC_S_2D) {
    const fc_axis &ay = pl.ax[0], &ax = pl.ax[1];
    // Rows of a polyphase-reduced transposed convolution: only every og-th output row carries data. The inverse
    // pass over y keeps the dense rows only (D0 of them) and the C2R pass scatters each to its output row and
    // fills the og-1 bias-only rows it owns.
    const bool row_lattice = ay.omap.og > 1;
    const int D0 = row_lattice ? (ay.Lout - 1 + ay.omap.ob) / ay.omap.og + 1 : ay.Lout;
    const int64_t nkx = (int64_t)ax.Nk * ax.seg_n;  // bins of all segments of a row
    fc_pass q = blank_pass(FC_C2C_INV, ay.N, tw);
    q.n_outer = n_outer;
    q.R = nkx;
    q.flat = 1;
    q.in_os = nkx * ay.N;
    q.in_rs = ay.N;
    q.in_es = 1;
    q.n_in = ay.N;
    q.n_out = D0;
    q.out_os = nkx * D0;  // [o][kx][jy]
    q.out_rs = D0;
    q.out_es = 1;
    q.omap = ay.omap;
    if (row_lattice) {
      q.omap.Lout = D0;
      q.omap.os = 1;
      q.omap.ob = 0;
      q.omap.og = 1;
    }
    finish_pass(q);
    out.push_back({q, FC_BUF_SPEC, FC_BUF_SA});
    fc_pass p = blank_pass(FC_C2R, ax.N, tw);
    p.n_outer = n_outer;
    p.R = D0;
    p.seg_n = ax.seg_n;
    p.seg_V = ax.seg_n > 1 ? ax.seg_V : ax.N;
    p.seg_off = ax.seg_off;
    p.in_os = nkx * D0;
    p.in_es = D0;
    p.in_rs = 1;
    p.in_rfast = 1;
    p.n_in = ax.Nk;
    p.n_out = ax.Lout;
    p.out_os = (int64_t)ay.Lout * ax.Lout;
    p.out_rs = ax.Lout;
    p.out_es = 1;
    p.omap = ax.omap;
    p.cout = pl.prob.cout;
    p.has_bias = has_bias;
    p.row_Lout = ay.Lout;
    if (row_lattice) {
      p.row_og = ay.omap.og;
      p.row_ob = ay.omap.ob;
    }
    finish_pass(p);
    out.push_back({p, FC_BUF_SA, FC_BUF_USER_OUT});
  } else {
    const fc_axis &az = pl.ax[0], &ay = pl.ax[1], &ax = pl.ax[2];
    fc_pass r = blank_pass(FC_C2C_INV, az.N, tw);
    r.n_outer = n_outer;
    r.R = (int64_t)ay.N * ax.Nk;
    r.flat = 1;
    r.in_os = (int64_t)ay.N * ax.Nk * az.N;
    r.in_rs = az.N;
    r.in_es = 1;
    r.n_in = az.N;
    r.n_out = az.Lout;
    r.out_os = (int64_t)ay.N * ax.Nk * az.Lout;  // [o][ky][kx][jz]
    r.out_rs = az.Lout;
    r.out_es = 1;
    r.omap = az.omap;
    finish_pass(r);
    out.push_back({r, FC_BUF_SPEC, FC_BUF_SA});
    fc_pass q = blank_pass(FC_C2C_INV, ay.N, tw);
    q.n_outer = n_outer;
    q.R = (int64_t)ax.Nk * az.Lout;  // lines (kx, jz)
    q.in_os = (int64_t)ay.N * ax.Nk * az.Lout;
    q.in_es = (int64_t)ax.Nk * az.Lout;
    q.in_rs = 1;
    q.in_rfast = 1;
    q.n_in = ay.N;
    q.n_out = ay.Lout;
    q.out_os = (int64_t)ax.Nk * az.Lout * ay.Lout;  // [o][kx][jz][jy]
    q.out_rs = ay.Lout;
    q.out_es = 1;
    q.omap = ay.omap;
    finish_pass(q);
    out.push_back({q, FC_BUF_SA, FC_BUF_SB});
    fc_pass p = blank_pass(FC_C2R, ax.N, tw);
    p.n_outer = n_outer;
    p.R = (int64_t)az.Lout * ay.Lout;  // lines (jz, jy)
    p.in_os = (int64_t)ax.Nk * az.Lout * ay.Lout;
    p.in_es = (int64_t)az.Lout * ay.Lout;
    p.in_rs = 1;
    p.in_rfast = 1;
    p.n_in = ax.Nk;
    p.n_out = ax.Lout;
    p.out_os = (int64_t)az.Lout * ay.Lout * ax.Lout;
    p.out_rs = ax.Lout;
    p.out_es = 1;
    p.omap = ax.omap;
    p.cout = pl.prob.cout;
    p.has_bias = has_bias;
    if (!p.row_Lout) p.row_Lout = (int32_t)p.R;
    finish_pass(p);
    out.push_back({p, FC_BUF_SB, FC_BUF_USER_OUT});
  }
}

int64_t step_out_bytes(const fc_step& s) { return s.pass.n_outer * s.pass.out_os * 8; }

}  // namespace

namespace {
int plan_build_core(fc_plan* pl, const fc_problem* prob, std::string* msg);

// ---- 1-d overlap-save with the segments as extra batch items ("batch segments").
// A 1-d fft_conv of a long line is cut into S windows of Ns points that start V = Vo*stride positions apart; each window is
// a *valid* convolution of its own and yields the Vo outputs [s*Vo, (s+1)*Vo) of the line. The plan is therefore the plan of
// the problem (batch B*S, length Ns, padding 0): only its first pass (which reads window s of the user's line, zero padding
// included) and its last pass (which writes run s of the user's output line) know about the segments (fc_pass::bseg_*).
// What it buys: the kernel spectrum and the contraction's operand traffic shrink from N to Ns bins per channel pair
// (BASELINE c4: 65536 -> 16384 points, kernel spectrum 17.2 -> 4.3 GB), and a line just above a power of two no longer
// pays for the next one (33000 points, K = 64: 5 x 8192 instead of 65536).
struct bseg_choice {
  int Ns, S, V, Vo;
};
bool choose_batch_segments(const fc_problem& P, bseg_choice* c) {
  if (P.ndim != 1 || P.transposed || (P.flags & (FC_FLAG_NO_SEGMENT | FC_FLAG_NO_FUSED))) return false;
  if (P.batch < 1 || P.cin < 1 || P.cout < 1 || P.groups < 1 || P.cin % P.groups || P.cout % P.groups) return false;
  const int64_t L = P.in_size[0], K = P.kernel_size[0], st = P.stride[0], d = P.dilation[0], pad = P.padding[0];
  if (L < 1 || K < 1 || st < 1 || d < 1 || pad < 0) return false;
  if (pad > 0 && P.padding_mode != FC_PAD_CONSTANT) return false;  // the windows see zeros outside the line
  if (gcd_i((int)st, (int)d) != 1) return false;                    // (polyphase-reduced lattices keep the one-transform plan)
  const int64_t Kd = (K - 1) * d + 1, Lp = L + 2 * pad;
  if (Lp < Kd) return false;
  const int64_t Lout = (Lp - Kd) / st + 1;
  const int64_t N0 = next_pow2(std::max<int64_t>(Lp, 2));
  // Cost in bytes moved, from the geometry and the channel counts only (the window length never depends on the batch, so
  // plans of different batch sizes that are both segmented share one kernel spectrum): the transform passes move ~24 bytes
  // per point and channel for a nominal batch of 16, the contraction reads the kernel spectrum once.
  const double Ig = (double)(P.cin / P.groups);
  auto cost = [&](double Ns, double S) { return S * Ns * (P.cin + P.cout) * 24.0 * 16.0 + (double)P.cout * Ig * (Ns / 2) * 8.0; };
  // Channel groups the fused axis kernel serves (<= 16 per group) keep the four-step layout and its specialised kernels:
  // windows of at least 64 x 256 points. (Measured, profiles/r2b_batch_segments_probe.txt: 8 channels, 33000 points, K = 64
  // runs 0.042 ms on one 65536-point transform against 0.055 ms on 17 windows of 2048 points through the generic passes.)
  const bool fusable = P.cin / P.groups <= 16 && P.cout / P.groups <= 16;
  // How much cheaper a segmented plan has to be. Wide channel groups (generic passes around the contraction): clearly, 25 %.
  // Fusable groups: any gain — and small problems (<= 2^22 point-channels) even at 1.6 x the bytes, because 16384-point
  // windows run the 64 x 256 kernels on S times as many CTAs and a small call is latency-bound (profiles/r2b_tiny_probe.txt:
  // BASELINE c1 34.8 -> 24.6 us on 3 windows; (1,4,131072) K = 513: 65.5 -> 28.7 us on 9). This one rule looks at the batch:
  // callers that need a batch-independent choice (the host pipeline's batch chunks) pin it with FC_FLAG_SEGMENT /
  // FC_FLAG_NO_SEGMENT. FC_FLAG_SEGMENT: take the cheapest window length whatever the gain.
  const bool small = fusable && (double)P.batch * (P.cin + P.cout) * (double)N0 <= (double)(1 << 22);
  double best = (P.flags & FC_FLAG_SEGMENT) ? 1e300 : (small ? 1.6 : fusable ? 1.0 : 0.75) * cost((double)N0, 1.0);
  bool found = false;
  for (int64_t Ns = fusable ? 16384 : 2048; Ns <= N0 / 2; Ns *= 2) {  // (from 2048 points a window runs on the four-step kernels)
    if (Ns < 2 * Kd) continue;  // at least half of a window is output
    const int64_t Vo = (Ns - Kd) / st + 1, S = (Lout + Vo - 1) / Vo;
    if (S < 2 || (int64_t)P.batch * S > (1 << 20)) continue;
    const double cst = cost((double)Ns, (double)S);
    if (cst < best) {
      best = cst;
      found = true;
      c->Ns = (int)Ns;
      c->S = (int)S;
      c->Vo = (int)Vo;
      c->V = (int)(Vo * st);
    }
  }
  return found;
}
}  // namespace

int fc_plan_build(fc_plan* pl, const fc_problem* prob, std::string* msg) {
  bseg_choice c;
  if (prob && choose_batch_segments(*prob, &c)) {
    const fc_problem& P = *prob;
    fc_problem P2 = P;
    P2.batch = P.batch * c.S;
    P2.in_size[0] = c.Ns;
    P2.padding[0] = 0;
    P2.padding_mode = FC_PAD_CONSTANT;
    std::string m2;
    if (plan_build_core(pl, &P2, &m2) == FC_OK && !pl->prog.empty() &&
        (pl->prog.front().type == FC_L_PASS || pl->prog.front().type == FC_L_COL_R2C || pl->prog.front().type == FC_L_LINE_R2C) &&
        pl->prog.front().pass.kind == FC_R2C &&
        (pl->prog.back().type == FC_L_PASS || pl->prog.back().type == FC_L_COL_C2R || pl->prog.back().type == FC_L_LINE_C2R) &&
        pl->prog.back().pass.kind == FC_C2R &&
        pl->ax[0].Lout == c.Vo && pl->ax[0].N == c.Ns) {
      const int64_t Kd = (int64_t)(P.kernel_size[0] - 1) * P.dilation[0] + 1;
      const int Lout = (int)(((int64_t)P.in_size[0] + 2 * P.padding[0] - Kd) / P.stride[0] + 1);
      auto patch_in = [&](fc_pass& p) {
        p.imap.L = P.in_size[0];
        p.imap.pad = P.padding[0];
        p.imap.mode = FC_PAD_CONSTANT;
        p.o_c2 = P.cin;
        p.o_q = c.S;
        p.o_sA = (int64_t)P.cin * P.in_size[0];
        p.o_sB = 0;
        p.o_sC = P.in_size[0];
        p.bseg_n = c.S;
        p.bseg_c = P.cin;
        p.bseg_V = c.V;
      };
      auto patch_out = [&](fc_pass& p) {
        p.bseg_n = c.S;
        p.bseg_c = P.cout;
        p.bseg_Vo = c.Vo;
        p.bseg_Lout = Lout;
      };
      patch_in(pl->sig_fwd.front().pass);
      patch_in(pl->prog.front().pass);
      patch_out(pl->inv.back().pass);
      patch_out(pl->prog.back().pass);
      pl->user_prob = P;
      pl->bseg_n = c.S;
      pl->info.out_size[0] = Lout;
      pl->info.out_elems = (int64_t)P.batch * P.cout * Lout;
      pl->info.segments = c.S;
      return FC_OK;
    }
    // (no program this code can drive with segments: fall through to the one-transform plan)
    *pl = fc_plan();
  }
  const int rc = plan_build_core(pl, prob, msg);
  if (prob) pl->user_prob = *prob;
  pl->bseg_n = 1;
  return rc;
}

namespace {
int plan_build_core(fc_plan* pl, const fc_problem* prob, std::string* msg) {
  auto fail = [&](int code, const std::string& m) {
    if (msg) *msg = m;
    return code;
  };
  if (!prob) return fail(FC_ENULL, "problem is NULL");
  const fc_problem& P = *prob;
  pl->prob = P;
  std::memset(&pl->info, 0, sizeof(pl->info));
  const int nd = P.ndim;
  if (nd < 1 || nd > FC_MAX_ND) return fail(FC_EUNSUPPORTED, "ndim must be 1, 2 or 3 (got " + std::to_string(nd) + ")");
  if (P.batch < 1 || P.cin < 1 || P.cout < 1 || P.groups < 1) return fail(FC_EINVAL, "batch, channels and groups must be positive");
  if (P.cin % P.groups || P.cout % P.groups)
    return fail(FC_EINVAL, "in_channels (" + std::to_string(P.cin) + ") and out_channels (" + std::to_string(P.cout) +
                               ") must be divisible by groups (" + std::to_string(P.groups) + ")");
  if (P.padding_mode < 0 || P.padding_mode > 3) return fail(FC_EINVAL, "unknown padding_mode");
  if (P.transposed && P.padding_mode != FC_PAD_CONSTANT) return fail(FC_EINVAL, "fft_conv_transpose has no padding_mode");
  pl->nd = nd;
  static const int env_threads = fc_tune_int("THREADS", 0);  // experiments
  pl->threads = P.threads > 0 ? P.threads : env_threads > 0 ? env_threads : 256;
  if (pl->threads % 32 || pl->threads > 1024) return fail(FC_EINVAL, "threads must be a multiple of 32, <= 1024");
  const bool poly = !(P.flags & FC_FLAG_NO_POLYPHASE);
  const int Ig_ = P.cin / P.groups, Og_ = P.cout / P.groups;
  static const char* env_seg = fc_tune_str("SEG");  // experiments: "Ny,Nx" forces the segment lengths (0 = automatic)
  int force_seg[2] = {0, 0};
  if (env_seg) std::sscanf(env_seg, "%d,%d", &force_seg[0], &force_seg[1]);
  // segments need the fused axis kernel (fc_plan_build_program: fuse_mid)
  const bool seg_ok = !(P.flags & (FC_FLAG_NO_SEGMENT | FC_FLAG_NO_FUSED | FC_FLAG_NO_FUSED_MID)) && Ig_ <= 16 && Og_ <= 16;

  int64_t bins = 1, out_vol = 1, in_vol = 1, k_vol = 1;
  double inv_scale = 1.0;
  for (int i = 0; i < nd; ++i) {
    fc_axis& a = pl->ax[i];
    a.L = P.in_size[i];
    a.K = P.kernel_size[i];
    a.stride = P.stride[i];
    a.pad = P.padding[i];
    a.dil = P.dilation[i];
    a.opad = P.transposed ? P.output_padding[i] : 0;
    const std::string ax_s = "axis " + std::to_string(i) + ": ";
    if (a.L < 1 || a.K < 1) return fail(FC_EINVAL, ax_s + "signal and kernel extents must be positive");
    if (a.stride < 1 || a.dil < 1) return fail(FC_EINVAL, ax_s + "stride and dilation must be >= 1");
    if (a.pad < 0 || a.opad < 0) return fail(FC_EINVAL, ax_s + "padding and output_padding must be >= 0");
    a.g = poly ? gcd_i(a.stride, a.dil) : 1;
    const int s2 = a.stride / a.g, d2 = a.dil / a.g;
    const int64_t Kd = (int64_t)(a.K - 1) * a.dil + 1;
    int64_t need;
    if (!P.transposed) {
      const int64_t Lp = (int64_t)a.L + 2 * a.pad;
      if (Lp < Kd)
        return fail(FC_EINVAL, ax_s + "dilated kernel extent (" + std::to_string(Kd) + ") exceeds the padded signal extent (" +
                                   std::to_string(Lp) + ")");
      if (P.padding_mode == FC_PAD_REFLECT && a.pad > a.L - 1) return fail(FC_EINVAL, ax_s + "reflect padding must be < signal extent");
      if (P.padding_mode == FC_PAD_CIRCULAR && a.pad > a.L) return fail(FC_EINVAL, ax_s + "circular padding must be <= signal extent");
      a.Lout = (int)((Lp - Kd) / a.stride + 1);
      a.imap_sig.L = a.L;
      a.imap_sig.mode = P.padding_mode;
      a.imap_sig.pad = a.pad;
      a.imap_sig.up = 1;
      a.imap_sig.sub = a.g;
      a.imap_sig.ext = (int)((Lp + a.g - 1) / a.g);
      a.omap.Lout = a.Lout;
      a.omap.os = s2;
      a.omap.ob = 0;
      a.omap.og = 1;
      need = a.imap_sig.ext;
    } else {
      const int64_t ext = (int64_t)(a.L - 1) * s2 + 1;
      const int64_t Kd2 = (int64_t)(a.K - 1) * d2 + 1;
      const int64_t dense = ext + Kd2 - 1;
      const int64_t Lout = (int64_t)(a.L - 1) * a.stride - 2 * (int64_t)a.pad + (int64_t)a.dil * (a.K - 1) + a.opad + 1;
      if (Lout < 1) return fail(FC_EINVAL, ax_s + "transposed convolution output extent would be " + std::to_string(Lout));
      a.Lout = (int)Lout;
      a.imap_sig.L = a.L;
      a.imap_sig.mode = FC_PAD_CONSTANT;
      a.imap_sig.pad = 0;
      a.imap_sig.up = s2;
      a.imap_sig.sub = 1;
      a.imap_sig.ext = (int)ext;
      a.omap.Lout = a.Lout;
      a.omap.os = 1;
      a.omap.ob = a.pad;
      a.omap.og = a.g;
      a.omap.lim = (int)dense;
      need = std::max<int64_t>(dense, (Lout - 1 + a.pad) / a.g + 1);
    }
    a.imap_ker.L = a.K;
    a.imap_ker.mode = FC_PAD_CONSTANT;
    a.imap_ker.pad = 0;
    a.imap_ker.up = d2;
    a.imap_ker.sub = 1;
    a.imap_ker.ext = (a.K - 1) * d2 + 1;
    int64_t N = next_pow2(std::max<int64_t>(need, 2));
    const bool last = (i == nd - 1);
    a.N_full = (int)std::min<int64_t>(N, int64_t(1) << 30);
    a.seg_n = 1;
    a.seg_V = 0;
    a.seg_off = 0;
    if (nd == 2 && i == 1 && seg_ok && (pl->ax[0].N == 256 || pl->ax[0].N == 512 || pl->ax[0].N == 1024) &&
        !(P.flags & (FC_FLAG_NO_FAST_R2C | FC_FLAG_NO_FAST_C2R)) && !(a.L & 1) &&
        (P.transposed ? s2 == 1 : (!(a.pad & 1) && a.g == 1 && P.padding_mode == FC_PAD_CONSTANT))) {
      // Overlap-save along x as well: the transposing row kernels K1 / K4 (which must both apply, see
      // fc_plan_build_program) treat a (row, segment) pair as a line. Segment extents are kept even so that the packed
      // real transforms load and store aligned pairs.
      const int64_t Kd2 = (int64_t)(a.K - 1) * d2 + 1;
      const int64_t n_need = ((int64_t)(a.Lout - 1) * a.omap.os + a.omap.ob) / a.omap.og + 1;
      const int64_t off = P.transposed ? ((Kd2 - 1 + 1) & ~int64_t(1)) : 0;
      // cost = bins per row (the work of every later kernel is proportional to it); it depends on the geometry only, never
      // on the batch, so the plans of the batch chunks of the host pipeline agree with the full-batch plan they share
      // the kernel spectrum with
      double best = N <= kMaxRealLine ? (double)(N / 2 + 1) : 1e300;
      int best_ns = 0;
      for (int Ns = 256; Ns <= 2048 && Ns < N; Ns *= 2) {  // (128-point segments measured slower: the short-row kernels lose more than the bins save)
        const int64_t V = P.transposed ? Ns - off : ((Ns - Kd2 + 1) & ~int64_t(1));
        if (V < Ns / 2) continue;
        const int64_t ns = (n_need + V - 1) / V;
        const double c = (double)ns * (Ns / 2 + 1);
        if (force_seg[1] ? Ns == force_seg[1] : c < best) {
          best = c;
          best_ns = Ns;
        }
      }
      if (best_ns) {
        a.seg_V = (int)(P.transposed ? best_ns - off : ((best_ns - Kd2 + 1) & ~int64_t(1)));
        a.seg_n = (int)((n_need + a.seg_V - 1) / a.seg_V);
        a.seg_off = (int)off;
        a.omap.lim = P.transposed ? a.omap.lim : a.seg_n * a.seg_V;
        N = best_ns;
      }
    }
    if (nd == 2 && i == 0 && seg_ok) {
      // Overlap-save along y: the fused axis kernel transforms segments of Ns points, of which
      // V = Ns - (Kd - 1) outputs are alias-free, so the kernel spectrum is Ns instead of N bins long on this axis
      // (BASELINE c5: 2048 -> 256, 17.2 GB -> 2.1 GB) and a long axis stays inside the fused kernel's line lengths.
      // Cost: points transformed per line (segments * Ns against the unsegmented N, when that is fusable at all); geometry
      // only, see the last axis below.
      const int64_t Kd2 = (int64_t)(a.K - 1) * d2 + 1;
      const int64_t n_need = ((int64_t)(a.Lout - 1) * a.omap.os + a.omap.ob) / a.omap.og + 1;  // dense outputs to produce
      double best = (N >= 256 && N <= 1024) ? (double)N : 1e300;  // unsegmented and fusable
      int best_ns = 0;
      for (int Ns = 256; Ns <= 1024 && Ns < N; Ns *= 2) {
        const int64_t V = Ns - Kd2 + 1;
        if (V < Ns / 2) continue;
        const int64_t ns = (n_need + V - 1) / V;
        const double c = (double)ns * Ns;
        if (force_seg[0] ? Ns == force_seg[0] : c < best) {
          best = c;
          best_ns = Ns;
        }
      }
      if (best_ns) {
        a.seg_V = (int)(best_ns - Kd2 + 1);
        a.seg_n = (int)((n_need + a.seg_V - 1) / a.seg_V);
        a.seg_off = P.transposed ? (int)(Kd2 - 1) : 0;
        a.omap.lim = P.transposed ? a.omap.lim : a.seg_n * a.seg_V;
        N = best_ns;
      }
    }
    if (nd > 1 && N > (last ? kMaxRealLine : kMaxComplexLine))
      return fail(FC_EUNSUPPORTED, ax_s + "transform extent " + std::to_string(N) + " exceeds the per-axis limit");
    if (nd == 1 && N > (int64_t)kMaxRealLine * kMaxComplexLine)
      return fail(FC_EUNSUPPORTED, ax_s + "transform extent " + std::to_string(N) + " exceeds the 1-d limit");
    a.N = (int)N;
    if (!P.transposed && a.seg_n == 1) a.omap.lim = a.N;
    a.Nk = last ? a.N / 2 + 1 : a.N;
    out_vol *= a.Lout;
    in_vol *= a.L;
    k_vol *= a.K;
    inv_scale *= (double)a.N;
  }

  // ---- structure
  pl->N1 = pl->N2 = 0;
  if (nd == 1) {
    // Lines of 2048 ... 8192 points take the four-step layout too (64 x 32 ... 128): its column and warp-engine passes run at
    // 4 - 5 TB/s where the generic block-level pass reaches ~1 TB/s on such lines, and the five specialised launches beat the
    // three generic ones even on small calls (profiles/r2b_short_split_probe.txt: faster in all twelve probed shapes, 16 - 47 %).
    const int64_t N0 = pl->ax[0].N;
    const bool short_split = !(P.flags & (FC_FLAG_NO_SHORT_SPLIT | FC_FLAG_NO_FUSED | FC_FLAG_NO_FAST_C2C | FC_FLAG_NO_FAST_R2C | FC_FLAG_NO_FAST_C2R)) &&
                             N0 >= 2048;
    if (N0 <= kMaxRealLine && !short_split) {
      pl->structure = FC_S_1D;
    } else {
      pl->structure = FC_S_1D_SPLIT;
      // Four-step split N = N1 * N2. Up to N = 64 * 2048 the strided pass is 64 points long: it then runs one thread
      // per column entirely in registers (fc_column.cuh) and the contiguous pass on the warp engine (N2 = 256..2048).
      // Longer lines fall back to a balanced split on the generic kernels.
      const int N = pl->ax[0].N, l2 = ilog2(N);
      pl->N1 = (N / 64 >= (short_split ? 32 : 256) && N / 64 <= 2048) ? 64 : 1 << ((l2 + 1) / 2);
      pl->N2 = N / pl->N1;
      pl->ax[0].Nk = 0;  // bins are (N1/2+1) x N2 for the split layout
    }
  } else {
    pl->structure = (nd == 2) ? FC_S_2D : FC_S_3D;
  }
  if (pl->structure == FC_S_1D_SPLIT)
    bins = (int64_t)(pl->N1 / 2 + 1) * pl->N2;
  else {
    bins = 1;
    for (int i = 0; i < nd; ++i) bins *= pl->ax[i].Nk;
  }
  int tw_len = 2;
  for (int i = 0; i < nd; ++i) tw_len = std::max(tw_len, pl->ax[i].N);
  if (pl->structure == FC_S_1D_SPLIT) tw_len = std::max(pl->N1, pl->N2);
  pl->tw_len = tw_len;

  // ---- forward programs
  const int Ig = P.cin / P.groups, Og = P.cout / P.groups;
  SrcDesc sig;
  sig.n_outer = (int64_t)P.batch * P.cin;
  sig.o_c2 = 1;
  sig.o_q = 1;
  sig.o_sA = in_vol;
  sig.o_sB = 0;
  sig.o_sC = 0;
  sig.conj = false;
  sig.scale = 1.f;
  sig.segmented_x = nd == 2 && pl->ax[1].seg_n > 1;
  SrcDesc ker;
  ker.n_outer = (int64_t)P.cout * Ig;
  ker.o_c2 = Ig;
  if (!P.transposed) {  // weight (Cout, Cin/g, K...)
    ker.o_q = 1;
    ker.o_sA = (int64_t)Ig * k_vol;
    ker.o_sB = 0;
    ker.o_sC = k_vol;
  } else {  // weight (Cin, Cout/g, K...): (co, ci_local) -> w[grp*Ig + ci_local][co_local]  (functional.py:109-114)
    ker.o_q = Og;
    ker.o_sA = (int64_t)Ig * Og * k_vol;
    ker.o_sB = k_vol;
    ker.o_sC = (int64_t)Og * k_vol;
  }
  ker.segmented_x = false;
  ker.conj = !P.transposed;
  ker.scale = (float)(1.0 / inv_scale);
  for (int i = 0; i < nd; ++i) {
    sig.imap[i] = pl->ax[i].imap_sig;
    sig.L[i] = pl->ax[i].L;
    ker.imap[i] = pl->ax[i].imap_ker;
    ker.L[i] = pl->ax[i].K;
  }
  pl->sig_fwd.clear();
  pl->ker_fwd.clear();
  pl->inv.clear();
  build_forward(*pl, sig, pl->sig_fwd);
  build_forward(*pl, ker, pl->ker_fwd);
  build_inverse(*pl, pl->inv);

  pl->contract.bins = bins;
  pl->contract.batch = P.batch;
  pl->contract.cin = P.cin;
  pl->contract.cout = P.cout;
  pl->contract.groups = P.groups;

  // ---- sizes
  fc_plan_info& I = pl->info;
  I.ndim = nd;
  for (int i = 0; i < nd; ++i) {
    I.out_size[i] = pl->ax[i].Lout;
    I.fft_size[i] = pl->ax[i].N;
  }
  I.bins = bins;
  I.out_elems = (int64_t)P.batch * P.cout * out_vol;
  I.xspec_bytes = (int64_t)P.batch * P.cin * bins * 8;
  I.kspec_bytes = (int64_t)P.cout * Ig * bins * 8;
  I.yspec_bytes = (int64_t)P.batch * P.cout * bins * 8;
  I.segments = (nd == 2) ? pl->ax[0].seg_n * pl->ax[1].seg_n : 1;
  if (I.segments > 1)  // the fused kernel writes all dense rows of the first axis into the product-spectrum buffer
    I.yspec_bytes = std::max<int64_t>(I.yspec_bytes, (int64_t)P.batch * P.cout * pl->inv[0].pass.R * pl->inv[0].pass.n_out * 8);
  int64_t sA = 0, sB = 0;
  auto scan = [&](const std::vector<fc_step>& v) {
    for (const fc_step& s : v) {
      if (s.dst == FC_BUF_SA) sA = std::max(sA, step_out_bytes(s));
      if (s.dst == FC_BUF_SB) sB = std::max(sB, step_out_bytes(s));
    }
  };
  scan(pl->sig_fwd);
  scan(pl->ker_fwd);
  scan(pl->inv);
  pl->off_xspec = 0;
  pl->off_yspec = align_up(pl->off_xspec + I.xspec_bytes, 256);
  pl->off_sA = align_up(pl->off_yspec + I.yspec_bytes, 256);
  pl->off_sB = align_up(pl->off_sA + sA, 256);
  pl->scratch_bytes = sA + sB;
  I.workspace_bytes = align_up(pl->off_sB + sB, 256) + 256;
  // tensor-core contraction (fc_tc.cuh): wide channel groups, small batch
  {
    const int Og = P.cout / P.groups;
    // batches run through the GEMM in chunks (fc_tc_chunk); the operand scratch is sized for one chunk, padded to a
    // multiple of 8
    const int bc = fc_tc_chunk(P.batch);
    const int bp = bc <= 8 ? 8 : (bc + 7) / 8 * 8;
    // (output groups that are multiples of 64 only run on 64-row A tiles at ~2.2 TB/s against 3.5 - 3.9 TB/s on full tiles: that
    // beats the SIMT contraction (0.8 - 1.3 TB/s at 64 channels) once there are enough bins to fill the pipeline,
    // profiles/r2b_tc_64row_tiles.txt)
    // With more bins the shared-memory-tiled SIMT contraction (fc_contract_tiled_kernel) is ahead again (profiles/r2b_contraction_paths.txt).
    pl->use_tc = !(P.flags & (FC_FLAG_NO_TC | FC_FLAG_NO_FUSED)) && Ig >= 32 && (2 * Ig) % 32 == 0 &&
                 (Og % 128 == 0 || (Og % 64 == 0 && bins >= 2048 && bins < 8192));
#ifdef FC_CPU_EMUL
    pl->use_tc = 0;  // tcgen05 cannot run in the host emulation
#endif
    pl->off_xtc = pl->off_ytc = 0;
    I.kspec_workspace_bytes = I.workspace_bytes;
    if (pl->use_tc) {
      pl->off_xtc = align_up(I.workspace_bytes, 256);
      const int64_t xtc = bins * P.groups * 2 * bp * 2 * Ig * 4;
      pl->off_ytc = align_up(pl->off_xtc + xtc, 256);
      const int64_t ytc = bins * P.cout * 2 * bp * 4;
      I.workspace_bytes = align_up(pl->off_ytc + ytc, 256) + 256;
      // building the cached kernel spectrum needs the pass-order spectrum as a temporary next to the usual scratch
      I.kspec_workspace_bytes = align_up(pl->off_sB + sB, 256) + 256 + align_up(I.kspec_bytes, 256);
      if (I.kspec_workspace_bytes < I.workspace_bytes) I.kspec_workspace_bytes = I.workspace_bytes;
    }
    I.tensor_core = pl->use_tc;
  }
  I.const_bytes = ((int64_t)tw_len + (pl->structure == FC_S_1D_SPLIT ? pl->N2 : 0)) * 8;
  I.n_launches = (int)(pl->sig_fwd.size() + 1 + pl->inv.size());
  I.n_launches_kspec = (int)pl->ker_fwd.size();
  I.fused = 0;
  I.algo_bytes_s1 = 4 * (int64_t)P.batch * P.cin * in_vol + I.xspec_bytes;
  I.algo_bytes_s2 = 4 * (int64_t)P.cout * Ig * k_vol + I.kspec_bytes;
  I.algo_bytes_s3 = I.xspec_bytes + I.kspec_bytes + I.yspec_bytes;
  I.algo_bytes_s4 = I.yspec_bytes + 4 * I.out_elems + 4 * (int64_t)P.cout;
  pl->pair = 0;
  fc_plan_build_program(pl);
  if (pl->pair) {
    // The packed batch-pair kernels keep both spectra as 16-byte slots {re0, re1, im0, im1} of the batch items 2p and
    // 2p + 1: an odd batch is rounded up to whole pairs. (The tensor-core operands never coexist with a fused program.)
    const int64_t bp2 = 2 * (int64_t)((P.batch + 1) / 2);
    I.xspec_bytes = I.xspec_bytes / P.batch * bp2;
    I.yspec_bytes = I.yspec_bytes / P.batch * bp2;
    // K1p writes and K4p reads the scratch buffers in the pair layout as well (whole pairs there too)
    sA = (sA + P.batch - 1) / P.batch * bp2;
    sB = (sB + P.batch - 1) / P.batch * bp2;
    if (pl->prog[1].fused.ystage > 0) {
      // K1p stores and K4p reads all N positions of the fused axis per bin (rows beyond the signal are zero), for every
      // row segment
      const fc_pass& k1p = pl->prog[0].pass;
      const fc_pass& k4p = pl->prog[2].pass;
      const int64_t need_x = k1p.n_outer * k1p.out_os * 16, need_y = k4p.n_outer * k4p.in_os * 16;
      if (pl->prog[0].dst == FC_BUF_SA) sA = std::max(sA, need_x);
      if (pl->prog[0].dst == FC_BUF_SB) sB = std::max(sB, need_x);
      if (pl->prog[0].dst == FC_BUF_SPEC) I.xspec_bytes = std::max(I.xspec_bytes, need_x);
      I.yspec_bytes = std::max(I.yspec_bytes, need_y);
    }
    pl->scratch_bytes = sA + sB;
    pl->off_yspec = align_up(pl->off_xspec + I.xspec_bytes, 256);
    pl->off_sA = align_up(pl->off_yspec + I.yspec_bytes, 256);
    pl->off_sB = align_up(pl->off_sA + sA, 256);
    I.workspace_bytes = align_up(pl->off_sB + sB, 256) + 256;
    I.kspec_workspace_bytes = I.workspace_bytes;
  }
  if (I.fused) {
    // The fused axis kernel reads the kernel spectrum bin-major: [group][line r][o_local][i][n], i.e. the Og*Ig channel
    // pairs of one (group, line) are adjacent lines (fc_fused_contract). The last pass of the kernel program writes it.
    fc_pass& k = pl->ker_fwd.back().pass;
    const int64_t OI = (int64_t)Og * Ig, Nl = k.n_out;
    k.out_oq = OI;
    k.out_osA = k.R * OI * Nl;
    k.out_os = Nl;
    k.out_rs = OI * Nl;
    if (pl->pair) {  // [group][line][o][i/2][n][i%2]: the values of two input channels of a bin are adjacent (fc_pair_contract)
      k.out_il = 2;
      k.out_es = 2;
      const int ys = pl->prog[1].fused.ystage;
      if (ys > 0) {  // [group][line][k1][o][i/2][k2][i%2], bin n = k1 + ys*k2 (fc_pair_fused64_kernel)
        const int S = pl->prog[1].fused.ystage_S;
        k.out_os = S;
        k.out_split = ys;
        k.out_split_stride = OI * S;
      }
    }
  }
  if (I.segments > 1 && !I.fused) return fail(FC_EUNSUPPORTED, "internal: segmented plan without the fused axis kernel");
  if (nd == 2 && pl->ax[1].seg_n > 1 && !((pl->prog.front().type == FC_L_FAST_R2C && pl->prog.back().type == FC_L_FAST_C2R) ||
                                       (pl->prog.front().type == FC_L_PAIR_R2C && pl->prog.back().type == FC_L_PAIR_C2R)))
    return fail(FC_EUNSUPPORTED, "internal: segmented rows without the transposing row kernels");
  return FC_OK;
}
}  // namespace

namespace {

const char* kKindName[] = {"r2c", "c2c_fwd", "c2c_inv", "c2r"};

int64_t pass_bytes(const fc_pass& p) {
  const int64_t lines = p.n_outer * p.R;
  const int64_t in_el = (p.kind == FC_R2C) ? 4 : 8, out_el = (p.kind == FC_C2R) ? 4 : 8;
  // four-step 1-d layout: the real side of the R2C / C2R pass is one tensor of n_in / n_out elements per outer item
  const int64_t in_b = (p.kind == FC_R2C && p.pos_r) ? p.n_outer * (int64_t)p.imap.L * in_el : lines * (int64_t)p.n_in * in_el;
  const int64_t out_b = (p.kind == FC_C2R && p.pos_r) ? p.n_outer * (int64_t)p.omap.Lout * out_el : lines * (int64_t)p.n_out * out_el;
  // overlap-save segments along the line (K1 / K4): the spectrum side holds seg_n half spectra per line
  if (p.seg_n > 1) return p.kind == FC_R2C ? in_b + out_b * p.seg_n : p.kind == FC_C2R ? in_b * p.seg_n + out_b : in_b + out_b;
  return in_b + out_b;
}

bool fast_line_len(int M) { return M == 32 || M == 64 || M == 128 || M == 256 || M == 512 || M == 1024; }

// Contiguous complex lines on both sides, a length the warp engine has: the pass can run on fc_fast_c2c_kernel.
bool fast_c2c_ok(const fc_pass& p) {
  return (p.kind == FC_C2C_FWD || p.kind == FC_C2C_INV) && !p.in_rfast && !p.out_rfast && p.in_es == 1 && p.out_es == 1 && !p.twiddle &&
         p.pos_n == 1 && p.pos_r == 0 && p.N >= 32 && p.N <= 2048;
}

// The strided pass of the four-step split with 64-point columns: one thread per column (fc_column.cuh).
bool column_pass_ok(const fc_pass& p) {
  if (p.N != 64 || !p.twiddle || !p.in_rfast || !p.out_rfast || p.pos_r != 1 || p.row_og > 1) return false;
  if (p.kind == FC_R2C) return p.in_rs == 0 && p.in_es == 1 && p.out_rs == 1;
  return p.kind == FC_C2R && p.in_rs == 1 && p.out_rs == 0 && p.out_es == 1;
}

// The real pass of a one-pass 1-d program (contiguous lines of 512 or 1024 real points on both sides): fc_line.cuh.
bool line_pass_ok(const fc_pass& p) {
  return (p.M == 256 || p.M == 512) && p.R == 1 && p.flat && p.in_es == 1 && p.out_es == 1 && !p.in_rfast && !p.out_rfast && !p.twiddle &&
         p.pos_n == 1 && p.pos_r == 0 && p.out_oq == 0 && p.seg_n <= 1;
}

bool plane_len(int n) { return n == 32 || n == 64 || n == 128; }
bool plain_gather(const fc_imap& m) { return m.mode == FC_PAD_CONSTANT && m.up == 1 && m.sub == 1; }
bool plain_crop(const fc_omap& m) { return m.og == 1 && m.os == 1 && m.ob >= 0; }

// 3-d program: the y pass (contiguous lines in, transposing store) and the z pass (contiguous) of the forward
// transform can run as one plane kernel when both extents are 32 or 64 and the gather maps are plain.
bool plane_fwd_ok(const fc_pass& a, const fc_pass& b, fc_plane_desc& d) {
  if (a.kind != FC_C2C_FWD || b.kind != FC_C2C_FWD || a.in_rfast || !a.out_rfast || b.in_rfast || b.out_rfast) return false;
  if (!plane_len(a.N) || !plane_len(b.N) || a.in_es != 1 || b.in_es != 1 || b.out_es != 1 || a.twiddle || b.twiddle) return false;
  if (!plain_gather(a.imap) || !plain_gather(b.imap) || a.pos_n != 1 || a.pos_r != 0 || b.pos_n != 1 || b.pos_r != 0) return false;
  const int Ly = a.imap.L, Lz = b.imap.L;
  if (b.R % a.N || a.in_rs != Ly || b.out_rs != b.N || a.n_outer != b.n_outer) return false;
  const int64_t nkx = b.R / a.N;
  if (a.R != nkx * Lz || a.in_os != nkx * Lz * Ly || a.scale != 1.f || a.conj_out) return false;
  std::memset(&d, 0, sizeof(d));
  d.ny = a.N;
  d.nz = b.N;
  d.nkx = (int32_t)nkx;
  d.conj_out = b.conj_out;
  d.scale = b.scale;
  d.n_outer = a.n_outer;
  d.in_os = a.in_os;
  d.out_os = b.out_os;
  d.imy = a.imap;
  d.imz = b.imap;
  return true;
}

// ... and likewise the z pass (contiguous) and the y pass (transposing load, contiguous cropped store) of the inverse.
bool plane_inv_ok(const fc_pass& c, const fc_pass& e, fc_plane_desc& d) {
  if (c.kind != FC_C2C_INV || e.kind != FC_C2C_INV || c.in_rfast || c.out_rfast || !e.in_rfast || e.out_rfast) return false;
  if (!plane_len(c.N) || !plane_len(e.N) || c.in_es != 1 || c.out_es != 1 || e.out_es != 1 || c.twiddle || e.twiddle) return false;
  if (!plain_crop(c.omap) || !plain_crop(e.omap) || c.pos_n != 1 || c.pos_r != 0 || e.pos_n != 1 || e.pos_r != 0) return false;
  const int Oy = e.omap.Lout, Oz = c.omap.Lout;
  if (c.R % e.N || c.in_rs != c.N || e.out_rs != Oy || c.n_outer != e.n_outer) return false;
  const int64_t nkx = c.R / e.N;
  if (e.R != nkx * Oz || e.out_os != nkx * Oz * Oy) return false;
  std::memset(&d, 0, sizeof(d));
  d.ny = e.N;
  d.nz = c.N;
  d.nkx = (int32_t)nkx;
  d.scale = 1.f;
  d.n_outer = c.n_outer;
  d.in_os = c.in_os;
  d.out_os = e.out_os;
  d.omy = e.omap;
  d.omz = c.omap;
  return true;
}

// Re-tile a pass for the transposing fast kernels: 16 lines per tile (128-byte segments on the transposed side;
// 32-line tiles measured the same on B200) for M >= 256; shorter lines are handled by groups of M/8 lanes, 2 lines per
// group and 8 warps, i.e. 4096/M lines per tile (fc_fast_tile_lines). Tiles never straddle an outer item.
void retile16(fc_pass& p) {
  const int T = fc_fast_tile_lines(p.M);
  p.T = T;
  p.log2T = ilog2(T);
  p.flat = 0;
  p.tiles_per_outer = (p.R + T - 1) / T * p.seg_n;  // tiles of one outer item: segment-major, then rows
  p.n_tiles = p.tiles_per_outer * p.n_outer;
}

}  // namespace

void fc_plan_build_program(fc_plan* pl) {
  const fc_problem& P = pl->prob;
  const int flags = P.flags;
  const bool allow = !(flags & FC_FLAG_NO_FUSED);
  pl->prog.clear();
  const int nf = (int)pl->sig_fwd.size(), ni = (int)pl->inv.size();
  const int Ig = P.cin / P.groups, Og = P.cout / P.groups;

  // fused middle: last forward pass + contraction + first inverse pass
  bool fuse_mid = allow && !(flags & FC_FLAG_NO_FUSED_MID) && nf >= 2 && ni >= 2;
  if (fuse_mid) {
    const fc_pass& f = pl->sig_fwd[nf - 1].pass;
    const fc_pass& b = pl->inv[0].pass;
    fuse_mid = f.kind == FC_C2C_FWD && b.kind == FC_C2C_INV && !f.in_rfast && !f.out_rfast && !b.in_rfast && !b.out_rfast && f.N == b.N &&
               (f.N == 256 || f.N == 512 || f.N == 1024) && Ig <= 16 && Og <= 16 && f.R == b.R && f.in_es == 1 && b.out_es == 1;
  }
  const fc_axis* seg_ax = (pl->structure == FC_S_2D && pl->ax[0].seg_n > 1) ? &pl->ax[0] : nullptr;
  const bool allow_plane = allow && !(flags & FC_FLAG_NO_FAST_C2C) && !fuse_mid && pl->structure == FC_S_3D && nf == 3 && ni == 3;
  for (int i = 0; i < nf; ++i) {
    if (fuse_mid && i == nf - 1) break;
    fc_launch L;
    std::memset(&L.plane, 0, sizeof(L.plane));
    if (allow_plane && i == 1 && plane_fwd_ok(pl->sig_fwd[1].pass, pl->sig_fwd[2].pass, L.plane)) {
      L.type = FC_L_PLANE_FWD;
      L.pass = pl->sig_fwd[1].pass;
      L.src = pl->sig_fwd[1].src;
      L.dst = pl->sig_fwd[2].dst;
      L.spec_is_y = 0;
      std::memset(&L.fused, 0, sizeof(L.fused));
      L.name = "plane_fwd_" + std::to_string(L.plane.ny) + "x" + std::to_string(L.plane.nz);
      L.bytes = 8 * (pl->sig_fwd[1].pass.R * pl->sig_fwd[1].pass.n_in + pl->sig_fwd[2].pass.R * pl->sig_fwd[2].pass.n_out) * L.plane.n_outer;
      pl->prog.push_back(L);
      break;
    }
    L.type = FC_L_PASS;
    L.pass = pl->sig_fwd[i].pass;
    L.src = pl->sig_fwd[i].src;
    L.dst = pl->sig_fwd[i].dst;
    L.spec_is_y = 0;
    std::memset(&L.fused, 0, sizeof(L.fused));
    const fc_pass& p = L.pass;
    if (i == 0 && allow && !(flags & FC_FLAG_NO_FAST_R2C) && p.kind == FC_R2C && !p.in_rfast && p.out_rfast && !p.twiddle && fast_line_len(p.M) &&
        p.imap.mode == FC_PAD_CONSTANT && !(p.imap.pad & 1) && p.imap.up == 1 && p.imap.sub == 1 && p.imap.ext == p.imap.L + 2 * p.imap.pad &&
        !(p.imap.L & 1) && !(p.in_rs & 1) &&
        !(p.o_sA & 1) && !(p.o_sB & 1) && !(p.o_sC & 1) && p.scale == 1.f && !p.conj_out && p.pos_n == 1 && p.pos_r == 0) {
      L.type = FC_L_FAST_R2C;
      retile16(L.pass);
    } else if (allow && !(flags & FC_FLAG_NO_FAST_C2C) && fast_c2c_ok(p)) {
      L.type = FC_L_FAST_C2C;
    } else if (allow && !(flags & FC_FLAG_NO_FAST_R2C) && p.kind == FC_R2C && column_pass_ok(p)) {
      L.type = FC_L_COL_R2C;
    } else if (allow && !(flags & FC_FLAG_NO_FAST_R2C) && p.kind == FC_R2C && line_pass_ok(p) && plain_gather(p.imap)) {
      L.type = FC_L_LINE_R2C;
    }
    L.name = L.type == FC_L_FAST_R2C   ? "fast_r2c_N" + std::to_string(p.N)
             : L.type == FC_L_LINE_R2C ? "line_r2c_N" + std::to_string(p.N)
             : L.type == FC_L_COL_R2C  ? "col_r2c_N" + std::to_string(p.N)
             : L.type == FC_L_FAST_C2C ? "fast_c2c_fwd_N" + std::to_string(p.N)
                                       : std::string("fwd_") + kKindName[p.kind] + "_N" + std::to_string(p.N);
    L.bytes = pass_bytes(p);
    pl->prog.push_back(L);
  }
  if (fuse_mid) {
    const fc_step& fs = pl->sig_fwd[nf - 1];
    const fc_step& bs = pl->inv[0];
    fc_launch L;
    L.type = FC_L_FUSED;
    L.pass = fs.pass;
    L.src = fs.src;
    L.dst = FC_BUF_SPEC;  // the product-spectrum buffer is free in the fused program; fs.src and bs.dst may alias
    L.spec_is_y = 1;
    L.fused.N = fs.pass.N;
    L.fused.ystage = 0;
    L.fused.ystage_S = 0;
    L.fused.fill_rows = 0;
    L.fused.n_in = fs.pass.n_in;
    L.fused.n_out = bs.pass.n_out;
    L.fused.n_seg = seg_ax ? seg_ax->seg_n : 1;
    L.fused.seg_V = seg_ax ? seg_ax->seg_V : fs.pass.N;
    L.fused.seg_off = seg_ax ? seg_ax->seg_off : 0;
    L.fused.ci = (Ig <= 8 && Og <= 8) ? 8 : 16;
    const int64_t items = (int64_t)P.batch * L.fused.n_seg;  // (batch, segment) pairs per bin
    L.fused.nb = (items >= 2 && fs.pass.N * L.fused.ci <= 4096) ? 2 : 1;
    L.fused.warps = 8;
    L.fused.occ = fs.pass.N * L.fused.nb * L.fused.ci <= 8192 ? 2 : 1;
    {
      const fc_imap& im = fs.pass.imap;
      const fc_omap& om = bs.pass.omap;
      L.fused.plain = im.mode == FC_PAD_CONSTANT && im.pad == 0 && im.up == 1 && im.sub == 1 && im.L >= fs.pass.N && im.ext >= fs.pass.N &&
                      om.og == 1 && om.os == 1 && om.ob == 0 && om.Lout <= om.lim && Ig == L.fused.ci && Og == L.fused.ci && L.fused.n_seg == 1;
    }
    if (const char* tune = fc_tune_str("TUNE")) {  // A/B timing knobs: "nb=1,warps=4"
      const char* q;
      if (L.fused.ci == 8) {
      if ((q = std::strstr(tune, "nb="))) L.fused.nb = std::atoi(q + 3) >= 2 && fs.pass.N <= 512 && items >= 2 ? 2 : 1;
      if ((q = std::strstr(tune, "warps="))) L.fused.warps = (std::atoi(q + 6) == 4 && L.fused.plain && fs.pass.N <= 512) ? 4 : 8;
      if (L.fused.warps == 4 && fs.pass.N == 256 && L.fused.nb == 1) L.fused.warps = 8;
      L.fused.occ = fs.pass.N * L.fused.nb <= 1024 ? (L.fused.warps == 4 ? 3 : 2) : 1;
      if ((q = std::strstr(tune, "occ=")) && L.fused.plain && fs.pass.N == 512 && L.fused.warps == 8) {
        const int o = std::atoi(q + 4);
        if (o == 3 || (o == 4 && L.fused.nb == 1)) L.fused.occ = o;
      }
      }
    }
    L.fused.R = fs.pass.R;
    L.fused.Rk = pl->structure == FC_S_2D ? pl->ax[1].Nk : fs.pass.R;
    L.fused.imap = fs.pass.imap;
    L.fused.omap = bs.pass.omap;
    L.name = "fused_axis_N" + std::to_string(fs.pass.N) + (L.fused.n_seg > 1 ? "_seg" + std::to_string(L.fused.n_seg) : "");
    L.bytes = 8 * ((int64_t)P.batch * P.cin * fs.pass.R * fs.pass.n_in + (int64_t)P.cout * Ig * L.fused.Rk * fs.pass.N +
                   (int64_t)P.batch * P.cout * bs.pass.R * bs.pass.n_out);
    pl->prog.push_back(L);
  } else if (pl->use_tc) {
    const char* names[3] = {"tc_relayout_x", "tc_gemm_3xtf32", "tc_relayout_y"};
    const int bc = fc_tc_chunk(P.batch);
    for (int b0 = 0; b0 < P.batch; b0 += bc) {  // batch chunks: pass.in_os = first batch, pass.n_outer = batches of the chunk
      const int nb = P.batch - b0 < bc ? P.batch - b0 : bc;
      const int bp = nb <= 8 ? 8 : (nb + 7) / 8 * 8;
      const int64_t xtc = pl->info.bins * P.groups * 2 * bp * 2 * Ig * 4, ytc = pl->info.bins * P.cout * 2 * bp * 4;
      const int64_t xs = pl->info.xspec_bytes / P.batch * nb, ys = pl->info.yspec_bytes / P.batch * nb;
      const int64_t bytes[3] = {xs + xtc, pl->info.kspec_bytes + xtc + ytc, ytc + ys};
      for (int j = 0; j < 3; ++j) {
        fc_launch L;
        L.type = FC_L_TC_X + j;
        std::memset(&L.pass, 0, sizeof(L.pass));
        std::memset(&L.fused, 0, sizeof(L.fused));
        L.pass.in_os = b0;
        L.pass.n_outer = nb;
        L.src = L.dst = FC_BUF_SPEC;
        L.spec_is_y = 0;
        L.name = std::string(names[j]) + (P.batch > bc ? "_b" + std::to_string(b0) : "");
        L.bytes = bytes[j];
        pl->prog.push_back(L);
      }
    }
  } else {
    fc_launch L;
    L.type = FC_L_CONTRACT;
    std::memset(&L.pass, 0, sizeof(L.pass));
    std::memset(&L.fused, 0, sizeof(L.fused));
    L.src = L.dst = FC_BUF_SPEC;
    L.spec_is_y = 0;
    L.name = "contract";
    L.bytes = pl->info.algo_bytes_s3;
    pl->prog.push_back(L);
  }
  for (int i = fuse_mid ? 1 : 0; i < ni; ++i) {
    fc_launch L;
    std::memset(&L.plane, 0, sizeof(L.plane));
    if (allow_plane && i == 0 && plane_inv_ok(pl->inv[0].pass, pl->inv[1].pass, L.plane)) {
      L.type = FC_L_PLANE_INV;
      L.pass = pl->inv[0].pass;
      L.src = pl->inv[0].src;
      L.dst = pl->inv[1].dst;
      L.spec_is_y = 1;
      std::memset(&L.fused, 0, sizeof(L.fused));
      L.name = "plane_inv_" + std::to_string(L.plane.ny) + "x" + std::to_string(L.plane.nz);
      L.bytes = 8 * (pl->inv[0].pass.R * pl->inv[0].pass.n_in + pl->inv[1].pass.R * pl->inv[1].pass.n_out) * L.plane.n_outer;
      pl->prog.push_back(L);
      ++i;  // the y pass is part of the plane kernel
      continue;
    }
    L.type = FC_L_PASS;
    L.pass = pl->inv[i].pass;
    L.src = (fuse_mid && i == 1) ? FC_BUF_SPEC : pl->inv[i].src;
    L.dst = pl->inv[i].dst;
    L.spec_is_y = 1;
    std::memset(&L.fused, 0, sizeof(L.fused));
    const fc_pass& p = L.pass;
    if (i == ni - 1 && allow && !(flags & FC_FLAG_NO_FAST_C2R) && p.kind == FC_C2R && p.in_rfast && !p.out_rfast && !p.twiddle && fast_line_len(p.M) &&
        p.pos_n == 1 && p.pos_r == 0 && p.out_es == 1) {
      L.type = FC_L_FAST_C2R;
      retile16(L.pass);
    } else if (allow && !(flags & FC_FLAG_NO_FAST_C2C) && fast_c2c_ok(p)) {
      L.type = FC_L_FAST_C2C;
    } else if (allow && !(flags & FC_FLAG_NO_FAST_C2R) && p.kind == FC_C2R && column_pass_ok(p)) {
      L.type = FC_L_COL_C2R;
    } else if (allow && !(flags & FC_FLAG_NO_FAST_C2R) && p.kind == FC_C2R && line_pass_ok(p) && p.omap.og == 1 && p.omap.os == 1 && p.row_og <= 1) {
      L.type = FC_L_LINE_C2R;
    }
    L.name = L.type == FC_L_FAST_C2R   ? "fast_c2r_N" + std::to_string(p.N)
             : L.type == FC_L_LINE_C2R ? "line_c2r_N" + std::to_string(p.N)
             : L.type == FC_L_COL_C2R  ? "col_c2r_N" + std::to_string(p.N)
             : L.type == FC_L_FAST_C2C ? "fast_c2c_inv_N" + std::to_string(p.N)
                                       : std::string("inv_") + kKindName[p.kind] + "_N" + std::to_string(p.N);
    L.bytes = pass_bytes(p);
    pl->prog.push_back(L);
  }
  // Tensor-core contraction with a single batch chunk between two contiguous complex passes: those passes write / read
  // the GEMM's operand layouts themselves (fc_tc.cuh), and the two relayout kernels and their round trips go away.
  if (pl->use_tc && allow && !(flags & FC_FLAG_NO_FAST_C2C) && Ig % 32 == 0 && fc_tc_chunk(P.batch) >= P.batch) {
    int ix = -1;
    for (size_t i = 0; i < pl->prog.size(); ++i)
      if (pl->prog[i].type == FC_L_TC_X) ix = (int)i;
    if (ix >= 1 && ix + 3 < (int)pl->prog.size() && pl->prog[ix + 1].type == FC_L_TC_GEMM && pl->prog[ix + 2].type == FC_L_TC_Y) {
      fc_launch& F = pl->prog[ix - 1];
      fc_launch& X = pl->prog[ix];
      fc_launch& Y = pl->prog[ix + 2];
      fc_launch& V = pl->prog[ix + 3];
      const fc_pass& f = F.pass;
      const fc_pass& v = V.pass;
      auto tile_len = [](int n) { return n == 256 || n == 512; };
      const bool f_ok = F.type == FC_L_FAST_C2C && f.kind == FC_C2C_FWD && tile_len(f.N) && plain_gather(f.imap) && f.imap.pad == 0 &&
                        f.imap.L >= f.N && f.imap.ext >= f.N && f.scale == 1.f && !f.conj_out && f.out_os == f.R * (int64_t)f.N &&
                        f.out_rs == f.N && f.n_outer == (int64_t)P.batch * P.cin;
      const bool v_ok = V.type == FC_L_FAST_C2C && v.kind == FC_C2C_INV && tile_len(v.N) && v.omap.og == 1 && v.omap.os == 1 && v.omap.ob == 0 &&
                        v.omap.Lout == v.N && v.omap.lim >= v.N && v.in_os == v.R * (int64_t)v.N && v.in_rs == v.N &&
                        v.n_outer == (int64_t)P.batch * P.cout && v.R == f.R && v.N == f.N;
      if (f_ok && v_ok) {
        F.type = FC_L_TC_FWD;
        F.name = "tc_c2c_fwd_N" + std::to_string(f.N);
        F.bytes = F.bytes / 2 + X.bytes - pl->info.xspec_bytes;  // lines in, Bt blobs out
        Y.type = FC_L_TC_INV;
        Y.pass = v;
        Y.src = V.src;
        Y.dst = V.dst;
        Y.spec_is_y = V.spec_is_y;
        Y.name = "tc_c2c_inv_N" + std::to_string(v.N);
        Y.bytes = Y.bytes - pl->info.yspec_bytes + V.bytes / 2;  // product in, lines out
        pl->prog.erase(pl->prog.begin() + ix + 3);
        pl->prog.erase(pl->prog.begin() + ix);
      }
    }
  }
  pl->info.n_launches = (int)pl->prog.size();
  pl->info.fused = fuse_mid ? 1 : 0;
  // bias-only rows of a row lattice: written by the fused kernel (DRAM idle there) instead of the last one (store-bound)
  if (fuse_mid && pl->prog.size() == 3 && pl->prog[2].type == FC_L_FAST_C2R && pl->prog[2].pass.row_og > 1 && !(flags & FC_FLAG_NO_ROW_FILL)) {
    pl->prog[1].fused.fill_rows = 1;
    pl->prog[2].pass.row_fill_skip = 1;
  }

  // ---- packed batch pairs (fc_pair.cuh): the whole fused 2-d program K1 -> KB -> K4 on 16-byte slots
  pl->pair = 0;
  if (fuse_mid && !(flags & FC_FLAG_NO_PAIR) && pl->prog.size() == 3 && pl->prog[0].type == FC_L_FAST_R2C && pl->prog[2].type == FC_L_FAST_C2R &&
      !pl->use_tc) {
    fc_launch& A = pl->prog[0];
    fc_launch& Bk = pl->prog[1];
    fc_launch& C = pl->prog[2];
    const int N = Bk.fused.N;
    auto pair_len = [](int M) { return M == 128 || M == 256 || M == 512 || M == 1024; };
    // opt-in (FC_FLAG_PAIR): at BASELINE c2 the whole step measured 138.6 us on this program (y stage, S = 128) against 132.1 us on
    // the one-line kernels, although its fused kernel is faster (57 / 44 us against 71 us): profiles/r2_c2_paths_same_box.txt
    const bool want = (flags & FC_FLAG_PAIR) != 0;
    const bool ok = want && Ig == Og && (Ig == 8 || Ig == 16) && pair_len(A.pass.M) && pair_len(C.pass.M) && (int64_t)N * Ig * 16 <= 128 * 1024;
    if (ok) {
      pl->pair = 1;
      const int BP = (P.batch + 1) / 2;
      auto retile_pair = [&](fc_pass& p, int channels) {
        const int T = fc_pair_tile_lines(p.M);
        p.T = T;
        p.log2T = ilog2(T);
        p.n_outer = (int64_t)BP * channels;
        p.tiles_per_outer = (p.R + T - 1) / T * p.seg_n;
        p.n_tiles = p.tiles_per_outer * p.n_outer;
      };
      A.type = FC_L_PAIR_R2C;
      retile_pair(A.pass, P.cin);
      A.name = "pair_r2c_N" + std::to_string(A.pass.N);
      Bk.type = FC_L_PAIR_FUSED;
      const int64_t items = (int64_t)BP * Bk.fused.n_seg;
      Bk.fused.ci = Ig;
      Bk.fused.nb = (N * Ig <= 2048 && items >= 2) ? 2 : 1;  // pair items per CTA
      Bk.fused.warps = 8;
      Bk.fused.occ = (int64_t)Bk.fused.nb * N * Ig * 16 <= 64 * 1024 ? 2 : 1;
      if (const char* t = fc_tune_str("PAIRKB")) {  // tuning builds: "np,warps,occ" (must name an instantiation of fc_api.cu)
        int np = Bk.fused.nb, wp = 8, oc = Bk.fused.occ;
        std::sscanf(t, "%d,%d,%d", &np, &wp, &oc);
        if (np <= items) {
          Bk.fused.nb = np;
          Bk.fused.warps = wp;
          Bk.fused.occ = oc;
        }
      }
      Bk.name = "pair_fused_N" + std::to_string(N) + (Bk.fused.n_seg > 1 ? "_seg" + std::to_string(Bk.fused.n_seg) : "");
      C.type = FC_L_PAIR_C2R;
      retile_pair(C.pass, P.cout);
      C.name = "pair_c2r_N" + std::to_string(C.pass.N);
      // y stage: N = 64*YS; K1p / K4p run the radix-YS stage of the fused axis, the fused kernel 64-point sub-problems
      // that share the kernel spectrum across eight batch items (fc_pair_fused64_kernel)
      // N = YS * S: radix YS <= 4 keeps the stage in K1p / K4p cheap (64-byte runs of 4 adjacent n2 per tile, a radix-4
      // butterfly); measured at BASELINE c2: S = 64 / YS = 8 makes K1p 63 us and K4p 47 us against 44 / 33 us
      int S = N >= 512 ? 128 : 64;
      if (const char* t = fc_tune_str("YSS")) S = std::atoi(t) == 64 ? 64 : 128;
      if (N / S < 4) S = 64;
      const int YS = N / S;
      auto ys_row = [](int M) { return M == 128 || M == 256; };  // K1p / K4p variants with 16-line tiles
      Bk.fused.ystage = 0;
      // identity gather on the fused axis (rows beyond the signal are zero), plain crop, one segment
      const fc_imap& imy = Bk.fused.imap;
      const fc_omap& omy = Bk.fused.omap;
      const bool plain_y = imy.mode == FC_PAD_CONSTANT && imy.pad == 0 && imy.up == 1 && imy.sub == 1 && omy.og == 1 && omy.os == 1 && omy.ob == 0 &&
                           omy.Lout <= omy.lim && Bk.fused.n_seg == 1;
      if (!(flags & FC_FLAG_NO_YSTAGE) && plain_y && (YS == 8 || YS == 4) && Ig == 8 && ys_row(A.pass.M) && ys_row(C.pass.M) &&
          A.pass.R <= N && C.pass.R <= N && C.pass.row_og == 1) {
        Bk.fused.ystage = YS;
        Bk.fused.ystage_S = S;
        Bk.type = FC_L_PAIR_FUSED64;
        Bk.name = "pair_fused64_N" + std::to_string(N);
        auto ys_pass = [&](fc_pass& p, bool fwd) {
          p.ystage = YS;
          p.ystage_N = N;
          p.ystage_S = S;
          p.T = 16;
          p.log2T = 4;
          p.tiles_per_outer = (int64_t)(N / 16) * p.seg_n;
          p.n_tiles = p.tiles_per_outer * p.n_outer;
          const int64_t lines = (int64_t)(p.M + 1) * p.seg_n;  // bins of this kernel's own axis per pair image
          if (fwd) {
            p.out_es = N;
            p.out_os = lines * N;
          } else {
            p.in_es = N;
            p.in_os = lines * N;
          }
        };
        ys_pass(A.pass, true);
        ys_pass(C.pass, false);
        A.name += "_ys" + std::to_string(YS);
        C.name += "_ys" + std::to_string(YS);
      }
    }
  }
}

std::string fc_plan_to_string(const fc_plan* pl) {
  std::ostringstream os;
  const fc_problem& P = pl->prob;
  if (pl->bseg_n > 1)
    os << "batch segments: user problem B=" << pl->user_prob.batch << " L=" << pl->user_prob.in_size[0] << " pad=" << pl->user_prob.padding[0]
       << " -> " << pl->bseg_n << " windows of " << pl->ax[0].N << " points, " << pl->prog.back().pass.bseg_Vo << " outputs each, run as:\n";
  os << "fft_conv" << (P.transposed ? "_transpose" : "") << " nd=" << pl->nd << " B=" << P.batch << " Cin=" << P.cin
     << " Cout=" << P.cout << " groups=" << P.groups << " structure=" << pl->structure;
  if (pl->structure == FC_S_1D_SPLIT) os << " N1=" << pl->N1 << " N2=" << pl->N2;
  os << " bins=" << pl->info.bins << " tw_len=" << pl->tw_len << "\n";
  for (int i = 0; i < pl->nd; ++i) {
    const fc_axis& a = pl->ax[i];
    os << "  axis" << i << ": L=" << a.L << " K=" << a.K << " s=" << a.stride << " p=" << a.pad << " d=" << a.dil
       << " op=" << a.opad << " g=" << a.g << " N=" << a.N << " Nk=" << a.Nk << " Lout=" << a.Lout << " imap(ext=" << a.imap_sig.ext
       << ",up=" << a.imap_sig.up << ",sub=" << a.imap_sig.sub << ",pad=" << a.imap_sig.pad << ") omap(os=" << a.omap.os
       << ",ob=" << a.omap.ob << ",og=" << a.omap.og << ",lim=" << a.omap.lim << ")";
    if (a.seg_n > 1) os << " segments(n=" << a.seg_n << ",V=" << a.seg_V << ",off=" << a.seg_off << ",N_full=" << a.N_full << ")";
    os << "\n";
  }
  auto dump = [&](const char* name, const std::vector<fc_step>& v) {
    for (size_t i = 0; i < v.size(); ++i) {
      const fc_pass& p = v[i].pass;
      static const char* kn[] = {"R2C", "C2C_FWD", "C2C_INV", "C2R"};
      os << "  " << name << "[" << i << "] " << kn[p.kind] << " N=" << p.N << " M=" << p.M << " T=" << p.T << " outer=" << p.n_outer
         << " R=" << p.R << " tiles=" << p.n_tiles << " flat=" << p.flat << " in(os=" << p.in_os << ",rs=" << p.in_rs
         << ",es=" << p.in_es << ",rfast=" << p.in_rfast << ") out(os=" << p.out_os << ",rs=" << p.out_rs << ",es=" << p.out_es
         << ",rfast=" << p.out_rfast << ") tw=" << p.twiddle << " conj=" << p.conj_out << " scale=" << p.scale << " buf " << v[i].src
         << "->" << v[i].dst << "\n";
    }
  };
  dump("sig", pl->sig_fwd);
  dump("ker", pl->ker_fwd);
  dump("inv", pl->inv);
  for (const fc_launch& L : pl->prog) os << "  launch " << L.name << " type=" << L.type << " bytes=" << L.bytes << "\n";
  os << "  fused=" << pl->info.fused << " workspace=" << pl->info.workspace_bytes << " launches=" << pl->info.n_launches << "\n";
  return os.str();
}

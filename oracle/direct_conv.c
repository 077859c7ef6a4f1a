/*
 * direct_conv.c — CPU oracle #2 for the fft_conv hot path: TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The convolution the reference's FFT path must equal, evaluated straight from its definition with double
 * accumulators (the reference's own tests pin fft_conv to exactly this quantity via
 * torch.nn.functional.conv{n}d / conv_transpose{n}d: reference tests/test_functional.py:33-59,
 * tests/test_functional_transpose.py:33-59).
 *
 *   forward     y[b,o,j] = bias[o] + sum_{i in group(o)} sum_m xpad[b,i, j*s + m*d] * w[o,i_local,m]
 *               xpad = F.pad(x, p, mode)                                   (reference functional.py:44-62, 76-87)
 *   transposed  y[b,o,j] = bias[o] + sum_{i in group(o)} sum_{q,m: q*t + m*d = j + p} x[b,i,q] * w[i,o_local,m]
 *                                                                          (reference functional.py:103-174)
 * Spatial arrays always have 3 entries (leading singleton axes for 1-d / 2-d). O(output * kernel) — small cases only.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu-baseline leg may load this.
 */
#include <stddef.h>

static int pad_index(int v, int L, int mode) {
  if (v >= 0 && v < L) return v;
  switch (mode) {
    case 0: return -1;                               /* constant (zeros) */
    case 1: return v < 0 ? -v : 2 * (L - 1) - v;     /* reflect */
    case 2: return v < 0 ? 0 : L - 1;                /* replicate */
    default: return v < 0 ? v + L : v - L;           /* circular */
  }
}

int fc_oracle_direct_conv(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int Cout, int groups,
                          const int* L, const int* K, const int* stride, const int* pad, const int* dil, int pad_mode) {
  int Ig = Cin / groups, Og = Cout / groups;
  int Lo[3];
  for (int a = 0; a < 3; ++a) {
    int Lp = L[a] + 2 * pad[a], Kd = (K[a] - 1) * dil[a] + 1;
    if (Lp < Kd) return -1;
    Lo[a] = (Lp - Kd) / stride[a] + 1;
  }
  size_t xvol = (size_t)L[0] * L[1] * L[2], kvol = (size_t)K[0] * K[1] * K[2], yvol = (size_t)Lo[0] * Lo[1] * Lo[2];
  for (int b = 0; b < B; ++b)
    for (int o = 0; o < Cout; ++o) {
      int g = o / Og;
      for (int j0 = 0; j0 < Lo[0]; ++j0)
        for (int j1 = 0; j1 < Lo[1]; ++j1)
          for (int j2 = 0; j2 < Lo[2]; ++j2) {
            double acc = bias ? (double)bias[o] : 0.0;
            for (int il = 0; il < Ig; ++il) {
              const float* xp = x + ((size_t)b * Cin + (size_t)g * Ig + il) * xvol;
              const float* wp = w + ((size_t)o * Ig + il) * kvol;
              for (int m0 = 0; m0 < K[0]; ++m0) {
                int u0 = pad_index(j0 * stride[0] + m0 * dil[0] - pad[0], L[0], pad_mode);
                if (u0 < 0) continue;
                for (int m1 = 0; m1 < K[1]; ++m1) {
                  int u1 = pad_index(j1 * stride[1] + m1 * dil[1] - pad[1], L[1], pad_mode);
                  if (u1 < 0) continue;
                  for (int m2 = 0; m2 < K[2]; ++m2) {
                    int u2 = pad_index(j2 * stride[2] + m2 * dil[2] - pad[2], L[2], pad_mode);
                    if (u2 < 0) continue;
                    acc += (double)xp[((size_t)u0 * L[1] + u1) * L[2] + u2] * (double)wp[((size_t)m0 * K[1] + m1) * K[2] + m2];
                  }
                }
              }
            }
            y[((size_t)b * Cout + o) * yvol + ((size_t)j0 * Lo[1] + j1) * Lo[2] + j2] = (float)acc;
          }
    }
  return 0;
}

int fc_oracle_direct_conv_transpose(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int Cout, int groups,
                                    const int* L, const int* K, const int* stride, const int* pad, const int* dil, const int* opad) {
  int Ig = Cin / groups, Og = Cout / groups;
  int Lo[3];
  for (int a = 0; a < 3; ++a) {
    Lo[a] = (L[a] - 1) * stride[a] - 2 * pad[a] + dil[a] * (K[a] - 1) + opad[a] + 1;
    if (Lo[a] < 1) return -1;
  }
  size_t xvol = (size_t)L[0] * L[1] * L[2], kvol = (size_t)K[0] * K[1] * K[2], yvol = (size_t)Lo[0] * Lo[1] * Lo[2];
  for (int b = 0; b < B; ++b)
    for (int o = 0; o < Cout; ++o) {
      int g = o / Og, ol = o % Og;
      for (int j0 = 0; j0 < Lo[0]; ++j0)
        for (int j1 = 0; j1 < Lo[1]; ++j1)
          for (int j2 = 0; j2 < Lo[2]; ++j2) {
            double acc = bias ? (double)bias[o] : 0.0;
            for (int il = 0; il < Ig; ++il) {
              int i = g * Ig + il;
              const float* xp = x + ((size_t)b * Cin + i) * xvol;
              const float* wp = w + ((size_t)i * Og + ol) * kvol; /* weight layout (Cin, Cout/groups, K...) */
              for (int m0 = 0; m0 < K[0]; ++m0) {
                int t0 = j0 + pad[0] - m0 * dil[0];
                if (t0 < 0 || t0 % stride[0] || t0 / stride[0] >= L[0]) continue;
                for (int m1 = 0; m1 < K[1]; ++m1) {
                  int t1 = j1 + pad[1] - m1 * dil[1];
                  if (t1 < 0 || t1 % stride[1] || t1 / stride[1] >= L[1]) continue;
                  for (int m2 = 0; m2 < K[2]; ++m2) {
                    int t2 = j2 + pad[2] - m2 * dil[2];
                    if (t2 < 0 || t2 % stride[2] || t2 / stride[2] >= L[2]) continue;
                    acc += (double)xp[((size_t)(t0 / stride[0]) * L[1] + t1 / stride[1]) * L[2] + t2 / stride[2]] *
                           (double)wp[((size_t)m0 * K[1] + m1) * K[2] + m2];
                  }
                }
              }
            }
            y[((size_t)b * Cout + o) * yvol + ((size_t)j0 * Lo[1] + j1) * Lo[2] + j2] = (float)acc;
          }
    }
  return 0;
}

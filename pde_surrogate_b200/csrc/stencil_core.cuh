// stencil_core.cuh — per-thread "column strip" arithmetic of the fused Darcy-loss kernels.
//
// Everything here is __host__ __device__ on plain pointers so that the exact index math the
// sm_100a kernels execute can also be driven thread-by-thread on a CPU
// (tests/host_emul/stencil_emul.cu) and compared with the oracle without a GPU.
//
// Operators (utils/image_gradient.py:50-92 of the reference, filter_size=3):
//   Dx f = (W/8) * S_y (x) d_x      Dy f = (H/8) * S_x (x) d_y
// S = [1,2,1] smoothing with replicate boundary; d = central difference with replicate
// boundary, whose first/last row is replaced by the 3-point one-sided difference
// [-3,4,-1] / [1,-4,3] when `correct` (the reference's `modifier` matrix, lines 43-46).
//
// Work split: a thread owns 4 consecutive columns (one float4) and a run of R rows and
// slides a 3-row register window down its strip.
#pragma once
#include "common.cuh"

namespace pdes {
namespace stencil {

struct Geom {
  int H, W, W4, nrs, R;
};

PDES_HD Geom make_geom(int H, int W, int nthreads) {
  Geom g;
  g.H = H;
  g.W = W;
  g.W4 = W / 4;
  int nrs = nthreads / g.W4;
  if (nrs < 1) nrs = 1;
  if (nrs > H) nrs = H;
  g.nrs = nrs;
  g.R = (H + nrs - 1) / nrs;
  return g;
}

PDES_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// six words f[y][4cs-1 .. 4cs+4], replicate (clamped) in both directions
PDES_HD void load6_clamp(const float* plane, int y, int cs, int H, int W, float e[6]) {
  y = clampi(y, 0, H - 1);
  const float* r = plane + (size_t)y * W;
  const int x0 = 4 * cs;
  const float4 c = *reinterpret_cast<const float4*>(r + x0);
  e[1] = c.x;
  e[2] = c.y;
  e[3] = c.z;
  e[4] = c.w;
  e[0] = r[x0 > 0 ? x0 - 1 : 0];
  e[5] = r[x0 + 4 < W ? x0 + 4 : W - 1];
}

// same, zero-extended outside the image (used by the adjoint)
PDES_HD void load6_zero(const float* plane, int y, int cs, int H, int W, float e[6]) {
  if (y < 0 || y >= H) {
#pragma unroll
    for (int j = 0; j < 6; ++j) e[j] = 0.f;
    return;
  }
  const float* r = plane + (size_t)y * W;
  const int x0 = 4 * cs;
  const float4 c = *reinterpret_cast<const float4*>(r + x0);
  e[1] = c.x;
  e[2] = c.y;
  e[3] = c.z;
  e[4] = c.w;
  e[0] = x0 > 0 ? r[x0 - 1] : 0.f;
  e[5] = x0 + 4 < W ? r[x0 + 4] : 0.f;
}

// d_x applied to one row window: h[k] for columns 4cs+k
PDES_HD void hdiff(const float e[6], int cs, int W4, bool correct, float h[4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) h[k] = e[k + 2] - e[k];
  if (correct) {
    if (cs == 0) h[0] = 4.f * (e[2] - e[1]) - (e[3] - e[1]);
    if (cs == W4 - 1) h[3] = 4.f * (e[4] - e[3]) - (e[4] - e[2]);
  }
}

// d_x^T applied to one zero-extended row window
PDES_HD void hdiff_T(const float e[6], int cs, int W4, bool correct, float t[4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) t[k] = e[k] - e[k + 2];
  if (cs == 0) {
    const float p0 = e[1];
    if (correct) {
      t[0] -= 3.f * p0;
      t[1] += 3.f * p0;
      t[2] -= p0;
    } else {
      t[0] -= p0;
    }
  }
  if (cs == W4 - 1) {
    const float pl = e[4];
    if (correct) {
      t[1] += pl;
      t[2] -= 3.f * pl;
      t[3] += 3.f * pl;
    } else {
      t[3] += pl;
    }
  }
}

struct FwdPartial {
  float c, d, dir, neu;
};

// Forward residual sums of one thread's strip.  If P planes are non-null the (scaled)
// residual fields needed by the backward pass are also written:
//   P1 = a*K*r1, P2 = a*K*r2, P3 = b*r3 (row-masked), Q1 = a*r1, Q2 = a*r2.
PDES_HD FwdPartial fwd_strip(const float* Kp, const float* up, const float* s1p,
                             const float* s2p, int H, int W, int tid, int nthreads,
                             bool correct, bool use_tb, float a, float b, float* P1,
                             float* P2, float* P3, float* Q1, float* Q2) {
  FwdPartial acc;
  acc.c = acc.d = acc.dir = acc.neu = 0.f;
  const Geom g = make_geom(H, W, nthreads);
  const int cs = tid % g.W4, rs = tid / g.W4;
  if (rs >= g.nrs) return acc;
  const int y0 = rs * g.R;
  const int y1 = (y0 + g.R < H) ? y0 + g.R : H;
  if (y0 >= y1) return acc;
  const float cW = (float)W * 0.125f, cH = (float)H * 0.125f;
  const bool hasK = (Kp != nullptr);

  float eu[3][6], e1[3][6], e2[3][6];
  load6_clamp(up, y0 - 1, cs, H, W, eu[0]);
  load6_clamp(s1p, y0 - 1, cs, H, W, e1[0]);
  load6_clamp(s2p, y0 - 1, cs, H, W, e2[0]);
  load6_clamp(up, y0, cs, H, W, eu[1]);
  load6_clamp(s1p, y0, cs, H, W, e1[1]);
  load6_clamp(s2p, y0, cs, H, W, e2[1]);

  for (int y = y0; y < y1; ++y) {
    load6_clamp(up, y + 1, cs, H, W, eu[2]);
    load6_clamp(s1p, y + 1, cs, H, W, e1[2]);
    load6_clamp(s2p, y + 1, cs, H, W, e2[2]);

    float dxu[4], dxs1[4];
    {
      float h0[4], h1[4], h2[4];
      hdiff(eu[0], cs, g.W4, correct, h0);
      hdiff(eu[1], cs, g.W4, correct, h1);
      hdiff(eu[2], cs, g.W4, correct, h2);
#pragma unroll
      for (int k = 0; k < 4; ++k) dxu[k] = cW * (h0[k] + 2.f * h1[k] + h2[k]);
      hdiff(e1[0], cs, g.W4, correct, h0);
      hdiff(e1[1], cs, g.W4, correct, h1);
      hdiff(e1[2], cs, g.W4, correct, h2);
#pragma unroll
      for (int k = 0; k < 4; ++k) dxs1[k] = cW * (h0[k] + 2.f * h1[k] + h2[k]);
    }
    float vu[6], v2[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      vu[j] = eu[2][j] - eu[0][j];
      v2[j] = e2[2][j] - e2[0][j];
    }
    if (correct && (y == 0 || y == H - 1)) {
      float xu[6], x2[6];
      const int yy = (y == 0) ? 2 : H - 3;
      load6_clamp(up, yy, cs, H, W, xu);
      load6_clamp(s2p, yy, cs, H, W, x2);
      if (y == 0) {
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          vu[j] = 4.f * (eu[2][j] - eu[1][j]) - (xu[j] - eu[1][j]);
          v2[j] = 4.f * (e2[2][j] - e2[1][j]) - (x2[j] - e2[1][j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          vu[j] = 4.f * (eu[1][j] - eu[0][j]) - (eu[1][j] - xu[j]);
          v2[j] = 4.f * (e2[1][j] - e2[0][j]) - (e2[1][j] - x2[j]);
        }
      }
    }
    float kk[4] = {0.f, 0.f, 0.f, 0.f};
    if (hasK) {
      const float4 kv = *reinterpret_cast<const float4*>(Kp + (size_t)y * W + 4 * cs);
      kk[0] = kv.x;
      kk[1] = kv.y;
      kk[2] = kv.z;
      kk[3] = kv.w;
    }
    const bool row_in = use_tb || (y >= 1 && y <= H - 2);
    float p1[4], p2[4], p3[4], q1[4], q2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float dyu = cH * (vu[k] + 2.f * vu[k + 1] + vu[k + 2]);
      const float dys2 = cH * (v2[k] + 2.f * v2[k + 1] + v2[k + 2]);
      float r1 = 0.f, r2 = 0.f;
      if (hasK) {
        r1 = e1[1][k + 1] + kk[k] * dxu[k];
        r2 = e2[1][k + 1] + kk[k] * dyu;
      }
      const float r3 = row_in ? (dxs1[k] + dys2) : 0.f;
      acc.c += r1 * r1 + r2 * r2;
      acc.d += r3 * r3;
      p1[k] = a * kk[k] * r1;
      p2[k] = a * kk[k] * r2;
      p3[k] = b * r3;
      q1[k] = a * r1;
      q2[k] = a * r2;
    }
    if (cs == 0) {
      const float t = eu[1][1] - 1.f;
      acc.dir += t * t;
    }
    if (cs == g.W4 - 1) acc.dir += eu[1][4] * eu[1][4];
    if (y == 0 || y == H - 1) {
#pragma unroll
      for (int k = 0; k < 4; ++k) acc.neu += e2[1][k + 1] * e2[1][k + 1];
    }
    if (P1 != nullptr) {
      const size_t o = (size_t)y * W + 4 * cs;
      *reinterpret_cast<float4*>(P1 + o) = make_float4(p1[0], p1[1], p1[2], p1[3]);
      *reinterpret_cast<float4*>(P2 + o) = make_float4(p2[0], p2[1], p2[2], p2[3]);
      *reinterpret_cast<float4*>(P3 + o) = make_float4(p3[0], p3[1], p3[2], p3[3]);
      *reinterpret_cast<float4*>(Q1 + o) = make_float4(q1[0], q1[1], q1[2], q1[3]);
      *reinterpret_cast<float4*>(Q2 + o) = make_float4(q2[0], q2[1], q2[2], q2[3]);
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      eu[0][j] = eu[1][j];
      eu[1][j] = eu[2][j];
      e1[0][j] = e1[1][j];
      e1[1][j] = e1[2][j];
      e2[0][j] = e2[1][j];
      e2[1][j] = e2[2][j];
    }
  }
  return acc;
}

// d_y^T of a plane at row y for the six window columns (zero-extended), before S_x^T.
PDES_HD void vdiff_T(const float* P, const float eprev[6], const float enext[6], int y, int cs,
                     int H, int W, bool correct, float t[6]) {
#pragma unroll
  for (int j = 0; j < 6; ++j) t[j] = eprev[j] - enext[j];
  const bool near_top = (y <= 2), near_bot = (y >= H - 3);
  if (near_top) {
    float r0[6];
    load6_zero(P, 0, cs, H, W, r0);
    float c = 0.f;
    if (correct) {
      if (y == 0) c += -3.f;
      if (y == 1) c += 3.f;
      if (y == 2) c += -1.f;
    } else {
      if (y == 0) c += -1.f;
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) t[j] += c * r0[j];
  }
  if (near_bot) {
    float rl[6];
    load6_zero(P, H - 1, cs, H, W, rl);
    float c = 0.f;
    if (correct) {
      if (y == H - 3) c += 1.f;
      if (y == H - 2) c += -3.f;
      if (y == H - 1) c += 3.f;
    } else {
      if (y == H - 1) c += 1.f;
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) t[j] += c * rl[j];
  }
}

// Second backward pass: applies the adjoint operators to the residual planes and adds the
// boundary-loss gradients.  du/ds1/ds2 may alias up/s1p/s2p (only own-position reads).
PDES_HD void bwd_strip_pass2(const float* P1, const float* P2, const float* P3, const float* Q1,
                             const float* Q2, const float* up, const float* s2p, float* du,
                             float* ds1, float* ds2, int H, int W, int tid, int nthreads,
                             bool correct, float cdir, float cneu) {
  const Geom g = make_geom(H, W, nthreads);
  const int cs = tid % g.W4, rs = tid / g.W4;
  if (rs >= g.nrs) return;
  const int y0 = rs * g.R;
  const int y1 = (y0 + g.R < H) ? y0 + g.R : H;
  if (y0 >= y1) return;
  const float cW = (float)W * 0.125f, cH = (float)H * 0.125f;

  float tx1[3][4], tx3[3][4];  // d_x^T rows of P1, P3
  float e2[3][6], e3[3][6];    // zero-extended rows of P2, P3
  {
    float e[6];
    load6_zero(P1, y0 - 1, cs, H, W, e);
    hdiff_T(e, cs, g.W4, correct, tx1[0]);
    load6_zero(P1, y0, cs, H, W, e);
    hdiff_T(e, cs, g.W4, correct, tx1[1]);
    load6_zero(P3, y0 - 1, cs, H, W, e3[0]);
    hdiff_T(e3[0], cs, g.W4, correct, tx3[0]);
    load6_zero(P3, y0, cs, H, W, e3[1]);
    hdiff_T(e3[1], cs, g.W4, correct, tx3[1]);
    load6_zero(P2, y0 - 1, cs, H, W, e2[0]);
    load6_zero(P2, y0, cs, H, W, e2[1]);
  }
  for (int y = y0; y < y1; ++y) {
    {
      float e[6];
      load6_zero(P1, y + 1, cs, H, W, e);
      hdiff_T(e, cs, g.W4, correct, tx1[2]);
    }
    load6_zero(P3, y + 1, cs, H, W, e3[2]);
    hdiff_T(e3[2], cs, g.W4, correct, tx3[2]);
    load6_zero(P2, y + 1, cs, H, W, e2[2]);

    const float edge = ((y == 0) ? 1.f : 0.f) + ((y == H - 1) ? 1.f : 0.f);
    float ox1[4], ox3[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ox1[k] = tx1[0][k] + (2.f + edge) * tx1[1][k] + tx1[2][k];
      ox3[k] = tx3[0][k] + (2.f + edge) * tx3[1][k] + tx3[2][k];
    }
    float ty2[6], ty3[6];
    vdiff_T(P2, e2[0], e2[2], y, cs, H, W, correct, ty2);
    vdiff_T(P3, e3[0], e3[2], y, cs, H, W, correct, ty3);

    const size_t o = (size_t)y * W + 4 * cs;
    const float4 q1 = *reinterpret_cast<const float4*>(Q1 + o);
    const float4 q2 = *reinterpret_cast<const float4*>(Q2 + o);
    const float4 uu = *reinterpret_cast<const float4*>(up + o);
    const float4 ss = *reinterpret_cast<const float4*>(s2p + o);
    const float q1a[4] = {q1.x, q1.y, q1.z, q1.w};
    const float q2a[4] = {q2.x, q2.y, q2.z, q2.w};
    const float ua[4] = {uu.x, uu.y, uu.z, uu.w};
    const float sa[4] = {ss.x, ss.y, ss.z, ss.w};
    float gu[4], g1[4], g2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int x = 4 * cs + k;
      const float xe = ((x == 0) ? 1.f : 0.f) + ((x == W - 1) ? 1.f : 0.f);
      const float oy2 = ty2[k] + (2.f + xe) * ty2[k + 1] + ty2[k + 2];
      const float oy3 = ty3[k] + (2.f + xe) * ty3[k + 1] + ty3[k + 2];
      gu[k] = cW * ox1[k] + cH * oy2;
      g1[k] = q1a[k] + cW * ox3[k];
      g2[k] = q2a[k] + cH * oy3;
      if (x == 0) gu[k] += cdir * (ua[k] - 1.f);
      if (x == W - 1) gu[k] += cdir * ua[k];
      if (y == 0 || y == H - 1) g2[k] += cneu * sa[k];
    }
    *reinterpret_cast<float4*>(du + o) = make_float4(gu[0], gu[1], gu[2], gu[3]);
    *reinterpret_cast<float4*>(ds1 + o) = make_float4(g1[0], g1[1], g1[2], g1[3]);
    *reinterpret_cast<float4*>(ds2 + o) = make_float4(g2[0], g2[1], g2[2], g2[3]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      tx1[0][k] = tx1[1][k];
      tx1[1][k] = tx1[2][k];
      tx3[0][k] = tx3[1][k];
      tx3[1][k] = tx3[2][k];
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      e2[0][j] = e2[1][j];
      e2[1][j] = e2[2][j];
      e3[0][j] = e3[1][j];
      e3[1][j] = e3[2][j];
    }
  }
}


// =======================================================================================
// Exact-fit fast path: every thread owns R rows (compile-time) x 4 columns, the strips tile the
// image exactly (H % R == 0, nthreads == (H/R)*(W/4)) and `correct` is true.  Same operators as
// fwd_strip / bwd_strip_pass2 above, but the R+2 row window is loaded once into registers and the
// row loop is fully unrolled, so each horizontal difference is formed once (the rolling version
// recomputes it three times and pays a register rotation per row) and the top/bottom one-sided
// rows come out of the window instead of extra loads.  Measured effect on B200: see DESIGN.md §4.
// =======================================================================================
PDES_HD void load6_cols(const float* r, int x0, int xl, int xr, float e[6]) {
  const float4 c = *reinterpret_cast<const float4*>(r + x0);
  e[1] = c.x;
  e[2] = c.y;
  e[3] = c.z;
  e[4] = c.w;
  e[0] = r[xl];
  e[5] = r[xr];
}
// zero-extended columns: the halo words are always loaded from an in-range (clamped) address and
// then masked, which keeps the loads unconditional (no branches around predicated LDS)
PDES_HD void load6_zcols(const float* r, int x0, int xl, int xr, float ml, float mr, float e[6]) {
  const float4 c = *reinterpret_cast<const float4*>(r + x0);
  e[1] = c.x;
  e[2] = c.y;
  e[3] = c.z;
  e[4] = c.w;
  e[0] = ml * r[xl];
  e[5] = mr * r[xr];
}
PDES_HD void zero6(float e[6]) {
#pragma unroll
  for (int j = 0; j < 6; ++j) e[j] = 0.f;
}
// d_x with the one-sided first/last column (first/last are thread-uniform flags)
PDES_HD void hdiff_c(const float e[6], bool first, bool last, float h[4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) h[k] = e[k + 2] - e[k];
  if (first) h[0] = 4.f * (e[2] - e[1]) - (e[3] - e[1]);
  if (last) h[3] = 4.f * (e[4] - e[3]) - (e[4] - e[2]);
}
PDES_HD void hdiff_Tc(const float e[6], bool first, bool last, float t[4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) t[k] = e[k] - e[k + 2];
  if (first) {
    const float p0 = e[1];
    t[0] -= 3.f * p0;
    t[1] += 3.f * p0;
    t[2] -= p0;
  }
  if (last) {
    const float pl = e[4];
    t[1] += pl;
    t[2] -= 3.f * pl;
    t[3] += 3.f * pl;
  }
}

template <int R, bool WRITE, bool SUMS>
PDES_HD FwdPartial fwd_strip_r(const float* Kp, const float* up, const float* s1p, const float* s2p,
                               int H, int W, int cs, int y0, bool use_tb, float a, float b,
                               float* P1, float* P2, float* P3, float* Q1, float* Q2, float* qreg = nullptr) {
  // qreg != nullptr: Q1 / Q2 (only ever read back at the thread's OWN pixels by the adjoint pass) stay in the
  // caller's thread-private array qreg[2][R][4] (registers) instead of making a shared-memory round trip
  static_assert(R >= 2, "the one-sided boundary rows must lie inside the window");
  FwdPartial acc;
  acc.c = acc.d = acc.dir = acc.neu = 0.f;
  const int W4 = W >> 2;
  const bool first = (cs == 0), last = (cs == W4 - 1);
  const bool top = (y0 == 0), bot = (y0 + R == H);
  const int x0 = 4 * cs;
  const int xl = first ? 0 : x0 - 1, xr = last ? W - 1 : x0 + 4;
  const float cW = (float)W * 0.125f, cH = (float)H * 0.125f;
  const bool hasK = (Kp != nullptr);
  // window row r <-> image row clamp(y0 - 1 + r)
  const int rtop = top ? 0 : y0 - 1, rbot = bot ? H - 1 : y0 + R;

  float dxu[R][4], dyu[R][4], dxs1[R][4], dys2[R][4];
  float uc[R][2], s1c[R][4], s2c[R][4];  // u at columns 4cs / 4cs+3 (Dirichlet), centre sigma values
  {
    float e[R + 2][6];
    load6_cols(up + (size_t)rtop * W, x0, xl, xr, e[0]);
#pragma unroll
    for (int r = 1; r <= R; ++r) load6_cols(up + (size_t)(y0 - 1 + r) * W, x0, xl, xr, e[r]);
    load6_cols(up + (size_t)rbot * W, x0, xl, xr, e[R + 1]);
    float h[R + 2][4];
#pragma unroll
    for (int r = 0; r < R + 2; ++r) hdiff_c(e[r], first, last, h[r]);
#pragma unroll
    for (int i = 0; i < R; ++i) {
      float v[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) v[j] = e[i + 2][j] - e[i][j];
      if (i == 0 && top) {
#pragma unroll
        for (int j = 0; j < 6; ++j) v[j] = 4.f * (e[2][j] - e[1][j]) - (e[3][j] - e[1][j]);
      }
      if (i == R - 1 && bot) {
#pragma unroll
        for (int j = 0; j < 6; ++j) v[j] = 4.f * (e[R][j] - e[R - 1][j]) - (e[R][j] - e[R - 2][j]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        dxu[i][k] = cW * (h[i][k] + 2.f * h[i + 1][k] + h[i + 2][k]);
        dyu[i][k] = cH * (v[k] + 2.f * v[k + 1] + v[k + 2]);
      }
      uc[i][0] = e[i + 1][1];
      uc[i][1] = e[i + 1][4];
    }
  }
  {
    float e[R + 2][6];
    load6_cols(s1p + (size_t)rtop * W, x0, xl, xr, e[0]);
#pragma unroll
    for (int r = 1; r <= R; ++r) load6_cols(s1p + (size_t)(y0 - 1 + r) * W, x0, xl, xr, e[r]);
    load6_cols(s1p + (size_t)rbot * W, x0, xl, xr, e[R + 1]);
    float h[R + 2][4];
#pragma unroll
    for (int r = 0; r < R + 2; ++r) hdiff_c(e[r], first, last, h[r]);
#pragma unroll
    for (int i = 0; i < R; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        dxs1[i][k] = cW * (h[i][k] + 2.f * h[i + 1][k] + h[i + 2][k]);
        s1c[i][k] = e[i + 1][k + 1];
      }
    }
  }
  {
    float e[R + 2][6];
    load6_cols(s2p + (size_t)rtop * W, x0, xl, xr, e[0]);
#pragma unroll
    for (int r = 1; r <= R; ++r) load6_cols(s2p + (size_t)(y0 - 1 + r) * W, x0, xl, xr, e[r]);
    load6_cols(s2p + (size_t)rbot * W, x0, xl, xr, e[R + 1]);
#pragma unroll
    for (int i = 0; i < R; ++i) {
      float v[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) v[j] = e[i + 2][j] - e[i][j];
      if (i == 0 && top) {
#pragma unroll
        for (int j = 0; j < 6; ++j) v[j] = 4.f * (e[2][j] - e[1][j]) - (e[3][j] - e[1][j]);
      }
      if (i == R - 1 && bot) {
#pragma unroll
        for (int j = 0; j < 6; ++j) v[j] = 4.f * (e[R][j] - e[R - 1][j]) - (e[R][j] - e[R - 2][j]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        dys2[i][k] = cH * (v[k] + 2.f * v[k + 1] + v[k + 2]);
        s2c[i][k] = e[i + 1][k + 1];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int y = y0 + i;
    float kk[4] = {0.f, 0.f, 0.f, 0.f};
    if (hasK) {
      const float4 kv = *reinterpret_cast<const float4*>(Kp + (size_t)y * W + x0);
      kk[0] = kv.x;
      kk[1] = kv.y;
      kk[2] = kv.z;
      kk[3] = kv.w;
    }
    const bool yedge = (i == 0 && top) || (i == R - 1 && bot);
    const bool row_in = use_tb || !yedge;
    float p1[4], p2[4], p3[4], q1[4], q2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float r1 = 0.f, r2 = 0.f;
      if (hasK) {
        r1 = s1c[i][k] + kk[k] * dxu[i][k];
        r2 = s2c[i][k] + kk[k] * dyu[i][k];
      }
      const float r3 = row_in ? (dxs1[i][k] + dys2[i][k]) : 0.f;
      if (SUMS) {
        acc.c += r1 * r1 + r2 * r2;
        acc.d += r3 * r3;
      }
      if (WRITE) {
        q1[k] = a * r1;
        q2[k] = a * r2;
        p1[k] = kk[k] * q1[k];
        p2[k] = kk[k] * q2[k];
        p3[k] = b * r3;
      }
    }
    if (SUMS) {
      if (first) {
        const float t = uc[i][0] - 1.f;
        acc.dir += t * t;
      }
      if (last) acc.dir += uc[i][1] * uc[i][1];
      if (yedge) {
#pragma unroll
        for (int k = 0; k < 4; ++k) acc.neu += s2c[i][k] * s2c[i][k];
      }
    }
    if (WRITE) {
      const size_t o = (size_t)y * W + x0;
      *reinterpret_cast<float4*>(P1 + o) = make_float4(p1[0], p1[1], p1[2], p1[3]);
      *reinterpret_cast<float4*>(P2 + o) = make_float4(p2[0], p2[1], p2[2], p2[3]);
      *reinterpret_cast<float4*>(P3 + o) = make_float4(p3[0], p3[1], p3[2], p3[3]);
      if (qreg != nullptr) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          qreg[i * 4 + k] = q1[k];
          qreg[(R + i) * 4 + k] = q2[k];
        }
      } else {
        *reinterpret_cast<float4*>(Q1 + o) = make_float4(q1[0], q1[1], q1[2], q1[3]);
        *reinterpret_cast<float4*>(Q2 + o) = make_float4(q2[0], q2[1], q2[2], q2[3]);
      }
    }
  }
  return acc;
}

// adjoint pass of the exact-fit path; du/ds1/ds2 may alias up/s1p/s2p (own-position reads only)
template <int R>
PDES_HD void bwd_strip_pass2_r(const float* P1, const float* P2, const float* P3, const float* Q1,
                               const float* Q2, const float* up, const float* s2p, float* du, float* ds1,
                               float* ds2, int H, int W, int cs, int y0, float cdir, float cneu,
                               const float* qreg = nullptr) {
  static_assert(R >= 2, "R >= 2");
  const int W4 = W >> 2;
  const bool first = (cs == 0), last = (cs == W4 - 1);
  const bool top = (y0 == 0), bot = (y0 + R == H);
  const int x0 = 4 * cs;
  const float cW = (float)W * 0.125f, cH = (float)H * 0.125f;
  const int xl = first ? x0 : x0 - 1, xr = last ? x0 + 3 : x0 + 4;
  const float ml = first ? 0.f : 1.f, mr = last ? 0.f : 1.f;

  float tx1[R + 2][4], tx3[R + 2][4], e2[R + 2][6], e3[R + 2][6];
#pragma unroll
  for (int r = 0; r < R + 2; ++r) {
    const int yy = y0 - 1 + r;
    const bool valid = !((r == 0 && top) || (r == R + 1 && bot));
    if (valid) {
      float e[6];
      load6_zcols(P1 + (size_t)yy * W, x0, xl, xr, ml, mr, e);
      hdiff_Tc(e, first, last, tx1[r]);
      load6_zcols(P3 + (size_t)yy * W, x0, xl, xr, ml, mr, e3[r]);
      hdiff_Tc(e3[r], first, last, tx3[r]);
      load6_zcols(P2 + (size_t)yy * W, x0, xl, xr, ml, mr, e2[r]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) tx1[r][k] = tx3[r][k] = 0.f;
      zero6(e2[r]);
      zero6(e3[r]);
    }
  }
  float ty2[R][6], ty3[R][6];
#pragma unroll
  for (int i = 0; i < R; ++i) {
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      ty2[i][j] = e2[i][j] - e2[i + 2][j];
      ty3[i][j] = e3[i][j] - e3[i + 2][j];
    }
  }
  // one-sided first/last image row of d_y: extra taps on rows 0 and H-1 for y <= 2 / y >= H-3
  if (y0 <= 2) {
    float r2[6], r3[6];
    load6_zcols(P2, x0, xl, xr, ml, mr, r2);
    load6_zcols(P3, x0, xl, xr, ml, mr, r3);
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int y = y0 + i;
      const float c = (y == 0) ? -3.f : (y == 1) ? 3.f : (y == 2) ? -1.f : 0.f;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        ty2[i][j] += c * r2[j];
        ty3[i][j] += c * r3[j];
      }
    }
  }
  if (y0 + R - 1 >= H - 3) {
    float r2[6], r3[6];
    load6_zcols(P2 + (size_t)(H - 1) * W, x0, xl, xr, ml, mr, r2);
    load6_zcols(P3 + (size_t)(H - 1) * W, x0, xl, xr, ml, mr, r3);
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int y = y0 + i;
      const float c = (y == H - 3) ? 1.f : (y == H - 2) ? -3.f : (y == H - 1) ? 3.f : 0.f;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        ty2[i][j] += c * r2[j];
        ty3[i][j] += c * r3[j];
      }
    }
  }
  const float wx0 = first ? 3.f : 2.f, wx3 = last ? 3.f : 2.f;
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int y = y0 + i;
    const bool yedge = (i == 0 && top) || (i == R - 1 && bot);
    const float wy = yedge ? 3.f : 2.f;
    const size_t o = (size_t)y * W + x0;
    float q1a[4], q2a[4];
    if (qreg != nullptr) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        q1a[k] = qreg[i * 4 + k];
        q2a[k] = qreg[(R + i) * 4 + k];
      }
    } else {
      const float4 q1 = *reinterpret_cast<const float4*>(Q1 + o);
      const float4 q2 = *reinterpret_cast<const float4*>(Q2 + o);
      q1a[0] = q1.x; q1a[1] = q1.y; q1a[2] = q1.z; q1a[3] = q1.w;
      q2a[0] = q2.x; q2a[1] = q2.y; q2a[2] = q2.z; q2a[3] = q2.w;
    }
    float gu[4], g1[4], g2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float wx = (k == 0) ? wx0 : (k == 3) ? wx3 : 2.f;
      const float ox1 = tx1[i][k] + wy * tx1[i + 1][k] + tx1[i + 2][k];
      const float ox3 = tx3[i][k] + wy * tx3[i + 1][k] + tx3[i + 2][k];
      const float oy2 = ty2[i][k] + wx * ty2[i][k + 1] + ty2[i][k + 2];
      const float oy3 = ty3[i][k] + wx * ty3[i][k + 1] + ty3[i][k + 2];
      gu[k] = cW * ox1 + cH * oy2;
      g1[k] = q1a[k] + cW * ox3;
      g2[k] = q2a[k] + cH * oy3;
    }
    if (first) gu[0] += cdir * (up[o] - 1.f);
    if (last) gu[3] += cdir * up[o + 3];
    if (yedge) {
      const float4 ss = *reinterpret_cast<const float4*>(s2p + o);
      g2[0] += cneu * ss.x;
      g2[1] += cneu * ss.y;
      g2[2] += cneu * ss.z;
      g2[3] += cneu * ss.w;
    }
    *reinterpret_cast<float4*>(du + o) = make_float4(gu[0], gu[1], gu[2], gu[3]);
    *reinterpret_cast<float4*>(ds1 + o) = make_float4(g1[0], g1[1], g1[2], g1[3]);
    *reinterpret_cast<float4*>(ds2 + o) = make_float4(g2[0], g2[1], g2[2], g2[3]);
  }
}

}  // namespace stencil
}  // namespace pdes

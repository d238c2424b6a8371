/*
 * darcy_oracle.c — plain-C, double-precision restatement of the Sobel stencils and the Darcy
 * mixed-residual loss of the reference, with the loss gradient obtained by transposing every
 * forward step (scatter).  TEST INFRASTRUCTURE ONLY: linked by tests/ and by the host-emulation
 * harness, never by the product library.
 *
 * Follows the reference literally (paths relative to the cics-nd/pde-surrogate root):
 *   utils/image_gradient.py:28-33   HSOBEL = [[-1,-2,-1],[0,0,0],[1,2,1]]/8, VSOBEL = HSOBEL^T
 *   utils/image_gradient.py:43-46   modifier = I; modifier[0:2,0]=[4,-1]; modifier[-2:,-1]=[-1,4]
 *   utils/image_gradient.py:68-73   grad_h: replicate pad 1 -> conv2d(VSOBEL) * W -> grad @ modifier
 *   utils/image_gradient.py:85-90   grad_v: replicate pad 1 -> conv2d(HSOBEL) * H -> modifier^T @ grad
 *   models/darcy.py:170-176         constitutive residual
 *   models/darcy.py:217-224         continuity residual (use_tb)
 *   models/darcy.py:227-233         boundary terms
 * Pinned against tests/golden/sobel_darcy.npz (tests/test_oracle_c.py).
 */
#include <stdlib.h>
#include <string.h>

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static const double HS[3][3] = {{-0.125, -0.25, -0.125}, {0, 0, 0}, {0.125, 0.25, 0.125}};

static double kern(int dir, int i, int j) { return dir == 0 ? HS[j][i] /* VSOBEL */ : HS[i][j]; }

/* g = conv(pad_rep(f), kernel) * scale */
static void conv_fwd(const double* f, double* g, int H, int W, int dir) {
  const double scale = dir == 0 ? (double)W : (double)H;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      double s = 0.0;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
          s += kern(dir, i, j) * f[clampi(y + i - 1, 0, H - 1) * W + clampi(x + j - 1, 0, W - 1)];
      g[y * W + x] = s * scale;
    }
}
static void conv_adj(const double* gbar, double* fbar, int H, int W, int dir) {
  const double scale = dir == 0 ? (double)W : (double)H;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x)
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
          fbar[clampi(y + i - 1, 0, H - 1) * W + clampi(x + j - 1, 0, W - 1)] +=
              kern(dir, i, j) * scale * gbar[y * W + x];
}
/* out = g @ M (dir 0) or M^T @ g (dir 1) */
static void modifier_fwd(const double* g, double* out, int H, int W, int dir) {
  memcpy(out, g, sizeof(double) * H * W);
  if (dir == 0) {
    for (int y = 0; y < H; ++y) {
      out[y * W + 0] = 4.0 * g[y * W + 0] - g[y * W + 1];
      out[y * W + W - 1] = 4.0 * g[y * W + W - 1] - g[y * W + W - 2];
    }
  } else {
    for (int x = 0; x < W; ++x) {
      out[0 * W + x] = 4.0 * g[0 * W + x] - g[1 * W + x];
      out[(H - 1) * W + x] = 4.0 * g[(H - 1) * W + x] - g[(H - 2) * W + x];
    }
  }
}
static void modifier_adj(const double* obar, double* gbar, int H, int W, int dir) {
  memcpy(gbar, obar, sizeof(double) * H * W);
  if (dir == 0) {
    for (int y = 0; y < H; ++y) {
      gbar[y * W + 0] = 4.0 * obar[y * W + 0];
      gbar[y * W + 1] += -obar[y * W + 0];
      gbar[y * W + W - 1] = 4.0 * obar[y * W + W - 1];
      gbar[y * W + W - 2] += -obar[y * W + W - 1];
    }
  } else {
    for (int x = 0; x < W; ++x) {
      gbar[0 * W + x] = 4.0 * obar[0 * W + x];
      gbar[1 * W + x] += -obar[0 * W + x];
      gbar[(H - 1) * W + x] = 4.0 * obar[(H - 1) * W + x];
      gbar[(H - 2) * W + x] += -obar[(H - 1) * W + x];
    }
  }
}

/* one plane: out = D_dir f */
void pdes_oracle_sobel_plane(const double* f, double* out, int H, int W, int dir, int correct) {
  double* g = (double*)malloc(sizeof(double) * H * W);
  conv_fwd(f, g, H, W, dir);
  if (correct)
    modifier_fwd(g, out, H, W, dir);
  else
    memcpy(out, g, sizeof(double) * H * W);
  free(g);
}
/* fbar += D_dir^T obar */
void pdes_oracle_sobel_plane_adj(const double* obar, double* fbar, int H, int W, int dir,
                                 int correct) {
  double* gbar = (double*)malloc(sizeof(double) * H * W);
  if (correct)
    modifier_adj(obar, gbar, H, W, dir);
  else
    memcpy(gbar, obar, sizeof(double) * H * W);
  conv_adj(gbar, fbar, H, W, dir);
  free(gbar);
}

void pdes_oracle_sobel(const double* img, double* out, long n_img, int H, int W, int dir,
                       int correct, int adjoint) {
  for (long n = 0; n < n_img; ++n) {
    if (!adjoint) {
      pdes_oracle_sobel_plane(img + n * H * W, out + n * H * W, H, W, dir, correct);
    } else {
      memset(out + n * H * W, 0, sizeof(double) * H * W);
      pdes_oracle_sobel_plane_adj(img + n * H * W, out + n * H * W, H, W, dir, correct);
    }
  }
}

/* K may be NULL (constitutive term skipped).  dout may be NULL (losses only).
 * dout = d/d(out) of sum_i gw[i]*loss4[i]. */
void pdes_oracle_darcy(const float* K, const float* out, int B, int H, int W, int use_tb,
                       const double gw[4], double loss4[4], double* dout) {
  const int HW = H * W;
  const double n_c = (double)B * HW, n_d = (double)B * (use_tb ? H : H - 2) * W;
  const double n_dir = (double)B * H, n_neu = (double)B * 2 * W;
  double acc[4] = {0, 0, 0, 0};
  double* buf = (double*)malloc(sizeof(double) * HW * 12);
  double *u = buf, *s1 = buf + HW, *s2 = buf + 2 * HW, *dxu = buf + 3 * HW, *dyu = buf + 4 * HW,
         *dxs1 = buf + 5 * HW, *dys2 = buf + 6 * HW, *t1 = buf + 7 * HW, *t2 = buf + 8 * HW,
         *t3 = buf + 9 * HW, *kk = buf + 10 * HW;
  for (int b = 0; b < B; ++b) {
    const float* ob = out + (size_t)b * 3 * HW;
    for (int p = 0; p < HW; ++p) {
      u[p] = ob[p];
      s1[p] = ob[HW + p];
      s2[p] = ob[2 * HW + p];
      kk[p] = K ? (double)K[(size_t)b * HW + p] : 0.0;
    }
    pdes_oracle_sobel_plane(u, dxu, H, W, 0, 1);
    pdes_oracle_sobel_plane(u, dyu, H, W, 1, 1);
    pdes_oracle_sobel_plane(s1, dxs1, H, W, 0, 1);
    pdes_oracle_sobel_plane(s2, dys2, H, W, 1, 1);
    double* gu = dout ? dout + (size_t)b * 3 * HW : NULL;
    if (gu) memset(gu, 0, sizeof(double) * 3 * HW);
    for (int p = 0; p < HW; ++p) {
      const int y = p / W, x = p % W;
      double r1 = 0, r2 = 0, r3 = 0;
      if (K) {
        r1 = s1[p] + kk[p] * dxu[p];
        r2 = s2[p] + kk[p] * dyu[p];
      }
      if (use_tb || (y >= 1 && y <= H - 2)) r3 = dxs1[p] + dys2[p];
      acc[0] += r1 * r1 + r2 * r2;
      acc[1] += r3 * r3;
      if (x == 0) acc[2] += (u[p] - 1.0) * (u[p] - 1.0) / n_dir;
      if (x == W - 1) acc[2] += u[p] * u[p] / n_dir;
      if (y == 0 || y == H - 1) acc[3] += s2[p] * s2[p];
      t1[p] = gw[0] * 2.0 / n_c * kk[p] * r1; /* d/d(Dx u) */
      t2[p] = gw[0] * 2.0 / n_c * kk[p] * r2; /* d/d(Dy u) */
      t3[p] = gw[1] * 2.0 / n_d * r3;         /* d/d(Dx s1), d/d(Dy s2) */
      if (gu) {
        gu[HW + p] += gw[0] * 2.0 / n_c * r1;
        gu[2 * HW + p] += gw[0] * 2.0 / n_c * r2;
        if (x == 0) gu[p] += gw[2] * 2.0 * (u[p] - 1.0) / n_dir;
        if (x == W - 1) gu[p] += gw[2] * 2.0 * u[p] / n_dir;
        if (y == 0 || y == H - 1) gu[2 * HW + p] += gw[3] * 2.0 * s2[p] / n_neu;
      }
    }
    if (gu) {
      pdes_oracle_sobel_plane_adj(t1, gu, H, W, 0, 1);
      pdes_oracle_sobel_plane_adj(t2, gu, H, W, 1, 1);
      pdes_oracle_sobel_plane_adj(t3, gu + HW, H, W, 0, 1);
      pdes_oracle_sobel_plane_adj(t3, gu + 2 * HW, H, W, 1, 1);
    }
  }
  loss4[0] = acc[0] / n_c;
  loss4[1] = acc[1] / n_d;
  loss4[2] = acc[2];
  loss4[3] = acc[3] / n_neu;
  free(buf);
}

// Host emulation of the tile kernels' per-thread strip arithmetic (stencil_core.cuh), thread by
// thread on the CPU, checked against the C oracle.  Lets the index math be verified without a GPU.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../pde_surrogate_b200/csrc/stencil_core.cuh"

extern "C" void pdes_oracle_darcy(const float* K, const float* out, int B, int H, int W, int use_tb,
                                  const double gw[4], double loss4[4], double* dout);

static double frand() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }

// fitR > 0: drive the exact-fit unrolled strips (fwd_strip_r / bwd_strip_pass2_r) instead
template <int R>
static void fast_sample(const float* st, int hasK, int H, int W, int use_tb, float a, float b, float cdir,
                        float cneu, float* scratch, double acc[4]) {
  using namespace pdes::stencil;
  const int HW = H * W, W4 = W / 4, nthreads = (H / R) * W4;
  float* s = const_cast<float*>(st);
  for (int t = 0; t < nthreads; ++t) {
    FwdPartial p = fwd_strip_r<R, false, true>(hasK ? s : nullptr, s + HW, s + 2 * HW, s + 3 * HW, H, W, t % W4,
                                               (t / W4) * R, use_tb != 0, 0.f, 0.f, nullptr, nullptr, nullptr,
                                               nullptr, nullptr);
    acc[0] += p.c; acc[1] += p.d; acc[2] += p.dir; acc[3] += p.neu;
  }
  float *P1 = scratch, *P2 = P1 + HW, *P3 = P2 + HW, *Q1 = P3 + HW, *Q2 = Q1 + HW;
  // the kernel keeps Q1 / Q2 in per-thread registers (qreg): emulated with one private array per thread
  std::vector<float> qreg((size_t)nthreads * 2 * R * 4);
  for (int t = 0; t < nthreads; ++t)
    (void)fwd_strip_r<R, true, false>(hasK ? s : nullptr, s + HW, s + 2 * HW, s + 3 * HW, H, W, t % W4,
                                      (t / W4) * R, use_tb != 0, a, b, P1, P2, P3, Q1, Q2,
                                      qreg.data() + (size_t)t * 2 * R * 4);
  for (int t = 0; t < nthreads; ++t)
    bwd_strip_pass2_r<R>(P1, P2, P3, Q1, Q2, s + HW, s + 3 * HW, s + HW, s + 2 * HW, s + 3 * HW, H, W, t % W4,
                         (t / W4) * R, cdir, cneu, qreg.data() + (size_t)t * 2 * R * 4);
}

static int run_case(int B, int H, int W, int use_tb, int hasK, int nthreads, int fitR = 0) {
  using namespace pdes::stencil;
  const int HW = H * W;
  std::vector<float> K((size_t)B * HW), out((size_t)B * 3 * HW);
  for (auto& v : K) v = (float)std::exp(0.5 * frand());
  for (auto& v : out) v = (float)frand();
  const double gw[4] = {0.7, 1.3, 10.0, 4.0};
  double l4_ref[4];
  std::vector<double> dref((size_t)B * 3 * HW);
  pdes_oracle_darcy(hasK ? K.data() : nullptr, out.data(), B, H, W, use_tb, gw, l4_ref, dref.data());

  const double n_c = (double)B * HW, n_d = (double)B * (use_tb ? H : H - 2) * W;
  const double n_dir = (double)B * H, n_neu = (double)B * 2 * W;
  double acc[4] = {0, 0, 0, 0};
  std::vector<float> dout((size_t)B * 3 * HW);
  std::vector<float> scratch((size_t)5 * HW), stage((size_t)4 * HW);
  const float a = hasK ? (float)(gw[0] * 2.0 / n_c) : 0.f, b = (float)(gw[1] * 2.0 / n_d);
  const float cdir = (float)(gw[2] * 2.0 / n_dir), cneu = (float)(gw[3] * 2.0 / n_neu);
  for (int s = 0; s < B; ++s) {
    for (int p = 0; p < HW; ++p) {
      stage[p] = K[(size_t)s * HW + p];
      for (int c = 0; c < 3; ++c) stage[(size_t)(c + 1) * HW + p] = out[((size_t)s * 3 + c) * HW + p];
    }
    float* st = stage.data();
    if (fitR > 0) {
      if (fitR == 2) fast_sample<2>(st, hasK, H, W, use_tb, a, b, cdir, cneu, scratch.data(), acc);
      else fast_sample<4>(st, hasK, H, W, use_tb, a, b, cdir, cneu, scratch.data(), acc);
      for (int p = 0; p < 3 * HW; ++p) dout[(size_t)s * 3 * HW + p] = st[HW + p];
      continue;
    }
    for (int t = 0; t < nthreads; ++t) {
      FwdPartial p = fwd_strip(hasK ? st : nullptr, st + HW, st + 2 * HW, st + 3 * HW, H, W, t, nthreads,
                               true, use_tb != 0, 0.f, 0.f, nullptr, nullptr, nullptr, nullptr, nullptr);
      acc[0] += p.c; acc[1] += p.d; acc[2] += p.dir; acc[3] += p.neu;
    }
    float *P1 = scratch.data(), *P2 = P1 + HW, *P3 = P2 + HW, *Q1 = P3 + HW, *Q2 = Q1 + HW;
    for (int t = 0; t < nthreads; ++t)
      (void)fwd_strip(hasK ? st : nullptr, st + HW, st + 2 * HW, st + 3 * HW, H, W, t, nthreads, true,
                      use_tb != 0, a, b, P1, P2, P3, Q1, Q2);
    for (int t = 0; t < nthreads; ++t)
      bwd_strip_pass2(P1, P2, P3, Q1, Q2, st + HW, st + 3 * HW, st + HW, st + 2 * HW, st + 3 * HW, H, W, t,
                      nthreads, true, cdir, cneu);
    for (int p = 0; p < 3 * HW; ++p) dout[(size_t)s * 3 * HW + p] = st[HW + p];
  }
  double l4[4] = {acc[0] / n_c, acc[1] / n_d, acc[2] / n_dir, acc[3] / n_neu};
  double worst = 0;
  for (int i = 0; i < 4; ++i) {
    const double e = std::fabs(l4[i] - l4_ref[i]) / std::fmax(std::fabs(l4_ref[i]), 1e-30);
    if (l4_ref[i] != 0.0 && e > worst) worst = e;
    if (l4_ref[i] == 0.0 && l4[i] != 0.0) worst = 1.0;
  }
  double num = 0, den = 0;
  for (size_t i = 0; i < dref.size(); ++i) {
    num += (dout[i] - dref[i]) * (dout[i] - dref[i]);
    den += dref[i] * dref[i];
  }
  const double ge = std::sqrt(num / std::fmax(den, 1e-300));
  const int ok = worst < 2e-5 && ge < 2e-5;
  printf("B=%d H=%d W=%d use_tb=%d hasK=%d nthr=%d fitR=%d : loss rel err %.2e, grad rel-L2 %.2e %s\n", B, H, W,
         use_tb, hasK, nthreads, fitR, worst, ge, ok ? "ok" : "FAIL");
  return ok ? 0 : 1;
}

int main() {
  srand(1234);
  int fails = 0;
  const int shapes[][2] = {{16, 16}, {32, 32}, {64, 64}, {8, 12}, {64, 32}, {5, 4}, {3, 8}, {7, 128}};
  for (auto& s : shapes)
    for (int tb = 0; tb < 2; ++tb)
      for (int hk = 0; hk < 2; ++hk) fails += run_case(2, s[0], s[1], tb, hk, 256);
  // exact-fit unrolled strips: the shapes the kernels use (64x64) and small/edge-heavy ones
  const int fshapes[][2] = {{64, 64}, {32, 32}, {4, 4}, {8, 12}, {6, 8}, {12, 16}, {64, 32}};
  for (auto& s : fshapes)
    for (int tb = 0; tb < 2; ++tb)
      for (int hk = 0; hk < 2; ++hk) {
        if (s[0] % 2 == 0) fails += run_case(2, s[0], s[1], tb, hk, 0, 2);
        if (s[0] % 4 == 0) fails += run_case(2, s[0], s[1], tb, hk, 0, 4);
      }
  fails += run_case(1, 64, 64, 1, 1, 128);
  fails += run_case(1, 32, 32, 1, 1, 64);
  printf("%s\n", fails ? "EMUL FAILED" : "EMUL PASSED");
  return fails ? 1 : 0;
}

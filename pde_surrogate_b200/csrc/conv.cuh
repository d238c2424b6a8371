// conv.cuh — argument blocks shared by the convolution kernels (SIMT fp32 and tcgen05) and the
// network executor.
#pragma once
#include "common.cuh"

namespace pdes {

// Where a kernel gets per-channel BatchNorm constants from.  Either explicit scale/shift arrays,
// or (training) the double-precision batch sums produced by the layer that wrote the channels,
// or (eval) the running statistics.  nn.BatchNorm2d semantics, reference models/codec.py:57-66.
struct BnSrc {
  const float* scale;   // explicit (unit tests); overrides everything else when non-null
  const float* shift;
  const double* sum;    // per-channel sum / sum of squares over B*H*W
  const double* sumsq;
  double inv_count;
  const float* gamma;
  const float* beta;
  const float* run_mean;
  const float* run_var;
  int use_running;
  float eps;
};

enum { IN_DIRECT = 0, IN_UPSAMPLE = 1, IN_ZEROINS = 2 };
enum { EPI_NHWC = 0, EPI_NCHW = 1, EPI_BNBWD = 2 };

struct ConvArgs {
  // ---- input operand -----------------------------------------------------------------
  const float* x;  // NHWC (ldx floats per pixel, first channel at x[0]) or NCHW if in_nchw
  int ldx, Cin, Hs, Ws, B;
  int in_mode;     // IN_DIRECT | IN_UPSAMPLE (nearest x2) | IN_ZEROINS (stride-2 transpose)
  int in_nchw;
  int pro;         // 1: a = max(0, x*scale+shift)
  BnSrc bn;
  // ---- weights packed [tap][CinP][CoP] ------------------------------------------------
  const float* w;
  int CinP, CoP, Cout;
  int KS, pad, stride;
  int Ho, Wo;      // conv output size (before the optional 2x2 sum-pool)
  // ---- epilogue -------------------------------------------------------------------------
  int epi;
  int pool;        // 1: 2x2 sum-pool of the conv output (adjoint of nearest upsampling)
  float* y;        // EPI_NHWC / EPI_NCHW destination
  int ldy, coff;
  double* o_sum;   // EPI_NHWC: per-channel sum / sumsq of what was stored (+=), or null
  double* o_sumsq;
  // EPI_BNBWD: acc is dL/d(a) of a consumer layer; fold ReLU mask and the BatchNorm backward
  const float* fx; // the consumer's input activations (NHWC, ldfx), dims Hf x Wf
  int ldfx, Hf, Wf;
  BnSrc fbn;       // the consumer's BatchNorm
  float* G;        // gradient buffer of the same tensor (ldG)
  int ldG, g_accum;
  double* bsum;    // [0,C): sum dZ ; [C,2C): sum dZ*xhat   (C = Cout of this "conv")
  unsigned* gmax;  // EPI_BNBWD: running max |G| written (float bits, atomicMax) or null
};

struct WgradArgs {
  const float* x;
  int ldx, Cin, Hs, Ws, B, in_mode, in_nchw, pro;
  BnSrc bn;
  const float* dy;  // NHWC (lddy, channel offset already applied) or NCHW
  int lddy, dy_nchw, Cout;
  int KS, pad, stride, Ho, Wo;
  float* dw;        // OIHW, accumulated with atomics
};

constexpr int kMaxConsumers = 24;
struct FixDyArgs {
  float* G;         // gradient slice base (channel offset applied), ldG
  const float* X;   // activation slice base, ldX
  int ldG, ldX, C;
  int64_t npix;
  const double* sum;  // stats of these channels (offset applied)
  const double* sumsq;
  double inv_count;
  float eps;
  int n_cons;
  const float* cons_gamma[kMaxConsumers];  // consumer BN weight, offset to this slice
  const double* cons_bsum[kMaxConsumers];  // consumer bsum base, offset to this slice
  int cons_C[kMaxConsumers];               // consumer channel count (stride to the 2nd half)
  // nn.Dropout2d behind the producing convolution (models/codec.py:70-71, 110-149, 171-172): the slice holds
  // y * mask; its gradient is multiplied by the same per-(sample, channel) mask.  null: no dropout
  const float* drop_mask;                  // [B][C], values 0 or 1/(1-p)
  int64_t pix_per_img;
};
// y[:, c] *= mask[b][c] in place on an NHWC slice, and per-channel sum / sum of squares of the result (+=)
int launch_dropout_fwd(float* y, int ld, int C, int64_t npix, int64_t pix_per_img, const float* mask,
                       double* o_sum, double* o_sumsq, cudaStream_t st);

// upsample='bilinear' (bilinear.cu): x = the layer's input (NHWC, H x W, ldx), up = the fp32 NHWC scratch at
// 2H x 2W (ldu): the upsampled activation in the forward pass, the convolution's data gradient in the backward
struct BilinearArgs {
  const float* x;
  int ldx, C, H, W, B;
  int pro;          // forward: apply BatchNorm + ReLU before interpolating
  BnSrc bn;
  float* up;
  int ldu;
  // backward only
  float* G;
  int ldG, g_accum;
  double* bsum;     // [0,C): sum dZ ; [C,2C): sum dZ*xhat
  unsigned* gmax;
  int zero_insert;  // 1: the "upsampling" is the zero insertion of a stride-2 transposed convolution
                    // (up[2y][2x] = a[y][x], zero elsewhere) instead of the bilinear interpolation
};
int launch_bilinear_up(const BilinearArgs& a, cudaStream_t st);
int launch_bilinear_bwd(const BilinearArgs& a, cudaStream_t st);
// nn.ConvTranspose2d(k, stride 2) <-> the equivalent Conv2d over the zero-inserted input (bilinear.cu):
//   fwd:  wc[co][ci][ky][kx] = wt[ci][co][K-1-ky][K-1-kx]
//   bwd:  gt[ci][co][K-1-ky][K-1-kx] += gc[co][ci][ky][kx];  gc = 0   (gc is a self-cleaning staging buffer)
int launch_convt_weight(const float* wt, float* wc, int Cin, int Cout, int KS, cudaStream_t st);
int launch_convt_weight_grad(float* gc, float* gt, int Cin, int Cout, int KS, cudaStream_t st);

// cGlow coupling network helpers (coupling.cu)
int launch_nchw_to_block(const float* x, float* act, int ld, int C, int B, int HW, double* o_sum, double* o_sumsq,
                         cudaStream_t st);
int launch_block_to_nchw(const float* g, int ld, int C, int B, int HW, float* dx, cudaStream_t st);
int launch_zeros_fwd(float* out, const float* bias, const float* scale, int B, int C, int HW, float* keep,
                     cudaStream_t st);
int launch_zeros_bwd(const float* dout, const float* out, const float* scale, int B, int C, int HW, float* dyg,
                     float* dbias, float* dscale, cudaStream_t st);

int launch_conv_simt(const ConvArgs& a, cudaStream_t st);
int launch_wgrad_simt(const WgradArgs& a, cudaStream_t st);
int launch_fix_dy(const FixDyArgs& a, cudaStream_t st);

// weight packing for one layer: OIHW -> fwd [tap][CinP][CoP] and bwd [tap'][CoutP][CiP]
struct PackDesc {
  const float* w;   // OIHW
  float* wf;        // [KS*KS][CinP][CoP]
  float* wb;        // [KS*KS][CoutP_b][CiP_b], taps flipped; may be null
  int Cout, Cin, KS, CinP, CoP, CoutPb, CiPb;
};
int launch_pack_weights(const PackDesc* dev_table, int n_layers, int max_elems, cudaStream_t st);

struct BnLayerDesc {
  const double* sum;   // stats of the BN input channels
  const double* sumsq;
  const double* bsum;  // backward sums of this BN (2*C)
  float* run_mean;
  float* run_var;
  float* dgamma;
  float* dbeta;
  double count;        // pixels per sample (H*W); the batch size is a launch argument
  int C;
};
int launch_bn_running_update(const BnLayerDesc* dev_table, int n, int maxC, float momentum,
                             int B, cudaStream_t st);
int launch_bn_param_grad(const BnLayerDesc* dev_table, int n, int maxC, cudaStream_t st);

}  // namespace pdes

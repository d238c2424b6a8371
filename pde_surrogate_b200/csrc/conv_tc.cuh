// conv_tc.cuh — interface of the tcgen05 convolution kernels.
#pragma once
#include <cuda_fp16.h>
#include "conv.cuh"

namespace pdes {

// ---- fp32-accurate tensor-core operands -----------------------------------------------------
// Every GEMM operand x is stored as TWO fp16 pieces of the power-of-two scaled value:
//   x * 2^s = h1 + h2,   h1 = fp16(x * 2^s),  h2 = fp16(x * 2^s - h1)      (11 + 11 significand bits)
// and a product a*w is evaluated as the three tensor-core products a1*w1 + a1*w2 + a2*w1 (the dropped
// a2*w2 is 2^-22 relative), accumulated in fp32 in separate TMEM column groups for the leading and
// the cross terms.  fp16 has 5 exponent bits, hence the scaling: static powers of two for
// activations and filters, a per-layer dynamic power of two for gradients (derived on the device
// from the running maximum of the gradient buffer the slice lives in); the epilogues multiply by
// the exact inverse.  Below 2^-2 (scaled) the second piece is subnormal: the absolute error of an
// element is max(2^-23 |x|, 2^-25 / 2^s).
using op16 = __half;   // 16-bit storage of a piece (fp16, or bf16 bits in the bf16 one-piece mode)
constexpr int kPieces = 2;
// Reduced-precision modes (BASELINE config 3, "bf16 tensor-core conv path"): ONE piece per operand and one
// tensor-core product per useful product instead of three.  The memory layouts stay those of the two-piece
// mode (the second piece is simply never written, loaded or multiplied).
enum { LOWP_NONE = 0, LOWP_FP16 = 1, LOWP_BF16 = 2 };
// instruction-descriptor bits selecting bf16 A/B operands (kind::f16: format 0 = fp16, 1 = bf16)
__host__ __device__ inline uint32_t idesc_fmt_bits(int lowp) { return lowp == LOWP_BF16 ? ((1u << 7) | (1u << 10)) : 0u; }
constexpr int kActScaleLog2 = 4;    // activations (post BatchNorm+ReLU, O(1)):   x * 16
constexpr int kWScaleLog2 = 8;      // filters (|w| << 256):                      w * 256
constexpr int kDyTargetLog2 = 10;   // gradients: buffer maximum scaled into [2^10, 2^11)
inline float pow2f(int e) {
  union { uint32_t u; float f; } v;
  v.u = (uint32_t)(e + 127) << 23;
  return v.f;
}

// elementwise pre-pass of the tensor-core wgrad: two fp16 piece planes of the operand
struct ActSplitArgs {
  const float* x;        // NHWC, ldx floats per pixel (channel offset already applied)
  int ldx, C, Hs, Ws, B;
  int up;                // 1: write the nearest x2 upsampled image; 2: zero-insert x2 (value at even positions)
  int nchw;              // 1: x is planar (B,C,Hs,Ws)
  int pro;               // 1: a = max(0, x*scale+shift) first
  BnSrc bn;
  op16* out;    // [2][B][Hv][Cp/8][Wv][8]
  int Cp;                // C rounded up to 8
  // fused dY correction (replaces a separate fix_dy launch): x is the gradient slice G, rewritten in
  // place as G - c1[c] - xhat*c2[c] before the split; fx.G is ignored
  int fix;
  FixDyArgs fx;
  // scaling of the pieces: x * scale, with scale = 2^kActScaleLog2 style constants, or — when dyn_max
  // is set — the power of two that brings the float whose bits are *dyn_max (running |gradient|
  // maximum of the buffer) to 2^kDyTargetLog2; the inverse is published in *dyn_inv for the
  // consumers' epilogues
  float scale;
  const unsigned* dyn_max;
  float* dyn_inv;
  int lowp;              // LOWP_*: one piece (fp16 / bf16) instead of two fp16 pieces
};

struct TcWgradArgs {
  const op16* planesA;  // [2][B][Hv][CpA/8][Wv][8]  activation pieces (after BN+ReLU / upsampling)
  const op16* planesB;  // [2][B][Ho][CpB/8][Wo][8]  dY pieces
  float* dwp;                    // staging gradient [tap][ci_pad][co_pad] (vector reductions)
  int B, Hv, Wv, Ho, Wo, Cin, Cout, KS, pad;
  int ci_pad, co_pad;
  float out_scale;               // exact inverse of the static operand scales
  const float* dyn_scale;        // optional device scalar multiplied in as well (dynamic dY scale)
  int lowp;                      // LOWP_*: a1 x d1 only (one pass)
};

struct TcWgradUnpack {
  float* dw;         // OIHW gradient (+=)
  float* dwp;        // staging buffer (read, then cleared)
  int Cout, Cin, KS, ci_pad, co_pad;
  int taps_in_n;     // > 0: staged as a 1x1 problem with N = (tap, co): dwp[ci][tap * Cout + co], taps_in_n = KS*KS
};

// Weight gradient of a convolution with very few output channels (the network's last layer): the
// GEMM N would be Cout padded to 16 per filter tap.  Instead the dY operand is expanded once into
// KS*KS shifted copies, n = tap * Cout + co:  planes[.][b][y][n/8][x][n%8] = dY[b][co][y-ky+pad][x-kx+pad],
// and the weight gradient becomes ONE 1x1-convolution GEMM with N = KS*KS*Cout.
struct DyIm2colArgs {
  const float* dy;   // planar (B, Cout, H, W) output gradient
  op16* out;         // [2][B][H][Np/8][W][8], Np = KS*KS*Cout rounded up to 8
  int B, H, W, Cout, KS, pad, Np;
  const unsigned* dyn_max;  // dynamic scale as in ActSplitArgs
  float* dyn_inv;
  int lowp;
};
int launch_dy_im2col(const DyIm2colArgs& a, cudaStream_t st);

bool wgrad_tc_supported(int KS, int stride);
void wgrad_tc_dims(int Cin, int Cout, int* ci_pad, int* co_pad);
int launch_wgrad_tc(const TcWgradArgs& t, cudaStream_t st);
int launch_act_split(const ActSplitArgs& a, cudaStream_t st);
size_t act_planes_bytes(int B, int H, int W, int C);
// max |x| over n floats as float bits, atomicMax-ed into *out (the caller clears it)
int launch_absmax(const float* x, size_t n, unsigned* out, cudaStream_t st);
// max_cin: largest Cin in the table; max_slab_floats: largest Cout * 8 * KS*KS in the table
int launch_wgrad_unpack(const TcWgradUnpack* dev_table, int n, int max_cin, int max_slab_floats, cudaStream_t st);

// ---- TMA-fed two-piece fp16 convolution (conv_tc2.cu) -----------------------------------------------
struct Tc2Plan {
  int KC, nchunks, ngroups, S, TS, AST, NB, TPB;
  size_t smem, pack_elems;  // pack_elems: 16-bit elements of the packed filter
};
struct Tc2Args {
  ConvArgs c;                  // geometry (B, Ho, Wo, KS, pad, Cout) and epilogue; x/w/prologue unused
  const op16* wpk;    // [chunk][tap][k-octet][piece][N][8]
  int N, KC, nchunks, ngroups, S, TS, AST, NB, TPB;
  long long* dbg;              // optional per-CTA phase timestamps (debug), 16 slots per CTA
  int osub;                    // 1: store only even output positions at (y/2, x/2)  (stride-2 as stride-1)
  float out_scale;             // exact inverse of the static operand scales, applied to the accumulator
  const float* dyn_scale;      // optional device scalar multiplied in as well (dynamic dY scale)
  int exp;                     // timing experiments (PDES_TC2_EXP): 1 = skip activation loads, 2 = skip filter loads
  int lowp;                    // LOWP_*: single product a1 x w1 (one accumulator group)
};
struct Tc2PackDesc {
  const float* w;  // OIHW
  op16* dst;
  int Cout, Cin, KS, N, KC, nchunks, transpose;
  int lowp;        // LOWP_*: only the first piece is written, in fp16 or bf16
  int dxn, CoP;    // dxn = 1: "dx in N" layout of conv_dense.cu, [chunk][ky][k-octet][piece][n = kx*CoP + co][8]
                   // dxn = 2: "dx in K" layout of conv_dense_bwd.cu, [jy][k-octet (jx, co octet)][piece][n = ci][8]
                   // dxn = 3: data gradient over the (tap, co)-expanded dY planes: conv_tc2 1x1 layout, k = tap*Cout + co
};
void tc2_plan(int KS, int Cin_k, int N, Tc2Plan* p, int lowp = 0);
bool tc2_supported(int KS, int stride, int Cin_k, int N);
// planes: [2][B][Hv][round8(Cin_k)/8][Wv][8] fp16 pieces of the GEMM-K operand (act_split_kernel)
int launch_conv_tc2(const Tc2Args& t, const op16* planes, int Hv, int Wv, int Cin_k, cudaStream_t st);
int launch_pack_tc2(const Tc2PackDesc* dev_table, int n, size_t max_elems, cudaStream_t st);

// ---- fused thin-layer forward (conv_dense.cu): BatchNorm + ReLU + fp16 split inside the convolution ----
struct DenseFwdArgs {
  const float* x;      // NHWC fp32 input (the dense block's buffer), ldx floats per pixel
  int ldx, Cin, H, W, B;
  int pro;             // 1: a = max(0, x*scale+shift) first
  BnSrc bn;
  const op16* wpk;     // packed filter, Tc2PackDesc::dxn layout
  int CoP, Cout;       // CoP = 16
  float* y;            // NHWC output slice: y[pixel*ldy + coff + co]
  int ldy, coff;
  double* o_sum;       // per-channel sum / sum of squares of what was stored (+=), offset to the slice; or null
  double* o_sumsq;
  op16* planes;        // optional [2][B][H][Cp/8][W][8]: the operand pieces, written once for the weight gradient
  int Cp;
  float out_scale;     // exact inverse of the static operand scales
  int b_early;         // 1: the filter may be fetched before griddepcontrol.wait (packed >= 2 launches ago)
  int lowp;            // LOWP_*
  long long* dbg;      // optional phase timestamps (clock64), 64 slots per CTA for the first 4 CTAs (diagnostics)
  int exp;             // timing experiments (PDES_DENSE_EXP; results are wrong): 1 converters skip their shared-memory
                       // traffic, 2 no a2 x w1 MMAs, 4 no raw TMA loads
};
// ---- fused thin-layer data gradient (conv_dense_bwd.cu) ----
struct DenseBwdArgs {
  FixDyArgs fx;          // gradient slice G / activation slice X of the layer's OUTPUT + the consumers' lazy
                         // BatchNorm-backward corrections (n_cons == 0: dY = G as it is)
  int H, W, B, Cout;
  const unsigned* dyn_max;   // running |G| maximum of the buffer (float bits) -> dynamic power-of-two scale
  float* dyn_inv;            // its exact inverse, published for the weight-gradient epilogue
  op16* planesB;             // optional: dY pieces [2][B][H][round8(Cout)/8][W][8] for the weight gradient
  const op16* wpk;           // packed filter, Tc2PackDesc::dxn == 2 layout: [jy][k-octet (jx, co octet)][piece][n = ci][8]
  int N, Cin;                // N = Cin rounded up to 16
  const float* x;            // the layer's INPUT activations (block buffer), ldx floats per pixel
  int ldx;
  BnSrc fbn;                 // the layer's BatchNorm
  float* G;                  // gradient buffer of the same tensor
  int ldG, g_accum;
  double* bsum;              // [0,Cin): sum dZ ; [Cin,2Cin): sum dZ*xhat
  unsigned* gmax;
  float out_scale;
  int b_early;               // 1: the filter may be fetched before griddepcontrol.wait (packed >= 2 launches ago)
  int lowp;                  // LOWP_*
  long long* dbg;            // optional phase timestamps (clock64), 64 slots per CTA for the first 4 CTAs
};
bool dense_bwd_supported(int KS, int stride, int pad, int up, int Cin, int Cout, int H, int W);
size_t dense_bwd_pack_elems(int N);
int launch_conv_dense_bwd(const DenseBwdArgs& a, cudaStream_t st);
bool dense_fwd_supported(int KS, int stride, int pad, int up, int Cin, int Cout, int H, int W);
size_t dense_pack_elems(int Cin, int CoP);   // 16-bit elements of the packed filter (both pieces)
int launch_conv_dense_fwd(const DenseFwdArgs& a, cudaStream_t st);

}  // namespace pdes

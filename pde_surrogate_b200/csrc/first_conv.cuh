// first_conv.cuh — interface of the dedicated first-convolution kernels (first_conv.cu).
#pragma once
#include "common.cuh"

namespace pdes {

struct FirstConvArgs {
  const float* x;   // planar network input (B, Cin, H, W)
  const float* w;   // OIHW filter (forward)
  float* y;         // NHWC output, ldy floats per pixel, first channel at y[coff] (forward)
  int ldy, coff, vec_ok;
  double* o_sum;    // per-channel sum / sum of squares of y (+=), or null
  double* o_sumsq;
  const float* dy;  // NHWC output gradient, lddy floats per pixel, channel offset applied (wgrad)
  int lddy;
  float* dw;        // OIHW filter gradient, accumulated with atomics (wgrad)
  int B, Cin, H, W, Cout, KS, stride, pad, Ho, Wo;
};

bool first_conv_supported(int Cin, int Cout, int KS, int stride);
int launch_first_conv_fwd(const FirstConvArgs& a, cudaStream_t st);
int launch_first_conv_wgrad(const FirstConvArgs& a, cudaStream_t st);

}  // namespace pdes

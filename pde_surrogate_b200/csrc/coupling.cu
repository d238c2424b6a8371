// coupling.cu — the pieces around the dense layers that make a cGlow coupling network
// (_DenseCoupling, models/glow_msc.py:276-294) out of the DenseED executor:
//   * the network input (planar NCHW) becomes the first channels of the dense block's NHWC buffer, with the
//     batch statistics the following BatchNorm layers need (there is no In_conv in front of the block);
//   * Conv2dZeros (models/glow_msc.py:240-255): y = (conv3x3(a) + bias) * exp(3 * scale); the convolution itself
//     is the executor's last-layer path (planar output), bias / gain are applied in place afterwards;
//   * backward of that gain (d conv = dout * gain, d bias, d scale) and the gradient w.r.t. the network input
//     (the coupling network sits inside a flow: its input depends on earlier parameters).
#include "conv.cuh"

namespace pdes {
namespace {

__global__ void __launch_bounds__(256) nchw_to_block_kernel(const float* __restrict__ x, float* act, int ld, int C,
                                                            int B, int HW, double* o_sum, double* o_sumsq) {
  griddep_wait();
  __shared__ float s1[256], s2[256];
  for (int c = threadIdx.x; c < C; c += blockDim.x) s1[c] = s2[c] = 0.f;
  __syncthreads();
  const int64_t total = (int64_t)B * C * HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const int64_t r = i / HW;
    const int c = (int)(r % C), b = (int)(r / C);
    const float v = x[i];
    act[((size_t)b * HW + p) * ld + c] = v;
    if (o_sum != nullptr) {
      atomicAdd(&s1[c], v);
      atomicAdd(&s2[c], v * v);
    }
  }
  __syncthreads();
  if (o_sum != nullptr)
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      atomicAdd(o_sum + c, (double)s1[c]);
      atomicAdd(o_sumsq + c, (double)s2[c]);
    }
}

__global__ void __launch_bounds__(256) block_to_nchw_kernel(const float* __restrict__ g, int ld, int C, int B, int HW,
                                                            float* dx) {
  griddep_wait();
  const int64_t total = (int64_t)B * C * HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const int64_t r = i / HW;
    const int c = (int)(r % C), b = (int)(r / C);
    dx[i] = g[((size_t)b * HW + p) * ld + c];
  }
}

// out = (out + bias[c]) * exp(3 scale[c]) in place (planar); keep = a copy for the backward pass (or null)
__global__ void __launch_bounds__(256) zeros_fwd_kernel(float* out, const float* __restrict__ bias,
                                                        const float* __restrict__ scale, int B, int C, int HW,
                                                        float* keep) {
  griddep_wait();
  const int64_t total = (int64_t)B * C * HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)((i / HW) % C);
    const float v = (out[i] + bias[c]) * expf(3.f * scale[c]);
    out[i] = v;
    if (keep != nullptr) keep[i] = v;
  }
}

// dyg = dout * gain;  dbias[c] += sum dyg;  dscale[c] += 3 * sum dout * out      (one block per channel)
__global__ void __launch_bounds__(256) zeros_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                        const float* __restrict__ scale, int B, int C, int HW,
                                                        float* dyg, float* dbias, float* dscale) {
  griddep_wait();
  const int c = blockIdx.x;
  const float gain = expf(3.f * scale[c]);
  double sb = 0.0, ss = 0.0;
  for (int b = 0; b < B; ++b) {
    const size_t base = ((size_t)b * C + c) * HW;
    for (int p = threadIdx.x; p < HW; p += blockDim.x) {
      const float d = dout[base + p];
      const float v = d * gain;
      dyg[base + p] = v;
      sb += (double)v;
      ss += (double)d * (double)out[base + p];
    }
  }
  __shared__ double r1[8], r2[8];
  sb = warp_sum(sb);
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) {
    r1[threadIdx.x >> 5] = sb;
    r2[threadIdx.x >> 5] = ss;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b2 = 0.0;
    for (int w = 0; w < 8; ++w) {
      a += r1[w];
      b2 += r2[w];
    }
    dbias[c] += (float)a;
    dscale[c] += (float)(3.0 * b2);
  }
}

int grid_of(int64_t total) {
  int blocks = (int)((total + 255) / 256);
  const int cap = sm_count() * 4;
  if (blocks > cap) blocks = cap;
  return blocks < 1 ? 1 : blocks;
}

}  // namespace

int launch_nchw_to_block(const float* x, float* act, int ld, int C, int B, int HW, double* o_sum, double* o_sumsq,
                         cudaStream_t st) {
  PDES_REQUIRE(x && act && C >= 1 && C <= 256, PDES_ERR_INVALID, "nchw_to_block: invalid arguments (C %d)", C);
  PDES_CUDA(launch_pdl(nchw_to_block_kernel, dim3(grid_of((int64_t)B * C * HW)), dim3(256), 0, st, x, act, ld, C, B, HW,
                       o_sum, o_sumsq));
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}
int launch_block_to_nchw(const float* g, int ld, int C, int B, int HW, float* dx, cudaStream_t st) {
  PDES_CUDA(launch_pdl(block_to_nchw_kernel, dim3(grid_of((int64_t)B * C * HW)), dim3(256), 0, st, g, ld, C, B, HW, dx));
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}
int launch_zeros_fwd(float* out, const float* bias, const float* scale, int B, int C, int HW, float* keep,
                     cudaStream_t st) {
  PDES_CUDA(launch_pdl(zeros_fwd_kernel, dim3(grid_of((int64_t)B * C * HW)), dim3(256), 0, st, out, bias, scale, B, C, HW,
                       keep));
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}
int launch_zeros_bwd(const float* dout, const float* out, const float* scale, int B, int C, int HW, float* dyg,
                     float* dbias, float* dscale, cudaStream_t st) {
  PDES_CUDA(launch_pdl(zeros_bwd_kernel, dim3(C), dim3(256), 0, st, dout, out, scale, B, C, HW, dyg, dbias, dscale));
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

}  // namespace pdes

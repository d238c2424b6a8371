// first_conv.cu — dedicated CUDA-core kernels for the network's first convolution
// (In_conv: Conv2d(in_channels -> init_features, k7, s2, p3|2, bias=False), models/codec.py:238-243).
//
// With one input channel the GEMM-K of this layer is 49: on the tensor-core path it had to be run
// as a stride-1 convolution with K padded 1 -> 16 per tap and three quarters of the outputs thrown
// away (~100 us), and its weight gradient on the generic SIMT kernel cost another ~130 us — for
// 0.3 % of the network's FLOPs.  Here the planar input tile and the whole filter live in shared
// memory, each thread owns one output pixel (forward) or a (4 output channels x filter row) strip
// of the weight gradient, and the arithmetic is plain fp32 FMA.
//   forward : y[b, oy, ox, coff + co] = sum_{ci,ky,kx} x[b, ci, S*oy+ky-pad, S*ox+kx-pad] * w[co,ci,ky,kx]
//             + per-channel sum / sum-of-squares of y for the consumers' BatchNorm (fp64 atomics)
//   wgrad   : dw[co,ci,ky,kx] += sum_{b,oy,ox} dy[b, oy, ox, co] * x[b, ci, S*oy+ky-pad, S*ox+kx-pad]
#include "conv.cuh"
#include "first_conv.cuh"

namespace pdes {
namespace {

constexpr int kKS = 7, kS = 2;
constexpr int kTOH = 8, kTOW = 16;                 // output pixels per tile
constexpr int kIH = (kTOH - 1) * kS + kKS;         // 21 input rows
constexpr int kIW = (kTOW - 1) * kS + kKS;         // 37 input columns
constexpr int kFwdThreads = kTOH * kTOW;           // 128: one thread per output pixel
constexpr int kWgThreads = 256;

__device__ __forceinline__ void stage_input(const FirstConvArgs& a, float* xs, int b, int oy0, int ox0) {
  const int iy0 = oy0 * kS - a.pad, ix0 = ox0 * kS - a.pad;
  for (int i = threadIdx.x; i < a.Cin * kIH * kIW; i += blockDim.x) {
    const int ci = i / (kIH * kIW), r = i - ci * (kIH * kIW);
    const int iy = iy0 + r / kIW, ix = ix0 + r % kIW;
    float v = 0.f;
    if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) v = __ldg(a.x + (((size_t)b * a.Cin + ci) * a.H + iy) * a.W + ix);
    xs[i] = v;
  }
}

__global__ void __launch_bounds__(kFwdThreads) first_conv_fwd_kernel(FirstConvArgs a) {
  griddep_wait();
  extern __shared__ __align__(16) float sm[];
  const int CoP = (a.Cout + 15) & ~15;
  float* ws = sm;                                      // [Cin*49][CoP]
  float* xs = ws + (size_t)a.Cin * kKS * kKS * CoP;    // [Cin][kIH][kIW]
  float* red = xs + ((a.Cin * kIH * kIW + 3) & ~3);    // [4 warps][CoP][2]
  const int tiles_x = (a.Wo + kTOW - 1) / kTOW;
  const int b = blockIdx.y;
  const int oy0 = (blockIdx.x / tiles_x) * kTOH, ox0 = (blockIdx.x % tiles_x) * kTOW;
  const int T = kKS * kKS;
  for (int i = threadIdx.x; i < a.Cin * T * CoP; i += blockDim.x) {
    const int co = i % CoP, r = i / CoP;  // r = ci*49 + tap
    ws[i] = co < a.Cout ? __ldg(a.w + ((size_t)co * a.Cin) * T + r) : 0.f;
  }
  stage_input(a, xs, b, oy0, ox0);
  __syncthreads();
  const int ty = threadIdx.x / kTOW, tx = threadIdx.x % kTOW;
  const int oy = oy0 + ty, ox = ox0 + tx;
  const bool valid = oy < a.Ho && ox < a.Wo;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool want_red = a.o_sum != nullptr;
  const int my_col = colsum16_col(lane);
  float* dst = a.y + (((size_t)b * a.Ho + oy) * a.Wo + ox) * a.ldy + a.coff;
  for (int g = 0; g < CoP; g += 16) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    for (int ci = 0; ci < a.Cin; ++ci) {
      const float* xr = xs + (size_t)ci * kIH * kIW + (ty * kS) * kIW + tx * kS;
      const float* wr = ws + (size_t)ci * T * CoP + g;
#pragma unroll
      for (int ky = 0; ky < kKS; ++ky) {
#pragma unroll
        for (int kx = 0; kx < kKS; ++kx) {
          const float xv = xr[ky * kIW + kx];
          const float4* w4 = reinterpret_cast<const float4*>(wr + (ky * kKS + kx) * CoP);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 w = w4[q];
            acc[4 * q + 0] = fmaf(xv, w.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(xv, w.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(xv, w.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(xv, w.w, acc[4 * q + 3]);
          }
        }
      }
    }
    if (valid) {
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        if (a.vec_ok && g + i + 3 < a.Cout) {
          *reinterpret_cast<float4*>(dst + g + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (g + i + k < a.Cout) dst[g + i + k] = acc[i + k];
        }
      }
    }
    if (want_red) {
      float s1[16], s2[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        s1[i] = valid ? acc[i] : 0.f;
        s2[i] = s1[i] * s1[i];
      }
      const float u = colsum16(s1, lane), w = colsum16(s2, lane);
      if ((lane & 1) == 0) {
        red[((size_t)warp * CoP + g + my_col) * 2 + 0] = u;
        red[((size_t)warp * CoP + g + my_col) * 2 + 1] = w;
      }
    }
  }
  if (want_red) {
    __syncthreads();
    for (int c = threadIdx.x; c < a.Cout; c += blockDim.x) {
      double u = 0.0, w = 0.0;
#pragma unroll
      for (int q = 0; q < kFwdThreads / 32; ++q) {
        u += (double)red[((size_t)q * CoP + c) * 2 + 0];
        w += (double)red[((size_t)q * CoP + c) * 2 + 1];
      }
      atomicAdd(a.o_sum + c, u);
      atomicAdd(a.o_sumsq + c, w);
    }
  }
}

// Weight gradient.  Work item = (group of 4 output channels, filter row ky): 7 x 4 accumulators per
// input channel in registers.  blockDim / n_items pixel subsets share a tile; a CTA walks several
// tiles before it reduces the subsets through shared memory and issues one atomic per weight.
__global__ void __launch_bounds__(kWgThreads) first_conv_wgrad_kernel(FirstConvArgs a, int n_tiles) {
  griddep_wait();
  extern __shared__ __align__(16) float sm[];
  const int Co4 = (a.Cout + 3) >> 2, CoP = Co4 * 4;
  const int n_items = Co4 * kKS;
  const int nsub = kWgThreads / n_items;
  float* dys = sm;                                        // [128 px][CoP]
  float* xs = dys + (size_t)kTOH * kTOW * CoP;            // [Cin][kIH][kIW]
  float* red = xs + ((a.Cin * kIH * kIW + 3) & ~3);       // [(nsub-1)][n_items][28]
  const int sub = threadIdx.x / n_items, item = threadIdx.x - sub * n_items;
  const bool active = sub < nsub;
  const int g = item / kKS, ky = item - g * kKS;
  const int tiles_x = (a.Wo + kTOW - 1) / kTOW, tiles_y = (a.Ho + kTOH - 1) / kTOH;
  const int T = kKS * kKS;
  for (int ci = 0; ci < a.Cin; ++ci) {
    float acc[kKS][4];
#pragma unroll
    for (int i = 0; i < kKS; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][k] = 0.f;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int b = tile / (tiles_x * tiles_y), r = tile - b * (tiles_x * tiles_y);
      const int oy0 = (r / tiles_x) * kTOH, ox0 = (r % tiles_x) * kTOW;
      __syncthreads();  // previous tile fully consumed
      stage_input(a, xs, b, oy0, ox0);  // (all channels; only channel ci is read in this pass)
      for (int i = threadIdx.x; i < kTOH * kTOW * CoP; i += blockDim.x) {
        const int p = i / CoP, c = i - p * CoP;
        const int oy = oy0 + p / kTOW, ox = ox0 + p % kTOW;
        float v = 0.f;
        if (oy < a.Ho && ox < a.Wo && c < a.Cout) v = a.dy[(((size_t)b * a.Ho + oy) * a.Wo + ox) * a.lddy + c];
        dys[i] = v;
      }
      __syncthreads();
      if (active) {
        const float* xc = xs + (size_t)ci * kIH * kIW;
        for (int p = sub; p < kTOH * kTOW; p += nsub) {
          const int ty = p / kTOW, tx = p - ty * kTOW;
          const float4 d = *reinterpret_cast<const float4*>(dys + (size_t)p * CoP + 4 * g);
          const float* xr = xc + (ty * kS + ky) * kIW + tx * kS;
#pragma unroll
          for (int kx = 0; kx < kKS; ++kx) {
            const float xv = xr[kx];
            acc[kx][0] = fmaf(xv, d.x, acc[kx][0]);
            acc[kx][1] = fmaf(xv, d.y, acc[kx][1]);
            acc[kx][2] = fmaf(xv, d.z, acc[kx][2]);
            acc[kx][3] = fmaf(xv, d.w, acc[kx][3]);
          }
        }
      }
    }
    // reduce the pixel subsets, then one atomic per weight and CTA
    __syncthreads();
    if (active && sub > 0) {
      float* r = red + ((size_t)(sub - 1) * n_items + item) * (kKS * 4);
#pragma unroll
      for (int kx = 0; kx < kKS; ++kx)
#pragma unroll
        for (int k = 0; k < 4; ++k) r[kx * 4 + k] = acc[kx][k];
    }
    __syncthreads();
    if (active && sub == 0) {
      for (int s = 1; s < nsub; ++s) {
        const float* r = red + ((size_t)(s - 1) * n_items + item) * (kKS * 4);
#pragma unroll
        for (int kx = 0; kx < kKS; ++kx)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[kx][k] += r[kx * 4 + k];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int co = 4 * g + k;
        if (co < a.Cout) {
          float* dw = a.dw + ((size_t)co * a.Cin + ci) * T + ky * kKS;
#pragma unroll
          for (int kx = 0; kx < kKS; ++kx) atomicAdd(dw + kx, acc[kx][k]);
        }
      }
    }
  }
}

size_t fwd_smem(int Cin, int Cout) {
  const int CoP = (Cout + 15) & ~15;
  return sizeof(float) * ((size_t)Cin * kKS * kKS * CoP + ((Cin * kIH * kIW + 3) & ~3) + (size_t)(kFwdThreads / 32) * CoP * 2);
}
size_t wg_smem(int Cin, int Cout) {
  const int Co4 = (Cout + 3) >> 2, CoP = Co4 * 4;
  const int n_items = Co4 * kKS;
  const int nsub = kWgThreads / n_items;
  return sizeof(float) * ((size_t)kTOH * kTOW * CoP + ((Cin * kIH * kIW + 3) & ~3) +
                          (size_t)(nsub > 1 ? nsub - 1 : 0) * n_items * kKS * 4);
}

}  // namespace

bool first_conv_supported(int Cin, int Cout, int KS, int stride) {
  if (KS != kKS || stride != kS || Cin < 1 || Cin > 4 || Cout < 1) return false;
  if (((Cout + 3) >> 2) * kKS > kWgThreads) return false;  // one wgrad work item per thread
  return fwd_smem(Cin, Cout) <= 200 * 1024 && wg_smem(Cin, Cout) <= 200 * 1024;
}

int launch_first_conv_fwd(const FirstConvArgs& a0, cudaStream_t st) {
  FirstConvArgs a = a0;
  PDES_REQUIRE(first_conv_supported(a.Cin, a.Cout, a.KS, a.stride), PDES_ERR_UNSUPPORTED,
               "first_conv: unsupported shape (Cin %d, Cout %d, k%d s%d)", a.Cin, a.Cout, a.KS, a.stride);
  a.vec_ok = (((a.ldy | a.coff) & 3) == 0 && ((uintptr_t)a.y & 15u) == 0) ? 1 : 0;
  const size_t smem = fwd_smem(a.Cin, a.Cout);
  PDES_ENSURE_SMEM(first_conv_fwd_kernel, smem);
  const dim3 grid(((a.Wo + kTOW - 1) / kTOW) * ((a.Ho + kTOH - 1) / kTOH), a.B);
  PDES_CUDA(launch_pdl(first_conv_fwd_kernel, grid, dim3(kFwdThreads), smem, st, a));
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

int launch_first_conv_wgrad(const FirstConvArgs& a, cudaStream_t st) {
  PDES_REQUIRE(first_conv_supported(a.Cin, a.Cout, a.KS, a.stride), PDES_ERR_UNSUPPORTED,
               "first_conv wgrad: unsupported shape (Cin %d, Cout %d, k%d s%d)", a.Cin, a.Cout, a.KS, a.stride);
  PDES_REQUIRE(a.dy && a.dw, PDES_ERR_INVALID, "first_conv wgrad: null pointer");
  const size_t smem = wg_smem(a.Cin, a.Cout);
  PDES_ENSURE_SMEM(first_conv_wgrad_kernel, smem);
  const int n_tiles = ((a.Wo + kTOW - 1) / kTOW) * ((a.Ho + kTOH - 1) / kTOH) * a.B;
  int grid = sm_count();
  if (grid > n_tiles) grid = n_tiles;
  PDES_CUDA(launch_pdl(first_conv_wgrad_kernel, dim3(grid), dim3(kWgThreads), smem, st, a, n_tiles));
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

}  // namespace pdes

// bilinear.cu — the `upsample='bilinear'` option of DenseED / Decoder (models/codec.py:33-40, 143-146, 177-178;
// train_codec_mixed_residual.py --upsample bilinear): F.interpolate(scale_factor=2, mode='bilinear',
// align_corners=True) between BatchNorm+ReLU and the following 3x3 convolution.
//
// Forward: a_up = bilinear(relu(bn(x))) is materialised once as fp32 NHWC (the convolution kernels then
// see a plain direct input).  Backward: the convolution's data gradient dA_up is gathered back through the
// transposed interpolation, followed by the same ReLU-mask / BatchNorm-backward bookkeeping the dgrad
// epilogues do (sum dZ, sum dZ*xhat, scale*dZ accumulated into the block's gradient buffer, running |G| max).
// Index arithmetic follows ATen's upsample_bilinear2d (float scale = (in-1)/(out-1), src = scale*dst).
//
// The same two kernels serve `upsample=None` (models/codec.py:139-142, nn.ConvTranspose2d(k3, s2, p1, op1)):
// with zero_insert the "interpolation" is the zero insertion of the transposed convolution, which then is a
// plain 3x3 / pad 1 convolution with the flipped, channel-transposed filter (convt_weight_kernel).
#include "conv.cuh"
#include "tc_common.cuh"

namespace pdes {
namespace {

struct Lerp {
  int i0, i1;
  float w0, w1;
};
__device__ __forceinline__ Lerp lerp_of(int dst, float scale, int in_size) {
  const float src = scale * (float)dst;
  Lerp l;
  l.i0 = (int)src;
  if (l.i0 > in_size - 1) l.i0 = in_size - 1;
  l.i1 = l.i0 + ((l.i0 < in_size - 1) ? 1 : 0);
  l.w1 = src - (float)l.i0;
  l.w0 = 1.f - l.w1;
  return l;
}

__global__ void __launch_bounds__(256) bilinear_up_kernel(BilinearArgs a) {
  griddep_wait();
  extern __shared__ float sm_b[];   // scale[C], shift[C]
  float* sc = sm_b;
  float* sh = sm_b + a.C;
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    float s = 1.f, h = 0.f, m, is;
    if (a.pro) tc::bn_consts_tc(a.bn, c, s, h, m, is);
    sc[c] = s;
    sh[c] = h;
  }
  __syncthreads();
  const int Ho = 2 * a.H, Wo = 2 * a.W;
  const float rh = a.H > 1 ? (float)(a.H - 1) / (float)(Ho - 1) : 0.f;
  const float rw = a.W > 1 ? (float)(a.W - 1) / (float)(Wo - 1) : 0.f;
  const int64_t total = (int64_t)a.B * Ho * Wo * a.C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % a.C);
    int64_t p = i / a.C;
    const int ox = (int)(p % Wo);
    p /= Wo;
    const int oy = (int)(p % Ho), b = (int)(p / Ho);
    const float* base = a.x + (size_t)b * a.H * a.W * a.ldx + c;
    auto act = [&](int y, int x) {
      const float v = base[((size_t)y * a.W + x) * a.ldx];
      return a.pro ? fmaxf(0.f, fmaf(v, sc[c], sh[c])) : v;
    };
    if (a.zero_insert) {
      a.up[(((size_t)b * Ho + oy) * Wo + ox) * a.ldu + c] = ((oy | ox) & 1) ? 0.f : act(oy >> 1, ox >> 1);
      continue;
    }
    const Lerp ly = lerp_of(oy, rh, a.H), lx = lerp_of(ox, rw, a.W);
    const float v = ly.w0 * (lx.w0 * act(ly.i0, lx.i0) + lx.w1 * act(ly.i0, lx.i1)) +
                    ly.w1 * (lx.w0 * act(ly.i1, lx.i0) + lx.w1 * act(ly.i1, lx.i1));
    a.up[(((size_t)b * Ho + oy) * Wo + ox) * a.ldu + c] = v;
  }
}

__global__ void __launch_bounds__(256) bilinear_bwd_kernel(BilinearArgs a) {
  griddep_wait();
  extern __shared__ float sm_b[];   // scale, shift, invstd, -mean*invstd, s1, s2   (6 * C)
  float* sc = sm_b;
  float* sh = sc + a.C;
  float* isd = sh + a.C;
  float* off = isd + a.C;
  float* s1 = off + a.C;
  float* s2 = s1 + a.C;
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    float s, h, m, is;
    tc::bn_consts_tc(a.bn, c, s, h, m, is);
    sc[c] = s;
    sh[c] = h;
    isd[c] = is;
    off[c] = -m * is;
    s1[c] = s2[c] = 0.f;
  }
  __syncthreads();
  const int Ho = 2 * a.H, Wo = 2 * a.W;
  const float rh = a.H > 1 ? (float)(a.H - 1) / (float)(Ho - 1) : 0.f;
  const float rw = a.W > 1 ? (float)(a.W - 1) / (float)(Wo - 1) : 0.f;
  float gmx = 0.f;
  const int64_t total = (int64_t)a.B * a.H * a.W * a.C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % a.C);
    int64_t p = i / a.C;
    const int x = (int)(p % a.W);
    p /= a.W;
    const int y = (int)(p % a.H), b = (int)(p / a.H);
    // transposed interpolation: output rows / columns whose two taps include this source row / column
    const int oy_lo = max(0, 2 * y - 3), oy_hi = min(Ho - 1, 2 * y + 3);
    const int ox_lo = max(0, 2 * x - 3), ox_hi = min(Wo - 1, 2 * x + 3);
    float acc = 0.f;
    if (a.zero_insert) acc = a.up[(((size_t)b * Ho + 2 * y) * Wo + 2 * x) * a.ldu + c];
    for (int oy = oy_lo; oy <= oy_hi && !a.zero_insert; ++oy) {
      const Lerp ly = lerp_of(oy, rh, a.H);
      const float wy = (ly.i0 == y ? ly.w0 : 0.f) + (ly.i1 == y ? ly.w1 : 0.f);
      if (wy == 0.f) continue;
      const float* row = a.up + (((size_t)b * Ho + oy) * Wo) * a.ldu + c;
      float r = 0.f;
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        const Lerp lx = lerp_of(ox, rw, a.W);
        const float wx = (lx.i0 == x ? lx.w0 : 0.f) + (lx.i1 == x ? lx.w1 : 0.f);
        if (wx != 0.f) r += wx * row[(size_t)ox * a.ldu];
      }
      acc += wy * r;
    }
    // ReLU mask + BatchNorm backward bookkeeping (what the dgrad epilogues do)
    const size_t pix = ((size_t)b * a.H + y) * a.W + x;
    const float xv = a.x[pix * a.ldx + c];
    const float z = fmaf(xv, sc[c], sh[c]);
    const float dz = z > 0.f ? acc : 0.f;
    const float xh = fmaf(xv, isd[c], off[c]);
    if (dz != 0.f) {
      atomicAdd(&s1[c], dz);
      atomicAdd(&s2[c], dz * xh);
    }
    const float o = sc[c] * dz;
    gmx = fmaxf(gmx, fabsf(o));
    float* g = a.G + pix * a.ldG + c;
    *g = a.g_accum ? *g + o : o;
  }
  if (a.gmax != nullptr) {
    const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(gmx));
    if ((threadIdx.x & 31) == 0 && m != 0u) atomicMax(a.gmax, m);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    atomicAdd(a.bsum + c, (double)s1[c]);
    atomicAdd(a.bsum + a.C + c, (double)s2[c]);
  }
}

__global__ void __launch_bounds__(256) convt_weight_kernel(const float* wt, float* wc, int Cin, int Cout, int KS) {
  griddep_wait();
  const int T = KS * KS, total = Cin * Cout * T;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int tap = i % T, r = i / T;
    const int ci = r % Cin, co = r / Cin;                       // i indexes wc[co][ci][tap]
    wc[i] = wt[((size_t)ci * Cout + co) * T + (T - 1 - tap)];   // flipping both axes = reversing the tap index
  }
}
__global__ void __launch_bounds__(256) convt_weight_grad_kernel(float* gc, float* gt, int Cin, int Cout, int KS) {
  griddep_wait();
  const int T = KS * KS, total = Cin * Cout * T;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int tap = i % T, r = i / T;
    const int ci = r % Cin, co = r / Cin;
    gt[((size_t)ci * Cout + co) * T + (T - 1 - tap)] += gc[i];
    gc[i] = 0.f;
  }
}

int grid_for(int64_t total) {
  int blocks = (int)((total + 255) / 256);
  const int cap = sm_count() * 8;
  if (blocks > cap) blocks = cap;
  return blocks < 1 ? 1 : blocks;
}

}  // namespace

int launch_bilinear_up(const BilinearArgs& a, cudaStream_t st) {
  PDES_REQUIRE(a.x && a.up && a.C >= 1 && a.C <= 1024, PDES_ERR_INVALID, "bilinear_up: invalid arguments");
  const int64_t total = (int64_t)a.B * 4 * a.H * a.W * a.C;
  PDES_CUDA(launch_pdl(bilinear_up_kernel, dim3(grid_for(total)), dim3(256), sizeof(float) * 2 * a.C, st, a));
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

int launch_bilinear_bwd(const BilinearArgs& a, cudaStream_t st) {
  PDES_REQUIRE(a.x && a.up && a.G && a.bsum && a.C >= 1 && a.C <= 1024, PDES_ERR_INVALID, "bilinear_bwd: invalid arguments");
  const int64_t total = (int64_t)a.B * a.H * a.W * a.C;
  PDES_CUDA(launch_pdl(bilinear_bwd_kernel, dim3(grid_for(total)), dim3(256), sizeof(float) * 6 * a.C, st, a));
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

int launch_convt_weight(const float* wt, float* wc, int Cin, int Cout, int KS, cudaStream_t st) {
  PDES_REQUIRE(wt && wc && Cin >= 1 && Cout >= 1 && KS >= 1, PDES_ERR_INVALID, "convt_weight: invalid arguments");
  PDES_CUDA(launch_pdl(convt_weight_kernel, dim3(grid_for((int64_t)Cin * Cout * KS * KS)), dim3(256), 0, st, wt, wc, Cin,
                       Cout, KS));
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

int launch_convt_weight_grad(float* gc, float* gt, int Cin, int Cout, int KS, cudaStream_t st) {
  PDES_REQUIRE(gc && gt && Cin >= 1 && Cout >= 1 && KS >= 1, PDES_ERR_INVALID, "convt_weight_grad: invalid arguments");
  PDES_CUDA(launch_pdl(convt_weight_grad_kernel, dim3(grid_for((int64_t)Cin * Cout * KS * KS)), dim3(256), 0, st, gc, gt,
                       Cin, Cout, KS));
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

}  // namespace pdes

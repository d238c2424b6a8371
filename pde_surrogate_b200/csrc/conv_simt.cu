// conv_simt.cu — fp32 CUDA-core convolution kernels (forward / dgrad via the same kernel, wgrad)
// plus the small per-channel helper kernels of the DenseED executor.
//
// These are the exact-fp32 path: every convolution of the network can run on them, and they are
// the on-device cross-check for the tcgen05 kernels.  NHWC activations; BatchNorm(+ReLU) of the
// producer is applied while the input halo tile is staged into shared memory, the nearest x2
// upsampling and the stride-2 transpose are folded into the staging address math, and the
// epilogue either stores a channel slice of a dense-block buffer and accumulates the batch
// statistics of the new channels, or (dgrad) applies the ReLU mask + BatchNorm backward and
// accumulates into the gradient buffer.
#include "conv.cuh"

namespace pdes {
namespace {

constexpr int kConvThreads = 128;
constexpr int kTile = 8;  // 8x8 output pixels per CTA

__device__ __forceinline__ void bn_consts(const BnSrc& s, int c, float& scale, float& shift,
                                          float& mean, float& invstd) {
  if (s.scale != nullptr) {
    scale = s.scale[c];
    shift = s.shift[c];
    mean = 0.f;
    invstd = 1.f;
    return;
  }
  double m, var;
  if (s.use_running) {
    m = (double)s.run_mean[c];
    var = (double)s.run_var[c];
  } else {
    m = s.sum[c] * s.inv_count;
    var = s.sumsq[c] * s.inv_count - m * m;
    if (var < 0.0) var = 0.0;
  }
  const double is = 1.0 / sqrt(var + (double)s.eps);
  const float g = s.gamma[c];
  invstd = (float)is;
  mean = (float)m;
  scale = g * invstd;
  shift = s.beta[c] - mean * scale;
}

// Stage the input halo tile of one channel chunk into shared memory, applying the prologue.
//   a_s[(hy*pitch + hx)*(KC+4) + k]
template <class A>
__device__ __forceinline__ void stage_halo(const A& a, float* a_s, const float* sc_s,
                                           const float* sh_s, int b, int iy0, int ix0, int HH,
                                           int pitch, int KC, int c0) {
  const int kq = KC >> 2;
  const int Hv = a.in_mode == IN_DIRECT ? a.Hs : 2 * a.Hs;
  const int Wv = a.in_mode == IN_DIRECT ? a.Ws : 2 * a.Ws;
  const int total = HH * HH * kq;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int hp = idx / kq, k4 = idx - hp * kq;
    const int hy = hp / HH, hx = hp - hy * HH;
    const int vy = iy0 + hy, vx = ix0 + hx;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    bool ok = vy >= 0 && vy < Hv && vx >= 0 && vx < Wv;
    if (ok && a.in_mode == IN_ZEROINS) ok = ((vy | vx) & 1) == 0;
    if (ok) {
      const int sy = a.in_mode == IN_DIRECT ? vy : (vy >> 1);
      const int sx = a.in_mode == IN_DIRECT ? vx : (vx >> 1);
      const int c = c0 + 4 * k4;
      if (a.in_nchw) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (c + i < a.Cin) v[i] = a.x[(((size_t)b * a.Cin + c + i) * a.Hs + sy) * a.Ws + sx];
      } else {
        const float* p = a.x + (((size_t)b * a.Hs + sy) * a.Ws + sx) * a.ldx + c;
        if (c + 3 < a.Cin && ((reinterpret_cast<uintptr_t>(p) & 15u) == 0)) {
          const float4 q = *reinterpret_cast<const float4*>(p);
          v[0] = q.x;
          v[1] = q.y;
          v[2] = q.z;
          v[3] = q.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (c + i < a.Cin) v[i] = p[i];
        }
      }
      if (a.pro) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ci = c + i;
          v[i] = ci < a.Cin ? fmaxf(0.f, fmaf(v[i], sc_s[ci], sh_s[ci])) : 0.f;
        }
      }
    }
    *reinterpret_cast<float4*>(a_s + (size_t)(hy * pitch + hx) * (KC + 4) + 4 * k4) =
        make_float4(v[0], v[1], v[2], v[3]);
  }
}

// ---------------------------------------------------------------------------------------
// forward / dgrad kernel: 8x8 output pixels x TN output channels per CTA, 128 threads,
// thread tile = 2 horizontally adjacent pixels x TN/4 channels (channels 16j + 4cg + i).
// ---------------------------------------------------------------------------------------
template <int TN, int KS>
__global__ void __launch_bounds__(kConvThreads) conv_simt_kernel(ConvArgs a) {
  constexpr int J = TN / 16;
  extern __shared__ __align__(16) float smem[];
  const int CinP4 = (a.Cin + 3) & ~3;
  const int KC = CinP4 < 16 ? CinP4 : 16;
  const int HH = (kTile - 1) * a.stride + KS;
  const int pitch = HH | 1;
  float* sc_s = smem;                       // CinP4
  float* sh_s = sc_s + CinP4;               // CinP4
  float* ep_s = sh_s + CinP4;               // 4*TN (scale, shift, mean, invstd of the epilogue BN)
  float* a_s = ep_s + 4 * TN;               // HH*pitch*(KC+4)
  float* w_s = a_s + HH * pitch * (KC + 4); // KS*KS*KC*TN
  float* red_s = w_s + KS * KS * KC * TN;   // 4*TN*2

  const int tiles_x = (a.Wo + kTile - 1) / kTile, tiles_y = (a.Ho + kTile - 1) / kTile;
  int bid = blockIdx.x;
  const int tx = bid % tiles_x;
  bid /= tiles_x;
  const int ty = bid % tiles_y;
  const int b = bid / tiles_y;
  const int n0 = blockIdx.y * TN;
  const int oy0 = ty * kTile, ox0 = tx * kTile;
  const int t = threadIdx.x;
  const int pp = t >> 2, cg = t & 3;
  const int row = pp >> 2, cp = pp & 3;

  if (a.pro) {
    for (int c = t; c < CinP4; c += kConvThreads) {
      float s = 0.f, h = 0.f, m, is;
      if (c < a.Cin) bn_consts(a.bn, c, s, h, m, is);
      sc_s[c] = s;
      sh_s[c] = h;
    }
  }
  if (a.epi == EPI_BNBWD) {
    for (int n = t; n < TN; n += kConvThreads) {
      float s = 0.f, h = 0.f, m = 0.f, is = 0.f;
      if (n0 + n < a.Cout) bn_consts(a.fbn, n0 + n, s, h, m, is);
      ep_s[n] = s;
      ep_s[TN + n] = h;
      ep_s[2 * TN + n] = m;
      ep_s[3 * TN + n] = is;
    }
  }

  float acc[2][J][4];
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[p][j][i] = 0.f;

  const int iy0 = oy0 * a.stride - a.pad, ix0 = ox0 * a.stride - a.pad;
  for (int c0 = 0; c0 < CinP4; c0 += KC) {
    __syncthreads();
    stage_halo(a, a_s, sc_s, sh_s, b, iy0, ix0, HH, pitch, KC, c0);
    {
      const int nq = TN / 4;
      const int total = KS * KS * KC * nq;
      for (int idx = t; idx < total; idx += kConvThreads) {
        const int n4 = idx % nq;
        const int r = idx / nq;
        const int k = r % KC, tap = r / KC;
        const int c = c0 + k, n = n0 + 4 * n4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < a.CinP && n < a.CoP)
          v = *reinterpret_cast<const float4*>(a.w + ((size_t)tap * a.CinP + c) * a.CoP + n);
        *reinterpret_cast<float4*>(w_s + (size_t)(tap * KC + k) * TN + 4 * n4) = v;
      }
    }
    __syncthreads();
    const int kq = KC >> 2;
#pragma unroll
    for (int ky = 0; ky < KS; ++ky) {
#pragma unroll
      for (int kx = 0; kx < KS; ++kx) {
        const float* ap0 =
            a_s + (size_t)((row * a.stride + ky) * pitch + (2 * cp) * a.stride + kx) * (KC + 4);
        const float* ap1 = ap0 + a.stride * (KC + 4);
        const float* wp = w_s + (size_t)((ky * KS + kx) * KC) * TN + 4 * cg;
        for (int k4 = 0; k4 < kq; ++k4) {
          const float4 a0 = *reinterpret_cast<const float4*>(ap0 + 4 * k4);
          const float4 a1 = *reinterpret_cast<const float4*>(ap1 + 4 * k4);
          const float a0v[4] = {a0.x, a0.y, a0.z, a0.w};
          const float a1v[4] = {a1.x, a1.y, a1.z, a1.w};
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int j = 0; j < J; ++j) {
              const float4 w4 = *reinterpret_cast<const float4*>(wp + (4 * k4 + kk) * TN + 16 * j);
              acc[0][j][0] = fmaf(a0v[kk], w4.x, acc[0][j][0]);
              acc[0][j][1] = fmaf(a0v[kk], w4.y, acc[0][j][1]);
              acc[0][j][2] = fmaf(a0v[kk], w4.z, acc[0][j][2]);
              acc[0][j][3] = fmaf(a0v[kk], w4.w, acc[0][j][3]);
              acc[1][j][0] = fmaf(a1v[kk], w4.x, acc[1][j][0]);
              acc[1][j][1] = fmaf(a1v[kk], w4.y, acc[1][j][1]);
              acc[1][j][2] = fmaf(a1v[kk], w4.z, acc[1][j][2]);
              acc[1][j][3] = fmaf(a1v[kk], w4.w, acc[1][j][3]);
            }
          }
        }
      }
    }
  }

  // ---- epilogue ----------------------------------------------------------------------
  // optional 2x2 sum-pool: horizontal in-thread, vertical across lanes (row r <-> r^1 = lane^16)
  int npx = 2;
  int py[2], px[2];
  bool act = true;
  float gmx = 0.f;
  if (a.pool) {
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v = acc[0][j][i] + acc[1][j][i];
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        acc[0][j][i] = v;
      }
    npx = 1;
    act = (row & 1) == 0;
    py[0] = (oy0 + row) >> 1;
    px[0] = (ox0 + 2 * cp) >> 1;
    py[1] = px[1] = 0;
  } else {
    py[0] = py[1] = oy0 + row;
    px[0] = ox0 + 2 * cp;
    px[1] = px[0] + 1;
  }
  const int Hd = a.pool ? a.Ho / 2 : a.Ho, Wd = a.pool ? a.Wo / 2 : a.Wo;

  float s1[J][4], s2[J][4];
#pragma unroll
  for (int j = 0; j < J; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) s1[j][i] = s2[j][i] = 0.f;

#pragma unroll
  for (int p = 0; p < 2; ++p) {
    if (p >= npx || !act || py[p] >= Hd || px[p] >= Wd) continue;
    const size_t pix = ((size_t)b * Hd + py[p]) * Wd + px[p];
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int nl = 16 * j + 4 * cg;
      const int n = n0 + nl;
      if (n >= a.Cout) continue;
      if (a.epi == EPI_NHWC) {
        float* dst = a.y + pix * a.ldy + a.coff + n;
        if (n + 3 < a.Cout && ((a.ldy | a.coff) & 3) == 0) {
          *reinterpret_cast<float4*>(dst) =
              make_float4(acc[p][j][0], acc[p][j][1], acc[p][j][2], acc[p][j][3]);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (n + i < a.Cout) dst[i] = acc[p][j][i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          s1[j][i] += acc[p][j][i];
          s2[j][i] += acc[p][j][i] * acc[p][j][i];
        }
      } else if (a.epi == EPI_NCHW) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (n + i < a.Cout)
            a.y[(((size_t)b * a.Cout + n + i) * Hd + py[p]) * Wd + px[p]] = acc[p][j][i];
      } else {  // EPI_BNBWD
        const float* xs = a.fx + pix * a.ldfx + n;
        float* gp = a.G + pix * a.ldG + n;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (n + i >= a.Cout) continue;
          const float xv = xs[i];
          const float z = fmaf(xv, ep_s[nl + i], ep_s[TN + nl + i]);
          const float dz = z > 0.f ? acc[p][j][i] : 0.f;
          const float xh = (xv - ep_s[2 * TN + nl + i]) * ep_s[3 * TN + nl + i];
          s1[j][i] += dz;
          s2[j][i] += dz * xh;
          const float g = ep_s[nl + i] * dz;
          const float o = a.g_accum ? gp[i] + g : g;
          gp[i] = o;
          gmx = fmaxf(gmx, fabsf(o));
        }
      }
    }
  }

  if (a.epi == EPI_BNBWD && a.gmax != nullptr) {  // running |G| maximum (dynamic fp16 scale of the dY pieces)
    const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(gmx));
    if ((t & 31) == 0 && m != 0u) atomicMax(a.gmax, m);
  }
  const bool want_red = (a.epi == EPI_NHWC && a.o_sum != nullptr) || a.epi == EPI_BNBWD;
  if (want_red) {
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float u = s1[j][i], v = s2[j][i];
#pragma unroll
        for (int o = 4; o <= 16; o <<= 1) {
          u += __shfl_xor_sync(0xffffffffu, u, o);
          v += __shfl_xor_sync(0xffffffffu, v, o);
        }
        s1[j][i] = u;
        s2[j][i] = v;
      }
    const int lane = t & 31, wid = t >> 5;
    if (lane < 4) {
#pragma unroll
      for (int j = 0; j < J; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int nl = 16 * j + 4 * cg + i;
          red_s[(wid * TN + nl) * 2 + 0] = s1[j][i];
          red_s[(wid * TN + nl) * 2 + 1] = s2[j][i];
        }
    }
    __syncthreads();
    for (int nl = t; nl < TN; nl += kConvThreads) {
      const int n = n0 + nl;
      if (n >= a.Cout) continue;
      double u = 0.0, v = 0.0;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        u += (double)red_s[(w * TN + nl) * 2 + 0];
        v += (double)red_s[(w * TN + nl) * 2 + 1];
      }
      if (a.epi == EPI_NHWC) {
        atomicAdd(a.o_sum + n, u);
        atomicAdd(a.o_sumsq + n, v);
      } else {
        atomicAdd(a.bsum + n, u);
        atomicAdd(a.bsum + a.Cout + n, v);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// wgrad: dW[co][ci][ky][kx] += sum_p dY[p,co] * a[p*s + tap - pad, ci]
// CTA = one image x 16 input channels x TN output channels x TR tap rows; thread = (ci, co4)
// accumulating TR*KS x 4 weights over the image's pixels, then atomics into the OIHW gradient.
// ---------------------------------------------------------------------------------------
template <int TN, int KS, int TR>
__global__ void __launch_bounds__(kConvThreads) wgrad_simt_kernel(WgradArgs a) {
  constexpr int NQ = TN / 4;
  constexpr int GROUP = 16 * NQ;            // threads per pixel group
  constexpr int NG = kConvThreads / GROUP;  // pixel groups
  extern __shared__ __align__(16) float smem[];
  const int CinP4 = (a.Cin + 3) & ~3;
  const int KC = CinP4 < 16 ? CinP4 : 16;
  const int HH = (kTile - 1) * a.stride + KS;
  const int pitch = HH | 1;
  float* sc_s = smem;
  float* sh_s = sc_s + CinP4;
  float* a_s = sh_s + CinP4;                   // HH*pitch*(KC+4)
  float* dy_s = a_s + HH * pitch * (KC + 4);   // 64*TN

  const int b = blockIdx.x;
  const int c0 = blockIdx.y * KC;
  constexpr int NROWP = KS / TR;
  const int n0 = (blockIdx.z / NROWP) * TN;
  const int ky0 = (blockIdx.z % NROWP) * TR;
  const int t = threadIdx.x;
  const int tl = t % GROUP, g = t / GROUP;
  const int co4 = tl % NQ, ci = tl / NQ;

  if (a.pro) {
    for (int c = t; c < CinP4; c += kConvThreads) {
      float s = 0.f, h = 0.f, m, is;
      if (c < a.Cin) bn_consts(a.bn, c, s, h, m, is);
      sc_s[c] = s;
      sh_s[c] = h;
    }
  }
  float acc[TR][KS][4];
#pragma unroll
  for (int r = 0; r < TR; ++r)
#pragma unroll
    for (int k = 0; k < KS; ++k)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[r][k][i] = 0.f;

  const int tiles_x = (a.Wo + kTile - 1) / kTile, tiles_y = (a.Ho + kTile - 1) / kTile;
  for (int tile = 0; tile < tiles_x * tiles_y; ++tile) {
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int oy0 = ty * kTile, ox0 = tx * kTile;
    __syncthreads();
    stage_halo(a, a_s, sc_s, sh_s, b, oy0 * a.stride - a.pad, ox0 * a.stride - a.pad, HH, pitch, KC,
               c0);
    for (int idx = t; idx < 64 * NQ; idx += kConvThreads) {
      const int p = idx / NQ, q = idx - p * NQ;
      const int oy = oy0 + (p >> 3), ox = ox0 + (p & 7);
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (oy < a.Ho && ox < a.Wo) {
        const int n = n0 + 4 * q;
        if (a.dy_nchw) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (n + i < a.Cout) v[i] = a.dy[(((size_t)b * a.Cout + n + i) * a.Ho + oy) * a.Wo + ox];
        } else {
          const float* src = a.dy + (((size_t)b * a.Ho + oy) * a.Wo + ox) * a.lddy + n;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (n + i < a.Cout) v[i] = src[i];
        }
      }
      *reinterpret_cast<float4*>(dy_s + (size_t)p * TN + 4 * q) = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncthreads();
    if (ci < KC) {
      for (int p = g; p < 64; p += NG) {
        const float4 d = *reinterpret_cast<const float4*>(dy_s + (size_t)p * TN + 4 * co4);
        const int pyy = (p >> 3) * a.stride, pxx = (p & 7) * a.stride;
#pragma unroll
        for (int r = 0; r < TR; ++r) {
#pragma unroll
          for (int k = 0; k < KS; ++k) {
            const float av = a_s[(size_t)((pyy + ky0 + r) * pitch + pxx + k) * (KC + 4) + ci];
            acc[r][k][0] = fmaf(av, d.x, acc[r][k][0]);
            acc[r][k][1] = fmaf(av, d.y, acc[r][k][1]);
            acc[r][k][2] = fmaf(av, d.z, acc[r][k][2]);
            acc[r][k][3] = fmaf(av, d.w, acc[r][k][3]);
          }
        }
      }
    }
  }
  const int c = c0 + ci;
  if (ci < KC && c < a.Cin) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + 4 * co4 + i;
      if (n >= a.Cout) continue;
#pragma unroll
      for (int r = 0; r < TR; ++r)
#pragma unroll
        for (int k = 0; k < KS; ++k)
          atomicAdd(a.dw + (((size_t)n * a.Cin + c) * KS + ky0 + r) * KS + k, acc[r][k][i]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// helper kernels
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fix_dy_kernel(FixDyArgs a) {
  griddep_wait();
  __shared__ float c1_s[256], c2_s[256], mean_s[256], is_s[256];
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    const double m = a.sum[c] * a.inv_count;
    double var = a.sumsq[c] * a.inv_count - m * m;
    if (var < 0.0) var = 0.0;
    const double is = 1.0 / sqrt(var + (double)a.eps);
    double c1 = 0.0, c2 = 0.0;
    for (int l = 0; l < a.n_cons; ++l) {
      const double sc = (double)a.cons_gamma[l][c] * (double)(float)is;
      c1 += sc * a.cons_bsum[l][c];
      c2 += sc * a.cons_bsum[l][a.cons_C[l] + c];
    }
    c1_s[c] = (float)(c1 * a.inv_count);
    c2_s[c] = (float)(c2 * a.inv_count);
    mean_s[c] = (float)m;
    is_s[c] = (float)is;
  }
  __syncthreads();
  const int64_t total = a.npix * a.C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / a.C;
    const int c = (int)(i - p * a.C);
    const float xh = (a.X[p * a.ldX + c] - mean_s[c]) * is_s[c];
    float* g = a.G + p * a.ldG + c;
    float v = *g - c1_s[c] - xh * c2_s[c];
    if (a.drop_mask != nullptr) v *= a.drop_mask[(p / a.pix_per_img) * a.C + c];   // d(y*mask)/dy
    *g = v;
  }
}

// nn.Dropout2d forward on a freshly written NHWC slice + the batch statistics the next BatchNorm needs
__global__ void __launch_bounds__(256) dropout_fwd_kernel(float* y, int ld, int C, int64_t npix, int64_t pix_per_img,
                                                          const float* __restrict__ mask, double* o_sum,
                                                          double* o_sumsq) {
  griddep_wait();
  __shared__ float s1[256], s2[256];
  for (int c = threadIdx.x; c < C; c += blockDim.x) s1[c] = s2[c] = 0.f;
  __syncthreads();
  const int64_t total = npix * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / C;
    const int c = (int)(i - p * C);
    float* q = y + p * ld + c;
    const float v = *q * mask[(p / pix_per_img) * C + c];
    *q = v;
    if (o_sum != nullptr && v != 0.f) {
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

__global__ void __launch_bounds__(256) pack_weights_kernel(const PackDesc* tab) {
  const PackDesc d = tab[blockIdx.y];
  const int taps = d.KS * d.KS;
  const int nf = taps * d.CinP * d.CoP;
  const int nb = d.wb ? taps * d.CoutPb * d.CiPb : 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nf + nb; i += gridDim.x * blockDim.x) {
    if (i < nf) {
      const int co = i % d.CoP;
      const int r = i / d.CoP;
      const int ci = r % d.CinP, tap = r / d.CinP;
      float v = 0.f;
      if (co < d.Cout && ci < d.Cin) v = d.w[((size_t)co * d.Cin + ci) * taps + tap];
      d.wf[i] = v;
    } else {
      const int k = i - nf;
      const int ci = k % d.CiPb;
      const int r = k / d.CiPb;
      const int co = r % d.CoutPb, tap = r / d.CoutPb;
      float v = 0.f;
      if (co < d.Cout && ci < d.Cin) v = d.w[((size_t)co * d.Cin + ci) * taps + (taps - 1 - tap)];
      d.wb[k] = v;
    }
  }
}

// running_mean = (1-m) rm + m mean ; running_var = (1-m) rv + m var*N/(N-1)
__global__ void bn_running_update_kernel(const BnLayerDesc* tab, float momentum, int B) {
  griddep_wait();
  BnLayerDesc d = tab[blockIdx.y];
  d.count *= (double)B;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d.C) return;
  const double mean = d.sum[c] / d.count;
  double var = d.sumsq[c] / d.count - mean * mean;
  if (var < 0.0) var = 0.0;
  const double unb = d.count > 1.0 ? var * d.count / (d.count - 1.0) : var;
  d.run_mean[c] = (float)((1.0 - (double)momentum) * (double)d.run_mean[c] + (double)momentum * mean);
  d.run_var[c] = (float)((1.0 - (double)momentum) * (double)d.run_var[c] + (double)momentum * unb);
}

__global__ void bn_param_grad_kernel(const BnLayerDesc* tab) {
  griddep_wait();
  const BnLayerDesc d = tab[blockIdx.y];
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d.C) return;
  d.dbeta[c] += (float)d.bsum[c];
  d.dgamma[c] += (float)d.bsum[d.C + c];
}

size_t conv_smem_bytes(int Cin, int KS, int stride, int TN) {
  const int CinP4 = (Cin + 3) & ~3;
  const int KC = CinP4 < 16 ? CinP4 : 16;
  const int HH = (kTile - 1) * stride + KS;
  const int pitch = HH | 1;
  return sizeof(float) * ((size_t)2 * CinP4 + 4 * TN + (size_t)HH * pitch * (KC + 4) +
                          (size_t)KS * KS * KC * TN + 4 * TN * 2);
}
size_t wgrad_smem_bytes(int Cin, int KS, int stride, int TN) {
  const int CinP4 = (Cin + 3) & ~3;
  const int KC = CinP4 < 16 ? CinP4 : 16;
  const int HH = (kTile - 1) * stride + KS;
  const int pitch = HH | 1;
  return sizeof(float) * ((size_t)2 * CinP4 + (size_t)HH * pitch * (KC + 4) + (size_t)64 * TN);
}

template <int TN, int KS>
int launch_conv_t(const ConvArgs& a, cudaStream_t st) {
  const size_t smem = conv_smem_bytes(a.Cin, KS, a.stride, TN);
  PDES_REQUIRE(smem <= 227 * 1024, PDES_ERR_UNSUPPORTED,
               "conv_simt: tile needs %zu bytes of shared memory", smem);
  PDES_ENSURE_SMEM((conv_simt_kernel<TN, KS>), smem);
  const int tiles = ((a.Wo + kTile - 1) / kTile) * ((a.Ho + kTile - 1) / kTile) * a.B;
  dim3 grid(tiles, (a.Cout + TN - 1) / TN);
  conv_simt_kernel<TN, KS><<<grid, kConvThreads, smem, st>>>(a);
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

template <int TN, int KS, int TR>
int launch_wgrad_t(const WgradArgs& a, cudaStream_t st) {
  const size_t smem = wgrad_smem_bytes(a.Cin, KS, a.stride, TN);
  PDES_REQUIRE(smem <= 227 * 1024, PDES_ERR_UNSUPPORTED,
               "wgrad_simt: tile needs %zu bytes of shared memory", smem);
  PDES_ENSURE_SMEM((wgrad_simt_kernel<TN, KS, TR>), smem);
  const int CinP4 = (a.Cin + 3) & ~3;
  const int KC = CinP4 < 16 ? CinP4 : 16;
  dim3 grid(a.B, (CinP4 + KC - 1) / KC, ((a.Cout + TN - 1) / TN) * (KS / TR));
  wgrad_simt_kernel<TN, KS, TR><<<grid, kConvThreads, smem, st>>>(a);
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

}  // namespace

int launch_conv_simt(const ConvArgs& a, cudaStream_t st) {
  PDES_REQUIRE(a.stride == 1 || a.stride == 2, PDES_ERR_UNSUPPORTED, "conv: stride %d", a.stride);
  PDES_REQUIRE(!a.pool || ((a.Ho | a.Wo) & 1) == 0, PDES_ERR_INVALID, "conv: pool needs even size");
  PDES_REQUIRE((a.CoP & 3) == 0, PDES_ERR_INVALID, "conv: CoP must be a multiple of 4");
  const bool wide = a.Cout > 32;
  switch (a.KS) {
    case 1: return wide ? launch_conv_t<64, 1>(a, st) : launch_conv_t<16, 1>(a, st);
    case 3: return wide ? launch_conv_t<64, 3>(a, st) : launch_conv_t<16, 3>(a, st);
    case 5: return wide ? launch_conv_t<64, 5>(a, st) : launch_conv_t<16, 5>(a, st);
    case 7: return wide ? launch_conv_t<64, 7>(a, st) : launch_conv_t<16, 7>(a, st);
    default: break;
  }
  set_error("conv: kernel size %d not supported (1,3,5,7)", a.KS);
  return PDES_ERR_UNSUPPORTED;
}

int launch_wgrad_simt(const WgradArgs& a, cudaStream_t st) {
  PDES_REQUIRE(a.stride == 1 || a.stride == 2, PDES_ERR_UNSUPPORTED, "wgrad: stride %d", a.stride);
  const bool wide = a.Cout > 16;
  switch (a.KS) {
    case 1: return wide ? launch_wgrad_t<32, 1, 1>(a, st) : launch_wgrad_t<16, 1, 1>(a, st);
    case 3: return wide ? launch_wgrad_t<32, 3, 3>(a, st) : launch_wgrad_t<16, 3, 3>(a, st);
    case 5: return wide ? launch_wgrad_t<32, 5, 1>(a, st) : launch_wgrad_t<16, 5, 1>(a, st);
    case 7: return wide ? launch_wgrad_t<32, 7, 1>(a, st) : launch_wgrad_t<16, 7, 1>(a, st);
    default: break;
  }
  set_error("wgrad: kernel size %d not supported (1,3,5,7)", a.KS);
  return PDES_ERR_UNSUPPORTED;
}

int launch_fix_dy(const FixDyArgs& a, cudaStream_t st) {
  PDES_REQUIRE(a.C <= 256, PDES_ERR_UNSUPPORTED, "fix_dy: slice of %d channels (max 256)", a.C);
  const int64_t total = a.npix * a.C;
  int blocks = (int)((total + 255) / 256);
  const int cap = sm_count() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  PDES_CUDA(launch_pdl(fix_dy_kernel, dim3(blocks), dim3(256), 0, st, a));
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

int launch_dropout_fwd(float* y, int ld, int C, int64_t npix, int64_t pix_per_img, const float* mask,
                       double* o_sum, double* o_sumsq, cudaStream_t st) {
  PDES_REQUIRE(y && mask && C >= 1 && C <= 256, PDES_ERR_INVALID, "dropout_fwd: invalid arguments (C %d)", C);
  int blocks = (int)((npix * C + 255) / 256);
  const int cap = sm_count() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  PDES_CUDA(launch_pdl(dropout_fwd_kernel, dim3(blocks), dim3(256), 0, st, y, ld, C, npix, pix_per_img, mask, o_sum, o_sumsq));
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

int launch_pack_weights(const PackDesc* dev_table, int n_layers, int max_elems, cudaStream_t st) {
  int bx = (max_elems + 255) / 256;
  if (bx > 64) bx = 64;
  if (bx < 1) bx = 1;
  pack_weights_kernel<<<dim3(bx, n_layers), 256, 0, st>>>(dev_table);
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

int launch_bn_running_update(const BnLayerDesc* dev_table, int n, int maxC, float momentum,
                             int B, cudaStream_t st) {
  if (n == 0) return PDES_OK;
  PDES_CUDA(launch_pdl(bn_running_update_kernel, dim3((maxC + 127) / 128, n), dim3(128), 0, st, dev_table, momentum, B));
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

int launch_bn_param_grad(const BnLayerDesc* dev_table, int n, int maxC, cudaStream_t st) {
  if (n == 0) return PDES_OK;
  PDES_CUDA(launch_pdl(bn_param_grad_kernel, dim3((maxC + 127) / 128, n), dim3(128), 0, st, dev_table));
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

}  // namespace pdes

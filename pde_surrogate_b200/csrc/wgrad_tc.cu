// wgrad_tc.cu — tcgen05 weight-gradient kernel.
//
//   dW[tap][ci][co] = sum over pixels p of  a[p + tap, ci] * dY[p, co]
//
// GEMM view per filter tap: D[M = 128 input channels, N = 16 output channels] += A^T * B with the
// reduction (GEMM-K) over PIXELS.  Both operands are staged pixel-major ([pixel][channel], the
// layout NHWC gives for free), i.e. they are MN-major UMMA operands: core matrix = 8 pixels x 16 B
// (8 bf16 channels).  One tcgen05.mma (K = 16) consumes two 8-pixel tile rows; the nine taps read
// the SAME staged halo tile through shifted descriptors and own nine accumulators in TMEM.
//
// Precision: tcgen05 only transposes 16-bit operands (kind::tf32 with MN-major operands returns
// zeros on sm_100a — measured), so fp32 accuracy comes from an exact three-way bf16 split
// x = b1 + b2 + b3 (8+8+8 mantissa bits) and the six products whose weight is >= 2^-16:
//   pass 0 CTAs stage a1 and multiply by [d1|d2|d3]   (one N'=48 instruction)
//   pass 1 CTAs stage a2 and multiply by [d1|d2]      (N'=32)
//   pass 2 CTAs stage a3 and multiply by  d1          (N =16)
// Every product lands in its own TMEM columns (big and small terms never share a truncating
// accumulator) and the epilogue adds them in fp32.  CTAs split the pixel range (split-K) and finish
// with coalesced vector reductions (red.global.add.v4.f32) into a [tap][ci][co] staging buffer
// that a small kernel folds into the OIHW gradient.
#include "conv.cuh"
#include "conv_tc.cuh"
#include "tc_common.cuh"
#include <stdlib.h>
#include <cuda_bf16.h>

namespace pdes {
namespace {

using namespace tc;

constexpr int kThreads = 192;
constexpr int kTH = 16, kTW = 8;
constexpr int kMC = 128;  // input channels per CTA (GEMM M)
constexpr int kNC = 16;   // output channels per CTA (GEMM N)
constexpr int kPasses = 3;

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// exact 3-way bf16 split; returns piece `which` (0,1,2) as raw bf16 bits
__device__ __forceinline__ uint32_t bf16_piece(float x, int which) {
  const __nv_bfloat16 b1 = __float2bfloat16_rn(x);
  if (which == 0) return (uint32_t)__bfloat16_as_ushort(b1);
  const float r1 = x - __bfloat162float(b1);
  const __nv_bfloat16 b2 = __float2bfloat16_rn(r1);
  if (which == 1) return (uint32_t)__bfloat16_as_ushort(b2);
  const float r2 = r1 - __bfloat162float(b2);
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(r2));
}
__device__ __forceinline__ uint32_t pack2(uint32_t lo, uint32_t hi) { return lo | (hi << 16); }

// instruction descriptor: D fp32, A/B bf16, both MN-major
__device__ __forceinline__ uint32_t make_idesc_bf16_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void bn_consts_w(const BnSrc& s, int c, float& scale, float& shift) {
  if (s.scale != nullptr) {
    scale = s.scale[c];
    shift = s.shift[c];
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
  const float invstd = (float)(1.0 / sqrt(var + (double)s.eps));
  scale = s.gamma[c] * invstd;
  shift = s.beta[c] - (float)m * scale;
}

template <int KS>
__global__ void __launch_bounds__(kThreads, 1) wgrad_tc_kernel(TcWgradArgs t) {
  constexpr int T = KS * KS;
  constexpr int HWp = kTW + KS - 1;
  constexpr int HP = (kTH + KS - 1) * HWp;
  constexpr int HPpad = HP | 1;
  constexpr int QA = kMC / 8;                       // channel octets (8 bf16 = 16 B) of the A tile
  constexpr uint32_t A_BYTES = QA * HPpad * 16u;    // one bf16 plane
  constexpr uint32_t B_OCT = 128u * 16u;            // one co-octet: 128 pixels x 16 B
  constexpr uint32_t B_STAGE = 6u * B_OCT;          // up to [d1|d2|d3] x 2 octets
  const WgradArgs& a = t.w;
  const int pass = blockIdx.z;                      // which bf16 piece of `a` this CTA stages
  const int NP = kPasses - pass;                    // dY pieces multiplied: 3, 2, 1
  const int NW = NP * kNC;                          // accumulator columns per tap

  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);        // [2]
  uint64_t* empty = full + 2;                                // [2]
  uint64_t* acc_full = full + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 5);
  float* sc_s = reinterpret_cast<float*>(smem + 128);        // kMC
  float* sh_s = sc_s + kMC;
  unsigned char* A_s = smem + 128 + 2 * kMC * 4;             // 2 stages
  unsigned char* B_s = A_s + 2 * (size_t)A_BYTES;            // 2 stages x 6 octets

  const int n_ci_tiles = (a.Cin + kMC - 1) / kMC;
  const int ci_tile = blockIdx.y % n_ci_tiles, co_tile = blockIdx.y / n_ci_tiles;
  const int c0 = ci_tile * kMC, n0 = co_tile * kNC;
  const int tiles_x = (a.Wo + kTW - 1) / kTW, tiles_y = (a.Ho + kTH - 1) / kTH;
  const int n_tiles = tiles_x * tiles_y * a.B;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = (uint32_t)(T * NW) <= 32u ? 32u : ((uint32_t)(T * NW) <= 64u ? 64u
                             : ((uint32_t)(T * NW) <= 128u ? 128u : ((uint32_t)(T * NW) <= 256u ? 256u : 512u)));

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 128);
      mbar_init(&empty[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  if (warp >= 2 && a.pro) {
    for (int c = threadIdx.x - 64; c < kMC; c += 128) {
      float s = 0.f, h = 0.f;
      if (c0 + c < a.Cin) bn_consts_w(a.bn, c0 + c, s, h);
      sc_s[c] = s;
      sh_s[c] = h;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // number of pixel tiles this CTA processes (split-K over pixels)
  int my_tiles = 0;
  for (int pt = blockIdx.x; pt < n_tiles; pt += gridDim.x) ++my_tiles;

  if (warp == 1) {
    if (lane == 0 && my_tiles > 0) {
      const uint32_t idesc = make_idesc_bf16_mn(128, NW);
      // MN-major canonical layout ((8,1,m),(8,k)) : ((1,8,SBO),(8,LBO)): SBO strides along the
      // channels (next octet), LBO along the pixels (next 8-pixel tile row)
      const uint32_t sbo_a = HPpad * 16u, lbo_a = HWp * 16u;
      const uint32_t sbo_b = B_OCT, lbo_b = 128u;
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it & 1;
        mbar_wait(&full[s], (uint32_t)((it >> 1) & 1));
        tc_fence_after();
        const uint32_t a_base = smem_u32(A_s + (size_t)s * A_BYTES);
        const uint32_t b_base = smem_u32(B_s + (size_t)s * B_STAGE);
        for (int r = 0; r < kTH; r += 2) {
          const uint64_t bd = make_desc(b_base + (uint32_t)r * 128u, lbo_b, sbo_b);
#pragma unroll
          for (int tap = 0; tap < T; ++tap) {
            const uint64_t ad =
                make_desc(a_base + (uint32_t)((r + tap / KS) * HWp + (tap % KS)) * 16u, lbo_a, sbo_a);
            umma_bf16(tmem_base + (uint32_t)(tap * NW), ad, bd, idesc, (it | r) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty[s]);
      }
      umma_commit(acc_full);
    }
  } else if (warp >= 2) {
    const int tt = threadIdx.x - 64;
    const int Hv = a.in_mode == IN_DIRECT ? a.Hs : 2 * a.Hs;
    const int Wv = a.in_mode == IN_DIRECT ? a.Ws : 2 * a.Ws;
    int it = 0;
    for (int pt = blockIdx.x; pt < n_tiles; pt += gridDim.x, ++it) {
      const int s = it & 1;
      int rem = pt;
      const int tx = rem % tiles_x;
      rem /= tiles_x;
      const int ty = rem % tiles_y;
      const int b = rem / tiles_y;
      const int oy0 = ty * kTH, ox0 = tx * kTW;
      const int iy0 = oy0 - a.pad, ix0 = ox0 - a.pad;
      mbar_wait(&empty[s], (uint32_t)(((it >> 1) & 1) ^ 1));
      unsigned char* As = A_s + (size_t)s * A_BYTES;
      unsigned char* Bs = B_s + (size_t)s * B_STAGE;
      // ---- A: BN+ReLU'd halo tile, bf16 piece `pass` of every value ----
      constexpr int TOTAL_A = HP * QA;
      constexpr int BATCH = 6;
      for (int base = 0; base < TOTAL_A; base += BATCH * 128) {
        float4 raw[BATCH][2];
        int meta[BATCH];
#pragma unroll
        for (int j = 0; j < BATCH; ++j) {
          const int i = base + tt + j * 128;
          raw[j][0] = raw[j][1] = make_float4(0.f, 0.f, 0.f, 0.f);
          meta[j] = -1;
          if (i < TOTAL_A) {
            const int q = i % QA, hp = i / QA;
            const int hy = hp / HWp, hx = hp - hy * HWp;
            const int vy = iy0 + hy, vx = ix0 + hx;
            const bool in = vy >= 0 && vy < Hv && vx >= 0 && vx < Wv;
            meta[j] = (hp << 8) | (q << 1) | (in ? 1 : 0);
            const int c = c0 + 8 * q;
            if (in && c < a.Cin) {
              const int sy = a.in_mode == IN_DIRECT ? vy : (vy >> 1);
              const int sx = a.in_mode == IN_DIRECT ? vx : (vx >> 1);
              const float* p = a.x + (((size_t)b * a.Hs + sy) * a.Ws + sx) * a.ldx + c;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                if (c + 4 * h + 3 < a.Cin) {
                  raw[j][h] = __ldg(reinterpret_cast<const float4*>(p + 4 * h));
                } else {
                  float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    if (c + 4 * h + k < a.Cin) v[k] = p[4 * h + k];
                  raw[j][h] = make_float4(v[0], v[1], v[2], v[3]);
                }
              }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < BATCH; ++j) {
          if (meta[j] < 0) continue;
          const int hp = meta[j] >> 8, q = (meta[j] >> 1) & 127;
          float v[8] = {raw[j][0].x, raw[j][0].y, raw[j][0].z, raw[j][0].w,
                        raw[j][1].x, raw[j][1].y, raw[j][1].z, raw[j][1].w};
          const bool in = meta[j] & 1;
          if (a.pro) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int cl = 8 * q + k;
              v[k] = (in && c0 + cl < a.Cin) ? fmaxf(0.f, fmaf(v[k], sc_s[cl], sh_s[cl])) : 0.f;
            }
          }
          uint32_t o[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = bf16_piece(v[k], pass);
          *reinterpret_cast<uint4*>(As + ((size_t)q * HPpad + hp) * 16) =
              make_uint4(pack2(o[0], o[1]), pack2(o[2], o[3]), pack2(o[4], o[5]), pack2(o[6], o[7]));
        }
      }
      // ---- B: dY tile, 128 pixels x 16 channels; octets [2*piece, 2*piece+1] hold bf16 piece ----
      for (int i = tt; i < 128 * (kNC / 8); i += 128) {
        const int q = i % (kNC / 8), p = i / (kNC / 8);
        const int oy = oy0 + (p >> 3), ox = ox0 + (p & 7);
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (oy < a.Ho && ox < a.Wo) {
          const int n = n0 + 8 * q;
          const float* src = a.dy + (((size_t)b * a.Ho + oy) * a.Wo + ox) * a.lddy + n;
          if (n + 7 < a.Cout && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0)) {
            const float4 f0 = __ldg(reinterpret_cast<const float4*>(src));
            const float4 f1 = __ldg(reinterpret_cast<const float4*>(src + 4));
            v[0] = f0.x; v[1] = f0.y; v[2] = f0.z; v[3] = f0.w;
            v[4] = f1.x; v[5] = f1.y; v[6] = f1.z; v[7] = f1.w;
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (n + k < a.Cout) v[k] = src[k];
          }
        }
        for (int piece = 0; piece < NP; ++piece) {
          uint32_t o[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = bf16_piece(v[k], piece);
          *reinterpret_cast<uint4*>(Bs + (size_t)(2 * piece + q) * B_OCT + (size_t)p * 16) =
              make_uint4(pack2(o[0], o[1]), pack2(o[2], o[3]), pack2(o[4], o[5]), pack2(o[6], o[7]));
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&full[s]);
    }
    // ---- epilogue: TMEM -> coalesced vector reductions into dWp[tap][ci][co] -----------
    if (my_tiles > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
      const int quarter = warp & 3;
      const int ci = c0 + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
      for (int tap = 0; tap < T; ++tap) {
        float v[16];
        tmem_ld16(taddr + (uint32_t)(tap * NW + (NP - 1) * kNC), v);  // smallest terms first
        for (int piece = NP - 2; piece >= 0; --piece) {
          float x[16];
          tmem_ld16(taddr + (uint32_t)(tap * NW + piece * kNC), x);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += x[i];
        }
        if (t.dbg & 2) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 1.0f;
        }
        if (ci < a.Cin) {
          float* dst = t.dwp + ((size_t)tap * t.ci_pad + ci) * t.co_pad + n0;
#pragma unroll
          for (int i = 0; i < 16; i += 4) red_add_v4(dst + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// dW_OIHW[co][ci][tap] += dWp[tap][ci][co]; dWp is cleared for the next step
__global__ void __launch_bounds__(256) wgrad_unpack_kernel(const TcWgradUnpack* tab) {
  const TcWgradUnpack d = tab[blockIdx.y];
  const int T = d.KS * d.KS;
  const int total = d.Cout * d.Cin * T;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int tap = i % T;
    const int r = i / T;
    const int ci = r % d.Cin, co = r / d.Cin;
    float* src = d.dwp + ((size_t)tap * d.ci_pad + ci) * d.co_pad + co;
    d.dw[i] += *src;
    *src = 0.f;
  }
}

}  // namespace

size_t wgrad_tc_smem(int KS) {
  const int HWp = kTW + KS - 1;
  const int HP = (kTH + KS - 1) * HWp;
  const int HPpad = HP | 1;
  return 128 + 2 * kMC * 4 + 2 * (size_t)(kMC / 8) * HPpad * 16 + 2 * (size_t)6 * 128 * 16;
}

bool wgrad_tc_supported(int KS, int stride) { return (KS == 1 || KS == 3) && stride == 1; }

void wgrad_tc_dims(int Cin, int Cout, int* ci_pad, int* co_pad) {
  *ci_pad = (Cin + kMC - 1) / kMC * kMC;
  *co_pad = (Cout + kNC - 1) / kNC * kNC;
}

int launch_wgrad_tc(const TcWgradArgs& t_in, cudaStream_t st) {
  TcWgradArgs t = t_in;
  {
    const char* e = getenv("PDES_WG_DBG");
    t.dbg = e ? atoi(e) : 0;
  }
  const WgradArgs& a = t.w;
  PDES_REQUIRE(wgrad_tc_supported(a.KS, a.stride) && !a.in_nchw && !a.dy_nchw, PDES_ERR_UNSUPPORTED,
               "wgrad_tc: stride-1 1x1/3x3 NHWC convolutions only");
  PDES_REQUIRE((a.ldx & 3) == 0 && ((uintptr_t)a.x & 15u) == 0, PDES_ERR_INVALID,
               "wgrad_tc: input must be 16-byte aligned with a pixel stride multiple of 4");
  const size_t smem = wgrad_tc_smem(a.KS);
  const int tiles = ((a.Wo + kTW - 1) / kTW) * ((a.Ho + kTH - 1) / kTH) * a.B;
  const int n_ci = (a.Cin + kMC - 1) / kMC, n_co = (a.Cout + kNC - 1) / kNC;
  // split-K over pixel tiles so that about two CTAs per SM are in flight over the whole launch
  int P = (2 * sm_count()) / (n_ci * n_co * kPasses);
  if (P < 1) P = 1;
  if (P > tiles) P = tiles;
  dim3 grid(P, n_ci * n_co, kPasses);
  if (a.KS == 3) {
    static size_t attr = 0;
    if (smem > attr) {
      PDES_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = smem;
    }
    wgrad_tc_kernel<3><<<grid, kThreads, smem, st>>>(t);
  } else {
    static size_t attr = 0;
    if (smem > attr) {
      PDES_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = smem;
    }
    wgrad_tc_kernel<1><<<grid, kThreads, smem, st>>>(t);
  }
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

int launch_wgrad_unpack(const TcWgradUnpack* dev_table, int n, int max_elems, cudaStream_t st) {
  if (n == 0) return PDES_OK;
  int bx = (max_elems + 255) / 256;
  if (bx > 64) bx = 64;
  if (bx < 1) bx = 1;
  wgrad_unpack_kernel<<<dim3(bx, n), 256, 0, st>>>(dev_table);
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

}  // namespace pdes

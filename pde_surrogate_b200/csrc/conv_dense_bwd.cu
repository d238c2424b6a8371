// conv_dense_bwd.cu — fused data gradient of a thin 3x3 layer (DenseNet layer, reference models/codec.py:65-69
// through autograd): ONE kernel instead of dY-split + dgrad.
//
//   dA[q, ci] = sum_{jy, jx, co} dY[q + (jy-1, jx-1), co] * W[co, ci, 2-jy, 2-jx]
//
//   * the GEMM-K operand is built inside the kernel: producer warps read the 16-channel gradient slice G and
//     the matching activation slice X of the layer's OUTPUT, apply the lazy BatchNorm-backward mean
//     corrections of every consumer (dY = G - c1 - xhat*c2, what act_split_kernel's `fix` mode did), scale by
//     the dynamic power of two, split into two fp16 pieces and write them - three horizontally shifted
//     copies ("dx in K") - into the tensor core's shared-memory image [k-octet (jx, co octet)][pixel][16 B]
//     of a (TR+2)-row tile; the vertical taps are descriptor offsets of W pixels.  They also emit the dY
//     planes the weight-gradient kernel reads.
//   * the packed filter ([jy][k-octet][piece][n = ci][8], <= 120 KB) is loaded once per CTA and stays resident.
//   * the BatchNorm-backward epilogue (ReLU mask, sum dZ, sum dZ*xhat, gradient accumulation into the block's
//     gradient buffer) no longer waits on global loads: the fp32 activations of the tile arrive through a
//     TMA ring (32-channel boxes, 128-byte swizzle) issued far ahead of the MMAs;
//   * and the gradient accumulation  G[pixel, ci] += scale * dZ  no longer goes through L2 atomics (measured:
//     ~2.5k clk per 32-channel stage, the kernel's limiter): the matching G box rides in the same ring stage,
//     is updated in shared memory and goes back with one TMA store per stage.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include "conv.cuh"
#include "conv_tc.cuh"
#include "tc_common.cuh"

namespace pdes {
namespace {

using namespace tc;

constexpr int kEpiWarps = 8;                 // warps 0..7: quarter = w & 3, column half = w >> 2
constexpr int kMmaWarp = 8;
constexpr int kTmaWarp = 9;
constexpr int kProdWarp0 = 10, kProdWarps = 4;
constexpr int kThreads = 32 * (kProdWarp0 + kProdWarps);   // 448
constexpr int kTS = 2;
constexpr int kMaxX = 3;
constexpr int kKO = 6;                       // k-octets per vertical tap: (jx, co octet)

__device__ __forceinline__ uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                            uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "setp.ne.b32 p, %6, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v) {
  return __half22float2(*reinterpret_cast<const __half2*>(&v));
}
__device__ __forceinline__ uint32_t pack_bf2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(smem_dst)),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, int c0, int c1, int c2, const void* smem_src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(tm), "r"(c0),
               "r"(c1), "r"(c2), "r"(smem_u32(smem_src))
               : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

struct BwdPlan {
  int XST, AST, ngroups;
  size_t hdr, smem;
};

__host__ __device__ inline size_t bwd_hdr_bytes(int N) {
  // barriers (256) | ep_s 4*N | red_s 4 quarters * N * 2 | fix_s 4*16 | fixd 32 doubles  (floats), rounded to 1024
  const size_t b = 256 + sizeof(float) * ((size_t)4 * N + (size_t)8 * N + 64 + 64);
  return (b + 1023) & ~(size_t)1023;
}

__global__ void __launch_bounds__(kThreads, 1)
conv_dense_bwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                      DenseBwdArgs a, int XST, int AST, int ngroups) {
  const int W = a.W, H = a.H, N = a.N;
  const int TR = 128 / W;
  const int HP = (TR + 2) * W;
  const int tiles_per_img = (H + TR - 1) / TR;
  const int n_tiles = tiles_per_img * a.B;
  const int wsh = W == 32 ? 5 : (W == 16 ? 4 : 3);
  const uint32_t a_piece_bytes = (uint32_t)kKO * HP * 16u;
  const uint32_t a_stage_bytes = 2u * a_piece_bytes;
  const uint32_t b_jy_bytes = (uint32_t)kKO * 2u * N * 16u;     // [k-octet][piece][n][16 B]
  const uint32_t x_plane_bytes = 128u * 128u;                  // [pixel][32 fp32]
  const uint32_t x_stage_bytes = 2u * x_plane_bytes;           // the activations and, behind them, the G box
  const int n_xst = (N + 31) >> 5;                             // 32-channel activation stages per tile

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem);   // [2]
  uint64_t* a_empty = a_full + 2;                         // [2]
  uint64_t* b_full = a_empty + 2;                         // [3]
  uint64_t* acc_full = b_full + 3;                        // [kTS]
  uint64_t* acc_empty = acc_full + kTS;                   // [kTS]
  uint64_t* x_full = acc_empty + kTS;                     // [kMaxX]
  uint64_t* x_empty = x_full + kMaxX;                     // [kMaxX]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_empty + kMaxX);
  float* ep_s = reinterpret_cast<float*>(smem + 256);     // [N][4]: scale, shift, invstd, -mean*invstd
  float* red_s = ep_s + 4 * N;                            // [4 quarters][N][2]
  float* fix_s = red_s + 8 * N;                           // [4][16]: c1, c2, mean, invstd of the dY channels
  double* fixd = reinterpret_cast<double*>(fix_s + 64);   // [2][16]: sum over consumers of scale*bsum (N % 2 == 0: aligned)
  const size_t hdr = bwd_hdr_bytes(N);
  unsigned char* X_s = smem + hdr;
  unsigned char* A_s = X_s + (size_t)XST * x_stage_bytes;
  unsigned char* B_s = A_s + (size_t)AST * a_stage_bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#define DBG(slot) do { if (a.dbg != nullptr && blockIdx.x < 4 && (slot) < 64) a.dbg[(size_t)blockIdx.x * 64 + (slot)] = clock64(); } while (0)
  if (threadIdx.x == 0) DBG(0);
  const uint32_t ts_cols = (uint32_t)(ngroups * N);
  uint32_t tmem_cols = 32;
  while (tmem_cols < ts_cols * kTS) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], kProdWarps);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 3; ++i) mbar_init(&b_full[i], 1);
    for (int i = 0; i < kTS; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], kEpiWarps);
    }
    for (int i = 0; i < kMaxX; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], kEpiWarps);
    }
    fence_mbar_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmG) : "memory");
    // The packed filter was written by the pack kernel at the head of the step, several launches ago.  Under
    // programmatic dependent launch this kernel starts once the PREVIOUS kernel has passed its own
    // griddepcontrol.wait, i.e. once everything before the previous kernel has completed: the resident
    // filter can be fetched while the previous kernel drains.  (b_early = 0: the unit-test entry point,
    // whose pack kernel is the immediately preceding launch.)
    if (a.b_early && (int)blockIdx.x < n_tiles) {
      for (int jy = 0; jy < 3; ++jy) {
        mbar_arrive_expect_tx(&b_full[jy], b_jy_bytes);
        tma_load_1d(B_s + (size_t)jy * b_jy_bytes,
                    reinterpret_cast<const unsigned char*>(a.wpk) + (size_t)jy * b_jy_bytes, b_jy_bytes, &b_full[jy]);
      }
    }
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
    fixd[lane] = 0.0;
  }
  tc_fence_before();
  __syncthreads();   // barriers, the TMEM base and the zeroed fix accumulators are visible before anybody uses them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();
  if (threadIdx.x == 0) DBG(1);
  // dynamic power-of-two scale of the gradient pieces (same rule as act_split_kernel)
  int dyn_e = 0;
  if (a.dyn_max != nullptr) {
    const unsigned mx = *a.dyn_max;
    dyn_e = mx == 0u ? 0 : kDyTargetLog2 - ((int)((mx >> 23) & 0xffu) - 127);
    for (int c = 1; c < a.fx.n_cons; c <<= 1) --dyn_e;
    dyn_e = dyn_e < -100 ? -100 : (dyn_e > 100 ? 100 : dyn_e);
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.dyn_inv != nullptr)
      *a.dyn_inv = __uint_as_float((uint32_t)(127 - dyn_e) << 23);
  }
  const float dmul = __uint_as_float((uint32_t)(dyn_e + 127) << 23);
  const float dinv = __uint_as_float((uint32_t)(127 - dyn_e) << 23);
  // ===== dY producers: loads of one tile (issued for the first tile before the prologue's constants exist) =====
  constexpr int kIt = 3, kSpStep = (kProdWarps * 32) >> 1;
  float vv[kIt][8], xx[kIt][8];
  const int pt = (int)threadIdx.x - kProdWarp0 * 32;   // 0..127 in the producer warps
  const int oct = pt & 1;                              // channel octet of the 16-channel slice
  const bool fix = a.fx.n_cons > 0;
  const bool vec = (a.fx.ldG & 3) == 0 && (a.fx.ldX & 3) == 0 && (reinterpret_cast<uintptr_t>(a.fx.G) & 15u) == 0 &&
                   (reinterpret_cast<uintptr_t>(a.fx.X) & 15u) == 0 && oct * 8 + 7 < a.Cout;
  auto load_dy = [&](int tile) {
    const FixDyArgs& f = a.fx;
    const int b = tile / tiles_per_img, r0 = (tile - b * tiles_per_img) * TR;
      // HP <= 192 source pixels over 64 thread pairs: at most kIt = 3 per thread.  All global loads of the
      // tile are issued before the first conversion (one memory latency per tile instead of three).
#pragma unroll
      for (int it = 0; it < kIt; ++it) {
        const int sp = (pt >> 1) + it * kSpStep;
        const int prow = sp >> wsh, col = sp & (W - 1);
        const int row = r0 - 1 + prow;
#pragma unroll
        for (int k = 0; k < 8; ++k) vv[it][k] = xx[it][k] = 0.f;
        if (sp < HP && row >= 0 && row < H && oct * 8 < a.Cout) {
          const size_t pix = ((size_t)b * H + row) * W + col;
          const float* gp = f.G + pix * f.ldG + oct * 8;
          const float* xp = f.X + pix * f.ldX + oct * 8;
          if (vec) {
            const float4 g0 = *reinterpret_cast<const float4*>(gp), g1 = *reinterpret_cast<const float4*>(gp + 4);
            vv[it][0] = g0.x; vv[it][1] = g0.y; vv[it][2] = g0.z; vv[it][3] = g0.w;
            vv[it][4] = g1.x; vv[it][5] = g1.y; vv[it][6] = g1.z; vv[it][7] = g1.w;
            if (fix) {
              const float4 x0 = __ldg(reinterpret_cast<const float4*>(xp)), x1 = __ldg(reinterpret_cast<const float4*>(xp + 4));
              xx[it][0] = x0.x; xx[it][1] = x0.y; xx[it][2] = x0.z; xx[it][3] = x0.w;
              xx[it][4] = x1.x; xx[it][5] = x1.y; xx[it][6] = x1.z; xx[it][7] = x1.w;
            }
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const bool in = oct * 8 + k < a.Cout;
              vv[it][k] = in ? gp[k] : 0.f;
              xx[it][k] = (in && fix) ? xp[k] : 0.f;
            }
          }
        }
      }
  };
  if (warp >= kProdWarp0 && (int)blockIdx.x < n_tiles) load_dy((int)blockIdx.x);
  if (warp < kEpiWarps) {
    const int tt = threadIdx.x;
    for (int n = tt; n < N; n += kEpiWarps * 32) {
      float s = 0.f, h = 0.f, m = 0.f, is = 0.f;
      if (n < a.Cin) bn_consts_tc(a.fbn, n, s, h, m, is);
      *reinterpret_cast<float4*>(ep_s + 4 * n) = make_float4(s, h, is, -m * is);
    }
    for (int i = tt; i < 8 * N; i += kEpiWarps * 32) red_s[i] = 0.f;
  } else {
    // lazy BatchNorm-backward corrections of the dY channels: c1 = sum_l scale_l * sum dZ_l / count,
    // c2 = sum_l scale_l * sum dZ_l xhat / count over the consumers l.  One thread per (consumer, channel):
    // every global load of the prologue is in flight at once (one memory latency instead of n_cons)
    const FixDyArgs& f = a.fx;
    const int t0 = threadIdx.x - kMmaWarp * 32, nt = kThreads - kMmaWarp * 32;
    for (int idx = t0; idx < f.n_cons * 16; idx += nt) {
      const int l = idx >> 4, c = idx & 15;
      if (c < a.Cout) {
        const double m = f.sum[c] * f.inv_count;
        double var = f.sumsq[c] * f.inv_count - m * m;
        const float gam = f.cons_gamma[l][c];
        const double b1 = f.cons_bsum[l][c], b2 = f.cons_bsum[l][f.cons_C[l] + c];
        if (var < 0.0) var = 0.0;
        const double isd = 1.0 / sqrt(var + (double)f.eps);
        const double sc = (double)gam * (double)(float)isd;
        atomicAdd(&fixd[c], sc * b1);
        atomicAdd(&fixd[16 + c], sc * b2);
        if (l == 0) {
          fix_s[32 + c] = (float)m;
          fix_s[48 + c] = (float)isd;
        }
      }
    }
  }
  // Two independent prologues: the epilogue warps only need their own BatchNorm constants, the producers (and
  // the MMA / TMA warps that helped computing them) only the dY corrections - the dY conversion of the first tile
  // does not wait for the epilogue's constants.
  if (warp < kEpiWarps)
    named_bar_sync(1, kEpiWarps * 32);
  else
    named_bar_sync(2, kThreads - kEpiWarps * 32);
  if (threadIdx.x == kEpiWarps * 32) DBG(2);

  if (warp >= kProdWarp0) {
    // ===== dY producers: G, X slices -> corrected, scaled, split -> three shifted copies in A_s =====
    const FixDyArgs& f = a.fx;
    const int CpB = (a.Cout + 7) & ~7, octB = CpB >> 3;
    const size_t planeB_elems = (size_t)a.B * H * W * CpB;
    float c1[8], c2[8], mn[8], isd[8];  // (filled below)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const bool on = fix && (oct * 8 + k < a.Cout);
      c1[k] = on ? (float)(fixd[oct * 8 + k] * f.inv_count) : 0.f;
      c2[k] = on ? (float)(fixd[16 + oct * 8 + k] * f.inv_count) : 0.f;
      mn[k] = on ? fix_s[32 + oct * 8 + k] : 0.f;
      isd[k] = on ? fix_s[48 + oct * 8 + k] : 0.f;
    }
    int t_it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t_it) {
      const int b = tile / tiles_per_img, r0 = (tile - b * tiles_per_img) * TR;
      const int s = t_it % AST;
      unsigned char* st = A_s + (size_t)s * a_stage_bytes;
      if (t_it > 0) load_dy(tile);
      if (lane == 0) mbar_wait(&a_empty[s], (uint32_t)(((t_it / AST) & 1) ^ 1));   // loads above are in flight
      __syncwarp();
#pragma unroll
      for (int it = 0; it < kIt; ++it) {
        const int sp = (pt >> 1) + it * kSpStep;
        if (sp >= HP) break;
        const int prow = sp >> wsh, col = sp & (W - 1);
        const int row = r0 - 1 + prow;
        uint4 h1 = make_uint4(0u, 0u, 0u, 0u), h2 = h1;
        if (row >= 0 && row < H && oct * 8 < a.Cout) {
          float* v = vv[it];
          const float* xv = xx[it];
          if (fix) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float xh = (xv[k] - mn[k]) * isd[k];
              v[k] = v[k] - c1[k] - xh * c2[k];
            }
          }
          uint32_t p1[4], p2[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float t0 = v[2 * k] * dmul, t1 = v[2 * k + 1] * dmul;
            if (a.lowp == LOWP_BF16) {
              p1[k] = pack_bf2(t0, t1);
              p2[k] = 0u;
            } else {
              p1[k] = pack_h2(t0, t1);
              const float2 fl = unpack_h2(p1[k]);
              p2[k] = a.lowp ? 0u : pack_h2(t0 - fl.x, t1 - fl.y);
            }
          }
          h1 = make_uint4(p1[0], p1[1], p1[2], p1[3]);
          h2 = make_uint4(p2[0], p2[1], p2[2], p2[3]);
          if (a.planesB != nullptr && prow >= 1 && prow <= TR && oct < octB) {
            op16* pd = a.planesB + ((((size_t)b * H + row) * octB + oct) * W + col) * 8;
            *reinterpret_cast<uint4*>(pd) = h1;
            *reinterpret_cast<uint4*>(pd + planeB_elems) = h2;
          }
        }
        // A[(row', col'), (jx, oct)] = dY[(row', col' + jx - 1)]: this source feeds col' = col - (jx - 1)
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        unsigned char* rowb = st + (size_t)(prow << wsh) * 16;
#pragma unroll
        for (int jx = 0; jx < 3; ++jx) {
          const int dc = col - (jx - 1);
          unsigned char* kb = rowb + (size_t)(jx * 2 + oct) * HP * 16;
          if (dc >= 0 && dc < W) {
            *reinterpret_cast<uint4*>(kb + (size_t)dc * 16) = h1;
            *reinterpret_cast<uint4*>(kb + (size_t)dc * 16 + a_piece_bytes) = h2;
          }
          // image border: the tap that would read column -1 / W sees the convolution's zero padding
          if ((jx == 0 && col == 0) || (jx == 2 && col == W - 1)) {
            *reinterpret_cast<uint4*>(kb + (size_t)col * 16) = z;
            *reinterpret_cast<uint4*>(kb + (size_t)col * 16 + a_piece_bytes) = z;
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[s]);
      if (threadIdx.x == kProdWarp0 * 32) DBG(3 + t_it);      // A operand of tile t_it published
    }
  } else if (warp == kTmaWarp) {
    // ===== activation tiles of the BatchNorm-backward epilogue: (32 channels, 128 pixels) boxes =====
    if (lane == 0) {
      // ring stage = [X box | G box].  Before a stage is refilled, the G box the epilogue warps updated in it
      // (iteration q - XST) is written back with one TMA store (out-of-range channels / pixels are clipped).
      int q_it = 0;
      int pend_c[kMaxX], pend_p[kMaxX], pend_b[kMaxX];
      for (int i = 0; i < kMaxX; ++i) pend_c[i] = -1;
      auto write_back = [&](int s) {
        if (pend_c[s] >= 0) {
          tma_store_3d(&tmG, pend_c[s], pend_p[s], pend_b[s], X_s + (size_t)s * x_stage_bytes + x_plane_bytes);
          tma_store_commit();
          tma_store_wait_read<0>();   // the store has read the stage: it may be refilled
          pend_c[s] = -1;
        }
      };
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_img, r0 = (tile - b * tiles_per_img) * TR;
        for (int j = 0; j < n_xst; ++j, ++q_it) {
          const int s = q_it % XST;
          mbar_wait(&x_empty[s], (uint32_t)(((q_it / XST) & 1) ^ 1));
          write_back(s);
          unsigned char* st = X_s + (size_t)s * x_stage_bytes;
          mbar_arrive_expect_tx(&x_full[s], a.g_accum ? x_stage_bytes : x_plane_bytes);
          tma_load_3d(st, &tmX, j * 32, r0 * W, b, &x_full[s]);
          if (a.g_accum) tma_load_3d(st + x_plane_bytes, &tmG, j * 32, r0 * W, b, &x_full[s]);
          pend_c[s] = j * 32;
          pend_p[s] = r0 * W;
          pend_b[s] = b;
        }
      }
      // drain: the last XST stages
      for (int k = 0; k < XST; ++k, ++q_it) {
        const int s = q_it % XST;
        if (pend_c[s] < 0) continue;
        mbar_wait(&x_empty[s], (uint32_t)(((q_it / XST) & 1) ^ 1));
        write_back(s);
      }
      tma_store_wait_all<0>();
    }
  } else if (warp == kMmaWarp) {
    if (!a.b_early && lane == 0 && (int)blockIdx.x < n_tiles) {
      for (int jy = 0; jy < 3; ++jy) {
        mbar_arrive_expect_tx(&b_full[jy], b_jy_bytes);
        tma_load_1d(B_s + (size_t)jy * b_jy_bytes,
                    reinterpret_cast<const unsigned char*>(a.wpk) + (size_t)jy * b_jy_bytes, b_jy_bytes, &b_full[jy]);
      }
    }
    const uint32_t idesc2 = make_idesc_f16(128, 2 * N) | idesc_fmt_bits(a.lowp);
    const uint32_t idesc1 = make_idesc_f16(128, N) | idesc_fmt_bits(a.lowp);
    const uint32_t lbo_a = (uint32_t)HP * 16u, sbo_a = 128u;
    const uint32_t lbo_b = 2u * N * 16u, sbo_b = 128u;
    const uint32_t kstep_a = (2u * lbo_a) >> 4, kstep_b = (2u * lbo_b) >> 4;
    const uint32_t a_piece_u = a_piece_bytes >> 4, b_jy_u = b_jy_bytes >> 4;
    const uint32_t n_u = (uint32_t)N;   // N rows * 16 B in 16-byte units: offset of the second filter piece
    int ts = 0, t_it = 0;
    uint32_t pt = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t_it) {
      const int sa = t_it % AST;
      mbar_wait(&acc_empty[ts], pt ^ 1u);
      mbar_wait(&a_full[sa], (uint32_t)((t_it / AST) & 1));
      if (t_it == 0)
        for (int jy = 0; jy < 3; ++jy) mbar_wait(&b_full[jy], 0);
      tc_fence_after();
      if (lane == 0) DBG(8 + t_it);                           // MMA issue of tile t_it starts
      const uint32_t d0 = tmem_base + (uint32_t)ts * ts_cols;
      const uint64_t ad0 = make_desc(smem_u32(A_s + (size_t)sa * a_stage_bytes), lbo_a, sbo_a);
      const uint64_t bd0 = make_desc(smem_u32(B_s), lbo_b, sbo_b);
      const uint32_t a_lo0 = (uint32_t)ad0, a_hi = (uint32_t)(ad0 >> 32);
      const uint32_t b_lo0 = (uint32_t)bd0, b_hi = (uint32_t)(bd0 >> 32);
      if (elect_one()) {
#pragma unroll
        for (int jy = 0; jy < 3; ++jy) {
#pragma unroll
          for (int k16 = 0; k16 < 3; ++k16) {
            const uint32_t a_k = a_lo0 + (uint32_t)(jy * W) + (uint32_t)k16 * kstep_a;
            const uint32_t b_k = b_lo0 + (uint32_t)jy * b_jy_u + (uint32_t)k16 * kstep_b;
            const uint32_t first = (jy == 0 && k16 == 0) ? 0u : 1u;
            if (a.lowp) {
              umma_f16_w(d0, a_k, a_hi, b_k, b_hi, idesc1, first);                           // one piece: a1 x w1
            } else if (ngroups == 2) {
              umma_f16_w(d0, a_k, a_hi, b_k, b_hi, idesc2, first);                           // a1 x [w1|w2] -> G0 | G1
              umma_f16_w(d0 + (uint32_t)N, a_k + a_piece_u, a_hi, b_k, b_hi, idesc1, 1u);    // a2 x w1 -> G1
            } else {
              umma_f16_w(d0, a_k, a_hi, b_k, b_hi, idesc1, first);                           // a1 x w1
              umma_f16_w(d0, a_k, a_hi, b_k + n_u, b_hi, idesc1, 1u);                        // a1 x w2
              umma_f16_w(d0, a_k + a_piece_u, a_hi, b_k, b_hi, idesc1, 1u);                  // a2 x w1
            }
          }
        }
        umma_commit(&a_empty[sa]);
        umma_commit(&acc_full[ts]);
      }
      __syncwarp();
      if (++ts == kTS) {
        ts = 0;
        pt ^= 1u;
      }
    }
    griddep_launch();
  } else {
    // ===== epilogue: ReLU mask + BatchNorm backward + gradient accumulation =====
    const int ew = warp, quarter = warp & 3, half = warp >> 2;
    const int m = quarter * 32 + lane;                 // tile pixel = TMEM lane
    const int prow = m >> wsh, px = m & (W - 1);
    const int my_col = colsum16_col(lane);
    const float osc = a.out_scale * dinv;
    float gmx = 0.f;
    int t_it = 0, q_it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t_it) {
      const int ts = t_it % kTS;
      const int b = tile / tiles_per_img, r0 = (tile - b * tiles_per_img) * TR;
      const int row = r0 + prow;
      const bool valid = row < H;
      mbar_wait(&acc_full[ts], (uint32_t)((t_it / kTS) & 1));
      tc_fence_after();
      if (threadIdx.x == 0) DBG(12 + 2 * t_it);               // accumulator of tile t_it complete
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)ts * ts_cols;
      for (int j = 0; j < n_xst; ++j, ++q_it) {
        const int xs = q_it % XST;
        mbar_wait(&x_full[xs], (uint32_t)((q_it / XST) & 1));
        if (threadIdx.x == 0) DBG(20 + q_it);                 // X stage q_it available to the epilogue
        const int n0 = j * 32 + half * 16;
        if (n0 < N) {
          // activations of this pixel: 16 channels = 4 swizzled 16-byte chunks of row m
          const unsigned char* xrow = X_s + (size_t)xs * x_stage_bytes + (size_t)m * 128;
          float xv[16];
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const float4 t = *reinterpret_cast<const float4*>(xrow + (((half * 4 + c4) ^ (m & 7)) << 4));
            xv[4 * c4] = t.x; xv[4 * c4 + 1] = t.y; xv[4 * c4 + 2] = t.z; xv[4 * c4 + 3] = t.w;
          }
          float v[16];
          if (ngroups == 2 && !a.lowp) {
            float w1[16];
            tmem_ld16(taddr + (uint32_t)(N + n0), v);     // cross terms first (small), then the leading ones
            tmem_ld16(taddr + (uint32_t)n0, w1);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += w1[i];
          } else {
            tmem_ld16(taddr + (uint32_t)n0, v);
          }
          float s1[16], s2[16], o[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float4 c = *reinterpret_cast<const float4*>(ep_s + 4 * (n0 + i));
            const float z = fmaf(xv[i], c.x, c.y);
            const float dz = (valid && z > 0.f) ? v[i] * osc : 0.f;
            const float xh = fmaf(xv[i], c.z, c.w);
            s1[i] = dz;
            s2[i] = dz * xh;
            o[i] = c.x * dz;
            gmx = fmaxf(gmx, fabsf(o[i]));
          }
          {
            // G box of this stage: row m, the four 16-byte chunks of this warp half (same swizzle as X)
            unsigned char* grow = X_s + (size_t)xs * x_stage_bytes + x_plane_bytes + (size_t)m * 128;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              float4* gp = reinterpret_cast<float4*>(grow + (((half * 4 + c4) ^ (m & 7)) << 4));
              float4 t = make_float4(o[4 * c4], o[4 * c4 + 1], o[4 * c4 + 2], o[4 * c4 + 3]);
              if (a.g_accum) {
                const float4 g = *gp;
                t.x += g.x; t.y += g.y; t.z += g.z; t.w += g.w;
              }
              *gp = t;
            }
          }
          const float u = colsum16(s1, lane), w = colsum16(s2, lane);
          if ((lane & 1) == 0) {
            float* r = red_s + ((size_t)quarter * N + n0 + my_col) * 2;
            r[0] += u;
            r[1] += w;
          }
        }
        fence_proxy_async_smem();   // the updated G box -> visible to the TMA store
        __syncwarp();
        if (lane == 0) mbar_arrive(&x_empty[xs]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[ts]);
      if (threadIdx.x == 0) DBG(13 + 2 * t_it);               // epilogue of tile t_it done
    }
    if (a.gmax != nullptr) {
      const unsigned mm = __reduce_max_sync(0xffffffffu, __float_as_uint(gmx));
      if (lane == 0 && mm != 0u) atomicMax(a.gmax, mm);
    }
    named_bar_sync(1, kEpiWarps * 32);
    for (int n = threadIdx.x; n < a.Cin; n += kEpiWarps * 32) {
      double u = 0.0, w = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        u += (double)red_s[((size_t)k * N + n) * 2 + 0];
        w += (double)red_s[((size_t)k * N + n) * 2 + 1];
      }
      atomicAdd(a.bsum + n, u);
      atomicAdd(a.bsum + a.Cin + n, w);
    }
    (void)ew;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) DBG(40);
#undef DBG
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

bool bwd_plan(int W, int N, BwdPlan* p) {
  const int TR = 128 / W, HP = (TR + 2) * W;
  const size_t a_stage = (size_t)2 * kKO * HP * 16, b_bytes = (size_t)3 * kKO * 2 * N * 16;
  p->ngroups = 4 * N <= 512 ? 2 : 1;
  p->hdr = bwd_hdr_bytes(N);
  for (int ast = 2; ast >= 1; --ast)
    for (int xst = kMaxX; xst >= 2; --xst) {
      const size_t s = 1024 + p->hdr + (size_t)xst * 2 * 128 * 128 + ast * a_stage + b_bytes;
      if (s <= 227 * 1024) {
        p->XST = xst;
        p->AST = ast;
        p->smem = s;
        return true;
      }
    }
  return false;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn get_encode_b() {
  static EncodeFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeFn>(p);
  return fn;
}

}  // namespace

bool dense_bwd_supported(int KS, int stride, int pad, int up, int Cin, int Cout, int H, int W) {
  if (KS != 3 || stride != 1 || pad != 1 || up) return false;
  if (!(W == 8 || W == 16 || W == 32) || H < 1) return false;
  if (Cout < 1 || Cout > 16) return false;
  const int N = (Cin + 15) / 16 * 16;
  if (N < 16 || N > 256) return false;
  BwdPlan p;
  return bwd_plan(W, N, &p);
}

size_t dense_bwd_pack_elems(int N) { return (size_t)3 * kKO * 2 * N * 8; }

int launch_conv_dense_bwd(const DenseBwdArgs& a, cudaStream_t st) {
  PDES_REQUIRE(dense_bwd_supported(3, 1, 1, 0, a.Cin, a.Cout, a.H, a.W), PDES_ERR_UNSUPPORTED,
               "conv_dense_bwd: unsupported shape (Cin %d, Cout %d, %dx%d)", a.Cin, a.Cout, a.H, a.W);
  PDES_REQUIRE(a.N == (a.Cin + 15) / 16 * 16, PDES_ERR_INVALID, "conv_dense_bwd: N must be Cin rounded up to 16");
  PDES_REQUIRE(((a.ldx | a.ldG) & 3) == 0 && ((uintptr_t)a.x & 15u) == 0 && ((uintptr_t)a.G & 15u) == 0,
               PDES_ERR_INVALID, "conv_dense_bwd: activation / gradient buffers misaligned");
  BwdPlan p;
  bwd_plan(a.W, a.N, &p);
  PDES_ENSURE_SMEM(conv_dense_bwd_kernel, p.smem);
  const int TR = 128 / a.W;
  const int tiles = ((a.H + TR - 1) / TR) * a.B;
  int grid = sm_count();
  if (grid > tiles) grid = tiles;
  EncodeFn enc = get_encode_b();
  PDES_REQUIRE(enc != nullptr, PDES_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  CUtensorMap tm;
  {
    const cuuint64_t gdim[3] = {(cuuint64_t)a.Cin, (cuuint64_t)a.H * a.W, (cuuint64_t)a.B};
    const cuuint64_t gstr[2] = {(cuuint64_t)a.ldx * 4, (cuuint64_t)a.H * a.W * a.ldx * 4};
    const cuuint32_t box[3] = {32, 128, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(a.x), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PDES_REQUIRE(r == CUDA_SUCCESS, PDES_ERR_CUDA, "cuTensorMapEncodeTiled (conv_dense_bwd) failed with code %d", (int)r);
  }
  CUtensorMap tg;
  {
    const cuuint64_t gdim[3] = {(cuuint64_t)a.Cin, (cuuint64_t)a.H * a.W, (cuuint64_t)a.B};
    const cuuint64_t gstr[2] = {(cuuint64_t)a.ldG * 4, (cuuint64_t)a.H * a.W * a.ldG * 4};
    const cuuint32_t box[3] = {32, 128, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(&tg, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, a.G, gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PDES_REQUIRE(r == CUDA_SUCCESS, PDES_ERR_CUDA, "cuTensorMapEncodeTiled (conv_dense_bwd, G) failed with code %d", (int)r);
  }
  PDES_CUDA(launch_pdl(conv_dense_bwd_kernel, dim3(grid), dim3(kThreads), p.smem, st, tm, tg, a, p.XST, p.AST, p.ngroups));
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

}  // namespace pdes

// wgrad_tc.cu — tcgen05 weight-gradient kernel fed by TMA tensor loads.
//
//   dW[tap][ci][co] = sum over pixels p of  a[p + tap, ci] * dY[p, co]
//
// GEMM view per filter tap: D[M = 128 input channels, N = 16 output channels] += A^T * B with the
// reduction (GEMM-K) over PIXELS.  Both operands are pixel-major ([pixel][channel], what NHWC
// gives for free), i.e. MN-major UMMA operands: core matrix = 8 pixels x 16 B (8 fp16 channels).
// One tcgen05.mma (K = 16) consumes two 8-pixel tile rows; the nine taps read the SAME staged halo
// tile through shifted descriptors and own nine accumulators in TMEM.
//
// Operands are NOT transformed inside this kernel.  A small elementwise pre-pass
// (act_split_kernel) writes the BN+ReLU'd (and nearest-upsampled) activations — and the
// corrected dY slice — once per layer as two fp16 planes of the power-of-two scaled values
// (conv_tc.cuh).  The kernel then streams 4-D TMA boxes (cp.async.bulk.tensor, zero fill
// outside the image = the convolution's padding) whose shared-memory image IS the UMMA layout:
// planes are stored [b][y][octet][x][8]; box (10 x * 8 ch, 18 y, 16 octets, 1) -> [octet][pixel][16 B],
// each box row 160 contiguous bytes.
//
// Precision: tcgen05 only transposes 16-bit operands (kind::tf32 with MN-major operands returns
// zeros on sm_100a — measured), hence 16-bit pieces: two fp16 pieces per (power-of-two scaled)
// operand and the three products of weight >= 2^-11 (conv_tc.cuh):
//   pass 0 CTAs stream a1 and multiply by [d1|d2]   (one N'=32 instruction per tap)
//   pass 1 CTAs stream a2 and multiply by  d1       (N =16)
// Every product lands in its own TMEM columns and the epilogue adds them in fp32.  CTAs split the
// pixel range (split-K) and finish with coalesced vector reductions (red.global.add.v4.f32) into a
// [tap][ci][co] staging buffer that a small kernel folds into the OIHW gradient.
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

constexpr int kThreads = 192;
constexpr int kTH = 16, kTW = 8;
constexpr int kMC = 128;  // input channels per CTA (GEMM M)
constexpr int kPasses = kPieces;
constexpr int kIm2colMaxN = 256;  // dy_im2col: padded column count (taps * Cout rounded up to 8)

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// two fp16 pieces of an (already scaled) value; saturating conversions keep overflow finite
__device__ __forceinline__ uint32_t f2h_sat(float x) {
  unsigned short h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
  return (uint32_t)h;
}
__device__ __forceinline__ void fp16_split2(float x, uint32_t& p0, uint32_t& p1) {
  p0 = f2h_sat(x);
  p1 = f2h_sat(x - __half2float(__ushort_as_half((unsigned short)p0)));
}
// the pieces of a value in the given precision mode (conv_tc.cuh): two fp16 pieces, or one fp16 / bf16 piece
__device__ __forceinline__ void split_mode(float x, int lowp, uint32_t& p0, uint32_t& p1) {
  if (lowp == LOWP_BF16) {
    p0 = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(x));
    p1 = 0u;
  } else if (lowp == LOWP_FP16) {
    p0 = f2h_sat(x);
    p1 = 0u;
  } else {
    fp16_split2(x, p0, p1);
  }
}

// instruction descriptor: D fp32, A/B fp16 (format 0), both MN-major
__device__ __forceinline__ uint32_t make_idesc_f16_mn(int M, int N) {
  return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
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
__device__ __forceinline__ bool elect_one_w() {
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
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2,
                                            int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, "
      "%5}], [%6];" ::"r"(smem_u32(smem_dst)),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
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

// ---------------------------------------------------------------------------------------
// pre-pass: planes[piece][b][y][octet][x][8] (fp16) = split2( scale * (pro ? relu(x*bn_scale+bn_shift) : x) ),
// optionally nearest-upsampled / zero-inserted x2.  Rows of one channel octet are x-contiguous so that
// a TMA box row is (TW+K-1)*16 contiguous bytes.
//
// Work item = one source image row (segment).  Phase 1 reads the NHWC row with consecutive threads on
// consecutive float4 of a pixel (coalesced: a pixel's channels are contiguous), applies the
// prologue / the lazy BatchNorm-backward correction, scales and splits, and drops the pieces into a
// shared-memory tile [piece][octet][x][16 B] (rows padded by 16 B against bank conflicts); phase 2
// streams every (piece, octet) row of the tile to global memory as x-contiguous 16-byte stores.
// ---------------------------------------------------------------------------------------
//
// Two kernels.  act_split_generic_kernel: any layout (planar input, unaligned rows), one work item per loop
// trip with direct global loads.  act_split_staged_kernel (the one the network uses): persistent CTAs; the
// NHWC rows of the NEXT work item are fetched with cp.async (zero-filled beyond C) into a double-buffered
// staging area while the current item is converted and stored, per-thread constants live in registers and
// the conversions are the packed f16x2 forms (the first version issued ~390 instructions per channel octet
// and ran issue-bound at ~30 % of the HBM roofline).
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2s(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf2s(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

__global__ void __launch_bounds__(256, 2) act_split_staged_kernel(ActSplitArgs a, int xs) {
  griddep_wait();
  griddep_launch();  // the consumer's CTAs may take their SM slots and run their prologue now
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int Cp = a.Cp;
  float* sc_s = reinterpret_cast<float*>(sm_raw);  // fused dY correction: c1 -> sc_s, c2 -> sh_s
  float* sh_s = sc_s + Cp;
  float* mean_s = sh_s + Cp;
  float* is_s = mean_s + Cp;
  unsigned char* tile = sm_raw + 16 * (size_t)Cp;
  if (a.fix) {
    const FixDyArgs& f = a.fx;
    for (int c = threadIdx.x; c < Cp; c += blockDim.x) {
      float c1 = 0.f, c2 = 0.f, mean = 0.f, is = 0.f;
      if (c < a.C) {
        const double m = f.sum[c] * f.inv_count;
        double var = f.sumsq[c] * f.inv_count - m * m;
        if (var < 0.0) var = 0.0;
        const double isd = 1.0 / sqrt(var + (double)f.eps);
        double d1 = 0.0, d2 = 0.0;
        for (int l = 0; l < f.n_cons; ++l) {
          const double sc = (double)f.cons_gamma[l][c] * (double)(float)isd;
          d1 += sc * f.cons_bsum[l][c];
          d2 += sc * f.cons_bsum[l][f.cons_C[l] + c];
        }
        c1 = (float)(d1 * f.inv_count);
        c2 = (float)(d2 * f.inv_count);
        mean = (float)m;
        is = (float)isd;
      }
      sc_s[c] = c1;
      sh_s[c] = c2;
      mean_s[c] = mean;
      is_s[c] = is;
    }
  } else if (a.pro) {
    for (int c = threadIdx.x; c < Cp; c += blockDim.x) {
      float s = 0.f, h = 0.f;
      if (c < a.C) bn_consts_w(a.bn, c, s, h);
      sc_s[c] = s;
      sh_s[c] = h;
    }
  }
  float mul = a.scale;
  if (a.dyn_max != nullptr) {
    // same rule as the generic kernel: power of two that brings the running |gradient| maximum (times the
    // number of summed consumer contributions) to 2^kDyTargetLog2
    const unsigned m = *a.dyn_max;
    int e = m == 0u ? 0 : kDyTargetLog2 - ((int)((m >> 23) & 0xffu) - 127);
    if (a.fix)
      for (int c = 1; c < a.fx.n_cons; c <<= 1) --e;
    e = e < -100 ? -100 : (e > 100 ? 100 : e);
    mul = __uint_as_float((uint32_t)(e + 127) << 23);
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.dyn_inv = __uint_as_float((uint32_t)(127 - e) << 23);
  }
  const int oct = Cp >> 3;
  const int up = a.up;
  const int Hv = up ? 2 * a.Hs : a.Hs, Wv = up ? 2 * a.Ws : a.Ws;
  const size_t plane = (size_t)a.B * Hv * Wv * Cp;  // elements per piece plane
  const int XO = up ? 2 * xs : xs;                  // output pixels per segment
  const int rstride = XO * 16 + 16;                 // bytes per (piece, octet) row of the tile
  const int nseg = (a.Ws + xs - 1) / xs;
  const int n_work = a.B * a.Hs * nseg;
  // phase-1 thread grid: qp (power of two >= octets) threads along the channels of a pixel, 256/qp pixels
  // in flight; phase-2 thread grid: xp (power of two >= output pixels) threads along x
  int qsh = 0;
  while ((1 << qsh) < oct) ++qsh;
  const int tq = threadIdx.x & ((1 << qsh) - 1), tp = threadIdx.x >> qsh, pstep = 256 >> qsh;
  const bool qon = tq < oct;
  const int c = 8 * tq;
  // staging: [buffer][x | fx.X][pixel][Cp] fp32 behind the tile
  const size_t tile_bytes = (size_t)kPieces * oct * rstride;
  const uint32_t stage_img = (uint32_t)xs * Cp * 4u;
  const uint32_t stage_buf = stage_img * (a.fix ? 2u : 1u);
  unsigned char* stage = tile + ((tile_bytes + 15) & ~(size_t)15);
  // valid bytes of this thread's two 16-byte chunks (channels beyond C are zero-filled)
  int nb0 = (a.C - c) * 4, nb1 = (a.C - c - 4) * 4;
  nb0 = nb0 < 0 ? 0 : (nb0 > 16 ? 16 : nb0);
  nb1 = nb1 < 0 ? 0 : (nb1 > 16 ? 16 : nb1);
  auto issue = [&](int work, int buf) {
    if (!qon) return;
    int seg = 0, row = work;   // row = b * Hs + sy
    if (nseg > 1) {
      seg = work % nseg;
      row = work / nseg;
    }
    const int x0 = seg * xs;
    const int nx = (a.Ws - x0) < xs ? (a.Ws - x0) : xs;
    const size_t pix = (size_t)row * a.Ws + x0 + tp;
    const float* src = a.x + pix * a.ldx + c;
    const float* srx = a.fix ? a.fx.X + pix * a.fx.ldX + c : nullptr;
    unsigned char* d = stage + (size_t)buf * stage_buf + ((size_t)tp * Cp + c) * 4;
    const size_t sstep = (size_t)pstep * a.ldx, xstep = a.fix ? (size_t)pstep * a.fx.ldX : 0;
    const uint32_t dstep = (uint32_t)pstep * Cp * 4u;
    for (int px = tp; px < nx; px += pstep) {
      cp_async16_zfill(d, nb0 ? src : a.x, nb0);
      cp_async16_zfill(d + 16, nb1 ? src + 4 : a.x, nb1);
      if (a.fix) {
        cp_async16_zfill(d + stage_img, nb0 ? srx : a.fx.X, nb0);
        cp_async16_zfill(d + stage_img + 16, nb1 ? srx + 4 : a.fx.X, nb1);
        srx += xstep;
      }
      src += sstep;
      d += dstep;
    }
  };
  if ((int)blockIdx.x < n_work) issue(blockIdx.x, 0);
  cp_async_commit();
  __syncthreads();   // constants ready
  // per-thread constants of this thread's channel octet (zero beyond C: those lanes produce zeros).
  //   pro: v = max(0, v*ka + kb) with the piece scale folded in (a power of two: exact)
  //   fix: v = v - ka - ((x - kc) * kd) * kb
  float ka[8], kb[8], kc[8], kd[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const bool in = qon && c + k < a.C;
    ka[k] = kb[k] = kc[k] = kd[k] = 0.f;
    if (a.fix) {
      ka[k] = in ? sc_s[c + k] : 0.f;
      kb[k] = in ? sh_s[c + k] : 0.f;
      kc[k] = in ? mean_s[c + k] : 0.f;
      kd[k] = in ? is_s[c + k] : 0.f;
    } else if (a.pro) {
      ka[k] = in ? sc_s[c + k] * mul : 0.f;
      kb[k] = in ? sh_s[c + k] * mul : 0.f;
    }
  }
  const bool full8 = c + 7 < a.C;
  const bool hs_pow2 = (a.Hs & (a.Hs - 1)) == 0;
  int hs_sh = 0;
  while ((1 << hs_sh) < a.Hs) ++hs_sh;
  // phase-2 lane grid of a full-width segment (recomputed only for a shorter last segment)
  int xsh_full = 0;
  while ((1 << xsh_full) < (up ? 2 * xs : xs)) ++xsh_full;
  int buf = 0;
  for (int work = blockIdx.x; work < n_work; work += gridDim.x) {
    int seg = 0, row = work;
    if (nseg > 1) {
      seg = work % nseg;
      row = work / nseg;
    }
    // (image sizes are powers of two in every configuration of the scripts: shift / mask instead of a division)
    const int b = hs_pow2 ? (row >> hs_sh) : row / a.Hs, sy = row - b * a.Hs;
    const int x0 = seg * xs;
    const int nx = (a.Ws - x0) < xs ? (a.Ws - x0) : xs;
    if (work + (int)gridDim.x < n_work) issue(work + gridDim.x, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();   // this item's rows have landed (this thread's share; the barrier publishes the rest)
    __syncthreads();      // previous tile drained / staging visible
    // ---- phase 1: staged rows -> transform, split -> shared tile ------------------------------
    if (qon) {
      const unsigned char* sp = stage + (size_t)buf * stage_buf + ((size_t)tp * Cp + c) * 4;
      const uint32_t sstep = (uint32_t)pstep * Cp * 4u;
      float* gp = const_cast<float*>(a.x) + ((size_t)row * a.Ws + x0 + tp) * a.ldx + c;
      const size_t gstep = (size_t)pstep * a.ldx;
      unsigned char* t0 = tile + (size_t)tq * rstride + (size_t)(up ? 2 * tp : tp) * 16;
      const uint32_t tstep = (uint32_t)(up ? 2 * pstep : pstep) * 16u;
      const uint32_t poff = (uint32_t)oct * rstride;
      for (int px = tp; px < nx; px += pstep) {
        const float4 f0 = *reinterpret_cast<const float4*>(sp);
        const float4 f1 = *reinterpret_cast<const float4*>(sp + 16);
        float v[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
        if (a.fix) {
          // dY = G - c1 - xhat*c2 (lazy BatchNorm-backward mean corrections of every consumer), written
          // back in fp32 for the non-tensor-core consumers of the slice
          const float4 x0q = *reinterpret_cast<const float4*>(sp + stage_img);
          const float4 x1q = *reinterpret_cast<const float4*>(sp + stage_img + 16);
          const float xv[8] = {x0q.x, x0q.y, x0q.z, x0q.w, x1q.x, x1q.y, x1q.z, x1q.w};
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float xh = (xv[k] - kc[k]) * kd[k];
            v[k] = v[k] - ka[k] - xh * kb[k];
          }
          if (full8) {
            *reinterpret_cast<float4*>(gp) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(gp + 4) = make_float4(v[4], v[5], v[6], v[7]);
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (c + k < a.C) gp[k] = v[k];
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] *= mul;
        } else if (a.pro) {
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = fmaxf(0.f, fmaf(v[k], ka[k], kb[k]));
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] *= mul;
        }
        uint32_t p1[4], p2[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (a.lowp == LOWP_BF16) {
            p1[k] = pack_bf2s(v[2 * k], v[2 * k + 1]);
            p2[k] = 0u;
          } else {
            p1[k] = pack_h2s(v[2 * k], v[2 * k + 1]);
            const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&p1[k]));
            p2[k] = a.lowp ? 0u : pack_h2s(v[2 * k] - fl.x, v[2 * k + 1] - fl.y);
          }
        }
        const uint4 w0 = make_uint4(p1[0], p1[1], p1[2], p1[3]);
        const uint4 w1 = make_uint4(p2[0], p2[1], p2[2], p2[3]);
        unsigned char* t1 = t0 + poff;
        *reinterpret_cast<uint4*>(t0) = w0;
        *reinterpret_cast<uint4*>(t1) = w1;
        if (up == 1) {  // nearest x2: the same value at 2x and 2x+1
          *reinterpret_cast<uint4*>(t0 + 16) = w0;
          *reinterpret_cast<uint4*>(t1 + 16) = w1;
        } else if (up == 2) {  // zero insertion: odd positions are zero
          *reinterpret_cast<uint4*>(t0 + 16) = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(t1 + 16) = make_uint4(0u, 0u, 0u, 0u);
        }
        sp += sstep;
        gp += gstep;
        t0 += tstep;
      }
    }
    buf ^= 1;
    __syncthreads();
    // ---- phase 2: tile rows -> global, x-contiguous 16-byte stores ---------------------------
    const int nxo = up ? 2 * nx : nx, xo0 = up ? 2 * x0 : x0;
    int xsh = xsh_full;
    if (nx != xs) {
      xsh = 0;
      while ((1 << xsh) < nxo) ++xsh;
    }
    const int ox = threadIdx.x & ((1 << xsh) - 1);
    if (ox < nxo) {
      const int nrows = (a.lowp ? 1 : kPieces) * oct, rstep = 256 >> xsh;
      op16* dst0 = a.out + ((((size_t)b * Hv + (up ? 2 * sy : sy)) * oct) * Wv + xo0 + ox) * 8;
      const size_t qstride = (size_t)Wv * 8, rowo = (size_t)oct * Wv * 8;
      const unsigned char* tp2 = tile + (size_t)ox * 16;
      for (int r = threadIdx.x >> xsh; r < nrows; r += rstep) {  // r = piece * oct + q
        const int piece = r >= oct ? 1 : 0, q = r - piece * oct;
        const uint4 val = *reinterpret_cast<const uint4*>(tp2 + (size_t)r * rstride);
        op16* dst = dst0 + (piece ? plane : 0) + (size_t)q * qstride;
        *reinterpret_cast<uint4*>(dst) = val;
        if (up) {  // second output row: a copy (nearest) or zeros (zero insertion)
          *reinterpret_cast<uint4*>(dst + rowo) = up == 1 ? val : make_uint4(0u, 0u, 0u, 0u);
        }
      }
    }
  }
}


__global__ void __launch_bounds__(256, 4) act_split_generic_kernel(ActSplitArgs a, int xs) {
  griddep_wait();
  griddep_launch();  // the consumer's CTAs may take their SM slots and run their prologue now
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int Cp = a.Cp;
  float* sc_s = reinterpret_cast<float*>(sm_raw);  // fused dY correction: c1 -> sc_s, c2 -> sh_s
  float* sh_s = sc_s + Cp;
  float* mean_s = sh_s + Cp;
  float* is_s = mean_s + Cp;
  unsigned char* tile = sm_raw + 16 * (size_t)Cp;
  if (a.fix) {
    const FixDyArgs& f = a.fx;
    for (int c = threadIdx.x; c < Cp; c += blockDim.x) {
      float c1 = 0.f, c2 = 0.f, mean = 0.f, is = 0.f;
      if (c < a.C) {
        const double m = f.sum[c] * f.inv_count;
        double var = f.sumsq[c] * f.inv_count - m * m;
        if (var < 0.0) var = 0.0;
        const double isd = 1.0 / sqrt(var + (double)f.eps);
        double d1 = 0.0, d2 = 0.0;
        for (int l = 0; l < f.n_cons; ++l) {
          const double sc = (double)f.cons_gamma[l][c] * (double)(float)isd;
          d1 += sc * f.cons_bsum[l][c];
          d2 += sc * f.cons_bsum[l][f.cons_C[l] + c];
        }
        c1 = (float)(d1 * f.inv_count);
        c2 = (float)(d2 * f.inv_count);
        mean = (float)m;
        is = (float)isd;
      }
      sc_s[c] = c1;
      sh_s[c] = c2;
      mean_s[c] = mean;
      is_s[c] = is;
    }
  } else if (a.pro) {
    for (int c = threadIdx.x; c < Cp; c += blockDim.x) {
      float s = 0.f, h = 0.f;
      if (c < a.C) bn_consts_w(a.bn, c, s, h);
      sc_s[c] = s;
      sh_s[c] = h;
    }
  }
  float mul = a.scale;
  if (a.dyn_max != nullptr) {
    // power of two that brings the buffer's running |gradient| maximum to 2^kDyTargetLog2
    // *dyn_max = largest single contribution any dgrad epilogue added to this gradient buffer; a slice
    // with n_cons consumers is bounded by n_cons times that (the contributions are summed)
    const unsigned m = *a.dyn_max;
    int e = m == 0u ? 0 : kDyTargetLog2 - ((int)((m >> 23) & 0xffu) - 127);
    if (a.fix)
      for (int c = 1; c < a.fx.n_cons; c <<= 1) --e;
    e = e < -100 ? -100 : (e > 100 ? 100 : e);
    mul = __uint_as_float((uint32_t)(e + 127) << 23);
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.dyn_inv = __uint_as_float((uint32_t)(127 - e) << 23);
  }
  const int oct = Cp >> 3;
  const int up = a.up;
  const int Hv = up ? 2 * a.Hs : a.Hs, Wv = up ? 2 * a.Ws : a.Ws;
  const size_t plane = (size_t)a.B * Hv * Wv * Cp;  // elements per piece plane
  const int XO = up ? 2 * xs : xs;                  // output pixels per segment
  const int rstride = XO * 16 + 16;                 // bytes per (piece, octet) row of the tile
  const int nseg = (a.Ws + xs - 1) / xs;
  const int n_work = a.B * a.Hs * nseg;
  const bool vec_in = !a.nchw && (a.ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(a.x) & 15u) == 0;
  const bool vec_fx = a.fix && (a.fx.ldX & 3) == 0 && (reinterpret_cast<uintptr_t>(a.fx.X) & 15u) == 0 && vec_in;
  // phase-1 thread grid: qp (power of two >= octets) threads along the channels of a pixel, 256/qp pixels
  // in flight; phase-2 thread grid: xp (power of two >= output pixels) threads along x
  int qsh = 0;
  while ((1 << qsh) < oct) ++qsh;
  const int tq = threadIdx.x & ((1 << qsh) - 1), tp = threadIdx.x >> qsh, pstep = 256 >> qsh;
  for (int work = blockIdx.x; work < n_work; work += gridDim.x) {
    const int seg = work % nseg, row = work / nseg;
    const int sy = row % a.Hs, b = row / a.Hs;
    const int x0 = seg * xs;
    const int nx = (a.Ws - x0) < xs ? (a.Ws - x0) : xs;
    __syncthreads();  // constants ready / previous tile drained
    // ---- phase 1: load, transform, split -> shared tile ------------------------------------
    if (tq < oct) {
      const int c = 8 * tq;
      for (int px = tp; px < nx; px += pstep) {
        const int sx = x0 + px;
        const size_t pix = ((size_t)b * a.Hs + sy) * a.Ws + sx;
        float v[8];
        if (a.nchw) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            v[k] = (c + k < a.C) ? a.x[(((size_t)b * a.C + c + k) * a.Hs + sy) * a.Ws + sx] : 0.f;
        } else {
          const float* p = a.x + pix * a.ldx + c;
          if (vec_in && c + 7 < a.C) {
            const float4 f0 = __ldg(reinterpret_cast<const float4*>(p));
            const float4 f1 = __ldg(reinterpret_cast<const float4*>(p + 4));
            v[0] = f0.x; v[1] = f0.y; v[2] = f0.z; v[3] = f0.w;
            v[4] = f1.x; v[5] = f1.y; v[6] = f1.z; v[7] = f1.w;
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = (c + k < a.C) ? p[k] : 0.f;
          }
        }
        if (a.pro) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            v[k] = (c + k < a.C) ? fmaxf(0.f, fmaf(v[k], sc_s[c + k], sh_s[c + k])) : 0.f;
        }
        if (a.fix) {
          // dY = G - c1 - xhat*c2 (lazy BatchNorm-backward mean corrections of every consumer), written
          // back in fp32 for the non-tensor-core consumers of the slice
          const float* xp = a.fx.X + pix * a.fx.ldX + c;
          float* gp = const_cast<float*>(a.x) + pix * a.ldx + c;
          if (vec_fx && c + 7 < a.C) {
            const float4 x0q = __ldg(reinterpret_cast<const float4*>(xp));
            const float4 x1q = __ldg(reinterpret_cast<const float4*>(xp + 4));
            const float xv[8] = {x0q.x, x0q.y, x0q.z, x0q.w, x1q.x, x1q.y, x1q.z, x1q.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float xh = (xv[k] - mean_s[c + k]) * is_s[c + k];
              v[k] = v[k] - sc_s[c + k] - xh * sh_s[c + k];
            }
            *reinterpret_cast<float4*>(gp) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(gp + 4) = make_float4(v[4], v[5], v[6], v[7]);
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              if (c + k < a.C) {
                const float xh = (xp[k] - mean_s[c + k]) * is_s[c + k];
                v[k] = v[k] - sc_s[c + k] - xh * sh_s[c + k];
                gp[k] = v[k];
              }
            }
          }
        }
        uint32_t h0[8], h1[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) split_mode(v[k] * mul, a.lowp, h0[k], h1[k]);
        const uint4 w0 = make_uint4(h0[0] | (h0[1] << 16), h0[2] | (h0[3] << 16), h0[4] | (h0[5] << 16), h0[6] | (h0[7] << 16));
        const uint4 w1 = make_uint4(h1[0] | (h1[1] << 16), h1[2] | (h1[3] << 16), h1[4] | (h1[5] << 16), h1[6] | (h1[7] << 16));
        const int ox = up ? 2 * px : px;
        unsigned char* t0 = tile + (size_t)tq * rstride + ox * 16;
        unsigned char* t1 = t0 + (size_t)oct * rstride;
        *reinterpret_cast<uint4*>(t0) = w0;
        *reinterpret_cast<uint4*>(t1) = w1;
        if (up == 1) {  // nearest x2: the same value at 2x and 2x+1
          *reinterpret_cast<uint4*>(t0 + 16) = w0;
          *reinterpret_cast<uint4*>(t1 + 16) = w1;
        } else if (up == 2) {  // zero insertion: odd positions are zero
          *reinterpret_cast<uint4*>(t0 + 16) = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(t1 + 16) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
    }
    __syncthreads();
    // ---- phase 2: tile rows -> global, x-contiguous 16-byte stores ---------------------------
    const int nxo = up ? 2 * nx : nx, xo0 = up ? 2 * x0 : x0;
    int xsh = 0;
    while ((1 << xsh) < nxo) ++xsh;
    const int ox = threadIdx.x & ((1 << xsh) - 1);
    if (ox < nxo) {
      for (int r = threadIdx.x >> xsh; r < (a.lowp ? 1 : kPieces) * oct; r += 256 >> xsh) {  // r = piece * oct + q
        const int piece = r >= oct ? 1 : 0, q = r - piece * oct;
        const uint4 val = *reinterpret_cast<const uint4*>(tile + (size_t)r * rstride + ox * 16);
        op16* dst = a.out + (size_t)piece * plane + ((((size_t)b * Hv + (up ? 2 * sy : sy)) * oct + q) * Wv + xo0 + ox) * 8;
        *reinterpret_cast<uint4*>(dst) = val;
        if (up) {  // second output row: a copy (nearest) or zeros (zero insertion)
          *reinterpret_cast<uint4*>(dst + (size_t)oct * Wv * 8) = up == 1 ? val : make_uint4(0u, 0u, 0u, 0u);
        }
      }
    }
  }
}


// ---------------------------------------------------------------------------------------
// the weight-gradient kernel
// ---------------------------------------------------------------------------------------
// CT = 16-channel output tiles per CTA (GEMM N = pieces * 16 * CT): layers with many output channels
// amortise the A-operand shared-memory read of an MMA over a 2-4x wider N; their accumulators
// (TG taps x N columns) then cover one filter row per CTA instead of the whole filter.
template <int KS, int CT>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                TcWgradArgs t) {
  constexpr int T = KS * KS;
  constexpr int kNC = 16 * CT;
  constexpr int kStages = CT == 1 ? 3 : 2;
  constexpr int HWp = kTW + KS - 1;
  constexpr int HH = kTH + KS - 1;
  constexpr int HP = HH * HWp;
  constexpr int QA = kMC / 8;                       // channel octets of the A tile
  constexpr uint32_t A_BYTES = QA * HP * 16u;       // one fp16 plane tile
  constexpr uint32_t B_OCT = 128u * 16u;            // one co-octet: 128 pixels x 16 B
  constexpr uint32_t B_PIECE = 2u * CT * B_OCT;     // 16*CT output channels of one piece
  constexpr uint32_t STAGE = (A_BYTES + (uint32_t)kPieces * B_PIECE + 127u) & ~127u;
  constexpr int TG = CT == 1 ? (T <= 9 ? T : 10) : (T < 3 ? T : 3);  // filter taps per CTA (TG accumulators in TMEM)
  const int passes = t.lowp ? 1 : kPasses;          // one-piece modes: the single product a1 x d1
  const int pass = blockIdx.z % passes;             // which fp16 piece of `a` this CTA streams
  const int tap0 = (blockIdx.z / passes) * TG;      // first tap of this CTA's tap group
  const int ntap = (T - tap0) < TG ? (T - tap0) : TG;
  const int NP = t.lowp ? 1 : kPasses - pass;       // dY pieces multiplied: 2, 1
  const int NW = NP * kNC;                          // accumulator columns per tap

  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);        // [kStages]
  uint64_t* empty = full + kStages;                          // [kStages]
  uint64_t* acc_full = full + 2 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 2 * kStages + 1);
  unsigned char* stage0 = smem + 128;

  const int n_ci_tiles = (t.Cin + kMC - 1) / kMC;
  const int ci_tile = blockIdx.y % n_ci_tiles, co_tile = blockIdx.y / n_ci_tiles;
  const int c0 = ci_tile * kMC, n0 = co_tile * kNC;
  const int tiles_x = (t.Wo + kTW - 1) / kTW, tiles_y = (t.Ho + kTH - 1) / kTH;
  const int n_tiles = tiles_x * tiles_y * t.B;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(TG * NW)) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();  // operand planes / staging gradient come from earlier kernels of the step

  int my_tiles = 0;
  for (int pt = blockIdx.x; pt < n_tiles; pt += gridDim.x) ++my_tiles;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int it = 0;
      for (int pt = blockIdx.x; pt < n_tiles; pt += gridDim.x, ++it) {
        const int s = it % kStages;
        int rem = pt;
        const int tx = rem % tiles_x;
        rem /= tiles_x;
        const int ty = rem % tiles_y;
        const int b = rem / tiles_y;
        const int oy0 = ty * kTH, ox0 = tx * kTW;
        mbar_wait(&empty[s], (uint32_t)(((it / kStages) & 1) ^ 1));
        unsigned char* st = stage0 + (size_t)s * STAGE;
        mbar_arrive_expect_tx(&full[s], A_BYTES + (uint32_t)NP * B_PIECE);
        tma_load_4d(st, &tmA, (ox0 - t.pad) * 8, oy0 - t.pad, c0 >> 3, pass * t.B + b, &full[s]);
        for (int piece = 0; piece < NP; ++piece)
          tma_load_4d(st + A_BYTES + (size_t)piece * B_PIECE, &tmB, ox0 * 8, oy0, n0 >> 3, piece * t.B + b,
                      &full[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: whole warp walks the uniform loop, one elected lane issues =====
    if (my_tiles > 0) {
      const uint32_t idesc = make_idesc_f16_mn(128, NW) | idesc_fmt_bits(t.lowp);
      // MN-major canonical layout ((8,1,m),(8,k)) : ((1,8,SBO),(8,LBO)): SBO strides along the
      // channels (next octet), LBO along the pixels (next 8-pixel tile row)
      const uint32_t sbo_a = HP * 16u, lbo_a = HWp * 16u;
      const uint32_t sbo_b = B_OCT, lbo_b = 128u;
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < my_tiles; ++it) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_base = smem_u32(stage0 + (size_t)s * STAGE);
        const uint64_t ad0 = make_desc(a_base, lbo_a, sbo_a);
        const uint64_t bd0 = make_desc(a_base + A_BYTES, lbo_b, sbo_b);
        const uint32_t a_lo0 = (uint32_t)ad0, a_hi = (uint32_t)(ad0 >> 32);
        const uint32_t b_lo0 = (uint32_t)bd0, b_hi = (uint32_t)(bd0 >> 32);
        if (elect_one_w()) {
#pragma unroll 1
          for (int r = 0; r < kTH; r += 2) {
            const uint32_t b_lo = b_lo0 + (uint32_t)(r * 8);  // r * 128 B
            const uint32_t acc = (it | r) != 0 ? 1u : 0u;
#pragma unroll
            for (int j = 0; j < TG; ++j) {
              const int tap = tap0 + j;
              if (j < ntap) {
                const uint32_t a_lo = a_lo0 + (uint32_t)((r + tap / KS) * HWp + (tap % KS));
                umma_f16_w(tmem_base + (uint32_t)(j * NW), a_lo, a_hi, b_lo, b_hi, idesc, acc);
              }
            }
          }
          umma_commit(&empty[s]);
          if (it == my_tiles - 1) umma_commit(acc_full);
        }
        __syncwarp();
        if (++s == kStages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
    griddep_launch();
  } else if (my_tiles > 0) {
    // ===== epilogue: TMEM -> coalesced vector reductions into dWp[tap][ci][co] =====
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int quarter = warp & 3;
    const int ci = c0 + quarter * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float osc = t.out_scale * (t.dyn_scale != nullptr ? *t.dyn_scale : 1.f);
    for (int j = 0; j < ntap; ++j) {
      const int tap = tap0 + j;
#pragma unroll 1
      for (int ct = 0; ct < CT; ++ct) {
        if (n0 + 16 * ct >= t.co_pad) break;  // warp-uniform
        float v[16];
        tmem_ld16(taddr + (uint32_t)(j * NW + (NP - 1) * kNC + 16 * ct), v);  // smallest terms first
        for (int piece = NP - 2; piece >= 0; --piece) {
          float x[16];
          tmem_ld16(taddr + (uint32_t)(j * NW + piece * kNC + 16 * ct), x);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += x[i];
        }
        if (ci < t.Cin) {
          float* dst = t.dwp + ((size_t)tap * t.ci_pad + ci) * t.co_pad + n0 + 16 * ct;
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            red_add_v4(dst + i, v[i] * osc, v[i + 1] * osc, v[i + 2] * osc, v[i + 3] * osc);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------
// Thin 3x3 layers (16 output channels, stride 1, pad 1): all nine taps in ONE GEMM-N.
//
//   dW[tap][ci][co] = sum_q act[q][ci] * dY[q - s_tap][co]          (q over the image, dY zero outside)
//
// wgrad_tc_kernel<3, 1> keeps dY fixed and shifts the activation tile per tap: 72 MMAs of N = 32 / 16 per
// 128-pixel tile, each paced by the shared-memory fetch of its M = 128 activation operand (~45 clk for
// 1-2 KB of B).  Here the dY tile is loaded NINE times at the shifted origins (TMA zero-fills outside the image;
// dY is the small operand: 4 KB per piece and tap) into consecutive N-chunks, so that one K = 16 step is
// a1 x [taps 0-7] (N = 256) + a1 x [tap 8] (N = 32), or a2 x [all taps of d1] (N = 144): 16 / 8 MMAs per tile and the
// activation tile needs no halo.  The accumulator columns are those of wgrad_tc_kernel<3, 1> (tap * NW +
// piece * 16 + co): same epilogue, same staging gradient.
// ---------------------------------------------------------------------------------------
constexpr int kTnStages = 2;
constexpr uint32_t kTnABytes = (kMC / 8) * 128u * 16u;   // 16 channel octets x 128 pixels x 16 B
constexpr uint32_t kTnBOct = 128u * 16u;                 // one (tap, piece, co-octet) chunk
constexpr uint32_t kTnStage = kTnABytes + 9u * kPieces * 2u * kTnBOct;

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcWgradArgs t) {
  constexpr int KS = 3, T = 9, kNC = 16;
  const int passes = t.lowp ? 1 : kPasses;
  const int pass = blockIdx.z % passes;             // which fp16 piece of `a` this CTA streams
  const int NP = t.lowp ? 1 : kPasses - pass;       // dY pieces multiplied: 2, 1
  const int NW = NP * kNC;                          // accumulator columns per tap

  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);        // [kTnStages]
  uint64_t* empty = full + kTnStages;                        // [kTnStages]
  uint64_t* acc_full = full + 2 * kTnStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 2 * kTnStages + 1);
  unsigned char* stage0 = smem + 128;

  const int c0 = blockIdx.y * kMC;
  const int tiles_x = (t.Wo + kTW - 1) / kTW, tiles_y = (t.Ho + kTH - 1) / kTH;
  const int n_tiles = tiles_x * tiles_y * t.B;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(T * NW)) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kTnStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();  // operand planes / staging gradient come from earlier kernels of the step

  int my_tiles = 0;
  for (int pt = blockIdx.x; pt < n_tiles; pt += gridDim.x) ++my_tiles;

  if (warp == 0) {
    // ===== TMA producer: the activation tile and the nine shifted dY tiles =====
    if (lane == 0) {
      int it = 0;
      for (int pt = blockIdx.x; pt < n_tiles; pt += gridDim.x, ++it) {
        const int s = it % kTnStages;
        int rem = pt;
        const int tx = rem % tiles_x;
        rem /= tiles_x;
        const int ty = rem % tiles_y;
        const int b = rem / tiles_y;
        const int y0 = ty * kTH, x0 = tx * kTW;
        mbar_wait(&empty[s], (uint32_t)(((it / kTnStages) & 1) ^ 1));
        unsigned char* st = stage0 + (size_t)s * kTnStage;
        mbar_arrive_expect_tx(&full[s], kTnABytes + (uint32_t)(T * NP) * 2u * kTnBOct);
        tma_load_4d(st, &tmA, x0 * 8, y0, c0 >> 3, pass * t.B + b, &full[s]);
        unsigned char* bs = st + kTnABytes;
        for (int tap = 0; tap < T; ++tap)
          for (int piece = 0; piece < NP; ++piece)
            tma_load_4d(bs + (size_t)(tap * NP + piece) * 2u * kTnBOct, &tmB, (x0 + t.pad - tap % KS) * 8,
                        y0 + t.pad - tap / KS, 0, piece * t.B + b, &full[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (my_tiles > 0) {
      // MN-major canonical layout: SBO strides along the channels (next octet), LBO along the pixels
      const uint32_t sbo = kTnBOct, lbo = 128u;
      const uint32_t ncols = (uint32_t)(T * NW);                 // 288 (a1 x [d1|d2]) or 144 (a2 x d1)
      const uint32_t n_hi = ncols > 256u ? 256u : ncols, n_lo = ncols - n_hi;
      const uint32_t idesc_hi = make_idesc_f16_mn(128, (int)n_hi) | idesc_fmt_bits(t.lowp);
      const uint32_t idesc_lo = n_lo ? (make_idesc_f16_mn(128, (int)n_lo) | idesc_fmt_bits(t.lowp)) : 0u;
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < my_tiles; ++it) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_base = smem_u32(stage0 + (size_t)s * kTnStage);
        const uint64_t ad0 = make_desc(a_base, lbo, sbo);
        const uint64_t bd0 = make_desc(a_base + kTnABytes, lbo, sbo);
        const uint32_t a_lo0 = (uint32_t)ad0, a_hi = (uint32_t)(ad0 >> 32);
        const uint32_t b_lo0 = (uint32_t)bd0, b_hi = (uint32_t)(bd0 >> 32);
        if (elect_one_w()) {
#pragma unroll 1
          for (int r = 0; r < kTH; r += 2) {                      // K = 16 pixels = two 8-pixel tile rows
            const uint32_t off = (uint32_t)(r * 8);              // r * 128 B in 16-byte units
            const uint32_t acc = (it | r) != 0 ? 1u : 0u;
            umma_f16_w(tmem_base, a_lo0 + off, a_hi, b_lo0 + off, b_hi, idesc_hi, acc);
            if (n_lo)   // chunks 32.. (= tap 8): 32 chunks x 2 KB further on
              umma_f16_w(tmem_base + n_hi, a_lo0 + off, a_hi, b_lo0 + off + ((32u * kTnBOct) >> 4), b_hi, idesc_lo, acc);
          }
          umma_commit(&empty[s]);
          if (it == my_tiles - 1) umma_commit(acc_full);
        }
        __syncwarp();
        if (++s == kTnStages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
    griddep_launch();
  } else if (my_tiles > 0) {
    // ===== epilogue: TMEM -> coalesced vector reductions into dWp[tap][ci][co] =====
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int quarter = warp & 3;
    const int ci = c0 + quarter * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float osc = t.out_scale * (t.dyn_scale != nullptr ? *t.dyn_scale : 1.f);
    for (int tap = 0; tap < T; ++tap) {
      float v[16];
      tmem_ld16(taddr + (uint32_t)(tap * NW + (NP - 1) * kNC), v);  // smallest terms first
      for (int piece = NP - 2; piece >= 0; --piece) {
        float x[16];
        tmem_ld16(taddr + (uint32_t)(tap * NW + piece * kNC), x);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += x[i];
      }
      if (ci < t.Cin) {
        float* dst = t.dwp + ((size_t)tap * t.ci_pad + ci) * t.co_pad;
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          red_add_v4(dst + i, v[i] * osc, v[i + 1] * osc, v[i + 2] * osc, v[i + 3] * osc);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// dY expansion for the taps-in-N weight gradient (DyIm2colArgs): one thread per (pixel, n-octet),
// x-contiguous gathers and 16-byte stores.
__global__ void __launch_bounds__(256) dy_im2col_kernel(DyIm2colArgs a) {
  // column n = tap * Cout + co reads dy[b][co][y - ty + pad][x - tx + pad]: the per-column offsets are the
  // same for every pixel, so they are tabulated once per CTA (the first version divided per element and ran
  // issue-bound)
  __shared__ int4 tab[kIm2colMaxN];   // {offset co*H*W + dy*W + dx, dy, dx, valid}
  const int nreal = a.KS * a.KS * a.Cout;
  for (int n = threadIdx.x; n < a.Np; n += blockDim.x) {
    int4 t = make_int4(0, 0, 0, 0);
    if (n < nreal) {
      const int tap = n / a.Cout, co = n - tap * a.Cout;
      const int dy = a.pad - tap / a.KS, dx = a.pad - tap % a.KS;
      t = make_int4((co * a.H + dy) * a.W + dx, dy, dx, 1);
    }
    tab[n] = t;
  }
  griddep_wait();
  float mul = 1.f;
  if (a.dyn_max != nullptr) {
    const unsigned m = *a.dyn_max;
    int e = m == 0u ? 0 : kDyTargetLog2 - ((int)((m >> 23) & 0xffu) - 127);
    e = e < -100 ? -100 : (e > 100 ? 100 : e);
    mul = __uint_as_float((uint32_t)(e + 127) << 23);
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.dyn_inv = __uint_as_float((uint32_t)(127 - e) << 23);
  }
  __syncthreads();
  const int oct = a.Np >> 3;
  const size_t plane = (size_t)a.B * a.H * a.W * a.Np;
  const size_t total = (size_t)a.B * a.H * oct * a.W;
  const size_t img = (size_t)a.Cout * a.H * a.W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % a.W);
    size_t r = i / a.W;
    const int q = (int)(r % oct);
    r /= oct;
    const int y = (int)(r % a.H), b = (int)(r / a.H);
    const float* base = a.dy + (size_t)b * img + (size_t)y * a.W + x;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int4 t = tab[8 * q + k];
      const bool in = t.w && (unsigned)(y + t.y) < (unsigned)a.H && (unsigned)(x + t.z) < (unsigned)a.W;
      v[k] = in ? __ldg(base + t.x) * mul : 0.f;
    }
    uint32_t p1[4], p2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (a.lowp == LOWP_BF16) {
        p1[k] = pack_bf2s(v[2 * k], v[2 * k + 1]);
        p2[k] = 0u;
      } else {
        p1[k] = pack_h2s(v[2 * k], v[2 * k + 1]);
        const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&p1[k]));
        p2[k] = a.lowp ? 0u : pack_h2s(v[2 * k] - fl.x, v[2 * k + 1] - fl.y);
      }
    }
    op16* dst = a.out + i * 8;  // (((b*H + y)*oct + q)*W + x)*8
    *reinterpret_cast<uint4*>(dst) = make_uint4(p1[0], p1[1], p1[2], p1[3]);
    *reinterpret_cast<uint4*>(dst + plane) = make_uint4(p2[0], p2[1], p2[2], p2[3]);
  }
}

// max |x| as float bits (non-negative floats order like unsigned integers)
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, size_t n, unsigned* out) {
  griddep_wait();
  unsigned m = 0u;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned u = __float_as_uint(x[i]) & 0x7fffffffu;
    m = u > m ? u : m;
  }
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && m != 0u) atomicMax(out, m);
}

// dW_OIHW[co][ci][tap] += dWp[tap][ci][co]; dWp is cleared for the next step.  One block per
// (layer, 8 input channels): the [tap][8 ci][co] slab is read with co-contiguous loads, transposed in
// shared memory to [co][8 ci][tap] — in OIHW that is one contiguous run of 8*T floats per output
// channel — and accumulated with consecutive threads on consecutive addresses.
constexpr int kUnpackCi = 8;
__global__ void __launch_bounds__(256) wgrad_unpack_kernel(const TcWgradUnpack* tab) {
  griddep_wait();
  extern __shared__ float tile_u[];  // [Cout][8][T]
  const TcWgradUnpack d = tab[blockIdx.y];
  const int ci0 = blockIdx.x * kUnpackCi;
  if (ci0 >= d.Cin) return;
  if (d.taps_in_n > 0) {
    // staged as dwp[ci][tap * Cout + co] (few output channels): a small scattered fold
    const int Tn = d.taps_in_n, nci = (d.Cin - ci0) < kUnpackCi ? (d.Cin - ci0) : kUnpackCi;
    for (int i = threadIdx.x; i < nci * Tn * d.Cout; i += blockDim.x) {
      const int n = i % (Tn * d.Cout), c = i / (Tn * d.Cout);
      const int tap = n / d.Cout, co = n - tap * d.Cout;
      float* src = d.dwp + (size_t)(ci0 + c) * d.co_pad + n;
      d.dw[((size_t)co * d.Cin + ci0 + c) * Tn + tap] += *src;
      *src = 0.f;
    }
    return;
  }
  const int T = d.KS * d.KS;
  const int nci = (d.Cin - ci0) < kUnpackCi ? (d.Cin - ci0) : kUnpackCi;
  const size_t tap_stride = (size_t)d.ci_pad * d.co_pad;
  const int rows = T * nci;  // (tap, ci) rows of Cout floats
  constexpr int U = 4;       // independent loads in flight per thread
  const int n1 = rows * d.Cout;
  for (int base = threadIdx.x; base < n1; base += U * blockDim.x) {
    float* src[U];
    float v[U];
    int dsti[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u * blockDim.x;
      src[u] = nullptr;
      if (i < n1) {
        const int co = i % d.Cout, r = i / d.Cout;
        const int c = r % nci, tap = r / nci;
        src[u] = d.dwp + (size_t)tap * tap_stride + (size_t)(ci0 + c) * d.co_pad + co;
        dsti[u] = (co * kUnpackCi + c) * T + tap;
        v[u] = *src[u];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (src[u] != nullptr) {
        tile_u[dsti[u]] = v[u];
        *src[u] = 0.f;
      }
    }
  }
  __syncthreads();
  const int run = nci * T;  // contiguous floats per output channel
  const int n2 = d.Cout * run;
  for (int base = threadIdx.x; base < n2; base += U * blockDim.x) {
    float* dst[U];
    float v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u * blockDim.x;
      dst[u] = nullptr;
      if (i < n2) {
        const int co = i / run, k = i - co * run;
        dst[u] = d.dw + ((size_t)co * d.Cin + ci0) * T + k;
        v[u] = *dst[u] + tile_u[(size_t)co * kUnpackCi * T + k];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (dst[u] != nullptr) *dst[u] = v[u];
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
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

// planes [2*B][H][Cp/8][W][8] fp16 viewed as (x*8+c8, y, octet, plane*B+b); box (bx*8, by, boct, 1)
// -> shared memory [octet][y][x][16 B]
int make_plane_map(CUtensorMap* tm, const op16* base, int B, int H, int W, int Cp, int bx, int by,
                   int boct) {
  EncodeFn enc = get_encode();
  PDES_REQUIRE(enc != nullptr, PDES_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t oct = (cuuint64_t)(Cp / 8);
  const cuuint64_t gdim[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, oct, (cuuint64_t)kPieces * B};
  const cuuint64_t gstr[3] = {oct * W * 16, (cuuint64_t)W * 16, (cuuint64_t)H * oct * W * 16};
  const cuuint32_t box[4] = {(cuuint32_t)bx * 8, (cuuint32_t)by, (cuuint32_t)boct, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<op16*>(base), gdim, gstr,
                         box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PDES_REQUIRE(r == CUDA_SUCCESS, PDES_ERR_CUDA, "cuTensorMapEncodeTiled failed with code %d", (int)r);
  return PDES_OK;
}

// CTAs per SM-slot the pixel range is split over (one CTA fits per SM: 1 = a single wave)
int wg_waves() {
  static int w = 0;
  if (w == 0) {
    const char* e = getenv("PDES_WG_WAVES");
    w = e ? atoi(e) : 1;
    if (w < 1) w = 1;
  }
  return w;
}

// thin 3x3 layers through wgrad_tn_kernel (PDES_WGRAD_TAPSN=0: wgrad_tc_kernel<3, 1>)
bool wg_taps_n_on() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PDES_WGRAD_TAPSN");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

template <int KS, int CT>
size_t wg_smem() {
  constexpr int HP = (kTH + KS - 1) * (kTW + KS - 1);
  const size_t stage = ((size_t)(kMC / 8) * HP * 16 + (size_t)kPieces * 2 * CT * 128 * 16 + 127) & ~(size_t)127;
  return 128 + (CT == 1 ? 3 : 2) * stage;
}

}  // namespace

bool wgrad_tc_supported(int KS, int stride) { return (KS == 1 || KS == 3 || KS == 5) && stride == 1; }

void wgrad_tc_dims(int Cin, int Cout, int* ci_pad, int* co_pad) {
  *ci_pad = (Cin + kMC - 1) / kMC * kMC;
  *co_pad = (Cout + 15) / 16 * 16;
}

size_t act_planes_bytes(int B, int H, int W, int C) {
  const int Cp = (C + 7) & ~7;
  return (size_t)kPieces * B * H * W * Cp * sizeof(op16);
}

int launch_dy_im2col(const DyIm2colArgs& a, cudaStream_t st) {
  PDES_REQUIRE(a.dy && a.out && a.Np % 8 == 0 && a.Np >= a.KS * a.KS * a.Cout && a.Np <= kIm2colMaxN, PDES_ERR_INVALID,
               "dy_im2col: invalid arguments");
  const size_t total = (size_t)a.B * a.H * (a.Np / 8) * a.W;
  int blocks = (int)((total + 255) / 256);
  const int cap = sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  PDES_CUDA(launch_pdl(dy_im2col_kernel, dim3(blocks), dim3(256), 0, st, a));
  return PDES_OK;
}

int launch_absmax(const float* x, size_t n, unsigned* out, cudaStream_t st) {
  int blocks = (int)((n / 4 + 255) / 256);
  if (blocks > 2 * sm_count()) blocks = 2 * sm_count();
  if (blocks < 1) blocks = 1;
  PDES_CUDA(launch_pdl(absmax_kernel, dim3(blocks), dim3(256), 0, st, x, n, out));
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

int launch_act_split(const ActSplitArgs& a, cudaStream_t st) {
  PDES_REQUIRE(a.Cp % 8 == 0 && a.Cp >= a.C && a.Cp <= 1024, PDES_ERR_INVALID,
               "act_split: padded channel count %d invalid", a.Cp);
  // pixels of a source row per work item: the whole row unless its tile would not fit 40 KB
  int xs = a.Ws;
  auto tile_bytes = [&](int x) { return (size_t)kPieces * (a.Cp / 8) * ((size_t)(a.up ? 2 * x : x) * 16 + 16); };
  while (xs > 1 && tile_bytes(xs) > 40 * 1024) xs = (xs + 1) / 2;
  // staged (cp.async) input: needs 16-byte addressable NHWC rows, also of the activations the fused dY
  // correction reads
  bool staged = !a.nchw && (a.ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(a.x) & 15u) == 0;
  if (a.fix) staged = staged && (a.fx.ldX & 3) == 0 && (reinterpret_cast<uintptr_t>(a.fx.X) & 15u) == 0;
  const size_t stage_bytes = staged ? (size_t)2 * (a.fix ? 2 : 1) * xs * a.Cp * sizeof(float) : 0;
  const size_t smem = 16 * (size_t)a.Cp + ((tile_bytes(xs) + 15) & ~(size_t)15) + stage_bytes;
  PDES_REQUIRE(smem <= (staged ? 200 : 48) * 1024, PDES_ERR_UNSUPPORTED, "act_split: %d channels need %zu bytes of shared memory", a.Cp, smem);
  if (staged) PDES_ENSURE_SMEM(act_split_staged_kernel, smem);
  const int n_work = a.B * a.Hs * ((a.Ws + xs - 1) / xs);
  // persistent CTAs: as many as fit an SM (shared memory, 2 x 256 threads), each looping over its items
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm);
  int blocks = n_work;
  const int cap = sm_count() * (staged ? per_sm : 8);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (staged)
    PDES_CUDA(launch_pdl(act_split_staged_kernel, dim3(blocks), dim3(256), smem, st, a, xs));
  else
    PDES_CUDA(launch_pdl(act_split_generic_kernel, dim3(blocks), dim3(256), smem, st, a, xs));
  return PDES_OK;
}

int launch_wgrad_tc(const TcWgradArgs& t, cudaStream_t st) {
  PDES_REQUIRE(wgrad_tc_supported(t.KS, 1), PDES_ERR_UNSUPPORTED, "wgrad_tc: 1x1 / 3x3 / 5x5 only");
  PDES_REQUIRE(t.planesA && t.planesB && t.dwp, PDES_ERR_INVALID, "wgrad_tc: null operand planes");
  const int CpA = (t.Cin + 7) & ~7, CpB = (t.Cout + 7) & ~7;
  const int n_ci = (t.Cin + kMC - 1) / kMC, n_co = (t.Cout + 15) / 16;
  const int CT = n_co >= 3 ? 4 : n_co;  // 16-channel output tiles per CTA: 1, 2 or 4
  CUtensorMap tmA, tmB;
  int rc = make_plane_map(&tmA, t.planesA, t.B, t.Hv, t.Wv, CpA, kTW + t.KS - 1, kTH + t.KS - 1, kMC / 8);
  if (rc) return rc;
  rc = make_plane_map(&tmB, t.planesB, t.B, t.Ho, t.Wo, CpB, kTW, kTH, 2 * CT);
  if (rc) return rc;
  const int tiles = ((t.Wo + kTW - 1) / kTW) * ((t.Ho + kTH - 1) / kTH) * t.B;
  if (wg_taps_n_on() && t.KS == 3 && t.pad == 1 && n_co == 1 && t.Hv == t.Ho && t.Wv == t.Wo) {
    // thin 3x3 layer: all taps in one GEMM-N (wgrad_tn_kernel); the activation tile has no halo
    rc = make_plane_map(&tmA, t.planesA, t.B, t.Hv, t.Wv, CpA, kTW, kTH, kMC / 8);
    if (rc) return rc;
    const int passes = t.lowp ? 1 : kPasses;
    int P = (wg_waves() * sm_count()) / (n_ci * passes);
    if (P < 1) P = 1;
    if (P > tiles) P = tiles;
    const size_t smem = 128 + (size_t)kTnStages * kTnStage;
    PDES_ENSURE_SMEM(wgrad_tn_kernel, smem);
    PDES_CUDA(launch_pdl(wgrad_tn_kernel, dim3(P, n_ci, passes), dim3(kThreads), smem, st, tmA, tmB, t));
    PDES_LAUNCH_CHECK();
    return PDES_OK;
  }
  const int n_cog = (n_co + CT - 1) / CT;
  const int T = t.KS * t.KS;
  const int TG = CT == 1 ? (T <= 9 ? T : 10) : (T < 3 ? T : 3);
  const int tap_groups = (T + TG - 1) / TG;
  const int passes = t.lowp ? 1 : kPasses;
  int P = (wg_waves() * sm_count()) / (n_ci * n_cog * passes * tap_groups);
  if (P < 1) P = 1;
  if (P > tiles) P = tiles;
  dim3 grid(P, n_ci * n_cog, passes * tap_groups);
#define PDES_WG_LAUNCH(KSV, CTV)                                                                             \
  {                                                                                                          \
    const size_t smem = wg_smem<KSV, CTV>();                                                                 \
    PDES_ENSURE_SMEM((wgrad_tc_kernel<KSV, CTV>), smem);                                                     \
    PDES_CUDA(launch_pdl(wgrad_tc_kernel<KSV, CTV>, grid, dim3(kThreads), smem, st, tmA, tmB, t));           \
  }
#define PDES_WG_CT(KSV)                      \
  {                                          \
    if (CT == 1) PDES_WG_LAUNCH(KSV, 1)      \
    else if (CT == 2) PDES_WG_LAUNCH(KSV, 2) \
    else PDES_WG_LAUNCH(KSV, 4)              \
  }
  if (t.KS == 3) PDES_WG_CT(3)
  else if (t.KS == 1) PDES_WG_CT(1)
  else PDES_WG_CT(5)
#undef PDES_WG_CT
#undef PDES_WG_LAUNCH
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

int launch_wgrad_unpack(const TcWgradUnpack* dev_table, int n, int max_cin, int max_slab_floats, cudaStream_t st) {
  if (n == 0) return PDES_OK;
  const size_t smem = sizeof(float) * (size_t)max_slab_floats;  // Cout * 8 * T of the largest layer
  if (smem > 48 * 1024) PDES_ENSURE_SMEM(wgrad_unpack_kernel, smem);
  PDES_CUDA(launch_pdl(wgrad_unpack_kernel, dim3((max_cin + kUnpackCi - 1) / kUnpackCi, n), dim3(256), smem, st, dev_table));
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

}  // namespace pdes

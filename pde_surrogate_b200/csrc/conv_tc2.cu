// conv_tc2.cu — TMA-fed, persistent, warp-specialised tcgen05 implicit-GEMM convolution
// (forward and dgrad) on two-piece fp16 operand splits.
//
// Operands never pass through registers: the activation side (BN+ReLU'd, optionally nearest-
// upsampled input — or the corrected dY slice for dgrad) is pre-split once per layer into two
// fp16 planes by act_split_kernel; 4-D TMA boxes (cp.async.bulk.tensor, zero fill outside the
// image = convolution padding) drop a (TH+2)x(TW+2) halo tile of one channel chunk into shared
// memory as [channel octet][halo pixel][16 B] — the canonical no-swizzle K-major UMMA layout —
// and the KSxKS filter taps are shifted descriptors into that one tile (im2col-free).  Filter
// tiles are pre-packed per (chunk, tap) in their shared-memory image and arrive by 1-D bulk TMA.
//
// fp32 parity: every operand is two fp16 pieces of its power-of-two scaled value (conv_tc.cuh); the
// three products of weight >= 2^-11
//   a1*[w1|w2], a2*[w1]
// are issued as two tcgen05.mma.kind::f16 instructions (three when 2N > 256) whose N-concatenated
// filter operand sends each product class to its own TMEM column group (G0 = a1w1,
// G1 = a1w2 + a2w1): tensor-core accumulation truncates, so big and small terms never share an
// accumulator, the K range is spread over S accumulator sets, and the epilogue adds everything
// with round-to-nearest fp32 and multiplies by the exact inverse of the operand scales.
//
// Warp roles (384 threads, persistent over pixel tiles): warp 0 = TMA producer of activation
// tiles, warp 1 = TMA producer of filter tiles, warp 2 = TMEM allocator + MMA issuer (one lane),
// warps 4-11 = epilogue (two warps per TMEM lane quarter).  With two TMEM accumulator stages the
// epilogue of tile i overlaps the main loop of tile i+1.
#include <cuda.h>
#include <stdlib.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include "conv.cuh"
#include "conv_tc.cuh"
#include "tc_common.cuh"

namespace pdes {
namespace {

using namespace tc;

constexpr int kThreads = 384;
constexpr int kTH = 16, kTW = 8;
constexpr int kEpiWarp0 = 4, kEpiWarps = 8;

__device__ __forceinline__ uint32_t make_idesc_f16(int M, int N) {  // D fp32, A/B fp16 (format 0), K-major
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
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
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
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
// same instruction with the 64-bit descriptors given as (lo, hi) words: all per-MMA address
// arithmetic happens on the 32-bit low words (start address / LBO fields)
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
__device__ __forceinline__ uint32_t f2h_sat(float x) {
  unsigned short h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
  return (uint32_t)h;
}
__device__ __forceinline__ void bn_consts2(const BnSrc& s, int c, float& scale, float& shift, float& mean,
                                           float& invstd) {
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
  invstd = (float)(1.0 / sqrt(var + (double)s.eps));
  mean = (float)m;
  scale = s.gamma[c] * invstd;
  shift = s.beta[c] - mean * scale;
}

// instruction list of one k16 step per MODE: (a piece, first filter piece, #pieces, first group)
//  MODE 0: 2N <= 256   a1x[w1|w2] -> G0..1, a2xw1 -> G1
//  MODE 1: 2N >  256   three single-piece instructions a1xw1 -> G0, a1xw2 -> G1, a2xw1 -> G1
// nibble-packed tables (entry i = bits [4i, 4i+4)) so that they fold to constants in device code
template <int MODE> struct Ops;
template <> struct Ops<0> {
  static constexpr int n = 2;
  static constexpr uint32_t A = 0x10u, Bp = 0x00u, NP = 0x12u, G = 0x10u;
};
template <> struct Ops<1> {
  static constexpr int n = 3;
  static constexpr uint32_t A = 0x100u, Bp = 0x010u, NP = 0x111u, G = 0x110u;
};
//  MODE 2: one-piece (fp16 / bf16) operands: the single product a1xw1 -> G0
template <> struct Ops<2> {
  static constexpr int n = 1;
  static constexpr uint32_t A = 0x0u, Bp = 0x0u, NP = 0x1u, G = 0x0u;
};
#define OPF(tab, i) ((int)(((tab) >> (4 * (i))) & 0xFu))
// ops that are the first writer of (all of) their column groups within one k16 step
template <int MODE> struct FirstW { static constexpr uint32_t mask = MODE == 1 ? 0x3u : 0x1u; };

template <int KS, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, Tc2Args t) {
  constexpr int T = KS * KS;
  constexpr int HWp = kTW + KS - 1;
  constexpr int HH = kTH + KS - 1;
  constexpr int HP = HH * HWp;
  const ConvArgs& a = t.c;
  const int N = t.N, KC = t.KC, NB = t.NB, TPB = t.TPB, AST = t.AST;
  const int koct = KC >> 3;
  const uint32_t a_piece_bytes = (uint32_t)koct * HP * 16u;
  const uint32_t a_stage_bytes = (uint32_t)kPieces * a_piece_bytes;
  const uint32_t b_tap_bytes = (uint32_t)koct * (uint32_t)kPieces * (uint32_t)N * 16u;
  const uint32_t b_stage_bytes = (uint32_t)TPB * b_tap_bytes;

  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint64_t* a_full = bars;        // [4]
  uint64_t* a_empty = bars + 4;   // [4]
  uint64_t* b_full = bars + 8;    // [8]
  uint64_t* b_empty = bars + 16;  // [8]
  uint64_t* acc_full = bars + 24; // [2]
  uint64_t* acc_empty = bars + 26;// [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);
  float* ep_s = reinterpret_cast<float*>(smem + 256);   // 4*N : scale, shift, mean, invstd
  float* red_s = ep_s + 4 * N;                          // kEpiWarps * N * 2
  const size_t hdr = (256 + sizeof(float) * (size_t)(4 + 2 * kEpiWarps) * N + 127) & ~(size_t)127;
  unsigned char* A_s = smem + hdr;
  unsigned char* B_s = A_s + (size_t)AST * a_stage_bytes;

  const int tiles_x = (a.Wo + kTW - 1) / kTW, tiles_y = (a.Ho + kTH - 1) / kTH;
  const int n_tiles = tiles_x * tiles_y * a.B;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = t.nchunks;
  const int bstages_per_chunk = T / TPB;
  const int TS = t.TS;
  const uint32_t ts_cols = (uint32_t)(t.S * t.ngroups * N);
  uint32_t tmem_cols = 32;
  while (tmem_cols < ts_cols * (uint32_t)TS) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < AST; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < NB; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], kEpiWarps);
    }
    fence_mbar_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  griddep_wait();  // everything below reads what earlier kernels of the step produced
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp >= kEpiWarp0) {
    // epilogue constants (BatchNorm of the consumer: dependent global loads + fp64 math): computed by
    // the epilogue warps only, behind the CTA barrier, so the producers and the MMA issuer start at once
    const int tt = threadIdx.x - kEpiWarp0 * 32;
    for (int n = tt; n < N; n += kEpiWarps * 32) {
      // per channel (scale, shift, invstd, -mean*invstd): z = x*scale + shift, xhat = x*invstd + c;
      // all zero for the padding channels, whose dz is then 0 without a bounds test
      float s = 0.f, h = 0.f, m = 0.f, is = 0.f;
      if (a.epi == EPI_BNBWD && n < a.Cout) bn_consts2(a.fbn, n, s, h, m, is);
      *reinterpret_cast<float4*>(ep_s + 4 * n) = make_float4(s, h, is, -m * is);
    }
    for (int i = tt; i < kEpiWarps * N * 2; i += kEpiWarps * 32) red_s[i] = 0.f;
    named_bar_sync(1, kEpiWarps * 32);
  }
#define DBG(slot) do { if (t.dbg) t.dbg[(size_t)blockIdx.x * 16 + (slot)] = clock64(); } while (0)
  if (threadIdx.x == 0) DBG(0);

  if (warp == 0) {
    // ===== TMA producer: activation halo tiles (two fp16 pieces per chunk) =====
    if (lane == 0) {
      int q = 0;  // global chunk counter
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int rem = tile;
        const int tx = rem % tiles_x;
        rem /= tiles_x;
        const int ty = rem % tiles_y;
        const int b = rem / tiles_y;
        const int ix0 = tx * kTW - a.pad, iy0 = ty * kTH - a.pad;
        for (int ch = 0; ch < nchunks; ++ch, ++q) {
          const int s = q % AST;
          mbar_wait(&a_empty[s], (uint32_t)(((q / AST) & 1) ^ 1));
          unsigned char* st = A_s + (size_t)s * a_stage_bytes;
          if (t.exp == 1) {  // timing experiment: no activation loads
            mbar_arrive(&a_full[s]);
            continue;
          }
          const int npl = MODE == 2 ? 1 : kPieces;   // one-piece modes never touch the second plane
          mbar_arrive_expect_tx(&a_full[s], (uint32_t)npl * a_piece_bytes);
#pragma unroll
          for (int p = 0; p < npl; ++p)
            tma_load_4d(st + (size_t)p * a_piece_bytes, &tmA, ix0 * 8, iy0, ch * koct, p * a.B + b, &a_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== TMA producer: filter tiles =====
    if (lane == 0) {
      int q = 0;  // global filter-stage counter
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int st = 0; st < nchunks * bstages_per_chunk; ++st, ++q) {
          const int s = q % NB;
          mbar_wait(&b_empty[s], (uint32_t)(((q / NB) & 1) ^ 1));
          if (t.exp == 2) {  // timing experiment: no filter loads
            mbar_arrive(&b_full[s]);
            continue;
          }
          mbar_arrive_expect_tx(&b_full[s], b_stage_bytes);
          tma_load_1d(B_s + (size_t)s * b_stage_bytes,
                      reinterpret_cast<const unsigned char*>(t.wpk) + (size_t)st * b_stage_bytes, b_stage_bytes,
                      &b_full[s]);
        }
      }
    }
  } else if (warp == 2) {
    // ===== MMA issuer: the whole warp walks the (warp-uniform) loop so that descriptors live in
    // uniform registers; one elected lane issues the tcgen05 instructions =====
    {
      using OP = Ops<MODE>;
      uint32_t idesc[OP::n];
#pragma unroll
      for (int i = 0; i < OP::n; ++i) idesc[i] = make_idesc_f16(128, OPF(OP::NP, i) * N) | idesc_fmt_bits(t.lowp);
      const uint32_t lbo_a = HP * 16u, sbo_a = HWp * 16u;            // K-major: LBO = next channel octet
      const uint32_t lbo_b = (uint32_t)kPieces * (uint32_t)N * 16u, sbo_b = 128u;   // [koct][piece][n][16 B]
      // per-op loop invariants: filter-piece offset (16-byte units) inside a k-octet block
      uint32_t boff[OP::n];
#pragma unroll
      for (int i = 0; i < OP::n; ++i) boff[i] = (uint32_t)(OPF(OP::Bp, i) * N);
      const uint32_t kstep_a = (2u * lbo_a) >> 4, kstep_b = (2u * lbo_b) >> 4;  // one k16 step, 16-byte units
      const uint32_t a_piece_u = a_piece_bytes >> 4;
      const int kc16 = KC / 16;
      // ring positions are kept as (stage, phase) counters: no integer division in the issue loop
      int sa = 0, sb = 0, ts = 0, tile_it = 0;
      uint32_t pa = 0, pb = 0, pt = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
        mbar_wait(&acc_empty[ts], pt ^ 1u);
        tc_fence_after();
        if (tile_it == 0 && lane == 0) DBG(1);
        const uint32_t d_base = tmem_base + (uint32_t)ts * ts_cols;
        int set = 0;
        for (int ch = 0; ch < nchunks; ++ch) {
          const bool fresh = ch < t.S;  // first visit of this accumulator set in this tile
          uint32_t dcol[OP::n];
#pragma unroll
          for (int i = 0; i < OP::n; ++i) dcol[i] = d_base + (uint32_t)((set * t.ngroups + OPF(OP::G, i)) * N);
          mbar_wait(&a_full[sa], pa);
          tc_fence_after();
          if (tile_it == 0 && ch == 0 && lane == 0) DBG(2);
          if (tile_it == 0 && ch == 1 && lane == 0) DBG(3);
          const uint64_t ad0 = make_desc(smem_u32(A_s + (size_t)sa * a_stage_bytes), lbo_a, sbo_a);
          const uint32_t a_lo0 = (uint32_t)ad0, a_hi = (uint32_t)(ad0 >> 32);
          if (TPB == T) {
            // the whole chunk's filter taps sit in one TMA stage: one wait, one elected issue block
            mbar_wait(&b_full[sb], pb);
            tc_fence_after();
            const uint64_t bd0 = make_desc(smem_u32(B_s + (size_t)sb * b_stage_bytes), lbo_b, sbo_b);
            const uint32_t b_lo0 = (uint32_t)bd0, b_hi = (uint32_t)(bd0 >> 32);
            const uint32_t b_tap_u = b_tap_bytes >> 4;
            if (elect_one()) {
#pragma unroll(KS <= 3 ? T : 1)
              for (int tap = 0; tap < T; ++tap) {
                const uint32_t a_tap = a_lo0 + (uint32_t)((tap / KS) * HWp + (tap % KS));
                const uint32_t b_tap = b_lo0 + (uint32_t)tap * b_tap_u;
#pragma unroll
                for (int k16 = 0; k16 < 2; ++k16) {
                  if (k16 < kc16) {
                    const uint32_t a_k = a_tap + (uint32_t)k16 * kstep_a;
                    const uint32_t b_k = b_tap + (uint32_t)k16 * kstep_b;
#pragma unroll
                    for (int i = 0; i < OP::n; ++i) {
                      const bool first = fresh && tap == 0 && k16 == 0 && ((FirstW<MODE>::mask >> i) & 1u);
                      umma_f16_w(dcol[i], a_k + (uint32_t)OPF(OP::A, i) * a_piece_u, a_hi, b_k + boff[i], b_hi,
                                  idesc[i], first ? 0u : 1u);
                    }
                  }
                }
              }
              umma_commit(&b_empty[sb]);
              umma_commit(&a_empty[sa]);
            }
            __syncwarp();
            if (++sb == NB) {
              sb = 0;
              pb ^= 1u;
            }
          } else {
            // filter taps arrive in groups of TPB per ring stage: one wait and one elected issue block
            // per stage (the tap loop stays rolled; its body is 2 x OP::n back-to-back MMAs)
            const uint32_t b_tap_u = b_tap_bytes >> 4;
            for (int tb = 0; tb < T; tb += TPB) {
              mbar_wait(&b_full[sb], pb);
              tc_fence_after();
              const uint64_t bd0 = make_desc(smem_u32(B_s + (size_t)sb * b_stage_bytes), lbo_b, sbo_b);
              const uint32_t b_lo0 = (uint32_t)bd0, b_hi = (uint32_t)(bd0 >> 32);
              if (elect_one()) {
#pragma unroll 1
                for (int j = 0; j < TPB; ++j) {
                  const int tap = tb + j;
                  const uint32_t a_tap = a_lo0 + (uint32_t)((tap / KS) * HWp + (tap % KS));
                  const uint32_t b_tap = b_lo0 + (uint32_t)j * b_tap_u;
#pragma unroll
                  for (int k16 = 0; k16 < 2; ++k16) {
                    if (k16 < kc16) {
                      const uint32_t a_k = a_tap + (uint32_t)k16 * kstep_a;
                      const uint32_t b_k = b_tap + (uint32_t)k16 * kstep_b;
#pragma unroll
                      for (int i = 0; i < OP::n; ++i) {
                        const bool first = fresh && tap == 0 && k16 == 0 && ((FirstW<MODE>::mask >> i) & 1u);
                        umma_f16_w(dcol[i], a_k + (uint32_t)OPF(OP::A, i) * a_piece_u, a_hi, b_k + boff[i], b_hi,
                                    idesc[i], first ? 0u : 1u);
                      }
                    }
                  }
                }
                umma_commit(&b_empty[sb]);
                if (tb + TPB >= T) umma_commit(&a_empty[sa]);
              }
              __syncwarp();
              if (++sb == NB) {
                sb = 0;
                pb ^= 1u;
              }
            }
          }
          if (++sa == AST) {
            sa = 0;
            pa ^= 1u;
          }
          if (++set == t.S) set = 0;
        }
        if (elect_one()) umma_commit(&acc_full[ts]);
        __syncwarp();
        if (lane == 0) {
          if (tile_it == 0) DBG(4);
          DBG(5);
        }
        if (++ts == TS) {
          ts = 0;
          pt ^= 1u;
        }
      }
      griddep_launch();  // all MMAs of this CTA are issued: the next kernel may start its prologue
    }
  } else if (warp >= kEpiWarp0) {
    // ===== epilogue =====
    const int ew = warp - kEpiWarp0;       // 0..7
    const int quarter = warp & 3;          // TMEM lane quarter of this warp
    const int half = ew >> 2;              // which half of the 16-column chunks
    const int row = quarter * 32 + lane;   // GEMM row = tile pixel
    const int py_t = row >> 3, px_t = row & 7;
    const bool want_red = (a.epi == EPI_NHWC && a.o_sum != nullptr) || a.epi == EPI_BNBWD;
    const int my_col = colsum16_col(lane);
    const int nsets = nchunks < t.S ? nchunks : t.S;
    const float osc = t.out_scale * (t.dyn_scale != nullptr ? *t.dyn_scale : 1.f);
    float gmx = 0.f;  // running max |G| written by this thread (dynamic fp16 scale of the next layers)
    int tile_it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
      const int ts = tile_it % TS;
      int rem = tile;
      const int tx = rem % tiles_x;
      rem /= tiles_x;
      const int ty = rem % tiles_y;
      const int b = rem / tiles_y;
      int oy = ty * kTH + py_t, ox = tx * kTW + px_t;
      bool valid = oy < a.Ho && ox < a.Wo;
      int Hd = a.Ho, Wd = a.Wo;
      if (a.pool || t.osub) {
        // pool: 2x2 sum (adjoint of nearest upsampling); osub: keep even positions only (a stride-2
        // convolution evaluated as a stride-1 convolution and subsampled)
        valid = valid && ((py_t | px_t) & 1) == 0;
        oy >>= 1;
        ox >>= 1;
        Hd = (Hd + 1) >> 1;
        Wd = (Wd + 1) >> 1;
      }
      const size_t pix = ((size_t)b * Hd + oy) * Wd + ox;
      // The BatchNorm-backward operands do not depend on the accumulator.  x of the consumer's input is
      // needed first (ReLU mask): it is software-pipelined one column chunk ahead — chunk 0 is fetched
      // while this tile's MMAs are still running — so its L2 round trip never sits on the chunk's path.
      const bool bnb = a.epi == EPI_BNBWD && valid;
      float xn[16];
      auto load_x = [&](int n0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) xn[i] = 0.f;
        if (bnb && n0 < N) {
          const float* xs = a.fx + pix * a.ldfx + n0;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            if (n0 + i + 3 < a.Cout) {
              const float4 f = __ldg(reinterpret_cast<const float4*>(xs + i));
              xn[i] = f.x; xn[i + 1] = f.y; xn[i + 2] = f.z; xn[i + 3] = f.w;
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (n0 + i + k < a.Cout) xn[i + k] = xs[i + k];
            }
          }
        }
      };
      load_x(half * 16);
      mbar_wait(&acc_full[ts], (uint32_t)((tile_it / TS) & 1));
      tc_fence_after();
      if (ew == 0 && lane == 0 && tile_it == 0) DBG(6);
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)ts * ts_cols;
      for (int n0 = half * 16; n0 < N; n0 += 32) {
        float xv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) xv[i] = xn[i];
        load_x(n0 + 32);  // next chunk of this warp
        // accumulator: small groups first, then the big one, over all sets
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
        for (int g = t.ngroups - 1; g >= 0; --g) {
          for (int sset = 0; sset < nsets; ++sset) {
            float w1[16];
            tmem_ld16(taddr + (uint32_t)((sset * t.ngroups + g) * N + n0), w1);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += w1[i];
          }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] *= osc;
        if (a.pool) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float s = v[i] + __shfl_xor_sync(0xffffffffu, v[i], 1);
            s += __shfl_xor_sync(0xffffffffu, s, 8);
            v[i] = s;
          }
        }
        float s1[16], s2[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) s1[i] = s2[i] = 0.f;
        if (valid) {
          if (a.epi == EPI_NHWC) {
            float* dst = a.y + pix * a.ldy + a.coff + n0;
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              if (n0 + i + 3 < a.Cout) {
                *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (n0 + i + k < a.Cout) dst[i + k] = v[i + k];
              }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const bool in = n0 + i < a.Cout;
              s1[i] = in ? v[i] : 0.f;
              s2[i] = in ? v[i] * v[i] : 0.f;
            }
          } else if (a.epi == EPI_NCHW) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (n0 + i < a.Cout) a.y[(((size_t)b * a.Cout + n0 + i) * Hd + oy) * Wd + ox] = v[i];
          } else {  // EPI_BNBWD
            float* gp = a.G + pix * a.ldG + n0;
            float o[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float4 c = *reinterpret_cast<const float4*>(ep_s + 4 * (n0 + i));
              const float z = fmaf(xv[i], c.x, c.y);
              const float dz = z > 0.f ? v[i] : 0.f;
              const float xh = fmaf(xv[i], c.z, c.w);
              s1[i] = dz;
              s2[i] = dz * xh;
              o[i] = c.x * dz;  // this consumer's contribution to dL/dx
              gmx = fmaxf(gmx, fabsf(o[i]));
            }
            // the first consumer to run stores, the others add with fire-and-forget vector reductions:
            // no read of the gradient row, nothing to wait for (one add per element and kernel, kernels
            // are stream-ordered: the sum is as deterministic as a read-modify-write)
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              if (n0 + i + 3 < a.Cout) {
                if (a.g_accum) red_add_v4(gp + i, o[i], o[i + 1], o[i + 2], o[i + 3]);
                else *reinterpret_cast<float4*>(gp + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (n0 + i + k < a.Cout) gp[i + k] = a.g_accum ? gp[i + k] + o[i + k] : o[i + k];
              }
            }
          }
        }
        if (want_red) {
          const float u = colsum16(s1, lane), w = colsum16(s2, lane);
          if ((lane & 1) == 0) {  // lanes l and l^1 hold the same column
            float* r = red_s + ((size_t)ew * N + n0 + my_col) * 2;
            r[0] += u;
            r[1] += w;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[ts]);
      if (ew == 0 && lane == 0) { if (tile_it == 0) DBG(7); DBG(8); }
    }
    if (a.epi == EPI_BNBWD && a.gmax != nullptr) {
      const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(gmx));
      if (lane == 0 && m != 0u) atomicMax(a.gmax, m);
    }
    if (want_red) {
      named_bar_sync(1, kEpiWarps * 32);
      for (int n = threadIdx.x - kEpiWarp0 * 32; n < a.Cout; n += kEpiWarps * 32) {
        double u = 0.0, w = 0.0;
#pragma unroll
        for (int q = 0; q < kEpiWarps; ++q) {
          u += (double)red_s[((size_t)q * N + n) * 2 + 0];
          w += (double)red_s[((size_t)q * N + n) * 2 + 1];
        }
        if (a.epi == EPI_NHWC) {
          atomicAdd(a.o_sum + n, u);
          atomicAdd(a.o_sumsq + n, w);
        } else {
          atomicAdd(a.bsum + n, u);
          atomicAdd(a.bsum + a.Cout + n, w);
        }
      }
    }
  }
  if (threadIdx.x == kEpiWarp0 * 32) DBG(9);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) DBG(10);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
#undef DBG
}

// ---------------------------------------------------------------------------------------
// filter packing: OIHW fp32 -> [chunk][tap][k-octet][piece][n][8] fp16 pieces of w * 2^kWScaleLog2
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_tc2_kernel(const Tc2PackDesc* tab) {
  griddep_wait();
  const Tc2PackDesc d = tab[blockIdx.y];
  // dxn: one "tap" per filter ROW, the columns live in n (1) / in k (2); 3: every tap lives in k (one 1x1 GEMM)
  const int T = d.dxn == 3 ? 1 : (d.dxn ? d.KS : d.KS * d.KS);
  const int koct = d.KC >> 3, KK = d.KS * d.KS;
  const size_t per_tap = (size_t)koct * d.N * 8;  // elements of ONE piece of one (chunk, tap)
  const int rows = d.nchunks * T * koct * d.N;    // 16-byte rows: 8 consecutive k of one (chunk, tap, k octet, n)
  const float wscale = (float)(1 << kWScaleLog2);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += gridDim.x * blockDim.x) {
    const int n = i % d.N;
    int r = i / d.N;
    const int ko = r % koct;
    r /= koct;                                  // = step
    const int chunk = r / T, tap = r - chunk * T;
    const int k0 = chunk * d.KC + ko * 8;
    float v[8];
#pragma unroll
    for (int k8 = 0; k8 < 8; ++k8) {
      const int k = k0 + k8;
      float x = 0.f;
      if (d.dxn == 3) {
        // data gradient over the (tap, co)-expanded dY: k = tap * Cout + co, n = ci; W[co, ci, tap] (no flip: the
        // expansion already reads dY at q - tap + pad)
        const int tp = k / d.Cout, co = k - tp * d.Cout;
        if (n < d.Cin && tp < KK) x = d.w[((size_t)co * d.Cin + n) * KK + tp];
      } else if (d.dxn == 2) {
        // data gradient: k = (jx, co), n = ci, tap = jy; flipped filter W[co, ci, 2-jy, 2-jx]
        const int jx = k >> 4, co = k & 15;
        if (n < d.Cin && co < d.Cout && jx < d.KS)
          x = d.w[((size_t)co * d.Cin + n) * KK + (d.KS - 1 - tap) * d.KS + (d.KS - 1 - jx)];
      } else if (d.dxn) {
        const int kx = n / d.CoP, co = n - kx * d.CoP;
        if (co < d.Cout && k < d.Cin) x = d.w[((size_t)co * d.Cin + k) * KK + tap * d.KS + kx];
      } else if (!d.transpose) {
        if (n < d.Cout && k < d.Cin) x = d.w[((size_t)n * d.Cin + k) * T + tap];
      } else {
        if (n < d.Cin && k < d.Cout) x = d.w[((size_t)k * d.Cin + n) * T + (T - 1 - tap)];
      }
      v[k8] = x * wscale;
    }
    uint32_t p1[4], p2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (d.lowp == LOWP_BF16) {
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p1[j]) : "f"(v[2 * j + 1]), "f"(v[2 * j]));
        p2[j] = 0u;
      } else {
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(p1[j]) : "f"(v[2 * j + 1]), "f"(v[2 * j]));
        const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&p1[j]));
        p2[j] = 0u;
        if (!d.lowp)
          asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(p2[j]) : "f"(v[2 * j + 1] - fl.y), "f"(v[2 * j] - fl.x));
      }
    }
    op16* base = d.dst + (size_t)r * per_tap * kPieces + (size_t)ko * kPieces * d.N * 8 + (size_t)n * 8;
    *reinterpret_cast<uint4*>(base) = make_uint4(p1[0], p1[1], p1[2], p1[3]);
    *reinterpret_cast<uint4*>(base + (size_t)d.N * 8) = make_uint4(p2[0], p2[1], p2[2], p2[3]);
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn get_encode2() {
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

size_t tc2_smem(int KS, int N, int KC, int AST, int NB, int TPB) {
  const int HP = (kTH + KS - 1) * (kTW + KS - 1);
  const size_t hdr = (256 + sizeof(float) * (size_t)(4 + 2 * kEpiWarps) * N + 127) & ~(size_t)127;
  return hdr + (size_t)AST * kPieces * (KC / 8) * HP * 16 + (size_t)NB * TPB * (KC / 8) * kPieces * N * 16;
}

}  // namespace

void tc2_plan(int KS, int Cin_k, int N, Tc2Plan* p, int lowp) {
  const int T = KS * KS;
  int KC = Cin_k >= 32 ? 32 : 16;
  p->KC = KC;
  p->nchunks = (Cin_k + KC - 1) / KC;
  p->ngroups = lowp ? 1 : 2;  // G0 = a1*w1, G1 = a1*w2 + a2*w1
  int S = 256 / (p->ngroups * N), TS = 2;
  if (S < 1) {
    S = 512 / (p->ngroups * N);
    TS = 1;
  }
  {
    const char* e = getenv("PDES_TC2_TS1");  // experiment: one accumulator stage, more K-spreading sets
    if (e && e[0] == '1') {
      S = 512 / (p->ngroups * N);
      TS = 1;
    }
  }
  int smax = 4;
  {
    const char* e = getenv("PDES_TC2_SMAX");
    if (e) smax = atoi(e);
  }
  if (S > p->nchunks) S = p->nchunks;
  if (S > smax) S = smax;
  if (S < 1) S = 1;
  p->S = S;
  p->TS = TS;
  const size_t tap_bytes = (size_t)(KC / 8) * kPieces * N * 16;
  int TPB = 1;
  for (int d = 1; d <= T; ++d)
    if (T % d == 0 && d * tap_bytes <= (d == T ? 56 : 32) * 1024) TPB = d;
  int NB = TPB * tap_bytes >= 16 * 1024 ? 3 : 4;
  int AST = 3;
  while (tc2_smem(KS, N, KC, AST, NB, TPB) > 224 * 1024 && (NB > 2 || AST > 2)) {
    if (NB > 2) --NB;
    else --AST;
  }
  p->AST = AST;
  p->NB = NB;
  p->TPB = TPB;
  p->smem = tc2_smem(KS, N, KC, AST, NB, TPB);
  p->pack_elems = (size_t)p->nchunks * T * (KC / 8) * kPieces * N * 8;
}

bool tc2_supported(int KS, int stride, int Cin_k, int N) {
  if (!(KS == 1 || KS == 3 || KS == 5 || KS == 7) || stride != 1) return false;
  if (N < 16 || N > 256 || (N & 15)) return false;
  Tc2Plan p;
  tc2_plan(KS, Cin_k, N, &p);
  return p.smem <= 225 * 1024 && (uint32_t)(p.S * p.ngroups * N * p.TS) <= 512u;
}

int launch_conv_tc2(const Tc2Args& t, const op16* planes, int Hv, int Wv, int Cin_k,
                    cudaStream_t st) {
  const ConvArgs& a = t.c;
  PDES_REQUIRE(a.KS == 1 || a.KS == 3 || a.KS == 5 || a.KS == 7, PDES_ERR_UNSUPPORTED, "conv_tc2: kernel size %d", a.KS);
  if (a.epi == EPI_NHWC)
    PDES_REQUIRE(((a.ldy | a.coff) & 3) == 0 && ((uintptr_t)a.y & 15u) == 0, PDES_ERR_INVALID,
                 "conv_tc2: output slice must be 16-byte aligned");
  if (a.epi == EPI_BNBWD)
    PDES_REQUIRE(((a.ldfx | a.ldG) & 3) == 0, PDES_ERR_INVALID, "conv_tc2: gradient buffers misaligned");
  PDES_REQUIRE(!a.pool || ((a.Ho | a.Wo) & 1) == 0, PDES_ERR_INVALID, "conv_tc2: pool needs even size");
  PDES_REQUIRE(!(a.pool && t.osub), PDES_ERR_INVALID, "conv_tc2: pool and subsample are exclusive");
  EncodeFn enc = get_encode2();
  PDES_REQUIRE(enc != nullptr, PDES_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const int Cp = (Cin_k + 7) & ~7;
  CUtensorMap tm;
  {
    // planes [2*B][Hv][Cp/8][Wv][8] viewed as (x*8+c8, y, octet, piece*B+b)
    const cuuint64_t oct = (cuuint64_t)(Cp / 8);
    const cuuint64_t gdim[4] = {(cuuint64_t)Wv * 8, (cuuint64_t)Hv, oct, (cuuint64_t)kPieces * a.B};
    const cuuint64_t gstr[3] = {oct * Wv * 16, (cuuint64_t)Wv * 16, (cuuint64_t)Hv * oct * Wv * 16};
    const cuuint32_t box[4] = {(cuuint32_t)(kTW + a.KS - 1) * 8, (cuuint32_t)(kTH + a.KS - 1), (cuuint32_t)(t.KC / 8), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<op16*>(planes), gdim,
                           gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PDES_REQUIRE(r == CUDA_SUCCESS, PDES_ERR_CUDA, "cuTensorMapEncodeTiled failed with code %d", (int)r);
  }
  const size_t smem = tc2_smem(a.KS, t.N, t.KC, t.AST, t.NB, t.TPB);
  PDES_REQUIRE(smem <= 227 * 1024, PDES_ERR_UNSUPPORTED, "conv_tc2: needs %zu bytes of shared memory", smem);
  const int tiles = ((a.Wo + kTW - 1) / kTW) * ((a.Ho + kTH - 1) / kTH) * a.B;
  int grid = sm_count();
  if (grid > tiles) grid = tiles;
  const int mode = t.lowp ? 2 : (2 * t.N <= 256 ? 0 : 1);
  PDES_REQUIRE(!t.lowp || t.ngroups == 1, PDES_ERR_INVALID, "conv_tc2: one-piece modes use one accumulator group");
#define PDES_TC2_LAUNCH(KSV, MODEV)                                                                          \
  {                                                                                                          \
    PDES_ENSURE_SMEM((conv_tc2_kernel<KSV, MODEV>), smem);                                                   \
    PDES_CUDA(launch_pdl(conv_tc2_kernel<KSV, MODEV>, dim3(grid), dim3(kThreads), smem, st, tm, t));         \
  }
#define PDES_TC2_MODES(KSV)                        \
  {                                                \
    if (mode == 0) PDES_TC2_LAUNCH(KSV, 0)         \
    else if (mode == 1) PDES_TC2_LAUNCH(KSV, 1)    \
    else PDES_TC2_LAUNCH(KSV, 2)                   \
  }
  if (a.KS == 3) PDES_TC2_MODES(3)
  else if (a.KS == 1) PDES_TC2_MODES(1)
  else if (a.KS == 5) PDES_TC2_MODES(5)
  else PDES_TC2_MODES(7)
#undef PDES_TC2_MODES
#undef PDES_TC2_LAUNCH
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

int launch_pack_tc2(const Tc2PackDesc* dev_table, int n, size_t max_elems, cudaStream_t st) {
  if (n == 0) return PDES_OK;
  int bx = (int)((max_elems / kPieces / 8 + 255) / 256);   // one thread per 16-byte row
  if (bx > 32) bx = 32;
  if (bx < 1) bx = 1;
  PDES_CUDA(launch_pdl(pack_tc2_kernel, dim3(bx, n), dim3(256), 0, st, dev_table));
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

}  // namespace pdes

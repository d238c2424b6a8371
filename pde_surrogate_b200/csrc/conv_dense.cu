// conv_dense.cu — fused forward of a thin 3x3 layer (a DenseNet layer: BatchNorm -> ReLU -> conv3x3 with a
// handful of output channels; reference models/codec.py:65-69), ONE kernel instead of operand-split +
// convolution:
//
//   * the operand never makes a round trip through global memory: producer warps read the fp32 NHWC rows
//     of the block buffer, apply BatchNorm + ReLU, split every value into two fp16 pieces (conv_tc.cuh)
//     and store them straight into the shared-memory image the tensor core reads
//     ([channel octet][pixel][16 B] = canonical K-major, no swizzle).  In training they also emit the
//     pieces once to global memory for the weight-gradient kernel (planes [piece][b][y][octet][x][8]).
//   * "dx in N": a pixel tile is 128/W full image rows.  An M=128, N<=48 tcgen05.mma is paced by the
//     shared-memory fetch of its A operand (~51 clk whatever N), so reading A once per tap is what made
//     thin layers slow.  Here the three horizontal taps are folded into GEMM-N,
//         D[(y, x'), (kx, co)] = sum_ky sum_ci  a[(y + ky - 1, x'), ci] * w[co, ci, ky, kx],
//     three A reads per k-step instead of nine (ky = a descriptor offset of W pixels into the (TR+2)-row
//     tile: a multiple of eight rows), and the epilogue finishes the convolution with two warp shuffles,
//         y[(y, x), co] = D[(y, x-1), (0, co)] + D[(y, x), (1, co)] + D[(y, x+1), (2, co)],
//     because a warp's 32 TMEM lanes are 32 consecutive pixels of full rows (W in {8, 16, 32}).
//   * the whole packed filter of the layer (<= 130 KB) is loaded ONCE per CTA by bulk TMA and stays
//     resident; activations stream through a 3-stage ring.
//
// fp32 parity as in conv_tc2.cu: a1 x [w1 | w2] (one N = 2*NQ instruction, leading and cross products in
// separate TMEM column groups) + a2 x w1.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include "conv.cuh"
#include "conv_tc.cuh"
#include "tc_common.cuh"

namespace pdes {
namespace {

using namespace tc;

constexpr int kEpiWarps = 4;                 // warps 0..3: TMEM lane quarter = warp index
constexpr int kMmaWarp = 4;                  // filter loads, TMEM allocation, MMA issue
constexpr int kProdWarp0 = 5, kProdWarps = 8;
constexpr int kThreads = 32 * (kProdWarp0 + kProdWarps);   // 416
constexpr int kKC = 32;                      // channels per activation stage
constexpr int kStages = 3;
constexpr int kMaxChunks = 8;
constexpr int kTS = 2;                       // accumulator stages in TMEM

__device__ __forceinline__ uint32_t make_idesc_f16(int M, int N) {  // D fp32, A/B fp16, K-major
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
// two floats -> packed fp16x2 (round to nearest, saturating), lo in bits [0,16)
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v) {
  return __half22float2(*reinterpret_cast<const __half2*>(&v));
}

struct Geo {
  int W, H, TR, HP, tiles_per_img, n_tiles;
};

__global__ void __launch_bounds__(kThreads, 1) conv_dense_fwd_kernel(DenseFwdArgs a) {
  const int W = a.W, H = a.H;
  const int TR = 128 / W;                 // output rows per tile (GEMM-M = 128 pixels)
  const int HP = (TR + 2) * W;            // pixels of the staged tile (one halo row above and below)
  const int tiles_per_img = (H + TR - 1) / TR;
  const int n_tiles = tiles_per_img * a.B;
  const int NQ = 3 * a.CoP;               // GEMM-N of one filter piece: (kx, co)
  const int nchunks = (a.Cin + kKC - 1) / kKC;
  const int n16_total = (a.Cin + 15) >> 4;
  const uint32_t a_piece_bytes = (uint32_t)(kKC / 8) * HP * 16u;
  const uint32_t a_stage_bytes = 2u * a_piece_bytes;
  const uint32_t b_ky_bytes = (uint32_t)(kKC / 8) * 2u * NQ * 16u;   // [k-octet][piece][n][16 B]
  const uint32_t b_chunk_bytes = 3u * b_ky_bytes;

  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem);   // [kStages]
  uint64_t* a_empty = a_full + kStages;                   // [kStages]
  uint64_t* b_full = a_empty + kStages;                   // [kMaxChunks]
  uint64_t* acc_full = b_full + kMaxChunks;               // [kTS]
  uint64_t* acc_empty = acc_full + kTS;                   // [kTS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kTS);
  float* sc_s = reinterpret_cast<float*>(smem + 256);     // [nchunks * 32] scale (x 2^kActScaleLog2)
  float* sh_s = sc_s + kMaxChunks * kKC;                  // shift
  float* red_s = sh_s + kMaxChunks * kKC;                 // [kEpiWarps][16][2]
  unsigned char* A_s = smem + 256 + sizeof(float) * (2 * kMaxChunks * kKC + kEpiWarps * 32);
  unsigned char* B_s = A_s + (size_t)kStages * a_stage_bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ts_cols = 2u * NQ;       // G0 = a1*w1 | G1 = a1*w2 + a2*w1
  uint32_t tmem_cols = 32;
  while (tmem_cols < ts_cols * kTS) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&a_full[i], kProdWarps);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < kMaxChunks; ++i) mbar_init(&b_full[i], 1);
    for (int i = 0; i < kTS; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  griddep_wait();  // x, the batch statistics and the packed filter come from earlier kernels of the step
  if (warp >= kProdWarp0) {
    // BatchNorm constants of this layer, the activation scale folded in: relu(s*x+h)*2^k = relu(2^k s x + 2^k h)
    const float mul = (float)(1 << kActScaleLog2);
    for (int c = threadIdx.x - kProdWarp0 * 32; c < nchunks * kKC; c += kProdWarps * 32) {
      float s = 0.f, h = 0.f, m, is;
      if (c < a.Cin) {
        if (a.pro) {
          bn_consts_tc(a.bn, c, s, h, m, is);
          s *= mul;
          h *= mul;
        } else {
          s = mul;
        }
      }
      sc_s[c] = s;
      sh_s[c] = h;
    }
  } else if (warp < kEpiWarps) {
    for (int i = lane; i < 32; i += 32) red_s[warp * 32 + i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= kProdWarp0) {
    // ===== producers: fp32 rows -> BN + ReLU -> two fp16 pieces -> UMMA shared-memory image =====
    // item = (8-pixel group g, channel octet q, pixel e): consecutive lanes walk e, then q: a warp reads
    // eight full 128-byte lines (8 pixels x 32 channels) and a quarter-warp stores 128 contiguous bytes.
    const int pt = threadIdx.x - kProdWarp0 * 32;
    const int n_items = HP * (kKC / 8);
    constexpr int kMaxIt = 3;                  // HP*4 <= 768 = 3 * 256 (W = 32: 6 rows x 32)
    const bool relu = a.pro != 0;
    const size_t plane_elems = (size_t)a.B * H * W * a.Cp;
    const int oct_total = a.Cp >> 3;
    const bool vec_ok = (a.ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(a.x) & 15u) == 0;
    float4 v0[kMaxIt], v1[kMaxIt];
    auto issue_loads = [&](int tile, int ch) {
      const int b = tile / tiles_per_img, r0 = (tile - b * tiles_per_img) * TR;
#pragma unroll
      for (int j = 0; j < kMaxIt; ++j) {
        const int it = pt + j * (kProdWarps * 32);
        v0[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        v1[j] = v0[j];
        if (it < n_items) {
          const int e = it & 7, q = (it >> 3) & 3, g = it >> 5;
          const int p = g * 8 + e;
          const int row = r0 - 1 + p / W, col = p % W;
          const int c = ch * kKC + q * 8;
          if (row >= 0 && row < H && c < a.Cin) {
            const float* src = a.x + ((size_t)(b * H + row) * W + col) * a.ldx + c;
            if (vec_ok && c + 7 < a.Cin) {
              v0[j] = __ldg(reinterpret_cast<const float4*>(src));
              v1[j] = __ldg(reinterpret_cast<const float4*>(src + 4));
            } else {
              float t[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) t[k] = (c + k < a.Cin) ? __ldg(src + k) : 0.f;
              v0[j] = make_float4(t[0], t[1], t[2], t[3]);
              v1[j] = make_float4(t[4], t[5], t[6], t[7]);
            }
          }
        }
      }
    };
    int q_it = 0;  // global (tile, chunk) counter of this CTA
    int tile = blockIdx.x, ch = 0;
    if (tile < n_tiles) issue_loads(tile, 0);
    while (tile < n_tiles) {
      const int s = q_it % kStages;
      const int b = tile / tiles_per_img, r0 = (tile - b * tiles_per_img) * TR;
      if (lane == 0) mbar_wait(&a_empty[s], (uint32_t)(((q_it / kStages) & 1) ^ 1));
      __syncwarp();
      unsigned char* st = A_s + (size_t)s * a_stage_bytes;
      // convert what is in registers
      uint4 h1[kMaxIt], h2[kMaxIt];
#pragma unroll
      for (int j = 0; j < kMaxIt; ++j) {
        const int it = pt + j * (kProdWarps * 32);
        if (it < n_items) {
          const int q = (it >> 3) & 3;
          // rows outside the image are the convolution's zero padding of the ACTIVATION (after BN + ReLU)
          const int prow_i = ((it >> 5) * 8 + (it & 7)) / W;
          const bool inside = (r0 - 1 + prow_i) >= 0 && (r0 - 1 + prow_i) < H;
          const float4 s0 = *reinterpret_cast<const float4*>(sc_s + ch * kKC + q * 8);
          const float4 s1 = *reinterpret_cast<const float4*>(sc_s + ch * kKC + q * 8 + 4);
          const float4 o0 = *reinterpret_cast<const float4*>(sh_s + ch * kKC + q * 8);
          const float4 o1 = *reinterpret_cast<const float4*>(sh_s + ch * kKC + q * 8 + 4);
          float t[8] = {fmaf(v0[j].x, s0.x, o0.x), fmaf(v0[j].y, s0.y, o0.y), fmaf(v0[j].z, s0.z, o0.z),
                        fmaf(v0[j].w, s0.w, o0.w), fmaf(v1[j].x, s1.x, o1.x), fmaf(v1[j].y, s1.y, o1.y),
                        fmaf(v1[j].z, s1.z, o1.z), fmaf(v1[j].w, s1.w, o1.w)};
          if (relu) {
#pragma unroll
            for (int k = 0; k < 8; ++k) t[k] = fmaxf(t[k], 0.f);
          }
          if (!inside) {
#pragma unroll
            for (int k = 0; k < 8; ++k) t[k] = 0.f;
          }
          uint32_t p1[4], p2[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            p1[k] = pack_h2(t[2 * k], t[2 * k + 1]);
            const float2 f = unpack_h2(p1[k]);
            p2[k] = pack_h2(t[2 * k] - f.x, t[2 * k + 1] - f.y);
          }
          h1[j] = make_uint4(p1[0], p1[1], p1[2], p1[3]);
          h2[j] = make_uint4(p2[0], p2[1], p2[2], p2[3]);
        }
      }
      // next (tile, chunk): its loads fly while this chunk is stored and the MMAs of earlier stages run
      int ntile = tile, nch = ch + 1;
      if (nch == nchunks) {
        nch = 0;
        ntile = tile + gridDim.x;
      }
      const int cur_ch = ch;
      if (ntile < n_tiles) issue_loads(ntile, nch);
#pragma unroll
      for (int j = 0; j < kMaxIt; ++j) {
        const int it = pt + j * (kProdWarps * 32);
        if (it < n_items) {
          const int e = it & 7, q = (it >> 3) & 3, g = it >> 5;
          const int p = g * 8 + e;
          unsigned char* dst = st + (size_t)q * HP * 16 + (size_t)p * 16;
          *reinterpret_cast<uint4*>(dst) = h1[j];
          *reinterpret_cast<uint4*>(dst + a_piece_bytes) = h2[j];
          if (a.planes != nullptr) {
            const int prow = p / W, col = p - prow * W;
            const int row = r0 - 1 + prow;
            const int oq = cur_ch * (kKC / 8) + q;
            if (prow >= 1 && prow <= TR && row < H && oq < oct_total) {
              op16* pd = a.planes + ((((size_t)b * H + row) * oct_total + oq) * W + col) * 8;
              *reinterpret_cast<uint4*>(pd) = h1[j];
              *reinterpret_cast<uint4*>(pd + plane_elems) = h2[j];
            }
          }
        }
      }
      fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[s]);
      ++q_it;
      tile = ntile;
      ch = nch;
    }
  } else if (warp == kMmaWarp) {
    // ===== resident filter (bulk TMA, once per CTA) + MMA issue =====
    if (lane == 0 && blockIdx.x < n_tiles) {
      for (int c = 0; c < nchunks; ++c) {
        mbar_arrive_expect_tx(&b_full[c], b_chunk_bytes);
        tma_load_1d(B_s + (size_t)c * b_chunk_bytes,
                    reinterpret_cast<const unsigned char*>(a.wpk) + (size_t)c * b_chunk_bytes, b_chunk_bytes,
                    &b_full[c]);
      }
    }
    const uint32_t idesc2 = make_idesc_f16(128, 2 * NQ), idesc1 = make_idesc_f16(128, NQ);
    const uint32_t lbo_a = (uint32_t)HP * 16u, sbo_a = 128u;
    const uint32_t lbo_b = 2u * NQ * 16u, sbo_b = 128u;
    const uint32_t kstep_a = (2u * lbo_a) >> 4, kstep_b = (2u * lbo_b) >> 4;
    const uint32_t a_piece_u = a_piece_bytes >> 4, b_ky_u = b_ky_bytes >> 4;
    const uint32_t ky_a = (uint32_t)W;   // W pixels * 16 B, in 16-byte units
    int sa = 0, ts = 0, tile_it = 0;
    uint32_t pa = 0, pt = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
      mbar_wait(&acc_empty[ts], pt ^ 1u);
      tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)ts * ts_cols;
      for (int ch = 0; ch < nchunks; ++ch) {
        if (tile_it == 0) mbar_wait(&b_full[ch], 0);
        mbar_wait(&a_full[sa], pa);
        tc_fence_after();
        const uint64_t ad0 = make_desc(smem_u32(A_s + (size_t)sa * a_stage_bytes), lbo_a, sbo_a);
        const uint64_t bd0 = make_desc(smem_u32(B_s + (size_t)ch * b_chunk_bytes), lbo_b, sbo_b);
        const uint32_t a_lo0 = (uint32_t)ad0, a_hi = (uint32_t)(ad0 >> 32);
        const uint32_t b_lo0 = (uint32_t)bd0, b_hi = (uint32_t)(bd0 >> 32);
        const int n16 = (n16_total - 2 * ch) < 2 ? (n16_total - 2 * ch) : 2;
        if (elect_one()) {
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
            for (int k16 = 0; k16 < 2; ++k16) {
              if (k16 < n16) {
                const uint32_t a_k = a_lo0 + (uint32_t)ky * ky_a + (uint32_t)k16 * kstep_a;
                const uint32_t b_k = b_lo0 + (uint32_t)ky * b_ky_u + (uint32_t)k16 * kstep_b;
                const bool first = ch == 0 && ky == 0 && k16 == 0;
                umma_f16_w(d0, a_k, a_hi, b_k, b_hi, idesc2, first ? 0u : 1u);                       // a1 x [w1|w2]
                umma_f16_w(d0 + (uint32_t)NQ, a_k + a_piece_u, a_hi, b_k, b_hi, idesc1, 1u);          // a2 x w1 -> G1
              }
            }
          }
          umma_commit(&a_empty[sa]);
          if (ch == nchunks - 1) umma_commit(&acc_full[ts]);
        }
        __syncwarp();
        if (++sa == kStages) {
          sa = 0;
          pa ^= 1u;
        }
      }
      if (++ts == kTS) {
        ts = 0;
        pt ^= 1u;
      }
    }
    griddep_launch();
  } else {
    // ===== epilogue: finish the convolution along x with shuffles, store the slice, batch statistics =====
    const int q = warp;                               // TMEM lane quarter
    const int m = q * 32 + lane;                      // tile pixel
    const int prow = m / W, px = m - prow * W;
    const bool has_l = px > 0, has_r = px < W - 1;
    const bool want_red = a.o_sum != nullptr;
    const int my_col = colsum16_col(lane);
    const float osc = a.out_scale;
    const int CoP = a.CoP;
    int tile_it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
      const int ts = tile_it % kTS;
      const int b = tile / tiles_per_img, r0 = (tile - b * tiles_per_img) * TR;
      const int row = r0 + prow;
      const bool valid = row < H;
      mbar_wait(&acc_full[ts], (uint32_t)((tile_it / kTS) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)ts * ts_cols;
      for (int c0 = 0; c0 < CoP; c0 += 16) {
        float o[16];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          float g1[16], g0[16];
          tmem_ld16(taddr + (uint32_t)(NQ + kx * CoP + c0), g1);   // cross terms first (small), then the leading ones
          tmem_ld16(taddr + (uint32_t)(kx * CoP + c0), g0);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float v = g1[i] + g0[i];
            if (kx == 0) {
              const float t = __shfl_up_sync(0xffffffffu, v, 1);      // D[(y, x-1), kx=0]
              o[i] = has_l ? t : 0.f;
            } else if (kx == 1) {
              o[i] += v;
            } else {
              const float t = __shfl_down_sync(0xffffffffu, v, 1);    // D[(y, x+1), kx=2]
              o[i] += has_r ? t : 0.f;
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] *= osc;
        float s1[16], s2[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const bool in = valid && (c0 + i < a.Cout);
          s1[i] = in ? o[i] : 0.f;
          s2[i] = in ? o[i] * o[i] : 0.f;
        }
        if (valid) {
          float* dst = a.y + ((size_t)(b * H + row) * W + px) * a.ldy + a.coff + c0;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            if (c0 + i + 3 < a.Cout) {
              *reinterpret_cast<float4*>(dst + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (c0 + i + k < a.Cout) dst[i + k] = o[i + k];
            }
          }
        }
        if (want_red) {
          const float u = colsum16(s1, lane), w = colsum16(s2, lane);
          if ((lane & 1) == 0 && c0 == 0) {   // (CoP == 16: one column block)
            red_s[(warp * 16 + my_col) * 2 + 0] += u;
            red_s[(warp * 16 + my_col) * 2 + 1] += w;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[ts]);
    }
    if (want_red) {
      named_bar_sync(1, kEpiWarps * 32);
      for (int n = threadIdx.x; n < a.Cout && n < 16; n += kEpiWarps * 32) {
        double u = 0.0, w = 0.0;
#pragma unroll
        for (int k = 0; k < kEpiWarps; ++k) {
          u += (double)red_s[(k * 16 + n) * 2 + 0];
          w += (double)red_s[(k * 16 + n) * 2 + 1];
        }
        atomicAdd(a.o_sum + n, u);
        atomicAdd(a.o_sumsq + n, w);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

size_t dense_smem(int W, int Cin, int CoP) {
  const int TR = 128 / W, HP = (TR + 2) * W;
  const int nchunks = (Cin + kKC - 1) / kKC;
  return 256 + sizeof(float) * (2 * kMaxChunks * kKC + kEpiWarps * 32) +
         (size_t)kStages * 2 * (kKC / 8) * HP * 16 + (size_t)nchunks * 3 * (kKC / 8) * 2 * (3 * CoP) * 16;
}

}  // namespace

bool dense_fwd_supported(int KS, int stride, int pad, int up, int Cin, int Cout, int H, int W) {
  if (KS != 3 || stride != 1 || pad != 1 || up) return false;
  if (!(W == 8 || W == 16 || W == 32) || H < 1) return false;
  if (Cout < 1 || Cout > 16) return false;
  if (Cin < 1 || (Cin + kKC - 1) / kKC > kMaxChunks) return false;
  return dense_smem(W, Cin, 16) <= 227 * 1024;
}

size_t dense_pack_elems(int Cin, int CoP) {
  const int nchunks = (Cin + kKC - 1) / kKC;
  return (size_t)nchunks * 3 * (kKC / 8) * 2 * (3 * CoP) * 8;
}

int launch_conv_dense_fwd(const DenseFwdArgs& a, cudaStream_t st) {
  PDES_REQUIRE(dense_fwd_supported(3, 1, 1, 0, a.Cin, a.Cout, a.H, a.W), PDES_ERR_UNSUPPORTED,
               "conv_dense: unsupported shape (Cin %d, Cout %d, %dx%d)", a.Cin, a.Cout, a.H, a.W);
  PDES_REQUIRE(a.CoP == 16, PDES_ERR_INVALID, "conv_dense: CoP must be 16");
  PDES_REQUIRE(((a.ldy | a.coff) & 3) == 0 && ((uintptr_t)a.y & 15u) == 0, PDES_ERR_INVALID,
               "conv_dense: output slice must be 16-byte aligned");
  PDES_REQUIRE(a.planes == nullptr || (a.Cp % 8 == 0 && a.Cp >= a.Cin), PDES_ERR_INVALID,
               "conv_dense: padded plane channels %d invalid", a.Cp);
  const size_t smem = dense_smem(a.W, a.Cin, a.CoP);
  PDES_ENSURE_SMEM(conv_dense_fwd_kernel, smem);
  const int TR = 128 / a.W;
  const int tiles = ((a.H + TR - 1) / TR) * a.B;
  int grid = sm_count();
  if (grid > tiles) grid = tiles;
  PDES_CUDA(launch_pdl(conv_dense_fwd_kernel, dim3(grid), dim3(kThreads), smem, st, a));
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

}  // namespace pdes

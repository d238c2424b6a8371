// conv_dense.cu — fused forward of a thin 3x3 layer (a DenseNet layer: BatchNorm -> ReLU -> conv3x3 with a
// handful of output channels; reference models/codec.py:65-69), ONE kernel instead of operand-split +
// convolution:
//
//   * the operand never makes a round trip through global memory: 3-D TMA boxes (32 channels x tile pixels,
//     128-byte swizzle, zero fill outside the image / beyond Cin) bring the raw fp32 NHWC rows of the block
//     buffer into a shared-memory ring two to three stages ahead; converter warps apply BatchNorm + ReLU,
//     split every value into two fp16 pieces (conv_tc.cuh) and store them straight into the shared-memory
//     image the tensor core reads ([channel octet][pixel][16 B] = canonical K-major, no swizzle).  In
//     training they also emit the pieces once to global memory for the weight-gradient kernel (planes
//     [piece][b][y][octet][x][8]).
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
#include <cuda_bf16.h>
#include <stdlib.h>
#include "conv.cuh"
#include "conv_tc.cuh"
#include "tc_common.cuh"

namespace pdes {
namespace {

using namespace tc;

constexpr int kEpiWarps = 4;                 // warps 0..3: TMEM lane quarter = warp index
constexpr int kMmaWarp = 4;                  // filter loads, TMEM allocation, MMA issue
constexpr int kTmaWarp = 5;                  // raw fp32 tile loads (one lane)
constexpr int kProdWarp0 = 6, kProdWarps = 12; // converters: 384 threads = exactly 2 items each for a 6 x 32 tile
constexpr int kThreads = 32 * (kProdWarp0 + kProdWarps);   // 576
constexpr int kKC = 32;                      // channels per activation stage (32 fp32 = one 128-byte swizzle row)
constexpr int kMaxStages = 4;                // fp16 operand stages (as many as fit: hides the MMA -> converter handshake)
constexpr int kMaxRaw = 3;                   // raw fp32 stages (2 when the resident filter leaves no room for 3)
constexpr int kMaxChunks = 8;
constexpr int kTS = 2;                       // accumulator stages in TMEM
constexpr int kHdrBytes = 3072;              // barriers + BatchNorm constants + reduction scratch (1024-aligned)

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

struct Geo {
  int W, H, TR, HP, tiles_per_img, n_tiles;
};

__global__ void __launch_bounds__(kThreads, 1)
conv_dense_fwd_kernel(const __grid_constant__ CUtensorMap tmX, DenseFwdArgs a, int RST, int AST) {
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

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // the 128-byte swizzle pattern of the raw stages is a function of the shared-memory address: 1024-byte base
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem);   // [kMaxStages]
  uint64_t* a_empty = a_full + kMaxStages;                // [kMaxStages]
  uint64_t* b_full = a_empty + kMaxStages;                // [kMaxChunks]
  uint64_t* acc_full = b_full + kMaxChunks;               // [kTS]
  uint64_t* acc_empty = acc_full + kTS;                   // [kTS]
  uint64_t* raw_full = acc_empty + kTS;                   // [kMaxRaw]
  uint64_t* raw_empty = raw_full + kMaxRaw;               // [kMaxRaw]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(raw_empty + kMaxRaw);
  float* sc_s = reinterpret_cast<float*>(smem + 256);     // [nchunks * 32] scale (x 2^kActScaleLog2)
  float* sh_s = sc_s + kMaxChunks * kKC;                  // shift
  float* red_s = sh_s + kMaxChunks * kKC;                 // [kEpiWarps][16][2]
  // raw stages first: the 128-byte swizzle pattern needs 1024-byte aligned stage bases
  const uint32_t raw_stage_bytes = (uint32_t)HP * 128u;   // [pixel][32 fp32], a multiple of 1024 (HP % 8 == 0)
  unsigned char* R_s = smem + kHdrBytes;
  unsigned char* A_s = R_s + (size_t)RST * raw_stage_bytes;
  unsigned char* B_s = A_s + (size_t)AST * a_stage_bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#define DBG(slot) do { if (a.dbg != nullptr && blockIdx.x < 4 && (slot) < 64) a.dbg[(size_t)blockIdx.x * 64 + (slot)] = clock64(); } while (0)
  if (threadIdx.x == 0) DBG(0);
  const uint32_t ts_cols = 2u * NQ;       // G0 = a1*w1 | G1 = a1*w2 + a2*w1
  uint32_t tmem_cols = 32;
  while (tmem_cols < ts_cols * kTS) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(&a_full[i], kProdWarps);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < kMaxChunks; ++i) mbar_init(&b_full[i], 1);
    for (int i = 0; i < kMaxRaw; ++i) {
      mbar_init(&raw_full[i], 1);
      mbar_init(&raw_empty[i], kProdWarps);
    }
    for (int i = 0; i < kTS; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], kEpiWarps);
    }
    fence_mbar_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    // resident filter, fetched before griddepcontrol.wait when it was packed >= 2 launches ago (see
    // conv_dense_bwd.cu); the unit-test entry point (b_early = 0) loads it behind the wait
    if (a.b_early && (int)blockIdx.x < n_tiles) {
      for (int c = 0; c < nchunks; ++c) {
        mbar_arrive_expect_tx(&b_full[c], b_chunk_bytes);
        tma_load_1d(B_s + (size_t)c * b_chunk_bytes,
                    reinterpret_cast<const unsigned char*>(a.wpk) + (size_t)c * b_chunk_bytes, b_chunk_bytes,
                    &b_full[c]);
      }
    }
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  __syncthreads();  // barrier initialisation visible: the first raw loads go out before the constants below
  griddep_wait();   // x, the batch statistics and the packed filter come from earlier kernels of the step
  if (threadIdx.x == 0) DBG(1);
  // the first RST raw chunks need no free-slot wait: issued now, they are in flight while the BatchNorm
  // constants are computed (saves one load latency per launch)
  int pre_issued = 0;
  if (warp == kTmaWarp && lane == 0 && !(a.exp & 4) && (int)blockIdx.x < n_tiles) {
    int tile = blockIdx.x, ch = 0;
    while (pre_issued < RST && tile < n_tiles) {
      const int b = tile / tiles_per_img, r0 = (tile - b * tiles_per_img) * TR;
      mbar_arrive_expect_tx(&raw_full[pre_issued], raw_stage_bytes);
      tma_load_3d(R_s + (size_t)pre_issued * raw_stage_bytes, &tmX, ch * kKC, (r0 - 1) * W, b, &raw_full[pre_issued]);
      ++pre_issued;
      if (++ch == nchunks) {
        ch = 0;
        tile += gridDim.x;
      }
    }
  }
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
  if (threadIdx.x == 0) DBG(2);

  if (warp == kTmaWarp) {
    // ===== raw tile loads: box (32 channels, HP pixels, 1 image) -> [pixel][128 B], 128-byte swizzle =====
    if (lane == 0) {
      int q_it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_img, r0 = (tile - b * tiles_per_img) * TR;
        for (int ch = 0; ch < nchunks; ++ch, ++q_it) {
          if (q_it < pre_issued) continue;   // issued before the prologue (stage q_it, first pass of the ring)
          const int s = q_it % RST;
          mbar_wait(&raw_empty[s], (uint32_t)(((q_it / RST) & 1) ^ 1));
          if (a.exp & 4) {
            mbar_arrive(&raw_full[s]);
            continue;
          }
          mbar_arrive_expect_tx(&raw_full[s], raw_stage_bytes);
          tma_load_3d(R_s + (size_t)s * raw_stage_bytes, &tmX, ch * kKC, (r0 - 1) * W, b, &raw_full[s]);
        }
      }
    }
  } else if (warp >= kProdWarp0) {
    // ===== converters: raw fp32 -> BN + ReLU -> two fp16 pieces -> UMMA shared-memory image =====
    // item = (8-pixel group g, channel octet q, pixel e): consecutive lanes walk e, then q.  A quarter-warp
    // reads one 16-byte chunk column of eight consecutive swizzled rows (conflict-free) and stores 128
    // contiguous bytes.
    const int pt = threadIdx.x - kProdWarp0 * 32;
    const int e = pt & 7, q = (pt >> 3) & 3, g0 = pt >> 5;   // a warp = one 8-pixel group x 4 channel octets
    const int n_groups = HP >> 3;
    const int wsh = W == 32 ? 5 : (W == 16 ? 4 : 3);
    const bool relu = a.pro != 0;
    const size_t plane_elems = (size_t)a.B * H * W * a.Cp;
    const int oct_total = a.Cp >> 3;
    int q_it = 0;  // global (tile, chunk) counter of this CTA
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int b = tile / tiles_per_img, r0 = (tile - b * tiles_per_img) * TR;
      for (int ch = 0; ch < nchunks; ++ch, ++q_it) {
        const int rs = q_it % RST, s = q_it % AST;
        if (lane == 0) {
          mbar_wait(&raw_full[rs], (uint32_t)((q_it / RST) & 1));
          if (warp == kProdWarp0) DBG(32 + q_it);
          mbar_wait(&a_empty[s], (uint32_t)(((q_it / AST) & 1) ^ 1));
        }
        __syncwarp();
        const unsigned char* raw = R_s + (size_t)rs * raw_stage_bytes;
        unsigned char* st = A_s + (size_t)s * a_stage_bytes + (size_t)q * HP * 16;
        const float4 s0 = *reinterpret_cast<const float4*>(sc_s + ch * kKC + q * 8);
        const float4 s1 = *reinterpret_cast<const float4*>(sc_s + ch * kKC + q * 8 + 4);
        const float4 o0 = *reinterpret_cast<const float4*>(sh_s + ch * kKC + q * 8);
        const float4 o1 = *reinterpret_cast<const float4*>(sh_s + ch * kKC + q * 8 + 4);
        const int oq = ch * (kKC / 8) + q;
        const bool want_plane = a.planes != nullptr && oq < oct_total;
#pragma unroll 2
        for (int g = g0; g < ((a.exp & 1) ? 0 : n_groups); g += kProdWarps) {
          const int p = g * 8 + e;
          const int prow = p >> wsh, row = r0 - 1 + prow;   // warp-uniform (8 divides W)
          uint4 h1 = make_uint4(0u, 0u, 0u, 0u), h2 = h1;
          // rows outside the image are the convolution's zero padding of the ACTIVATION (after BN + ReLU)
          if (row >= 0 && row < H) {
            // swizzled row p: logical 16-byte chunk c sits at chunk c ^ (p & 7)
            const unsigned char* rrow = raw + (size_t)p * 128;
            const float4 x0 = *reinterpret_cast<const float4*>(rrow + (((2 * q) ^ e) << 4));
            const float4 x1 = *reinterpret_cast<const float4*>(rrow + (((2 * q + 1) ^ e) << 4));
            float t[8] = {fmaf(x0.x, s0.x, o0.x), fmaf(x0.y, s0.y, o0.y), fmaf(x0.z, s0.z, o0.z),
                          fmaf(x0.w, s0.w, o0.w), fmaf(x1.x, s1.x, o1.x), fmaf(x1.y, s1.y, o1.y),
                          fmaf(x1.z, s1.z, o1.z), fmaf(x1.w, s1.w, o1.w)};
            if (relu) {
#pragma unroll
              for (int k = 0; k < 8; ++k) t[k] = fmaxf(t[k], 0.f);
            }
            uint32_t p1[4], p2[4];
            if (a.lowp == LOWP_BF16) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                p1[k] = pack_bf2(t[2 * k], t[2 * k + 1]);
                p2[k] = 0u;
              }
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                p1[k] = pack_h2(t[2 * k], t[2 * k + 1]);
                const float2 f = unpack_h2(p1[k]);
                p2[k] = a.lowp ? 0u : pack_h2(t[2 * k] - f.x, t[2 * k + 1] - f.y);
              }
            }
            h1 = make_uint4(p1[0], p1[1], p1[2], p1[3]);
            h2 = make_uint4(p2[0], p2[1], p2[2], p2[3]);
            if (want_plane && prow >= 1 && prow <= TR) {
              const int col = p & (W - 1);
              op16* pd = a.planes + ((((size_t)b * H + row) * oct_total + oq) * W + col) * 8;
              *reinterpret_cast<uint4*>(pd) = h1;
              if (!a.lowp) *reinterpret_cast<uint4*>(pd + plane_elems) = h2;
            }
          }
          *reinterpret_cast<uint4*>(st + (size_t)p * 16) = h1;
          if (!a.lowp) *reinterpret_cast<uint4*>(st + (size_t)p * 16 + a_piece_bytes) = h2;
        }
        fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&raw_empty[rs]);
          mbar_arrive(&a_full[s]);
          if (warp == kProdWarp0) DBG(48 + q_it);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===== resident filter (bulk TMA, once per CTA) + MMA issue =====
    if (!a.b_early && lane == 0 && (int)blockIdx.x < n_tiles) {
      for (int c = 0; c < nchunks; ++c) {
        mbar_arrive_expect_tx(&b_full[c], b_chunk_bytes);
        tma_load_1d(B_s + (size_t)c * b_chunk_bytes,
                    reinterpret_cast<const unsigned char*>(a.wpk) + (size_t)c * b_chunk_bytes, b_chunk_bytes,
                    &b_full[c]);
      }
    }
    const uint32_t idesc2 = make_idesc_f16(128, 2 * NQ) | idesc_fmt_bits(a.lowp);
    const uint32_t idesc1 = make_idesc_f16(128, NQ) | idesc_fmt_bits(a.lowp);
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
        if (lane == 0) DBG(3 + tile_it * nchunks + ch);
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
                if (a.lowp) {
                  umma_f16_w(d0, a_k, a_hi, b_k, b_hi, idesc1, first ? 0u : 1u);                     // one piece: a1 x w1
                } else {
                  umma_f16_w(d0, a_k, a_hi, b_k, b_hi, idesc2, first ? 0u : 1u);                     // a1 x [w1|w2]
                  if (!(a.exp & 2))
                    umma_f16_w(d0 + (uint32_t)NQ, a_k + a_piece_u, a_hi, b_k, b_hi, idesc1, 1u);      // a2 x w1 -> G1
                }
              }
            }
          }
          umma_commit(&a_empty[sa]);
          if (ch == nchunks - 1) umma_commit(&acc_full[ts]);
        }
        __syncwarp();
        if (++sa == AST) {
          sa = 0;
          pa ^= 1u;
        }
      }
      if (++ts == kTS) {
        ts = 0;
        pt ^= 1u;
      }
    }
    if (lane == 0) DBG(20);
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
      if (threadIdx.x == 0) DBG(24 + 2 * tile_it);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)ts * ts_cols;
      for (int c0 = 0; c0 < CoP; c0 += 16) {
        float o[16];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          float g1[16], g0[16];
          if (a.lowp) {
#pragma unroll
            for (int i = 0; i < 16; ++i) g1[i] = 0.f;
          } else {
            tmem_ld16(taddr + (uint32_t)(NQ + kx * CoP + c0), g1);   // cross terms first (small), then the leading ones
          }
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
      if (threadIdx.x == 0) DBG(25 + 2 * tile_it);
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
  if (threadIdx.x == 0) DBG(30);
#undef DBG
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

static_assert(256 + sizeof(float) * (2 * kMaxChunks * kKC + kEpiWarps * 32) <= kHdrBytes, "header overflow");

size_t dense_smem(int W, int Cin, int CoP, int RST, int AST) {
  const int TR = 128 / W, HP = (TR + 2) * W;
  const int nchunks = (Cin + kKC - 1) / kKC;
  return 1024 /* alignment slack of the dynamic window */ + kHdrBytes + (size_t)RST * HP * 128 +
         (size_t)AST * 2 * (kKC / 8) * HP * 16 + (size_t)nchunks * 3 * (kKC / 8) * 2 * (3 * CoP) * 16;
}
// stage counts that fit: operand stages first (they hide the MMA -> converter handshake), then raw stages
bool dense_stages(int W, int Cin, int* RST, int* AST) {
  for (int a = kMaxStages; a >= 2; --a)
    for (int r = (a > 2 ? 2 : kMaxRaw); r >= 2; --r)
      if (dense_smem(W, Cin, 16, r, a) <= 227 * 1024) {
        *RST = r;
        *AST = a;
        return true;
      }
  return false;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn get_encode_d() {
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

bool dense_fwd_supported(int KS, int stride, int pad, int up, int Cin, int Cout, int H, int W) {
  if (KS != 3 || stride != 1 || pad != 1 || up) return false;
  if (!(W == 8 || W == 16 || W == 32) || H < 1) return false;
  if (Cout < 1 || Cout > 16) return false;
  if (Cin < 1 || (Cin + kKC - 1) / kKC > kMaxChunks) return false;
  int r, a;
  return dense_stages(W, Cin, &r, &a);
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
  PDES_REQUIRE((a.ldx & 3) == 0 && ((uintptr_t)a.x & 15u) == 0, PDES_ERR_INVALID,
               "conv_dense: input rows must be 16-byte aligned (ldx %d)", a.ldx);
  int RST = 2, AST = 2;
  dense_stages(a.W, a.Cin, &RST, &AST);
  const size_t smem = dense_smem(a.W, a.Cin, a.CoP, RST, AST);
  PDES_ENSURE_SMEM(conv_dense_fwd_kernel, smem);
  const int TR = 128 / a.W, HP = (TR + 2) * a.W;
  const int tiles = ((a.H + TR - 1) / TR) * a.B;
  int grid = sm_count();
  if (grid > tiles) grid = tiles;
  EncodeFn enc = get_encode_d();
  PDES_REQUIRE(enc != nullptr, PDES_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  CUtensorMap tm;
  {
    // x viewed as (channel, pixel of one image, image): reads beyond Cin and outside the image are zero-filled
    const cuuint64_t gdim[3] = {(cuuint64_t)a.Cin, (cuuint64_t)a.H * a.W, (cuuint64_t)a.B};
    const cuuint64_t gstr[2] = {(cuuint64_t)a.ldx * 4, (cuuint64_t)a.H * a.W * a.ldx * 4};
    const cuuint32_t box[3] = {(cuuint32_t)kKC, (cuuint32_t)HP, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(a.x), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PDES_REQUIRE(r == CUDA_SUCCESS, PDES_ERR_CUDA, "cuTensorMapEncodeTiled (conv_dense) failed with code %d", (int)r);
  }
  PDES_CUDA(launch_pdl(conv_dense_fwd_kernel, dim3(grid), dim3(kThreads), smem, st, tm, a, RST, AST));
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

}  // namespace pdes

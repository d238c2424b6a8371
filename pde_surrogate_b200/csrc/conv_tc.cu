// conv_tc.cu — tcgen05 (5th-gen tensor core) implicit-GEMM convolution for sm_100a.
//
// One CTA computes a 16x8 pixel tile (M = 128 rows of the GEMM) for ALL output channels
// (N <= 256, a multiple of 16) with the accumulator in tensor memory (TMEM).  The kernel is
// im2col-free: the BN+ReLU'd input halo tile (18x10 pixels for a 3x3 filter) of one channel chunk
// is written ONCE into shared memory in the canonical no-swizzle K-major core-matrix layout, and
// each of the 9 filter taps is the SAME buffer addressed through a shifted UMMA shared-memory
// descriptor (start address + (ky*10+kx)*16 B, stride-byte-offset = one halo row).  Filter tiles
// are pre-packed in global memory in exactly the shared-memory image and arrive by TMA bulk copies
// (cp.async.bulk + mbarrier complete_tx).
//
// fp32 parity on TF32 tensor cores: 3xTF32 split.  x = hi + lo with hi = x & 0xffffe000 (exactly
// representable in TF32, so hardware rounding mode is irrelevant); D += Ahi*Bhi + Ahi*Blo + Alo*Bhi
// accumulates in fp32 in TMEM.  (1xTF32 misses the 1e-4 bar by 36x)
//
// Warp roles (192 threads): warp 0 = TMA producer of filter tiles + TMEM allocator,
// warp 1 = MMA issuer (one elected lane), warps 2-5 = operand transform (global -> BN+ReLU ->
// hi/lo -> smem) during the main loop, then the epilogue (tcgen05.ld -> store / statistics or
// ReLU-mask + BatchNorm backward).
#include "conv.cuh"
#include "conv_tc.cuh"
#include "tc_common.cuh"

namespace pdes {
namespace {

using namespace tc;

constexpr int kTcThreads = 192;
constexpr int kTH = 16, kTW = 8;  // pixel tile (rows x cols) = 128 GEMM rows

// ---------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------
template <int KS>
__global__ void __launch_bounds__(kTcThreads, 1) conv_tc_kernel(TcConvArgs t) {
  constexpr int T = KS * KS;
  constexpr int HWp = kTW + KS - 1;          // halo row pitch in pixels
  constexpr int HP = (kTH + KS - 1) * HWp;   // halo pixels
  constexpr int HPpad = HP | 1;              // odd -> conflict-free transform stores
  const ConvArgs& a = t.c;
  const int N = t.N, KC = t.KC, NB = t.NB;
  const int kq = KC >> 2;

  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint64_t* a_full = bars;            // [2]
  uint64_t* a_empty = bars + 2;       // [2]
  uint64_t* b_full = bars + 4;        // [NB] (NB <= 8)
  uint64_t* b_empty = bars + 12;      // [NB]
  uint64_t* acc_full = bars + 20;     // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);
  const int CinP4 = (a.Cin + 3) & ~3;
  float* sc_s = reinterpret_cast<float*>(smem + 256);  // CinP4
  float* sh_s = sc_s + CinP4;                          // CinP4
  float* ep_s = sh_s + CinP4;                          // 4*N
  float* red_s = ep_s + 4 * N;                         // 4*N*2
  const size_t hdr = (256 + sizeof(float) * ((size_t)2 * CinP4 + 12 * (size_t)N) + 127) & ~(size_t)127;
  const uint32_t a_stage_bytes = 2u * HPpad * KC * 4u;  // hi + lo
  const int TPB = t.TPB;                                 // filter taps per TMA stage
  const uint32_t b_tap_bytes = 2u * N * KC * 4u;        // hi + lo of one tap
  const uint32_t b_stage_bytes = (uint32_t)TPB * b_tap_bytes;
  unsigned char* A_s = smem + hdr;
  unsigned char* B_s = A_s + 2 * (size_t)a_stage_bytes;

  const int tiles_x = (a.Wo + kTW - 1) / kTW, tiles_y = (a.Ho + kTH - 1) / kTH;
  int bid = blockIdx.x;
  const int tx = bid % tiles_x;
  bid /= tiles_x;
  const int ty = bid % tiles_y;
  const int b = bid / tiles_y;
  const int oy0 = ty * kTH, ox0 = tx * kTW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = t.nchunks;
  const int nsteps = nchunks * (T / TPB);               // TMA stages over the whole K loop
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(t.S * 2 * N)) tmem_cols <<= 1;

  // ---- setup -------------------------------------------------------------------------
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 128);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < NB; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  if (warp >= 2) {
    const int tt = threadIdx.x - 64;
    if (a.pro) {
      for (int c = tt; c < CinP4; c += 128) {
        float s = 0.f, h = 0.f, m, is;
        if (c < a.Cin) bn_consts_tc(a.bn, c, s, h, m, is);
        sc_s[c] = s;
        sh_s[c] = h;
      }
    }
    if (a.epi == EPI_BNBWD) {
      for (int n = tt; n < N; n += 128) {
        float s = 0.f, h = 0.f, m = 0.f, is = 0.f;
        if (n < a.Cout) bn_consts_tc(a.fbn, n, s, h, m, is);
        ep_s[n] = s;
        ep_s[N + n] = h;
        ep_s[2 * N + n] = m;
        ep_s[3 * N + n] = is;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: filter tiles =====
    if (lane == 0) {
      for (int step = 0; step < nsteps; ++step) {
        const int sb = step % NB;
        mbar_wait(&b_empty[sb], (uint32_t)(((step / NB) & 1) ^ 1));
        mbar_arrive_expect_tx(&b_full[sb], b_stage_bytes);
        tma_load_1d(B_s + (size_t)sb * b_stage_bytes,
                    t.wtc + (size_t)step * (b_stage_bytes / 4), b_stage_bytes, &b_full[sb]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // Accumulator layout in TMEM: set s (= chunk % S) owns columns [s*2N, s*2N+N) for the
      // hi*hi products and [s*2N+N, s*2N+2N) for the cross terms hi*lo + lo*hi.  Tensor-core
      // accumulation truncates, so the error grows with the number of MMAs that feed one
      // accumulator: big (hi*hi) and small (cross) terms are kept apart and the K range is
      // spread over S accumulators; the epilogue adds them with round-to-nearest fp32.
      const bool fused = (2 * N <= 256);  // Ahi x [Bhi|Blo] as ONE N'=2N instruction
      const uint32_t idesc_n = make_idesc(128, N);
      const uint32_t idesc_2n = make_idesc(128, 2 * N);
      const uint32_t lbo_a = HPpad * 16u, sbo_a = HWp * 16u;
      const uint32_t lbo_b = 2u * (uint32_t)N * 16u, sbo_b = 128u;  // k-quad stride: [hi N rows | lo N rows]
      const uint32_t a_lo_off = (uint32_t)kq * HPpad * 16u;
      const uint32_t b_lo_off = (uint32_t)N * 16u;
      uint32_t used = 0;  // bit s: accumulator set s already holds data
      int step = 0;
      for (int ch = 0; ch < nchunks; ++ch) {
        const int sa = ch & 1;
        const int set = ch % t.S;
        const uint32_t d_hh = tmem_base + (uint32_t)(set * 2 * N);
        const uint32_t d_x = d_hh + (uint32_t)N;
        mbar_wait(&a_full[sa], (uint32_t)((ch >> 1) & 1));
        const uint32_t a_base = smem_u32(A_s + (size_t)sa * a_stage_bytes);
        const uint64_t a_desc0 = make_desc(a_base, lbo_a, sbo_a);
        for (int tap = 0; tap < T; ++tap) {
          const int sb = step % NB;
          const int tin = tap % TPB;
          if (tin == 0) {
            mbar_wait(&b_full[sb], (uint32_t)((step / NB) & 1));
            tc_fence_after();
          }
          const uint32_t b_base = smem_u32(B_s + (size_t)sb * b_stage_bytes) + (uint32_t)tin * b_tap_bytes;
          const uint64_t b_desc0 = make_desc(b_base, lbo_b, sbo_b);
          const uint64_t a_tap = a_desc0 + (uint64_t)((uint32_t)((tap / KS) * HWp + (tap % KS)));  // +16 B units
          for (int ks = 0; ks < KC / 8; ++ks) {
            const uint32_t acc = (used >> set) & 1u;
            const uint64_t ahi = a_tap + (uint64_t)((2u * ks * lbo_a) >> 4);
            const uint64_t bhi = b_desc0 + (uint64_t)((2u * ks * lbo_b) >> 4);
            if (t.prec != 0) {
              umma_tf32(d_hh, ahi, bhi, idesc_n, acc);
            } else {
              const uint64_t alo = ahi + (uint64_t)(a_lo_off >> 4);
              if (fused) {
                umma_tf32(d_hh, ahi, bhi, idesc_2n, acc);  // [hh | x] += Ahi * [Bhi | Blo]
              } else {
                const uint64_t blo = bhi + (uint64_t)(b_lo_off >> 4);
                umma_tf32(d_hh, ahi, bhi, idesc_n, acc);
                umma_tf32(d_x, ahi, blo, idesc_n, acc);
              }
              umma_tf32(d_x, alo, bhi, idesc_n, 1);
            }
            used |= 1u << set;
          }
          if (tin == TPB - 1) {
            umma_commit(&b_empty[sb]);
            ++step;
          }
        }
        umma_commit(&a_empty[sa]);
      }
      umma_commit(acc_full);
    }
  } else {
    // ===== operand transform (warps 2..5) =====
    const int tt = threadIdx.x - 64;
    const int Hv = a.in_mode == IN_DIRECT ? a.Hs : 2 * a.Hs;
    const int Wv = a.in_mode == IN_DIRECT ? a.Ws : 2 * a.Ws;
    const int iy0 = oy0 - a.pad, ix0 = ox0 - a.pad;
    for (int ch = 0; ch < nchunks; ++ch) {
      const int sa = ch & 1;
      mbar_wait(&a_empty[sa], (uint32_t)(((ch >> 1) & 1) ^ 1));
      unsigned char* As = A_s + (size_t)sa * a_stage_bytes;
      const int c0 = ch * KC;
      // all global loads of the chunk are issued before any is consumed (memory-level
      // parallelism: one L2/HBM latency per chunk instead of one per item)
      constexpr int MAXI = (HP * 8 + 127) / 128;
      float4 raw[MAXI];
      int meta[MAXI];  // (hp << 8) | (q << 1) | in-image ; -1 = no item
#pragma unroll
      for (int j = 0; j < MAXI; ++j) {
        const int i = tt + j * 128;
        raw[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        meta[j] = -1;
        if (i < HP * kq) {
          const int q = i % kq, hp = i / kq;
          const int hy = hp / HWp, hx = hp - hy * HWp;
          const int vy = iy0 + hy, vx = ix0 + hx;
          const bool in = vy >= 0 && vy < Hv && vx >= 0 && vx < Wv;
          meta[j] = (hp << 8) | (q << 1) | (in ? 1 : 0);
          if (in) {
            const int sy = a.in_mode == IN_DIRECT ? vy : (vy >> 1);
            const int sx = a.in_mode == IN_DIRECT ? vx : (vx >> 1);
            const int c = c0 + 4 * q;
            const float* p = a.x + (((size_t)b * a.Hs + sy) * a.Ws + sx) * a.ldx + c;
            if (c + 3 < a.Cin) {
              raw[j] = __ldg(reinterpret_cast<const float4*>(p));
            } else {
              float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (c + k < a.Cin) v[k] = p[k];
              raw[j] = make_float4(v[0], v[1], v[2], v[3]);
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < MAXI; ++j) {
        if (meta[j] < 0) continue;
        const int hp = meta[j] >> 8, q = (meta[j] >> 1) & 127;
        float v[4] = {raw[j].x, raw[j].y, raw[j].z, raw[j].w};
        if (a.pro) {
          const int c = c0 + 4 * q;
          const bool in = meta[j] & 1;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int cc = c + k;
            v[k] = (in && cc < a.Cin) ? fmaxf(0.f, fmaf(v[k], sc_s[cc], sh_s[cc])) : 0.f;
          }
        }
        float hi[4], lo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) split_tf32(v[k], hi[k], lo[k]);
        unsigned char* dst = As + ((size_t)q * HPpad + hp) * 16;
        *reinterpret_cast<float4*>(dst) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(dst + (size_t)kq * HPpad * 16) = make_float4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async_smem();
      mbar_arrive(&a_full[sa]);
    }

    // ===== epilogue =====
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int quarter = warp & 3;            // TMEM lane quarter this warp may read
    const int row = quarter * 32 + lane;     // GEMM row = tile pixel
    const int py_t = row >> 3, px_t = row & 7;
    int oy = oy0 + py_t, ox = ox0 + px_t;
    bool valid = oy < a.Ho && ox < a.Wo;
    int Hd = a.Ho, Wd = a.Wo;
    if (a.pool) {
      valid = valid && ((py_t | px_t) & 1) == 0;
      oy >>= 1;
      ox >>= 1;
      Hd >>= 1;
      Wd >>= 1;
    }
    const size_t pix = ((size_t)b * Hd + oy) * Wd + ox;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const bool want_red = (a.epi == EPI_NHWC && a.o_sum != nullptr) || a.epi == EPI_BNBWD;
    for (int n0 = 0; n0 < N; n0 += 16) {
      float v[16];
      {
        const int nsets = nchunks < t.S ? nchunks : t.S;
        float xs[16];
        tmem_ld16(taddr + (uint32_t)n0, v);
        if (t.prec == 0) tmem_ld16(taddr + (uint32_t)(N + n0), xs);
        for (int sset = 1; sset < nsets; ++sset) {
          float w1[16];
          tmem_ld16(taddr + (uint32_t)(sset * 2 * N + n0), w1);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += w1[i];
          if (t.prec == 0) {
            tmem_ld16(taddr + (uint32_t)(sset * 2 * N + N + n0), w1);
#pragma unroll
            for (int i = 0; i < 16; ++i) xs[i] += w1[i];
          }
        }
        if (t.prec == 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += xs[i];
        }
      }
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
        } else {  // EPI_BNBWD
          const float* xs = a.fx + pix * a.ldfx + n0;
          float* gp = a.G + pix * a.ldG + n0;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float xv[4] = {0.f, 0.f, 0.f, 0.f}, gv[4] = {0.f, 0.f, 0.f, 0.f};
            const bool full = n0 + i + 3 < a.Cout;
            if (full) {
              const float4 f = *reinterpret_cast<const float4*>(xs + i);
              xv[0] = f.x; xv[1] = f.y; xv[2] = f.z; xv[3] = f.w;
              if (a.g_accum) {
                const float4 g4 = *reinterpret_cast<const float4*>(gp + i);
                gv[0] = g4.x; gv[1] = g4.y; gv[2] = g4.z; gv[3] = g4.w;
              }
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (n0 + i + k < a.Cout) {
                  xv[k] = xs[i + k];
                  if (a.g_accum) gv[k] = gp[i + k];
                }
            }
            float out4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int nl = n0 + i + k;
              const float z = fmaf(xv[k], ep_s[nl], ep_s[N + nl]);
              const float dz = (nl < a.Cout && z > 0.f) ? v[i + k] : 0.f;
              const float xh = (xv[k] - ep_s[2 * N + nl]) * ep_s[3 * N + nl];
              s1[i + k] = dz;
              s2[i + k] = dz * xh;
              out4[k] = gv[k] + ep_s[nl] * dz;
            }
            if (full) {
              *reinterpret_cast<float4*>(gp + i) = make_float4(out4[0], out4[1], out4[2], out4[3]);
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (n0 + i + k < a.Cout) gp[i + k] = out4[k];
            }
          }
        }
      }
      if (want_red) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float u = s1[i], w = s2[i];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            u += __shfl_xor_sync(0xffffffffu, u, o);
            w += __shfl_xor_sync(0xffffffffu, w, o);
          }
          if (lane == 0) {
            red_s[((size_t)quarter * N + n0 + i) * 2 + 0] = u;
            red_s[((size_t)quarter * N + n0 + i) * 2 + 1] = w;
          }
        }
      }
    }
    if (want_red) {
      named_bar_sync(1, 128);
      for (int n = threadIdx.x - 64; n < a.Cout; n += 128) {
        double u = 0.0, w = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
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
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------
// filter packing: OIHW fp32 -> per (chunk, tap) shared-memory images, hi and lo TF32 halves
//   dst[(((chunk*T + tap)*(KC/4) + k/4)*2 + h)*N*4 + n*4 + k%4]      (h = 0 hi, 1 lo)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_tc_kernel(const TcPackDesc* tab) {
  const TcPackDesc d = tab[blockIdx.y];
  const int T = d.KS * d.KS;
  const size_t per_step = (size_t)2 * d.N * d.KC;
  const size_t total = (size_t)d.nchunks * T * per_step;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t step = i / per_step;
    size_t r = i - step * per_step;
    const int kq = (int)(r / ((size_t)2 * d.N * 4));
    r -= (size_t)kq * 2 * d.N * 4;
    const int h = (int)(r / ((size_t)d.N * 4));
    r -= (size_t)h * d.N * 4;
    const int n = (int)(r >> 2), k4 = (int)(r & 3);
    const int chunk = (int)(step / T), tap = (int)(step % T);
    const int k = chunk * d.KC + kq * 4 + k4;
    float v = 0.f;
    if (!d.transpose) {
      // forward: n = output channel, k = input channel
      if (n < d.Cout && k < d.Cin) v = d.w[((size_t)n * d.Cin + k) * T + tap];
    } else {
      // dgrad: n = input channel, k = output channel, taps flipped
      if (n < d.Cin && k < d.Cout) v = d.w[((size_t)k * d.Cin + n) * T + (T - 1 - tap)];
    }
    float hi, lo;
    split_tf32(v, hi, lo);
    d.dst[i] = h == 0 ? hi : lo;
  }
}

size_t tc_smem_bytes(int KS, int Cin, int N, int KC, int NB, int TPB) {
  const int HWp = kTW + KS - 1;
  const int HP = (kTH + KS - 1) * HWp;
  const int HPpad = HP | 1;
  const int CinP4 = (Cin + 3) & ~3;
  const size_t hdr = (256 + sizeof(float) * ((size_t)2 * CinP4 + 12 * (size_t)N) + 127) & ~(size_t)127;
  return hdr + 2 * (size_t)(2 * HPpad * KC * 4) + (size_t)NB * TPB * (2 * (size_t)N * KC * 4);
}

}  // namespace

void tc_plan(int KS, int Cin_k, int N, TcPlan* p) {
  // channel chunk: 32 unless the operand is thin or the filter tile would not fit
  int KC = 32;
  if (Cin_k <= 16 || N > 128) KC = 16;
  if (Cin_k <= 8) KC = 8;
  p->KC = KC;
  p->nchunks = (Cin_k + KC - 1) / KC;
  // filter taps per TMA stage: few large bulk copies instead of many latency-bound small ones
  const int T = KS * KS;
  int TPB = 1, NB = 4;
  const size_t tap_bytes = 2 * (size_t)N * KC * 4;
  if (T * tap_bytes <= 40 * 1024) {
    TPB = T;
    NB = 2;
  } else if (T % 3 == 0 && 3 * tap_bytes <= 50 * 1024) {
    TPB = 3;
    NB = 2;
  }
  while (NB > 2 && tc_smem_bytes(KS, Cin_k, N, KC, NB, TPB) > 220 * 1024) --NB;
  p->NB = NB;
  p->TPB = TPB;
  p->smem = tc_smem_bytes(KS, Cin_k, N, KC, NB, TPB);
  int S = 512 / (2 * N);
  if (S < 1) S = 1;
  if (S > p->nchunks) S = p->nchunks;
  if (S > 8) S = 8;
  p->S = S;
  p->pack_floats = (size_t)p->nchunks * KS * KS * 2 * N * KC;
}

bool tc_supported(int KS, int stride, int Cin_k, int N) {
  if (!(KS == 1 || KS == 3) || stride != 1) return false;
  if (N < 16 || N > 256 || (N & 15)) return false;  // 2N accumulator columns must fit the 512 of TMEM
  TcPlan p;
  tc_plan(KS, Cin_k, N, &p);
  return p.smem <= 225 * 1024;
}

int launch_conv_tc(const TcConvArgs& t, cudaStream_t st) {
  const ConvArgs& a = t.c;
  PDES_REQUIRE(a.KS == 1 || a.KS == 3, PDES_ERR_UNSUPPORTED, "conv_tc: kernel size %d", a.KS);
  PDES_REQUIRE(a.stride == 1 && !a.in_nchw && a.epi != EPI_NCHW, PDES_ERR_UNSUPPORTED,
               "conv_tc: stride-1 NHWC convolutions only");
  PDES_REQUIRE((a.ldx & 3) == 0 && ((uintptr_t)a.x & 15u) == 0, PDES_ERR_INVALID,
               "conv_tc: input must be 16-byte aligned with a pixel stride multiple of 4");
  if (a.epi == EPI_NHWC)
    PDES_REQUIRE(((a.ldy | a.coff) & 3) == 0 && ((uintptr_t)a.y & 15u) == 0, PDES_ERR_INVALID,
                 "conv_tc: output slice must be 16-byte aligned");
  if (a.epi == EPI_BNBWD)
    PDES_REQUIRE(((a.ldfx | a.ldG) & 3) == 0, PDES_ERR_INVALID, "conv_tc: gradient buffers misaligned");
  PDES_REQUIRE(!a.pool || ((a.Ho | a.Wo) & 1) == 0, PDES_ERR_INVALID, "conv_tc: pool needs even size");
  const size_t smem = tc_smem_bytes(a.KS, a.Cin, t.N, t.KC, t.NB, t.TPB);
  PDES_REQUIRE(smem <= 227 * 1024, PDES_ERR_UNSUPPORTED, "conv_tc: needs %zu bytes of shared memory", smem);
  const int tiles = ((a.Wo + kTW - 1) / kTW) * ((a.Ho + kTH - 1) / kTH) * a.B;
  if (a.KS == 3) {
    static size_t attr = 0;
    if (smem > attr) {
      PDES_CUDA(cudaFuncSetAttribute(conv_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = smem;
    }
    conv_tc_kernel<3><<<tiles, kTcThreads, smem, st>>>(t);
  } else {
    static size_t attr = 0;
    if (smem > attr) {
      PDES_CUDA(cudaFuncSetAttribute(conv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = smem;
    }
    conv_tc_kernel<1><<<tiles, kTcThreads, smem, st>>>(t);
  }
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

int launch_pack_tc(const TcPackDesc* dev_table, int n, size_t max_elems, cudaStream_t st) {
  if (n == 0) return PDES_OK;
  int bx = (int)((max_elems + 255) / 256);
  if (bx > 128) bx = 128;
  if (bx < 1) bx = 1;
  pack_tc_kernel<<<dim3(bx, n), 256, 0, st>>>(dev_table);
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

}  // namespace pdes

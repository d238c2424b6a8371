// tc_common.cuh — tcgen05 / TMEM / UMMA-descriptor helpers shared by the tensor-core kernels.
#pragma once
#include "conv.cuh"

namespace pdes {
namespace tc {

// ---- tcgen05 / TMEM PTX wrappers -----------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// UMMA shared-memory descriptor, SWIZZLE_NONE, K-major canonical layout
//   8 rows x 16 B core matrices (8 fp16 elements per row); SBO = next 8-row group, LBO = next k octet
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
  return d;
}
__device__ __forceinline__ void bn_consts_tc(const BnSrc& s, int c, float& scale, float& shift,
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
  invstd = (float)is;
  mean = (float)m;
  scale = s.gamma[c] * invstd;
  shift = s.beta[c] - mean * scale;
}


}  // namespace tc
}  // namespace pdes

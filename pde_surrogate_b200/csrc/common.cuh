// common.cuh — shared helpers for the sm_100a kernels and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include "../../include/pdes_b200.h"

#ifndef PDES_HD
#define PDES_HD __host__ __device__ __forceinline__
#endif

namespace pdes {

// ---- error plumbing --------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define PDES_CUDA(expr)                                                          \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess) return ::pdes::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define PDES_REQUIRE(cond, code, ...)      \
  do {                                     \
    if (!(cond)) {                         \
      ::pdes::set_error(__VA_ARGS__);      \
      return (code);                       \
    }                                      \
  } while (0)

#define PDES_LAUNCH_CHECK() PDES_CUDA(cudaGetLastError())

int sm_count();      // of the CURRENT device (cached per device)
int cur_device();    // cudaGetDevice(), clamped to [0, kMaxDevices)
constexpr int kMaxDevices = 64;
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute: the high-water
// mark already applied is tracked per device, so that a process driving several GPUs raises it on each
struct SmemAttr {
  size_t v[kMaxDevices];
  SmemAttr() { for (int i = 0; i < kMaxDevices; ++i) v[i] = 0; }
};
#define PDES_ENSURE_SMEM(kernel, bytes)                                                                   \
  do {                                                                                                    \
    static ::pdes::SmemAttr _pdes_attr;                                                                   \
    size_t& _cur = _pdes_attr.v[::pdes::cur_device()];                                                    \
    if ((size_t)(bytes) > _cur) {                                                                         \
      PDES_CUDA(cudaFuncSetAttribute((kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
      _cur = (size_t)(bytes);                                                                             \
    }                                                                                                     \
  } while (0)
bool pdl_enabled();  // programmatic dependent launch of the kernel chain (opt-in: env PDES_PDL=1)

// Launch with the programmatic-stream-serialization attribute: the kernel may start (and run its
// prologue up to griddep_wait()) while the previous kernel in the stream is still draining.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- device helpers ----------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
               : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// TMA 1-D bulk copy shared -> global (bulk async-group completion).
__device__ __forceinline__ void tma_store_1d(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// programmatic dependent launch: wait for the previous grid's results / let the next grid start
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// column sums over the 32 lanes of a warp for 16 values per lane, by recursive halving:
// 8+4+2+1+1 = 16 shuffles.  On return lane l holds in `out` the sum of column col_of_lane(l).
__device__ __forceinline__ float colsum16(const float v[16], int lane) {
  float a[8];
  const bool up16 = lane & 16;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = up16 ? v[i] : v[i + 8];
    const float keep = up16 ? v[i + 8] : v[i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  float b[4];
  const bool up8 = lane & 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = up8 ? a[i] : a[i + 4];
    const float keep = up8 ? a[i + 4] : a[i];
    b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  float c[2];
  const bool up4 = lane & 4;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = up4 ? b[i] : b[i + 2];
    const float keep = up4 ? b[i + 2] : b[i];
    c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const bool up2 = lane & 2;
  const float send = up2 ? c[0] : c[1];
  const float keep = up2 ? c[1] : c[0];
  float d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  return d;
}
// column index held by `lane` after colsum16
__device__ __forceinline__ int colsum16_col(int lane) {
  return ((lane & 16) ? 8 : 0) + ((lane & 8) ? 4 : 0) + ((lane & 4) ? 2 : 0) + ((lane & 2) ? 1 : 0);
}

#endif  // __CUDACC__

}  // namespace pdes

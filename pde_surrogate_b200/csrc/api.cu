// api.cu — error plumbing, device queries and the fused flat Adam step of the C-ABI.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>

namespace pdes {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d in `%s`", (int)e, cudaGetErrorString(e), file, line, what);
  return PDES_ERR_CUDA;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    // Programmatic dependent launch of the kernel chain: the next kernel's CTAs start (barrier init,
    // TMEM allocation, descriptor prefetch) while the previous grid drains, and block in
    // griddepcontrol.wait until it has completed.  Measured on B200 with the round-1 kernels of
    // this file set: 2.46 -> 2.24 ms per graph-replayed step (it lost 3 % with the early, slower
    // kernels, whose waiting dependents only took SM resources).  PDES_PDL=0 turns it off.
    const char* e = getenv("PDES_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

int cur_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return dev < 0 ? 0 : (dev >= kMaxDevices ? kMaxDevices - 1 : dev);
}

int sm_count() {
  static int cached[kMaxDevices];  // zero-initialised
  const int dev = cur_device();
  if (cached[dev] > 0) return cached[dev];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
    return 148;
  cached[dev] = n;
  return n;
}

// torch.optim.Adam (amsgrad=False, maximize=False): g += wd*p; m = b1 m + (1-b1) g;
// v = b2 v + (1-b2) g^2; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float wd,
            float gscale, float step_size, float inv_sqrt_bc2) {
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pa = &pp.x;
    const float* ga = &gg.x;
    float* ma = &mm.x;
    float* va = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gk = ga[k] * gscale;
      if (wd != 0.f) gk += wd * pa[k];
      ma[k] = b1 * ma[k] + (1.f - b1) * gk;
      va[k] = b2 * va[k] + (1.f - b2) * gk * gk;
      const float denom = sqrtf(va[k]) * inv_sqrt_bc2 + eps;
      pa[k] -= step_size * (ma[k] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail
  const int64_t t0 = n4 << 2;
  for (int64_t i = t0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    float gk = g[i] * gscale;
    if (wd != 0.f) gk += wd * p[i];
    const float mk = b1 * m[i] + (1.f - b1) * gk;
    const float vk = b2 * v[i] + (1.f - b2) * gk * gk;
    m[i] = mk;
    v[i] = vk;
    p[i] -= step_size * (mk / (sqrtf(vk) * inv_sqrt_bc2 + eps));
  }
}

__global__ void __launch_bounds__(256)
adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                float* __restrict__ v, int64_t n, const float* __restrict__ hyper) {
  griddep_wait();
  const float step_size = hyper[0], inv_sqrt_bc2 = hyper[1], b1 = hyper[2], b2 = hyper[3], eps = hyper[4],
              wd = hyper[5], gscale = hyper[6];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    float gk = g[i] * gscale;
    if (wd != 0.f) gk += wd * p[i];
    const float mk = b1 * m[i] + (1.f - b1) * gk;
    const float vk = b2 * v[i] + (1.f - b2) * gk * gk;
    m[i] = mk;
    v[i] = vk;
    p[i] -= step_size * (mk / (sqrtf(vk) * inv_sqrt_bc2 + eps));
  }
}

}  // namespace pdes

using namespace pdes;

extern "C" const char* pdes_last_error(void) { return g_err; }
extern "C" int pdes_abi_version(void) { return PDES_ABI_VERSION; }

extern "C" int pdes_device_info(int* smc, int* cc_major, int* cc_minor) {
  int dev = 0;
  PDES_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  PDES_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (smc) *smc = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return PDES_OK;
}

extern "C" int pdes_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr,
                              float beta1, float beta2, float eps, float weight_decay,
                              float grad_scale, int64_t step, void* stream) {
  PDES_REQUIRE(p && g && m && v, PDES_ERR_INVALID, "pdes_adam_step: null pointer");
  PDES_REQUIRE(n >= 0 && step >= 1, PDES_ERR_INVALID, "pdes_adam_step: n>=0 and step>=1 required");
  PDES_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15u) == 0,
               PDES_ERR_INVALID, "pdes_adam_step: buffers must be 16-byte aligned");
  if (n == 0) return PDES_OK;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  int blocks = (int)(((n >> 2) + 255) / 256);
  if (blocks < 1) blocks = 1;
  const int cap = sm_count() * 8;
  if (blocks > cap) blocks = cap;
  adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps,
                                                        weight_decay, grad_scale, step_size,
                                                        inv_sqrt_bc2);
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

extern "C" int pdes_adam_hyper(float* h, float lr, float beta1, float beta2, float eps, float weight_decay,
                               float grad_scale, int64_t step) {
  PDES_REQUIRE(h != nullptr && step >= 1, PDES_ERR_INVALID, "pdes_adam_hyper: null pointer or step < 1");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  h[0] = (float)((double)lr / bc1);
  h[1] = (float)(1.0 / sqrt(bc2));
  h[2] = beta1;
  h[3] = beta2;
  h[4] = eps;
  h[5] = weight_decay;
  h[6] = grad_scale;
  return PDES_OK;
}

extern "C" int pdes_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper,
                                  void* stream) {
  PDES_REQUIRE(p && g && m && v && hyper, PDES_ERR_INVALID, "pdes_adam_step_dev: null pointer");
  if (n <= 0) return PDES_OK;
  int blocks = (int)((n + 255) / 256);
  const int cap = sm_count() * 8;
  if (blocks > cap) blocks = cap;
  PDES_CUDA(launch_pdl(adam_dev_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, n, hyper));
  PDES_LAUNCH_CHECK();
  return PDES_OK;
}

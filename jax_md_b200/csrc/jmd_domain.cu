// Slab domain decomposition helpers: face / migration selection
// and gather-pack of halo and migration payloads.  The exchange itself is NCCL
// send/recv issued by the host on the same stream (jax_md_b200/domain.py).
#include <cuda_runtime.h>
#include "jmd_common.cuh"

namespace {

template <typename T>
__global__ void k_dd_select(int dim, int n, const T* pos, int axis, T lo, T L, T thr_a, T thr_b, int* list_a,
                            int* list_b, int* counters, int cap) {
  const int stride = gridDim.x * blockDim.x;
  const T half = L * T(0.5);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    T d = pos[(size_t)i * dim + axis] - lo;
    // periodic ring along the decomposition axis
    if (d >= half) d -= L;
    else if (d < -half) d += L;
    if (d < thr_a) {
      int p = atomicAdd(&counters[0], 1);
      if (p < cap) list_a[p] = i;
    }
    if (d >= thr_b) {
      int p = atomicAdd(&counters[1], 1);
      if (p < cap) list_b[p] = i;
    }
  }
}

template <typename T>
__global__ void k_dd_pack(int ncomp, int n_idx, const int* idx, const T* src, T* dst) {
  const long long total = (long long)n_idx * ncomp;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += stride) {
    const int i = (int)(e / ncomp), c = (int)(e % ncomp);
    dst[e] = src[(size_t)idx[i] * ncomp + c];
  }
}

inline int blocks_for(long long n) {
  long long g = (n + 255) / 256;
  if (g < 1) g = 1;
  if (g > JMD_SM_COUNT * 8) g = JMD_SM_COUNT * 8;
  return (int)g;
}

}  // namespace

extern "C" {

int jmd_dd_select(int dtype, int dim, int n, const void* position, int axis, double lo, double L, double thr_a,
                  double thr_b, int32_t* list_a, int32_t* list_b, int32_t* counters, int cap, void* stream) {
  if (!position || !list_a || !list_b || !counters || axis < 0 || axis >= dim) return JMD_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == JMD_F32)
    k_dd_select<float><<<blocks_for(n), 256, 0, s>>>(dim, n, (const float*)position, axis, (float)lo, (float)L,
                                                     (float)thr_a, (float)thr_b, list_a, list_b, counters, cap);
  else if (dtype == JMD_F64)
    k_dd_select<double><<<blocks_for(n), 256, 0, s>>>(dim, n, (const double*)position, axis, lo, L, thr_a, thr_b,
                                                      list_a, list_b, counters, cap);
  else return JMD_EINVAL;
  JMD_LAUNCH_CHECK();
  return 0;
}

int jmd_dd_pack(int dtype, int ncomp, int n_idx, const int32_t* idx, const void* src, void* dst, void* stream) {
  if (n_idx == 0) return 0;
  if (!idx || !src || !dst || ncomp < 1) return JMD_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  const int g = blocks_for((long long)n_idx * ncomp);
  if (dtype == JMD_F32) k_dd_pack<float><<<g, 256, 0, s>>>(ncomp, n_idx, idx, (const float*)src, (float*)dst);
  else if (dtype == JMD_F64) k_dd_pack<double><<<g, 256, 0, s>>>(ncomp, n_idx, idx, (const double*)src, (double*)dst);
  else return JMD_EINVAL;
  JMD_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"

// Slab domain decomposition helpers: face / migration selection
// and gather-pack of halo and migration payloads.  The exchange itself is NCCL
// send/recv issued by the host on the same stream (jax_md_b200/domain.py).
#include <cuda_runtime.h>
#include "jmd_common.cuh"

namespace {

template <typename T>
__global__ void k_dd_select(int dim, int n, const int* n_dev, const T* pos, int axis, T lo, T L, T thr_a, T thr_b,
                            int* list_a, int* list_b, int* counters, int cap) {
  if (n_dev) n = min(n, *n_dev);
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

// ---- rebuild-time pipeline without host round trips ----------------------------------
// Counts stay on the device: every message has a fixed capacity and carries its own
// element count, `info` collects what the host needs and is read back once.

// Ordered compaction by one CTA: out[k] = val(t) for the k-th t in [0, n) with pred(t).
template <typename Pred, typename Val>
__device__ int cta_compact(int n, Pred pred, Val val, int* out) {
  __shared__ int warp_sums[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int carry = 0;
  for (int base = 0; base < n; base += blockDim.x) {
    const int t = base + threadIdx.x;
    const bool p = (t < n) && pred(t);
    const unsigned ballot = __ballot_sync(0xffffffffu, p);
    if (lane == 0) warp_sums[w] = __popc(ballot);
    __syncthreads();
    int off = 0, tot = 0;
    for (int k = 0; k < nw; ++k) {
      const int s = warp_sums[k];
      if (k < w) off += s;
      tot += s;
    }
    if (p) out[carry + off + __popc(ballot & ((1u << lane) - 1u))] = val(t);
    carry += tot;
    __syncthreads();
  }
  return carry;
}

// Migration payload of both directions (blockIdx.y): R | P | F rows and the global ids;
// gid_msg[0] is the number of atoms in the message.
template <typename T>
__global__ void k_dd_pack_migrate(int dim, int cap_mig, const int* list_a, const int* list_b, const int* counters,
                                  const T* R, const T* P, const T* F, const long long* gid, T* pay_a, T* pay_b,
                                  long long* gm_a, long long* gm_b) {
  const int dir = blockIdx.y;
  const int* list = dir ? list_b : list_a;
  T* pay = dir ? pay_b : pay_a;
  long long* gm = dir ? gm_b : gm_a;
  const int n = min(counters[dir], cap_mig);
  const int row = 3 * dim;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  if (tid == 0) gm[0] = n;
  for (int e = tid; e < n * row; e += stride) {
    const int i = e / row, c = e % row, arr = c / dim, cc = c % dim;
    const T* src = arr == 0 ? R : (arr == 1 ? P : F);
    pay[e] = src[(size_t)list[i] * dim + cc];
  }
  for (int i = tid; i < n; i += stride) gm[1 + i] = gid[list[i]];
}

// One CTA: remove the leavers (sorted lists a, b) by moving the staying tail atoms into
// the holes they leave in the head, then append the arrivals.  info[0] = n_own in/out.
template <typename T>
__global__ void __launch_bounds__(1024) k_dd_compact(int dim, int cap_own, int cap_mig, const int* list_a,
                                                     const int* list_b, const int* counters, const T* in_l,
                                                     const long long* gin_l, const T* in_r, const long long* gin_r,
                                                     T* R, T* P, T* F, long long* gid, int* scratch, int* info) {
  const int n_old = info[0];
  const int na = min(counters[0], cap_mig), nb = min(counters[1], cap_mig);
  int err = (counters[0] > cap_mig || counters[1] > cap_mig) ? JMD_DD_ELIST : 0;
  const int n_leave = na + nb, n_keep = n_old - n_leave;
  int n_in_l = (int)gin_l[0], n_in_r = (int)gin_r[0];
  int* flag = scratch;
  int* fillers = scratch + 2 * cap_mig;
  int* holes = scratch + 4 * cap_mig;
  auto leaver = [&](int t) { return t < na ? list_a[t] : list_b[t - na]; };
  for (int t = threadIdx.x; t < n_leave; t += blockDim.x) flag[t] = 0;
  __syncthreads();
  for (int t = threadIdx.x; t < n_leave; t += blockDim.x) {
    const int i = leaver(t);
    if (i >= n_keep) flag[i - n_keep] = 1;
  }
  __syncthreads();
  cta_compact(n_leave, [&](int t) { return flag[t] == 0; }, [&](int t) { return n_keep + t; }, fillers);
  const int nh = cta_compact(n_leave, [&](int t) { return leaver(t) < n_keep; }, leaver, holes);
  __syncthreads();
  for (int e = threadIdx.x; e < nh * dim; e += blockDim.x) {
    const int k = e / dim, c = e % dim;
    const size_t src = (size_t)fillers[k] * dim + c, dst = (size_t)holes[k] * dim + c;
    R[dst] = R[src];
    P[dst] = P[src];
    F[dst] = F[src];
  }
  for (int k = threadIdx.x; k < nh; k += blockDim.x) gid[holes[k]] = gid[fillers[k]];
  __syncthreads();     // the arrivals below overwrite the tail rows the fillers were read from
  int n_new = n_keep + n_in_l + n_in_r;
  if (n_new > cap_own) {
    err |= JMD_DD_ECAP;
    n_in_l = n_in_r = 0;
    n_new = n_keep;
  }
  const int row = 3 * dim;
  for (int e = threadIdx.x; e < (n_in_l + n_in_r) * row; e += blockDim.x) {
    const int i = e / row, c = e % row, arr = c / dim, cc = c % dim;
    const T v = i < n_in_l ? in_l[(size_t)i * row + c] : in_r[(size_t)(i - n_in_l) * row + c];
    T* dst = arr == 0 ? R : (arr == 1 ? P : F);
    dst[(size_t)(n_keep + i) * dim + cc] = v;
  }
  for (int i = threadIdx.x; i < n_in_l + n_in_r; i += blockDim.x)
    gid[n_keep + i] = i < n_in_l ? gin_l[1 + i] : gin_r[1 + i - n_in_l];
  __syncthreads();
  if (threadIdx.x == 0) {
    info[JMD_DD_N_OWN] = n_new;
    info[JMD_DD_ERROR] |= err;
    info[JMD_DD_MIG_L] = na;
    info[JMD_DD_MIG_R] = nb;
    info[JMD_DD_IN_L] = n_in_l;
    info[JMD_DD_IN_R] = n_in_r;
  }
}

// dst row 0 = [count, ...], rows 1.. = src[idx[i], :]
template <typename T>
__global__ void k_dd_pack_counted(int ncomp, int cap, const int* idx, const int* count, const T* src, T* dst) {
  const int n = min(*count, cap);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  if (tid == 0) dst[0] = (T)n;
  for (int e = tid; e < n * ncomp; e += stride) {
    const int i = e / ncomp, c = e % ncomp;
    dst[ncomp + e] = src[(size_t)idx[i] * ncomp + c];
  }
}

// Ghost rows behind the owned atoms: [from left | from right]; fills `info`.
template <typename T>
__global__ void k_dd_place(int dim, int cap_total, int cap_list, const int* counters, const T* recv_l,
                           const T* recv_r, T* R, int* info) {
  const int n_own = info[JMD_DD_N_OWN];
  const int nl = (int)recv_l[0], nr = (int)recv_r[0];
  const bool ok = n_own + nl + nr <= cap_total;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  if (ok) {
    for (int e = tid; e < (nl + nr) * dim; e += stride) {
      const int k = e / dim, c = e % dim;
      R[(size_t)(n_own + k) * dim + c] =
          k < nl ? recv_l[(size_t)(1 + k) * dim + c] : recv_r[(size_t)(1 + k - nl) * dim + c];
    }
  }
  if (tid == 0) {
    info[JMD_DD_FACE_L] = min(counters[0], cap_list);
    info[JMD_DD_FACE_R] = min(counters[1], cap_list);
    info[JMD_DD_FROM_L] = ok ? nl : 0;
    info[JMD_DD_FROM_R] = ok ? nr : 0;
    info[JMD_DD_ERROR] |= (ok ? 0 : JMD_DD_ECAP) | ((counters[0] > cap_list || counters[1] > cap_list) ? JMD_DD_ELIST : 0);
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

int jmd_dd_select(int dtype, int dim, int n, const int32_t* n_dev, const void* position, int axis, double lo, double L,
                  double thr_a, double thr_b, int32_t* list_a, int32_t* list_b, int32_t* counters, int cap,
                  void* stream) {
  if (!position || !list_a || !list_b || !counters || axis < 0 || axis >= dim) return JMD_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == JMD_F32)
    k_dd_select<float><<<blocks_for(n), 256, 0, s>>>(dim, n, n_dev, (const float*)position, axis, (float)lo, (float)L,
                                                     (float)thr_a, (float)thr_b, list_a, list_b, counters, cap);
  else if (dtype == JMD_F64)
    k_dd_select<double><<<blocks_for(n), 256, 0, s>>>(dim, n, n_dev, (const double*)position, axis, lo, L, thr_a, thr_b,
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

int jmd_dd_pack_migrate(int dtype, int dim, int cap_mig, const int32_t* list_a, const int32_t* list_b,
                        const int32_t* counters, const void* R, const void* P, const void* F, const int64_t* gid,
                        void* pay_a, void* pay_b, int64_t* gid_a, int64_t* gid_b, void* stream) {
  if (!list_a || !list_b || !counters || !R || !P || !F || !gid || !pay_a || !pay_b || !gid_a || !gid_b)
    return JMD_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  dim3 g(8, 2);
  if (dtype == JMD_F32)
    k_dd_pack_migrate<float><<<g, 256, 0, s>>>(dim, cap_mig, list_a, list_b, counters, (const float*)R, (const float*)P,
                                               (const float*)F, (const long long*)gid, (float*)pay_a, (float*)pay_b,
                                               (long long*)gid_a, (long long*)gid_b);
  else if (dtype == JMD_F64)
    k_dd_pack_migrate<double><<<g, 256, 0, s>>>(dim, cap_mig, list_a, list_b, counters, (const double*)R,
                                                (const double*)P, (const double*)F, (const long long*)gid,
                                                (double*)pay_a, (double*)pay_b, (long long*)gid_a, (long long*)gid_b);
  else return JMD_EINVAL;
  JMD_LAUNCH_CHECK();
  return 0;
}

int jmd_dd_compact(int dtype, int dim, int cap_own, int cap_mig, const int32_t* list_a, const int32_t* list_b,
                   const int32_t* counters, const void* in_l, const int64_t* gid_in_l, const void* in_r,
                   const int64_t* gid_in_r, void* R, void* P, void* F, int64_t* gid, int32_t* scratch, int32_t* info,
                   void* stream) {
  if (!list_a || !list_b || !counters || !in_l || !in_r || !gid_in_l || !gid_in_r || !R || !P || !F || !gid ||
      !scratch || !info)
    return JMD_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == JMD_F32)
    k_dd_compact<float><<<1, 1024, 0, s>>>(dim, cap_own, cap_mig, list_a, list_b, counters, (const float*)in_l,
                                           (const long long*)gid_in_l, (const float*)in_r, (const long long*)gid_in_r,
                                           (float*)R, (float*)P, (float*)F, (long long*)gid, scratch, info);
  else if (dtype == JMD_F64)
    k_dd_compact<double><<<1, 1024, 0, s>>>(dim, cap_own, cap_mig, list_a, list_b, counters, (const double*)in_l,
                                            (const long long*)gid_in_l, (const double*)in_r, (const long long*)gid_in_r,
                                            (double*)R, (double*)P, (double*)F, (long long*)gid, scratch, info);
  else return JMD_EINVAL;
  JMD_LAUNCH_CHECK();
  return 0;
}

int jmd_dd_pack_counted(int dtype, int ncomp, int cap, const int32_t* idx, const int32_t* count, const void* src,
                        void* dst, void* stream) {
  if (!idx || !count || !src || !dst || ncomp < 1) return JMD_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  const int g = blocks_for((long long)cap * ncomp);
  if (dtype == JMD_F32)
    k_dd_pack_counted<float><<<g, 256, 0, s>>>(ncomp, cap, idx, count, (const float*)src, (float*)dst);
  else if (dtype == JMD_F64)
    k_dd_pack_counted<double><<<g, 256, 0, s>>>(ncomp, cap, idx, count, (const double*)src, (double*)dst);
  else return JMD_EINVAL;
  JMD_LAUNCH_CHECK();
  return 0;
}

int jmd_dd_place(int dtype, int dim, int cap_total, int cap_list, const int32_t* counters, const void* recv_l,
                 const void* recv_r, void* R, int32_t* info, void* stream) {
  if (!counters || !recv_l || !recv_r || !R || !info) return JMD_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  const int g = blocks_for((long long)cap_list * dim * 2);
  if (dtype == JMD_F32)
    k_dd_place<float><<<g, 256, 0, s>>>(dim, cap_total, cap_list, counters, (const float*)recv_l,
                                        (const float*)recv_r, (float*)R, info);
  else if (dtype == JMD_F64)
    k_dd_place<double><<<g, 256, 0, s>>>(dim, cap_total, cap_list, counters, (const double*)recv_l,
                                         (const double*)recv_r, (double*)R, info);
  else return JMD_EINVAL;
  JMD_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"

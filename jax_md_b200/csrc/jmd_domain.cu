// Slab domain decomposition helpers: face / migration selection
// and gather-pack of halo and migration payloads.  The exchange itself is NCCL
// send/recv issued by the host on the same stream (jax_md_b200/domain.py).
#include <cuda_runtime.h>
#include <string.h>
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
    info[JMD_DD_N_LOC] = n_own + (ok ? nl + nr : 0);
    info[JMD_DD_N_ROWS] = n_own;
    info[JMD_DD_ERROR] |= (ok ? 0 : JMD_DD_ECAP) | ((counters[0] > cap_list || counters[1] > cap_list) ? JMD_DD_ELIST : 0);
  }
}

// ---- ordered selection (single pass, decoupled look-back) -----------------------------
// Tiles of 2048 atoms handed out by an atomic counter; a tile's running prefix packs both
// list counts into one 64-bit look-back word: [2 flag | 31 count_b | 31 count_a].
template <typename T>
__global__ void __launch_bounds__(256) k_dd_select_ordered(int dim, int n, const int* n_dev, const T* pos, int axis,
                                                           T lo, T L, T thr_a, T thr_b, int* list_a, int* list_b,
                                                           int* counters, int cap, unsigned long long* lb,
                                                           int tiles) {
  if (n_dev) n = min(n, *n_dev);
  __shared__ int s_tile;
  __shared__ unsigned long long s_pre;
  __shared__ unsigned long long wsum[8];
  constexpr unsigned long long F_AGG = 1ull << 62, F_PRE = 2ull << 62, MASK = (1ull << 62) - 1ull;
  unsigned long long* counter = lb + tiles;
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(counter, 1ull);
  __syncthreads();
  const int tile = s_tile;
  const T half = L * T(0.5);
  const int base = tile * 2048 + threadIdx.x * 8;
  unsigned fa = 0, fb = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int i = base + j;
    if (i < n) {
      T d = pos[(size_t)i * dim + axis] - lo;
      if (d >= half) d -= L;
      else if (d < -half) d += L;
      if (d < thr_a) fa |= 1u << j;
      if (d >= thr_b) fb |= 1u << j;
    }
  }
  // exclusive scan of (count_a, count_b) packed as a | b << 31 over the block
  const unsigned long long mine = (unsigned long long)__popc(fa) | ((unsigned long long)__popc(fb) << 31);
  unsigned long long incl = mine;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  unsigned long long wbase = 0, tot = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k < w) wbase += wsum[k];
    tot += wsum[k];
  }
  if (threadIdx.x == 0) {
    volatile unsigned long long* st = lb;
    unsigned long long run = 0;
    if (tile > 0) {
      st[tile] = F_AGG | tot;
      int p = tile - 1;
      while (true) {
        unsigned long long x;
        do { x = st[p]; } while ((x >> 62) == 0ull);
        run += x & MASK;
        if ((x >> 62) == 2ull) break;
        --p;
      }
    }
    __threadfence();
    st[tile] = F_PRE | (run + tot);
    s_pre = run;
    if (tile == tiles - 1) {            // totals (entries beyond `cap` are counted, not stored)
      const unsigned long long all = run + tot;
      counters[0] = (int)(all & 0x7fffffffull);
      counters[1] = (int)((all >> 31) & 0x7fffffffull);
    }
  }
  __syncthreads();
  const unsigned long long excl = s_pre + wbase + incl - mine;
  int pa = (int)(excl & 0x7fffffffull), pb = (int)((excl >> 31) & 0x7fffffffull);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (fa & (1u << j)) { if (pa < cap) list_a[pa] = base + j; ++pa; }
    if (fb & (1u << j)) { if (pb < cap) list_b[pb] = base + j; ++pb; }
  }
}

// ---- per-step exchange through peer memory ---------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// spins until *p >= want (or *p >> 1 >= want for flag words); false on timeout (~8 s)
__device__ __forceinline__ bool spin_until(const unsigned long long* p, unsigned long long want, int shift,
                                           unsigned long long* last) {
  const long long t0 = clock64();
  unsigned long long v;
  while (((v = ld_acquire_sys(p)) >> shift) < want) {
    if (clock64() - t0 > 16000000000ll) { *last = v; return false; }   // ~8 s
    __nanosleep(64);
  }
  *last = v;
  return true;
}

template <typename T>
struct DdP {
  int dim, rank, world, cap_list, always_rebuild;
  const int *face_l, *face_r, *face_counts;
  int* info;
  unsigned long long* epoch;
  unsigned int* ticket;
  const int* skin_blk;
  T* land;
  unsigned long long *signal, *flags;
  T *peer_land_l, *peer_land_r;
  unsigned long long *peer_signal_l, *peer_signal_r;
  unsigned long long* const* peer_flags;
  unsigned long long* host_flag;
};

template <typename T>
__global__ void __launch_bounds__(256) k_dd_comm_push(DdP<T> D, const T* R) {
  const unsigned long long ep = *D.epoch + 1ull;
  const int parity = (int)(ep & 1ull);
  const int dim = D.dim;
  const size_t side_rows = (size_t)D.cap_list * dim;
  // face atoms -> the neighbours' landing rows (we are the left neighbour's RIGHT side and
  // the right neighbour's LEFT side)
  const int nl = min(D.face_counts[0], D.cap_list), nr = min(D.face_counts[1], D.cap_list);
  const long long total = (long long)(nl + nr) * dim;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += stride) {
    const int k = (int)(e / dim), c = (int)(e % dim);
    if (k < nl) {
      T* dst = D.peer_land_l + ((size_t)parity * 2 + 1) * side_rows;
      dst[(size_t)k * dim + c] = R[(size_t)D.face_l[k] * dim + c];
    } else {
      T* dst = D.peer_land_r + ((size_t)parity * 2 + 0) * side_rows;
      dst[(size_t)(k - nl) * dim + c] = R[(size_t)D.face_r[k - nl] * dim + c];
    }
  }
  // this rank's rebuild flag -> every rank's flag word (block 0)
  if (blockIdx.x == 0) {
    const int n_own = D.info[JMD_DD_N_OWN];
    const int nblk = (n_own + 255) / 256;
    int moved = D.always_rebuild;
    for (int b = threadIdx.x; b < nblk; b += blockDim.x) moved |= (D.skin_blk[b] != 0);
    const int any = __syncthreads_or(moved);
    // (flag words are double-buffered by step parity like the landing rows: a rank that
    // runs one step ahead never overwrites a word a slower rank still has to read)
    if (threadIdx.x < D.world)
      st_release_sys(D.peer_flags[threadIdx.x] + (size_t)parity * D.world + D.rank,
                     (ep << 1) | (unsigned long long)(any ? 1 : 0));
  }
  // release the neighbours' signals once every block's rows are out
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&D.ticket[0], 1u);
    if (t == gridDim.x - 1) {
      D.ticket[0] = 0u;
      __threadfence_system();
      st_release_sys(D.peer_signal_l + 1, ep);
      st_release_sys(D.peer_signal_r + 0, ep);
    }
  }
}

template <typename T, int DIM>
__global__ void __launch_bounds__(256) k_dd_comm_wait(DdP<T> D, T* R, typename Vec4<T>::type* pos_sorted,
                                                      const int* inv_perm) {
  const unsigned long long ep = *D.epoch + 1ull;
  const int parity = (int)(ep & 1ull);
  __shared__ int ok_s;
  if (threadIdx.x == 0) {
    unsigned long long v;
    bool ok = spin_until(D.signal + 0, ep, 0, &v);
    ok = spin_until(D.signal + 1, ep, 0, &v) && ok;
    ok_s = ok ? 1 : 0;
    if (!ok) atomicOr(&D.info[JMD_DD_ERROR], JMD_DD_ETIMEOUT);
  }
  __syncthreads();
  const int n_own = D.info[JMD_DD_N_OWN], nl = D.info[JMD_DD_FROM_L], nr = D.info[JMD_DD_FROM_R];
  const size_t side_rows = (size_t)D.cap_list * DIM;
  const T* from_l = D.land + ((size_t)parity * 2 + 0) * side_rows;
  const T* from_r = D.land + ((size_t)parity * 2 + 1) * side_rows;
  if (ok_s) {
    const int stride = gridDim.x * blockDim.x;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nl + nr; k += stride) {
      const T* src = k < nl ? from_l + (size_t)k * DIM : from_r + (size_t)(k - nl) * DIM;
      T r[3] = {T(0), T(0), T(0)};
#pragma unroll
      for (int c = 0; c < DIM; ++c) {
        r[c] = src[c];
        R[(size_t)(n_own + k) * DIM + c] = r[c];
      }
      if (pos_sorted) {
        typename Vec4<T>::type v;
        v.x = r[0]; v.y = r[1]; v.z = r[2]; v.w = T(0);
        pos_sorted[inv_perm[n_own + k]] = v;
      }
    }
  }
  // global rebuild decision -> mapped host word (block 0); last block advances the epoch
  if (blockIdx.x == 0) {
    int any = 0;
    if (threadIdx.x < D.world) {
      unsigned long long v = 0;
      if (!spin_until(D.flags + (size_t)parity * D.world + threadIdx.x, ep, 1, &v))
        atomicOr(&D.info[JMD_DD_ERROR], JMD_DD_ETIMEOUT);
      any = (int)(v & 1ull);
    }
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) {
      st_release_sys(D.host_flag, (ep << 1) | (unsigned long long)(any ? 1 : 0));
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(&D.ticket[1], 1u);
    if (t == gridDim.x - 1) {
      D.ticket[1] = 0u;
      *D.epoch = ep;
    }
  }
}

template <typename T>
int fill_dd(DdP<T>& D, const jmd_dd_t* dd) {
  if (!dd || dd->world < 1 || dd->world > 256 || dd->cap_list < 1) return JMD_EINVAL;
  if (!dd->face_l || !dd->face_r || !dd->face_counts || !dd->info || !dd->epoch || !dd->ticket || !dd->land ||
      !dd->signal || !dd->flags || !dd->peer_land_l || !dd->peer_land_r || !dd->peer_signal_l ||
      !dd->peer_signal_r || !dd->peer_flags || !dd->host_flag || !dd->skin_blk)
    return JMD_EINVAL;
  D.dim = dd->dim; D.rank = dd->rank; D.world = dd->world; D.cap_list = dd->cap_list;
  D.always_rebuild = dd->always_rebuild;
  D.face_l = dd->face_l; D.face_r = dd->face_r; D.face_counts = dd->face_counts; D.info = dd->info;
  D.epoch = (unsigned long long*)dd->epoch; D.ticket = dd->ticket; D.skin_blk = dd->skin_blk;
  D.land = (T*)dd->land; D.signal = (unsigned long long*)dd->signal; D.flags = (unsigned long long*)dd->flags;
  D.peer_land_l = (T*)dd->peer_land_l; D.peer_land_r = (T*)dd->peer_land_r;
  D.peer_signal_l = (unsigned long long*)dd->peer_signal_l; D.peer_signal_r = (unsigned long long*)dd->peer_signal_r;
  D.peer_flags = (unsigned long long* const*)dd->peer_flags;
  D.host_flag = (unsigned long long*)dd->host_flag;
  return 0;
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

int jmd_dd_select_ordered(int dtype, int dim, int n, const int32_t* n_dev, const void* position, int axis, double lo,
                          double L, double thr_a, double thr_b, int32_t* list_a, int32_t* list_b,
                          int32_t* counters, int cap, uint64_t* scratch, void* stream) {
  if (!position || !list_a || !list_b || !counters || !scratch || axis < 0 || axis >= dim || n < 0) return JMD_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  const int tiles = n / 2048 + 1;
  cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(uint64_t) * (tiles + 1), s);
  if (e != cudaSuccess) return (int)e;
  if (dtype == JMD_F32)
    k_dd_select_ordered<float><<<tiles, 256, 0, s>>>(dim, n, n_dev, (const float*)position, axis, (float)lo, (float)L,
                                                     (float)thr_a, (float)thr_b, list_a, list_b, counters, cap,
                                                     (unsigned long long*)scratch, tiles);
  else if (dtype == JMD_F64)
    k_dd_select_ordered<double><<<tiles, 256, 0, s>>>(dim, n, n_dev, (const double*)position, axis, lo, L, thr_a,
                                                      thr_b, list_a, list_b, counters, cap,
                                                      (unsigned long long*)scratch, tiles);
  else return JMD_EINVAL;
  JMD_LAUNCH_CHECK();
  return 0;
}

int jmd_p2p_alloc(int64_t bytes, void** ptr, uint8_t* handle64) {
  if (!ptr || !handle64 || bytes <= 0) return JMD_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemset(*ptr, 0, (size_t)bytes);
  if (e != cudaSuccess) return (int)e;
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, *ptr);
  if (e != cudaSuccess) return (int)e;
  memcpy(handle64, &h, 64);
  return 0;
}

int jmd_p2p_open(const uint8_t* handle64, void** ptr) {
  if (!ptr || !handle64) return JMD_EINVAL;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  return (int)cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
}

int jmd_p2p_close(void* ptr) { return ptr ? (int)cudaIpcCloseMemHandle(ptr) : 0; }
int jmd_p2p_free(void* ptr) { return ptr ? (int)cudaFree(ptr) : 0; }

int jmd_host_flag_alloc(uint64_t** host_ptr, uint64_t** dev_ptr) {
  if (!host_ptr || !dev_ptr) return JMD_EINVAL;
  cudaError_t e = cudaHostAlloc((void**)host_ptr, 64, cudaHostAllocMapped | cudaHostAllocPortable);
  if (e != cudaSuccess) return (int)e;
  memset(*host_ptr, 0, 64);
  return (int)cudaHostGetDevicePointer((void**)dev_ptr, *host_ptr, 0);
}

int jmd_host_flag_free(uint64_t* host_ptr) { return host_ptr ? (int)cudaFreeHost(host_ptr) : 0; }

int jmd_dd_comm_push(const jmd_dd_t* dd, const void* R, void* stream) {
  if (!dd || !R) return JMD_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  const int g = blocks_for((long long)dd->cap_list * dd->dim * 2);
  int rc;
  if (dd->dtype == JMD_F32) {
    DdP<float> D;
    if ((rc = fill_dd(D, dd))) return rc;
    k_dd_comm_push<float><<<g, 256, 0, s>>>(D, (const float*)R);
  } else if (dd->dtype == JMD_F64) {
    DdP<double> D;
    if ((rc = fill_dd(D, dd))) return rc;
    k_dd_comm_push<double><<<g, 256, 0, s>>>(D, (const double*)R);
  } else return JMD_EINVAL;
  JMD_LAUNCH_CHECK();
  return 0;
}

int jmd_dd_comm_wait(const jmd_dd_t* dd, const jmd_nbr_t* nb, void* R, void* stream) {
  if (!dd || !R) return JMD_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  const int g = blocks_for((long long)dd->cap_list * 2);
  int rc;
#define JMD_WAIT(T, DIM)                                                                              \
  {                                                                                                   \
    DdP<T> D;                                                                                         \
    if ((rc = fill_dd(D, dd))) return rc;                                                             \
    k_dd_comm_wait<T, DIM><<<g, 256, 0, s>>>(D, (T*)R, nb ? (typename Vec4<T>::type*)nb->pos_sorted : nullptr, \
                                             nb ? nb->inv_perm : nullptr);                            \
  }
  if (dd->dtype == JMD_F32 && dd->dim == 3) JMD_WAIT(float, 3)
  else if (dd->dtype == JMD_F32 && dd->dim == 2) JMD_WAIT(float, 2)
  else if (dd->dtype == JMD_F64 && dd->dim == 3) JMD_WAIT(double, 3)
  else if (dd->dtype == JMD_F64 && dd->dim == 2) JMD_WAIT(double, 2)
  else return JMD_EINVAL;
#undef JMD_WAIT
  JMD_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"

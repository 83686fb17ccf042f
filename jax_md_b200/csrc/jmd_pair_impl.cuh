// Shared implementation of the pair-force kernels; included twice:
//   jmd_pair.cu          JMD_PAIR_STAGED 0  neighbour positions gathered from global memory
//   jmd_pair_staged.cu   JMD_PAIR_STAGED 1  ... from a per-block shared-memory staging buffer
//   jmd_pair_tric.cu     JMD_PAIR_TRIC 1    global gathers, minimum image through the full box matrix
//                                            (space.periodic_general with off-diagonal elements)
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <type_traits>
#include "jmd_common.cuh"

#ifndef JMD_PAIR_TRIC
#define JMD_PAIR_TRIC 0
#endif
#if JMD_PAIR_STAGED
#define JMD_PAIR_KERNEL k_pair_force_staged
#define JMD_PAIR_LAUNCHER launch_pair_staged_impl
#elif JMD_PAIR_TRIC
#define JMD_PAIR_KERNEL k_pair_force_tric
#define JMD_PAIR_LAUNCHER launch_pair_tric_impl
#else
#define JMD_PAIR_KERNEL k_pair_force
#define JMD_PAIR_LAUNCHER launch_pair_impl
#endif

namespace {

// Block size / residency of the global-gather kernel, measured on B200 (LJ, N=1M):
// 256 threads x 4 blocks (61 regs) 0.257 ms, 128 x 8 0.261 ms, 512 x 2 0.276 ms,
// 128 x 10 (48 regs) 0.247 ms.  The staged kernel is tied to JMD_STAGE_BLOCK.
#ifndef JMD_PAIR_BLOCK
#if JMD_PAIR_STAGED
#define JMD_PAIR_BLOCK JMD_STAGE_BLOCK
#else
#define JMD_PAIR_BLOCK 128
#endif
#endif
constexpr int PAIR_BLOCK = JMD_PAIR_BLOCK;
#ifndef JMD_PAIR_BATCH
#define JMD_PAIR_BATCH 4
#endif
constexpr int PAIR_BATCH = JMD_PAIR_BATCH;   // row entries in flight per loop trip (power of two)
#ifndef JMD_PAIR_IDXTMA
#define JMD_PAIR_IDXTMA 0
#endif
#ifndef JMD_PAIR_IDXCH
#define JMD_PAIR_IDXCH 16
#endif
constexpr int IDX_CH = JMD_PAIR_IDXCH;       // rows per TMA stage of the index stream
#ifndef JMD_PAIR_ALWAYS_WRAP
#define JMD_PAIR_ALWAYS_WRAP 1   /* measured: branch-free rint() form 2 % faster than the |d| > L/2 test */
#endif
#ifndef JMD_PAIR_MIN_BLOCKS
#define JMD_PAIR_MIN_BLOCKS 10
#endif

template <typename T, int DIM>
struct PairP {
  int n, m_int, n_rows;
  long long n_pad;
  Space<T, DIM> sp;
  const typename Vec4<T>::type* pos_sorted;
  const int* nl;
  const int* cnt;
  const int* perm;
  // potential
  int kind, has_cutoff, n_species, transposed, dparam_rows;
  int mode[3];
  T scalar[3];
  const T* array[3];
  T r_onset2, r_cutoff2, inv_denom;   // onset^2, cutoff^2, 1/(rc^2-ro^2)^3
  T r_cutoff, r_onset;
  // outputs
  T* force;
  T* e_atom;
  double* red;
  double* dparam;
  double* partials;
  // fused kick
  T* momentum;
  const T* mass;
  int mass_is_array;
  T dt_2;
  const T* dt_dev;
  // public-idx variant
  const int* idx;
  long long idx_m;
  const T* position;
  const int* species;
  // shared-memory staging (jmd_common.cuh)
  const unsigned short* nl16;
  const int* blk_table;
};

// read-only (non-coherent) 16/32-byte position gathers
__device__ __forceinline__ float4 ld_pos(const float4* p) { return __ldg(p); }
__device__ __forceinline__ double4 ld_pos(const double4* p) {
  const double2 a = __ldg(reinterpret_cast<const double2*>(p));
  const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}

// 16 / 32-byte shared-memory load at a 32-bit shared address (no generic->shared
// conversion per access)
template <typename T> __device__ __forceinline__ typename Vec4<T>::type lds_v4(unsigned addr);
template <> __device__ __forceinline__ float4 lds_v4<float>(unsigned addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
template <> __device__ __forceinline__ double4 lds_v4<double>(unsigned addr) {
  double4 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.z), "=d"(v.w) : "r"(addr + 16u));
  return v;
}

// 1/x: MUFU.RCP + one Newton step in f32 (~1 ulp, branch-free); IEEE in f64.
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y * (2.0f - x * y);
}
__device__ __forceinline__ double fast_rcp(double x) { return 1.0 / x; }
__device__ __forceinline__ float fast_rsqrt(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y * (1.5f - 0.5f * x * y * y);
}
__device__ __forceinline__ double fast_rsqrt(double x) { return 1.0 / sqrt(x); }

// U, (dU/dr)/r, dU/dsigma, dU/depsilon of the (switched) potential at r2.
// Written select-style (no data-dependent branches on the Lennard-Jones path)
// so the unrolled neighbour loop stays one basic block and its loads pipeline.
template <typename T, int POT, bool WANT_E>
__device__ __forceinline__ void pair_eval(int has_cutoff, T r2, T sigma, T eps, T alpha, T ro2, T rc2,
                                          T inv_denom, T& u, T& du_r, T& dus, T& due) {
  dus = T(0); due = T(0);
  // reference: distance() has zero gradient at r = 0 (util.safe_mask, space.py:246)
  const bool pos = r2 > T(0);
  bool live = pos;
  if (POT == JMD_POT_LJ) {
    if (has_cutoff) live = live && (r2 < rc2);
    const T ir2 = fast_rcp(r2);
    const T x2 = sigma * sigma * ir2;
    const T x6 = x2 * x2 * x2;
    const T x12 = x6 * x6;
    const T e4 = T(4) * eps;
    u = e4 * (x12 - x6);
    du_r = T(-6) * e4 * (T(2) * x12 - x6) * ir2;
    if (WANT_E) {
      dus = e4 * (T(12) * x12 - T(6) * x6) / sigma;
      due = T(4) * (x12 - x6);
    }
  } else if (POT == JMD_POT_SOFT_SPHERE) {
    const T ir = fast_rsqrt(r2);
    const T r = r2 * ir;
    const T x = r / sigma;
    live = live && (x < T(1));
    const T b = live ? T(1) - x : T(0);
    const T bm1 = (alpha == T(2)) ? b : ((alpha == T(2.5)) ? b * sqrt(b) : pow(b, alpha - T(1)));
    u = eps / alpha * bm1 * b;
    du_r = -(eps / sigma) * bm1 * ir;
    if (WANT_E) {
      dus = eps * bm1 * r / (sigma * sigma);
      due = bm1 * b / alpha;
      if (!pos) { u = eps / alpha; due = T(1) / alpha; dus = T(0); }   // r = 0: U = eps/alpha
    }
  } else {
    if (has_cutoff) live = live && (r2 < rc2);
    const T ir = fast_rsqrt(r2);
    const T r = pos ? r2 * ir : T(0);
    const T m = exp(-alpha * (r - sigma));
    const T om = T(1) - m;
    u = eps * om * om - eps;
    const T dudr = T(2) * eps * alpha * m * om;
    du_r = dudr * ir;
    if (WANT_E) {
      dus = -dudr;
      due = om * om - T(1);
    }
  }
  if (has_cutoff) {
    // energy.py:562-574: S = (rc2-r2)^2 (rc2 + 2 r2 - 3 ro2) / (rc2-ro2)^3 on [ro, rc)
    const bool sw = r2 >= ro2;
    const T a = rc2 - r2;
    const T ai = a * inv_denom;
    const T S = sw ? ai * a * (T(2) * r2 + (rc2 - T(3) * ro2)) : T(1);
    const T dS_r = sw ? T(12) * ai * (ro2 - r2) : T(0);                   // (dS/dr)/r
    du_r = dS_r * u + S * du_r;
    u = S * u;
    dus = S * dus;
    due = S * due;
  }
  du_r = live ? du_r : T(0);
  if (POT == JMD_POT_LJ) {
    u = live ? u : T(0); dus = live ? dus : T(0); due = live ? due : T(0);
  } else if (POT == JMD_POT_SOFT_SPHERE) {
    if (pos && !live) { u = T(0); dus = T(0); due = T(0); }
  } else {
    // Morse at r = 0 keeps its (finite) energy; beyond the cutoff everything is 0
    if (has_cutoff && !(r2 < rc2)) { u = T(0); dus = T(0); due = T(0); }
    if (!pos) dus = T(0);
  }
}

template <typename T, int DIM>
__device__ __forceinline__ T lookup(const PairP<T, DIM>& Q, int k, int ai, int aj, int si, int sj) {
  switch (Q.mode[k]) {
    case JMD_PARAM_SCALAR: return Q.scalar[k];
    case JMD_PARAM_PER_ATOM: return T(0.5) * (Q.array[k][ai] + Q.array[k][aj]);   // smap.py:836
    case JMD_PARAM_SPECIES:
      return Q.transposed ? Q.array[k][sj * Q.n_species + si] : Q.array[k][si * Q.n_species + sj];
    default:
      return Q.transposed ? Q.array[k][(size_t)aj * Q.n_species + ai] : Q.array[k][(size_t)ai * Q.n_species + aj];
  }
}

// RED: 0 none, 1 kick sums (KE, FF, PP, FP), 2 energy block + kick sums
template <int RED> struct RedN { static constexpr int value = RED == 0 ? 1 : (RED == 1 ? 4 : 13); };

template <typename T, int DIM, int POT, bool SCALAR, int RED, bool KICK>
// (the register cap of JMD_PAIR_MIN_BLOCKS residency is for the f32 force / kick
// variants; energy + virial + parameter-gradient and f64 variants need more registers)
__global__ void __launch_bounds__(PAIR_BLOCK, JMD_PAIR_STAGED ? ((sizeof(T) == 4 && RED < 2) ? 4 : 2)
                                                            : ((sizeof(T) == 4 && RED < 2) ? JMD_PAIR_MIN_BLOCKS : 4))
JMD_PAIR_KERNEL(PairP<T, DIM> Q) {
  using V4 = typename Vec4<T>::type;
#if JMD_PAIR_STAGED
  // Stage the positions of every atom in the 3^d stencils of this block's home
  // cells into shared memory: a few contiguous ranges of the cell-sorted array
  // (plan: ph_plan in jmd_neighbor.cu).
  extern __shared__ __align__(16) unsigned char stage_raw[];
  V4* const stage = reinterpret_cast<V4*>(stage_raw);
  __shared__ int tbl[JMD_TBL_INTS];
  if (threadIdx.x < JMD_TBL_INTS) tbl[threadIdx.x] = Q.blk_table[(size_t)blockIdx.x * JMD_TBL_INTS + threadIdx.x];
  __syncthreads();
  const bool staged = tbl[JMD_TBL_MODE] != 0;
  // TMA bulk copies (cp.async.bulk global -> shared, completion on an mbarrier):
  // one thread issues every range, nothing passes through registers, and the
  // copies of all ranges are in flight together.
  __shared__ __align__(8) unsigned long long stage_bar;
  if (staged) {
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&stage_bar);
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      const unsigned bytes = (unsigned)tbl[JMD_TBL_TOTAL] * (unsigned)sizeof(V4);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
#pragma unroll 1
      for (int e = 0; e < JMD_TBL_NENTRIES; ++e) {
        const int len = tbl[JMD_TBL_ENTRIES + 3 * e + 1];
        if (len <= 0) continue;
        const int g0 = tbl[JMD_TBL_ENTRIES + 3 * e], l0 = tbl[JMD_TBL_ENTRIES + 3 * e + 2];
        const unsigned dst = (unsigned)__cvta_generic_to_shared(stage + l0);
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
            ::"r"(dst), "l"(__cvta_generic_to_global(Q.pos_sorted + g0)), "r"((unsigned)len * (unsigned)sizeof(V4)), "r"(bar)
            : "memory");
      }
    }
    __syncthreads();                      // the barrier is initialised for everyone
    unsigned done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done) : "r"(bar) : "memory");
    }
  }
#endif
  constexpr bool WANT_E = RED == 2;
  constexpr int NV = RedN<RED>::value;
  const int t = blockIdx.x * PAIR_BLOCK + threadIdx.x;
  double rv[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) rv[i] = 0.0;

  const int ai = t < Q.n ? Q.perm[t] : 0x7fffffff;
  const bool valid = ai < Q.n_rows;         // ghosts (ids >= n_rows) have no row
  {
    const V4 pi = valid ? Q.pos_sorted[t] : V4();
    const int cnt = valid ? min(Q.cnt[t], Q.m_int) : 0;
    const int si = (int)pi.w;
    T f[3] = {T(0), T(0), T(0)};
    // energy / virial / parameter-gradient sums are carried in f64 per thread (the
    // force itself stays in the position dtype like the reference's)
    double e = 0.0, ds = 0.0, de = 0.0;
    double vir[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    // species-table gradients without atomics (jmd_pair_t.dparam_rows): this atom's sums per
    // NEIGHBOUR species; the host folds the rows of equal home species in a fixed order
    constexpr int DP = (WANT_E && !SCALAR) ? JMD_DPARAM_MAX_SPECIES : 1;
    double dps[DP], dpe[DP];
    const bool dp_rows = WANT_E && !SCALAR && Q.dparam && Q.dparam_rows;
    if (dp_rows) {
#pragma unroll
      for (int k = 0; k < DP; ++k) { dps[k] = 0.0; dpe[k] = 0.0; }
    }
    const int* col = Q.nl + t;
    const T sig0 = Q.scalar[0], eps0 = Q.scalar[1], alp0 = Q.scalar[2];
    // free space: half = +inf, never "far"
    const T hx = Q.sp.periodic ? Q.sp.half[0] : (T)INFINITY;
    const T hy = Q.sp.periodic ? Q.sp.half[1] : (T)INFINITY;
    const T hz = Q.sp.periodic ? Q.sp.half[DIM - 1] : (T)INFINITY;
#if JMD_PAIR_STAGED
    // has_cutoff is uniform: both specialisations live in the kernel and the loop
    // body is compiled with it as a constant
    auto run = [&](auto cut_tag) {
    constexpr int CUT = decltype(cut_tag)::value ? 1 : 0;
#else
    // (measured: with has_cutoff a compile-time constant ptxas hoists the whole
    // batch of loads above the arithmetic and the global-gather kernel gets 20 %
    // slower; the runtime test keeps loads and arithmetic interleaved)
    const int CUT = Q.has_cutoff;
#endif
    // one neighbour: displacement, potential, accumulation
    auto pair = [&](const int j, const V4& pj) {
      // minimum image (tolerance-level, handles unwrapped positions): a raw
      // difference within half a box side on every axis IS the minimum image --
      // the case for every pair of an atom away from the box faces -- so the
      // rint() form runs only behind a rarely taken branch.
      T d[3];
      d[0] = pi.x - pj.x;
      d[1] = pi.y - pj.y;
      d[2] = DIM == 3 ? pi.z - pj.z : T(0);
      bool far = fabs(d[0]) > hx || fabs(d[1]) > hy;
      if (DIM == 3) far = far || fabs(d[2]) > hz;
#if JMD_PAIR_TRIC
      if (far) Q.sp.wrap_tric(d);      // full-matrix box: d - H rint(H^-1 d); half[] holds the bounds
                                       // under which the raw difference already is that image
#else
      if (JMD_PAIR_ALWAYS_WRAP || far) {
        d[0] = Q.sp.wrap_fast(d[0], 0);
        d[1] = Q.sp.wrap_fast(d[1], 1);
        if (DIM == 3) d[2] = Q.sp.wrap_fast(d[2], DIM - 1);
      }
#endif
      const T r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
      T sigma = sig0, eps = eps0, alpha = alp0;
      if (!SCALAR) {
        const int sj = (int)pj.w;
        const int aj = ((Q.mode[0] | Q.mode[1] | Q.mode[2]) & 1) ? Q.perm[j] : 0;
        sigma = lookup(Q, 0, ai, aj, si, sj);
        eps = lookup(Q, 1, ai, aj, si, sj);
        alpha = lookup(Q, 2, ai, aj, si, sj);
      }
      T u, du_r, dus, due;
      pair_eval<T, POT, WANT_E>(CUT, r2, sigma, eps, alpha, Q.r_onset2, Q.r_cutoff2, Q.inv_denom,
                                u, du_r, dus, due);
      f[0] -= du_r * d[0];
      f[1] -= du_r * d[1];
      if (DIM == 3) f[2] -= du_r * d[2];
      if (WANT_E) {
        e += (double)u;
        ds += (double)dus;
        de += (double)due;
        vir[0] += (double)(du_r * d[0] * d[0]);
        vir[1] += (double)(du_r * d[1] * d[1]);
        vir[3] += (double)(du_r * d[0] * d[1]);
        if (DIM == 3) {
          vir[2] += (double)(du_r * d[2] * d[2]);
          vir[4] += (double)(du_r * d[0] * d[2]);
          vir[5] += (double)(du_r * d[1] * d[2]);
        }
        if (!SCALAR && Q.dparam) {
          const int sj = (int)pj.w;
          if (dp_rows) {
            const int b = sj < DP ? sj : DP - 1;        // (n_species <= DP is checked at launch)
            dps[b] += 0.5 * (double)dus;
            dpe[b] += 0.5 * (double)due;
          } else {
            const int cell = Q.transposed ? sj * Q.n_species + si : si * Q.n_species + sj;
            if (Q.mode[0] == JMD_PARAM_SPECIES) atomicAdd(&Q.dparam[cell], 0.5 * (double)dus);
            if (Q.mode[1] == JMD_PARAM_SPECIES)
              atomicAdd(&Q.dparam[Q.n_species * Q.n_species + cell], 0.5 * (double)due);
          }
        }
      }
    };
    const size_t np = (size_t)Q.n_pad;
#if JMD_PAIR_STAGED
    if (staged) {
      // 16-bit staging indices (streamed once, 64 B per warp and row); positions
      // from shared memory: a batch of indices, then the batch's ld.shared.v4,
      // then the arithmetic.  Full batches first: no bounds tests in the hot loop.
      const unsigned short* col16 = Q.nl16 + t;
      const unsigned sbase = (unsigned)__cvta_generic_to_shared(stage);
      const int full = cnt & ~(PAIR_BATCH - 1);
      int k = 0;
#pragma unroll 1
      for (; k < full; k += PAIR_BATCH) {
        unsigned cc[PAIR_BATCH];
        V4 pj[PAIR_BATCH];
#pragma unroll
        for (int u = 0; u < PAIR_BATCH; ++u) cc[u] = __ldcs(col16 + (size_t)(k + u) * np);
#pragma unroll
        for (int u = 0; u < PAIR_BATCH; ++u) pj[u] = lds_v4<T>(sbase + cc[u] * (unsigned)sizeof(V4));
#pragma unroll
        for (int u = 0; u < PAIR_BATCH; ++u) pair(0, pj[u]);
      }
#pragma unroll 1
      for (; k < cnt; ++k) {
        const unsigned c = __ldcs(col16 + (size_t)k * np);
        pair(0, lds_v4<T>(sbase + c * (unsigned)sizeof(V4)));
      }
    } else
#endif
    {
#if JMD_PAIR_IDXTMA && !JMD_PAIR_STAGED
      // The row entries are a pure stream (read once per step) and the kernel sits on
      // the L1 sector rate of the position gathers, so the stream is taken off the
      // LSU/L1 path: one thread issues TMA bulk copies (cp.async.bulk) of the block's
      // next IDX_CH rows (one contiguous 1 KB segment each) into a two-stage
      // shared-memory ring while the block evaluates the current ones.
      __shared__ __align__(128) int ibuf[2][IDX_CH][PAIR_BLOCK];
      __shared__ __align__(8) unsigned long long ibar[2];
      __shared__ int kmax_s;
      if (threadIdx.x == 0) {
        kmax_s = 0;
        for (int b = 0; b < 2; ++b)
          asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&ibar[b])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      }
      __syncthreads();
      {
        int wmax = cnt;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
        if ((threadIdx.x & 31) == 0 && wmax > 0) atomicMax(&kmax_s, wmax);
      }
      __syncthreads();
      const int kmax = kmax_s;
      const int nch = (kmax + IDX_CH - 1) / IDX_CH;
      const long long blk0 = (long long)blockIdx.x * PAIR_BLOCK;
      const int cols = (int)min((long long)PAIR_BLOCK, Q.n_pad - blk0);     // ints per row segment
      auto issue = [&](int c) {            // thread 0: rows [c*IDX_CH, ...) -> ibuf[c & 1]
        const int rows = min(IDX_CH, kmax - c * IDX_CH);
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&ibar[c & 1]);
        const unsigned bytes = (unsigned)cols * 4u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes * rows) : "memory");
        const int* src = Q.nl + (size_t)(c * IDX_CH) * np + blk0;
        for (int r = 0; r < rows; ++r) {
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
              ::"r"((unsigned)__cvta_generic_to_shared(&ibuf[c & 1][r][0])),
                "l"(__cvta_generic_to_global(src + (size_t)r * np)), "r"(bytes), "r"(bar)
              : "memory");
        }
      };
      if (threadIdx.x == 0 && nch > 0) issue(0);
      for (int c = 0; c < nch; ++c) {
        if (threadIdx.x == 0 && c + 1 < nch) issue(c + 1);   // its buffer was released by the barrier below
        {
          const unsigned bar = (unsigned)__cvta_generic_to_shared(&ibar[c & 1]);
          const unsigned parity = (unsigned)(c >> 1) & 1u;
          unsigned done = 0;
          while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done) : "r"(bar), "r"(parity) : "memory");
          }
        }
        const int kend = min(cnt - c * IDX_CH, IDX_CH);
        const int* rowp = &ibuf[c & 1][0][threadIdx.x];
#pragma unroll(PAIR_BATCH)
        for (int r = 0; r < kend; ++r) {
          const int j = rowp[r * PAIR_BLOCK];
          pair(j, ld_pos(&Q.pos_sorted[j]));
        }
        __syncthreads();
      }
#else
#pragma unroll(PAIR_BATCH)
      for (int k = 0; k < cnt; ++k) {
        const int j = __ldcs(col + (size_t)k * np);        // streamed once: keep it out of L1
        pair(j, ld_pos(&Q.pos_sorted[j]));
      }
#endif
    }
#if JMD_PAIR_STAGED
    };   // run
    if (Q.has_cutoff) run(std::true_type()); else run(std::false_type());
#endif
    if (valid) {
    T* fo = Q.force + (size_t)ai * DIM;
#pragma unroll
    for (int k = 0; k < DIM; ++k) fo[k] = f[k];
    if (WANT_E) {
      if (Q.e_atom) Q.e_atom[ai] = (T)(0.5 * e);        // smap.py:955-958: / normalization
      if (!SCALAR && Q.dparam) {
        if (dp_rows) {
          const int S = Q.n_species;
          for (int k = 0; k < S && k < DP; ++k) {
            Q.dparam[(size_t)ai * S + k] = dps[k];
            Q.dparam[((size_t)Q.n + ai) * S + k] = dpe[k];
          }
        } else {
          if (Q.mode[0] == JMD_PARAM_PER_ATOM) Q.dparam[ai] = 0.5 * ds;
          if (Q.mode[1] == JMD_PARAM_PER_ATOM) Q.dparam[Q.n + ai] = 0.5 * de;
        }
      }
    }
    T ke = T(0), pp = T(0), fp = T(0), ff = T(0);
    if (KICK) {
      T* po = Q.momentum + (size_t)ai * DIM;
      const T m = Q.mass_is_array ? Q.mass[ai] : Q.mass[0];
      // FIRE passes a traced dt (minimize.py:185): dt_2 = f32(f32(dt) / 2)
      const T dt_2 = Q.dt_dev ? (T)(float)((T)(float)(*Q.dt_dev) / T(2)) : Q.dt_2;
#pragma unroll
      for (int k = 0; k < DIM; ++k) {
        T p = po[k] + dt_2 * f[k];                  // simulate.py:168-173
        po[k] = p;
        ke += p * p / m;                              // quantity.py:152
        pp += p * p;
        fp += f[k] * p;
        ff += f[k] * f[k];
      }
    }
    if (RED >= 1) {
      const int o = RED == 2 ? 9 : 0;
      rv[o + 0] = 0.5 * (double)ke;
      rv[o + 1] = (double)ff;
      rv[o + 2] = (double)pp;
      rv[o + 3] = (double)fp;
    }
    if (RED == 2) {
      rv[0] = 0.5 * e;
#pragma unroll
      for (int k = 0; k < 6; ++k) rv[1 + k] = 0.5 * vir[k];
      rv[7] = 0.5 * ds;
      rv[8] = 0.5 * de;
    }
    }   // valid
  }
  if (RED >= 1) {
    __shared__ double sm[NV * (PAIR_BLOCK / 32)];
    __shared__ int slots[NV];
    if (threadIdx.x == 0) {
      if (RED == 2) {
        slots[0] = JMD_RED_ENERGY;
        for (int k = 0; k < 6; ++k) slots[1 + k] = JMD_RED_VIRIAL + k;
        slots[7] = JMD_RED_DSIGMA;
        slots[8] = JMD_RED_DEPSILON;
      }
      const int o = RED == 2 ? 9 : 0;
      slots[o + 0] = JMD_RED_KINETIC;
      slots[o + 1] = JMD_RED_FF;
      slots[o + 2] = JMD_RED_PP;
      slots[o + 3] = JMD_RED_FP;
    }
    __syncthreads();
    grid_reduce_finish<NV, PAIR_BLOCK>(rv, Q.partials + 2, (unsigned int*)Q.partials, Q.red, slots, sm);
  }
}

#if JMD_PAIR_STAGED
// dynamic shared memory above 48 KB needs the opt-in, once per kernel
#define JMD_PAIR_LAUNCH(R, K)                                                                   \
  do {                                                                                          \
    static int optin[64] = {0};                                                                 \
    auto kern = JMD_PAIR_KERNEL<T, DIM, POT, SCALAR, R, K>;                                     \
    int dev = 0;                                                                                \
    cudaGetDevice(&dev);                                                                        \
    if (dev >= 64 || !optin[dev]) {                                                             \
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, JMD_STAGE_BYTES); \
      if (dev < 64) optin[dev] = 1;                                                             \
    }                                                                                           \
    kern<<<grid, PAIR_BLOCK, JMD_STAGE_BYTES, s>>>(Q);                                          \
  } while (0)
#else
#define JMD_PAIR_LAUNCH(R, K) JMD_PAIR_KERNEL<T, DIM, POT, SCALAR, R, K><<<grid, PAIR_BLOCK, 0, s>>>(Q)
#endif

template <typename T, int DIM, int POT, bool SCALAR>
int launch_variants(const PairP<T, DIM>& Q, bool want_e, bool kick, cudaStream_t s) {
  const int grid = (int)jmd_div_up(Q.n > 0 ? Q.n : 1, PAIR_BLOCK);
  if (want_e) {
    if (kick) JMD_PAIR_LAUNCH(2, true);
    else JMD_PAIR_LAUNCH(2, false);
  } else if (kick) {
    JMD_PAIR_LAUNCH(1, true);
  } else {
    JMD_PAIR_LAUNCH(0, false);
  }
  JMD_LAUNCH_CHECK();
  return 0;
}

template <typename T, int DIM>
int JMD_PAIR_LAUNCHER(const jmd_nbr_t* nb, const jmd_pair_t* pp, void* force, void* e_atom, double* red,
                double* dparam, double* partials, void* momentum, const void* mass, int mass_is_array,
                double dt_2, const void* dt_dev, bool want_e, cudaStream_t s) {
  PairP<T, DIM> Q;
  Q.n = nb->n; Q.m_int = nb->m_int; Q.n_pad = nb->n_pad;
  Q.n_rows = (nb->n_rows > 0 && nb->n_rows < nb->n) ? nb->n_rows : nb->n;
  Q.sp.init(nb->space);
  Q.pos_sorted = (const typename Vec4<T>::type*)nb->pos_sorted;
  Q.nl = nb->nl; Q.cnt = nb->cnt; Q.perm = nb->perm;
  Q.kind = pp->kind; Q.has_cutoff = pp->has_cutoff; Q.n_species = pp->n_species;
  Q.transposed = pp->transposed;
  Q.dparam_rows = pp->dparam_rows;
  if (Q.dparam_rows && (pp->n_species > JMD_DPARAM_MAX_SPECIES || pp->n_species < 1)) return JMD_EINVAL;
  bool scalar = true;
  for (int k = 0; k < 3; ++k) {
    Q.mode[k] = pp->mode[k];
    Q.scalar[k] = (T)pp->scalar[k];
    Q.array[k] = (const T*)pp->array[k];
    if (pp->mode[k] != JMD_PARAM_SCALAR) {
      scalar = false;
      if (!pp->array[k]) return JMD_EINVAL;
    }
  }
  T ro = (T)pp->r_onset, rc = (T)pp->r_cutoff;
  Q.r_onset = ro; Q.r_cutoff = rc;
  Q.r_onset2 = (T)pp->r_onset2; Q.r_cutoff2 = (T)pp->r_cutoff2;
  T den3 = (T)pp->switch_denom;
  Q.inv_denom = pp->has_cutoff ? T(1) / den3 : T(0);
  Q.force = (T*)force; Q.e_atom = (T*)e_atom; Q.red = red; Q.dparam = dparam; Q.partials = partials;
  Q.momentum = (T*)momentum; Q.mass = (const T*)mass; Q.mass_is_array = mass_is_array; Q.dt_2 = (T)dt_2;
  Q.dt_dev = (const T*)dt_dev;
  Q.idx = nullptr; Q.idx_m = 0; Q.position = nullptr; Q.species = nullptr;
  Q.nl16 = nb->nl16; Q.blk_table = nb->blk_table;
  const bool kick = momentum != nullptr;
  if ((kick || want_e) && (!red || !partials)) return JMD_EINVAL;
  if (kick && !mass) return JMD_EINVAL;
  if (!force) return JMD_EINVAL;
#define JMD_POT_CASE(POT)                                                              \
  case POT:                                                                            \
    return scalar ? launch_variants<T, DIM, POT, true>(Q, want_e, kick, s)             \
                  : launch_variants<T, DIM, POT, false>(Q, want_e, kick, s);
  switch (pp->kind) {
    JMD_POT_CASE(JMD_POT_LJ)
    JMD_POT_CASE(JMD_POT_SOFT_SPHERE)
    JMD_POT_CASE(JMD_POT_MORSE)
    default: return JMD_EINVAL;
  }
#undef JMD_POT_CASE
}

}  // namespace

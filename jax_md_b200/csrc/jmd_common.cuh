// Shared device helpers for libjmd_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "jmd_b200.h"

#define JMD_WARP 32
#define JMD_SM_COUNT 148

#define JMD_LAUNCH_CHECK()                         \
  do {                                             \
    cudaError_t e__ = cudaGetLastError();          \
    if (e__ != cudaSuccess) return (int)e__;       \
  } while (0)

// ---- shared-memory staging of neighbour positions (force kernel) -----------------
// A force-kernel block owns JMD_STAGE_BLOCK consecutive cell-sorted slots.  With cells
// stored x-fastest, the 3^d stencils of its home cells are a few contiguous slot
// ranges: <= JMD_STAGE_SEGS row segments x 9 (dy, dz) stencil rows x 2 pieces (periodic
// wrap in x).
// The neighbour build writes that plan into a per-block table and a second copy of
// the neighbour rows as 16-bit indices into the block's staging buffer; the force
// kernel copies the ranges into shared memory once and gathers from there.
#define JMD_STAGE_BLOCK 256
#define JMD_STAGE_BYTES 57344          /* 3584 float4 / 1792 double4; 4 blocks per SM */
#define JMD_STAGE_SEGS 4               /* consecutive cell rows a block may span */
#define JMD_TBL_INTS 256               /* per-block table stride */
#define JMD_TBL_MODE 0                 /* 1: staged, 0: gather from global (32-bit rows) */
#define JMD_TBL_TOTAL 1                /* staged entries */
#define JMD_TBL_ROW0 2                 /* row id (cell / cells_x) of segment 0; segment s is row ROW0 + s */
#define JMD_TBL_NSEG 3
#define JMD_TBL_ENTRIES 8              /* then SEGS x 18 x {first slot, length, first staging index} */
#define JMD_TBL_NENTRIES (JMD_STAGE_SEGS * 18)
#define JMD_TBL_ENTRY(seg, dy, dz, piece) (JMD_TBL_ENTRIES + 3 * (((((seg) * 3 + (dy) + 1) * 3 + (dz) + 1) * 2) + (piece)))

template <typename T> struct Vec4;
template <> struct Vec4<float> { typedef float4 type; };
template <> struct Vec4<double> { typedef double4 type; };

// ---- separately rounded IEEE ops: the compiler must not contract these into
// FMAs (the oracle / reference spell the ops out; SURVEY 7 hard-part 1). -------
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

// jnp.mod(t, L) for L > 0 (space.py:224,252): fmod, then +L when the remainder
// is negative.  The three fast branches are exact restatements of that for
// t in (-L, 2L); everything else takes the fmod path.
template <typename T>
__device__ __forceinline__ T mod_pos(T t, T L) {
  if (t >= T(0) && t < L) return t;
  if (t >= L && t < L + L) return sub_rn(t, L);          // exact (Sterbenz)
  if (t < T(0) && t > -L) return add_rn(t, L);           // fmod(t, L) == t
  T m = fmod(t, L);
  if (m != T(0) && m < T(0)) m = add_rn(m, L);
  return m;
}

// round-to-nearest-integer by the magic-constant trick (2 full-rate adds, no
// branch): valid for |x| < 2^22 (f32) / 2^51 (f64).
__device__ __forceinline__ float rint_magic(float x) {
  return __fadd_rn(__fadd_rn(x, 12582912.0f), -12582912.0f);
}
__device__ __forceinline__ double rint_magic(double x) {
  return __dadd_rn(__dadd_rn(x, 6755399441055744.0), -6755399441055744.0);
}

// Full-matrix box of space.periodic_general (row-major H: real_i = sum_j H[i][j] frac_j) and its
// inverse.  The exact metric for it is deliberately NOT inlined into the neighbour-list kernels:
// inlined, it pushed the Dense stencil scan past ptxas' inlining budget, the kernel-parameter
// struct got materialised in local memory and the scan went from 0.8 to 4.8 ms for ORTHORHOMBIC
// boxes.  As a call it costs the common path one uniform branch.
template <typename T, int DIM>
struct TricM {
  T H[DIM * DIM];
  T Hi[DIM * DIM];
  int frac;
};

template <typename T, int DIM>
__device__ __forceinline__ void tric_matvec(const T* M, const T* v, T* out) {
  // space.raw_transform's einsum (space.py:128-150): XLA does not pin the summation order; here
  // (and in the oracle) it is j = 0, 1, 2 with separately rounded products and sums
  if (DIM == 2) {
    out[0] = add_rn(mul_rn(M[0], v[0]), mul_rn(M[1], v[1]));
    out[1] = add_rn(mul_rn(M[2], v[0]), mul_rn(M[3], v[1]));
  } else {
    constexpr int Q = DIM * DIM, Z = 2 % DIM;      // (keeps indices in range where DIM == 2 is compiled)
    out[0] = add_rn(add_rn(mul_rn(M[0], v[0]), mul_rn(M[1], v[1])), mul_rn(M[2 % Q], v[Z]));
    out[1] = add_rn(add_rn(mul_rn(M[3], v[0]), mul_rn(M[4 % Q], v[1])), mul_rn(M[5 % Q], v[Z]));
    out[Z] = add_rn(add_rn(mul_rn(M[6 % Q], v[0]), mul_rn(M[7 % Q], v[1])), mul_rn(M[8 % Q], v[Z]));
  }
}

// exact squared distance of periodic_general with a matrix box (space.py:419-433):
// |H (mod(ua - ub + 1/2, 1) - 1/2)|^2, ua = a (unit cube) or H^-1 a
template <typename T, int DIM>
__device__ __noinline__ T dist2_tric_call(TricM<T, DIM> m, T a0, T a1, T a2, T b0, T b1, T b2) {
  const T a[3] = {a0, a1, a2}, b[3] = {b0, b1, b2};
  T ua[3] = {a0, a1, a2}, ub[3] = {b0, b1, b2}, w[3] = {T(0), T(0), T(0)}, g[3] = {T(0), T(0), T(0)};
  if (!m.frac) {
    tric_matvec<T, DIM>(m.Hi, a, ua);
    tric_matvec<T, DIM>(m.Hi, b, ub);
  }
#pragma unroll
  for (int k = 0; k < DIM; ++k)
    w[k] = sub_rn(mod_pos(add_rn(sub_rn(ua[k], ub[k]), T(0.5)), T(1)), T(0.5));
  tric_matvec<T, DIM>(m.H, w, g);
  T acc = mul_rn(g[0], g[0]);
#pragma unroll
  for (int k = 1; k < DIM; ++k) acc = add_rn(acc, mul_rn(g[k], g[k]));
  return acc;
}

template <typename T, int DIM>
struct Space {
  T side[DIM];
  T half[DIM];
  T quarter[DIM];    // half / 2: |d| <= quarter takes the exact fast path
  T inv_side[DIM];   // 1 / side, or 0 for free space (then disp_fast is a no-op)
  T side_f[DIM];     // side, or 0 for free space
  int periodic;
  int wrapped;
  int general;       // space.periodic_general, orthorhombic (see jmd_space_t)
  int frac;          // ... with positions stored in the unit cube
  T ibox[DIM];       // 1 / box
  int tric;          // ... with a full box matrix: tm.H and its inverse replace side / ibox
  TricM<T, DIM> tm;
  __host__ void init(const jmd_space_t& s) {
    periodic = s.kind == JMD_SPACE_PERIODIC;
    wrapped = s.wrapped;
    general = periodic && s.general;
    frac = general && s.fractional;
    tric = general && s.triclinic;
    for (int i = 0; i < DIM; ++i)
      for (int j = 0; j < DIM; ++j) {
        tm.H[i * DIM + j] = tric ? (T)s.box_m[i * 3 + j] : T(i == j);
        tm.Hi[i * DIM + j] = tric ? (T)s.inv_box_m[i * 3 + j] : T(i == j);
      }
    tm.frac = frac;
    for (int k = 0; k < DIM; ++k) ibox[k] = general ? (T)s.inv_box[k] : T(1);
    for (int k = 0; k < DIM; ++k) {
      side[k] = (T)s.side[k];
      half[k] = (T)s.half[k];
      quarter[k] = periodic ? (T)(s.half[k] * 0.5) : (T)0;
      inv_side[k] = periodic && s.side[k] != 0 ? (T)(1.0 / s.side[k]) : (T)0;
      side_f[k] = periodic ? (T)s.side[k] : (T)0;
    }
  }
  // Exact displacement component d(a, b)_k (space.py:213-224):
  //   mod(fl(a - b) + h, L) - h.
  // For |a - b| <= h/2 the sum t = fl(d + h) lies in [h/2, 3h/2], strictly inside
  // (0, L), so jnp.mod returns t itself and the result is fl(t - h): two adds.
  __device__ __forceinline__ T disp(T a, T b, int k) const {
    T d = sub_rn(a, b);
    if (periodic) {
      T t = add_rn(d, half[k]);
      if (fabs(d) <= quarter[k]) return sub_rn(t, half[k]);
      d = sub_rn(mod_pos(t, side[k]), half[k]);
    }
    return d;
  }
  // Branch-free minimum image for the force kernels (tolerance-level, not
  // bit-level; handles unwrapped positions): d - L * rint(d / L).
  __device__ __forceinline__ T disp_fast(T a, T b, int k) const {
    T d = a - b;
    T r = rint_magic(d * inv_side[k]);
    return d - side_f[k] * r;
  }
  // the same on an already formed difference
  __device__ __forceinline__ T wrap_fast(T d, int k) const {
    T r = rint_magic(d * inv_side[k]);
    return d - side_f[k] * r;
  }
  // shift_fn (space.py:250-252 / 268-270)
  __device__ __forceinline__ T shift(T r, T dr, int k) const {
    if (general) return shift_general(r, dr, k);
    T s = add_rn(r, dr);
    if (periodic && wrapped) s = mod_pos(s, side[k]);
    return s;
  }
  // periodic_general (space.py:437-470): dR -> inv_box * dR, the update happens on the unit
  // cube (R itself is transformed there and back when it is stored in real space)
  __device__ __forceinline__ T shift_general(T r, T dr, int k) const {
    const T du = mul_rn(dr, ibox[k]);
    if (!frac && !wrapped) return add_rn(r, dr);
    T u = frac ? r : mul_rn(r, ibox[k]);
    u = add_rn(u, du);
    if (wrapped) u = mod_pos(u, T(1));
    return frac ? u : mul_rn(u, side[k]);
  }
  // exact displacement component of periodic_general (space.py:419-433):
  // box * (mod(ua - ub + 1/2, 1) - 1/2), ua = a (fractional) or inv_box * a
  __device__ __forceinline__ T disp_general(T a, T b, int k) const {
    const T ua = frac ? a : mul_rn(a, ibox[k]);
    const T ub = frac ? b : mul_rn(b, ibox[k]);
    const T d = sub_rn(ua, ub);
    const T m = sub_rn(mod_pos(add_rn(d, T(0.5)), T(1)), T(0.5));
    return mul_rn(m, side[k]);
  }
  // fractional -> real (what the cell-sorted copy holds)
  __device__ __forceinline__ T to_real(T r, int k) const { return frac ? mul_rn(r, side[k]) : r; }

  // ---- full-matrix boxes: all components at once (drift kernel, packing) -----------------
  __device__ __forceinline__ void to_real_v(const T* r, T* out) const {
    if (tric && frac) { tric_matvec<T, DIM>(tm.H, r, out); return; }
#pragma unroll
    for (int k = 0; k < DIM; ++k) out[k] = to_real(r[k], k);
  }
  __device__ __forceinline__ void shift_v(const T* r, const T* dr, T* out) const {
    if (!tric) {
#pragma unroll
      for (int k = 0; k < DIM; ++k) out[k] = shift(r[k], dr[k], k);
      return;
    }
    if (!frac && !wrapped) {
#pragma unroll
      for (int k = 0; k < DIM; ++k) out[k] = add_rn(r[k], dr[k]);
      return;
    }
    T du[DIM], u[DIM];
    tric_matvec<T, DIM>(tm.Hi, dr, du);
    if (frac) {
#pragma unroll
      for (int k = 0; k < DIM; ++k) u[k] = r[k];
    } else {
      tric_matvec<T, DIM>(tm.Hi, r, u);
    }
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
      u[k] = add_rn(u[k], du[k]);
      if (wrapped) u[k] = mod_pos(u[k], T(1));
    }
    if (frac) {
#pragma unroll
      for (int k = 0; k < DIM; ++k) out[k] = u[k];
    } else {
      tric_matvec<T, DIM>(tm.H, u, out);
    }
  }
  // minimum image of a real-space difference, tolerance-level (force kernels):
  // d - H rint(H^-1 d)
  __device__ __forceinline__ void wrap_tric(T* d) const {
    if (DIM == 2) {
      const T f0 = rint_magic(tm.Hi[0] * d[0] + tm.Hi[1] * d[1]);
      const T f1 = rint_magic(tm.Hi[2] * d[0] + tm.Hi[3] * d[1]);
      d[0] -= tm.H[0] * f0 + tm.H[1] * f1;
      d[1] -= tm.H[2] * f0 + tm.H[3] * f1;
    } else {
      constexpr int Q = DIM * DIM, Z = 2 % DIM;      // (keeps the indices in range when DIM == 2 is compiled)
      const T f0 = rint_magic(tm.Hi[0] * d[0] + tm.Hi[1] * d[1] + tm.Hi[2 % Q] * d[Z]);
      const T f1 = rint_magic(tm.Hi[3] * d[0] + tm.Hi[4 % Q] * d[1] + tm.Hi[5 % Q] * d[Z]);
      const T f2 = rint_magic(tm.Hi[6 % Q] * d[0] + tm.Hi[7 % Q] * d[1] + tm.Hi[8 % Q] * d[Z]);
      d[0] -= tm.H[0] * f0 + tm.H[1] * f1 + tm.H[2 % Q] * f2;
      d[1] -= tm.H[3] * f0 + tm.H[4 % Q] * f1 + tm.H[5 % Q] * f2;
      d[Z] -= tm.H[6 % Q] * f0 + tm.H[7 % Q] * f1 + tm.H[8 % Q] * f2;
    }
  }
};

// Exact squared distance sum_k d_k^2, sequential (space.py:227-235).
// TRIC_OK = false compiles the full-matrix branch out: the pre-filtered stencil scans never see a
// triclinic box (their filter is off for it), and even as a rarely taken CALL the branch cost the
// Dense scan its register allocation (0.8 -> 4.9 ms at N = 1M).
template <typename T, int DIM, bool TRIC_OK = true>
__device__ __forceinline__ T dist2_exact(const Space<T, DIM>& sp, const T* a, const T* b) {
  if (TRIC_OK && sp.tric)
    return dist2_tric_call<T, DIM>(sp.tm, a[0], a[1], DIM == 3 ? a[DIM - 1] : T(0), b[0], b[1],
                                   DIM == 3 ? b[DIM - 1] : T(0));
  if (sp.general) {
    T g = sp.disp_general(a[0], b[0], 0);
    T acc = mul_rn(g, g);
#pragma unroll
    for (int k = 1; k < DIM; ++k) {
      g = sp.disp_general(a[k], b[k], k);
      acc = add_rn(acc, mul_rn(g, g));
    }
    return acc;
  }
  T dx = sp.disp(a[0], b[0], 0);
  T acc = mul_rn(dx, dx);
#pragma unroll
  for (int k = 1; k < DIM; ++k) {
    T d = sp.disp(a[k], b[k], k);
    acc = add_rn(acc, mul_rn(d, d));
  }
  return acc;
}

// ---- block reduction of NV doubles, deterministic ----------------------------
template <int NV, int BLOCK>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double* smem /*[NV*BLOCK/32]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    if (lane == 0) smem[i * (BLOCK / 32) + warp] = v[i];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double x = lane < BLOCK / 32 ? smem[i * (BLOCK / 32) + lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      v[i] = x;
    }
  }
  __syncthreads();
}

// Per-block partials -> final sum by the last block to finish (fixed order, so
// results are bitwise reproducible).  partials: [gridDim.x, NV] then a ticket.
template <int NV, int BLOCK>
__device__ __forceinline__ void grid_reduce_finish(double (&v)[NV], double* partials,
                                                   unsigned int* ticket, double* out,
                                                   const int* slots, double* smem) {
  block_reduce<NV, BLOCK>(v, smem);
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) partials[(size_t)blockIdx.x * NV + i] = v[i];
    __threadfence();
    unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.0;
  for (unsigned int b = threadIdx.x; b < gridDim.x; b += BLOCK) {
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] += __ldcg(&partials[(size_t)b * NV + i]);
  }
  block_reduce<NV, BLOCK>(acc, smem);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) out[slots[i]] = acc[i];
    *ticket = 0u;
  }
}

__host__ inline int64_t jmd_div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

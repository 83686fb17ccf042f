// Fused neighbour-list pair force / energy / virial / parameter-gradient kernel.
//
// Replaces smap.pair_neighbor_list (smap.py:922-979) + jax.grad of it
// (quantity.py:58-60, a gather fwd + scatter-add bwd in XLA) for the three pair
// potentials energy.py:125-173 (soft sphere), :246-272 (Lennard-Jones),
// :346-371 (Morse) with energy.py:534-580's multiplicative cutoff.
//
// One thread per atom in cell-sorted slot order, full (both-direction) rows, no
// atomics: row entries are read coalesced from the transposed list
// nl[k][slot]; neighbour positions are gathered as one 16-byte float4 each from
// the cell-sorted copy (L1/L2 resident: adjacent threads share most
// neighbours).  Optionally fuses the second velocity-Verlet half kick
// (simulate.py:241) and the KE / |F|^2 / |P|^2 / F.P reductions.
#include <cuda_runtime.h>
#include <math.h>
#include <type_traits>
#include "jmd_common.cuh"

namespace {

#ifndef JMD_PAIR_BLOCK
#define JMD_PAIR_BLOCK 256
#endif
constexpr int PAIR_BLOCK = JMD_PAIR_BLOCK;
#ifndef JMD_PAIR_UNROLL
#define JMD_PAIR_UNROLL 4
#endif
constexpr int PAIR_UNROLL = JMD_PAIR_UNROLL;
#ifndef JMD_PAIR_BATCH
#define JMD_PAIR_BATCH 0
#endif
#ifndef JMD_PAIR_ALWAYS_WRAP
#define JMD_PAIR_ALWAYS_WRAP 0
#endif
#ifndef JMD_PAIR_MIN_BLOCKS
#define JMD_PAIR_MIN_BLOCKS 1
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
  int kind, has_cutoff, n_species, transposed;
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
};

// read-only (non-coherent) 16/32-byte position gathers
__device__ __forceinline__ float4 ld_pos(const float4* p) { return __ldg(p); }
__device__ __forceinline__ double4 ld_pos(const double4* p) {
  const double2 a = __ldg(reinterpret_cast<const double2*>(p));
  const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
  return make_double4(a.x, a.y, b.x, b.y);
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
__global__ void __launch_bounds__(PAIR_BLOCK, JMD_PAIR_MIN_BLOCKS) k_pair_force(PairP<T, DIM> Q) {
  using V4 = typename Vec4<T>::type;
  constexpr bool WANT_E = RED == 2;
  constexpr int NV = RedN<RED>::value;
  const int t = blockIdx.x * PAIR_BLOCK + threadIdx.x;
  double rv[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) rv[i] = 0.0;

  const int ai = t < Q.n ? Q.perm[t] : 0x7fffffff;
  if (ai < Q.n_rows) {                      // ghosts (ids >= n_rows) have no row
    const V4 pi = Q.pos_sorted[t];
    const int cnt = min(Q.cnt[t], Q.m_int);
    const int si = (int)pi.w;
    T f[3] = {T(0), T(0), T(0)};
    T e = T(0), ds = T(0), de = T(0);
    T vir[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
    const int* col = Q.nl + t;
    const T sig0 = Q.scalar[0], eps0 = Q.scalar[1], alp0 = Q.scalar[2];
    // free space: half = +inf, never "far"
    const T hx = Q.sp.periodic ? Q.sp.half[0] : (T)INFINITY;
    const T hy = Q.sp.periodic ? Q.sp.half[1] : (T)INFINITY;
    const T hz = Q.sp.periodic ? Q.sp.half[DIM - 1] : (T)INFINITY;
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
      if (JMD_PAIR_ALWAYS_WRAP || far) {
        d[0] = Q.sp.wrap_fast(d[0], 0);
        d[1] = Q.sp.wrap_fast(d[1], 1);
        if (DIM == 3) d[2] = Q.sp.wrap_fast(d[2], DIM - 1);
      }
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
      pair_eval<T, POT, WANT_E>(Q.has_cutoff, r2, sigma, eps, alpha, Q.r_onset2, Q.r_cutoff2, Q.inv_denom,
                                u, du_r, dus, due);
      f[0] -= du_r * d[0];
      f[1] -= du_r * d[1];
      if (DIM == 3) f[2] -= du_r * d[2];
      if (WANT_E) {
        e += u;
        ds += dus;
        de += due;
        vir[0] += du_r * d[0] * d[0];
        vir[1] += du_r * d[1] * d[1];
        vir[3] += du_r * d[0] * d[1];
        if (DIM == 3) {
          vir[2] += du_r * d[2] * d[2];
          vir[4] += du_r * d[0] * d[2];
          vir[5] += du_r * d[1] * d[2];
        }
        if (!SCALAR && Q.dparam) {
          const int sj = (int)pj.w;
          const int cell = Q.transposed ? sj * Q.n_species + si : si * Q.n_species + sj;
          if (Q.mode[0] == JMD_PARAM_SPECIES) atomicAdd(&Q.dparam[cell], 0.5 * (double)dus);
          if (Q.mode[1] == JMD_PARAM_SPECIES)
            atomicAdd(&Q.dparam[Q.n_species * Q.n_species + cell], 0.5 * (double)due);
        }
      }
    };
#if JMD_PAIR_BATCH > 0
    // Explicit batches: JMD_PAIR_BATCH row entries, then their JMD_PAIR_BATCH
    // position gathers, all in flight before the first pair is evaluated.
    int k = 0;
    for (; k + JMD_PAIR_BATCH <= cnt; k += JMD_PAIR_BATCH) {
      int jj[JMD_PAIR_BATCH];
      V4 pp4[JMD_PAIR_BATCH];
#pragma unroll
      for (int u = 0; u < JMD_PAIR_BATCH; ++u) jj[u] = __ldcs(col + (size_t)(k + u) * Q.n_pad);
#pragma unroll
      for (int u = 0; u < JMD_PAIR_BATCH; ++u) pp4[u] = ld_pos(&Q.pos_sorted[jj[u]]);
#pragma unroll
      for (int u = 0; u < JMD_PAIR_BATCH; ++u) pair(jj[u], pp4[u]);
    }
    for (; k < cnt; ++k) {
      const int j = __ldcs(col + (size_t)k * Q.n_pad);
      pair(j, ld_pos(&Q.pos_sorted[j]));
    }
#else
#pragma unroll(PAIR_UNROLL)
    for (int k = 0; k < cnt; ++k) {
      const int j = __ldcs(col + (size_t)k * Q.n_pad);   // streamed once: keep it out of L1
      pair(j, ld_pos(&Q.pos_sorted[j]));
    }
#endif
    T* fo = Q.force + (size_t)ai * DIM;
#pragma unroll
    for (int k = 0; k < DIM; ++k) fo[k] = f[k];
    if (WANT_E) {
      if (Q.e_atom) Q.e_atom[ai] = T(0.5) * e;          // smap.py:955-958: / normalization
      if (!SCALAR && Q.dparam) {
        if (Q.mode[0] == JMD_PARAM_PER_ATOM) Q.dparam[ai] = 0.5 * (double)ds;
        if (Q.mode[1] == JMD_PARAM_PER_ATOM) Q.dparam[Q.n + ai] = 0.5 * (double)de;
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
      rv[0] = 0.5 * (double)e;
#pragma unroll
      for (int k = 0; k < 6; ++k) rv[1 + k] = 0.5 * (double)vir[k];
      rv[7] = 0.5 * (double)ds;
      rv[8] = 0.5 * (double)de;
    }
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

template <typename T, int DIM, int POT, bool SCALAR>
int launch_variants(const PairP<T, DIM>& Q, bool want_e, bool kick, cudaStream_t s) {
  const int grid = (int)jmd_div_up(Q.n > 0 ? Q.n : 1, PAIR_BLOCK);
  if (want_e) {
    if (kick) k_pair_force<T, DIM, POT, SCALAR, 2, true><<<grid, PAIR_BLOCK, 0, s>>>(Q);
    else k_pair_force<T, DIM, POT, SCALAR, 2, false><<<grid, PAIR_BLOCK, 0, s>>>(Q);
  } else if (kick) {
    k_pair_force<T, DIM, POT, SCALAR, 1, true><<<grid, PAIR_BLOCK, 0, s>>>(Q);
  } else {
    k_pair_force<T, DIM, POT, SCALAR, 0, false><<<grid, PAIR_BLOCK, 0, s>>>(Q);
  }
  JMD_LAUNCH_CHECK();
  return 0;
}

template <typename T, int DIM>
int launch_pair(const jmd_nbr_t* nb, const jmd_pair_t* pp, void* force, void* e_atom, double* red,
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

extern "C" {

int64_t jmd_red_scratch_doubles(int64_t n) { return 2 + 16 * (jmd_div_up(n > 0 ? n : 1, 64) + 1); }

int jmd_pair_force(const jmd_nbr_t* nb, const jmd_pair_t* pp, void* force, void* e_atom, double* red,
                   double* dparam, double* partials, void* momentum, const void* mass, int mass_is_array,
                   double dt_2, const void* dt_dev, int want_energy, void* stream) {
  if (!nb || !pp) return JMD_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  const int dim = nb->space.dim;
  const bool we = want_energy != 0;
#define JMD_ARGS nb, pp, force, e_atom, red, dparam, partials, momentum, mass, mass_is_array, dt_2, dt_dev, we, s
  if (nb->dtype == JMD_F32 && dim == 3) return launch_pair<float, 3>(JMD_ARGS);
  if (nb->dtype == JMD_F32 && dim == 2) return launch_pair<float, 2>(JMD_ARGS);
  if (nb->dtype == JMD_F64 && dim == 3) return launch_pair<double, 3>(JMD_ARGS);
  if (nb->dtype == JMD_F64 && dim == 2) return launch_pair<double, 2>(JMD_ARGS);
#undef JMD_ARGS
  return JMD_EINVAL;
}

}  // extern "C"

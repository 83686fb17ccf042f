// Integrator kernels: velocity-Verlet halves, Nose-Hoover chain scalars, FIRE.
//
// Replaces simulate.py:168-243 (momentum_step / position_step / velocity_verlet),
// simulate.py:444-507 (NHC sub-steps; all scalar work runs in ONE single-thread
// kernel, the momentum rescale exp(-d/2 p_xi0/Q0) is applied once, fused into
// the next kick) and minimize.py:184-224 (FIRE mixing and schedule).
#include <cuda_runtime.h>
#include <math.h>
#include "jmd_common.cuh"

namespace {

constexpr int IB = 256;

template <typename T, int DIM>
struct StepP {
  int n;
  Space<T, DIM> sp;
  const T *r_in, *p_in, *f_in, *mass;
  int mass_is_array;
  T dt, dt_2;
  const T* dt_dev;
  const T* scale_dev;
  T *r_out, *p_out;
  typename Vec4<T>::type* pos_sorted;
  const int* inv_perm;
  const int* species;
  // fused skin predicate (partition.py:1146-1154) against the list's reference positions
  const T* ref;
  int* skin_blk;
  int n_rows;
  T threshold_sq;
  Space<T, DIM> nsp;       // the neighbour list's metric
  const int* n_dev;        // {n, n_rows} on the device (domain decomposition): drift the owned rows
};

template <typename T>
__device__ __forceinline__ typename Vec4<T>::type mk4(T x, T y, T z, T w);
template <> __device__ __forceinline__ float4 mk4<float>(float x, float y, float z, float w) { return make_float4(x, y, z, w); }
template <> __device__ __forceinline__ double4 mk4<double>(double x, double y, double z, double w) { return make_double4(x, y, z, w); }

// simulate.py:235-239: dt = f32(dt), dt_2 = f32(dt/2); p += dt_2 F; R = shift(R, dt p / m)
template <typename T, int DIM>
__global__ void __launch_bounds__(IB) k_kick_drift(StepP<T, DIM> S) {
  const int a = blockIdx.x * IB + threadIdx.x;
  bool moved = false;
  const int s_n = S.n_dev ? __ldg(S.n_dev + 1) : S.n;          // (no writes to the parameter struct)
  const int s_rows = S.n_dev ? s_n : S.n_rows;
  if (a < s_n) {
    T dt = S.dt, dt_2 = S.dt_2;
    if (S.dt_dev) {
      dt = (T)(float)(*S.dt_dev);
      dt_2 = (T)(float)(dt / T(2));
    }
    const T scale = S.scale_dev ? *S.scale_dev : T(1);
    const T m = S.mass_is_array ? S.mass[a] : S.mass[0];
    T r[3] = {T(0), T(0), T(0)};
    if (!S.sp.tric) {
#pragma unroll
      for (int k = 0; k < DIM; ++k) {
        const size_t o = (size_t)a * DIM + k;
        T p = S.p_in[o];
        if (S.scale_dev) p *= scale;
        p = p + dt_2 * S.f_in[o];
        S.p_out[o] = p;
        r[k] = S.sp.shift(S.r_in[o], dt * p / m, k);
        S.r_out[o] = r[k];
      }
    } else {                        // full-matrix periodic_general: the shift couples the components
      T r0[DIM], dr[DIM];
#pragma unroll
      for (int k = 0; k < DIM; ++k) {
        const size_t o = (size_t)a * DIM + k;
        T p = S.p_in[o];
        if (S.scale_dev) p *= scale;
        p = p + dt_2 * S.f_in[o];
        S.p_out[o] = p;
        r0[k] = S.r_in[o];
        dr[k] = dt * p / m;
      }
      S.sp.shift_v(r0, dr, r);
#pragma unroll
      for (int k = 0; k < DIM; ++k) S.r_out[(size_t)a * DIM + k] = r[k];
    }
    if (S.pos_sorted) {
      T w = S.species ? (T)S.species[a] : T(0);
      T q[3] = {T(0), T(0), T(0)};
      S.sp.to_real_v(r, q);
      S.pos_sorted[S.inv_perm[a]] = mk4<T>(q[0], q[1], DIM == 3 ? q[DIM - 1] : T(0), w);
    }
    if (S.skin_blk && a < s_rows) {
      // the next NeighborList.update(R') asks: did any atom move further than skin/2
      // from its reference position (strict >, min-image metric, exact arithmetic)?
      T b[DIM];
#pragma unroll
      for (int k = 0; k < DIM; ++k) b[k] = S.ref[(size_t)a * DIM + k];
      moved = dist2_exact<T, DIM>(S.nsp, r, b) > S.threshold_sq;
    }
  }
  if (S.skin_blk) {
    const int any = __syncthreads_or(moved ? 1 : 0);
    if (threadIdx.x == 0) S.skin_blk[blockIdx.x] = any;      // plain store: overwritten every drift
  }
}

template <typename T>
__global__ void __launch_bounds__(IB) k_kick_reduce(int n, int dim, T* p, const T* f, const T* mass,
                                                    int mass_is_array, T dt_2, const T* dt_dev, double* red,
                                                    double* partials) {
  const int a = blockIdx.x * IB + threadIdx.x;
  double rv[4] = {0, 0, 0, 0};
  if (a < n) {
    if (dt_dev) dt_2 = (T)(float)((T)(float)(*dt_dev) / T(2));
    const T m = mass_is_array ? mass[a] : mass[0];
    T ke = 0, ff = 0, pp = 0, fp = 0;
    for (int k = 0; k < dim; ++k) {
      const size_t o = (size_t)a * dim + k;
      T fk = f[o];
      T pk = p[o] + dt_2 * fk;
      p[o] = pk;
      ke += pk * pk / m; ff += fk * fk; pp += pk * pk; fp += fk * pk;
    }
    rv[0] = 0.5 * (double)ke; rv[1] = ff; rv[2] = pp; rv[3] = fp;
  }
  __shared__ double sm[4 * (IB / 32)];
  __shared__ int slots[4];
  if (threadIdx.x == 0) { slots[0] = JMD_RED_KINETIC; slots[1] = JMD_RED_FF; slots[2] = JMD_RED_PP; slots[3] = JMD_RED_FP; }
  __syncthreads();
  grid_reduce_finish<4, IB>(rv, partials + 2, (unsigned int*)partials, red, slots, sm);
}

template <typename T>
__global__ void k_scale(long long count, T* p, const T* scale_dev) {
  const T s = *scale_dev;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += stride) p[i] *= s;
}

__constant__ double SY3[3] = {0.828981543588751, -0.657963087177502, 0.828981543588751};
__constant__ double SY5[5] = {0.2967324292201065, 0.2967324292201065, -0.186929716880426, 0.2967324292201065,
                              0.2967324292201065};
__constant__ double SY7[7] = {0.784513610477560, 0.235573213359357, -1.17767998417887, 1.31518632068391,
                              -1.17767998417887, 0.235573213359357, 0.784513610477560};

// chain layout (T): xi[cl] | p_xi[cl] | Q[cl] | KE | scale
template <typename T>
__global__ void k_nhc_half_step(int cl, int chain_steps, int sy_steps, T dt, T tau, long long dof,
                                const T* kT_dev, const T* chain_in, T* chain, const double* ke_red, T* scale_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  // out of place: the caller's previous chain state stays intact (the reference returns a
  // fresh NoseHooverChain every step, simulate.py:492-507)
  if (chain_in != chain)
    for (int m = 0; m < 3 * cl + 1; ++m) chain[m] = chain_in[m];
  T* xi = chain;
  T* p_xi = chain + cl;
  T* Q = chain + 2 * cl;
  T* KEp = chain + 3 * cl;
  const T kT = *kT_dev;
  // update_mass (simulate.py:509-515): Q = kT tau^2 ones(f32); Q[0] *= dof
  for (int m = 0; m < cl; ++m) Q[m] = (T)((float)kT * ((float)tau * (float)tau));
  Q[0] = (T)((float)Q[0] * (float)dof);
  T KE = ke_red ? (T)(*ke_red) : *KEp;
  const T DOF = (T)dof;
  const int M = cl - 1;
  T total_scale = T(1);
  const int nsub = (chain_steps == 1 && sy_steps == 1) ? 1 : chain_steps * sy_steps;
  for (int it = 0; it < nsub; ++it) {
    T delta;
    if (nsub == 1 && chain_steps == 1 && sy_steps == 1) {
      delta = dt;
    } else {
      double w = sy_steps == 1 ? 1.0 : (sy_steps == 3 ? SY3[it % 3] : (sy_steps == 5 ? SY5[it % 5] : SY7[it % 7]));
      // simulate.py:497-500: delta = dt / chain_steps (f32); d = f32(delta * ws[i]) with
      // ws f32 unless x64 is on (then the product is formed in f64 first)
      const float delta_f = (float)dt / (float)chain_steps;
      delta = sizeof(T) == 4 ? (T)(delta_f * (float)w) : (T)(float)((double)delta_f * w);
    }
    const T d2 = delta / T(2), d4 = d2 / T(2), d8 = d4 / T(2);
    // simulate.py:456-466 backward sweep (uses the OLD p_xi[m-1])
    T G = p_xi[M - 1] * p_xi[M - 1] / Q[M - 1] - kT;
    p_xi[M] += d4 * G;
    T carry = p_xi[M];
    T old_prev;   // p_xi[m-1] is untouched until we get there, so read in place
    for (int m = M - 1; m >= 1; --m) {
      old_prev = p_xi[m - 1];
      G = old_prev * old_prev / Q[m - 1] - kT;
      T s = exp(-d8 * carry / Q[m + 1]);
      carry = s * (s * p_xi[m] + d4 * G);
      p_xi[m] = carry;
    }
    G = T(2) * KE - DOF * kT;
    T s = exp(-d8 * p_xi[1] / Q[1]);
    p_xi[0] = s * (s * p_xi[0] + d4 * G);
    s = exp(-d2 * p_xi[0] / Q[0]);
    KE = KE * s * s;
    total_scale *= s;
    for (int m = 0; m < cl; ++m) xi[m] += d2 * p_xi[m] / Q[m];
    G = T(2) * KE - DOF * kT;
    for (int m = 0; m < M; ++m) {
      T sc = exp(-d8 * p_xi[m + 1] / Q[m + 1]);
      p_xi[m] = sc * (sc * p_xi[m] + d4 * G);
      G = p_xi[m] * p_xi[m] / Q[m] - kT;
    }
    p_xi[M] += d4 * G;
  }
  *KEp = KE;
  *scale_out = total_scale;
}

// minimize.py:190-224.  fire_in/out = [dt, alpha]; n_pos in/out.
template <typename T>
__global__ void __launch_bounds__(IB) k_fire_mix(long long count, T* p, const T* f, const double* red,
                                                 const T* fire_in, T* fire_out, const int* npos_in,
                                                 int* npos_out, T dt_max, T n_min, T f_inc, T f_dec,
                                                 T alpha_start, T f_alpha) {
  const T FF = (T)red[JMD_RED_FF], PP = (T)red[JMD_RED_PP], FP = (T)red[JMD_RED_FP];
  const T Fn = sqrt(FF + T(1e-6));
  const T Pn = sqrt(PP);
  const T alpha = fire_in[1];
  const T keep = FP >= T(0) ? T(1) : T(0);
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += stride) {
    T pk = p[i];
    pk = pk + alpha * (f[i] * Pn / Fn - pk);
    p[i] = keep * pk;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    T dt = fire_in[0];
    T al = alpha;
    int np = FP >= T(0) ? *npos_in + 1 : 0;
    if (FP > T(0) && (T)np > n_min) {
      T inc = dt * f_inc;
      dt = inc < dt_max ? inc : dt_max;
      al = al * f_alpha;
    }
    if (FP < T(0)) { dt = dt * f_dec; al = alpha_start; }
    fire_out[0] = dt;
    fire_out[1] = al;
    *npos_out = np;
  }
}

template <typename T, int DIM>
int launch_kick_drift(const jmd_space_t* sp, int n, const jmd_nbr_t* nb, const void* r_in, const void* p_in,
                      const void* f_in, const void* mass, int mass_is_array, double dt, const void* dt_dev,
                      const void* scale_dev, void* r_out, void* p_out, cudaStream_t s) {
  StepP<T, DIM> S;
  S.n = n;
  S.sp.init(*sp);
  S.r_in = (const T*)r_in; S.p_in = (const T*)p_in; S.f_in = (const T*)f_in; S.mass = (const T*)mass;
  S.mass_is_array = mass_is_array;
  float dtf = (float)dt;                 // simulate.py:235-236
  S.dt = (T)dtf; S.dt_2 = (T)(float)(dtf / 2);
  S.dt_dev = (const T*)dt_dev; S.scale_dev = (const T*)scale_dev;
  S.r_out = (T*)r_out; S.p_out = (T*)p_out;
  S.pos_sorted = nb ? (typename Vec4<T>::type*)nb->pos_sorted : nullptr;
  S.inv_perm = nb ? nb->inv_perm : nullptr;
  S.species = nb ? nb->species : nullptr;
  S.ref = nullptr; S.skin_blk = nullptr; S.n_rows = 0; S.threshold_sq = T(0);
  S.nsp.init(nb ? nb->space : *sp);
  S.n_dev = nb ? nb->n_dev : nullptr;
  if (nb && nb->n_dev && nb->skin_blk && nb->reference_position) {
    S.ref = (const T*)nb->reference_position;
    S.skin_blk = nb->skin_blk;
    S.threshold_sq = (T)nb->threshold_sq;
  } else if (nb && nb->skin_blk && nb->reference_position) {
    // drift over all atoms of the list, or over its owned rows (domain decomposition:
    // ghosts are refreshed by the halo exchange and have no skin check)
    const int rows = (nb->n_rows > 0 && nb->n_rows < nb->n) ? nb->n_rows : nb->n;
    if (n == nb->n || n == rows) {
      S.ref = (const T*)nb->reference_position;
      S.skin_blk = nb->skin_blk;
      S.n_rows = rows < n ? rows : n;
      S.threshold_sq = (T)nb->threshold_sq;
    }
  }
  k_kick_drift<T, DIM><<<(int)jmd_div_up(n > 0 ? n : 1, IB), IB, 0, s>>>(S);
  JMD_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" {

int jmd_nve_kick_drift(const jmd_space_t* sp, int dtype, int n, const jmd_nbr_t* nb, const void* r_in,
                       const void* p_in, const void* f_in, const void* mass, int mass_is_array, double dt,
                       const void* dt_dev, const void* scale_dev, void* r_out, void* p_out, void* stream) {
  if (!sp || !r_in || !p_in || !f_in || !mass || !r_out || !p_out) return JMD_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
#define A sp, n, nb, r_in, p_in, f_in, mass, mass_is_array, dt, dt_dev, scale_dev, r_out, p_out, s
  if (dtype == JMD_F32 && sp->dim == 3) return launch_kick_drift<float, 3>(A);
  if (dtype == JMD_F32 && sp->dim == 2) return launch_kick_drift<float, 2>(A);
  if (dtype == JMD_F64 && sp->dim == 3) return launch_kick_drift<double, 3>(A);
  if (dtype == JMD_F64 && sp->dim == 2) return launch_kick_drift<double, 2>(A);
#undef A
  return JMD_EINVAL;
}

int jmd_kick_reduce(int dtype, int n, int dim, void* momentum, const void* force, const void* mass,
                    int mass_is_array, double dt_2, const void* dt_dev, double* red, double* partials,
                    void* stream) {
  if (!momentum || !force || !mass || !red || !partials) return JMD_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  int grid = (int)jmd_div_up(n > 0 ? n : 1, IB);
  if (dtype == JMD_F32)
    k_kick_reduce<float><<<grid, IB, 0, s>>>(n, dim, (float*)momentum, (const float*)force, (const float*)mass,
                                             mass_is_array, (float)dt_2, (const float*)dt_dev, red, partials);
  else if (dtype == JMD_F64)
    k_kick_reduce<double><<<grid, IB, 0, s>>>(n, dim, (double*)momentum, (const double*)force,
                                              (const double*)mass, mass_is_array, (double)(float)dt_2,
                                              (const double*)dt_dev, red, partials);
  else return JMD_EINVAL;
  JMD_LAUNCH_CHECK();
  return 0;
}

int jmd_scale_momentum(int dtype, int64_t count, void* momentum, const void* scale_dev, void* stream) {
  if (!momentum || !scale_dev) return JMD_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  int grid = (int)(jmd_div_up(count > 0 ? count : 1, 256) < JMD_SM_COUNT * 8 ? jmd_div_up(count > 0 ? count : 1, 256)
                                                                              : JMD_SM_COUNT * 8);
  if (dtype == JMD_F32) k_scale<float><<<grid, 256, 0, s>>>(count, (float*)momentum, (const float*)scale_dev);
  else if (dtype == JMD_F64) k_scale<double><<<grid, 256, 0, s>>>(count, (double*)momentum, (const double*)scale_dev);
  else return JMD_EINVAL;
  JMD_LAUNCH_CHECK();
  return 0;
}

int jmd_nhc_half_step(int dtype, int chain_length, int chain_steps, int sy_steps, double dt, double tau,
                      int64_t dof, const void* kT_dev, const void* chain, void* chain_out, const double* ke_red,
                      void* scale_out, void* stream) {
  if (!kT_dev || !chain || !scale_out || chain_length < 2) return JMD_EINVAL;
  if (!chain_out) chain_out = (void*)chain;
  if (sy_steps != 1 && sy_steps != 3 && sy_steps != 5 && sy_steps != 7) return JMD_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == JMD_F32)
    k_nhc_half_step<float><<<1, 32, 0, s>>>(chain_length, chain_steps, sy_steps, (float)dt, (float)tau, dof,
                                            (const float*)kT_dev, (const float*)chain, (float*)chain_out, ke_red, (float*)scale_out);
  else if (dtype == JMD_F64)
    k_nhc_half_step<double><<<1, 32, 0, s>>>(chain_length, chain_steps, sy_steps, (double)(float)dt,
                                             (double)(float)tau, dof, (const double*)kT_dev, (const double*)chain,
                                             (double*)chain_out, ke_red, (double*)scale_out);
  else return JMD_EINVAL;
  JMD_LAUNCH_CHECK();
  return 0;
}

int jmd_fire_mix(int dtype, int64_t count, void* momentum, const void* force, const double* red,
                 const void* fire_in, void* fire_out, const int32_t* npos_in, int32_t* npos_out, double dt_max,
                 double n_min, double f_inc, double f_dec, double alpha_start, double f_alpha, void* stream) {
  if (!momentum || !force || !red || !fire_in || !fire_out || !npos_in || !npos_out) return JMD_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  long long blocks = jmd_div_up(count > 0 ? count : 1, IB);
  int grid = (int)(blocks < JMD_SM_COUNT * 8 ? blocks : JMD_SM_COUNT * 8);
  if (dtype == JMD_F32)
    k_fire_mix<float><<<grid, IB, 0, s>>>(count, (float*)momentum, (const float*)force, red,
                                          (const float*)fire_in, (float*)fire_out, npos_in, npos_out,
                                          (float)dt_max, (float)n_min, (float)f_inc, (float)f_dec,
                                          (float)alpha_start, (float)f_alpha);
  else if (dtype == JMD_F64)
    k_fire_mix<double><<<grid, IB, 0, s>>>(count, (double*)momentum, (const double*)force, red,
                                           (const double*)fire_in, (double*)fire_out, npos_in, npos_out, dt_max,
                                           n_min, f_inc, f_dec, alpha_start, f_alpha);
  else return JMD_EINVAL;
  JMD_LAUNCH_CHECK();
  return 0;
}

const char* jmd_version(void) { return "jmd_b200 0.1 (sm_100a)"; }

}  // extern "C"

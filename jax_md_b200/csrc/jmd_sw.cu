// Stillinger-Weber force / energy over Dense neighbour rows, atomics-free.
//
// Replaces energy.py:842-893 (+ :994-1012) and its autodiff gradient.  The
// reference evaluates the full M x M neighbour square per atom (97 % masked for
// diamond Si).  Two kernels, one thread per atom:
//   k_sw_compact  walks the Verlet row once, applies the exact in-range test
//                 (r < cutoff; 4 of ~16-27 entries in cold Si, up to ~12 at 300 K) and
//                 writes, per in-range neighbour, a 2 x 16-byte record to compact
//                 TRANSPOSED rows: (dx, dy, dz, slot) and (r, h(r), h'(r), 1/r) with
//                 h = exp(gamma / (r/sigma - a)).  Uniform trip counts, predicated appends.
//   k_sw          works on the compact records only -- no position gathers, no min-image,
//                 no sqrt / exp and one reciprocal per triplet in the loops:
//     (1) the unordered pairs (a, b) of its own compact row: triplets centred
//         on i (energy + force on i + virial); the records are coalesced loads,
//     (2) for every compact neighbour j, j's compact row: triplets centred on j
//         that have i as an end atom (force on i only); two 16-byte gathers per entry.
// Every force component is produced by exactly one thread in a fixed order: no
// atomics, bitwise reproducible.  Needs symmetric rows (the Dense build
// guarantees that; d(i,j) = -d(j,i) exactly, so r, h and the in-range decision agree).
// History: walking full rows inside the triplet loops: 5.3 of 32 lanes active, 2.2 ms at
// N=512k; compact index rows with positions re-gathered and h re-evaluated per triplet
// (three times per triplet overall): 0.24 ms on the cold lattice but 0.74-0.97 ms at 300 K.
#include <cuda_runtime.h>
#include <math.h>
#include "jmd_common.cuh"

namespace {

constexpr int SWB = 128;

template <typename T>
struct SwP {
  int n, m_int, n_rows;
  long long n_pad;
  Space<T, 3> sp;
  const typename Vec4<T>::type* pos_sorted;
  const int* nl;
  const int* cnt;
  const int* perm;
  int* ccnt;                              // [n_pad] in-range neighbours per slot
  typename Vec4<T>::type* geo;            // [m_int, n_pad] (dx, dy, dz, slot of the neighbour)
  typename Vec4<T>::type* hd;             // [m_int, n_pad] (r, h, dh/dr, 1/r)
  T sigma, A, B, lam, gamma, eps, tbs, cutoff, a, cutoff2;
  T* force;
  double* red;
  double* partials;
  T* momentum;
  const T* mass;
  int mass_is_array;
  T dt_2;
  const T* dt_dev;
  int kick;
};

// h(r) = exp(gamma / (r/sigma - a)) and dh/dr, r < cutoff
template <typename T>
__device__ __forceinline__ void sw_h(const SwP<T>& S, T r, T& h, T& dh) {
  T x = r / S.sigma - S.a;
  h = exp(S.gamma / x);
  dh = -h * S.gamma / (S.sigma * x * x);
}

// triplet term g = h1 h2 (cos + 1/3)^2 with cos = d1.d2 / ((r1+1e-7)(r2+1e-7)),
// clipped to [-1, 1] (quantity.py:285-289).  Returns g and dg/dd1, dg/dd2.
// i1 = 1/r1, i2 = 1/r2 (from the records), m1 = 1/(r1+1e-7), m2 = 1/(r2+1e-7): the gradient
//   dg/dd1 = h1' h2 ct^2 d1/r1 + 2 h1 h2 ct [ d2/(n1 n2) - c d1/(n1 r1) ]
// is evaluated with these four reciprocals instead of ~18 IEEE divides per triplet.
template <typename T, bool WANT_G2>
__device__ __forceinline__ T sw_triplet(const T* d1, T h1, T dh1, T i1, T m1,
                                        const T* d2, T h2, T dh2, T i2, T m2, T* g1, T* g2) {
  const T dot = d1[0] * d2[0] + d1[1] * d2[1] + d1[2] * d2[2];
  const T mm = m1 * m2;
  const T c = dot * mm;
  const bool live = (c >= T(-1)) && (c <= T(1));
  const T cc = c < T(-1) ? T(-1) : (c > T(1) ? T(1) : c);
  const T ct = cc + T(1.0 / 3.0);
  const T hh = h1 * h2;
  const T ct2 = ct * ct;
  const T g = hh * ct2;
  const T pre = live ? T(2) * hh * ct : T(0);
  const T cross = pre * mm;                                // coefficient of the other vector
  const T own1 = dh1 * h2 * ct2 * i1 - pre * c * m1 * i1;  // coefficient of d1 in dg/dd1
#pragma unroll
  for (int k = 0; k < 3; ++k) g1[k] = own1 * d1[k] + cross * d2[k];
  if (WANT_G2) {
    const T own2 = dh2 * h1 * ct2 * i2 - pre * c * m2 * i2;
#pragma unroll
    for (int k = 0; k < 3; ++k) g2[k] = own2 * d2[k] + cross * d1[k];
  }
  return g;
}

// the neighbour's slot travels in the 4th component: bit pattern for float, value for double
__device__ __forceinline__ float sw_pack_idx(int j, float) { return __int_as_float(j); }
__device__ __forceinline__ double sw_pack_idx(int j, double) { return (double)j; }
__device__ __forceinline__ int sw_unpack_idx(float w) { return __float_as_int(w); }
__device__ __forceinline__ int sw_unpack_idx(double w) { return (int)w; }
__device__ __forceinline__ float4 sw_ld(const float4* p) { return __ldg(p); }
__device__ __forceinline__ double4 sw_ld(const double4* p) {
  const double2 a = __ldg(reinterpret_cast<const double2*>(p));
  const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}

template <typename T>
__global__ void __launch_bounds__(SWB) k_sw_compact(SwP<T> S) {
  using V4 = typename Vec4<T>::type;
  const int t = blockIdx.x * SWB + threadIdx.x;
  if (t >= S.n) return;
  int kc = 0;
  if (S.perm[t] < S.n_rows) {
    const V4 pi = S.pos_sorted[t];
    const int cnt = min(S.cnt[t], S.m_int);
    const int* col = S.nl + t;
    V4* og = S.geo + t;
    V4* oh = S.hd + t;
#pragma unroll 4
    for (int k = 0; k < cnt; ++k) {
      const int j = __ldcs(col + (size_t)k * S.n_pad);
      const V4 pj = S.pos_sorted[j];
      T dx, dy, dz;
      if (S.sp.tric) {                        // full-matrix box (periodic_general)
        T d[3] = {pj.x - pi.x, pj.y - pi.y, pj.z - pi.z};
        S.sp.wrap_tric(d);
        dx = d[0]; dy = d[1]; dz = d[2];
      } else {
        dx = S.sp.disp_fast(pj.x, pi.x, 0); dy = S.sp.disp_fast(pj.y, pi.y, 1);
        dz = S.sp.disp_fast(pj.z, pi.z, 2);
      }
      const T r2 = dx * dx + dy * dy + dz * dz;
      if (r2 > T(0) && r2 < S.cutoff2) {      // conservative, skips the sqrt of skin-only entries
        const T r = sqrt(r2);
        if (r < S.cutoff) {                   // energy.py:868-870
          T h, dh;
          sw_h(S, r, h, dh);
          V4 g, q;
          g.x = dx; g.y = dy; g.z = dz; g.w = sw_pack_idx(j, T(0));
          q.x = r; q.y = h; q.z = dh; q.w = T(1) / r;
          og[(size_t)kc * S.n_pad] = g;
          oh[(size_t)kc * S.n_pad] = q;
          ++kc;
        }
      }
    }
  }
  S.ccnt[t] = kc;
}

template <typename T>
__global__ void __launch_bounds__(SWB) k_sw(SwP<T> S) {
  using V4 = typename Vec4<T>::type;
  const int t = blockIdx.x * SWB + threadIdx.x;
  double rv[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  // virial dU/d(eps_ab) for the strain (I + eps) of every displacement (xx yy zz xy xz yz):
  // what quantity.pressure / stress and npt_nose_hoover obtain in the reference by
  // differentiating through `perturbation=` (quantity.py:226-282, simulate.py:848-855)
  double vir[6] = {0, 0, 0, 0, 0, 0};
  if (t < S.n && S.perm[t] < S.n_rows) {
    const int cnt = S.ccnt[t];
    const V4* gcol = S.geo + t;
    const V4* hcol = S.hd + t;
    T f[3] = {0, 0, 0};
    T e2 = 0, e3 = 0;
    const T w = S.eps * S.lam * S.tbs;
    for (int ka = 0; ka < cnt; ++ka) {
      const V4 ga = gcol[(size_t)ka * S.n_pad];
      const V4 ha = hcol[(size_t)ka * S.n_pad];
      const int j = sw_unpack_idx(ga.w);
      const T da[3] = {ga.x, ga.y, ga.z};
      const T ra = ha.x;
      const T ma = T(1) / (ra + T(1e-7));
      // two-body, energy.py:883-893: [B (r/s)^-4 - 1] exp(1/(r/s - a))
      {
        const T x = ra / S.sigma;
        const T x2 = x * x;
        const T t1 = S.B / (x2 * x2) - T(1);
        const T xa = x - S.a;
        const T t2 = exp(T(1) / xa);
        e2 += t1 * t2;
        const T df = (T(-4) * S.B / (x2 * x2 * x) * t2 - t1 * t2 / (xa * xa)) / S.sigma;
        const T c2 = S.eps * S.A * df / ra;       // d/dR_i of (1/2)(f_ij + f_ji) = -df * da/ra
        f[0] += c2 * da[0]; f[1] += c2 * da[1]; f[2] += c2 * da[2];
        const double h = 0.5 * (double)c2;        // the pair (i, j) appears in both rows
        vir[0] += h * da[0] * da[0]; vir[1] += h * da[1] * da[1]; vir[2] += h * da[2] * da[2];
        vir[3] += h * da[0] * da[1]; vir[4] += h * da[0] * da[2]; vir[5] += h * da[1] * da[2];
      }
      // (1) triplets centred on i: unordered pairs (a, b), b > a
      for (int kb = ka + 1; kb < cnt; ++kb) {
        const V4 gb = gcol[(size_t)kb * S.n_pad];
        const V4 hb = hcol[(size_t)kb * S.n_pad];
        const T db[3] = {gb.x, gb.y, gb.z};
        const T s0 = da[0] - db[0], s1 = da[1] - db[1], s2 = da[2] - db[2];
        if (!(s0 * s0 + s1 * s1 + s2 * s2 > T(1e-5) * T(1e-5))) continue;    // energy.py:872-874 (|d_ab| > 1e-5)
        T g1[3], g2[3];
        e3 += sw_triplet<T, true>(da, ha.y, ha.z, ha.w, ma, db, hb.y, hb.z, hb.w, T(1) / (hb.x + T(1e-7)), g1, g2);
        // d_a = R_j - R_i, so dE/dR_i = -(g1 + g2); force = +w (g1 + g2)
        f[0] += w * (g1[0] + g2[0]); f[1] += w * (g1[1] + g2[1]); f[2] += w * (g1[2] + g2[2]);
        // own-centre triplets carry the whole three-body energy: dU/d(eps_ab) = w (g1_a d1_b + g2_a d2_b)
        const double wd = (double)w;
        vir[0] += wd * (g1[0] * da[0] + g2[0] * db[0]);
        vir[1] += wd * (g1[1] * da[1] + g2[1] * db[1]);
        vir[2] += wd * (g1[2] * da[2] + g2[2] * db[2]);
        vir[3] += 0.5 * wd * (g1[0] * da[1] + g1[1] * da[0] + g2[0] * db[1] + g2[1] * db[0]);
        vir[4] += 0.5 * wd * (g1[0] * da[2] + g1[2] * da[0] + g2[0] * db[2] + g2[2] * db[0]);
        vir[5] += 0.5 * wd * (g1[1] * da[2] + g1[2] * da[1] + g2[1] * db[2] + g2[2] * db[1]);
      }
      // (2) triplets centred on j with i as an end atom: d1 = R_i - R_j = -da, same r and h
      {
        const T d1[3] = {-da[0], -da[1], -da[2]};
        const int cj = S.ccnt[j];
        const V4* gj = S.geo + j;
        const V4* hj = S.hd + j;
        for (int kc = 0; kc < cj; ++kc) {
          const V4 gc = sw_ld(gj + (size_t)kc * S.n_pad);
          if (sw_unpack_idx(gc.w) == t) continue;
          const V4 hc = sw_ld(hj + (size_t)kc * S.n_pad);
          const T dc[3] = {gc.x, gc.y, gc.z};
          const T s0 = d1[0] - dc[0], s1 = d1[1] - dc[1], s2 = d1[2] - dc[2];
          if (!(s0 * s0 + s1 * s1 + s2 * s2 > T(1e-5) * T(1e-5))) continue;
          T g1[3];
          sw_triplet<T, false>(d1, ha.y, ha.z, ha.w, ma, dc, hc.y, hc.z, hc.w, T(1) / (hc.x + T(1e-7)), g1, (T*)nullptr);
          f[0] -= w * g1[0]; f[1] -= w * g1[1]; f[2] -= w * g1[2];
        }
      }
    }
    const int ai = S.perm[t];
    T* fo = S.force + (size_t)ai * 3;
    fo[0] = f[0]; fo[1] = f[1]; fo[2] = f[2];
    // E = eps (A/2 sum f2 + tbs lam sum_unordered g)   (energy.py:1001-1012)
    rv[0] = (double)(S.eps * (S.A * T(0.5) * e2 + S.tbs * S.lam * e3));
#pragma unroll
    for (int k = 0; k < 6; ++k) rv[5 + k] = vir[k];
    if (S.kick) {
      T* po = S.momentum + (size_t)ai * 3;
      const T m = S.mass_is_array ? S.mass[ai] : S.mass[0];
      T ke = 0, ff = 0, pp = 0, fp = 0;
      const T dt_2 = S.dt_dev ? (T)(float)((T)(float)(*S.dt_dev) / T(2)) : S.dt_2;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        T p = po[k] + dt_2 * f[k];
        po[k] = p;
        ke += p * p / m; ff += f[k] * f[k]; pp += p * p; fp += f[k] * p;
      }
      rv[1] = 0.5 * (double)ke; rv[2] = ff; rv[3] = pp; rv[4] = fp;
    }
  }
  __shared__ double sm[11 * (SWB / 32)];
  __shared__ int slots[11];
  if (threadIdx.x == 0) {
    slots[0] = JMD_RED_ENERGY; slots[1] = JMD_RED_KINETIC; slots[2] = JMD_RED_FF; slots[3] = JMD_RED_PP;
    slots[4] = JMD_RED_FP;
    for (int k = 0; k < 6; ++k) slots[5 + k] = JMD_RED_VIRIAL + k;
  }
  __syncthreads();
  grid_reduce_finish<11, SWB>(rv, S.partials + 2, (unsigned int*)S.partials, S.red, slots, sm);
}

template <typename T>
int launch_sw(const jmd_nbr_t* nb, const jmd_sw_t* sw, int* scratch, void* force, double* red, double* partials,
              void* momentum, const void* mass, int mass_is_array, double dt_2, const void* dt_dev,
              cudaStream_t s) {
  SwP<T> S;
  S.n = nb->n; S.m_int = nb->m_int; S.n_pad = nb->n_pad;
  S.n_rows = (nb->n_rows > 0 && nb->n_rows < nb->n) ? nb->n_rows : nb->n;
  S.sp.init(nb->space);
  S.pos_sorted = (const typename Vec4<T>::type*)nb->pos_sorted;
  S.nl = nb->nl; S.cnt = nb->cnt; S.perm = nb->perm;
  S.sigma = (T)sw->sigma; S.A = (T)sw->A; S.B = (T)sw->B; S.lam = (T)sw->lam; S.gamma = (T)sw->gamma;
  S.eps = (T)sw->epsilon; S.tbs = (T)sw->three_body_strength; S.cutoff = (T)sw->cutoff;
  S.a = S.cutoff / S.sigma;
  S.cutoff2 = S.cutoff * S.cutoff * (T)(1.0 + 1e-6);   // conservative pre-filter; exact test follows
  S.force = (T*)force; S.red = red; S.partials = partials;
  S.momentum = (T*)momentum; S.mass = (const T*)mass; S.mass_is_array = mass_is_array; S.dt_2 = (T)dt_2;
  S.dt_dev = (const T*)dt_dev;
  S.kick = momentum != nullptr;
  // scratch: ccnt | geo | hd (jmd_sw_scratch_ints)
  S.ccnt = scratch;
  S.geo = (typename Vec4<T>::type*)(scratch + nb->n_pad);
  S.hd = S.geo + (size_t)nb->m_int * nb->n_pad;
  const int grid = (int)jmd_div_up(S.n > 0 ? S.n : 1, SWB);
  k_sw_compact<T><<<grid, SWB, 0, s>>>(S);
  k_sw<T><<<grid, SWB, 0, s>>>(S);
  JMD_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" int64_t jmd_sw_scratch_ints(const jmd_nbr_t* nb) {
  if (!nb) return 0;
  const int64_t rec = nb->dtype == JMD_F64 ? 16 : 8;           // two 4-vectors per entry, in int32 units
  return (int64_t)nb->n_pad + (int64_t)nb->m_int * nb->n_pad * rec;
}

extern "C" int jmd_sw_force(const jmd_nbr_t* nb, const jmd_sw_t* sw, int32_t* scratch, void* force, double* red,
                            double* partials,
                            void* momentum, const void* mass, int mass_is_array, double dt_2,
                            const void* dt_dev, void* stream) {
  if (!nb || !sw || !force || !red || !partials || !scratch) return JMD_EINVAL;
  if (nb->space.dim != 3 || nb->format != JMD_DENSE) return JMD_EINVAL;   // energy.py:1007-1010
  if (momentum && !mass) return JMD_EINVAL;
  if (nb->dtype == JMD_F32)
    return launch_sw<float>(nb, sw, scratch, force, red, partials, momentum, mass, mass_is_array, dt_2, dt_dev, (cudaStream_t)stream);
  if (nb->dtype == JMD_F64)
    return launch_sw<double>(nb, sw, scratch, force, red, partials, momentum, mass, mass_is_array, dt_2, dt_dev, (cudaStream_t)stream);
  return JMD_EINVAL;
}

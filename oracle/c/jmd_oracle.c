/* oracle/c/jmd_oracle.c -- plain-C restatement of the reference's
 * short-range MD step for the CPU baseline and for parity checks at sizes the
 * NumPy oracle cannot reach.  TEST INFRASTRUCTURE ONLY: only tests/, smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * It follows the same reference lines as the NumPy oracle (f32, 3-D, periodic cube):
 *   cell list       partition.py:146-188, 421-460 (trunc(R / cell) mod cps, x fastest)
 *   candidates      partition.py:911-951 (3^3 stencil, d2(R_i, R_j) < cutoff^2,
 *                   exact op order of space.py:213-235: mod(d + L/2, L) - L/2,
 *                   separately rounded mul/add)
 *   Dense re-test   partition.py:960-980 (reverse orientation, map_neighbor)
 *   skin predicate  partition.py:1146-1154
 *   LJ + switch     energy.py:246-272, 534-580 ; force = -grad (closed form)
 *   NVE             simulate.py:227-243
 * Rows are kept as a full (both-direction) list like the B200 path; the order
 * inside a row is stencil order then ascending id (sets are what is compared).
 * Build with -ffp-contract=off so the compiler cannot fuse the exact ops.
 * Threading: every O(N) loop is exposed as a [i0, i1) range function; the Python
 * wrapper (oracle/cport.py) runs the ranges on a thread pool (ctypes drops the
 * GIL), because this image ships gcc without libgomp.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int n, cps, m;          /* atoms, cells per side, row capacity */
  float L, half, cell, cutoff_sq, thr_sq;
  int dense;              /* 1: keep a pair only if both orientations pass */
  int* cell_start;        /* [cps^3 + 1] */
  int* cell_atoms;        /* [n] ids sorted by cell, ascending id inside a cell */
  int* hash;              /* [n] */
  int* nl;                /* [n, m] row-major, padded with n */
  int* cnt;               /* [n] */
  float* ref;             /* [n, 3] */
  int max_row, max_cell, overflow, builds;
} jo_nbr_t;

static inline float jo_mod(float t, float L) {          /* jnp.mod, L > 0 */
  float m = fmodf(t, L);
  if (m != 0.0f && m < 0.0f) m = m + L;
  return m;
}
static inline float jo_disp(float a, float b, float L, float h) {
  float d = a - b;
  float t = d + h;
  return jo_mod(t, L) - h;
}
static inline float jo_d2(const float* a, const float* b, float L, float h) {
  float x = jo_disp(a[0], b[0], L, h), y = jo_disp(a[1], b[1], L, h), z = jo_disp(a[2], b[2], L, h);
  float s = x * x;
  s = s + y * y;
  s = s + z * z;
  return s;
}

jo_nbr_t* jo_nbr_create(int n, float L, float r_cutoff, float skin, int m, int dense) {
  jo_nbr_t* nb = (jo_nbr_t*)calloc(1, sizeof(jo_nbr_t));
  float cutoff = r_cutoff + skin;
  nb->n = n; nb->L = L; nb->half = L * 0.5f;
  nb->cps = (int)floorf(L / cutoff);
  if (nb->cps < 3) { free(nb); return NULL; }
  nb->cell = L / (float)nb->cps;
  nb->cutoff_sq = cutoff * cutoff;
  nb->thr_sq = (skin / 2.0f) * (skin / 2.0f);
  nb->m = m; nb->dense = dense;
  int nc = nb->cps * nb->cps * nb->cps;
  nb->cell_start = (int*)malloc(sizeof(int) * (nc + 1));
  nb->cell_atoms = (int*)malloc(sizeof(int) * n);
  nb->hash = (int*)malloc(sizeof(int) * n);
  nb->nl = (int*)malloc(sizeof(int) * (size_t)n * m);
  nb->cnt = (int*)malloc(sizeof(int) * n);
  nb->ref = (float*)malloc(sizeof(float) * 3 * (size_t)n);
  return nb;
}

void jo_nbr_free(jo_nbr_t* nb) {
  if (!nb) return;
  free(nb->cell_start); free(nb->cell_atoms); free(nb->hash); free(nb->nl); free(nb->cnt); free(nb->ref);
  free(nb);
}

int* jo_nbr_rows(jo_nbr_t* nb) { return nb->nl; }
int* jo_nbr_counts(jo_nbr_t* nb) { return nb->cnt; }
int jo_nbr_max_row(jo_nbr_t* nb) { return nb->max_row; }
int jo_nbr_max_cell(jo_nbr_t* nb) { return nb->max_cell; }
int jo_nbr_overflow(jo_nbr_t* nb) { return nb->overflow; }
int jo_nbr_builds(jo_nbr_t* nb) { return nb->builds; }

/* serial part of a rebuild: hash + counting sort (ascending id inside a cell) */
void jo_nbr_bin(jo_nbr_t* nb, const float* R) {
  const int n = nb->n, cps = nb->cps, nc = cps * cps * cps;
  memset(nb->cell_start, 0, sizeof(int) * (nc + 1));
  for (int i = 0; i < n; ++i) {
    int h = 0, mult = 1;
    for (int k = 0; k < 3; ++k) {
      int ci = (int)(R[3 * i + k] / nb->cell);
      ci %= cps; if (ci < 0) ci += cps;
      h += ci * mult; mult *= cps;
    }
    nb->hash[i] = h;
    nb->cell_start[h + 1]++;
  }
  int mc = 0;
  for (int c = 0; c < nc; ++c) { if (nb->cell_start[c + 1] > mc) mc = nb->cell_start[c + 1]; nb->cell_start[c + 1] += nb->cell_start[c]; }
  nb->max_cell = mc;
  int* cur = (int*)malloc(sizeof(int) * nc);
  memcpy(cur, nb->cell_start, sizeof(int) * nc);
  for (int i = 0; i < n; ++i) nb->cell_atoms[cur[nb->hash[i]]++] = i;
  free(cur);
  memcpy(nb->ref, R, sizeof(float) * 3 * (size_t)n);
  nb->builds++;
  nb->max_row = 0;
}

/* rows of atoms [i0, i1); returns the longest row of the range */
int jo_nbr_rows_range(jo_nbr_t* nb, const float* R, int i0, int i1) {
  const int n = nb->n, cps = nb->cps;
  int max_row = 0;
  for (int i = i0; i < i1; ++i) {
    const float* ri = R + 3 * i;
    int h = nb->hash[i];
    int cx = h % cps, cy = (h / cps) % cps, cz = h / (cps * cps);
    int k = 0;
    int* row = nb->nl + (size_t)i * nb->m;
    for (int sx = -1; sx <= 1; ++sx) for (int sy = -1; sy <= 1; ++sy) for (int sz = -1; sz <= 1; ++sz) {
      int x = (cx + sx + cps) % cps, y = (cy + sy + cps) % cps, z = (cz + sz + cps) % cps;
      int c = x + cps * (y + cps * z);
      for (int p = nb->cell_start[c]; p < nb->cell_start[c + 1]; ++p) {
        int j = nb->cell_atoms[p];
        if (j == i) continue;
        const float* rj = R + 3 * j;
        int keep = jo_d2(ri, rj, nb->L, nb->half) < nb->cutoff_sq;
        if (keep && nb->dense) keep = jo_d2(rj, ri, nb->L, nb->half) < nb->cutoff_sq;
        if (keep) { if (k < nb->m) row[k] = j; ++k; }
      }
    }
    nb->cnt[i] = k;
    for (int q = k; q < nb->m; ++q) row[q] = n;
    if (k > max_row) max_row = k;
  }
  return max_row;
}

void jo_nbr_set_max_row(jo_nbr_t* nb, int max_row) {
  nb->max_row = max_row;
  if (max_row > nb->m) nb->overflow = 1;
}

/* skin predicate over [i0, i1): any |d(R_i, ref_i)|^2 > (skin/2)^2 */
int jo_skin_range(const jo_nbr_t* nb, const float* R, int i0, int i1) {
  for (int i = i0; i < i1; ++i)
    if (jo_d2(R + 3 * i, nb->ref + 3 * i, nb->L, nb->half) > nb->thr_sq) return 1;
  return 0;
}

/* LJ with the multiplicative switch; full list: F_i = -sum_j U'(r) dR/r.
 * Returns the potential energy of rows [i0, i1) (sum over rows / 2). */
double jo_lj_force_range(const jo_nbr_t* nb, const float* R, float sigma, float eps, float r_onset,
                         float r_cutoff, float* F, int i0, int i1) {
  const float ro2 = r_onset * r_onset, rc2 = r_cutoff * r_cutoff;
  const float den = (rc2 - ro2) * (rc2 - ro2) * (rc2 - ro2);
  double etot = 0.0;
  for (int i = i0; i < i1; ++i) {
    const float* ri = R + 3 * i;
    const int* row = nb->nl + (size_t)i * nb->m;
    int c = nb->cnt[i] < nb->m ? nb->cnt[i] : nb->m;
    float fx = 0, fy = 0, fz = 0, e = 0;
    for (int k = 0; k < c; ++k) {
      const float* rj = R + 3 * row[k];
      float dx = jo_disp(ri[0], rj[0], nb->L, nb->half);
      float dy = jo_disp(ri[1], rj[1], nb->L, nb->half);
      float dz = jo_disp(ri[2], rj[2], nb->L, nb->half);
      float r2 = dx * dx + dy * dy + dz * dz;
      if (!(r2 > 0.0f) || !(r2 < rc2)) continue;
      float ir2 = 1.0f / r2, x2 = sigma * sigma * ir2, x6 = x2 * x2 * x2, x12 = x6 * x6;
      float u = 4.0f * eps * (x12 - x6);
      float du_r = -24.0f * eps * (2.0f * x12 - x6) * ir2;
      if (r2 >= ro2) {
        float a = rc2 - r2;
        float S = a * a * (rc2 + 2.0f * r2 - 3.0f * ro2) / den;
        float dS_r = 12.0f * a * (ro2 - r2) / den;
        du_r = dS_r * u + S * du_r;
        u = S * u;
      }
      fx -= du_r * dx; fy -= du_r * dy; fz -= du_r * dz;
      e += u;
    }
    F[3 * i] = fx; F[3 * i + 1] = fy; F[3 * i + 2] = fz;
    etot += 0.5 * (double)e;
  }
  return etot;
}

/* simulate.py:238-239 over atoms [i0, i1) */
void jo_kick_drift_range(const jo_nbr_t* nb, float* R, float* P, const float* F, float mass, float dt,
                         int i0, int i1) {
  const float dt_2 = dt / 2.0f;
  for (int i = 3 * i0; i < 3 * i1; ++i) {
    P[i] = P[i] + dt_2 * F[i];
    R[i] = jo_mod(R[i] + dt * P[i] / mass, nb->L);
  }
}

void jo_kick_range(float* P, const float* F, float dt, int i0, int i1) {
  const float dt_2 = dt / 2.0f;
  for (int i = 3 * i0; i < 3 * i1; ++i) P[i] = P[i] + dt_2 * F[i];
}

// Neighbour-list build / update for sm_100a.
//
// Replaces the reference pipeline partition.py:349-471 (cell list by argsort +
// scatter) and partition.py:911-1154 (candidate gather, distance mask, cumsum
// compaction, skin predicate under lax.cond) with:
//   k_skin        max-displacement predicate, last block latches the decision
//                 and (tail_launch mode) launches the rebuild from the device,
//                 so the steady-state step pays no launches for the lax.cond.
//   k_zero/k_hash/k_scan*/k_scatter/k_rank_sort
//                 counting sort of cell hashes -> cell-ordered float4 positions
//   k_build_*     stencil scan, one warp per home cell, candidates staged in
//                 shared memory once per cell, warp-ballot compaction into
//                 capacity-bounded transposed rows (order == reference order)
//   k_export / k_finalize
//                 public idx (Dense / Sparse / OrderedSparse), error bits.
#include <cuda_runtime.h>
#include <stdio.h>
#include "jmd_common.cuh"

namespace {

enum { ST_REBUILD = JMD_ST_REBUILD, ST_MAX_CELL = JMD_ST_MAX_CELL_OCC,
       ST_MAX_ROW = JMD_ST_MAX_ROW, ST_TOTAL = JMD_ST_TOTAL,
       ST_BUILDS = JMD_ST_BUILDS, ST_TICKET = JMD_ST_SCAN_TICKET,
       ST_PENDING = 6 };

#ifndef JMD_BUILD_WARP_PER_CELL
#define JMD_BUILD_WARP_PER_CELL 0
#endif
constexpr int SCAN_TILE = 2048;   // 256 threads x 8
constexpr int BUILD_WARPS = 4;
constexpr int BUILD_CAP = 1024;   // candidates staged per warp per chunk
constexpr int RANK_LIMIT = 4096;  // cells above this keep arrival order

template <typename T, int DIM>
struct NbrP {
  int n, format, use_cells, mask_self, always_rebuild, n_cells, cell_capacity, m_int;
  int cps[3];
  int count_only, tail_launch, two_sided, rev_only;
  long long n_pad, max_occupancy;
  T cell_size[DIM];
  T cutoff_sq, threshold_sq, band, far;
  Space<T, DIM> sp;
  int *cell_count, *cell_start, *cell_cursor, *scan_tmp, *hash, *tmp_ids, *perm, *inv_perm;
  typename Vec4<T>::type* pos_sorted;
  int* nl;
  int* cnt;
  int* cnt_lower;
  long long* offsets;
  T* ref;
  int* idx;
  uint8_t* error;
  long long* state;
  const int* species;
  const T* position;
};

#define GATE(P, gated) if ((gated) && (P).state[ST_REBUILD] == 0) return

template <typename T>
__device__ __forceinline__ typename Vec4<T>::type make_v4(T x, T y, T z, T w);
template <>
__device__ __forceinline__ float4 make_v4<float>(float x, float y, float z, float w) {
  return make_float4(x, y, z, w);
}
template <>
__device__ __forceinline__ double4 make_v4<double>(double x, double y, double z, double w) {
  return make_double4(x, y, z, w);
}

// species id carried in .w (exact for ids < 2^24)
template <typename T, int DIM>
__device__ __forceinline__ typename Vec4<T>::type load_atom(const NbrP<T, DIM>& P, int i) {
  const T* r = P.position + (size_t)i * DIM;
  T z = DIM == 3 ? r[DIM - 1] : T(0);
  T w = P.species ? (T)P.species[i] : T(0);
  return make_v4<T>(r[0], r[1], z, w);
}

// ---- binning -------------------------------------------------------------------

template <typename T, int DIM>
__global__ void k_zero(NbrP<T, DIM> P, int gated) {
  GATE(P, gated);
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int stride = gridDim.x * blockDim.x;
  for (int c = i; c <= P.n_cells; c += stride) {
    P.cell_count[c] = 0;
    if (c < P.n_cells) P.cell_cursor[c] = 0;
  }
  if (i == 0) P.state[ST_MAX_CELL] = 0;
}

template <typename T, int DIM>
__global__ void k_build_reset(NbrP<T, DIM> P, int gated) {
  GATE(P, gated);
  P.state[ST_MAX_ROW] = 0;
  P.state[ST_TOTAL] = 0;
}

// partition.py:421-423: int32(R / cell_size) (truncation), mod cells_per_side,
// hash = x + y*cx + z*cx*cy.
template <typename T, int DIM>
__global__ void k_hash(NbrP<T, DIM> P, int gated) {
  GATE(P, gated);
  int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
    const T* r = P.position + (size_t)i * DIM;
    int h = 0, mult = 1;
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
      int ci = (int)div_rn(r[k], P.cell_size[k]);
      ci %= P.cps[k];
      if (ci < 0) ci += P.cps[k];
      h += ci * mult;
      mult *= P.cps[k];
    }
    P.hash[i] = h;
    atomicAdd(&P.cell_count[h], 1);
  }
}

// three-pass exclusive scan: tile sums, scan of tile sums, apply.
template <typename TIn, typename TOut>
__device__ __forceinline__ void scan_tile_load(const TIn* in, long long n, long long base,
                                               TOut (&v)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    long long i = base + (long long)threadIdx.x * 8 + j;
    v[j] = i < n ? (TOut)in[i] : TOut(0);
  }
}

template <typename TOut>
__device__ __forceinline__ TOut block_excl_scan_256(TOut x, TOut* total, TOut* smem /*[8]*/) {
  // exclusive scan of one value per thread across 256 threads
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  TOut incl = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    TOut y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) smem[warp] = incl;
  __syncthreads();
  TOut wbase = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    TOut s = smem[w];
    if (w < warp) wbase += s;
    tot += s;
  }
  __syncthreads();
  *total = tot;
  return wbase + incl - x;
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) k_scan_tiles(const TIn* in, long long n, TOut* tile_sums,
                                                    long long* max_out, const long long* gate) {
  if (gate && *gate == 0) return;
  __shared__ TOut sm[8];
  TOut v[8];
  scan_tile_load<TIn, TOut>(in, n, (long long)blockIdx.x * SCAN_TILE, v);
  TOut s = 0, mx = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { s += v[j]; mx = v[j] > mx ? v[j] : mx; }
  TOut tot;
  block_excl_scan_256<TOut>(s, &tot, sm);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
  if (max_out) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      TOut y = __shfl_xor_sync(0xffffffffu, mx, o);
      mx = y > mx ? y : mx;
    }
    if ((threadIdx.x & 31) == 0 && mx > 0) atomicMax((unsigned long long*)max_out, (unsigned long long)mx);
  }
}

template <typename TOut>
__global__ void __launch_bounds__(256) k_scan_top(TOut* tile_sums, int n_tiles, const long long* gate) {
  if (gate && *gate == 0) return;
  __shared__ TOut sm[8];
  TOut carry = 0;
  for (int base = 0; base < n_tiles; base += 256) {
    int i = base + threadIdx.x;
    TOut x = i < n_tiles ? tile_sums[i] : TOut(0);
    TOut tot;
    TOut e = block_excl_scan_256<TOut>(x, &tot, sm);
    if (i < n_tiles) tile_sums[i] = carry + e;
    carry += tot;
  }
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) k_scan_apply(const TIn* in, long long n, const TOut* tile_sums,
                                                    TOut* out /*[n+1]*/, const long long* gate) {
  if (gate && *gate == 0) return;
  __shared__ TOut sm[8];
  TOut v[8];
  long long base = (long long)blockIdx.x * SCAN_TILE;
  scan_tile_load<TIn, TOut>(in, n, base, v);
  TOut s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += v[j];
  TOut tot;
  TOut e = block_excl_scan_256<TOut>(s, &tot, sm) + tile_sums[blockIdx.x];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    long long i = base + (long long)threadIdx.x * 8 + j;
    if (i < n) out[i] = e;
    e += v[j];
    if (i == n - 1) out[n] = e;
  }
}

template <typename T, int DIM>
__global__ void k_scatter(NbrP<T, DIM> P, int gated) {
  GATE(P, gated);
  int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
    int h = P.hash[i];
    int pos = P.cell_start[h] + atomicAdd(&P.cell_cursor[h], 1);
    P.tmp_ids[pos] = i;
    P.inv_perm[i] = pos;
  }
}

// Stable order inside a cell (== stable argsort of hashes, partition.py:432):
// rank = #atoms of the same cell with a smaller id.
template <typename T, int DIM>
__global__ void k_rank_sort(NbrP<T, DIM> P, int gated) {
  GATE(P, gated);
  int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
    int h = P.hash[i];
    int s = P.cell_start[h];
    int c = P.cell_start[h + 1] - s;
    int rank;
    if (c <= RANK_LIMIT) {
      rank = 0;
      for (int j = 0; j < c; ++j) rank += (P.tmp_ids[s + j] < i);
    } else {
      rank = P.inv_perm[i] - s;
    }
    int dst = s + rank;
    P.perm[dst] = i;
    P.pos_sorted[dst] = load_atom(P, i);
  }
}

// inv_perm is read by k_rank_sort (arrival position) so it is rewritten after.
template <typename T, int DIM>
__global__ void k_inv_perm(NbrP<T, DIM> P, int gated) {
  GATE(P, gated);
  int stride = gridDim.x * blockDim.x;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < P.n; t += stride) P.inv_perm[P.perm[t]] = t;
}

// all-pairs path: identity order.
template <typename T, int DIM>
__global__ void k_identity_sort(NbrP<T, DIM> P, int gated) {
  GATE(P, gated);
  int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
    P.perm[i] = i;
    P.inv_perm[i] = i;
    P.pos_sorted[i] = load_atom(P, i);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) P.state[ST_MAX_CELL] = 0;
}

template <typename T, int DIM>
__global__ void k_pack(NbrP<T, DIM> P) {
  int stride = gridDim.x * blockDim.x;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < P.n; t += stride)
    P.pos_sorted[t] = load_atom(P, P.perm[t]);
}

// ---- candidate test -------------------------------------------------------------
// Reference semantics (oracle/partition.py): the cell path tests
// d2(R_i, R_c) < cutoff^2 (partition.py:945-951); Dense then re-tests with the
// opposite orientation d2(R_c, R_i) (prune_neighbor_list_dense via map_neighbor,
// partition.py:960-980, space.py:494-502).  The two differ only by rounding, so
// the second one is evaluated only inside a rounding band around the cutoff.
template <typename T, int DIM>
__device__ __forceinline__ bool candidate_test(const NbrP<T, DIM>& P, const T* hp, const T* cp) {
  bool keep;
  if (P.rev_only) {
    keep = dist2_exact<T, DIM>(P.sp, cp, hp) < P.cutoff_sq;
  } else {
    T d1 = dist2_exact<T, DIM>(P.sp, hp, cp);
    keep = d1 < P.cutoff_sq;
    if (P.two_sided) {
      bool near = fabs(d1 - P.cutoff_sq) <= P.band;
#pragma unroll
      for (int k = 0; k < DIM; ++k) near = near || !(fabs(hp[k] - cp[k]) <= P.far);
      if (near) keep = keep && (dist2_exact<T, DIM>(P.sp, cp, hp) < P.cutoff_sq);
    }
  }
  return keep;
}

template <typename T, int DIM>
__device__ __forceinline__ void append_rows(const NbrP<T, DIM>& P, int slot, int hid, bool keep,
                                            int cslot, int cid, int& k, int& kl) {
  const unsigned lane = threadIdx.x & 31;
  unsigned b = __ballot_sync(0xffffffffu, keep);
  if (keep && !P.count_only) {
    int pos = k + __popc(b & ((1u << lane) - 1u));
    if (pos < P.m_int) P.nl[(size_t)pos * P.n_pad + slot] = cslot;
  }
  k += __popc(b);
  kl += __popc(__ballot_sync(0xffffffffu, keep && cid < hid));
}

template <typename T, int DIM>
__global__ void __launch_bounds__(BUILD_WARPS * 32) k_build_cells(NbrP<T, DIM> P, int gated) {
  GATE(P, gated);
  using V4 = typename Vec4<T>::type;
  constexpr int NS = DIM == 3 ? 27 : 9;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  V4* cand_pos = reinterpret_cast<V4*>(smem_raw) + (size_t)wib * BUILD_CAP;
  int* ibase = reinterpret_cast<int*>(smem_raw + sizeof(V4) * BUILD_CAP * BUILD_WARPS);
  int* cand_slot = ibase + (size_t)wib * (2 * BUILD_CAP + 4 * 32);
  int* cand_id = cand_slot + BUILD_CAP;
  int* st_start = cand_id + BUILD_CAP;
  int* st_count = st_start + 32;
  int* st_rot = st_count + 32;
  int* st_off = st_rot + 32;

  const int nwarps = gridDim.x * BUILD_WARPS;
  long long wmax = 0, wtotal = 0;
  const int cx_n = P.cps[0], cy_n = P.cps[1];
  for (int c = blockIdx.x * BUILD_WARPS + wib; c < P.n_cells; c += nwarps) {
    const int hs = P.cell_start[c];
    const int hn = P.cell_start[c + 1] - hs;
    if (hn == 0) continue;
    int cc[3];
    cc[0] = c % cx_n;
    cc[1] = (c / cx_n) % cy_n;
    cc[2] = DIM == 3 ? c / (cx_n * cy_n) : 0;
    // stencil, reference order: first coordinate slowest (partition.py:232-240)
    int my_start = 0, my_cnt = 0, my_rot = 0;
    if (lane < NS) {
      int sh[3];
      if (DIM == 3) { sh[0] = lane / 9 - 1; sh[1] = (lane / 3) % 3 - 1; sh[2] = lane % 3 - 1; }
      else { sh[0] = lane / 3 - 1; sh[1] = lane % 3 - 1; sh[2] = 0; }
      int h = 0, mult = 1;
#pragma unroll
      for (int k = 0; k < DIM; ++k) {
        int v = cc[k] + sh[k];
        v = v < 0 ? v + P.cps[k] : (v >= P.cps[k] ? v - P.cps[k] : v);
        h += v * mult;
        mult *= P.cps[k];
      }
      my_start = P.cell_start[h];
      my_cnt = P.cell_start[h + 1] - my_start;
      // slot = sorted_rank mod capacity (partition.py:441): slot order inside a
      // cell is arrival order rotated by `rot`.
      int cap = P.cell_capacity > 0 ? P.cell_capacity : 1;
      int room = cap - my_start % cap;
      my_rot = my_cnt < room ? my_cnt : room;
    }
    int incl = my_cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    __syncwarp();
    st_start[lane] = my_start;
    st_count[lane] = my_cnt;
    st_rot[lane] = my_rot;
    st_off[lane] = incl - my_cnt;
    __syncwarp();

    for (int chunk = 0; chunk < total; chunk += BUILD_CAP) {
      const int nchunk = min(BUILD_CAP, total - chunk);
      for (int q = lane; q < nchunk; q += 32) {
        int ci = chunk + q;
        int s = 0;
        while (s + 1 < NS && st_off[s + 1] <= ci) ++s;
        int within = ci - st_off[s];
        int cntc = st_count[s];
        int r = within + st_rot[s];
        r = r >= cntc ? r - cntc : r;
        int rank = st_start[s] + r;
        cand_pos[q] = P.pos_sorted[rank];
        cand_slot[q] = rank;
        cand_id[q] = P.perm[rank];
      }
      __syncwarp();
      const bool last_chunk = chunk + BUILD_CAP >= total;
      for (int h = 0; h < hn; ++h) {
        const int slot = hs + h;
        const V4 hv = P.pos_sorted[slot];
        const T hp[3] = {hv.x, hv.y, hv.z};
        const int hid = P.perm[slot];
        int k = 0, kl = 0;
        if (chunk != 0) { k = P.cnt[slot]; kl = P.cnt_lower[slot]; }
        for (int q0 = 0; q0 < nchunk; q0 += 32) {
          const int q = q0 + lane;
          bool keep = false;
          int cslot = 0, cid = 0;
          if (q < nchunk) {
            const V4 cv = cand_pos[q];
            const T cp[3] = {cv.x, cv.y, cv.z};
            cslot = cand_slot[q];
            cid = cand_id[q];
            keep = candidate_test<T, DIM>(P, hp, cp);
            if (P.mask_self && cslot == slot) keep = false;
          }
          append_rows<T, DIM>(P, slot, hid, keep, cslot, cid, k, kl);
        }
        if (lane == 0) { P.cnt[slot] = k; P.cnt_lower[slot] = kl; }
        if (last_chunk) {
          wmax = k > wmax ? k : wmax;
          wtotal += (P.format == JMD_ORDERED_SPARSE) ? kl : k;
        }
      }
      __syncwarp();
    }
  }
  if (lane == 0) {
    if (wmax > 0) atomicMax((unsigned long long*)&P.state[ST_MAX_ROW], (unsigned long long)wmax);
    if (wtotal > 0) atomicAdd((unsigned long long*)&P.state[ST_TOTAL], (unsigned long long)wtotal);
  }
}

// Thread-per-atom stencil scan (the default).  Thread t owns sorted slot t and
// walks the 3^d stencil cells of its own cell in reference order; candidates
// of a cell are a contiguous range of the cell-sorted float4 array, so the 32
// lanes of a warp (2-3 adjacent cells) issue loads that hit 2-3 distinct
// addresses (L1 broadcast).  Every row is appended in candidate order by its
// own thread: no ballots, no atomics, order == reference order.  Row k of the
// transposed list is written by neighbouring lanes at neighbouring addresses.
constexpr int TPA_BLOCK = 128;

template <typename T, int DIM, bool ORDERED>
__global__ void __launch_bounds__(TPA_BLOCK) k_build_cells_tpa(NbrP<T, DIM> P, int gated) {
  GATE(P, gated);
  using V4 = typename Vec4<T>::type;
  constexpr int NS = DIM == 3 ? 27 : 9;
  const int slot = blockIdx.x * TPA_BLOCK + threadIdx.x;
  long long my_k = 0, my_tot = 0;
  if (slot < P.n) {
    const V4 hv = P.pos_sorted[slot];
    const T hp[3] = {hv.x, hv.y, hv.z};
    const int hid = P.perm[slot];
    // own cell from the stored hash of this atom
    const int c = P.hash[hid];
    const int cx_n = P.cps[0], cy_n = P.cps[1];
    int cc[3];
    cc[0] = c % cx_n;
    cc[1] = (c / cx_n) % cy_n;
    cc[2] = DIM == 3 ? c / (cx_n * cy_n) : 0;
    const int cap = P.cell_capacity > 0 ? P.cell_capacity : 1;
    int k = 0, kl = 0;
    int* out = P.nl + slot;
    for (int s = 0; s < NS; ++s) {
      int sh[3];
      if (DIM == 3) { sh[0] = s / 9 - 1; sh[1] = (s / 3) % 3 - 1; sh[2] = s % 3 - 1; }
      else { sh[0] = s / 3 - 1; sh[1] = s % 3 - 1; sh[2] = 0; }
      int h = 0, mult = 1;
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        int v = cc[d] + sh[d];
        v = v < 0 ? v + P.cps[d] : (v >= P.cps[d] ? v - P.cps[d] : v);
        h += v * mult;
        mult *= P.cps[d];
      }
      const int start = __ldg(&P.cell_start[h]);
      const int count = __ldg(&P.cell_start[h + 1]) - start;
      // slot = sorted_rank mod capacity (partition.py:441): rotated arrival order
      const int room = cap - start % cap;
      int r = count < room ? count : room;       // first rank offset in slot order
      if (r == count) r = 0;
      for (int q = 0; q < count; ++q) {
        const int rank = start + r;
        r = r + 1 == count ? 0 : r + 1;
        const V4 cv = P.pos_sorted[rank];
        const T cp[3] = {cv.x, cv.y, cv.z};
        bool keep = candidate_test<T, DIM>(P, hp, cp);
        if (P.mask_self && rank == slot) keep = false;
        if (keep) {
          if (!P.count_only && k < P.m_int) out[(size_t)k * P.n_pad] = rank;
          ++k;
          if (ORDERED) kl += (__ldg(&P.perm[rank]) < hid);
        }
      }
    }
    P.cnt[slot] = k;
    P.cnt_lower[slot] = kl;
    my_k = k;
    my_tot = ORDERED ? kl : k;
  }
  // block max / total -> one atomic each per warp
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    long long y = __shfl_xor_sync(0xffffffffu, my_k, o);
    my_k = y > my_k ? y : my_k;
    my_tot += __shfl_xor_sync(0xffffffffu, my_tot, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (my_k > 0) atomicMax((unsigned long long*)&P.state[ST_MAX_ROW], (unsigned long long)my_k);
    if (my_tot > 0) atomicAdd((unsigned long long*)&P.state[ST_TOTAL], (unsigned long long)my_tot);
  }
}

// all-pairs candidates (partition.py:904-909): one warp per atom, candidates in
// id order.
template <typename T, int DIM>
__global__ void __launch_bounds__(128) k_build_all_pairs(NbrP<T, DIM> P, int gated) {
  GATE(P, gated);
  using V4 = typename Vec4<T>::type;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  long long wmax = 0, wtotal = 0;
  for (int i = warp; i < P.n; i += nwarps) {
    const V4 hv = P.pos_sorted[i];
    const T hp[3] = {hv.x, hv.y, hv.z};
    int k = 0, kl = 0;
    for (int j0 = 0; j0 < P.n; j0 += 32) {
      const int j = j0 + lane;
      bool keep = false;
      if (j < P.n) {
        const V4 cv = P.pos_sorted[j];
        const T cp[3] = {cv.x, cv.y, cv.z};
        keep = candidate_test<T, DIM>(P, hp, cp);
        if (P.mask_self && j == i) keep = false;
      }
      append_rows<T, DIM>(P, i, i, keep, j, j, k, kl);
    }
    if (lane == 0) { P.cnt[i] = k; P.cnt_lower[i] = kl; }
    wmax = k > wmax ? k : wmax;
    wtotal += (P.format == JMD_ORDERED_SPARSE) ? kl : k;
  }
  if (lane == 0) {
    if (wmax > 0) atomicMax((unsigned long long*)&P.state[ST_MAX_ROW], (unsigned long long)wmax);
    if (wtotal > 0) atomicAdd((unsigned long long*)&P.state[ST_TOTAL], (unsigned long long)wtotal);
  }
}

// ---- export to the public formats ---------------------------------------------------

// per-atom (user order) number of public sparse entries
template <typename T, int DIM>
__global__ void k_sparse_counts(NbrP<T, DIM> P, int gated) {
  GATE(P, gated);
  int stride = gridDim.x * blockDim.x;
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < P.n; a += stride) {
    int t = P.inv_perm[a];
    // rows longer than m_int are truncated; lower-count is then recomputed on export
    int c = P.format == JMD_ORDERED_SPARSE ? P.cnt_lower[t] : min(P.cnt[t], P.m_int);
    P.tmp_ids[a] = c;
  }
}

// Dense: idx[a, k] (partition.py:960-980, 1105); Sparse: idx[0]=receivers,
// idx[1]=senders ordered by sender then candidate order (partition.py:1010-1032).
template <typename T, int DIM>
__global__ void __launch_bounds__(128) k_export(NbrP<T, DIM> P, int gated) {
  GATE(P, gated);
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const long long cap = P.max_occupancy;
  for (int t = warp; t < P.n; t += nwarps) {
    const int a = P.perm[t];
    const int c = min(P.cnt[t], P.m_int);
    if (P.format == JMD_DENSE) {
      int* row = P.idx + (size_t)a * cap;
      for (int k = lane; k < cap; k += 32) {
        int v = P.n;
        if (k < c) v = P.perm[P.nl[(size_t)k * P.n_pad + t]];
        row[k] = v;
      }
    } else {
      long long off = P.offsets[a];
      int* recv = P.idx;
      int* send = P.idx + cap;
      int kk = 0;
      for (int k0 = 0; k0 < c; k0 += 32) {
        const int k = k0 + lane;
        int v = P.n;
        bool keep = false;
        if (k < c) {
          v = P.perm[P.nl[(size_t)k * P.n_pad + t]];
          keep = P.format == JMD_SPARSE || v < a;
        }
        unsigned b = __ballot_sync(0xffffffffu, keep);
        if (keep) {
          long long pos = off + kk + __popc(b & ((1u << lane) - 1u));
          if (pos < cap) { recv[pos] = v; send[pos] = a; }
        }
        kk += __popc(b);
      }
    }
  }
}

// pad the tail of the sparse arrays with N (partition.py:1024 `N * ones`)
template <typename T, int DIM>
__global__ void k_sparse_pad(NbrP<T, DIM> P, int gated) {
  GATE(P, gated);
  const long long cap = P.max_occupancy;
  long long start = P.offsets[P.n];
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = start + blockIdx.x * (long long)blockDim.x + threadIdx.x; p < cap; p += stride) {
    P.idx[p] = P.n;
    P.idx[cap + p] = P.n;
  }
}

template <typename T, int DIM>
__global__ void k_finalize(NbrP<T, DIM> P, int gated) {
  GATE(P, gated);
  long long stride = (long long)gridDim.x * blockDim.x;
  long long total = (long long)P.n * DIM;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride)
    P.ref[i] = P.position[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned e = *P.error;
    // partition.py:1066 CELL_LIST_OVERFLOW, :1110 NEIGHBOR_LIST_OVERFLOW
    if (P.use_cells && P.state[ST_MAX_CELL] > P.cell_capacity) e |= JMD_ERR_CELL_LIST_OVERFLOW;
    long long occ = P.format == JMD_DENSE ? P.state[ST_MAX_ROW] : P.state[ST_TOTAL];
    if (occ > P.max_occupancy) e |= JMD_ERR_NEIGHBOR_LIST_OVERFLOW;
    // internal rows are capacity-bounded too (documented in DESIGN.md)
    if (P.state[ST_MAX_ROW] > P.m_int) e |= JMD_ERR_NEIGHBOR_LIST_OVERFLOW;
    *P.error = (uint8_t)e;
    P.state[ST_BUILDS] += 1;
  }
}

// ---- launch plan (host or device) ----------------------------------------------------

template <typename T, int DIM>
__host__ __device__ inline size_t build_smem_bytes() {
  return (sizeof(typename Vec4<T>::type) * BUILD_CAP + sizeof(int) * (2 * BUILD_CAP + 4 * 32)) * BUILD_WARPS;
}

__host__ __device__ inline int grid_for(long long n, int block, int cap_blocks) {
  long long g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap_blocks) g = cap_blocks;
  return (int)g;
}

#ifdef __CUDA_ARCH__
#define JMD_STREAM cudaStreamTailLaunch
#else
#define JMD_STREAM stream
#endif

template <typename T, int DIM>
__host__ __device__ void launch_bin(const NbrP<T, DIM>& P, int gated, cudaStream_t stream) {
  const int G = JMD_SM_COUNT * 8;
  if (!P.use_cells) {
    k_identity_sort<T, DIM><<<grid_for(P.n, 256, G), 256, 0, JMD_STREAM>>>(P, gated);
    return;
  }
  const long long* gate = gated ? &P.state[ST_REBUILD] : nullptr;
  int tiles = (int)((P.n_cells + SCAN_TILE - 1) / SCAN_TILE);
  k_zero<T, DIM><<<grid_for(P.n_cells + 1, 256, G), 256, 0, JMD_STREAM>>>(P, gated);
  k_hash<T, DIM><<<grid_for(P.n, 256, G), 256, 0, JMD_STREAM>>>(P, gated);
  k_scan_tiles<int, int><<<tiles, 256, 0, JMD_STREAM>>>(P.cell_count, P.n_cells, P.scan_tmp,
                                                        &P.state[ST_MAX_CELL], gate);
  k_scan_top<int><<<1, 256, 0, JMD_STREAM>>>(P.scan_tmp, tiles, gate);
  k_scan_apply<int, int><<<tiles, 256, 0, JMD_STREAM>>>(P.cell_count, P.n_cells, P.scan_tmp,
                                                        P.cell_start, gate);
  k_scatter<T, DIM><<<grid_for(P.n, 256, G), 256, 0, JMD_STREAM>>>(P, gated);
  k_rank_sort<T, DIM><<<grid_for(P.n, 128, G * 2), 128, 0, JMD_STREAM>>>(P, gated);
  k_inv_perm<T, DIM><<<grid_for(P.n, 256, G), 256, 0, JMD_STREAM>>>(P, gated);
}

template <typename T, int DIM>
__host__ __device__ void launch_build(const NbrP<T, DIM>& P, int gated, cudaStream_t stream) {
  k_build_reset<T, DIM><<<1, 1, 0, JMD_STREAM>>>(P, gated);
  if (P.use_cells) {
#if JMD_BUILD_WARP_PER_CELL
    int g = grid_for(P.n_cells, BUILD_WARPS, JMD_SM_COUNT * 16);
    k_build_cells<T, DIM><<<g, BUILD_WARPS * 32, build_smem_bytes<T, DIM>(), JMD_STREAM>>>(P, gated);
#else
    int g = (P.n + TPA_BLOCK - 1) / TPA_BLOCK;
    if (g < 1) g = 1;
    if (P.format == JMD_ORDERED_SPARSE)
      k_build_cells_tpa<T, DIM, true><<<g, TPA_BLOCK, 0, JMD_STREAM>>>(P, gated);
    else
      k_build_cells_tpa<T, DIM, false><<<g, TPA_BLOCK, 0, JMD_STREAM>>>(P, gated);
#endif
  } else {
    k_build_all_pairs<T, DIM><<<grid_for((long long)P.n * 32, 128, JMD_SM_COUNT * 16), 128, 0, JMD_STREAM>>>(P, gated);
  }
}

template <typename T, int DIM>
__host__ __device__ void launch_export(const NbrP<T, DIM>& P, int gated, cudaStream_t stream) {
  const int G = JMD_SM_COUNT * 8;
  const long long* gate = gated ? &P.state[ST_REBUILD] : nullptr;
  if (P.format != JMD_DENSE) {
    int tiles = (int)(((long long)P.n + SCAN_TILE - 1) / SCAN_TILE);
    k_sparse_counts<T, DIM><<<grid_for(P.n, 256, G), 256, 0, JMD_STREAM>>>(P, gated);
    long long* tile_sums = (long long*)P.scan_tmp;   // free again here; host sizes it for n/2048 int64
    k_scan_tiles<int, long long><<<tiles, 256, 0, JMD_STREAM>>>(P.tmp_ids, P.n, tile_sums, nullptr, gate);
    k_scan_top<long long><<<1, 256, 0, JMD_STREAM>>>(tile_sums, tiles, gate);
    k_scan_apply<int, long long><<<tiles, 256, 0, JMD_STREAM>>>(P.tmp_ids, P.n, tile_sums, P.offsets, gate);
  }
  k_export<T, DIM><<<grid_for((long long)P.n * 32, 128, JMD_SM_COUNT * 16), 128, 0, JMD_STREAM>>>(P, gated);
  if (P.format != JMD_DENSE)
    k_sparse_pad<T, DIM><<<grid_for(P.max_occupancy / 4 + 1, 256, G), 256, 0, JMD_STREAM>>>(P, gated);
  k_finalize<T, DIM><<<grid_for((long long)P.n * DIM, 256, G), 256, 0, JMD_STREAM>>>(P, gated);
}

// Skin predicate (partition.py:1146-1154).  The last block to finish latches
// the decision into state[REBUILD]; in tail_launch mode it also enqueues the
// whole rebuild from the device (CUDA dynamic parallelism, tail-launch stream),
// which runs before the next kernel of the host stream starts.
template <typename T, int DIM>
__global__ void __launch_bounds__(256) k_skin(NbrP<T, DIM> P) {
  bool moved = false;
  int stride = gridDim.x * blockDim.x;
  if (!P.always_rebuild) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
      T a[DIM], b[DIM];
#pragma unroll
      for (int k = 0; k < DIM; ++k) { a[k] = P.position[(size_t)i * DIM + k]; b[k] = P.ref[(size_t)i * DIM + k]; }
      T d2 = dist2_exact<T, DIM>(P.sp, a, b);
      moved = moved || (d2 > P.threshold_sq);
    }
  }
  int any = __syncthreads_or(moved ? 1 : 0);
  if (threadIdx.x == 0) {
    if (any) atomicOr((unsigned long long*)&P.state[ST_PENDING], 1ull);
    __threadfence();
    unsigned long long t = atomicAdd((unsigned long long*)&P.state[ST_TICKET], 1ull);
    if (t == gridDim.x - 1) {
      __threadfence();
      long long reb = (atomicExch((unsigned long long*)&P.state[ST_PENDING], 0ull) != 0ull) || P.always_rebuild;
      P.state[ST_REBUILD] = reb;
      P.state[ST_TICKET] = 0;
      if (reb && P.tail_launch) {
        launch_bin<T, DIM>(P, 0, 0);
        launch_build<T, DIM>(P, 0, 0);
        launch_export<T, DIM>(P, 0, 0);
      }
    }
  }
}

// ---- host side ---------------------------------------------------------------------

template <typename T, int DIM>
int fill(NbrP<T, DIM>& P, const jmd_nbr_t* nb, const void* position) {
  if (!nb || nb->n < 0) return JMD_EINVAL;
  P.n = nb->n; P.format = nb->format; P.use_cells = nb->use_cells; P.mask_self = nb->mask_self;
  P.always_rebuild = nb->always_rebuild; P.n_cells = nb->n_cells; P.cell_capacity = nb->cell_capacity;
  P.m_int = nb->m_int;
  for (int k = 0; k < 3; ++k) P.cps[k] = nb->cps[k] > 0 ? nb->cps[k] : 1;
  P.count_only = 0; P.tail_launch = 0;
  P.n_pad = nb->n_pad; P.max_occupancy = nb->max_occupancy;
  for (int k = 0; k < DIM; ++k) P.cell_size[k] = (T)nb->cell_size[k];
  P.cutoff_sq = (T)nb->cutoff_sq; P.threshold_sq = (T)nb->threshold_sq;
  P.sp.init(nb->space);
  const bool periodic = nb->space.kind == JMD_SPACE_PERIODIC;
  // Dense re-tests with the opposite orientation (see candidate_test).
  P.two_sided = (nb->format == JMD_DENSE) && nb->use_cells && periodic;
  P.rev_only = (nb->format == JMD_DENSE) && !nb->use_cells;
  // rounding band: |d2(i,j) - d2(j,i)| <= band whenever either is near cutoff^2
  double u = sizeof(T) == 4 ? 5.9604644775390625e-08 : 1.1102230246251565e-16;
  double Lmax = 0;
  for (int k = 0; k < DIM; ++k) Lmax = nb->space.side[k] > Lmax ? nb->space.side[k] : Lmax;
  double c = sqrt(nb->cutoff_sq > 0 ? nb->cutoff_sq : 0.0);
  double delta = 10.0 * u * (Lmax > c ? Lmax : c);
  double band = DIM * (2.0 * (c + delta) + delta) * delta + 2.0 * (DIM + 1) * u * 1.01 * nb->cutoff_sq;
  P.band = (T)(2.0 * band);
  P.far = (T)Lmax;
  P.cell_count = nb->cell_count; P.cell_start = nb->cell_start; P.cell_cursor = nb->cell_cursor;
  P.scan_tmp = nb->scan_tmp; P.hash = nb->hash; P.tmp_ids = nb->tmp_ids; P.perm = nb->perm;
  P.inv_perm = nb->inv_perm;
  P.pos_sorted = (typename Vec4<T>::type*)nb->pos_sorted;
  P.nl = nb->nl; P.cnt = nb->cnt; P.cnt_lower = nb->cnt_lower; P.offsets = (long long*)nb->offsets;
  P.ref = (T*)nb->reference_position; P.idx = nb->idx; P.error = nb->error;
  P.state = (long long*)nb->state; P.species = nb->species;
  P.position = (const T*)position;
  return 0;
}

template <typename F>
int dispatch(const jmd_nbr_t* nb, F&& f) {
  if (!nb) return JMD_EINVAL;
  int dim = nb->space.dim;
  if (nb->dtype == JMD_F32 && dim == 3) return f(float(), std::integral_constant<int, 3>());
  if (nb->dtype == JMD_F32 && dim == 2) return f(float(), std::integral_constant<int, 2>());
  if (nb->dtype == JMD_F64 && dim == 3) return f(double(), std::integral_constant<int, 3>());
  if (nb->dtype == JMD_F64 && dim == 2) return f(double(), std::integral_constant<int, 2>());
  return JMD_EINVAL;
}

template <typename T, int DIM>
int ensure_smem() {
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(k_build_cells<T, DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)build_smem_bytes<T, DIM>());
    if (e != cudaSuccess) return (int)e;
    done = true;
  }
  return 0;
}

}  // namespace

extern "C" {

int jmd_nbr_skin_check(const jmd_nbr_t* nb, const void* position, int tail_launch, void* stream) {
  return dispatch(nb, [&](auto t, auto d) -> int {
    using T = decltype(t);
    constexpr int DIM = decltype(d)::value;
    NbrP<T, DIM> P;
    int rc = fill(P, nb, position);
    if (rc) return rc;
    P.tail_launch = tail_launch;
    if (tail_launch && (rc = ensure_smem<T, DIM>())) return rc;
    k_skin<T, DIM><<<grid_for(P.n, 256, JMD_SM_COUNT * 4), 256, 0, (cudaStream_t)stream>>>(P);
    JMD_LAUNCH_CHECK();
    return 0;
  });
}

int jmd_nbr_bin(const jmd_nbr_t* nb, const void* position, int gated, void* stream) {
  return dispatch(nb, [&](auto t, auto d) -> int {
    using T = decltype(t);
    constexpr int DIM = decltype(d)::value;
    NbrP<T, DIM> P;
    int rc = fill(P, nb, position);
    if (rc) return rc;
    launch_bin<T, DIM>(P, gated, (cudaStream_t)stream);
    JMD_LAUNCH_CHECK();
    return 0;
  });
}

int jmd_nbr_build(const jmd_nbr_t* nb, const void* position, int count_only, int gated, void* stream) {
  return dispatch(nb, [&](auto t, auto d) -> int {
    using T = decltype(t);
    constexpr int DIM = decltype(d)::value;
    NbrP<T, DIM> P;
    int rc = fill(P, nb, position);
    if (rc) return rc;
    if ((rc = ensure_smem<T, DIM>())) return rc;
    P.count_only = count_only;
    launch_build<T, DIM>(P, gated, (cudaStream_t)stream);
    JMD_LAUNCH_CHECK();
    return 0;
  });
}

int jmd_nbr_export(const jmd_nbr_t* nb, const void* position, int gated, void* stream) {
  return dispatch(nb, [&](auto t, auto d) -> int {
    using T = decltype(t);
    constexpr int DIM = decltype(d)::value;
    NbrP<T, DIM> P;
    int rc = fill(P, nb, position);
    if (rc) return rc;
    launch_export<T, DIM>(P, gated, (cudaStream_t)stream);
    JMD_LAUNCH_CHECK();
    return 0;
  });
}

int jmd_nbr_pack(const jmd_nbr_t* nb, const void* position, void* stream) {
  return dispatch(nb, [&](auto t, auto d) -> int {
    using T = decltype(t);
    constexpr int DIM = decltype(d)::value;
    NbrP<T, DIM> P;
    int rc = fill(P, nb, position);
    if (rc) return rc;
    k_pack<T, DIM><<<grid_for(P.n, 256, JMD_SM_COUNT * 8), 256, 0, (cudaStream_t)stream>>>(P);
    JMD_LAUNCH_CHECK();
    return 0;
  });
}

int jmd_nbr_state_host(const jmd_nbr_t* nb, int64_t* out, void* stream) {
  if (!nb || !out) return JMD_EINVAL;
  cudaError_t e = cudaMemcpyAsync(out, nb->state, sizeof(int64_t) * JMD_ST_COUNT, cudaMemcpyDeviceToHost,
                                  (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  e = cudaStreamSynchronize((cudaStream_t)stream);
  return (int)e;
}

}  // extern "C"

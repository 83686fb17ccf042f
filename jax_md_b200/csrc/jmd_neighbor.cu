// Neighbour-list build / update for sm_100a.
//
// Replaces the reference pipeline partition.py:349-471 (cell list by argsort +
// scatter) and partition.py:911-1154 (candidate gather, distance mask, cumsum
// compaction, skin predicate under lax.cond).  Every stage is a grid-stride
// "phase" (a __device__ function); the phases are used two ways:
//
//   k_update   a persistent kernel (one co-resident wave, hand-rolled grid
//              barrier) per NeighborList.update(): phase 0 is the skin
//              predicate (max displacement vs threshold); if no atom moved past
//              skin/2 every block returns, otherwise the same kernel bins and
//              sorts the atoms with grid-wide barriers between phases.  The
//              stencil scan (k_nbr_stencil_scan) and the export (k_update_c)
//              follow as gated launches that are empty on non-rebuild steps.
//              This is the lax.cond of partition.py:1146 with no host round
//              trip.
//   k_phase<>  one ordinary kernel per phase for the host-driven allocate path
//              (which needs occupancies on the host between stages) and for the
//              "gated" fallback update mode.
//
// Phases:  zero / hash+histogram / 3-pass exclusive scan / scatter / in-cell
// order (reference slot order) / stencil scan with exact reference arithmetic /
// sparse offsets / tile-transposed export to Dense|Sparse|OrderedSparse / error
// bits + reference positions.
#include <cuda_runtime.h>
#include <math.h>
#include <type_traits>
#include "jmd_common.cuh"

namespace {

enum { ST_EXPORT = JMD_ST_EXPORT_PENDING };
enum { ST_REBUILD = JMD_ST_REBUILD, ST_MAX_CELL = JMD_ST_MAX_CELL_OCC,
       ST_MAX_ROW = JMD_ST_MAX_ROW, ST_TOTAL = JMD_ST_TOTAL,
       ST_BUILDS = JMD_ST_BUILDS, ST_TICKET = JMD_ST_SCAN_TICKET,
       ST_PENDING = 6, ST_BARRIER = 7 };

#ifndef JMD_EXPORT_BATCH
#define JMD_EXPORT_BATCH 16     /* row entries per lane in flight while a tile is loaded */
#endif
#ifndef JMD_EXPORT_MIN_BLOCKS
#define JMD_EXPORT_MIN_BLOCKS 5   /* measured: 16 in flight at 5 blocks/SM, export 0.39 -> 0.33 ms */
#endif
#ifndef JMD_SCAN_UNROLL
#define JMD_SCAN_UNROLL 4
#endif
#ifndef JMD_SCAN_MIN_BLOCKS
#define JMD_SCAN_MIN_BLOCKS 4
#endif
constexpr int NB = 256;           // threads per block, every kernel here
constexpr int NWARP = NB / 32;
constexpr int SCAN_TILE = NB * 8;
constexpr int RANK_LIMIT = 4096;  // cells above this keep arrival order
constexpr int CELL_DIRTY = 1 << 30;  // cell_cursor bit: the cell holds an atom outside its cell's geometry

template <typename T, int DIM>
struct NbrP {
  int n, format, use_cells, mask_self, always_rebuild, n_cells, cell_capacity, m_int;
  int cps[3];          // INTERNAL (fine) search grid
  int bs, nb[3], rotate;   // storage order: bricks of (1 << bs)^DIM cells, nb bricks per side
  int* ref_start;          // [n_ref_cells + 1] exclusive scan of the counts in REFERENCE hash order
  int lazy_idx;            // update(): leave the public idx stale (EXPORT_PENDING) instead of exporting
  const int* skin_blk;     // per-drift-block predicate flags (jmd_nve_kick_drift)
  int skin_pre;            // skin_blk describes exactly this `position`
  int staged;              // maintain the force kernel's staging plan (blk_table, nl16)
  int stage_cap;           // staging entries per block (JMD_STAGE_BYTES / sizeof(Vec4<T>))
  int* blk_table;
  unsigned short* nl16;
  int ref_cps[3], n_ref_cells, sw;   // reference grid (capacity flag only), stencil half width
  int count_only, two_sided, rev_only, n_rows, no_public_idx;
  // warp-per-cell scan (jmd_nbr_cellscan.cuh)
  int cellscan, cs_chunks, cs_batches;
  unsigned* cs_bits;               // [n_cells][cs_batches][cs_chunks][32] accept masks
  unsigned long long* cs_lb;       // look-back words of the sparse offsets scan
  long long n_pad, max_occupancy;
  T cell_size[DIM];    // fine cell size
  T ref_cell_size[DIM];
  T cutoff_sq, threshold_sq, band, far;
  T f_lo, f_hi;        // pre-filter: a2 < f_lo accepts, a2 > f_hi rejects, else exact test
  int filter;          // pre-filter usable (periodic, cps >= 2*sw+3 on every axis)
  Space<T, DIM> sp;
  int *cell_count, *cell_start, *cell_cursor, *scan_tmp, *hash, *tmp_ids, *perm, *inv_perm, *ref_count;
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

struct Smem {
  int tile[NWARP][32][33];   // export: one padded 32x32 tile per warp
  long long scan[NWARP];
};

__device__ __forceinline__ int gtid() { return blockIdx.x * NB + threadIdx.x; }
// gated launches: 0 always run, 1 only when this update() decided to rebuild,
// 2 only when the public idx is stale (lazy materialisation)
__device__ __forceinline__ bool gate_closed(const long long* state, int gated) {
  return gated && state[gated == 2 ? (int)ST_EXPORT : (int)ST_REBUILD] == 0;
}
__device__ __forceinline__ int gthreads() { return gridDim.x * NB; }

// Grid-wide barrier for the persistent update kernels.  They are launched with
// exactly one co-resident wave of blocks (occupancy API x SM count) as ORDINARY
// launches: cudaLaunchCooperativeKernel costs ~35 us of device time per launch
// on B200 (measured), which is more than the skin predicate itself.  `bar` is a
// pair of zero-initialised counters {arrivals, exits}; the last block to leave
// the kernel (grid_exit) zeroes them for the next launch.
// The barrier assumes that every block of the launch is resident (the grid is sized by the
// occupancy API for the current device).  Should that ever not hold -- an SM-limited MPS
// partition, a debugger -- the spin gives up after ~2 s instead of hanging the device: the
// caller abandons the rebuild and both overflow bits are set, so the user-visible
// `did_buffer_overflow` asks for a fresh allocate() (which uses ordinary launches only).
__device__ __forceinline__ bool grid_sync(unsigned int* bar, unsigned int& target) {
  __shared__ int ok_s;
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(&bar[0], 1u);
    const long long t0 = clock64();
    int ok = 1;
    while (*((volatile unsigned int*)&bar[0]) < target) {
      if (clock64() - t0 > 4000000000ll) { ok = 0; break; }
    }
    __threadfence();
    ok_s = ok;
  }
  __syncthreads();
  return ok_s != 0;
}

template <typename T, int DIM>
__device__ __forceinline__ void grid_fault(const NbrP<T, DIM>& P, unsigned int* bar) {
  if (threadIdx.x == 0) {
    atomicOr((unsigned int*)((size_t)P.error & ~(size_t)3),
             (unsigned)(JMD_ERR_NEIGHBOR_LIST_OVERFLOW | JMD_ERR_CELL_LIST_OVERFLOW) << (8 * ((size_t)P.error & 3)));
    bar[0] = 0u;
    bar[1] = 0u;
  }
}
#define JMD_GRID_SYNC(P, bar, target) \
  do { if (!grid_sync(bar, target)) { grid_fault(P, bar); return; } } while (0)

__device__ __forceinline__ void grid_exit(unsigned int* bar) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&bar[1], 1u) == gridDim.x - 1) {
      bar[0] = 0u;
      bar[1] = 0u;
      __threadfence();
    }
  }
}

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
  // (periodic_general with fractional coordinates: the sorted copy is in real space)
  T q[3] = {T(0), T(0), T(0)};
  P.sp.to_real_v(r, q);
  T w = P.species ? (T)P.species[i] : T(0);
  return make_v4<T>(q[0], q[1], DIM == 3 ? q[DIM - 1] : T(0), w);
}


// ---- cell storage order ---------------------------------------------------------
// The reference hashes cells x-fastest (partition.py:212-224); that order only
// matters for (a) the order candidates are visited in -- the stencil walk below
// follows it explicitly -- and (b) the slot rotation of partition.py:441, which
// needs the reference-order prefix sums (ref_start).  WHERE a cell's atoms live in
// the cell-sorted arrays is ours to choose: cells are stored brick by brick
// ((1 << bs)^DIM cells per brick, bricks x-fastest) so that the 32 slots of a warp
// and the 256 of a block are spatially compact and the neighbour gathers of the
// force kernel stay inside the L1.  bs == 0 is the plain reference order.
template <typename T, int DIM>
__device__ __forceinline__ int cell_id(const NbrP<T, DIM>& P, const int (&v)[3]) {
  if (P.bs == 0) return v[0] + P.cps[0] * (v[1] + (DIM == 3 ? P.cps[1] * v[2] : 0));
  const int bs = P.bs, m = (1 << bs) - 1;
  const int brick = (v[0] >> bs) + P.nb[0] * ((v[1] >> bs) + (DIM == 3 ? P.nb[1] * (v[2] >> bs) : 0));
  const int inner = (v[0] & m) | ((v[1] & m) << bs) | (DIM == 3 ? (v[2] & m) << (2 * bs) : 0);
  return (brick << (DIM * bs)) | inner;
}

template <typename T, int DIM>
__device__ __forceinline__ void cell_coords(const NbrP<T, DIM>& P, int id, int (&v)[3]) {
  if (P.bs == 0) {
    v[0] = id % P.cps[0];
    v[1] = (id / P.cps[0]) % P.cps[1];
    v[2] = DIM == 3 ? id / (P.cps[0] * P.cps[1]) : 0;
    return;
  }
  const int bs = P.bs, m = (1 << bs) - 1;
  const int inner = id & ((1 << (DIM * bs)) - 1);
  const int brick = id >> (DIM * bs);
  const int b0 = brick % P.nb[0], b1 = (brick / P.nb[0]) % P.nb[1];
  const int b2 = DIM == 3 ? brick / (P.nb[0] * P.nb[1]) : 0;
  v[0] = (b0 << bs) | (inner & m);
  v[1] = (b1 << bs) | ((inner >> bs) & m);
  v[2] = DIM == 3 ? (b2 << bs) | ((inner >> (2 * bs)) & m) : 0;
}

// ---- phase: skin predicate (partition.py:1146-1154) ------------------------------
template <typename T, int DIM>
__device__ bool ph_skin(const NbrP<T, DIM>& P) {
  bool moved = false;
  if (P.always_rebuild) return true;
  for (int i = gtid(); i < P.n_rows; i += gthreads()) {
    T a[DIM], b[DIM];
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
      a[k] = P.position[(size_t)i * DIM + k];
      b[k] = P.ref[(size_t)i * DIM + k];
    }
    T d2 = dist2_exact<T, DIM>(P.sp, a, b);
    moved = moved || (d2 > P.threshold_sq);      // strict >, min-image metric
  }
  return moved;
}

// ---- phases: binning -----------------------------------------------------------------
template <typename T, int DIM>
__device__ void ph_zero(const NbrP<T, DIM>& P) {
  for (int c = gtid(); c <= P.n_cells; c += gthreads()) {
    P.cell_count[c] = 0;
    if (c < P.n_cells) P.cell_cursor[c] = 0;
    if (c < P.n_ref_cells) P.ref_count[c] = 0;
  }
  if (gtid() == 0) P.state[ST_MAX_CELL] = 0;
}

// partition.py:421-423: int32(R / cell_size) (truncation), mod cells_per_side,
// hash = x + y*cx + z*cx*cy.  The reference grid only feeds the occupancy
// histogram behind cell_list_capacity / CELL_LIST_OVERFLOW; atoms are binned on
// the finer internal grid (same formula, cell size / fine).
template <typename T, int DIM>
__device__ void ph_hash(const NbrP<T, DIM>& P) {
  for (int i = gtid(); i < P.n; i += gthreads()) {
    const T* r = P.position + (size_t)i * DIM;
    int hr = 0, multr = 1;
    int cv[3] = {0, 0, 0};
    bool regular = true;
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
      int ci = (int)div_rn(r[k], P.cell_size[k]);
      // "regular": the coordinate lies in the primary box and in the cell it is
      // binned to (no wrap by the mod below) -- what the scan's pre-filter assumes
      regular = regular && (r[k] >= T(0)) && (ci >= 0) && (ci < P.cps[k]);
      ci %= P.cps[k];
      if (ci < 0) ci += P.cps[k];
      cv[k] = ci;
      int cr = (int)div_rn(r[k], P.ref_cell_size[k]);
      cr %= P.ref_cps[k];
      if (cr < 0) cr += P.ref_cps[k];
      hr += cr * multr;
      multr *= P.ref_cps[k];
    }
    const int h = cell_id(P, cv);
    P.hash[i] = h;
    atomicAdd(&P.cell_count[h], 1);
    atomicAdd(&P.ref_count[hr], 1);
    if (!regular) atomicOr(&P.cell_cursor[h], CELL_DIRTY);   // exact test for this cell
  }
}

// max occupancy of a REFERENCE cell (partition.py:243-250, 458-460)
template <typename T, int DIM>
__device__ void ph_ref_max(const NbrP<T, DIM>& P) {
  int mx = 0;
  for (int c = gtid(); c < P.n_ref_cells; c += gthreads()) mx = max(mx, P.ref_count[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0)
    atomicMax((unsigned long long*)&P.state[ST_MAX_CELL], (unsigned long long)mx);
}

// exclusive scan of one value per thread across the block
template <typename TOut>
__device__ __forceinline__ TOut block_excl_scan(TOut x, TOut* total, long long* smem /*[NWARP]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  TOut incl = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    TOut y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) smem[warp] = (long long)incl;
  __syncthreads();
  TOut wbase = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < NWARP; ++w) {
    TOut s = (TOut)smem[w];
    if (w < warp) wbase += s;
    tot += s;
  }
  __syncthreads();
  *total = tot;
  return wbase + incl - x;
}

template <typename TIn, typename TOut>
__device__ __forceinline__ void scan_tile_load(const TIn* in, long long n, long long base, TOut (&v)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    long long i = base + (long long)threadIdx.x * 8 + j;
    v[j] = i < n ? (TOut)in[i] : TOut(0);
  }
}

// three-pass exclusive scan of in[0..n) -> out[0..n]; tile_sums scratch
template <typename TIn, typename TOut>
__device__ void ph_scan_tiles(const TIn* in, long long n, TOut* tile_sums, long long* max_out, Smem& sm) {
  const long long tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
    TOut v[8];
    scan_tile_load<TIn, TOut>(in, n, t * SCAN_TILE, v);
    TOut s = 0, mx = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { s += v[j]; mx = v[j] > mx ? v[j] : mx; }
    TOut tot;
    block_excl_scan<TOut>(s, &tot, sm.scan);
    if (threadIdx.x == 0) tile_sums[t] = tot;
    if (max_out) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        TOut y = __shfl_xor_sync(0xffffffffu, mx, o);
        mx = y > mx ? y : mx;
      }
      if ((threadIdx.x & 31) == 0 && mx > 0)
        atomicMax((unsigned long long*)max_out, (unsigned long long)mx);
    }
  }
}

template <typename TOut>
__device__ void ph_scan_top(TOut* tile_sums, long long n, Smem& sm) {
  if (blockIdx.x != 0) return;
  const int tiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
  TOut carry = 0;
  for (int base = 0; base < tiles; base += NB) {
    int i = base + threadIdx.x;
    TOut x = i < tiles ? tile_sums[i] : TOut(0);
    TOut tot;
    TOut e = block_excl_scan<TOut>(x, &tot, sm.scan);
    if (i < tiles) tile_sums[i] = carry + e;
    carry += tot;
  }
}

template <typename TIn, typename TOut>
__device__ void ph_scan_apply(const TIn* in, long long n, const TOut* tile_sums, TOut* out, Smem& sm) {
  const long long tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
    TOut v[8];
    const long long base = t * SCAN_TILE;
    scan_tile_load<TIn, TOut>(in, n, base, v);
    TOut s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];
    TOut tot;
    TOut e = block_excl_scan<TOut>(s, &tot, sm.scan) + tile_sums[t];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      long long i = base + (long long)threadIdx.x * 8 + j;
      if (i < n) out[i] = e;
      e += v[j];
      if (i == n - 1) out[n] = e;
    }
  }
  if (n == 0 && gtid() == 0) out[0] = 0;
}

template <typename T, int DIM>
__device__ void ph_scatter(const NbrP<T, DIM>& P) {
  for (int i = gtid(); i < P.n; i += gthreads()) {
    int h = P.hash[i];
    int pos = P.cell_start[h] + (atomicAdd(&P.cell_cursor[h], 1) & (CELL_DIRTY - 1));
    P.tmp_ids[pos] = i;
    P.inv_perm[i] = pos;
  }
}

// Order inside a cell.  The reference sorts atoms by hash with a stable argsort
// (partition.py:432: arrival order == id order) and then places sorted rank r
// in slot `r mod cell_capacity` (partition.py:441), so reading a cell in slot
// order yields the arrival order ROTATED by r0 = cap - start % cap (when that
// is < count).  When the search grid IS the reference grid (fine == 1) we store
// each cell directly in that slot order, so the stencil scan walks plain
// contiguous ranges and emits candidates in the reference's order; on a finer
// grid the order is ascending id (sets stay identical, order is ours).
template <typename T, int DIM>
__device__ void ph_rank_sort(const NbrP<T, DIM>& P) {
  const int cap = P.cell_capacity > 0 ? P.cell_capacity : 1;
  const bool rotate = P.rotate != 0;
  for (int i = gtid(); i < P.n; i += gthreads()) {
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
    if (rotate) {
      // sorted rank of the cell's first atom in the REFERENCE's hash order
      int s_ref = s;
      if (P.bs > 0) {
        int v[3];
        cell_coords(P, h, v);
        s_ref = P.ref_start[v[0] + P.cps[0] * (v[1] + (DIM == 3 ? P.cps[1] * v[2] : 0))];
      }
      const int room = cap - s_ref % cap;
      const int r0 = room < c ? room : 0;
      rank -= r0;
      rank = rank < 0 ? rank + c : rank;
    }
    int dst = s + rank;
    P.perm[dst] = i;
    P.pos_sorted[dst] = load_atom(P, i);
  }
}

// inv_perm held the arrival position for ph_rank_sort; now the final slot.
template <typename T, int DIM>
__device__ void ph_inv_perm(const NbrP<T, DIM>& P) {
  for (int t = gtid(); t < P.n; t += gthreads()) P.inv_perm[P.perm[t]] = t;
}


// ---- phase: staging plan of the force kernel (see jmd_common.cuh) ---------------------
// One thread per 256-slot block: the block's home cells are a run of consecutive
// cells in the x-fastest order, i.e. a few row segments; every (dy, dz)
// stencil row of a segment is one contiguous slot range, or two when the
// x-interval wraps around the box.  Blocks that do not fit (more than
// JMD_STAGE_SEGS segments, more atoms than the staging buffer) are marked "direct".
template <typename T, int DIM>
__device__ void ph_plan(const NbrP<T, DIM>& P) {
  if (!P.staged) return;
  const int nblk = (P.n + JMD_STAGE_BLOCK - 1) / JMD_STAGE_BLOCK;
  for (int b = gtid(); b < nblk; b += gthreads()) {
    int* tb = P.blk_table + (size_t)b * JMD_TBL_INTS;
    for (int i = 0; i < JMD_TBL_INTS; ++i) tb[i] = 0;
    tb[JMD_TBL_ROW0] = -1;
    if (!P.use_cells || P.sw != 1 || P.bs != 0) continue;          // mode 0: direct
    const int s_first = b * JMD_STAGE_BLOCK;
    const int s_last = min(s_first + JMD_STAGE_BLOCK, P.n) - 1;
    const int c_first = P.hash[P.perm[s_first]], c_last = P.hash[P.perm[s_last]];
    const int cx = P.cps[0], cy = P.cps[1], cz = DIM == 3 ? P.cps[2] : 1;
    const int r_first = c_first / cx, r_last = c_last / cx;
    if (r_last - r_first >= JMD_STAGE_SEGS) continue;
    const int nseg = r_last - r_first + 1;
    tb[JMD_TBL_ROW0] = r_first;
    tb[JMD_TBL_NSEG] = nseg;
    int total = 0;
    bool fits = true;
    for (int seg = 0; seg < nseg; ++seg) {
      const int row = r_first + seg;
      const int y = row % cy, z = row / cy;
      const int x0 = (seg == 0) ? c_first % cx : 0;
      const int x1 = (row == r_last) ? c_last % cx : cx - 1;
      // x-interval [x0-1, x1+1], periodic: piece 0 / piece 1
      int xa[2], xb[2];
      xa[1] = 0; xb[1] = -1;
      if (x1 - x0 + 3 >= cx) { xa[0] = 0; xb[0] = cx - 1; }
      else if (x0 - 1 < 0) { xa[0] = 0; xb[0] = x1 + 1; xa[1] = cx - 1; xb[1] = cx - 1; }
      else if (x1 + 1 >= cx) { xa[0] = x0 - 1; xb[0] = cx - 1; xa[1] = 0; xb[1] = 0; }
      else { xa[0] = x0 - 1; xb[0] = x1 + 1; }
      for (int dy = -1; dy <= 1; ++dy)
      for (int dz = (DIM == 3 ? -1 : 0); dz <= (DIM == 3 ? 1 : 0); ++dz) {
        int yy = y + dy, zz = z + dz;
        yy = yy < 0 ? yy + cy : (yy >= cy ? yy - cy : yy);
        zz = zz < 0 ? zz + cz : (zz >= cz ? zz - cz : zz);
        const int rowbase = (zz * cy + yy) * cx;
        for (int piece = 0; piece < 2; ++piece) {
          if (xb[piece] < xa[piece]) continue;
          const int g0 = P.cell_start[rowbase + xa[piece]];
          const int len = P.cell_start[rowbase + xb[piece] + 1] - g0;
          int* e = tb + JMD_TBL_ENTRY(seg, dy, dz, piece);
          e[0] = g0; e[1] = len; e[2] = total;
          total += len;
          fits = fits && total <= P.stage_cap;
        }
      }
    }
    tb[JMD_TBL_TOTAL] = total;
    tb[JMD_TBL_MODE] = fits ? 1 : 0;
  }
}

template <typename T, int DIM>
__device__ void ph_build_reset(const NbrP<T, DIM>& P) {
  if (gtid() == 0) {
    P.state[ST_MAX_ROW] = 0;
    P.state[ST_TOTAL] = 0;
    P.state[9] = 0;                       // ST_LB_TILE: tile counter of the offsets scan
  }
  if (P.cs_lb) {                          // its look-back words start from zero
    const int tiles = (P.n + SCAN_TILE - 1) / SCAN_TILE + 1;
    for (int i = gtid(); i < tiles; i += gthreads()) P.cs_lb[i] = 0ull;
  }
}

// all-pairs path: identity order.
template <typename T, int DIM>
__device__ void ph_identity_sort(const NbrP<T, DIM>& P) {
  for (int i = gtid(); i < P.n; i += gthreads()) {
    P.perm[i] = i;
    P.inv_perm[i] = i;
    P.pos_sorted[i] = load_atom(P, i);
  }
  if (gtid() == 0) P.state[ST_MAX_CELL] = 0;
}

template <typename T, int DIM>
__global__ void k_pack_range(NbrP<T, DIM> P, int first, int count) {
  const int stride = gridDim.x * blockDim.x;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < count; g += stride) {
    const int a = first + g;
    P.pos_sorted[P.inv_perm[a]] = load_atom(P, a);
  }
}

template <typename T, int DIM>
__device__ void ph_pack(const NbrP<T, DIM>& P) {
  for (int t = gtid(); t < P.n; t += gthreads()) P.pos_sorted[t] = load_atom(P, P.perm[t]);
}

// ---- candidate test -------------------------------------------------------------------
// Reference semantics (oracle/partition.py): the cell path tests
// d2(R_i, R_c) < cutoff^2 (partition.py:945-951); Dense then re-tests with the
// opposite orientation d2(R_c, R_i) (prune_neighbor_list_dense via map_neighbor,
// partition.py:960-980, space.py:494-502).  The two differ only by rounding, so
// the second one is evaluated only inside a rounding band around the cutoff.
template <typename T, int DIM>
__device__ __forceinline__ bool candidate_test(const NbrP<T, DIM>& P, const T* hp, const T* cp) {
  bool keep;
  // (periodic_general: callers pass the user's own coordinates, see general_keep)
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

__device__ __forceinline__ void publish_counts(long long* state, long long my_k, long long my_tot) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    long long y = __shfl_xor_sync(0xffffffffu, my_k, o);
    my_k = y > my_k ? y : my_k;
    my_tot += __shfl_xor_sync(0xffffffffu, my_tot, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (my_k > 0) atomicMax((unsigned long long*)&state[ST_MAX_ROW], (unsigned long long)my_k);
    if (my_tot > 0) atomicAdd((unsigned long long*)&state[ST_TOTAL], (unsigned long long)my_tot);
  }
}

// Thread-per-atom stencil scan.  A thread owns sorted slot `slot` and walks the
// (2w+1)^d stencil of its own FINE cell (cells whose gap to the home cell
// exceeds the cutoff are skipped: uniform over the warp); the candidates of a
// cell are a contiguous range of the cell-sorted float4 array, so the lanes of
// a warp issue loads that hit a handful of distinct addresses (L1 broadcast).
// The fine grid (reference cell / 2) cuts the candidates per atom from ~533
// to ~300 at the LJ benchmark density while the accepted SET stays exactly the
// reference's: the distance test is the reference's arithmetic, bit for bit.
// Every row is appended by its own thread: no ballots, no atomics,
// deterministic order (stencil order, then atom id).  Row k of the transposed
// list is written by neighbouring lanes at neighbouring addresses.
// MODE: 0 = forward test only (Sparse formats), 1 = forward + reverse inside
// the rounding band (Dense).
// Pins a loop-invariant value in a register: the empty asm makes it opaque, so
// ptxas cannot rematerialise it (recompute / reload constants) inside the loop.
__device__ __forceinline__ void keep_in_register(float& x) { asm volatile("" : "+f"(x)); }
__device__ __forceinline__ void keep_in_register(double& x) { asm volatile("" : "+d"(x)); }
__device__ __forceinline__ void keep_in_register(int& x) { asm volatile("" : "+r"(x)); }

// keep = (a2 < lo) && rank != self;  if (keep && off < off_end) { nl[off] = rank; off += n_pad; }
// k += keep.  `off` is a 32-bit ELEMENT offset into nl (m_int * n_pad < 2^32 is checked
// on the host).  Spelled in PTX so the append stays a handful of predicated instructions
// (ptxas otherwise turns the selects into a chain of moves or a branch).
#define JMD_APPEND_ASM(FT, FC)                                                                   \
  if (STAGE) {                                                                                   \
    asm volatile(                                                                                \
        "{\n\t.reg .pred pk, pw;\n\t.reg .u64 ad;\n\t.reg .u32 cd;\n\t.reg .u16 ch;\n\t"     \
        "setp.lt." FT " pk, %2, %3;\n\t"                                                        \
        "setp.ne.and.s32 pk, %4, %5, pk;\n\t"                                                   \
        "setp.lt.and.u32 pw, %0, %6, pk;\n\t"                                                   \
        "mad.wide.u32 ad, %0, 4, %8;\n\t"                                                       \
        "@pw st.global.s32 [ad], %4;\n\t"                                                       \
        "mad.wide.u32 ad, %0, 2, %9;\n\t"                                                       \
        "add.s32 cd, %4, %10;\n\t"                                                              \
        "cvt.u16.u32 ch, cd;\n\t"                                                               \
        "@pw st.global.u16 [ad], ch;\n\t"                                                       \
        "@pw add.u32 %0, %0, %7;\n\t"                                                           \
        "@pk add.s32 %1, %1, 1;\n\t}"                                                           \
        : "+r"(off), "+r"(k)                                                                     \
        : FC(a2), FC(lo), "r"(rank), "r"(self), "r"(off_end), "r"(n_pad), "l"(nl), "l"(nl16),    \
          "r"(lbase)                                                                             \
        : "memory");                                                                             \
  } else {                                                                                       \
    asm volatile(                                                                                \
        "{\n\t.reg .pred pk, pw;\n\t.reg .u64 ad;\n\t"                                         \
        "setp.lt." FT " pk, %2, %3;\n\t"                                                        \
        "setp.ne.and.s32 pk, %4, %5, pk;\n\t"                                                   \
        "setp.lt.and.u32 pw, %0, %6, pk;\n\t"                                                   \
        "mad.wide.u32 ad, %0, 4, %8;\n\t"                                                       \
        "@pw st.global.s32 [ad], %4;\n\t"                                                       \
        "@pw add.u32 %0, %0, %7;\n\t"                                                           \
        "@pk add.s32 %1, %1, 1;\n\t}"                                                           \
        : "+r"(off), "+r"(k)                                                                     \
        : FC(a2), FC(lo), "r"(rank), "r"(self), "r"(off_end), "r"(n_pad), "l"(nl)                \
        : "memory");                                                                             \
  }
template <bool STAGE>
__device__ __forceinline__ void append_if_below(float a2, float lo, int rank, int self, unsigned off_end,
                                                unsigned n_pad, int* nl, unsigned short* nl16, int lbase,
                                                unsigned& off, int& k) {
  JMD_APPEND_ASM("f32", "f");
}
template <bool STAGE>
__device__ __forceinline__ void append_if_below(double a2, double lo, int rank, int self, unsigned off_end,
                                                unsigned n_pad, int* nl, unsigned short* nl16, int lbase,
                                                unsigned& off, int& k) {
  JMD_APPEND_ASM("f64", "d");
}
#undef JMD_APPEND_ASM

// periodic_general: the reference evaluates the metric on the USER's coordinates (unit cube
// or real space, space.py:419-433), not on the real-space sorted copy: fetch them by id.
// Dense (MODE 1) keeps a pair only if both orientations pass (partition.py:960-980).
template <typename T, int DIM, int MODE, bool TRIC_OK = true>
__device__ __forceinline__ bool general_keep(const NbrP<T, DIM>& P, int home_id, int cand_id, T c2) {
  T a[DIM], b[DIM];
#pragma unroll
  for (int k = 0; k < DIM; ++k) {
    a[k] = P.position[(size_t)home_id * DIM + k];
    b[k] = P.position[(size_t)cand_id * DIM + k];
  }
  bool keep = dist2_exact<T, DIM, TRIC_OK>(P.sp, a, b) < c2;
  if (MODE == 1) keep = keep && (dist2_exact<T, DIM, TRIC_OK>(P.sp, b, a) < c2);
  return keep;
}

// The reference's candidate test on one (home, candidate) pair, bit for bit:
// forward d2(R_i, R_c) < cutoff^2 (partition.py:945-951) and, for Dense (MODE 1),
// the reverse-orientation re-test inside the rounding band (partition.py:960-980).
template <typename T, int DIM, int MODE, bool PERIODIC, bool TRIC_OK = true>
__device__ __forceinline__ bool exact_keep(const NbrP<T, DIM>& P, const T (&hp)[3],
                                           const typename Vec4<T>::type& cv, const T (&hh)[3],
                                           const T (&qq)[3], T c2) {
  // forward displacement d(R_i, R_c), exact (space.py:213-235)
  T dd[3];
  dd[0] = sub_rn(hp[0], cv.x);
  dd[1] = sub_rn(hp[1], cv.y);
  dd[2] = DIM == 3 ? sub_rn(hp[2], cv.z) : T(0);
  T d2;
  bool slow = false;
  if (!PERIODIC) {
    d2 = mul_rn(dd[0], dd[0]);
#pragma unroll
    for (int d = 1; d < DIM; ++d) d2 = add_rn(d2, mul_rn(dd[d], dd[d]));
  } else {
#pragma unroll
    for (int d = 0; d < DIM; ++d) slow = slow || !(fabs(dd[d]) <= qq[d]);
    if (!slow) {
      // |d| <= L/4: mod() is the identity on fl(d + h) (see Space::disp)
      T m0 = sub_rn(add_rn(dd[0], hh[0]), hh[0]);
      d2 = mul_rn(m0, m0);
#pragma unroll
      for (int d = 1; d < DIM; ++d) {
        T m = sub_rn(add_rn(dd[d], hh[d]), hh[d]);
        d2 = add_rn(d2, mul_rn(m, m));
      }
    } else {
      const T cp[3] = {cv.x, cv.y, cv.z};
      d2 = dist2_exact<T, DIM, TRIC_OK>(P.sp, hp, cp);
    }
  }
  bool keep = d2 < c2;
  if (MODE == 1 && PERIODIC) {
    if (slow || (fabs(d2 - c2) <= P.band)) {
      const T cp[3] = {cv.x, cv.y, cv.z};
      keep = keep && (dist2_exact<T, DIM, TRIC_OK>(P.sp, cp, hp) < c2);
    }
  }
  return keep;
}

template <typename T, int DIM, int MODE, bool ORDERED, bool PERIODIC, int WSTAT, bool FILTER, bool COUNT, bool STAGE>
__device__ void ph_build_cells(const NbrP<T, DIM>& P) {
  using V4 = typename Vec4<T>::type;
  const int lane = threadIdx.x & 31;
  const int self_on = P.mask_self;
  const T c2 = P.cutoff_sq;
  const int kmax = P.count_only ? 0 : P.m_int;
  const int w = WSTAT > 0 ? WSTAT : P.sw;      // WSTAT == 1: the reference 3^d stencil, unrolled
  const int wz = DIM == 3 ? w : 0;
  // gap^2 between the home cell and a stencil cell, with a safety margin so the
  // pruning can never drop a cell that holds an acceptable candidate
  const float gap_limit = (float)c2 * 1.0001f + 1e-6f;
  T hh[3], qq[3];
#pragma unroll
  for (int d = 0; d < DIM; ++d) { hh[d] = P.sp.half[d]; qq[d] = P.sp.quarter[d]; }
  const V4* __restrict__ const pos = P.pos_sorted;
  __shared__ int tbl[JMD_TBL_INTS];              // this block's staging plan (ph_plan)
  for (int blk = blockIdx.x; blk * NB < P.n; blk += gridDim.x) {
    const int slot = blk * NB + threadIdx.x;
    (void)lane;
    long long my_k = 0, my_tot = 0;
    const int hid = slot < P.n ? P.perm[slot] : 0x7fffffff;
    if (STAGE) {
      __syncthreads();
      if (threadIdx.x < JMD_TBL_INTS) tbl[threadIdx.x] = P.blk_table[(size_t)blk * JMD_TBL_INTS + threadIdx.x];
      __syncthreads();
    }
    if (slot < P.n && hid >= P.n_rows) {      // ghost atom: candidate only, no row
      P.cnt[slot] = 0;
      P.cnt_lower[slot] = 0;
    }
    if (hid < P.n_rows) {
      const V4 hv = P.pos_sorted[slot];
      const T hp[3] = {hv.x, hv.y, hv.z};
      const int c = P.hash[hid];              // own (fine) cell
      int cc[3];
      cell_coords(P, c, cc);
      // staging: which row segment of the block this home cell belongs to
      const bool stage_on = STAGE && tbl[JMD_TBL_MODE] != 0;
      const int seg = STAGE ? c / P.cps[0] - tbl[JMD_TBL_ROW0] : 0;
      int self = self_on ? slot : -1;
      keep_in_register(self);
      const bool home_dirty = FILTER ? (__ldg(&P.cell_cursor[c]) & CELL_DIRTY) != 0 : false;
      int k = 0, kl = 0;
      int* const out = P.nl + slot;
      // 32-bit element offsets of this slot's next / past-the-end row entry
      unsigned off = (unsigned)slot;
      int off_end_i = (int)((unsigned)kmax * (unsigned)P.n_pad + (unsigned)slot);
      keep_in_register(off_end_i);
      const unsigned off_end = (unsigned)off_end_i;
      // One stencil cell: its slot range, dirty flag and the home position shifted
      // into its periodic image.
      struct Cell { int start, end, dirty, s1, s2; T hs0, hs1, hs2; bool skip; };
      auto lookup = [&](const int s0, const int s1, const int s2) -> Cell {
        const int sh[3] = {s0, s1, s2};
        float gap2 = 0.f;
        int sv[3] = {0, 0, 0};
        // FILTER: home position moved into the stencil cell's periodic image, so
        // hs - R_c is the minimum-image displacement up to rounding (valid for
        // regular atoms: |hs - R_c| <= (w+1) cell sizes < L/2)
        T hs[3] = {hp[0], hp[1], hp[2]};
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
          if (WSTAT != 1) {
            const int e = abs(sh[d]) - 1;
            const float g = e > 0 ? (float)e * (float)P.cell_size[d] : 0.f;
            gap2 += g * g;
          }
          int v = cc[d] + sh[d];
          if (v < 0) { v += P.cps[d]; hs[d] = hp[d] + P.sp.side[d]; }
          else if (v >= P.cps[d]) { v -= P.cps[d]; hs[d] = hp[d] - P.sp.side[d]; }
          sv[d] = v;
        }
        Cell C;
        C.s1 = s1; C.s2 = s2;
        C.hs0 = hs[0]; C.hs1 = hs[1]; C.hs2 = hs[2];
        C.skip = WSTAT != 1 && gap2 > gap_limit;
        C.start = C.end = C.dirty = 0;
        if (!C.skip) {
          const int h = cell_id(P, sv);
          C.start = __ldg(&P.cell_start[h]);
          C.end = __ldg(&P.cell_start[h + 1]);
          C.dirty = FILTER ? (__ldg(&P.cell_cursor[h]) & CELL_DIRTY) : 0;
        }
        return C;
      };
      auto scan_cell = [&](const Cell& C) {
        if (C.skip) return;
        const int start = C.start, end = C.end, s1 = C.s1, s2 = C.s2;
        const T hs[3] = {C.hs0, C.hs1, C.hs2};
        (void)s1; (void)s2;
        // staging index of a candidate = lbase + its slot (see ph_plan)
        int lbase = 0;
        if (STAGE && stage_on) {
          const int* e = tbl + JMD_TBL_ENTRY(seg, WSTAT == 1 ? s1 : 0, (WSTAT == 1 && DIM == 3) ? s2 : 0, 0);
          if (!(start >= e[0] && start < e[0] + e[1])) e += 3;         // second piece (x wrap)
          lbase = e[2] - e[0];
        }
        T lo_c = P.f_lo, hi_c = P.f_hi;
        if (FILTER) {
          if (home_dirty || C.dirty) {
            lo_c = T(-1);                       // a2 >= 0: never accepted unseen
            hi_c = (T)INFINITY;                 // never rejected unseen
          }
        }
        if (FILTER) {
          // Branch-free common path: contracted arithmetic, predicated append.  The
          // exact reference arithmetic runs only inside the rounding band around
          // cutoff^2 (or for cells flagged dirty), a rarely taken branch.
          T h0 = hs[0], h1 = hs[1], h2 = hs[2];
          keep_in_register(h0); keep_in_register(h1); keep_in_register(h2);
          keep_in_register(lo_c); keep_in_register(hi_c);
          auto process = [&](const int rank, const V4& cv) {
            const T ax = h0 - cv.x, ay = h1 - cv.y;
            T a2 = ax * ax + ay * ay;
            if (DIM == 3) { const T az = h2 - cv.z; a2 += az * az; }
            if (!(a2 < lo_c) && !(a2 > hi_c))     // rare: inside the band / dirty cell
              a2 = (P.sp.general ? general_keep<T, DIM, MODE, !FILTER>(P, hid, __ldg(&P.perm[rank]), c2)
                                 : exact_keep<T, DIM, MODE, PERIODIC, !FILTER>(P, hp, cv, hh, qq, c2)) ? T(-2) : (T)INFINITY;
            const int k_before = k;
            append_if_below<STAGE>(a2, lo_c, rank, self, off_end, (unsigned)P.n_pad, P.nl, P.nl16, lbase, off, k);
            if (ORDERED && COUNT) { if (k != k_before) kl += (__ldg(&P.perm[rank]) < hid); }
          };
          int rank = start;
          // JMD_SCAN_UNROLL candidates per trip: their loads are all in flight
          // before the first one is tested (measured on the rebuild: 1 -> 2 -14 %, 2 -> 4 -4 %)
          for (; rank + (JMD_SCAN_UNROLL - 1) < end; rank += JMD_SCAN_UNROLL) {
            V4 cvs[JMD_SCAN_UNROLL];
#pragma unroll
            for (int u = 0; u < JMD_SCAN_UNROLL; ++u) cvs[u] = pos[rank + u];
#pragma unroll
            for (int u = 0; u < JMD_SCAN_UNROLL; ++u) process(rank + u, cvs[u]);
          }
          for (; rank < end; ++rank) process(rank, pos[rank]);
        } else {
          for (int rank = start; rank < end; ++rank) {
            const V4 cv = pos[rank];
            bool keep = P.sp.general ? general_keep<T, DIM, MODE, !FILTER>(P, hid, __ldg(&P.perm[rank]), c2)
                                     : exact_keep<T, DIM, MODE, PERIODIC, !FILTER>(P, hp, cv, hh, qq, c2);
            keep = keep && (rank != self);
            if (keep) {
              if (off < off_end) {
                P.nl[off] = rank;
                if (STAGE) P.nl16[off] = (unsigned short)(lbase + rank);
                off += (unsigned)P.n_pad;
              }
              ++k;
              if (ORDERED && COUNT) kl += (__ldg(&P.perm[rank]) < hid);
            }
          }
        }
      
      };
      if (WSTAT == 1) {
        // The cell-table loads of stencil cell s+1 are issued before the candidates
        // of cell s are scanned (27 dependent load chains per atom otherwise).
        constexpr int NS = DIM == 3 ? 27 : 9;
        Cell nxt = lookup(-1, -1, DIM == 3 ? -1 : 0);
#pragma unroll 1
        for (int si = 0; si < NS; ++si) {
          const Cell cur = nxt;
          if (si + 1 < NS) {
            const int t = si + 1;
            if (DIM == 3) nxt = lookup(t / 9 - 1, (t / 3) % 3 - 1, t % 3 - 1);
            else nxt = lookup(t / 3 - 1, t % 3 - 1, 0);
          }
          scan_cell(cur);
        }
      } else {
        for (int s0 = -w; s0 <= w; ++s0)
        for (int s1 = -w; s1 <= w; ++s1)
        for (int s2 = -wz; s2 <= wz; ++s2) scan_cell(lookup(s0, s1, s2));
      }
      if (ORDERED && !COUNT) {
        // entries whose atom id is below the row's id (OrderedSparse keeps those,
        // partition.py:1021-1022): counted from the finished row, off the hot loop
        const int kk_end = k < kmax ? k : kmax;
#pragma unroll 4
        for (int kk = 0; kk < kk_end; ++kk)
          kl += (__ldg(&P.perm[__ldcg(out + (size_t)kk * P.n_pad)]) < hid);
      }
      P.cnt[slot] = k;
      P.cnt_lower[slot] = kl;
      my_k = k;
      my_tot = ORDERED ? kl : k;
    }
    publish_counts(P.state, my_k, my_tot);
  }
}

// all-pairs candidates (partition.py:904-909): one warp per atom, candidates in
// id order, warp-ballot compaction.
template <typename T, int DIM>
__device__ void ph_build_all_pairs(const NbrP<T, DIM>& P) {
  using V4 = typename Vec4<T>::type;
  const unsigned lane = threadIdx.x & 31;
  const int warp = gtid() >> 5;
  const int nwarps = gthreads() >> 5;
  long long wmax = 0, wtotal = 0;
  for (int i = warp; i < P.n; i += nwarps) {
    if (i >= P.n_rows) {                        // ghost: no row (identity order: slot == id)
      if (lane == 0) { P.cnt[i] = 0; P.cnt_lower[i] = 0; }
      continue;
    }
    const V4 hv = P.pos_sorted[i];
    const T hp[3] = {hv.x, hv.y, hv.z};
    int k = 0, kl = 0;
    for (int j0 = 0; j0 < P.n; j0 += 32) {
      const int j = j0 + lane;
      bool keep = false;
      if (j < P.n) {
        const V4 cv = P.pos_sorted[j];
        const T cp[3] = {cv.x, cv.y, cv.z};
        if (P.sp.general) {                       // identity order: slot == atom id
          T a[DIM], b[DIM];
#pragma unroll
          for (int k = 0; k < DIM; ++k) { a[k] = P.position[(size_t)i * DIM + k]; b[k] = P.position[(size_t)j * DIM + k]; }
          keep = P.rev_only ? dist2_exact<T, DIM>(P.sp, b, a) < P.cutoff_sq : dist2_exact<T, DIM>(P.sp, a, b) < P.cutoff_sq;
        } else {
          keep = candidate_test<T, DIM>(P, hp, cp);
        }
        if (P.mask_self && j == i) keep = false;
      }
      unsigned b = __ballot_sync(0xffffffffu, keep);
      if (keep && !P.count_only) {
        int pos = k + __popc(b & ((1u << lane) - 1u));
        if (pos < P.m_int) P.nl[(size_t)pos * P.n_pad + i] = j;
      }
      k += __popc(b);
      kl += __popc(__ballot_sync(0xffffffffu, keep && j < i));
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

// ---- phases: export to the public formats ------------------------------------------------

// per-atom (user order) number of public sparse entries
template <typename T, int DIM>
__device__ void ph_sparse_counts(const NbrP<T, DIM>& P) {
  for (int a = gtid(); a < P.n; a += gthreads()) {
    int t = P.inv_perm[a];
    P.tmp_ids[a] = P.format == JMD_ORDERED_SPARSE ? P.cnt_lower[t] : min(P.cnt[t], P.m_int);
  }
}

// Dense: idx[a, k] (partition.py:960-980, 1105); Sparse: idx[0]=receivers,
// idx[1]=senders ordered by sender then candidate order (partition.py:1010-1032).
// The internal list is transposed ([k][slot]).  Each WARP moves 32 slots x 32 k
// through its own padded shared-memory tile: 32 independent coalesced row loads
// in flight, then 32 independent perm gathers, then coalesced writes along k.
// No block barriers.
template <typename T, int DIM>
__device__ void ph_export(const NbrP<T, DIM>& P, Smem& sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int (*tile)[33] = sm.tile[w];
  const long long cap = P.max_occupancy;
  const bool dense = P.format == JMD_DENSE;
  const bool ordered = P.format == JMD_ORDERED_SPARSE;
  const int nwarps = gthreads() >> 5;
  const int ntiles = (P.n + 31) / 32;
  for (int tI = gtid() >> 5; tI < ntiles; tI += nwarps) {
    const int t0 = tI * 32;
    // lane r keeps the state of slot t0 + r
    const int slot = t0 + lane;
    int a_l = -1, c_l = 0, kk_l = 0;
    long long off_l = 0;
    if (slot < P.n) {
      a_l = P.perm[slot];
      c_l = min(P.cnt[slot], P.m_int);
      if (!dense) off_l = P.offsets[a_l];
    }
    int cmax = c_l;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
    const long long kend = dense ? cap : (long long)cmax;
    // sparse formats: entries that are not exported (past the row, or id >= row id
    // for OrderedSparse) are marked -1 when the tile is loaded; the write phase then
    // needs one running 32-bit output position per slot (64-bit only past 2^32 entries)
    const bool narrow = !dense && cap < (1ll << 32);
    unsigned pos_l = (unsigned)off_l;
    for (int k0 = 0; k0 < kend; k0 += 32) {
      const int rows = min(32, cmax - k0);       // rows of this tile that hold data
      __syncwarp();
      if (slot < P.n) {
        // row entries (coalesced along slots) -> atom ids (perm gather), 8 independent
        // load->gather chains in flight per lane; the tile holds ATOM IDS
        const int* src = P.nl + (size_t)k0 * P.n_pad + slot;
        const int drop = dense ? P.n : -1;
        const int id_limit = ordered ? a_l : 0x7fffffff;
#pragma unroll 1
        for (int r0 = 0; r0 < rows; r0 += JMD_EXPORT_BATCH) {
          int jv[JMD_EXPORT_BATCH];
#pragma unroll
          for (int u = 0; u < JMD_EXPORT_BATCH; ++u)     // entries past this slot's own row are garbage
            jv[u] = (k0 + r0 + u < c_l) ? __ldcs(src + (size_t)(r0 + u) * P.n_pad) : -1;
#pragma unroll
          for (int u = 0; u < JMD_EXPORT_BATCH; ++u) {
            int v = jv[u] >= 0 ? __ldg(&P.perm[jv[u]]) : drop;
            if (!dense && v >= id_limit) v = -1;
            jv[u] = v;
          }
#pragma unroll
          for (int u = 0; u < JMD_EXPORT_BATCH; ++u) tile[r0 + u][lane] = jv[u];
        }
      }
      __syncwarp();
      const int k = k0 + lane;
      if (dense) {
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
          const int a = __shfl_sync(0xffffffffu, a_l, r);
          if (a < 0) break;                            // slots beyond n (uniform)
          const int c = __shfl_sync(0xffffffffu, c_l, r);
          const int v = k < c ? tile[lane][r] : P.n;
          if (k < cap) P.idx[(size_t)a * cap + k] = v;
        }
      } else if (narrow) {
        const unsigned ucap = (unsigned)cap;
        const unsigned lt = (1u << lane) - 1u;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
          const int a = __shfl_sync(0xffffffffu, a_l, r);
          if (a < 0) break;
          const int v = (lane < rows) ? tile[lane][r] : -1;
          const unsigned b = __ballot_sync(0xffffffffu, v >= 0);
          const unsigned base = __shfl_sync(0xffffffffu, pos_l, r);
          const unsigned pos = base + __popc(b & lt);
          if (v >= 0 && pos < ucap) { P.idx[pos] = v; P.idx[cap + pos] = a; }
          if (lane == r) pos_l += __popc(b);
        }
      } else {
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
          const int a = __shfl_sync(0xffffffffu, a_l, r);
          if (a < 0) break;
          const int v = (lane < rows) ? tile[lane][r] : -1;
          const bool keep = v >= 0;
          const unsigned b = __ballot_sync(0xffffffffu, keep);
          const long long off = __shfl_sync(0xffffffffu, off_l, r);
          const int kk = __shfl_sync(0xffffffffu, kk_l, r);
          if (keep) {
            const long long pos = off + kk + __popc(b & ((1u << lane) - 1u));
            if (pos < cap) { P.idx[pos] = v; P.idx[cap + pos] = a; }
          }
          if (lane == r) kk_l += __popc(b);
        }
      }
    }
  }
}

// pad the tail of the sparse arrays with N (partition.py:1024 `N * ones`)
template <typename T, int DIM>
__device__ void ph_sparse_pad(const NbrP<T, DIM>& P) {
  const long long cap = P.max_occupancy;
  const long long start = P.offsets[P.n];
  for (long long p = start + gtid(); p < cap; p += gthreads()) {
    P.idx[p] = P.n;
    P.idx[cap + p] = P.n;
  }
}

template <typename T, int DIM>
__device__ void ph_finalize(const NbrP<T, DIM>& P) {
  const long long total = (long long)P.n * DIM;
  for (long long i = gtid(); i < total; i += gthreads()) P.ref[i] = P.position[i];
  if (gtid() == 0) {
    unsigned e = *P.error;
    // partition.py:1066 CELL_LIST_OVERFLOW, :1110 NEIGHBOR_LIST_OVERFLOW
    if (P.use_cells && P.state[ST_MAX_CELL] > P.cell_capacity) e |= JMD_ERR_CELL_LIST_OVERFLOW;
    long long occ = P.format == JMD_DENSE ? P.state[ST_MAX_ROW] : P.state[ST_TOTAL];
    if (occ > P.max_occupancy) e |= JMD_ERR_NEIGHBOR_LIST_OVERFLOW;
    // internal rows are capacity-bounded too (documented in DESIGN.md)
    if (P.state[ST_MAX_ROW] > P.m_int) e |= JMD_ERR_NEIGHBOR_LIST_OVERFLOW;
    *P.error = (uint8_t)e;
    P.state[ST_BUILDS] += 1;
    P.state[ST_EXPORT] = (P.lazy_idx && !P.no_public_idx) ? 1 : 0;
  }
}

// the reference-order prefix sums are needed only when cells are stored in brick
// order AND the slot rotation applies (search grid == reference grid)
template <typename T, int DIM>
__device__ __forceinline__ bool ref_scan(const NbrP<T, DIM>& P) { return P.bs > 0 && P.rotate; }
template <typename T, int DIM>
__device__ __forceinline__ int* ref_sums(const NbrP<T, DIM>& P) {
  return P.scan_tmp + (P.n_cells / SCAN_TILE + 2);     // behind the storage-order tile sums
}

#include "jmd_nbr_cellscan.cuh"

// ---- one ordinary kernel per phase (allocate path / gated fallback) -----------------------
enum Phase { PH_ZERO, PH_HASH, PH_SCAN1, PH_SCAN2, PH_SCAN3, PH_SCATTER, PH_RANK, PH_INVPERM,
             PH_IDENTITY, PH_PACK, PH_BUILD_RESET, PH_BUILD, PH_SP_COUNTS, PH_SP_SCAN1,
             PH_SP_SCAN2, PH_SP_SCAN3, PH_EXPORT, PH_SP_PAD, PH_FINALIZE, PH_EXPORT_DONE };

template <typename T, int DIM, int PHASE>
__global__ void __launch_bounds__(NB, 3) k_phase(NbrP<T, DIM> P, int gated) {
  if (gate_closed(P.state, gated)) return;
  __shared__ Smem sm;
  long long* sp_sums = (long long*)P.scan_tmp;
  switch (PHASE) {
    case PH_ZERO: ph_zero(P); break;
    case PH_HASH: ph_hash(P); break;
    case PH_SCAN1:
      ph_scan_tiles<int, int>(P.cell_count, P.n_cells, P.scan_tmp, nullptr, sm);
      if (ref_scan(P)) ph_scan_tiles<int, int>(P.ref_count, P.n_ref_cells, ref_sums(P), nullptr, sm);
      ph_ref_max(P);
      break;
    case PH_SCAN2:
      ph_scan_top<int>(P.scan_tmp, P.n_cells, sm);
      if (ref_scan(P)) ph_scan_top<int>(ref_sums(P), P.n_ref_cells, sm);
      break;
    case PH_SCAN3:
      ph_scan_apply<int, int>(P.cell_count, P.n_cells, P.scan_tmp, P.cell_start, sm);
      if (ref_scan(P)) ph_scan_apply<int, int>(P.ref_count, P.n_ref_cells, ref_sums(P), P.ref_start, sm);
      break;
    case PH_SCATTER: ph_scatter(P); break;
    case PH_RANK: ph_rank_sort(P); break;
    case PH_INVPERM: ph_inv_perm(P); break;
    case PH_IDENTITY: ph_identity_sort(P); break;
    case PH_PACK: ph_pack(P); break;
    case PH_BUILD_RESET: ph_build_reset(P); break;
    case PH_BUILD: ph_plan(P); break;   // the scans are their own kernels: k_nbr_stencil_scan / k_nbr_all_pairs
    case PH_SP_COUNTS: ph_sparse_counts(P); break;
    case PH_SP_SCAN1: ph_scan_tiles<int, long long>(P.tmp_ids, P.n, sp_sums, nullptr, sm); break;
    case PH_SP_SCAN2: ph_scan_top<long long>(sp_sums, P.n, sm); break;
    case PH_SP_SCAN3: ph_scan_apply<int, long long>(P.tmp_ids, P.n, sp_sums, P.offsets, sm); break;
    case PH_EXPORT: ph_export(P, sm); break;
    case PH_SP_PAD: ph_sparse_pad(P); break;
    case PH_FINALIZE: ph_finalize(P); break;
    case PH_EXPORT_DONE: if (gtid() == 0) P.state[ST_EXPORT] = 0; break;
  }
}

inline int grid_for(long long work_items, int per_block, int cap_blocks) {
  long long g = (work_items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > cap_blocks) g = cap_blocks;
  return (int)g;
}

// The two heavy phases get their own kernel names so profiles are readable.
// FMT: 0 Dense (forward + reverse test), 1 Sparse, 2 OrderedSparse.  Each variant
// is its own kernel so that it gets its own register allocation.
template <typename T, int DIM, int FMT, bool PERIODIC, int W, bool FILTER>
__global__ void __launch_bounds__(NB, sizeof(T) == 8 ? 2 : JMD_SCAN_MIN_BLOCKS) k_nbr_stencil_scan(const __grid_constant__ NbrP<T, DIM> P, int gated) {
  if (gate_closed(P.state, gated)) return;
  constexpr int MODE = (FMT == 0 && PERIODIC) ? 1 : 0;
  // COUNT (OrderedSparse occupancy pass of allocate: no rows are written, so the
  // id-lower count has to be taken inside the candidate loop); STAGE: also emit
  // the 16-bit staging rows of the force kernel (reference stencil only)
  if (FMT == 2 && P.count_only)
    ph_build_cells<T, DIM, MODE, FMT == 2, PERIODIC, W, FILTER, true, false>(P);
  else if (W == 1 && P.staged && !P.count_only)
    ph_build_cells<T, DIM, MODE, FMT == 2, PERIODIC, W, FILTER, false, W == 1>(P);
  else
    ph_build_cells<T, DIM, MODE, FMT == 2, PERIODIC, W, FILTER, false, false>(P);
}

template <typename T, int DIM>
__global__ void __launch_bounds__(NB, 3) k_nbr_all_pairs(NbrP<T, DIM> P, int gated) {
  if (gate_closed(P.state, gated)) return;
  ph_build_all_pairs<T, DIM>(P);
}

template <typename T, int DIM, int FMT, bool PERIODIC>
void launch_scan_wf(const NbrP<T, DIM>& P, int gated, int grid, cudaStream_t stream) {
  if (PERIODIC && P.filter) {
    if (P.sw == 1) k_nbr_stencil_scan<T, DIM, FMT, PERIODIC, 1, PERIODIC><<<grid, NB, 0, stream>>>(P, gated);
    else k_nbr_stencil_scan<T, DIM, FMT, PERIODIC, 0, PERIODIC><<<grid, NB, 0, stream>>>(P, gated);
  } else {
    if (P.sw == 1) k_nbr_stencil_scan<T, DIM, FMT, PERIODIC, 1, false><<<grid, NB, 0, stream>>>(P, gated);
    else k_nbr_stencil_scan<T, DIM, FMT, PERIODIC, 0, false><<<grid, NB, 0, stream>>>(P, gated);
  }
}

template <typename T, int DIM>
void launch_scan(const NbrP<T, DIM>& P, int gated, cudaStream_t stream) {
  if (P.cellscan) {
    launch_cell_test<T, DIM>(P, gated, stream);
    return;
  }
  if (!P.use_cells) {
    k_nbr_all_pairs<T, DIM><<<grid_for((long long)P.n * 32, NB, JMD_SM_COUNT * 16), NB, 0, stream>>>(P, gated);
    return;
  }
  const int grid = grid_for(P.n, NB, 1 << 30);
  const int fmt = P.format == JMD_DENSE ? 0 : (P.format == JMD_SPARSE ? 1 : 2);
  if (P.sp.periodic) {
    if (fmt == 0) launch_scan_wf<T, DIM, 0, true>(P, gated, grid, stream);
    else if (fmt == 1) launch_scan_wf<T, DIM, 1, true>(P, gated, grid, stream);
    else launch_scan_wf<T, DIM, 2, true>(P, gated, grid, stream);
  } else {
    // free space: Dense has no reverse test (MODE 0), same rows as Sparse
    if (fmt == 2) launch_scan_wf<T, DIM, 2, false>(P, gated, grid, stream);
    else launch_scan_wf<T, DIM, 1, false>(P, gated, grid, stream);
  }
}

template <typename T, int DIM>
__global__ void __launch_bounds__(NB, JMD_EXPORT_MIN_BLOCKS) k_nbr_export(NbrP<T, DIM> P, int gated) {
  if (gate_closed(P.state, gated)) return;
  if (P.no_public_idx) return;
  __shared__ Smem sm;
  ph_export(P, sm);
}

// export + sparse tail padding + (fin) reference positions / error bits in one launch;
// the last block to finish does the scalar part of ph_finalize
template <typename T, int DIM>
__global__ void __launch_bounds__(NB, JMD_EXPORT_MIN_BLOCKS) k_nbr_export_fin(NbrP<T, DIM> P, int gated, int fin) {
  if (gate_closed(P.state, gated)) return;
  __shared__ Smem sm;
  ph_export(P, sm);
  if (P.format != JMD_DENSE) ph_sparse_pad(P);
  if (fin) {
    const long long total = (long long)P.n * DIM;
    for (long long i = gtid(); i < total; i += gthreads()) P.ref[i] = P.position[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long t = atomicAdd((unsigned long long*)&P.state[ST_CS_TICKET], 1ull);
    if (t == (unsigned long long)gridDim.x - 1ull) {
      P.state[ST_CS_TICKET] = 0;
      if (fin) {
        unsigned e = *P.error;
        if (P.use_cells && P.state[ST_MAX_CELL] > P.cell_capacity) e |= JMD_ERR_CELL_LIST_OVERFLOW;
        const long long occ = P.format == JMD_DENSE ? P.state[ST_MAX_ROW] : P.state[ST_TOTAL];
        if (occ > P.max_occupancy) e |= JMD_ERR_NEIGHBOR_LIST_OVERFLOW;
        if (P.state[ST_MAX_ROW] > P.m_int) e |= JMD_ERR_NEIGHBOR_LIST_OVERFLOW;
        *P.error = (uint8_t)e;
        P.state[ST_BUILDS] += 1;
      }
      P.state[ST_EXPORT] = 0;
    }
  }
}

#define LAUNCH(PH, grid) k_phase<T, DIM, PH><<<(grid), NB, 0, stream>>>(P, gated)

template <typename T, int DIM>
void launch_bin(const NbrP<T, DIM>& P, int gated, cudaStream_t stream) {
  const int G = JMD_SM_COUNT * 8;
  if (!P.use_cells) {
    LAUNCH(PH_IDENTITY, grid_for(P.n, NB, G));
    return;
  }
  const int tiles = grid_for(P.n_cells, SCAN_TILE, G);
  LAUNCH(PH_ZERO, grid_for(P.n_cells + 1, NB, G));
  LAUNCH(PH_HASH, grid_for(P.n, NB, G));
  LAUNCH(PH_SCAN1, tiles);
  LAUNCH(PH_SCAN2, 1);
  LAUNCH(PH_SCAN3, tiles);
  LAUNCH(PH_SCATTER, grid_for(P.n, NB, G));
  LAUNCH(PH_RANK, grid_for(P.n, NB, G * 2));
  LAUNCH(PH_INVPERM, grid_for(P.n, NB, G));
}

template <typename T, int DIM>
void launch_build(const NbrP<T, DIM>& P, int gated, cudaStream_t stream) {
  LAUNCH(PH_BUILD_RESET, 1);
  if (P.staged) LAUNCH(PH_BUILD, grid_for((P.n + JMD_STAGE_BLOCK - 1) / JMD_STAGE_BLOCK, NB, JMD_SM_COUNT * 8));
  launch_scan<T, DIM>(P, gated, stream);
}

template <typename T, int DIM>
void launch_export(const NbrP<T, DIM>& P, int gated, cudaStream_t stream) {
  const int G = JMD_SM_COUNT * 8;
  if (P.cs_lb && !(P.no_public_idx || (P.lazy_idx && gated == 1))) {
    // cell-scan lists: look-back offsets + one export kernel that also pads the sparse
    // tail, stores the reference positions and sets the error bits (no grid barrier)
    if (P.format != JMD_DENSE)
      k_nbr_offsets<T, DIM><<<(P.n + SCAN_TILE - 1) / SCAN_TILE + (P.n == 0 ? 1 : 0), NB, 0, stream>>>(P, gated);
    k_nbr_export_fin<T, DIM><<<grid_for((long long)P.n, NB, 1 << 30), NB, 0, stream>>>(P, gated, gated == 2 ? 0 : 1);
    return;
  }
  // lazy materialisation: an update() only stores the reference positions and the
  // error bits and marks idx stale; the export itself runs (gated == 2) when the
  // host reads NeighborList.idx
  if (P.no_public_idx || (P.lazy_idx && gated == 1)) {
    LAUNCH(PH_FINALIZE, grid_for((long long)P.n * DIM, NB, G));
    return;
  }
  if (P.format != JMD_DENSE) {
    const int tiles = grid_for(P.n, SCAN_TILE, G);
    LAUNCH(PH_SP_COUNTS, grid_for(P.n, NB, G));
    LAUNCH(PH_SP_SCAN1, tiles);
    LAUNCH(PH_SP_SCAN2, 1);
    LAUNCH(PH_SP_SCAN3, tiles);
  }
  k_nbr_export<T, DIM><<<grid_for((long long)P.n, NB, 1 << 30), NB, 0, stream>>>(P, gated);
  if (P.format != JMD_DENSE) LAUNCH(PH_SP_PAD, grid_for(P.max_occupancy / 4 + 1, NB, G));
  if (gated == 2) LAUNCH(PH_EXPORT_DONE, 1);
  else LAUNCH(PH_FINALIZE, grid_for((long long)P.n * DIM, NB, G));
}

// Skin predicate alone (gated mode): the last block latches the decision.
template <typename T, int DIM>
__global__ void __launch_bounds__(NB) k_skin(NbrP<T, DIM> P) {
  const bool moved = ph_skin(P);
  const int any = __syncthreads_or(moved ? 1 : 0);
  if (threadIdx.x == 0) {
    if (any) atomicOr((unsigned long long*)&P.state[ST_PENDING], 1ull);
    __threadfence();
    unsigned long long t = atomicAdd((unsigned long long*)&P.state[ST_TICKET], 1ull);
    if (t == gridDim.x - 1) {
      __threadfence();
      long long reb = (atomicExch((unsigned long long*)&P.state[ST_PENDING], 0ull) != 0ull) || P.always_rebuild;
      P.state[ST_REBUILD] = reb;
      P.state[ST_TICKET] = 0;
    }
  }
}

// NeighborList.update(), part A (cooperative, persistent): skin predicate; when
// it fires the same kernel bins and sorts the atoms with grid-wide barriers
// between the phases.  Part B is the stencil scan as an ordinary (gated) launch
// so it runs at one thread per atom and full occupancy; part C (k_update_c)
// exports.  On the common no-rebuild step A returns after the predicate and B,
// C are two empty launches.
template <typename T, int DIM>
__global__ void __launch_bounds__(NB, 4) k_update(NbrP<T, DIM> P) {
  __shared__ Smem sm;
  unsigned int* bar = reinterpret_cast<unsigned int*>(&P.state[ST_BARRIER]);
  unsigned int target = 0;
  bool rebuild;
  if (P.skin_pre) {
    // the drift kernel already evaluated the predicate for exactly these positions:
    // every block ORs the per-drift-block flags itself (a few KB out of L2), no
    // pass over the positions and no grid barrier on the common no-rebuild step
    const int nblk = (P.n_rows + 255) / 256;
    bool moved = false;
    for (int b = threadIdx.x; b < nblk; b += NB) moved = moved || (__ldcg(&P.skin_blk[b]) != 0);
    rebuild = __syncthreads_or(moved ? 1 : 0) != 0 || P.always_rebuild;
  } else {
    const bool moved = ph_skin(P);
    const int any = __syncthreads_or(moved ? 1 : 0);
    if (threadIdx.x == 0 && any) atomicOr((unsigned long long*)&P.state[ST_PENDING], 1ull);
    JMD_GRID_SYNC(P, bar, target);
    rebuild = __ldcg(&P.state[ST_PENDING]) != 0 || P.always_rebuild;
  }
  if (gtid() == 0) P.state[ST_REBUILD] = rebuild ? 1 : 0;
  if (!rebuild) { grid_exit(bar); return; }       // uniform over the whole grid
  if (P.use_cells) {
    ph_zero(P);
    JMD_GRID_SYNC(P, bar, target);
    if (gtid() == 0) P.state[ST_PENDING] = 0;    // everyone has read it
    ph_hash(P);
    JMD_GRID_SYNC(P, bar, target);
    ph_scan_tiles<int, int>(P.cell_count, P.n_cells, P.scan_tmp, nullptr, sm);
    if (ref_scan(P)) ph_scan_tiles<int, int>(P.ref_count, P.n_ref_cells, ref_sums(P), nullptr, sm);
    ph_ref_max(P);
    JMD_GRID_SYNC(P, bar, target);
    ph_scan_top<int>(P.scan_tmp, P.n_cells, sm);
    if (ref_scan(P)) ph_scan_top<int>(ref_sums(P), P.n_ref_cells, sm);
    JMD_GRID_SYNC(P, bar, target);
    ph_scan_apply<int, int>(P.cell_count, P.n_cells, P.scan_tmp, P.cell_start, sm);
    if (ref_scan(P)) ph_scan_apply<int, int>(P.ref_count, P.n_ref_cells, ref_sums(P), P.ref_start, sm);
    JMD_GRID_SYNC(P, bar, target);
    ph_scatter(P);
    JMD_GRID_SYNC(P, bar, target);
    ph_rank_sort(P);
    ph_build_reset(P);
    JMD_GRID_SYNC(P, bar, target);
    ph_inv_perm(P);
    ph_plan(P);
  } else {
    JMD_GRID_SYNC(P, bar, target);
    if (gtid() == 0) P.state[ST_PENDING] = 0;
    ph_identity_sort(P);
    ph_build_reset(P);
    ph_plan(P);                                   // all blocks "direct" (needs no perm)
  }
  grid_exit(bar);
}

// part C: sparse offsets (needs a scan, hence barriers) + export + error bits.
// Dense needs no barrier and is launched with one thread per atom instead.
template <typename T, int DIM>
__global__ void __launch_bounds__(NB, JMD_EXPORT_MIN_BLOCKS) k_update_c(NbrP<T, DIM> P) {
  if (P.state[ST_REBUILD] == 0) return;
  __shared__ Smem sm;
  if (P.no_public_idx || P.lazy_idx) { ph_finalize(P); return; }
  if (P.format != JMD_DENSE) {
    unsigned int* bar = reinterpret_cast<unsigned int*>(&P.state[ST_BARRIER]);
    unsigned int target = 0;
    long long* sp_sums = (long long*)P.scan_tmp;
    ph_sparse_counts(P);
    JMD_GRID_SYNC(P, bar, target);
    ph_scan_tiles<int, long long>(P.tmp_ids, P.n, sp_sums, nullptr, sm);
    JMD_GRID_SYNC(P, bar, target);
    ph_scan_top<long long>(sp_sums, P.n, sm);
    JMD_GRID_SYNC(P, bar, target);
    ph_scan_apply<int, long long>(P.tmp_ids, P.n, sp_sums, P.offsets, sm);
    JMD_GRID_SYNC(P, bar, target);
    ph_export(P, sm);
    ph_sparse_pad(P);
    ph_finalize(P);
    grid_exit(bar);
  } else {
    ph_export(P, sm);
    ph_finalize(P);
  }
}

// ---- host side -------------------------------------------------------------------------

template <typename T, int DIM>
int fill(NbrP<T, DIM>& P, const jmd_nbr_t* nb, const void* position) {
  if (!nb || nb->n < 0) return JMD_EINVAL;
  P.n = nb->n; P.format = nb->format; P.use_cells = nb->use_cells; P.mask_self = nb->mask_self;
  P.always_rebuild = nb->always_rebuild; P.cell_capacity = nb->cell_capacity;
  P.m_int = nb->m_int;
  // reference grid (flags) and the internal fine search grid
  P.n_ref_cells = nb->n_cells;
  P.n_cells = nb->n_fine_cells;       // storage cells (padded to whole bricks)
  P.bs = nb->use_cells ? nb->brick_shift : 0;
  if (P.bs < 0 || P.bs > 3) return JMD_EINVAL;
  P.rotate = 1;
  P.ref_start = nb->ref_start;
  P.lazy_idx = 0;
  P.skin_blk = nb->skin_blk;
  P.skin_pre = (nb->skin_pre && nb->skin_blk) ? 1 : 0;
  P.staged = (nb->staged && nb->blk_table && nb->nl16) ? 1 : 0;
  P.stage_cap = JMD_STAGE_BYTES / (int)sizeof(typename Vec4<T>::type);
  P.cs_bits = (unsigned*)nb->cs_bits;
  P.cs_lb = (unsigned long long*)nb->cs_lb;
  P.cs_chunks = nb->cs_chunks;
  P.cs_batches = nb->cs_batches;
  P.cellscan = 0;
  P.blk_table = nb->blk_table;
  P.nl16 = nb->nl16;
  P.sw = nb->stencil_w > 0 ? nb->stencil_w : 1;
  for (int k = 0; k < 3; ++k) {
    P.ref_cps[k] = nb->cps[k] > 0 ? nb->cps[k] : 1;
    P.cps[k] = nb->fine_cps[k] > 0 ? nb->fine_cps[k] : 1;
    P.nb[k] = (P.cps[k] + (1 << P.bs) - 1) >> P.bs;
    if (P.cps[k] != P.ref_cps[k]) P.rotate = 0;       // finer search grid: ascending-id cells
  }
  if (nb->use_cells) {
    long long storage = 1;
    for (int k = 0; k < DIM; ++k) storage *= (long long)P.nb[k] << P.bs;
    if (storage != P.n_cells) return JMD_EINVAL;
    if (P.bs > 0 && P.rotate && !P.ref_start) return JMD_EINVAL;
  }
  if (nb->use_cells) {
    for (int k = 0; k < DIM; ++k)
      if (P.cps[k] < 2 * P.sw + 1) return JMD_EINVAL;       // stencil cells would alias
  }
  // the stencil scan addresses nl with 32-bit element offsets
  if ((unsigned long long)nb->m_int * (unsigned long long)nb->n_pad >= (1ull << 32)) return JMD_EINVAL;
  // warp-per-cell scan: reference grid in the reference's storage order only
  if (nb->cell_scan && nb->use_cells && P.bs == 0 && P.sw == 1 && P.rotate && !P.staged && P.cs_lb &&
      P.cs_chunks > 0 && P.cs_batches > 0 && cs_test_smem(P.cs_chunks) <= CS_SMEM_MAX)
    P.cellscan = 1;
  P.count_only = 0;
  P.n_rows = (nb->n_rows > 0 && nb->n_rows < nb->n) ? nb->n_rows : nb->n;
  P.no_public_idx = nb->no_public_idx;
  P.n_pad = nb->n_pad; P.max_occupancy = nb->max_occupancy;
  for (int k = 0; k < DIM; ++k) {
    P.cell_size[k] = (T)nb->fine_cell_size[k];
    P.ref_cell_size[k] = (T)nb->cell_size[k];
  }
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
  // Pre-filter (see ph_build_cells): the contracted image-shifted d2 differs from
  // either exact orientation by less than `band` (delta above covers 10 half-ulps
  // of L per component; the exact path accumulates <= 4, the filter <= 2.4).
  P.f_lo = (T)(nb->cutoff_sq - 2.0 * band);
  P.f_hi = (T)(nb->cutoff_sq + 2.0 * band);
  P.filter = (nb->use_cells && periodic && !nb->no_filter) ? 1 : 0;
  if (nb->space.general) {
    // the pre-filter runs on the real-space copy (fractional * side: one more rounding of
    // size ulp(L)/2 per coordinate than the derivation above counts): double the band
    P.band = (T)(4.0 * band);
    P.f_lo = (T)(nb->cutoff_sq - 4.0 * band);
    P.f_hi = (T)(nb->cutoff_sq + 4.0 * band);
  }
  for (int k = 0; k < DIM; ++k)
    if (P.cps[k] < 2 * P.sw + 3) P.filter = 0;
  // full-matrix boxes: every candidate takes the exact metric (the pre-filter's image shift and
  // band are derived for axis-aligned cells)
  if (nb->space.general && nb->space.triclinic) P.filter = 0;
  P.cell_count = nb->cell_count; P.cell_start = nb->cell_start; P.cell_cursor = nb->cell_cursor;
  P.scan_tmp = nb->scan_tmp; P.hash = nb->hash; P.tmp_ids = nb->tmp_ids; P.perm = nb->perm;
  P.inv_perm = nb->inv_perm; P.ref_count = nb->ref_count;
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

// one co-resident wave of blocks of a persistent kernel on the current device
template <typename K>
int coop_grid(K kernel, int* cached /*[64]*/, int* grid) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev < 64 && cached[dev] > 0) { *grid = cached[dev]; return 0; }
  int per_sm = 0, sms = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, NB, 0);
  if (e != cudaSuccess) return (int)e;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return (int)e;
  if (per_sm < 1) return JMD_EINVAL;
  *grid = per_sm * sms;
  if (dev < 64) cached[dev] = *grid;
  return 0;
}

template <typename T, int DIM>
int launch_update(NbrP<T, DIM>& P, cudaStream_t stream) {
  static int cache_a[64] = {0}, cache_c[64] = {0};
  int grid = 0, rc;
  if ((rc = coop_grid(k_update<T, DIM>, cache_a, &grid))) return rc;
  k_update<T, DIM><<<grid, NB, 0, stream>>>(P);
  launch_scan<T, DIM>(P, 1, stream);
  if (P.cs_lb) {
    // offsets by a look-back scan + one export kernel: ordinary launches, no grid barrier
    launch_export<T, DIM>(P, 1, stream);
    JMD_LAUNCH_CHECK();
    return 0;
  }
  if (P.format == JMD_DENSE) {
    k_update_c<T, DIM><<<grid_for(P.n, NB, 1 << 30), NB, 0, stream>>>(P);
  } else {
    if ((rc = coop_grid(k_update_c<T, DIM>, cache_c, &grid))) return rc;
    k_update_c<T, DIM><<<grid, NB, 0, stream>>>(P);
  }
  JMD_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" {

int jmd_nbr_update(const jmd_nbr_t* nb, const void* position, void* stream) {
  return dispatch(nb, [&](auto t, auto d) -> int {
    using T = decltype(t);
    constexpr int DIM = decltype(d)::value;
    NbrP<T, DIM> P;
    int rc = fill(P, nb, position);
    if (rc) return rc;
    P.lazy_idx = nb->lazy_idx ? 1 : 0;
    return launch_update<T, DIM>(P, (cudaStream_t)stream);
  });
}

int jmd_nbr_skin_check(const jmd_nbr_t* nb, const void* position, void* stream) {
  return dispatch(nb, [&](auto t, auto d) -> int {
    using T = decltype(t);
    constexpr int DIM = decltype(d)::value;
    NbrP<T, DIM> P;
    int rc = fill(P, nb, position);
    if (rc) return rc;
    k_skin<T, DIM><<<grid_for(P.n, NB, JMD_SM_COUNT * 4), NB, 0, (cudaStream_t)stream>>>(P);
    JMD_LAUNCH_CHECK();
    return 0;
  });
}

int jmd_nbr_bin(const jmd_nbr_t* nb, const void* position, int gated, void* stream_) {
  return dispatch(nb, [&](auto t, auto d) -> int {
    using T = decltype(t);
    constexpr int DIM = decltype(d)::value;
    NbrP<T, DIM> P;
    int rc = fill(P, nb, position);
    if (rc) return rc;
    launch_bin<T, DIM>(P, gated, (cudaStream_t)stream_);
    JMD_LAUNCH_CHECK();
    return 0;
  });
}

int jmd_nbr_build(const jmd_nbr_t* nb, const void* position, int count_only, int gated, void* stream_) {
  return dispatch(nb, [&](auto t, auto d) -> int {
    using T = decltype(t);
    constexpr int DIM = decltype(d)::value;
    NbrP<T, DIM> P;
    int rc = fill(P, nb, position);
    if (rc) return rc;
    P.count_only = count_only;
    launch_build<T, DIM>(P, gated, (cudaStream_t)stream_);
    JMD_LAUNCH_CHECK();
    return 0;
  });
}

int jmd_nbr_export(const jmd_nbr_t* nb, const void* position, int gated, void* stream_) {
  return dispatch(nb, [&](auto t, auto d) -> int {
    using T = decltype(t);
    constexpr int DIM = decltype(d)::value;
    NbrP<T, DIM> P;
    int rc = fill(P, nb, position);
    if (rc) return rc;
    P.lazy_idx = (nb->lazy_idx && gated == 1) ? 1 : 0;     // gated == 1: part of an update()
    launch_export<T, DIM>(P, gated, (cudaStream_t)stream_);
    JMD_LAUNCH_CHECK();
    return 0;
  });
}

int jmd_nbr_pack(const jmd_nbr_t* nb, const void* position, void* stream_) {
  return dispatch(nb, [&](auto t, auto d) -> int {
    using T = decltype(t);
    constexpr int DIM = decltype(d)::value;
    NbrP<T, DIM> P;
    int rc = fill(P, nb, position);
    if (rc) return rc;
    k_phase<T, DIM, PH_PACK><<<grid_for(P.n, NB, JMD_SM_COUNT * 8), NB, 0, (cudaStream_t)stream_>>>(P, 0);
    JMD_LAUNCH_CHECK();
    return 0;
  });
}

int jmd_nbr_pack_range(const jmd_nbr_t* nb, const void* position, int first, int count, void* stream_) {
  if (count <= 0) return 0;
  return dispatch(nb, [&](auto t, auto d) -> int {
    using T = decltype(t);
    constexpr int DIM = decltype(d)::value;
    NbrP<T, DIM> P;
    int rc = fill(P, nb, position);
    if (rc) return rc;
    if (first < 0 || first + count > P.n) return JMD_EINVAL;
    k_pack_range<T, DIM><<<grid_for(count, NB, JMD_SM_COUNT * 8), NB, 0, (cudaStream_t)stream_>>>(P, first, count);
    JMD_LAUNCH_CHECK();
    return 0;
  });
}

int jmd_nbr_state_host(const jmd_nbr_t* nb, int64_t* out, void* stream_) {
  if (!nb || !out) return JMD_EINVAL;
  cudaError_t e = cudaMemcpyAsync(out, nb->state, sizeof(int64_t) * JMD_ST_COUNT, cudaMemcpyDeviceToHost,
                                  (cudaStream_t)stream_);
  if (e != cudaSuccess) return (int)e;
  e = cudaStreamSynchronize((cudaStream_t)stream_);
  return (int)e;
}

}  // extern "C"

// Warp-cooperative candidate scan ("cell scan") -- included by jmd_neighbor.cu inside
// its anonymous namespace, after NbrP / exact_keep / publish_counts.
//
// Replaces partition.py:911-1032 (candidate gather, distance mask, cumsum compaction)
// and the export to the public formats for lists whose cells hold enough atoms to keep
// a warp busy.  One WARP owns one home cell (all of the reference's 3^d stencil):
//
//   k_nbr_cell_test    lanes run over the CONCATENATED candidate stream of the stencil
//                      cells (reference order: stencil cells x-slowest, slot order
//                      inside a cell), 32 candidates per chunk, each candidate shifted
//                      once into the home cell's periodic image.  Home atoms are
//                      broadcast from shared memory; one FSETP + one VOTE.BALLOT per
//                      (chunk, home atom) yields the accept mask of that home atom over
//                      the chunk; after the chunk every lane (= home atom) appends the
//                      set bits of its mask to its row of the transposed internal list.
//                      The reference's exact arithmetic runs only for chunks that saw a
//                      candidate inside the rounding band around cutoff^2 or a "dirty"
//                      cell.
//   k_nbr_offsets      (sparse formats) exclusive scan of the per-atom entry counts in
//                      atom-id order: single pass, decoupled look-back.
// The export to the public idx is jmd_neighbor.cu's tile-transposing ph_export.
//
// Measured on B200 (LJ, N = 1M, profiles/r02_*): the test kernel needs 0.47 warp
// instructions per candidate test against 1.1 for the thread-per-atom scan.  Two
// variants that kept the accept masks in global memory and expanded them in a second
// kernel (rows staged in shared memory as 16-bit codes; a lock-step per-lane walk with
// coalesced row stores) were measured at 0.73 ms and 1.05 ms for the expansion alone and
// dropped.
//
// The candidate ORDER is the reference's, so `idx` stays element-exact.
#pragma once

constexpr int CS_WARPS = 4;                 // home cells per block
enum { ST_LB_TILE = 9, ST_CS_TICKET = 10 }; // state[] slots used by the offsets scan / expand finalize

template <typename T>
struct CsWarpSmem {
  int pre[32];        // exclusive prefix of the stencil cells' atom counts (entries >= NS: total)
  int cstart[32];     // first slot of stencil cell s
  int sdirty[32];     // stencil cell holds an irregular atom (exact test)
  T shift[3][32];     // image shift added to the candidates of stencil cell s
  typename Vec4<T>::type home[32];   // home atoms of the current batch: x y z id
  uint2 bits[32];     // .x accept mask of home atom h over the current chunk; .y the same
                      // restricted to candidates with a smaller atom id (OrderedSparse)
};

__device__ __forceinline__ int id_of(float w) { return __float_as_int(w); }
__device__ __forceinline__ int id_of(double w) { return (int)w; }            // exact below 2^53
__device__ __forceinline__ float id_as(float, int id) { return __int_as_float(id); }
__device__ __forceinline__ double id_as(double, int id) { return (double)id; }

// Stencil of `cell` in the reference's order; returns the candidate total.
template <typename T, int DIM>
__device__ __forceinline__ int cs_setup(const NbrP<T, DIM>& P, int cell, int lane, CsWarpSmem<T>& sm) {
  constexpr int NS = DIM == 3 ? 27 : 9;
  int cc[3];
  cell_coords(P, cell, cc);
  int cnt = 0, start = 0, dirty = 0;
  T sh[3] = {T(0), T(0), T(0)};
  if (lane < NS) {
    int s3[3];
    if (DIM == 3) { s3[0] = lane / 9 - 1; s3[1] = (lane / 3) % 3 - 1; s3[2] = lane % 3 - 1; }
    else { s3[0] = lane / 3 - 1; s3[1] = lane % 3 - 1; s3[2] = 0; }
    int sv[3] = {0, 0, 0};
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      int v = cc[d] + s3[d];
      // a stencil cell beyond a face is the periodic image of the cell at the other end:
      // its atoms are moved by -side / +side so that home - candidate is the minimum image
      if (v < 0) { v += P.cps[d]; sh[d] = -P.sp.side[d]; }
      else if (v >= P.cps[d]) { v -= P.cps[d]; sh[d] = P.sp.side[d]; }
      sv[d] = v;
    }
    const int h = cell_id(P, sv);
    start = __ldg(&P.cell_start[h]);
    cnt = __ldg(&P.cell_start[h + 1]) - start;
    dirty = __ldg(&P.cell_cursor[h]) & CELL_DIRTY;
  }
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  sm.pre[lane] = incl - cnt;             // lanes >= NS: the total
  sm.cstart[lane] = start;
  sm.sdirty[lane] = dirty;
#pragma unroll
  for (int d = 0; d < 3; ++d) sm.shift[d][lane] = sh[d];
  return __shfl_sync(0xffffffffu, incl, 31);
}

// largest s with pre[s] <= t (pre non-decreasing, 32 entries)
__device__ __forceinline__ int cs_find(const int* pre, int t) {
  int s = 0;
#pragma unroll
  for (int step = 16; step > 0; step >>= 1)
    if (pre[s + step] <= t) s += step;
  return s;
}

template <typename T, int DIM, int MODE, bool ORDERED, bool PERIODIC, bool FILTER>
__global__ void __launch_bounds__(CS_WARPS * 32) k_nbr_cell_test(NbrP<T, DIM> P, int gated) {
  if (gate_closed(P.state, gated)) return;
  using V4 = typename Vec4<T>::type;
  constexpr int NS = DIM == 3 ? 27 : 9;
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ CsWarpSmem<T> smem[CS_WARPS];
  extern __shared__ unsigned cs_dyn[];       // per warp: masks [cs_chunks + 1][32] | ranks [cs_chunks][32]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  CsWarpSmem<T>& sm = smem[w];
  unsigned* const masks_s = cs_dyn + (size_t)w * (2 * P.cs_chunks + 1) * 32;
  int* const ranks_s = reinterpret_cast<int*>(masks_s + (size_t)(P.cs_chunks + 1) * 32);
  const int cell = blockIdx.x * CS_WARPS + w;
  // the offsets scan of this rebuild starts from cleared look-back words
  {
    const int tiles = (P.n + SCAN_TILE - 1) / SCAN_TILE + 1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < tiles; i += gridDim.x * blockDim.x) P.cs_lb[i] = 0ull;
    if (blockIdx.x == 0 && threadIdx.x == 0) P.state[ST_LB_TILE] = 0;
  }
  long long my_k = 0, my_tot = 0;
  const V4* __restrict__ const pos = P.pos_sorted;
  int hs = 0, he = 0;
  if (cell < P.n_cells) { hs = P.cell_start[cell]; he = P.cell_start[cell + 1]; }
  if (he > hs) {
    const bool home_dirty = FILTER ? (__ldg(&P.cell_cursor[cell]) & CELL_DIRTY) != 0 : false;
    const int total = cs_setup<T, DIM>(P, cell, lane, sm);
    __syncwarp();
    const int nchunks = min((total + 31) >> 5, P.cs_chunks);
    const T c2 = P.cutoff_sq;
    const T bw = P.f_hi - c2;               // half width of the rounding band (2 * band)
    T hh[3], qq[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) { hh[d] = d < DIM ? P.sp.half[d] : T(0); qq[d] = d < DIM ? P.sp.quarter[d] : T(0); }
    const int self_pre = sm.pre[NS / 2];    // the centre stencil cell is the home cell itself
    for (int hb = hs; hb < he; hb += 32) {  // home atoms, 32 at a time
      const int batch = (hb - hs) >> 5;
      const int nh = min(32, he - hb);
      const int slot = hb + lane;
      if (batch >= P.cs_batches) {          // more atoms than cell_capacity: CELL_LIST_OVERFLOW is flagged
        if (lane < nh) { P.cnt[slot] = 0; P.cnt_lower[slot] = 0; }
        continue;
      }
      int hid = 0x7fffffff;
      V4 hv = make_v4<T>(T(0), T(0), T(0), T(0));
      if (lane < nh) { hv = pos[slot]; hid = P.perm[slot]; }
      const bool has_row = hid < P.n_rows;
      __syncwarp();
      sm.home[lane] = make_v4<T>(hv.x, hv.y, hv.z, id_as(T(0), hid));
      __syncwarp();
      int k = 0, kl = 0;
      if (__any_sync(FULL, has_row)) {
        for (int q = 0; q < nchunks; ++q) {
          const int t = q * 32 + lane;
          const bool valid = t < total;
          const int s = cs_find(sm.pre, valid ? t : 0);
          const int rank = sm.cstart[s] + (t - sm.pre[s]);
          V4 cv = make_v4<T>(T(0), T(0), T(0), T(0));
          int cid = 0x7fffffff;
          bool lane_exact = false;
          // invalid lanes sit far away: never accepted, never inside the band
          T cx = T(1e18), cy = T(1e18), cz = T(0);
          if (valid) {
            cv = pos[rank];
            if (ORDERED) cid = __ldg(&P.perm[rank]);
            lane_exact = FILTER ? (home_dirty || sm.sdirty[s] != 0) : true;
            cx = cv.x + sm.shift[0][s];
            cy = cv.y + sm.shift[1][s];
            if (DIM == 3) cz = cv.z + sm.shift[2][s];
          }
          bool band = false;
          if (FILTER) {
            // common path: contracted arithmetic on the image-shifted candidate; one
            // compare and one ballot per (chunk, home atom)
#pragma unroll 4
            for (int h = 0; h < nh; ++h) {
              const V4 hp = sm.home[h];
              const T ax = hp.x - cx, ay = hp.y - cy;
              T a2 = ax * ax + ay * ay;
              if (DIM == 3) { const T az = hp.z - cz; a2 += az * az; }
              const bool keep = a2 < c2;
              band = band || (fabs(a2 - c2) <= bw);
              const unsigned b = __ballot_sync(FULL, keep);
              if (ORDERED) {
                const unsigned bl = __ballot_sync(FULL, keep && cid < id_of(hp.w));
                if (lane == 0) sm.bits[h] = make_uint2(b, bl);
              } else {
                if (lane == 0) sm.bits[h].x = b;
              }
            }
          }
          if (!FILTER || __any_sync(FULL, band || lane_exact)) {
            // rare: a candidate inside the rounding band or a dirty cell -- the
            // reference's exact op sequence decides those pairs
#pragma unroll 1
            for (int h = 0; h < nh; ++h) {
              const V4 hp = sm.home[h];
              const T ax = hp.x - cx, ay = hp.y - cy;
              T a2 = ax * ax + ay * ay;
              if (DIM == 3) { const T az = hp.z - cz; a2 += az * az; }
              bool keep = a2 < c2;
              if (valid && (lane_exact || fabs(a2 - c2) <= bw)) {
                const T hp3[3] = {hp.x, hp.y, hp.z};
                keep = P.sp.general ? general_keep<T, DIM, MODE>(P, id_of(hp.w), __ldg(&P.perm[rank]), c2)
                                    : exact_keep<T, DIM, MODE, PERIODIC>(P, hp3, cv, hh, qq, c2);
              }
              keep = keep && valid;
              const unsigned b = __ballot_sync(FULL, keep);
              if (ORDERED) {
                const unsigned bl = __ballot_sync(FULL, keep && cid < id_of(hp.w));
                if (lane == 0) sm.bits[h] = make_uint2(b, bl);
              } else {
                if (lane == 0) sm.bits[h].x = b;
              }
            }
          }
          __syncwarp();
          unsigned m = 0u, ml = 0u;
          if (lane < nh) {
            if (ORDERED) { const uint2 mm = sm.bits[lane]; m = mm.x; ml = mm.y; }
            else m = sm.bits[lane].x;
          }
          __syncwarp();
          if (P.mask_self) {                 // partition.py:953-958
            const int t_self = self_pre + (slot - hs);
            if ((t_self >> 5) == q) m &= ~(1u << (t_self & 31));
          }
          if (!has_row) { m = 0u; ml = 0u; }
          kl += __popc(ml);
          k += __popc(m);
          if (!P.count_only) {                 // parked for the row walk below
            masks_s[q * 32 + lane] = m;
            ranks_s[q * 32 + lane] = rank;
          }
        }
        if (!P.count_only) {
          // Row walk: every lane (= home atom) pops ONE accepted candidate of its own masks
          // per iteration, so iteration i yields entry i of every row at once and the store
          // into the transposed list nl[i][slot] is coalesced over the lanes.  A sentinel
          // word of ones behind the last chunk ends the search for the next non-empty mask.
          masks_s[nchunks * 32 + lane] = 0xffffffffu;
          __syncwarp();
          const int c_l = min(k, P.m_int);
          int cmax = c_l;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) cmax = max(cmax, __shfl_xor_sync(FULL, cmax, o));
          int wq = 0;
          unsigned wm = masks_s[lane];
          int* dst = P.nl + slot;
#pragma unroll 1
          for (int i = 0; i < cmax; ++i) {
            if (i < c_l) {
              while (wm == 0u) wm = masks_s[(++wq) * 32 + lane];
              const int j = __ffs(wm) - 1;
              wm &= wm - 1u;
              *dst = ranks_s[wq * 32 + j];
            }
            dst += P.n_pad;
          }
          __syncwarp();
        }
      }
      if (lane < nh) {
        P.cnt[slot] = k;
        P.cnt_lower[slot] = kl;
        my_k = k > my_k ? k : my_k;
        my_tot += ORDERED ? kl : k;
      }
    }
  }
  publish_counts(P.state, my_k, my_tot);
}

// ---- sparse offsets: exclusive scan of the per-atom entry counts (atom-id order) ----
// single pass, decoupled look-back; tiles are handed out by an atomic counter so a
// tile only ever waits for tiles that already run.
template <typename T, int DIM>
__global__ void __launch_bounds__(NB) k_nbr_offsets(NbrP<T, DIM> P, int gated) {
  if (gate_closed(P.state, gated)) return;
  __shared__ long long sscan[NWARP];
  __shared__ int s_tile;
  __shared__ long long s_prefix;
  constexpr unsigned long long F_AGG = 1ull << 62, F_PRE = 2ull << 62, MASK = (1ull << 62) - 1ull;
  if (threadIdx.x == 0) s_tile = (int)atomicAdd((unsigned long long*)&P.state[ST_LB_TILE], 1ull);
  __syncthreads();
  const int tile = s_tile;
  const long long base = (long long)tile * SCAN_TILE;
  const bool ordered = P.format == JMD_ORDERED_SPARSE;
  int v[8];
  long long s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const long long a = base + (long long)threadIdx.x * 8 + j;
    int x = 0;
    if (a < P.n) {
      const int t = P.inv_perm[a];
      x = ordered ? P.cnt_lower[t] : min(P.cnt[t], P.m_int);
    }
    v[j] = x;
    s += x;
  }
  long long tot;
  long long e = block_excl_scan<long long>(s, &tot, sscan);
  if (threadIdx.x == 0) {
    volatile unsigned long long* st = P.cs_lb;
    long long run = 0;
    if (tile > 0) {
      st[tile] = F_AGG | (unsigned long long)tot;
      __threadfence();
      int p = tile - 1;
      while (true) {
        unsigned long long x;
        do { x = st[p]; } while ((x >> 62) == 0ull);
        run += (long long)(x & MASK);
        if ((x >> 62) == 2ull) break;
        --p;
      }
    }
    __threadfence();
    st[tile] = F_PRE | (unsigned long long)(run + tot);
    s_prefix = run;
  }
  __syncthreads();
  e += s_prefix;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const long long a = base + (long long)threadIdx.x * 8 + j;
    if (a < P.n) P.offsets[a] = e;
    e += v[j];
    if (a == (long long)P.n - 1) P.offsets[P.n] = e;
  }
  if (P.n == 0 && tile == 0 && threadIdx.x == 0) P.offsets[0] = 0;
}

// dynamic shared memory of one test block: per warp the parked masks (+ sentinel) and slots
constexpr size_t CS_SMEM_MAX = 160 * 1024;
inline size_t cs_test_smem(int cs_chunks) { return (size_t)CS_WARPS * (2 * cs_chunks + 1) * 32 * sizeof(unsigned); }

template <typename T, int DIM, int FMT, bool PERIODIC>
void launch_cell_test_f(const NbrP<T, DIM>& P, int gated, cudaStream_t stream) {
  constexpr int MODE = (FMT == 0 && PERIODIC) ? 1 : 0;
  const int grid = (P.n_cells + CS_WARPS - 1) / CS_WARPS;
  const size_t bytes = cs_test_smem(P.cs_chunks);
  static int optin[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 64 || !optin[dev]) {      // static + dynamic shared memory can pass 48 KB
    cudaFuncSetAttribute(k_nbr_cell_test<T, DIM, MODE, FMT == 2, PERIODIC, PERIODIC>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CS_SMEM_MAX);
    cudaFuncSetAttribute(k_nbr_cell_test<T, DIM, MODE, FMT == 2, PERIODIC, false>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CS_SMEM_MAX);
    if (dev < 64) optin[dev] = 1;
  }
  if (PERIODIC && P.filter)
    k_nbr_cell_test<T, DIM, MODE, FMT == 2, PERIODIC, PERIODIC><<<grid, CS_WARPS * 32, bytes, stream>>>(P, gated);
  else
    k_nbr_cell_test<T, DIM, MODE, FMT == 2, PERIODIC, false><<<grid, CS_WARPS * 32, bytes, stream>>>(P, gated);
}

template <typename T, int DIM>
void launch_cell_test(const NbrP<T, DIM>& P, int gated, cudaStream_t stream) {
  const int fmt = P.format == JMD_DENSE ? 0 : (P.format == JMD_SPARSE ? 1 : 2);
  if (P.sp.periodic) {
    if (fmt == 0) launch_cell_test_f<T, DIM, 0, true>(P, gated, stream);
    else if (fmt == 1) launch_cell_test_f<T, DIM, 1, true>(P, gated, stream);
    else launch_cell_test_f<T, DIM, 2, true>(P, gated, stream);
  } else {
    if (fmt == 2) launch_cell_test_f<T, DIM, 2, false>(P, gated, stream);
    else launch_cell_test_f<T, DIM, 1, false>(P, gated, stream);
  }
}


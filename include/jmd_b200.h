/* jmd_b200.h -- C ABI of libjmd_b200.so: B200 (sm_100a) kernels for the JAX MD
 * short-range hot path (neighbour list -> pair / Stillinger-Weber forces ->
 * velocity-Verlet / Nose-Hoover / FIRE step).
 *
 * Boundary rules:
 *   - extern "C", plain pointers and sizes, no torch / XLA types.
 *   - every entry point only ENQUEUES work on `stream` (a cudaStream_t passed as
 *     void*); none of them synchronises or allocates, so they are legal inside
 *     XLA FFI handlers and CUDA-graph capture.  The only exceptions are the
 *     `jmd_*_host` queries used by the (non-jittable) `allocate` path, which
 *     the reference also runs eagerly with host syncs (partition.py:249,1094).
 *   - data conditions (capacity overflow) never fail a call: they set bits in
 *     the device-side error code exactly like the reference's
 *     PartitionErrorCode (partition.py:494-561).
 *   - return value: 0 on success, a cudaError_t (>0) from the launch, or
 *     JMD_EINVAL (-1) for an invalid descriptor.
 *
 * All device pointers are owned by the caller (XLA / torch).  Positions,
 * momenta and forces are row-major [n, dim] arrays of `dtype`.
 */
#ifndef JMD_B200_H_
#define JMD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JMD_EINVAL (-1)

enum { JMD_F32 = 0, JMD_F64 = 1 };
/* partition.py:641-657 NeighborListFormat */
enum { JMD_DENSE = 0, JMD_SPARSE = 1, JMD_ORDERED_SPARSE = 2 };
/* partition.py:494-520 PartitionErrorCode */
enum { JMD_ERR_NEIGHBOR_LIST_OVERFLOW = 1, JMD_ERR_CELL_LIST_OVERFLOW = 2,
       JMD_ERR_CELL_SIZE_TOO_SMALL = 4, JMD_ERR_MALFORMED_BOX = 8 };
enum { JMD_SPACE_FREE = 0, JMD_SPACE_PERIODIC = 1 };
enum { JMD_POT_LJ = 0, JMD_POT_SOFT_SPHERE = 1, JMD_POT_MORSE = 2 };
/* smap.py:697-846 parameter modes */
enum { JMD_PARAM_SCALAR = 0, JMD_PARAM_PER_ATOM = 1, JMD_PARAM_SPECIES = 2,
       JMD_PARAM_MATRIX = 3 };
#define JMD_DPARAM_MAX_SPECIES 8   /* jmd_pair_t.dparam_rows */

/* space.py:258-329: free() / periodic(side).  `side` and `half` hold the
 * values of `side` and `f32(0.5) * side` in the position dtype, widened to
 * double (exact).  replaces: space.py:213-224 periodic_displacement,
 * space.py:250-252 periodic_shift. */
typedef struct {
  int32_t dim;        /* 2 or 3 */
  int32_t kind;       /* JMD_SPACE_* */
  int32_t wrapped;    /* shift_fn wraps into [0, side) */
  int32_t general;    /* 0: space.periodic / free; 1: space.periodic_general (space.py:332-472);
                         with an orthorhombic box (scalar, vector or diagonal matrix) the
                         exact displacement is box * (mod(sa - sb + 1/2, 1) - 1/2) on the
                         unit cube; `side` is the box diagonal, `inv_box` = 1 / box in the
                         position dtype */
  double side[3];
  double half[3];
  int32_t fractional; /* general: positions are stored in the unit cube (fractional_coordinates=True);
                         the cell-sorted float4 copy and everything the force kernels see stay in
                         real space (fractional * side) */
  int32_t triclinic;  /* general: the box is a full matrix (off-diagonal elements).  `box_m` / `inv_box_m`
                         (row-major, real_i = sum_j box_m[i][j] * frac_j) replace side / inv_box in
                         the metric; `half` then holds per-axis bounds under which a raw real-space
                         difference is already the minimum image (sum_j |inv_box_m[i][j]| half[j] <= 1/2) */
  double inv_box[3];
  double box_m[9];
  double inv_box_m[9];
} jmd_space_t;

/* Slots of the device scalar block `jmd_nbr_t.state` (int64 each). */
enum {
  JMD_ST_REBUILD = 0,      /* 1 when the next gated rebuild must run */
  JMD_ST_MAX_CELL_OCC = 1, /* max atoms in one cell (last binning) */
  JMD_ST_MAX_ROW = 2,      /* max neighbours in one row (last build) */
  JMD_ST_TOTAL = 3,        /* total entries in requested sparse format */
  JMD_ST_BUILDS = 4,       /* number of rebuilds executed */
  JMD_ST_SCAN_TICKET = 5,
  JMD_ST_EXPORT_PENDING = 8, /* public idx is stale (lazy_idx): set by a rebuild, cleared by the export */
  JMD_ST_COUNT = 16
};

/* Neighbour-list workspace: the hidden part of partition.NeighborList
 * (partition.py:684-737).  Filled by the host; all buffers device-resident. */
typedef struct {
  /* static configuration */
  int32_t n;               /* atoms */
  int32_t dtype;           /* JMD_F32 / JMD_F64 */
  int32_t format;          /* requested public format */
  int32_t use_cells;       /* 1: cell list path, 0: all-pairs candidates */
  int32_t mask_self;
  int32_t always_rebuild;  /* dr_threshold == 0, partition.py:892 */
  int32_t cps[3];          /* cells per side (x, y, z) */
  int32_t n_cells;
  int32_t cell_capacity;   /* reference cell_list_capacity (slot rotation, flag) */
  int32_t m_int;           /* internal row capacity (rows of `nl`) */
  int32_t n_rows;          /* atoms [0, n_rows) own rows / forces / skin checks; the rest
                              (ghosts of a domain decomposition) are candidates only.
                              0 means n. */
  int32_t no_public_idx;   /* skip the export of `idx` (distributed driver) */
  int64_t n_pad;           /* row stride of `nl` (n rounded up to 32) */
  int64_t max_occupancy;   /* public capacity: per row (Dense) / total (Sparse) */
  double cell_size[3];     /* f32 cell size, partition.py:162-163 */
  double cutoff_sq;        /* (r_cutoff + dr_threshold)^2, partition.py:900 */
  double threshold_sq;     /* (dr_threshold / 2)^2, partition.py:901 */
  jmd_space_t space;       /* metric of the user's displacement function */
  /* device buffers */
  int32_t* cell_count;     /* [n_fine_cells + 1] */
  int32_t* cell_start;     /* [n_fine_cells + 1] exclusive scan of cell_count */
  int32_t* cell_cursor;    /* [n_fine_cells] */
  int32_t* scan_tmp;       /* [>= n_cells / 1024 + 2] */
  int32_t* hash;           /* [n] cell hash per atom (user order) */
  int32_t* tmp_ids;        /* [n] */
  int32_t* perm;           /* [n_pad] sorted slot -> atom id */
  int32_t* inv_perm;       /* [n] atom id -> sorted slot */
  void* pos_sorted;        /* [n_pad] float4 / double4: x y z species */
  int32_t* nl;             /* [m_int, n_pad] transposed rows, slot indices */
  int32_t* cnt;            /* [n_pad] entries per slot row */
  int32_t* cnt_lower;      /* [n_pad] entries with id < row id */
  int64_t* offsets;        /* [n + 1] sparse export offsets (user order) */
  void* reference_position;/* [n, dim] positions at last build */
  int32_t* idx;            /* public idx: [n, max_occ] or [2, max_occ] */
  uint8_t* error;          /* PartitionError.code */
  int64_t* state;          /* [JMD_ST_COUNT] */
  const int32_t* species;  /* [n] or NULL (copied into pos_sorted.w) */
  /* internal search grid: the reference cells split `fine` times per side.  The
   * reference grid above only feeds cell_list_capacity / CELL_LIST_OVERFLOW;
   * cell_count / cell_start / cell_cursor are sized for the fine grid. */
  int32_t fine_cps[3];     /* fine cells per side */
  int32_t n_fine_cells;    /* storage cells: prod_k ceil(fine_cps[k] / brick) * brick */
  int32_t stencil_w;       /* stencil half width in fine cells (1, 2 or 3) */
  int32_t no_filter;       /* 1: disable the stencil scan's contracted pre-filter (exact test on every candidate) */
  double fine_cell_size[3];
  int32_t* ref_count;      /* [n_cells] atoms per REFERENCE cell */
  /* storage order of the cells: bricks of (1 << brick_shift)^dim cells (0 = the
   * reference's x-fastest order).  n_fine_cells counts STORAGE cells, i.e. the
   * grid padded to whole bricks.  ref_start: [n_cells + 1] scratch (prefix sums
   * in reference order, for the slot rotation of partition.py:441). */
  int32_t brick_shift;
  int32_t staged;          /* 1: nl16 / blk_table are allocated and maintained by the build */
  int32_t* ref_start;
  /* shared-memory staging plan of the force kernel (see csrc/jmd_common.cuh):
   * nl16 [m_int, n_pad] uint16 = the rows of `nl` as indices into the staging
   * buffer of the 256-slot block that owns the row; blk_table [n_pad / 256 + 1, 256]. */
  uint16_t* nl16;
  int32_t* blk_table;
  /* Skin predicate fused into the drift (jmd_nve_kick_drift): skin_blk[b] != 0 iff an
   * atom of drift block b (256 atoms, user order) moved further than skin/2 from its
   * reference position.  skin_pre = 1 tells jmd_nbr_update that skin_blk describes
   * exactly the `position` it is given (the host checks tensor identity), so the
   * predicate pass over the positions is skipped. */
  int32_t* skin_blk;       /* [n_pad / 256 + 1] or NULL */
  int32_t skin_pre;
  /* 1: jmd_nbr_update leaves the public `idx` stale after a rebuild (state[EXPORT_PENDING])
   * and jmd_nbr_export(gated = 2) materialises it on demand (NeighborList.idx read). */
  int32_t lazy_idx;
  /* Warp-per-cell candidate scan (csrc/jmd_nbr_cellscan.cuh), used when cell_scan != 0
   * and the search grid is the reference grid: one warp owns a home cell and tests the
   * concatenated candidate stream of its 3^d stencil 32 candidates at a time (cs_batches =
   * ceil(cell_capacity / 32) groups of home atoms, cs_chunks = ceil(3^d * cell_capacity /
   * 32) chunks of candidates).  cs_lb: [n / 2048 + 2] look-back words of the sparse
   * offsets scan.  cs_bits is reserved (must be NULL). */
  int32_t cell_scan;
  int32_t cs_chunks;
  int32_t cs_batches;
  int32_t _pad2;
  uint32_t* cs_bits;
  uint64_t* cs_lb;
  /* Device-resident atom counts {n, n_rows} (or NULL): when set, jmd_nve_kick_drift and
   * the jmd_dd_comm_* kernels take the counts from here instead of the host fields, so a
   * captured CUDA graph of the step stays valid when a rebuild of the domain
   * decomposition changes the local atom count (grids are sized for `n`, which then is
   * the capacity).  jmd_pair_force keeps the host counts: a decomposed list marks slots
   * without a row by cnt = 0 / perm >= n_rows (see jax_md_b200/domain.py). */
  const int32_t* n_dev;
} jmd_nbr_t;

/* ---- neighbour list (replaces partition.py:349-471, 911-1154) ------------- */

/* NeighborList.update (partition.py:1119-1154) as ONE cooperative kernel:
 * skin predicate any_i |d(R_i, ref_i)|^2 > threshold_sq (or always_rebuild);
 * when it is false every block returns, otherwise the same kernel bins, scans
 * the stencil, exports idx in nb->format, ORs the error bits and stores the new
 * reference positions, with grid-wide barriers between phases.  This is the
 * lax.cond of partition.py:1146 without a host round trip.  Leaves
 * state[REBUILD] = decision. */
int jmd_nbr_update(const jmd_nbr_t* nb, const void* position, void* stream);

/* The skin predicate alone: state[REBUILD] = decision.  Used with the gated
 * bin/build/export below as a fallback to jmd_nbr_update. */
int jmd_nbr_skin_check(const jmd_nbr_t* nb, const void* position, void* stream);

/* Bin atoms into cells and sort them (partition.py:421-460).  gated != 0:
 * kernels exit early unless state[REBUILD] is set (the lax.cond of
 * partition.py:1146).  Leaves state[MAX_CELL_OCC]. */
int jmd_nbr_bin(const jmd_nbr_t* nb, const void* position, int gated,
                void* stream);

/* Candidate scan + compaction (partition.py:911-1032) into the internal rows.
 * count_only != 0 just measures state[MAX_ROW] / state[TOTAL] (allocate). */
int jmd_nbr_build(const jmd_nbr_t* nb, const void* position, int count_only,
                  int gated, void* stream);

/* Export internal rows to the public idx in nb->format, update the error
 * code (partition.py:1066,1110), store reference_position, clear REBUILD. */
/* gated: 0 run, 1 run if this update() rebuilt, 2 run if idx is stale (lazy_idx). */
int jmd_nbr_export(const jmd_nbr_t* nb, const void* position, int gated,
                   void* stream);

/* allocate-time host query: copies state[] to `out[JMD_ST_COUNT]` after
 * synchronising the stream (partition.py:249,1094 do the same via int()). */
int jmd_nbr_state_host(const jmd_nbr_t* nb, int64_t* out, void* stream);

/* Refresh pos_sorted from user-order positions (no rebuild). */
int jmd_nbr_pack(const jmd_nbr_t* nb, const void* position, void* stream);

/* Same for atoms [first, first + count) only (ghost positions after a halo
 * exchange). */
int jmd_nbr_pack_range(const jmd_nbr_t* nb, const void* position, int first,
                       int count, void* stream);

/* ---- pair potentials (replaces smap.py:922-979 + energy.py:125-371,534-580) */

typedef struct {
  int32_t kind;            /* JMD_POT_* */
  int32_t has_cutoff;      /* multiplicative_isotropic_cutoff applied */
  int32_t mode[3];         /* JMD_PARAM_* for sigma, epsilon, alpha */
  int32_t n_species;       /* table side for SPECIES / MATRIX modes */
  int32_t transposed;      /* table lookups as p[neighbour, row] (Sparse formats:
                              smap.py:716,792 index with idx[0]=receiver first) */
  int32_t dparam_rows;     /* SPECIES-mode gradients: 0 = `dparam` holds two [S, S] tables filled with
                              atomicAdd (summation order varies from run to run); 1 = `dparam` holds
                              two [n, S] arrays, row i = atom i's sums per NEIGHBOUR species, no
                              atomics: table[s_i][s_j] = sum of the rows of species s_i, folded by
                              the caller in a fixed order (needs S <= JMD_DPARAM_MAX_SPECIES; swap
                              the table's axes when `transposed`) */
  double scalar[3];        /* sigma, epsilon, alpha when SCALAR */
  const void* array[3];    /* device arrays (position dtype) otherwise */
  double r_onset, r_cutoff;
  double r_onset2, r_cutoff2; /* r ** f32(2) in the reference's dtype, energy.py:558-559 */
  double switch_denom;        /* (r_c2 - r_o2) ** 3 in that same dtype, energy.py:566 */
} jmd_pair_t;

/* Slots of the double reduction block written by the force kernels. */
enum {
  JMD_RED_ENERGY = 0, JMD_RED_KINETIC = 1, JMD_RED_VIRIAL = 2, /* 2..7: xx yy zz xy xz yz */
  JMD_RED_DSIGMA = 8, JMD_RED_DEPSILON = 9, JMD_RED_FF = 10, JMD_RED_PP = 11,
  JMD_RED_FP = 12, JMD_RED_COUNT = 16
};

/* Fused force (+energy/virial/parameter gradients) over the internal rows.
 *   force      [n, dim] out (user order)
 *   e_atom     [n] out or NULL: per-atom energy (reduce_axis=(1,), smap.py:958)
 *   red        double[JMD_RED_COUNT] out or NULL (energy, virial, dE/dparam)
 *   dparam     out or NULL: dE/dsigma then dE/depsilon (SPECIES mode: 2 * n_species^2
 *              doubles, or 2 * n * n_species with pp->dparam_rows; PER_ATOM: 2 * n doubles)
 *   partials   scratch double[>= jmd_red_scratch_doubles(n)]
 * want_energy != 0 also produces red[ENERGY, VIRIAL.., DSIGMA, DEPSILON]
 * (dparam tables must be zeroed by the caller).
 * If momentum != NULL the second half-kick of velocity Verlet is fused
 * (simulate.py:241: p += dt_2 * F) and red[KINETIC] = 0.5 sum p^2/m; dt_dev
 * (device scalar, position dtype) overrides dt_2 with f32(f32(*dt_dev)/2). */
int jmd_pair_force(const jmd_nbr_t* nb, const jmd_pair_t* pp, void* force,
                   void* e_atom, double* red, double* dparam, double* partials,
                   void* momentum, const void* mass, int mass_is_array,
                   double dt_2, const void* dt_dev, int want_energy,
                   void* stream);

int64_t jmd_red_scratch_doubles(int64_t n);

/* ---- Stillinger-Weber (replaces energy.py:842-893, 994-1012) -------------- */
typedef struct {
  double sigma, A, B, lam, gamma, epsilon, three_body_strength, cutoff;
} jmd_sw_t;

/* scratch: int32[jmd_sw_scratch_ints(nb)], 16-byte aligned: the in-range neighbours of every atom as
 * compact transposed records (count | displacement + slot | r, h(r), h'(r)). */
int64_t jmd_sw_scratch_ints(const jmd_nbr_t* nb);
int jmd_sw_force(const jmd_nbr_t* nb, const jmd_sw_t* sw, int32_t* scratch, void* force,
                 double* red, double* partials, void* momentum,
                 const void* mass, int mass_is_array, double dt_2,
                 const void* dt_dev, void* stream);

/* ---- integrators (replaces simulate.py:168-243, 444-517; minimize.py:184-224) */

/* First half of velocity Verlet: p = scale*p + dt_2*F;  R = shift(R, dt*p/m)
 * (simulate.py:238-239), out of place; also refreshes nb->pos_sorted when
 * nb != NULL.  `dt_dev`/`scale_dev` (device scalars of the position dtype)
 * override dt/scale when non-NULL (FIRE's traced dt, minimize.py:185; the NHC
 * momentum scale, simulate.py:471-473). */
int jmd_nve_kick_drift(const jmd_space_t* sp, int dtype, int n,
                       const jmd_nbr_t* nb, const void* r_in, const void* p_in,
                       const void* f_in, const void* mass, int mass_is_array,
                       double dt, const void* dt_dev, const void* scale_dev,
                       void* r_out, void* p_out, void* stream);

/* Generic second half-kick for force functions we do not own:
 * p += dt_2 * F, red[KINETIC] = 0.5 sum p^2/m, plus FF / PP / FP sums. */
int jmd_kick_reduce(int dtype, int n, int dim, void* momentum, const void* force,
                    const void* mass, int mass_is_array, double dt_2,
                    const void* dt_dev, double* red, double* partials,
                    void* stream);

/* p *= *scale_dev (NHC second half step). */
int jmd_scale_momentum(int dtype, int64_t count, void* momentum,
                       const void* scale_dev, void* stream);

/* Nose-Hoover chain half step on device scalars (simulate.py:444-507).
 * chain = [xi[cl] | p_xi[cl] | Q[cl] | KE] of the position dtype.  ke_red:
 * NULL -> use the chain's own KE, else the double KE slot written by the
 * force+kick kernel (simulate.py:662).  Writes the product of the sub-step
 * momentum scales to scale_out (position dtype).  chain_out: where the updated
 * chain goes (NULL or == chain: in place); the reference returns a new chain. */
int jmd_nhc_half_step(int dtype, int chain_length, int chain_steps, int sy_steps,
                      double dt, double tau, int64_t dof, const void* kT_dev,
                      const void* chain, void* chain_out, const double* ke_red,
                      void* scale_out, void* stream);

/* FIRE momentum mixing + schedule (minimize.py:190-224) from red[FF,PP,FP].
 * fire_in/out = [dt, alpha] (position dtype), n_pos int32; out of place so
 * every thread mixes with the OLD alpha. */
int jmd_fire_mix(int dtype, int64_t count, void* momentum, const void* force,
                 const double* red, const void* fire_in, void* fire_out,
                 const int32_t* npos_in, int32_t* npos_out, double dt_max,
                 double n_min, double f_inc, double f_dec, double alpha_start,
                 double f_alpha, void* stream);

/* ---- slab domain decomposition helpers (no reference equivalent) */

/* `info` words of the rebuild-time pipeline (device int32[JMD_DD_INFO_COUNT]); the host
 * writes JMD_DD_N_OWN / clears JMD_DD_ERROR and reads everything back once per rebuild. */
enum {
  JMD_DD_N_OWN = 0,   /* owned atoms (in: before migration, out: after) */
  JMD_DD_FACE_L = 1,  /* face atoms sent to the left / right neighbour every step */
  JMD_DD_FACE_R = 2,
  JMD_DD_FROM_L = 3,  /* ghost atoms received from the left / right neighbour */
  JMD_DD_FROM_R = 4,
  JMD_DD_ERROR = 5,   /* sticky: JMD_DD_ELIST | JMD_DD_ECAP */
  JMD_DD_MIG_L = 6,   /* atoms that left to the left / right in the last migration */
  JMD_DD_MIG_R = 7,
  JMD_DD_IN_L = 8,    /* atoms that arrived from the left / right */
  JMD_DD_IN_R = 9,
  JMD_DD_N_LOC = 10,  /* owned + ghost atoms; {N_LOC, N_ROWS} is the jmd_nbr_t.n_dev pair */
  JMD_DD_N_ROWS = 11, /* == N_OWN */
  JMD_DD_INFO_COUNT = 16
};
#define JMD_DD_ELIST 1 /* a selection list exceeded its capacity */
#define JMD_DD_ECAP 2  /* owned + ghost atoms exceed the array capacity */
#define JMD_DD_ETIMEOUT 4 /* a peer's halo / flag did not arrive (spin-wait timed out) */

/* For atoms i < n (i < min(n, *n_dev) when n_dev != NULL, so the count can stay on the
 * device): d = (pos[i, axis] - lo) wrapped into [-L/2, L/2).  Appends i to
 * list_a when d < thr_a and to list_b when d >= thr_b (unordered; counters[0..1]
 * must be zeroed by the caller; entries beyond `cap` are counted, not stored).
 * Migration uses thr_a = 0, thr_b = width; ghost selection thr_a = ghost_width,
 * thr_b = width - ghost_width. */
int jmd_dd_select(int dtype, int dim, int n, const int32_t* n_dev, const void* position,
                  int axis, double lo, double L, double thr_a, double thr_b,
                  int32_t* list_a, int32_t* list_b, int32_t* counters, int cap,
                  void* stream);

/* dst[i, :] = src[idx[i], :] for rows of `ncomp` elements (per-step halo pack). */
int jmd_dd_pack(int dtype, int ncomp, int n_idx, const int32_t* idx,
                const void* src, void* dst, void* stream);

/* Fixed-capacity messages that carry their own count, so a rebuild needs no count
 * exchange and no host round trip before the data exchange:
 *   pack_migrate: pay_x[i, :] = R | P | F of atom list_x[i] (3*dim values),
 *                 gid_x[0] = min(counters[x], cap_mig), gid_x[1 + i] = gid[list_x[i]]
 *                 (x = a: to the left neighbour, b: to the right).
 *   compact:      removes the leavers (lists sorted ascending) from the owned arrays by
 *                 moving staying tail atoms into the holes, appends the arrivals
 *                 (in_l then in_r, counts in gid_in_x[0]) and updates info[JMD_DD_N_OWN].
 *                 scratch: int32[6 * cap_mig].
 *   pack_counted: dst[0, 0] = min(*count, cap); dst[1 + i, :] = src[idx[i], :].
 *   place:        ghost rows R[n_own + k] = [recv_l rows | recv_r rows] (counts in
 *                 recv_x[0, 0]); fills info FACE_x / FROM_x / ERROR. */
int jmd_dd_pack_migrate(int dtype, int dim, int cap_mig, const int32_t* list_a,
                        const int32_t* list_b, const int32_t* counters, const void* R,
                        const void* P, const void* F, const int64_t* gid, void* pay_a,
                        void* pay_b, int64_t* gid_a, int64_t* gid_b, void* stream);
int jmd_dd_compact(int dtype, int dim, int cap_own, int cap_mig, const int32_t* list_a,
                   const int32_t* list_b, const int32_t* counters, const void* in_l,
                   const int64_t* gid_in_l, const void* in_r, const int64_t* gid_in_r,
                   void* R, void* P, void* F, int64_t* gid, int32_t* scratch,
                   int32_t* info, void* stream);
int jmd_dd_pack_counted(int dtype, int ncomp, int cap, const int32_t* idx,
                        const int32_t* count, const void* src, void* dst, void* stream);
int jmd_dd_place(int dtype, int dim, int cap_total, int cap_list, const int32_t* counters,
                 const void* recv_l, const void* recv_r, void* R, int32_t* info,
                 void* stream);

/* Ordered variant of jmd_dd_select: the lists come out sorted by atom index (single
 * pass, decoupled look-back), so ghost order and hence summation order are reproducible.
 * scratch: uint64[n / 2048 + 2] look-back words + 1 tile counter, zeroed by the call. */
int jmd_dd_select_ordered(int dtype, int dim, int n, const int32_t* n_dev, const void* position,
                          int axis, double lo, double L, double thr_a, double thr_b,
                          int32_t* list_a, int32_t* list_b, int32_t* counters, int cap,
                          uint64_t* scratch, void* stream);

/* ---- per-step exchange over NVLink peer memory (CUDA IPC; one process per GPU) --------
 * Every rank owns one "shared block" (jmd_p2p_alloc) that its ring neighbours map
 * (jmd_p2p_open): two parities x two sides of landing rows for ghost positions, a signal
 * word per side and one rebuild-flag word per rank.  A step is then two kernels with no
 * host involvement and no NCCL call, so it can live inside a CUDA graph:
 *   jmd_dd_comm_push  gathers the face atoms and STORES them straight into the
 *                     neighbours' landing rows through the peer mapping, then releases
 *                     the neighbours' signal words; writes this rank's rebuild flag (OR
 *                     of the drift's skin flags) into every rank's flag word.
 *   jmd_dd_comm_wait  waits for both signals, copies the landed rows behind the owned
 *                     atoms and into the cell-sorted float4 array; block 0 ORs all
 *                     ranks' flags and publishes (epoch << 1 | any) to a mapped host word,
 *                     which the host polls while the force kernel still runs.
 * Spin-waits give up after ~8 s and set JMD_DD_ETIMEOUT in info[JMD_DD_ERROR]. */
int jmd_p2p_alloc(int64_t bytes, void** ptr, uint8_t* handle64);
int jmd_p2p_open(const uint8_t* handle64, void** ptr);
int jmd_p2p_close(void* ptr);
int jmd_p2p_free(void* ptr);
int jmd_host_flag_alloc(uint64_t** host_ptr, uint64_t** dev_ptr);   /* mapped pinned word */
int jmd_host_flag_free(uint64_t* host_ptr);

typedef struct {
  int32_t dtype, dim;
  int32_t rank, world;
  int32_t cap_list;            /* rows per landing buffer */
  int32_t always_rebuild;
  const int32_t* face_l;       /* [cap_list] owned-atom indices sent to the left neighbour */
  const int32_t* face_r;
  const int32_t* face_counts;  /* [2] device-resident list lengths */
  int32_t* info;               /* JMD_DD_* words */
  uint64_t* epoch;             /* [1] device step counter */
  unsigned int* ticket;        /* [2] zero-initialised scratch */
  const int32_t* skin_blk;     /* drift's skin flags (jmd_nbr_t.skin_blk) */
  /* this rank's shared block: */
  void* land;                  /* [2][2][cap_list][dim]: parity, side (0 = from left) */
  uint64_t* signal;            /* [2] epoch of the last complete push from the left / right */
  uint64_t* flags;             /* [2 parities][world] (epoch << 1 | flag), slot r written by rank r */
  /* peers: */
  void* peer_land_l;           /* left neighbour's `land` (we are ITS right side) */
  void* peer_land_r;
  uint64_t* peer_signal_l;     /* left neighbour's `signal` */
  uint64_t* peer_signal_r;
  uint64_t* const* peer_flags; /* device array [world] of every rank's `flags` */
  uint64_t* host_flag;         /* device address of the mapped host word */
} jmd_dd_t;

int jmd_dd_comm_push(const jmd_dd_t* dd, const void* R, void* stream);
int jmd_dd_comm_wait(const jmd_dd_t* dd, const jmd_nbr_t* nb, void* R, void* stream);

const char* jmd_version(void);

#ifdef __cplusplus
}
#endif
#endif  /* JMD_B200_H_ */

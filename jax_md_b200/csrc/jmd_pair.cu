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
#define JMD_PAIR_STAGED 0
#include "jmd_pair_impl.cuh"

// staged variants live in jmd_pair_staged.cu (compiled in parallel)
template <typename T, int DIM>
int jmd_launch_pair_staged(const jmd_nbr_t* nb, const jmd_pair_t* pp, void* force, void* e_atom, double* red,
                           double* dparam, double* partials, void* momentum, const void* mass,
                           int mass_is_array, double dt_2, const void* dt_dev, bool want_e, cudaStream_t s);

// full-matrix periodic_general boxes live in jmd_pair_tric.cu
template <typename T, int DIM>
int jmd_launch_pair_tric(const jmd_nbr_t* nb, const jmd_pair_t* pp, void* force, void* e_atom, double* red,
                         double* dparam, double* partials, void* momentum, const void* mass,
                         int mass_is_array, double dt_2, const void* dt_dev, bool want_e, cudaStream_t s);

namespace {
// The staged kernel serves scalar and per-species parameters (the species id
// travels in pos.w); per-atom / matrix parameters need the neighbour's atom id.
template <typename T, int DIM>
int launch_pair_any(const jmd_nbr_t* nb, const jmd_pair_t* pp, void* force, void* e_atom, double* red,
                    double* dparam, double* partials, void* momentum, const void* mass, int mass_is_array,
                    double dt_2, const void* dt_dev, bool want_e, cudaStream_t s) {
  if (nb->space.general && nb->space.triclinic)
    return jmd_launch_pair_tric<T, DIM>(nb, pp, force, e_atom, red, dparam, partials, momentum, mass,
                                        mass_is_array, dt_2, dt_dev, want_e, s);
  bool stage = nb->staged && nb->nl16 && nb->blk_table && nb->use_cells;
  for (int k = 0; k < 3; ++k)
    if (pp->mode[k] == JMD_PARAM_PER_ATOM || pp->mode[k] == JMD_PARAM_MATRIX) stage = false;
  if (stage)
    return jmd_launch_pair_staged<T, DIM>(nb, pp, force, e_atom, red, dparam, partials, momentum, mass,
                                          mass_is_array, dt_2, dt_dev, want_e, s);
  return launch_pair_impl<T, DIM>(nb, pp, force, e_atom, red, dparam, partials, momentum, mass,
                                  mass_is_array, dt_2, dt_dev, want_e, s);
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
  if (nb->dtype == JMD_F32 && dim == 3) return launch_pair_any<float, 3>(JMD_ARGS);
  if (nb->dtype == JMD_F32 && dim == 2) return launch_pair_any<float, 2>(JMD_ARGS);
  if (nb->dtype == JMD_F64 && dim == 3) return launch_pair_any<double, 3>(JMD_ARGS);
  if (nb->dtype == JMD_F64 && dim == 2) return launch_pair_any<double, 2>(JMD_ARGS);
#undef JMD_ARGS
  return JMD_EINVAL;
}

}  // extern "C"

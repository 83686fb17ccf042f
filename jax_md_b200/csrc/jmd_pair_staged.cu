// Pair-force kernels with neighbour positions staged through shared memory
// (see jmd_pair_impl.cuh / jmd_common.cuh).  Separate unit so it compiles in
// parallel with jmd_pair.cu.
#define JMD_PAIR_STAGED 1
#include "jmd_pair_impl.cuh"

template <typename T, int DIM>
int jmd_launch_pair_staged(const jmd_nbr_t* nb, const jmd_pair_t* pp, void* force, void* e_atom, double* red,
                           double* dparam, double* partials, void* momentum, const void* mass,
                           int mass_is_array, double dt_2, const void* dt_dev, bool want_e, cudaStream_t s) {
  return launch_pair_staged_impl<T, DIM>(nb, pp, force, e_atom, red, dparam, partials, momentum, mass,
                                         mass_is_array, dt_2, dt_dev, want_e, s);
}

#define JMD_INST(T, DIM)                                                                              \
  template int jmd_launch_pair_staged<T, DIM>(const jmd_nbr_t*, const jmd_pair_t*, void*, void*,     \
                                              double*, double*, double*, void*, const void*, int,    \
                                              double, const void*, bool, cudaStream_t);
JMD_INST(float, 2)
JMD_INST(float, 3)
JMD_INST(double, 2)
JMD_INST(double, 3)
#undef JMD_INST

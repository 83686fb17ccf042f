// Pair-force kernels for space.periodic_general with a full-matrix (triclinic) box: the same
// kernel body as jmd_pair.cu with the minimum image taken through the box matrix
// (jmd_common.cuh: Space::wrap_tric).  A separate unit so the orthorhombic kernels keep their
// exact instruction stream (see DESIGN.md 3 on how sensitive that loop is) and so it compiles
// in parallel.
#define JMD_PAIR_STAGED 0
#define JMD_PAIR_TRIC 1
#define JMD_PAIR_MIN_BLOCKS 8      // 64 registers: room for the two 3x3 matrices without spills
#include "jmd_pair_impl.cuh"

template <typename T, int DIM>
int jmd_launch_pair_tric(const jmd_nbr_t* nb, const jmd_pair_t* pp, void* force, void* e_atom, double* red,
                         double* dparam, double* partials, void* momentum, const void* mass,
                         int mass_is_array, double dt_2, const void* dt_dev, bool want_e, cudaStream_t s) {
  return launch_pair_tric_impl<T, DIM>(nb, pp, force, e_atom, red, dparam, partials, momentum, mass,
                                       mass_is_array, dt_2, dt_dev, want_e, s);
}

#define JMD_INST(T, DIM)                                                                            \
  template int jmd_launch_pair_tric<T, DIM>(const jmd_nbr_t*, const jmd_pair_t*, void*, void*,     \
                                            double*, double*, double*, void*, const void*, int,    \
                                            double, const void*, bool, cudaStream_t);
JMD_INST(float, 2)
JMD_INST(float, 3)
JMD_INST(double, 2)
JMD_INST(double, 3)
#undef JMD_INST

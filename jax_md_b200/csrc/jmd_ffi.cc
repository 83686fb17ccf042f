// jmd_ffi.cc -- XLA FFI handlers over the C ABI of libjmd_b200.so (include/jmd_b200.h).
//
// This is the plug-in surface a JAX MD maintainer binds (jax >= 0.5):
//     jax.ffi.register_ffi_target("jmd_nbr_update", jax.ffi.pycapsule(lib.jmd_ffi_nbr_update),
//                                 platform="CUDA")
// and calls through jax.ffi.ffi_call from the `neighbor_list_fn=` / `pair_neighbor_list_fn=`
// plug-in points of jax_md/energy.py:211-212, 312-313, 413-414, 976 (see
// jax_md_b200/_jax_binding.py for the Python half and INTEGRATION.md for the walk-through).
//
// It is compiled only where the XLA FFI headers exist (jax.ffi.include_dir(); build.py
// looks for $JMD_XLA_FFI_INCLUDE or an importable jax) into libjmd_b200_ffi.so, which
// links against libjmd_b200.so.  This image has neither jax nor the headers, so here the
// file is checked for syntax against a stand-in of the API (tests/ffi_stub) and for
// coverage of every launcher (tests/test_ffi_binding.py).
//
// Conventions (one handler per C entry point, same name with the jmd_ffi_ prefix):
//   * POD descriptors (jmd_nbr_t, jmd_pair_t, jmd_sw_t, jmd_space_t, jmd_dd_t) arrive as
//     byte-string attributes holding the struct exactly as the host code fills it (ctypes
//     mirror in jax_md_b200/_lib.py) with every pointer field zero; the handler patches the
//     pointers from XLA buffers.
//   * The neighbour-list workspace is a fixed, ordered bundle of buffers
//     (JMD_NBR_WORKSPACE below).  Each of them is an operand AND a result, aliased one to
//     one by `input_output_aliases`, so XLA updates them in place; the handler reads the
//     RESULT pointers.  The bundle is the hidden, dynamic half of the NeighborList pytree.
//   * Handlers only enqueue on XLA's stream, never synchronise, never allocate.  Data
//     conditions (capacity overflow) are bits in the `error` buffer, not ffi::Errors.
#include <cstdint>
#include <cstring>
#include <string_view>

#include <cuda_runtime.h>

#include "xla/ffi/api/ffi.h"

#include "jmd_b200.h"

namespace ffi = xla::ffi;

namespace {

// Order of the workspace bundle == order of the pointer fields of jmd_nbr_t.
#define JMD_NBR_WORKSPACE(X)                                                                  \
  X(cell_count) X(cell_start) X(cell_cursor) X(scan_tmp) X(hash) X(tmp_ids) X(perm)           \
  X(inv_perm) X(pos_sorted) X(nl) X(cnt) X(cnt_lower) X(offsets) X(reference_position)        \
  X(idx) X(error) X(state) X(ref_count) X(ref_start) X(skin_blk) X(cs_lb)
constexpr int kNbrWorkspace = 21;

ffi::Error Invalid(const char* what) { return ffi::Error(ffi::ErrorCode::kInvalidArgument, what); }

ffi::Error Status(int rc, const char* what) {
  if (rc == 0) return ffi::Error::Success();
  if (rc == JMD_EINVAL) return Invalid(what);
  return ffi::Error(ffi::ErrorCode::kInternal, cudaGetErrorString((cudaError_t)rc));
}

template <typename S>
bool Unpack(std::string_view bytes, S* out) {
  if (bytes.size() != sizeof(S)) return false;
  std::memcpy(out, bytes.data(), sizeof(S));
  return true;
}

// jmd_nbr_t from its byte image + the aliased workspace results.
ffi::Error FillNbr(std::string_view desc, ffi::RemainingRets& ws, size_t first, jmd_nbr_t* nb) {
  if (!Unpack(desc, nb)) return Invalid("nbr descriptor has the wrong size");
  if (ws.size() < first + kNbrWorkspace) return Invalid("neighbour workspace bundle is incomplete");
  size_t i = first;
#define X(field)                                                           \
  {                                                                        \
    auto b = ws.get<ffi::AnyBuffer>(i++);                                  \
    if (!b.has_value()) return Invalid("workspace buffer " #field);        \
    nb->field = reinterpret_cast<decltype(nb->field)>((*b)->untyped_data()); \
  }
  JMD_NBR_WORKSPACE(X)
#undef X
  return ffi::Error::Success();
}

void* Opt(ffi::RemainingArgs& args, size_t i) {      // optional operand: absent or empty -> NULL
  if (i >= args.size()) return nullptr;
  auto b = args.get<ffi::AnyBuffer>(i);
  if (!b.has_value() || b->element_count() == 0) return nullptr;
  return b->untyped_data();
}

void* OptRet(ffi::RemainingRets& rets, size_t i) {
  if (i >= rets.size()) return nullptr;
  auto b = rets.get<ffi::AnyBuffer>(i);
  if (!b.has_value() || (*b)->element_count() == 0) return nullptr;
  return (*b)->untyped_data();
}

// ---- neighbour list (partition.py:1037-1154) --------------------------------------------
// operands: position [N, dim], species [N] (optional), then the workspace bundle
// results:  the workspace bundle (aliased)
enum NbrOp { kUpdate, kSkinCheck, kBin, kBuild, kExport, kPack, kPackRange };

template <NbrOp OP>
ffi::Error NbrImpl(cudaStream_t stream, std::string_view desc, int32_t a0, int32_t a1, ffi::AnyBuffer position,
                   ffi::RemainingArgs args, ffi::RemainingRets ws) {
  jmd_nbr_t nb;
  if (auto e = FillNbr(desc, ws, 0, &nb); e.failure()) return e;
  nb.species = static_cast<const int32_t*>(Opt(args, 0));
  const void* pos = position.untyped_data();
  switch (OP) {
    case kUpdate: return Status(jmd_nbr_update(&nb, pos, stream), "jmd_nbr_update");
    case kSkinCheck: return Status(jmd_nbr_skin_check(&nb, pos, stream), "jmd_nbr_skin_check");
    case kBin: return Status(jmd_nbr_bin(&nb, pos, /*gated=*/a0, stream), "jmd_nbr_bin");
    case kBuild: return Status(jmd_nbr_build(&nb, pos, /*count_only=*/a0, /*gated=*/a1, stream), "jmd_nbr_build");
    case kExport: return Status(jmd_nbr_export(&nb, pos, /*gated=*/a0, stream), "jmd_nbr_export");
    case kPack: return Status(jmd_nbr_pack(&nb, pos, stream), "jmd_nbr_pack");
    case kPackRange: return Status(jmd_nbr_pack_range(&nb, pos, /*first=*/a0, /*count=*/a1, stream), "jmd_nbr_pack_range");
  }
  return Invalid("unknown neighbour op");
}

#define JMD_NBR_BINDING                                        \
  ffi::Ffi::Bind()                                             \
      .Ctx<ffi::PlatformStream<cudaStream_t>>()                \
      .Attr<std::string_view>("desc")                          \
      .Attr<int32_t>("a0")                                     \
      .Attr<int32_t>("a1")                                     \
      .Arg<ffi::AnyBuffer>() /* position */                    \
      .RemainingArgs()       /* [species], workspace... */     \
      .RemainingRets()       /* workspace (aliased) */

// ---- fused pair force (smap.py:922-979 + quantity.py:58-60 + simulate.py:241) ---------------
// operands: mass, dt_dev (optional), sigma/epsilon/alpha arrays (optional, PairT modes),
//           species (optional), momentum (aliased to result 4 when kicked), workspace...
// results:  force [N, dim], e_atom [N] (or empty), red f64[16], dparam (or empty), partials,
//           momentum (or empty), workspace...
ffi::Error PairForceImpl(cudaStream_t stream, std::string_view nbr_desc, std::string_view pair_desc,
                         int32_t mass_is_array, double dt_2, int32_t want_energy, int32_t kick,
                         ffi::RemainingArgs args, ffi::RemainingRets rets) {
  jmd_nbr_t nb;
  if (auto e = FillNbr(nbr_desc, rets, 6, &nb); e.failure()) return e;
  jmd_pair_t pp;
  if (!Unpack(pair_desc, &pp)) return Invalid("pair descriptor has the wrong size");
  const void* mass = Opt(args, 0);
  const void* dt_dev = Opt(args, 1);
  for (int k = 0; k < 3; ++k) pp.array[k] = Opt(args, 2 + k);
  nb.species = static_cast<const int32_t*>(Opt(args, 5));
  void* momentum = kick ? OptRet(rets, 5) : nullptr;
  return Status(jmd_pair_force(&nb, &pp, OptRet(rets, 0), OptRet(rets, 1), static_cast<double*>(OptRet(rets, 2)),
                               static_cast<double*>(OptRet(rets, 3)), static_cast<double*>(OptRet(rets, 4)),
                               momentum, mass, mass_is_array, dt_2, dt_dev, want_energy, stream),
                "jmd_pair_force");
}

// ---- Stillinger-Weber (energy.py:842-893, 994-1012) --------------------------------------------
// operands: mass, dt_dev (optional), momentum (aliased), workspace...
// results:  force, red, partials, scratch i32[jmd_sw_scratch_ints(nb)], momentum (or empty), workspace...
ffi::Error SwForceImpl(cudaStream_t stream, std::string_view nbr_desc, std::string_view sw_desc,
                       int32_t mass_is_array, double dt_2, int32_t kick, ffi::RemainingArgs args,
                       ffi::RemainingRets rets) {
  jmd_nbr_t nb;
  if (auto e = FillNbr(nbr_desc, rets, 5, &nb); e.failure()) return e;
  jmd_sw_t sw;
  if (!Unpack(sw_desc, &sw)) return Invalid("sw descriptor has the wrong size");
  void* momentum = kick ? OptRet(rets, 4) : nullptr;
  return Status(jmd_sw_force(&nb, &sw, static_cast<int32_t*>(OptRet(rets, 3)), OptRet(rets, 0),
                             static_cast<double*>(OptRet(rets, 1)), static_cast<double*>(OptRet(rets, 2)), momentum,
                             Opt(args, 0), mass_is_array, dt_2, Opt(args, 1), stream),
                "jmd_sw_force");
}

// ---- integrators (simulate.py:168-243, 444-517; minimize.py:184-224) ------------------------
// kick_drift operands: r_in, p_in, f_in, mass, dt_dev (optional), scale_dev (optional), workspace...
//            results:  r_out, p_out, workspace...   (nb_desc empty: no list to refresh)
ffi::Error KickDriftImpl(cudaStream_t stream, std::string_view space_desc, std::string_view nbr_desc, int32_t dtype,
                         int32_t n, int32_t mass_is_array, double dt, ffi::AnyBuffer r_in, ffi::AnyBuffer p_in,
                         ffi::AnyBuffer f_in, ffi::AnyBuffer mass, ffi::RemainingArgs args,
                         ffi::RemainingRets rets) {
  jmd_space_t sp;
  if (!Unpack(space_desc, &sp)) return Invalid("space descriptor has the wrong size");
  jmd_nbr_t nb;
  const bool with_list = !nbr_desc.empty();
  if (with_list)
    if (auto e = FillNbr(nbr_desc, rets, 2, &nb); e.failure()) return e;
  return Status(jmd_nve_kick_drift(&sp, dtype, n, with_list ? &nb : nullptr, r_in.untyped_data(), p_in.untyped_data(),
                                   f_in.untyped_data(), mass.untyped_data(), mass_is_array, dt, Opt(args, 0),
                                   Opt(args, 1), OptRet(rets, 0), OptRet(rets, 1), stream),
                "jmd_nve_kick_drift");
}

// kick_reduce: operands force, mass, dt_dev (optional), momentum (aliased); results momentum, red, partials
ffi::Error KickReduceImpl(cudaStream_t stream, int32_t dtype, int32_t n, int32_t dim, int32_t mass_is_array,
                          double dt_2, ffi::AnyBuffer force, ffi::AnyBuffer mass, ffi::RemainingArgs args,
                          ffi::Result<ffi::AnyBuffer> momentum, ffi::Result<ffi::AnyBuffer> red,
                          ffi::Result<ffi::AnyBuffer> partials) {
  return Status(jmd_kick_reduce(dtype, n, dim, momentum->untyped_data(), force.untyped_data(), mass.untyped_data(),
                                mass_is_array, dt_2, Opt(args, 0), static_cast<double*>(red->untyped_data()),
                                static_cast<double*>(partials->untyped_data()), stream),
                "jmd_kick_reduce");
}

ffi::Error ScaleMomentumImpl(cudaStream_t stream, int32_t dtype, ffi::AnyBuffer scale,
                             ffi::Result<ffi::AnyBuffer> momentum /* aliased to operand 1 */) {
  return Status(jmd_scale_momentum(dtype, (int64_t)momentum->element_count(), momentum->untyped_data(),
                                   scale.untyped_data(), stream),
                "jmd_scale_momentum");
}

// nhc_half_step: operands kT, chain_in, ke_red f64[16] (optional); results chain_out, scale
ffi::Error NhcHalfStepImpl(cudaStream_t stream, int32_t dtype, int32_t chain_length, int32_t chain_steps,
                           int32_t sy_steps, double dt, double tau, int64_t dof, ffi::AnyBuffer kT,
                           ffi::AnyBuffer chain_in, ffi::RemainingArgs args, ffi::Result<ffi::AnyBuffer> chain_out,
                           ffi::Result<ffi::AnyBuffer> scale) {
  const double* red = static_cast<const double*>(Opt(args, 0));
  return Status(jmd_nhc_half_step(dtype, chain_length, chain_steps, sy_steps, dt, tau, dof, kT.untyped_data(),
                                  chain_in.untyped_data(), chain_out->untyped_data(),
                                  red ? red + JMD_RED_KINETIC : nullptr, scale->untyped_data(), stream),
                "jmd_nhc_half_step");
}

// fire_mix: operands force, red, fire_in, npos_in, momentum (aliased); results momentum, fire_out, npos_out
ffi::Error FireMixImpl(cudaStream_t stream, int32_t dtype, double dt_max, double n_min, double f_inc, double f_dec,
                       double alpha_start, double f_alpha, ffi::AnyBuffer force, ffi::AnyBuffer red,
                       ffi::AnyBuffer fire_in, ffi::AnyBuffer npos_in, ffi::Result<ffi::AnyBuffer> momentum,
                       ffi::Result<ffi::AnyBuffer> fire_out, ffi::Result<ffi::AnyBuffer> npos_out) {
  return Status(jmd_fire_mix(dtype, (int64_t)momentum->element_count(), momentum->untyped_data(), force.untyped_data(),
                             static_cast<const double*>(red.untyped_data()), fire_in.untyped_data(),
                             fire_out->untyped_data(), static_cast<const int32_t*>(npos_in.untyped_data()),
                             static_cast<int32_t*>(npos_out->untyped_data()), dt_max, n_min, f_inc, f_dec,
                             alpha_start, f_alpha, stream),
                "jmd_fire_mix");
}

// ---- slab domain decomposition helpers (per-device shard_map bodies) ------------------------
// select / select_ordered: operand position; results list_a, list_b, counters[, scratch]
template <bool ORDERED>
ffi::Error DdSelectImpl(cudaStream_t stream, int32_t dtype, int32_t dim, int32_t n, int32_t axis, double lo, double L,
                        double thr_a, double thr_b, int32_t cap, ffi::AnyBuffer position, ffi::RemainingArgs args,
                        ffi::RemainingRets rets) {
  const int32_t* n_dev = static_cast<const int32_t*>(Opt(args, 0));
  int32_t* la = static_cast<int32_t*>(OptRet(rets, 0));
  int32_t* lb = static_cast<int32_t*>(OptRet(rets, 1));
  int32_t* counters = static_cast<int32_t*>(OptRet(rets, 2));
  if (ORDERED)
    return Status(jmd_dd_select_ordered(dtype, dim, n, n_dev, position.untyped_data(), axis, lo, L, thr_a, thr_b, la,
                                        lb, counters, cap, static_cast<uint64_t*>(OptRet(rets, 3)), stream),
                  "jmd_dd_select_ordered");
  return Status(jmd_dd_select(dtype, dim, n, n_dev, position.untyped_data(), axis, lo, L, thr_a, thr_b, la, lb,
                              counters, cap, stream),
                "jmd_dd_select");
}

ffi::Error DdPackImpl(cudaStream_t stream, int32_t dtype, int32_t ncomp, ffi::AnyBuffer idx, ffi::AnyBuffer src,
                      ffi::Result<ffi::AnyBuffer> dst) {
  return Status(jmd_dd_pack(dtype, ncomp, (int)idx.element_count(), static_cast<const int32_t*>(idx.untyped_data()),
                            src.untyped_data(), dst->untyped_data(), stream),
                "jmd_dd_pack");
}

ffi::Error DdPackCountedImpl(cudaStream_t stream, int32_t dtype, int32_t ncomp, int32_t cap, ffi::AnyBuffer idx,
                             ffi::AnyBuffer count, ffi::AnyBuffer src, ffi::Result<ffi::AnyBuffer> dst) {
  return Status(jmd_dd_pack_counted(dtype, ncomp, cap, static_cast<const int32_t*>(idx.untyped_data()),
                                    static_cast<const int32_t*>(count.untyped_data()), src.untyped_data(),
                                    dst->untyped_data(), stream),
                "jmd_dd_pack_counted");
}

ffi::Error DdPackMigrateImpl(cudaStream_t stream, int32_t dtype, int32_t dim, int32_t cap_mig, ffi::AnyBuffer list_a,
                             ffi::AnyBuffer list_b, ffi::AnyBuffer counters, ffi::AnyBuffer R, ffi::AnyBuffer P,
                             ffi::AnyBuffer F, ffi::AnyBuffer gid, ffi::Result<ffi::AnyBuffer> pay_a,
                             ffi::Result<ffi::AnyBuffer> pay_b, ffi::Result<ffi::AnyBuffer> gid_a,
                             ffi::Result<ffi::AnyBuffer> gid_b) {
  return Status(jmd_dd_pack_migrate(dtype, dim, cap_mig, static_cast<const int32_t*>(list_a.untyped_data()),
                                    static_cast<const int32_t*>(list_b.untyped_data()),
                                    static_cast<const int32_t*>(counters.untyped_data()), R.untyped_data(),
                                    P.untyped_data(), F.untyped_data(), static_cast<const int64_t*>(gid.untyped_data()),
                                    pay_a->untyped_data(), pay_b->untyped_data(),
                                    static_cast<int64_t*>(gid_a->untyped_data()),
                                    static_cast<int64_t*>(gid_b->untyped_data()), stream),
                "jmd_dd_pack_migrate");
}

// compact: results R, P, F, gid, info are aliased to the operands of the same name
ffi::Error DdCompactImpl(cudaStream_t stream, int32_t dtype, int32_t dim, int32_t cap_own, int32_t cap_mig,
                         ffi::AnyBuffer list_a, ffi::AnyBuffer list_b, ffi::AnyBuffer counters, ffi::AnyBuffer in_l,
                         ffi::AnyBuffer gid_in_l, ffi::AnyBuffer in_r, ffi::AnyBuffer gid_in_r,
                         ffi::RemainingArgs aliased, ffi::Result<ffi::AnyBuffer> R, ffi::Result<ffi::AnyBuffer> P,
                         ffi::Result<ffi::AnyBuffer> F, ffi::Result<ffi::AnyBuffer> gid,
                         ffi::Result<ffi::AnyBuffer> info, ffi::Result<ffi::AnyBuffer> scratch) {
  (void)aliased;
  return Status(jmd_dd_compact(dtype, dim, cap_own, cap_mig, static_cast<const int32_t*>(list_a.untyped_data()),
                               static_cast<const int32_t*>(list_b.untyped_data()),
                               static_cast<const int32_t*>(counters.untyped_data()), in_l.untyped_data(),
                               static_cast<const int64_t*>(gid_in_l.untyped_data()), in_r.untyped_data(),
                               static_cast<const int64_t*>(gid_in_r.untyped_data()), R->untyped_data(),
                               P->untyped_data(), F->untyped_data(), static_cast<int64_t*>(gid->untyped_data()),
                               static_cast<int32_t*>(scratch->untyped_data()),
                               static_cast<int32_t*>(info->untyped_data()), stream),
                "jmd_dd_compact");
}

ffi::Error DdPlaceImpl(cudaStream_t stream, int32_t dtype, int32_t dim, int32_t cap_total, int32_t cap_list,
                       ffi::AnyBuffer counters, ffi::AnyBuffer recv_l, ffi::AnyBuffer recv_r,
                       ffi::RemainingArgs aliased, ffi::Result<ffi::AnyBuffer> R, ffi::Result<ffi::AnyBuffer> info) {
  (void)aliased;
  return Status(jmd_dd_place(dtype, dim, cap_total, cap_list, static_cast<const int32_t*>(counters.untyped_data()),
                             recv_l.untyped_data(), recv_r.untyped_data(), R->untyped_data(),
                             static_cast<int32_t*>(info->untyped_data()), stream),
                "jmd_dd_place");
}

// comm_push / comm_wait: the jmd_dd_t descriptor carries PEER pointers (CUDA-IPC mappings
// owned by the library, not XLA buffers), so it is passed by address of a host-resident,
// library-owned struct (attribute "dd" = the address as int64).  R is an XLA buffer.
ffi::Error DdCommPushImpl(cudaStream_t stream, int64_t dd, ffi::AnyBuffer R) {
  return Status(jmd_dd_comm_push(reinterpret_cast<const jmd_dd_t*>(dd), R.untyped_data(), stream), "jmd_dd_comm_push");
}

ffi::Error DdCommWaitImpl(cudaStream_t stream, int64_t dd, std::string_view nbr_desc, ffi::RemainingArgs args,
                          ffi::RemainingRets rets) {
  (void)args;
  jmd_nbr_t nb;
  if (auto e = FillNbr(nbr_desc, rets, 1, &nb); e.failure()) return e;
  return Status(jmd_dd_comm_wait(reinterpret_cast<const jmd_dd_t*>(dd), &nb, OptRet(rets, 0), stream),
                "jmd_dd_comm_wait");
}

}  // namespace

// ---- exported handler symbols: jmd_ffi_<entry point> ----------------------------------------------
XLA_FFI_DEFINE_HANDLER_SYMBOL(jmd_ffi_nbr_update, NbrImpl<kUpdate>, JMD_NBR_BINDING);
XLA_FFI_DEFINE_HANDLER_SYMBOL(jmd_ffi_nbr_skin_check, NbrImpl<kSkinCheck>, JMD_NBR_BINDING);
XLA_FFI_DEFINE_HANDLER_SYMBOL(jmd_ffi_nbr_bin, NbrImpl<kBin>, JMD_NBR_BINDING);
XLA_FFI_DEFINE_HANDLER_SYMBOL(jmd_ffi_nbr_build, NbrImpl<kBuild>, JMD_NBR_BINDING);
XLA_FFI_DEFINE_HANDLER_SYMBOL(jmd_ffi_nbr_export, NbrImpl<kExport>, JMD_NBR_BINDING);
XLA_FFI_DEFINE_HANDLER_SYMBOL(jmd_ffi_nbr_pack, NbrImpl<kPack>, JMD_NBR_BINDING);
XLA_FFI_DEFINE_HANDLER_SYMBOL(jmd_ffi_nbr_pack_range, NbrImpl<kPackRange>, JMD_NBR_BINDING);

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    jmd_ffi_pair_force, PairForceImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Attr<std::string_view>("nbr")
        .Attr<std::string_view>("pair")
        .Attr<int32_t>("mass_is_array")
        .Attr<double>("dt_2")
        .Attr<int32_t>("want_energy")
        .Attr<int32_t>("kick")
        .RemainingArgs()
        .RemainingRets());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    jmd_ffi_sw_force, SwForceImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Attr<std::string_view>("nbr")
        .Attr<std::string_view>("sw")
        .Attr<int32_t>("mass_is_array")
        .Attr<double>("dt_2")
        .Attr<int32_t>("kick")
        .RemainingArgs()
        .RemainingRets());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    jmd_ffi_nve_kick_drift, KickDriftImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Attr<std::string_view>("space")
        .Attr<std::string_view>("nbr")
        .Attr<int32_t>("dtype")
        .Attr<int32_t>("n")
        .Attr<int32_t>("mass_is_array")
        .Attr<double>("dt")
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .RemainingArgs()
        .RemainingRets());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    jmd_ffi_kick_reduce, KickReduceImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Attr<int32_t>("dtype")
        .Attr<int32_t>("n")
        .Attr<int32_t>("dim")
        .Attr<int32_t>("mass_is_array")
        .Attr<double>("dt_2")
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .RemainingArgs()
        .Ret<ffi::AnyBuffer>()
        .Ret<ffi::AnyBuffer>()
        .Ret<ffi::AnyBuffer>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    jmd_ffi_scale_momentum, ScaleMomentumImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Attr<int32_t>("dtype")
        .Arg<ffi::AnyBuffer>()
        .Ret<ffi::AnyBuffer>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    jmd_ffi_nhc_half_step, NhcHalfStepImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Attr<int32_t>("dtype")
        .Attr<int32_t>("chain_length")
        .Attr<int32_t>("chain_steps")
        .Attr<int32_t>("sy_steps")
        .Attr<double>("dt")
        .Attr<double>("tau")
        .Attr<int64_t>("dof")
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .RemainingArgs()
        .Ret<ffi::AnyBuffer>()
        .Ret<ffi::AnyBuffer>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    jmd_ffi_fire_mix, FireMixImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Attr<int32_t>("dtype")
        .Attr<double>("dt_max")
        .Attr<double>("n_min")
        .Attr<double>("f_inc")
        .Attr<double>("f_dec")
        .Attr<double>("alpha_start")
        .Attr<double>("f_alpha")
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .Ret<ffi::AnyBuffer>()
        .Ret<ffi::AnyBuffer>()
        .Ret<ffi::AnyBuffer>());

#define JMD_DD_SELECT_BINDING                       \
  ffi::Ffi::Bind()                                  \
      .Ctx<ffi::PlatformStream<cudaStream_t>>()     \
      .Attr<int32_t>("dtype")                       \
      .Attr<int32_t>("dim")                         \
      .Attr<int32_t>("n")                           \
      .Attr<int32_t>("axis")                        \
      .Attr<double>("lo")                           \
      .Attr<double>("L")                            \
      .Attr<double>("thr_a")                        \
      .Attr<double>("thr_b")                        \
      .Attr<int32_t>("cap")                         \
      .Arg<ffi::AnyBuffer>()                        \
      .RemainingArgs()                              \
      .RemainingRets()
XLA_FFI_DEFINE_HANDLER_SYMBOL(jmd_ffi_dd_select, DdSelectImpl<false>, JMD_DD_SELECT_BINDING);
XLA_FFI_DEFINE_HANDLER_SYMBOL(jmd_ffi_dd_select_ordered, DdSelectImpl<true>, JMD_DD_SELECT_BINDING);

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    jmd_ffi_dd_pack, DdPackImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Attr<int32_t>("dtype")
        .Attr<int32_t>("ncomp")
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .Ret<ffi::AnyBuffer>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    jmd_ffi_dd_pack_counted, DdPackCountedImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Attr<int32_t>("dtype")
        .Attr<int32_t>("ncomp")
        .Attr<int32_t>("cap")
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .Ret<ffi::AnyBuffer>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    jmd_ffi_dd_pack_migrate, DdPackMigrateImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Attr<int32_t>("dtype")
        .Attr<int32_t>("dim")
        .Attr<int32_t>("cap_mig")
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .Ret<ffi::AnyBuffer>()
        .Ret<ffi::AnyBuffer>()
        .Ret<ffi::AnyBuffer>()
        .Ret<ffi::AnyBuffer>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    jmd_ffi_dd_compact, DdCompactImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Attr<int32_t>("dtype")
        .Attr<int32_t>("dim")
        .Attr<int32_t>("cap_own")
        .Attr<int32_t>("cap_mig")
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .RemainingArgs()
        .Ret<ffi::AnyBuffer>()
        .Ret<ffi::AnyBuffer>()
        .Ret<ffi::AnyBuffer>()
        .Ret<ffi::AnyBuffer>()
        .Ret<ffi::AnyBuffer>()
        .Ret<ffi::AnyBuffer>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    jmd_ffi_dd_place, DdPlaceImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Attr<int32_t>("dtype")
        .Attr<int32_t>("dim")
        .Attr<int32_t>("cap_total")
        .Attr<int32_t>("cap_list")
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .RemainingArgs()
        .Ret<ffi::AnyBuffer>()
        .Ret<ffi::AnyBuffer>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    jmd_ffi_dd_comm_push, DdCommPushImpl,
    ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("dd").Arg<ffi::AnyBuffer>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    jmd_ffi_dd_comm_wait, DdCommWaitImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Attr<int64_t>("dd")
        .Attr<std::string_view>("nbr")
        .RemainingArgs()
        .RemainingRets());

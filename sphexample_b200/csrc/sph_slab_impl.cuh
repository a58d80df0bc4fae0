// sph_slab_impl.cuh — slab-mode member functions of Sim<T, D> (included by sphb200.cu).
#pragma once

namespace sph {
int slab_unique_id(uint8_t *id_out) {
    (void)id_out;
    return SPHB200_ENCCL;
}
}  // namespace sph

namespace {
template <class T, int D>
int Sim<T, D>::comm_init(const uint8_t *, int, int, int) { return fail(SPHB200_ENCCL, "slab mode not available in this build"); }
template <class T, int D>
int Sim<T, D>::set_slab(int64_t, int64_t) { return fail(SPHB200_ENCCL, "slab mode not available in this build"); }
template <class T, int D>
int Sim<T, D>::column_histogram(int, int64_t *, int64_t *, int64_t *, int64_t) { return fail(SPHB200_ENCCL, "slab mode not available in this build"); }
template <class T, int D>
int Sim<T, D>::run_steps_slab(int64_t, bool) { return fail(SPHB200_ENCCL, "slab mode not available in this build"); }
}  // namespace

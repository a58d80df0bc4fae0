// sph_slab.cuh — multi-GPU slab decomposition state (one process per GPU, NCCL over NVLink).
#pragma once

#include "sph_device.cuh"

namespace sph {

struct SlabComm {
    bool active = false;
    int rank = 0, world = 1;
    template <class S>
    int allreduce_ctl(S *, Ctl *, cudaStream_t) { return 0; }
};

int slab_unique_id(uint8_t *id_out);

}  // namespace sph

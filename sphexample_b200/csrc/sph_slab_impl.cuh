// sph_slab_impl.cuh — slab-mode member functions of Sim<T, D> (included by sphb200.cu).
// See sph_slab.cuh for the decomposition and the exchange protocol.
#pragma once

namespace sph {

int slab_unique_id(uint8_t *id_out) {
    std::string err;
    if (!id_out || !nccl::api().load(err)) return SPHB200_ENCCL;
    nccl::UniqueId id;
    memset(&id, 0, sizeof id);
    if (nccl::api().GetUniqueId(&id) != nccl::Success) return SPHB200_ENCCL;
    memcpy(id_out, &id, 128);
    return SPHB200_OK;
}

// gather `cnt` table records listed in `list` into dst[off .. off+cnt)
template <class T, int D>
__global__ void k_slab_pack(const int *__restrict__ list, int cnt, int off, Table<T, D> src, Table<T, D> dst) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < cnt; k += gridDim.x * blockDim.x) {
        int s = list[k], p = off + k;
        dst.A[p] = src.A[s];
        dst.B[p] = src.B[s];
        dst.acc[p] = src.acc[s];
        dst.id[p] = src.id[s];
        dst.group[p] = src.group[s];
        dst.okey[p] = src.okey[s];
        dst.type[p] = src.type[s];
    }
}

}  // namespace sph

namespace {

#define CKS(call)                                                                                         \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail(SPHB200_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define NCK(call)                                                                                         \
    do {                                                                                                  \
        int r_ = (call);                                                                                  \
        if (r_ != nccl::Success)                                                                          \
            return fail(SPHB200_ENCCL, "%s failed: %s (%s:%d)", #call, nccl::api().GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

template <class T, int D>
int Sim<T, D>::comm_init(const uint8_t *uid, int rank, int world, int axis) {
    if (!uid || world < 1 || rank < 0 || rank >= world) return fail(SPHB200_EINVAL, "comm_init: bad rank/world");
    if (axis < 0 || axis >= D) return fail(SPHB200_EINVAL, "comm_init: the slab axis must be 0..%d", D - 1);
    if (slab.active) return fail(SPHB200_ESTATE, "comm_init: communicator already initialised");
    if (uploaded) return fail(SPHB200_ESTATE, "comm_init must precede upload");
    CKS(cudaSetDevice(device));
    std::string e;
    if (!nccl::api().load(e)) return fail(SPHB200_ENCCL, "%s", e.c_str());
    nccl::UniqueId id;
    memcpy(&id, uid, 128);
    NCK(nccl::api().CommInitRank(&slab.comm, world, id, rank));
    slab.rank = rank;
    slab.world = world;
    slab.left = rank > 0 ? rank - 1 : -1;
    slab.right = rank + 1 < world ? rank + 1 : -1;
    {
        int prio_lo = 0, prio_hi = 0;
        CKS(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CKS(cudaStreamCreateWithPriority(&slab.xstream, cudaStreamNonBlocking, prio_hi));
    }
    CKS(cudaEventCreateWithFlags(&slab.ev_bnd, cudaEventDisableTiming));
    CKS(cudaEventCreateWithFlags(&slab.ev_x, cudaEventDisableTiming));
    CKS(cudaMalloc((void **)&slab.d_flag, sizeof(unsigned)));
    CKS(cudaMemset(slab.d_flag, 0, sizeof(unsigned)));
    {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        // EXPERIMENTAL, off by default: the first r1 attempt (exchange enqueued before the launch that
        // raises the flag) hung on 4 GPUs — presumably NCCL's first-use connection setup synchronises
        // the stream it is given, which then waits on a flag no launched kernel will raise.  The
        // order below (launch first) has not been validated on hardware yet.
        if (getenv("SPHB200_SLAB_WAITVALUE") &&
            cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            *(void **)(&slab.wait_value32) = fn;
        cudaGetLastError();
    }
    CKS(cudaMalloc((void **)&slab.d_counts, 8 * sizeof(int)));
    CKS(cudaMallocHost((void **)&slab.h_counts, 8 * sizeof(int)));
    // the slab axis is the slowest key component (a slab layer is one contiguous range of the table); of the
    // other axes the lower component stays the fastest (x, or y when the slabs are cut along x)
    am.ax_s = axis;
    if (D == 3) {
        am.ax_f = axis == 0 ? 1 : 0;
        am.ax_m = axis == 2 ? 1 : 2;
    } else {
        am.ax_f = axis == 0 ? 1 : 0;
        am.ax_m = am.ax_f;
    }
    slab.active = true;
    return SPHB200_OK;
}

template <class T, int D>
int Sim<T, D>::set_slab(int64_t lo, int64_t hi) {
    if (!slab.active) return fail(SPHB200_ESTATE, "set_slab before comm_init");
    if (hi - lo < 2 && lo != INT64_MIN && hi != INT64_MAX) return fail(SPHB200_EINVAL, "a slab must be at least 2 cell layers wide");
    own_lo = lo <= (int64_t)INT_MIN ? INT_MIN : (int)lo;
    own_hi = hi >= (int64_t)INT_MAX ? INT_MAX : (int)hi;
    if (slab.left < 0) own_lo = INT_MIN;
    if (slab.right < 0) own_hi = INT_MAX;
    return SPHB200_OK;
}

// SimpleMDBC across slabs: the global, static ghost-node table (k_mdbc_nodes, sph_step.cuh).  Every rank
// passes the same table: points[ng][D] of T and the IDs of the particles the nodes belong to, ascending.
template <class T, int D>
int Sim<T, D>::set_ghost_nodes(int64_t ng, const void *points, const int64_t *ids) {
    if (!slab.active) return fail(SPHB200_ESTATE, "set_ghost_nodes is the slab-mode form of the ghost columns: call comm_init first (one GPU: pass ghost_points to upload)");
    if (!prm.mdbc) return fail(SPHB200_EINVAL, "set_ghost_nodes: the handle was created with NoMDBC");
    if (ng < 0 || ng > (int64_t)INT_MAX / 8 || (ng > 0 && (!points || !ids))) return fail(SPHB200_EINVAL, "set_ghost_nodes: bad arguments");
    for (int64_t g = 1; g < ng; ++g)
        if (ids[g] <= ids[g - 1]) return fail(SPHB200_EINVAL, "set_ghost_nodes: particle IDs must be strictly ascending (entry %lld)", (long long)g);
    CKS(cudaSetDevice(device));
    CKS(cudaStreamSynchronize(stream));
    CKS(g_point.alloc((size_t)ng + 1));
    CKS(g_id.alloc((size_t)ng + 1));
    CKS(g_sol.alloc((size_t)ng * (D + 2) + 1));
    if (ng > 0) {
        std::vector<TV> hp((size_t)ng);
        const T *src = (const T *)points;
        memset(hp.data(), 0, hp.size() * sizeof(TV));
        for (int64_t g = 0; g < ng; ++g)
            for (int k = 0; k < D; ++k) ((T *)&hp[(size_t)g])[k] = src[(size_t)g * D + k];   // TV = D (2D) or 4 (3D) packed T
        CKS(cudaMemcpyAsync(g_point.p, hp.data(), (size_t)ng * sizeof(TV), cudaMemcpyHostToDevice, stream));
        CKS(cudaMemcpyAsync(g_id.p, ids, (size_t)ng * sizeof(long long), cudaMemcpyHostToDevice, stream));
        CKS(cudaStreamSynchronize(stream));
    }
    n_ghost_nodes = (int)ng;
    return SPHB200_OK;
}

// S6 in slab mode: solve the nodes whose cell this rank owns, all-reduce the solves, extrapolate to
// every particle held here (owned and halo copies alike)
template <class T, int D>
int Sim<T, D>::slab_enqueue_mdbc() {
    if (n_ghost_nodes < 0) return fail(SPHB200_ESTATE, "SimpleMDBC in slab mode needs the ghost-node table (sphb200_set_ghost_nodes) before the first step");
    const int ng = n_ghost_nodes;
    if (ng == 0) return SPHB200_OK;
    k_mdbc_nodes<T, D><<<grid_for((int64_t)ng * 32, 128), 128, 0, stream>>>(A.p, g_point.p, ng, type.p, cell_start.p, d_grid.p, am, ph, prm.H_inv,
                                                              own_lo, own_hi, g_sol.p, d_ctl.p);
    ++launches;
    CKS(cudaGetLastError());
    NCK(nccl::api().AllReduce(g_sol.p, g_sol.p, (size_t)ng * (D + 2), nccl::Float64, nccl::Sum, slab.comm, stream));
    k_mdbc_apply_nodes<T, D><<<grid_for(n), 256, 0, stream>>>(A.p, RN.p, type.p, id.p, (int)n, g_point.p, g_id.p, ng, g_sol.p, ph.rho0,
                                                              d_ctl.p);
    ++launches;
    CKS(cudaGetLastError());
    return SPHB200_OK;
}

template <class T, int D>
int Sim<T, D>::column_histogram(int axis, int64_t *cell_min, int64_t *n_columns, int64_t *counts, int64_t cap) {
    if (!uploaded) return fail(SPHB200_ESTATE, "column_histogram before upload");
    if (axis < 0 || axis >= D || !cell_min || !n_columns) return fail(SPHB200_EINVAL, "column_histogram: bad arguments");
    CKS(cudaSetDevice(device));
    // range of cell coordinates along `axis`: from a fresh bounding box of the owned particles
    const int p0 = slab.active ? slab.own_p0 : 0, p1 = slab.active ? slab.own_p1 : (int)n;
    std::vector<T> hx((size_t)(p1 - p0) * (sizeof(TA) / sizeof(T)));
    CKS(cudaMemcpyAsync(hx.data(), A.p + p0, (size_t)(p1 - p0) * sizeof(TA), cudaMemcpyDeviceToHost, stream));
    CKS(cudaStreamSynchronize(stream));
    const int stride = sizeof(TA) / sizeof(T);
    auto cell = [&](double x) { double t = trunc(fma(fabs(x), prm.H_inv, 0.5)); return (int64_t)(((x > 0) - (x < 0)) * t); };
    int64_t cmin = INT64_MAX, cmax = INT64_MIN;
    for (int i = 0; i < p1 - p0; ++i) {
        int64_t c = cell((double)hx[(size_t)i * stride + axis]);
        cmin = std::min(cmin, c);
        cmax = std::max(cmax, c);
    }
    if (p1 <= p0) { cmin = 0; cmax = -1; }
    *cell_min = cmin;
    *n_columns = cmax - cmin + 1;
    if (counts) {
        if (cap < *n_columns) return fail(SPHB200_ECAPACITY, "column_histogram: need room for %lld columns", (long long)*n_columns);
        for (int64_t k = 0; k < *n_columns; ++k) counts[k] = 0;
        for (int i = 0; i < p1 - p0; ++i) counts[cell((double)hx[(size_t)i * stride + axis]) - cmin] += 1;
    }
    return SPHB200_OK;
}

// ---- small exchanges --------------------------------------------------------------------------
// d_counts[0], [1] -> left, right neighbour; their values arrive in d_counts[2] (from left), [3] (from right)
template <class T, int D>
int Sim<T, D>::slab_exchange_counts(int to_left, int to_right, int *from_left, int *from_right) {
    slab.h_counts[0] = to_left;
    slab.h_counts[1] = to_right;
    slab.h_counts[2] = slab.h_counts[3] = 0;
    CKS(cudaMemcpyAsync(slab.d_counts, slab.h_counts, 4 * sizeof(int), cudaMemcpyHostToDevice, stream));
    NCK(nccl::api().GroupStart());
    if (slab.left >= 0) {
        NCK(nccl::api().Send(slab.d_counts + 0, 1, nccl::Int32, slab.left, slab.comm, stream));
        NCK(nccl::api().Recv(slab.d_counts + 2, 1, nccl::Int32, slab.left, slab.comm, stream));
    }
    if (slab.right >= 0) {
        NCK(nccl::api().Send(slab.d_counts + 1, 1, nccl::Int32, slab.right, slab.comm, stream));
        NCK(nccl::api().Recv(slab.d_counts + 3, 1, nccl::Int32, slab.right, slab.comm, stream));
    }
    NCK(nccl::api().GroupEnd());
    CKS(cudaMemcpyAsync(slab.h_counts, slab.d_counts, 4 * sizeof(int), cudaMemcpyDeviceToHost, stream));
    {
        int rcw = wait_stream();
        if (rcw) return rcw;
    }
    *from_left = slab.h_counts[2];
    *from_right = slab.h_counts[3];
    return SPHB200_OK;
}

// full records: src ranges [sl0, sl0+nl) -> left, [sr0, sr0+nr) -> right of table `from`;
// arrivals are appended to table(false) at dst0 (left's block first)
template <class T, int D>
int Sim<T, D>::slab_exchange_records(Table<T, D> from, int sl0, int nl, int sr0, int nr, int dst0, int rl, int rr) {
    Table<T, D> to = table(false);
    NCK(nccl::api().GroupStart());
    auto xfer = [&](auto *src, auto *dst) -> int {
        const size_t es = sizeof(*src);
        if (slab.left >= 0) {
            if (nl) NCK(nccl::api().Send(src + sl0, (size_t)nl * es, nccl::Int8, slab.left, slab.comm, stream));
            if (rl) NCK(nccl::api().Recv(dst + dst0, (size_t)rl * es, nccl::Int8, slab.left, slab.comm, stream));
        }
        if (slab.right >= 0) {
            if (nr) NCK(nccl::api().Send(src + sr0, (size_t)nr * es, nccl::Int8, slab.right, slab.comm, stream));
            if (rr) NCK(nccl::api().Recv(dst + dst0 + rl, (size_t)rr * es, nccl::Int8, slab.right, slab.comm, stream));
        }
        return 0;
    };
    int rc;
    if ((rc = xfer(from.A, to.A))) return rc;
    if ((rc = xfer(from.B, to.B))) return rc;
    if ((rc = xfer(from.acc, to.acc))) return rc;
    if ((rc = xfer(from.id, to.id))) return rc;
    if ((rc = xfer(from.group, to.group))) return rc;
    if ((rc = xfer(from.okey, to.okey))) return rc;
    if ((rc = xfer(from.type, to.type))) return rc;
    NCK(nccl::api().GroupEnd());
    return SPHB200_OK;
}

// the half-step exchange: boundary layers of two packed arrays to the neighbours' halo ranges
template <class T, int D>
int Sim<T, D>::slab_exchange_halo(TA *a, TB *b, cudaStream_t stream) {
    const SlabComm &s = slab;
    const int nf = s.l1 - s.own_p0, nl = s.own_p1 - s.l2, hl = s.own_p0, hr = (int)n - s.own_p1;
    NCK(nccl::api().GroupStart());
    if (s.left >= 0) {
        if (nf) {
            NCK(nccl::api().Send(a + s.own_p0, (size_t)nf * sizeof(TA), nccl::Int8, s.left, s.comm, stream));
            NCK(nccl::api().Send(b + s.own_p0, (size_t)nf * sizeof(TB), nccl::Int8, s.left, s.comm, stream));
        }
        if (hl) {
            NCK(nccl::api().Recv(a, (size_t)hl * sizeof(TA), nccl::Int8, s.left, s.comm, stream));
            NCK(nccl::api().Recv(b, (size_t)hl * sizeof(TB), nccl::Int8, s.left, s.comm, stream));
        }
    }
    if (s.right >= 0) {
        if (nl) {
            NCK(nccl::api().Send(a + s.l2, (size_t)nl * sizeof(TA), nccl::Int8, s.right, s.comm, stream));
            NCK(nccl::api().Send(b + s.l2, (size_t)nl * sizeof(TB), nccl::Int8, s.right, s.comm, stream));
        }
        if (hr) {
            NCK(nccl::api().Recv(a + s.own_p1, (size_t)hr * sizeof(TA), nccl::Int8, s.right, s.comm, stream));
            NCK(nccl::api().Recv(b + s.own_p1, (size_t)hr * sizeof(TB), nccl::Int8, s.right, s.comm, stream));
        }
    }
    NCK(nccl::api().GroupEnd());
    return SPHB200_OK;
}

template <class T, int D>
int Sim<T, D>::slab_allreduce_ctl() {
    k_slab_pre_allreduce<<<1, 1, 0, stream>>>(d_ctl.p);
    ++launches;
    NCK(nccl::api().AllReduce(&d_ctl.p->red_disp2, &d_ctl.p->red_disp2, 5, nccl::Uint64, nccl::Max, slab.comm, stream));
    return SPHB200_OK;
}

// one sort of the live part of the table (retrying after a dense-grid growth)
template <class T, int D>
int Sim<T, D>::slab_sort(const SlabFilter &flt, int count_rebuild) {
    int rc;
    for (int attempt = 0; attempt < 3; ++attempt) {
        if ((rc = force_flag_rebuild())) return rc;
        if ((rc = enqueue_rebuild(flt, count_rebuild))) return rc;
        if ((rc = sync_ctl())) return rc;
        if (h_ctl->error != SPHB200_ECAPACITY) break;
        if ((rc = recover_capacity())) return rc;
    }
    if (h_ctl->error) return fail(h_ctl->error, "slab rebuild: device reported error %d", h_ctl->error);
    return SPHB200_OK;
}

// UpdateNeighbors! across slabs: migrate, sort the owned set, swap boundary layers, sort again
template <class T, int D>
int Sim<T, D>::slab_rebuild() {
    int rc;
    SlabComm &s = slab;
    const int n_old = (int)n;
    // 1. owned particles that left the slab
    CKS(cudaMemsetAsync(s.d_counts + 4, 0, 2 * sizeof(int), stream));
    k_slab_classify<T, D><<<grid_for(s.own_p1 - s.own_p0), 256, 0, stream>>>(A.p, s.own_p0, s.own_p1, prm.H_inv, am.ax_s, own_lo,
                                                                            own_hi, tmp_idx.p, perm.p, s.d_counts + 4, d_ctl.p);
    ++launches;
    CKS(cudaMemcpyAsync(s.h_counts + 4, s.d_counts + 4, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
    CKS(cudaStreamSynchronize(stream));
    const int ml = s.h_counts[4], mr = s.h_counts[5];
    // deterministic order of the migrants: ascending table index (the atomics append in any order)
    auto sort_list = [&](int *dlist, int cnt) -> int {
        if (cnt < 2) return 0;
        std::vector<int> h((size_t)cnt);
        CKS(cudaMemcpyAsync(h.data(), dlist, (size_t)cnt * 4, cudaMemcpyDeviceToHost, stream));
        CKS(cudaStreamSynchronize(stream));
        std::sort(h.begin(), h.end());
        CKS(cudaMemcpyAsync(dlist, h.data(), (size_t)cnt * 4, cudaMemcpyHostToDevice, stream));
        CKS(cudaStreamSynchronize(stream));
        return 0;
    };
    if ((rc = sort_list(tmp_idx.p, ml))) return rc;
    if ((rc = sort_list(perm.p, mr))) return rc;
    if (ml) k_slab_pack<T, D><<<grid_for(ml), 256, 0, stream>>>(tmp_idx.p, ml, 0, table(false), table(true));
    if (mr) k_slab_pack<T, D><<<grid_for(mr), 256, 0, stream>>>(perm.p, mr, ml, table(false), table(true));
    launches += (ml > 0) + (mr > 0);
    int rl = 0, rr = 0;
    if ((rc = slab_exchange_counts(ml, mr, &rl, &rr))) return rc;
    if ((rc = grow_particles((int64_t)n_old + rl + rr))) return rc;
    if ((rc = slab_exchange_records(table(true), 0, ml, ml, mr, n_old, rl, rr))) return rc;
    if (rl + rr) {
        k_slab_check_arrivals<T, D><<<grid_for(rl + rr), 256, 0, stream>>>(A.p, n_old, n_old + rl + rr, prm.H_inv, am.ax_s, own_lo,
                                                                           own_hi, d_ctl.p);
        ++launches;
    }
    s.n_migrated += ml + mr;
    // 2. sort A: the owned set only (old halo copies and the migrants that left are dropped)
    n = n_old + rl + rr;
    SlabFilter fa = {1, s.own_p0, s.own_p1, n_old, own_lo, own_hi};
    if ((rc = slab_sort(fa, 0))) return rc;
    n = h_grid->n_total;
    if (n < 1) return fail(SPHB200_ESTATE, "rank %d owns no particles (slab [%d, %d))", s.rank, own_lo, own_hi);
    const int cf = h_grid->own_l1 - h_grid->own_p0, cl = h_grid->own_p1 - h_grid->own_l2;
    const int l2 = h_grid->own_l2;
    // 3. boundary layers -> neighbours' halos (full records), then sort B over owned + halo
    int hl = 0, hr = 0;
    if ((rc = slab_exchange_counts(cf, cl, &hl, &hr))) return rc;
    if ((rc = grow_particles(n + hl + hr))) return rc;
    if ((rc = slab_exchange_records(table(false), 0, cf, l2, cl, (int)n, hl, hr))) return rc;
    n += hl + hr;
    SlabFilter fb = {0, 0, 0, 0, 0, 0};
    if ((rc = slab_sort(fb, 1))) return rc;
    s.own_p0 = h_grid->own_p0;
    s.own_p1 = h_grid->own_p1;
    s.l1 = h_grid->own_l1;
    s.l2 = h_grid->own_l2;
    if (s.own_p0 != hl || (int)n - s.own_p1 != hr || h_grid->n_total != (int)n)
        return fail(SPHB200_ESTATE, "slab rebuild: halo layers are not where expected (left %d/%d, right %d/%d)", s.own_p0, hl,
                    (int)n - s.own_p1, hr);
    s.halo_bytes_per_step = 2ll * ((s.l1 - s.own_p0) * (s.left >= 0) + (s.own_p1 - s.l2) * (s.right >= 0)) * (long long)(sizeof(TA) + sizeof(TB));
    return SPHB200_OK;
}

// One interaction pass over the owned bricks with its halo exchange hidden behind the interior
// bricks.  Nothing the exchange writes (the halo ranges of xa/xb) is read, and nothing it reads is
// written, before the next pass.  xev (optional, 4 events, stage_times): [0] pass start (main
// stream), [1] exchange released, [2] exchange complete (both on the exchange stream), [3] interior
// bricks done (main stream).
template <class T, int D>
int Sim<T, D>::slab_pass(int pass, TA *xa, TB *xb, cudaEvent_t *xev) {
    int rc;
    if (xev) CKS(cudaEventRecord(xev[0], stream));
    if (slab.wait_value32) {
        // one launch: boundary bricks first, the flag releases the exchange, interior bricks go on
        const unsigned epoch = ++slab.epoch;
        bnd_flag = slab.d_flag;
        bnd_epoch = epoch;
        rc = launch_interact(pass, EPI_FUSED);
        bnd_flag = nullptr;
        if (rc) return rc;
        k_slab_signal<<<1, 1, 0, stream>>>(slab.d_flag, epoch);
        ++launches;
        if (xev) CKS(cudaEventRecord(xev[3], stream));
        // everything that raises the flag is in flight before anything waits on it
        if (slab.wait_value32(slab.xstream, (unsigned long long)(uintptr_t)slab.d_flag, epoch, 0u /* GEQ */) != 0)
            return fail(SPHB200_ECUDA, "cuStreamWaitValue32 failed");
        if (xev) CKS(cudaEventRecord(xev[1], slab.xstream));
        if ((rc = slab_exchange_halo(xa, xb, slab.xstream))) return rc;
        CKS(cudaEventRecord(slab.ev_x, slab.xstream));
        if (xev) CKS(cudaEventRecord(xev[2], slab.xstream));
        CKS(cudaStreamWaitEvent(stream, slab.ev_x, 0));
        return SPHB200_OK;
    }
    // default: two launches, boundary bricks then interior bricks
    brick_part = 1;
    rc = launch_interact(pass, EPI_FUSED);
    if (!rc) {
        CKS(cudaEventRecord(slab.ev_bnd, stream));
        CKS(cudaStreamWaitEvent(slab.xstream, slab.ev_bnd, 0));
        if (xev) CKS(cudaEventRecord(xev[1], slab.xstream));
        rc = slab_exchange_halo(xa, xb, slab.xstream);
    }
    if (!rc) {
        CKS(cudaEventRecord(slab.ev_x, slab.xstream));
        if (xev) CKS(cudaEventRecord(xev[2], slab.xstream));
        brick_part = 2;
        rc = launch_interact(pass, EPI_FUSED);
        if (xev) CKS(cudaEventRecord(xev[3], stream));
    }
    brick_part = 0;
    if (rc) return rc;
    CKS(cudaStreamWaitEvent(stream, slab.ev_x, 0));
    return SPHB200_OK;
}

// the step that raised do_rebuild paused itself (k_step_control): clear the pause, rebuild across
// ranks; the caller then runs the body of the still-open step
template <class T, int D>
int Sim<T, D>::slab_resume_after_pause() {
    h_ctl->done = 0;
    h_ctl->paused = 0;
    int rc = push_ctl();
    if (rc) return rc;
    return slab_rebuild();
}

// S2 .. S19 of one step in slab mode.  ev (optional, 9 events): the stage boundaries of
// enqueue_step_body; xev (optional, 8 events): the exchange probes of the two passes (slab_pass).
template <class T, int D>
int Sim<T, D>::slab_step_body(cudaEvent_t *ev, bool host_synced, cudaEvent_t *xev) {
    int rc;
#define EV(k) if (ev) CKS(cudaEventRecord(ev[k], stream))
    EV(0);
    if (host_synced && h_ctl->do_rebuild && (rc = slab_resume_after_pause())) return rc;   // S2 (needs the host: exchange sizes)
    EV(1);
    if ((rc = enqueue_motion(-1.0))) return rc;                       // S3
    if ((rc = enqueue_snapshots())) return rc;
    EV(2);
    if (prm.mdbc && (rc = slab_enqueue_mdbc())) return rc;            // S6
    EV(3);
    if ((rc = enqueue_list_build())) return rc;
    EV(4);
    if ((rc = slab_pass(0, Ah.p, Bh.p, xev))) return rc;              // S4-S10, S13 + halo state n+1/2
    EV(5);
    if ((rc = enqueue_motion(-1.0))) return rc;                       // S12
    EV(6);
    if ((rc = slab_pass(1, A.p, B.p, xev ? xev + 4 : nullptr))) return rc;   // S11, S14-S18 + halo state n+1
    EV(7);
    k_step_end<<<1, 1, 0, stream>>>(d_ctl.p, d_grid.p);                         // S19
    ++launches;
    EV(8);
#undef EV
    have_half = true;
    have_cells = true;
    return SPHB200_OK;
}

template <class T, int D>
int Sim<T, D>::slab_check_head(bool *stop, bool until_target) {
    int rc;
    if ((rc = enqueue_step_head())) return rc;
    if ((rc = sync_ctl())) return rc;
    if (h_ctl->error == SPHB200_ENUMERIC)
        return fail(SPHB200_ENUMERIC, "non-finite state or time step at iteration %lld (t = %g)", h_ctl->iteration, h_ctl->total_time);
    if (h_ctl->error) return fail(h_ctl->error, "device reported error %d (rank %d)", h_ctl->error, slab.rank);
    *stop = until_target && h_ctl->done && !h_ctl->do_rebuild;
    return SPHB200_OK;
}

template <class T, int D>
int Sim<T, D>::run_steps_slab(int64_t nsteps, bool until_target) {
    int rc = sync_ctl();
    if (rc) return rc;
    const int64_t it0 = h_ctl->iteration;
    int64_t done_steps = 0;
    const int64_t slab_batch = std::max(1, std::min(opt_batch, 16));
    while (until_target || done_steps < nsteps) {
        int64_t batch = slab_batch;
        if (!until_target) batch = std::min<int64_t>(batch, nsteps - done_steps);
        else if (h_ctl->current_dt > 0.0) {
            double rem = (h_ctl->target_time - h_ctl->total_time) / h_ctl->current_dt;
            batch = std::max<int64_t>(1, std::min<int64_t>(batch, (int64_t)(rem * 1.02) + 2));
        } else {
            batch = 1;
        }
        // enqueue a batch blind; a step that needs a rebuild pauses itself and everything behind it
        for (int64_t k = 0; k < batch; ++k) {
            if ((rc = enqueue_step_head())) return rc;
            if ((rc = slab_step_body(nullptr, false))) return rc;
        }
        if ((rc = sync_ctl())) return rc;
        if (h_ctl->error == SPHB200_ENUMERIC)
            return fail(SPHB200_ENUMERIC, "non-finite state or time step at iteration %lld (t = %g)", h_ctl->iteration, h_ctl->total_time);
        if (h_ctl->error) return fail(h_ctl->error, "device reported error %d (rank %d)", h_ctl->error, slab.rank);
        if (h_ctl->done && h_ctl->do_rebuild) {
            // paused at a rebuild: every rank sees the same flags (identical all-reduced inputs)
            if ((rc = slab_resume_after_pause())) return rc;
            if ((rc = slab_step_body(nullptr, false))) return rc;     // the open step
            if ((rc = sync_ctl())) return rc;
            if (h_ctl->error) return fail(h_ctl->error, "device reported error %d (rank %d)", h_ctl->error, slab.rank);
        }
        done_steps = h_ctl->iteration - it0;
        if (until_target && h_ctl->done && !h_ctl->do_rebuild) break;
    }
    return SPHB200_OK;
}

// slab-mode stage times of ONE extra step (collective), SPHB200_STAGE_* of include/sphb200.h.  The
// head is split at the all-reduce: what a rank waits there for the slowest rank is reported as
// SPHB200_STAGE_ALLREDUCE, not as reduction time.  A pass's time includes whatever of its halo
// exchange the interior bricks did not hide; the exchange itself is timed on the exchange stream.
template <class T, int D>
int Sim<T, D>::slab_stage_times(double *ms_out, int cnt) {
    cudaEvent_t ev[9], xev[8], hev[4];
    for (auto &e : ev) CKS(cudaEventCreate(&e));
    for (auto &e : xev) CKS(cudaEventCreate(&e));
    for (auto &e : hev) CKS(cudaEventCreate(&e));
    int rc;
    // head: reductions | all-reduce | control
    const int p0 = slab.own_p0, p1 = slab.own_p1;
    CKS(cudaEventRecord(hev[0], stream));
    k_reduce_dt_dx<T, D><<<grid_for(p1 - p0), 256, 0, stream>>>(A.p, B.p, Ah.p, acc.p, p0, p1, ph.h, ph.eta2, have_half ? 1 : 0, d_ctl.p);
    ++launches;
    CKS(cudaEventRecord(hev[1], stream));
    if ((rc = slab_allreduce_ctl())) return rc;
    CKS(cudaEventRecord(hev[2], stream));
    k_step_control<T><<<1, 1, 0, stream>>>(d_ctl.p, d_grid.p, ph.h, ph.c0, (T)prm.cfl, lists_on() ? opt_skin * prm.H : 0.0,
                                           motion_vmax(), 1, opt_list_local, 0ull, 0ull);
    ++launches;
    CKS(cudaEventRecord(hev[3], stream));
    if ((rc = sync_ctl())) return rc;
    if (h_ctl->error) return fail(h_ctl->error, "device reported error %d (rank %d)", h_ctl->error, slab.rank);
    rc = slab_step_body(ev, true, xev);
    if (rc) return rc;
    CKS(cudaStreamSynchronize(stream));
    CKS(cudaStreamSynchronize(slab.xstream));
    double st[SPHB200_N_STAGES] = {0};
    auto el = [&](cudaEvent_t a, cudaEvent_t b) -> double {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) { cudaGetLastError(); return 0.0; }
        return (double)ms;
    };
    st[SPHB200_STAGE_TIMESTEP] = el(hev[0], hev[1]) + el(hev[2], hev[3]);
    st[SPHB200_STAGE_ALLREDUCE] = el(hev[1], hev[2]);
    for (int k = 0; k < 8; ++k) st[SPHB200_STAGE_REBUILD + k] = el(ev[k], ev[k + 1]);
    for (int p = 0; p < 2; ++p) {
        cudaEvent_t *x = xev + 4 * p;
        const bool have_x = slab.left >= 0 || slab.right >= 0;
        const double dur = have_x ? el(x[1], x[2]) : 0.0;
        // exposed: how long after the interior bricks the exchange still ran
        const double tail = have_x ? el(x[0], x[2]) - el(x[0], x[3]) : 0.0;
        st[p ? SPHB200_STAGE_HALO2 : SPHB200_STAGE_HALO1] = dur;
        st[p ? SPHB200_STAGE_HALO2_EXPOSED : SPHB200_STAGE_HALO1_EXPOSED] = tail > 0.0 ? tail : 0.0;
    }
    for (int k = 0; k < cnt && k < SPHB200_N_STAGES; ++k) ms_out[k] = st[k];
    for (auto &e : ev) cudaEventDestroy(e);
    for (auto &e : xev) cudaEventDestroy(e);
    for (auto &e : hev) cudaEventDestroy(e);
    if ((rc = sync_ctl())) return rc;
    if (h_ctl->error) return fail(h_ctl->error, "device reported error %d", h_ctl->error);
    return SPHB200_OK;
}

#undef CKS
#undef NCK
}  // namespace

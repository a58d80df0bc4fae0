// sph_listorder.h — bank-aware ("rainbow") ordering of one particle's neighbour list (host + device;
// the CPU tests drive this very code through tests/physics_shim.cpp).
//
// The list kernel gathers one 16-byte candidate record per lane and list position with LDS.128.
// The 8 lanes of a quarter warp are served together; two lanes that want DIFFERENT records in the
// same 16-byte bank group (record index mod 8) cost an extra wavefront.  With lists in window order
// the indices are effectively random: 9.3 wavefronts per warp-wide LDS.128 in the CPU model
// (scripts/sim_list_conflicts.py), 12.5 measured on B200 (profiles/r1m) — the list kernel was bound
// by exactly that.  The order of the entries inside a particle's list is free (it only changes the
// summation order), so lane q (= particle index in its brick, mod 8) stores its list as "rainbow"
// chunks of 8: slot u of EVERY chunk holds an entry of bank group (q + u) mod 8.  At any list
// position the 8 lanes of a quarter warp then read 8 different bank groups: conflict-free.
// A particle's entries are not spread evenly over the 8 groups; entries of an over-full group
// (more than the list has chunks) go into the holes the under-full groups leave, and what holes
// remain are padded with the sentinel of THEIR group (8 sentinel records, one per group, sit behind
// the staged window).  CPU model of the result: 4.4 wavefronts per LDS.128 (4 is the floor).
#pragma once

#if defined(__CUDACC__)
#define SPH_LO_HD __host__ __device__ __forceinline__
#else
#define SPH_LO_HD inline
#endif

namespace sph {

constexpr unsigned LIST_INDEX_MASK = 0x7fffu;   // entry = window index | role << 15

// One particle.  in(k), k < n_slots (a multiple of 8): the list as built, window order, padded with
// entries whose index is >= total8.  out(c, u): slot u of chunk c of the reordered list
// (n_slots / 8 chunks).  ovf(p), p < ovf_cap: scratch for the entries of over-full groups.
// q: the particle's lane phase.  Returns false (and leaves `out` unusable) when the scratch is too
// small — the caller then keeps the list as built, which is valid, only slower.
template <class In, class Out, class Ovf>
SPH_LO_HD bool rainbow_order(int n_slots, int q, unsigned total8, In in, Out out, Ovf ovf, int ovf_cap) {
    const int nc = n_slots >> 3;
    unsigned cnt_lo = 0u, cnt_hi = 0u;   // 8 x 8-bit entry counts per bank group (a list holds < 256 per group)
    int novf = 0;
    bool ok = true;
    for (int k = 0; k < n_slots; ++k) {
        const unsigned e = in(k);
        const unsigned idx = e & LIST_INDEX_MASK;
        if (idx >= total8) continue;   // padding of the build
        const unsigned r = idx & 7u;
        const unsigned sh = (r & 3u) * 8u;
        const unsigned c = (((r & 4u) ? cnt_hi : cnt_lo) >> sh) & 0xffu;
        if (r & 4u) cnt_hi += 1u << sh;
        else cnt_lo += 1u << sh;
        if ((int)c < nc) {
            out((int)c, (int)((r - (unsigned)q) & 7u)) = (unsigned short)e;
        } else if (novf < ovf_cap) {
            ovf(novf++) = (unsigned short)e;
        } else {
            ok = false;
        }
    }
    if (!ok) return false;
    // holes: chunks [count of group r, nc) of slot (r - q) & 7; over-full groups' entries first, then
    // the group's own sentinel (window index total8 + r: same bank group as the slot expects)
    int po = 0;
    for (unsigned r = 0; r < 8u; ++r) {
        const unsigned c0 = (((r & 4u) ? cnt_hi : cnt_lo) >> ((r & 3u) * 8u)) & 0xffu;
        const int u = (int)((r - (unsigned)q) & 7u);
        for (int c = (int)c0; c < nc; ++c) out(c, u) = (po < novf) ? (unsigned short)ovf(po++) : (unsigned short)(total8 + r);
    }
    return true;
}

}  // namespace sph

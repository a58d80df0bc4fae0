// sph_listorder.h — bank-aware ordering of neighbour-list entries (host + device, shared with the
// CPU tests through tests/physics_shim.cpp).
//
// The list kernel gathers one 16-byte candidate record per lane and list position with LDS.128.
// The 8 lanes of a quarter warp are served together; two lanes that want DIFFERENT records in the
// same 16-byte bank group (record index mod 8) cost an extra wavefront.  With lists in window order
// the indices are effectively random: 3.1 wavefronts per quarter-warp load measured on B200
// (profiles/r1m), 2.5 in the CPU model (scripts/sim_list_conflicts.py).  The order of the entries
// inside a particle's list is free (it only changes the summation order), so the build can hand
// lane q (= lane & 7) an order whose k-th entry lies in bank group (q + k) mod 8 whenever it still
// has one: the 8 lanes then hit 8 different groups.  CPU model: 1.3 wavefronts per load.
#pragma once

#if defined(__CUDACC__)
#define SPH_LO_HD __host__ __device__ __forceinline__
#else
#define SPH_LO_HD inline
#endif

namespace sph {

// Reorders a batch of m <= 64 entries.  `in(k)` reads entry k of the batch (window order),
// `tmp(p)` is a reference to scratch slot p (m slots).  After prepare(), pull(k) returns the entry
// for position k = 0 .. m-1 of the batch, each entry exactly once.
struct BankRotator {
    unsigned long long next, rem;   // 8 x 8-bit: next scratch slot / entries left, per bank group
    int q;

    template <class In, class Tmp>
    SPH_LO_HD void prepare(int m, int lane_q, In in, Tmp tmp) {
        q = lane_q & 7;
        unsigned long long cnt = 0;
        for (int k = 0; k < m; ++k) cnt += 1ull << (8 * (in(k) & 7u));
        unsigned long long startp = 0;
        unsigned acc = 0;
        for (int r = 0; r < 8; ++r) {
            startp |= (unsigned long long)acc << (8 * r);
            acc += (unsigned)((cnt >> (8 * r)) & 0xffull);
        }
        unsigned long long fill = startp;
        for (int k = 0; k < m; ++k) {            // stable bucket copy into the scratch column
            const unsigned e = in(k);
            const unsigned r = e & 7u;
            tmp((int)((fill >> (8 * r)) & 0xffull)) = (unsigned short)e;
            fill += 1ull << (8 * r);
        }
        next = startp;
        rem = cnt;
    }

    template <class Tmp>
    SPH_LO_HD unsigned pull(int k, Tmp tmp) {
        unsigned r = (unsigned)(q + k) & 7u;
        if (((rem >> (8 * r)) & 0xffull) == 0) {   // that group is exhausted: take from the fullest one
            unsigned best = 0, bc = 0;
            for (unsigned r2 = 0; r2 < 8; ++r2) {
                const unsigned c = (unsigned)((rem >> (8 * r2)) & 0xffull);
                if (c > bc) {
                    bc = c;
                    best = r2;
                }
            }
            r = best;
        }
        const int p = (int)((next >> (8 * r)) & 0xffull);
        next += 1ull << (8 * r);
        rem -= 1ull << (8 * r);
        return (unsigned)tmp(p);
    }
};

}  // namespace sph

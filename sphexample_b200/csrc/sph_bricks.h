// sph_bricks.h — how one row of cells is cut into bricks (the work units of the interaction
// kernels).  Plain C++ (host + device) so that CPU tests can drive the product's own logic
// (tests/physics_shim.cpp); used on the device by k_build_bricks (sph_cells.cuh).
#pragma once

#if defined(__CUDACC__)
#define SPH_BRICK_HD __host__ __device__ __forceinline__
#else
#define SPH_BRICK_HD inline
#endif

namespace sph {

struct Brick { int t0, t1; };   // particles [t0, t1) of the cell-sorted table, all in one row of cells

// Walk the cells of one row (cell_start[rowbase + cx], cx = 1 .. nx-2; cells 0 and nx-1 are the
// empty padding of the dense grid) and emit its bricks in order.  A brick closes
//   * when it holds `bt` particles — possibly in the middle of a cell — or
//   * before a cell whose inclusion would push the brick's candidate window beyond `wlimit`:
//     window(first, last) = sum over the NR neighbouring rows (roff[q] = their first cell) of the
//     particles in cells [first-1, last+1], i.e. what the interaction kernels stage.
// A single cell whose own window exceeds wlimit cannot be helped here (the list build notices).
template <int NR, class Emit>
SPH_BRICK_HD void walk_row_bricks(const int *cell_start, int rowbase, int nx, const int *roff, int bt, int wlimit, Emit emit) {
    const int p0 = cell_start[rowbase], p1 = cell_start[rowbase + nx];
    int t0 = p0, cfirst = -1, wlo = 0;
    for (int cx = 1; cx < nx - 1; ++cx) {
        const int cs = cell_start[rowbase + cx], ce = cell_start[rowbase + cx + 1];
        if (ce <= cs) continue;
        if (cfirst < 0) {
            cfirst = cx;
            wlo = 0;
            for (int q = 0; q < NR; ++q) wlo += cell_start[roff[q] + cx - 1];
        } else {
            int whi = 0;
            for (int q = 0; q < NR; ++q) whi += cell_start[roff[q] + cx + 2];
            if (whi - wlo > wlimit && cs > t0) {   // close before this cell
                emit(t0, cs);
                t0 = cs;
                cfirst = cx;
                wlo = 0;
                for (int q = 0; q < NR; ++q) wlo += cell_start[roff[q] + cx - 1];
            }
        }
        while (ce - t0 >= bt) {                     // full bricks, possibly ending mid-cell
            emit(t0, t0 + bt);
            t0 += bt;
            cfirst = (t0 < ce) ? cx : -1;          // the next brick starts inside this cell, or afresh
            if (cfirst >= 0) {
                wlo = 0;
                for (int q = 0; q < NR; ++q) wlo += cell_start[roff[q] + cx - 1];
            }
        }
    }
    if (t0 < p1) emit(t0, p1);
}

}  // namespace sph

"""The product's brick builder (csrc/sph_bricks.h, compiled for the host through the shim): every
owned row's particle range is cut into an exact partition of bricks that respect the size and the
candidate-window limits.  A hole or an overlap here would silently drop or double-count particles."""
import ctypes as C

import numpy as np
import pytest

from test_physics_host import shim  # noqa: F401  (fixture)


def _grid(rng, dim, nx, nm, ns, fill, dense):
    """random cell occupancy on a padded dense grid; returns cell_start (len ncell + 1)"""
    shape = (ns, nm, nx)
    cnt = rng.poisson(dense, size=shape) * (rng.random(shape) < fill)
    cnt[:, :, 0] = cnt[:, :, -1] = 0           # x padding
    cnt[0] = cnt[-1] = 0                       # s padding
    if dim == 3:
        cnt[:, 0] = cnt[:, -1] = 0             # m padding
    start = np.concatenate([[0], np.cumsum(cnt.ravel())]).astype(np.int32)
    return cnt, start


def _window(start, nx, nm, dim, r, c0, c1):
    rows = [(ds * nm + dm) for ds in (-1, 0, 1) for dm in ((-1, 0, 1) if dim == 3 else (0,))]
    return sum(int(start[(r + o) * nx + c1 + 2] - start[(r + o) * nx + c0 - 1]) for o in rows)


@pytest.mark.parametrize("dim,nx,nm,ns,fill,dense,bt,wlimit", [
    (3, 40, 6, 6, 0.9, 40, 128, 1900), (3, 60, 5, 5, 0.15, 12, 128, 1900), (3, 30, 5, 5, 1.0, 64, 128, 600),
    (2, 80, 1, 8, 0.9, 16, 128, 400), (2, 200, 1, 5, 0.05, 16, 128, 400), (3, 25, 4, 4, 0.5, 3, 128, 10 ** 9),
])
def test_bricks_partition_every_row(shim, dim, nx, nm, ns, fill, dense, bt, wlimit):  # noqa: F811
    rng = np.random.default_rng(dim * 1000 + nx)
    cnt, start = _grid(rng, dim, nx, nm, ns, fill, dense)
    shim.shim_row_bricks.argtypes = [C.c_void_p] + [C.c_int] * 6 + [C.c_void_p, C.c_int]
    shim.shim_row_bricks.restype = C.c_int
    out = np.zeros(2 * 4096, np.int32)
    nrows = nm * ns
    seen = 0
    for r in range(nm, nrows - nm):             # owned rows: not the s padding layers
        if dim == 3 and (r % nm in (0, nm - 1)):
            continue                            # m padding rows are empty anyway
        nb = shim.shim_row_bricks(start.ctypes.data, nx, nm, dim, r, bt, wlimit, out.ctypes.data, 4096)
        p0, p1 = int(start[r * nx]), int(start[(r + 1) * nx])
        if p1 == p0:
            assert nb == 0
            continue
        b = out[:2 * nb].reshape(nb, 2)
        assert b[0, 0] == p0 and b[-1, 1] == p1                    # covers the row ...
        assert np.array_equal(b[1:, 0], b[:-1, 1])                 # ... without holes or overlaps
        sizes = b[:, 1] - b[:, 0]
        assert sizes.min() >= 1 and sizes.max() <= bt
        # cell of a sorted index within this row
        cell_of = lambda t: int(np.searchsorted(start[r * nx:(r + 1) * nx + 1], t, side="right") - 1)
        for t0, t1 in b:
            c0, c1 = cell_of(t0), cell_of(t1 - 1)
            w = _window(start, nx, nm, dim, r, c0, c1)
            if c1 > c0:
                # multi-cell bricks respect the window limit, unless dropping the last cell would not have
                # helped either (the brick was forced to start where a full brick ended)
                assert w <= wlimit or _window(start, nx, nm, dim, r, c0, c0) > wlimit or t1 - t0 == bt
        seen += int(sizes.sum())
    assert seen == int(cnt[1:-1].sum())


def test_huge_window_limit_gives_ceil_count_over_bt(shim):  # noqa: F811
    rng = np.random.default_rng(5)
    cnt, start = _grid(rng, 3, 30, 5, 5, 0.8, 20)
    shim.shim_row_bricks.argtypes = [C.c_void_p] + [C.c_int] * 6 + [C.c_void_p, C.c_int]
    shim.shim_row_bricks.restype = C.c_int
    out = np.zeros(2 * 1024, np.int32)
    for r in range(5, 20):
        n_row = int(start[(r + 1) * 30] - start[r * 30])
        nb = shim.shim_row_bricks(start.ctypes.data, 30, 5, 3, r, 128, 10 ** 9, out.ctypes.data, 1024)
        assert nb == -(-n_row // 128)


def test_sparse_row_is_cut_at_the_gap(shim):  # noqa: F811
    """two tank walls 100 cells apart in one row next to a dense row: one brick would stage the whole
    neighbouring row; the window rule cuts it in two"""
    nx, nm, ns = 110, 3, 3
    cnt = np.zeros((ns, nm, nx), np.int64)
    cnt[1, 1, 2] = 12
    cnt[1, 1, 105] = 12
    cnt[1, 1, 3:105] = 0
    start = np.concatenate([[0], np.cumsum(cnt.ravel())]).astype(np.int32)
    shim.shim_row_bricks.argtypes = [C.c_void_p] + [C.c_int] * 6 + [C.c_void_p, C.c_int]
    shim.shim_row_bricks.restype = C.c_int
    out = np.zeros(64, np.int32)
    r = 1 * nm + 1
    assert shim.shim_row_bricks(start.ctypes.data, nx, nm, 3, r, 128, 2000, out.ctypes.data, 32) == 1   # empty surroundings: fine
    cnt[1, 1, 3:105] = 0
    cnt[1, 2, 2:106] = 0
    cnt[2, 1, 2:106] = 40                                                                              # a dense row above
    start = np.concatenate([[0], np.cumsum(cnt.ravel())]).astype(np.int32)
    nb = shim.shim_row_bricks(start.ctypes.data, nx, nm, 3, r, 128, 2000, out.ctypes.data, 32)
    assert nb == 2 and out[1] == out[2]

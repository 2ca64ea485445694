"""Host logic of the work-list IDWT (trinerflet_b200/idwt_plan.py): the per-level block maps must be closed under the
data dependencies of the level kernels, and the run lists must tile every level exactly once."""
import numpy as np
import torch

from trinerflet_b200.idwt_plan import MAX_ACTIVE_RUN, MAX_CLEAN_RUN, _items, level_maps


def _flags(T, seed, density):
    g = torch.Generator().manual_seed(seed)
    f = torch.rand(3, T, T, generator=g) < density
    f[:, T // 4: T // 2, T // 3: T // 2] = True      # a blob, like an object in the middle
    return f


def _cover(items, nb):
    cov = np.zeros((3, nb, nb), dtype=np.int32)
    for p, m0, lo, hi in items:
        assert m0 % 16 == 0 and lo % 16 == 0 and hi % 16 == 0 and hi > lo
        cov[p, lo // 16: hi // 16, m0 // 16] += 1
    return cov


def test_items_tile_each_level_once():
    for T, levels in ((64, 5), (16, 3), (4, 1)):
        fwd, bwd = level_maps(_flags(T, T, 0.05), levels)
        for maps in (fwd, bwd):
            for l, m in enumerate(maps):
                nb = T >> (levels - 1 - l)
                m = m.numpy()
                assert m.shape == (3, nb, nb)
                act, cln = _items(m, True, MAX_ACTIVE_RUN), _items(m, False, MAX_CLEAN_RUN)
                assert np.array_equal(_cover(act, nb), m.astype(np.int32))
                assert np.array_equal(_cover(cln, nb), (~m).astype(np.int32))
                assert all((hi - lo) // 16 <= MAX_ACTIVE_RUN for _, _, lo, hi in act)
                assert all((hi - lo) // 16 <= MAX_CLEAN_RUN for _, _, lo, hi in cln)


def test_maps_are_closed_under_the_kernel_halos():
    T, levels = 32, 4
    f = _flags(T, 1, 0.03)
    fwd, bwd = level_maps(f, levels)
    assert torch.equal(fwd[-1], f)                                     # top level reconstructs exactly the marked tiles
    for l in range(levels - 1, 0, -1):
        up, lo = fwd[l].numpy(), fwd[l - 1].numpy()
        nb = up.shape[1]
        for p, by, bx in zip(*np.nonzero(up)):
            # an active block of level l reads level l-1 output pixels [16b-4, 16b+20) per axis
            for y in range(max(16 * by - 4, 0), min(16 * by + 20, 16 * nb)):
                for x in (max(16 * bx - 4, 0), min(16 * bx + 19, 16 * nb - 1)):
                    assert lo[p, y // 32, x // 32]
    # backward: a block is active iff a tile with incoming gradient intersects its 8-pixel input halo; its own g_x
    # then counts as incoming gradient of the level below
    g = f.numpy()
    for l in range(levels - 1, -1, -1):
        b = bwd[l].numpy()
        nb = b.shape[1]
        for p in range(3):
            for by in range(nb):
                for bx in range(nb):
                    lo_y, hi_y = max(32 * by - 8, 0) // 32, min(32 * by + 39, 32 * nb - 1) // 32
                    lo_x, hi_x = max(32 * bx - 8, 0) // 32, min(32 * bx + 39, 32 * nb - 1) // 32
                    assert b[p, by, bx] == g[p, lo_y:hi_y + 1, lo_x:hi_x + 1].any()
        g = b.reshape(3, nb // 2, 2, nb // 2, 2).any(axis=(2, 4)) if nb > 1 else b

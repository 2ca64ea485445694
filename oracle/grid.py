"""TEST INFRASTRUCTURE ONLY (oracle) -- CPU restatement of the density-grid refresh
NeRFRenderer.update_extra_state (/root/reference/reconstruction/nerf/renderer.py:448-542) with the random draws made explicit:

  full sweep    (iter_density < 16, :460-488): every cell of the 128^3 grid in custom_meshgrid('ij') order, per cascade
                jittered cell centres  xyz = (2*coord/(H-1) - 1) * (bound_c - hgs) + (2*noise - 1) * hgs,  hgs = bound_c / H
  partial sweep (:492-518): per cascade  coords = randint(0, H, (N, 3));  occupied = nonzero(grid[cas] > 0);
                pick = randint(0, len(occupied), [N]);  cells = cat(morton(coords), occupied[pick]);  noise = rand(2N, 3)
  then          tmp_grid[cas, cells] = sigma * density_scale  (duplicates: the LAST writer wins here; a race in the reference),
                grid = max(grid * decay, tmp) where both >= 0 (:526-527), mean = mean(clamp(grid, 0)) (:528),
                thresh = min(mean, density_thresh) (:533), bitfield = packbits(grid > thresh) (:534),
                mean_count = int(sum(step_counter[:min(16, local_step), 0]) / total_step) (:537-540).

The caller supplies the draws (tests replay the product's torch.rand / torch.randint call sequence on the same seeded
generator) and the density function (oracle/field.py).  Pinned by tests/test_reference_host_over_dropin.py, where the
reference's own update_extra_state runs over the drop-in modules, and by the packbits / Morton known-answers of
tests/test_oracle_raymarch.py.  Never imported by the product package."""
import numpy as np
import torch

from . import raymarch as orc


def all_cells(H=128):
    a = torch.arange(H, dtype=torch.int32)
    xx, yy, zz = torch.meshgrid(a, a, a, indexing="ij")
    coords = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=-1)
    return coords, torch.from_numpy(orc.morton3D(coords.numpy())).long()


def cell_positions(coords, cas, bound, H, noise):
    """renderer.py:474-483 with torch's CUDA scalar rule (tensor / python-scalar == tensor * fp32(1/scalar))"""
    bound_c = min(2 ** cas, bound)
    hgs = bound_c / H
    xyzs = 2 * coords.float() * torch.tensor(1.0 / (H - 1), dtype=torch.float32) - 1
    return xyzs * (bound_c - hgs) + (noise * 2 - 1) * hgs


def update_extra_state(density_grid, iter_density, density_fn, draws, bound=1.5, H=128, decay=0.95, density_thresh=10.0,
                       density_scale=1.0, step_counter=None, local_step=0):
    """density_grid [cascade, H^3] fp32 (Morton order, not modified); density_fn(xyz [n,3]) -> sigma [n];
    draws: full sweep  -> list over cascades of noise [H^3, 3];
           partial     -> list over cascades of (coords [N,3] int64, pick [N] int64 or None, noise [n,3]).
    Returns dict(grid, mean_density, bitfield (uint8 numpy), mean_count, sampled (bool [cascade, H^3]), duplicated (same))."""
    cascade = density_grid.shape[0]
    grid = density_grid.clone()
    tmp = -torch.ones_like(grid)
    sampled = torch.zeros_like(grid, dtype=torch.bool)
    dup = torch.zeros_like(grid, dtype=torch.bool)
    if iter_density < 16:
        coords, indices = all_cells(H)
        for cas in range(cascade):
            sig = density_fn(cell_positions(coords, cas, bound, H, draws[cas])).reshape(-1).float() * density_scale
            tmp[cas, indices] = sig
            sampled[cas] = True
    else:
        for cas in range(cascade):
            coords, pick, noise = draws[cas]
            indices = torch.from_numpy(orc.morton3D(coords.numpy().astype(np.int32))).long()
            occ = torch.nonzero(grid[cas] > 0).squeeze(-1)
            if occ.shape[0] > 0:
                occ_idx = occ[pick]
                occ_coords = torch.from_numpy(orc.morton3D_invert(occ_idx.numpy().astype(np.int32))).long()
                indices = torch.cat([indices, occ_idx], dim=0)
                coords = torch.cat([coords, occ_coords], dim=0)
            sig = density_fn(cell_positions(coords, cas, bound, H, noise)).reshape(-1).float() * density_scale
            row = tmp[cas].numpy()
            row[indices.numpy()] = sig.numpy()                     # numpy fancy assignment: the last duplicate wins
            cnt = torch.bincount(indices, minlength=grid.shape[1])
            sampled[cas] = cnt > 0
            dup[cas] = cnt > 1
    valid = (grid >= 0) & (tmp >= 0)
    grid[valid] = torch.maximum(grid[valid] * decay, tmp[valid])
    mean = float(torch.mean(grid.clamp(min=0)))
    thresh = min(mean, density_thresh)
    bits = np.packbits(grid.reshape(-1).numpy() > np.float32(thresh), bitorder="little")
    total_step = min(16, local_step)
    mean_count = int(step_counter[:total_step, 0].sum().item() / total_step) if total_step > 0 else None
    return dict(grid=grid, mean_density=mean, thresh=thresh, bitfield=bits, mean_count=mean_count, sampled=sampled, duplicated=dup)

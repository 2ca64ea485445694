"""torchrun --nproc-per-node N profiles/check_render_sharded.py [max_steps]
Full-frame 800x800 render of the base config sharded into round-robin ray tiles over N GPUs (parallel.render_frame_sharded:
replicated planes, device-driven marching loop, no collective until the final NCCL gather) against the same
frame rendered on rank 0 alone.  Prints one JSON line on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trinerflet_b200 import parallel, scene  # noqa: E402
from trinerflet_b200.network import NeRFNetwork  # noqa: E402


def main():
    max_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    cfg = scene.CONFIGS["base_light"]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    net = NeRFNetwork(bound=1.5, cuda_ray=True, density_thresh=10, min_near=0.2, triplane_channels=cfg["C"], triplane_resolution=cfg["R"],
                      triplane_wavelet_levels=cfg["S"]).to(dev)
    scene.init_model_(net, seed=0)
    scene.install_ball_occupancy(net, 0.75)
    net.eval()
    net.infer_chunk = 8
    sc = scene.make_scene()
    ro, rd = (t.to(dev) for t in scene.full_frame(sc, 3))
    kw = dict(bg_color=1, max_steps=max_steps, dt_gamma=0)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        out = parallel.render_frame_sharded(net, ro, rd, rank, world, **kw)          # warm-up (+ planes)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            out = parallel.render_frame_sharded(net, ro, rd, rank, world, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 3], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        res = None
        if rank == 0:
            full = net.render(ro.unsqueeze(0), rd.unsqueeze(0), staged=True, perturb=False, **kw)
            d_img = (full["image"].reshape(-1, 3) - out["image"]).abs().max().item()
            d_ws = (full["weights_sum"].reshape(-1) - out["weights_sum"]).abs().max().item()
            fin = torch.isfinite(full["depth"].reshape(-1))
            d_dep = (full["depth"].reshape(-1)[fin] - out["depth"][fin]).abs().max().item()
            res = {"what": "render_sharded_check", "world": world, "max_steps": max_steps, "rays": int(ro.shape[0]), "ms_per_frame_sharded": round(float(ms), 3),
                   "frames_per_s": round(1e3 / float(ms), 2), "max_abs_diff_image": d_img, "max_abs_diff_weights_sum": d_ws, "max_abs_diff_depth": d_dep,
                   "ok": bool(d_img <= 5e-3 and d_ws <= 5e-3), "covered": float(out["weights_sum"].sum())}
    if rank == 0:
        print(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

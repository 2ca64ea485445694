"""torchrun --nproc-per-node N profiles/check_peer_exchange.py [config]
Multi-GPU check of parallel.PeerGradExchange (own multimem / P2P all-reduce kernels over symmetric memory, captured in one CUDA
graph) against the NCCL exchange with fp32 transport on the same rays: loss and every parameter gradient must agree (both are
exact fp32 sums up to the order of the additions); then times both exchanges alone and prints one JSON line on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trinerflet_b200 import scene, trainer  # noqa: E402
from trinerflet_b200.network import NeRFNetwork  # noqa: E402


def rel_l2(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def main():
    cfg = scene.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "base_light"]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sc = scene.make_scene()
    n_rays = 8192
    ro, rd, tgt = (t.to(dev) for t in scene.sample_batch(sc, n_rays, torch.Generator().manual_seed(100 + rank)))
    res = {}
    for mode in ("nccl", "peer"):
        net = NeRFNetwork(bound=1.5, cuda_ray=True, density_thresh=10, min_near=0.2, triplane_channels=cfg["C"], triplane_resolution=cfg["R"],
                          triplane_wavelet_levels=cfg["S"], hidden_dim=cfg["hidden"], hidden_dim_color=cfg["hidden"]).to(dev)
        scene.init_model_(net, seed=0)
        scene.install_ball_occupancy(net, 0.75)
        net.train()
        ts = trainer.TrainStep(net, trainer.default_opt(), None, world_size=world, transport=torch.float32, exchange=mode)
        torch.manual_seed(7 + rank)
        loss = ts.forward_backward(ro, rd, tgt, update_grid=False)
        grads = [p.grad.detach().clone() for p in net.parameters()]
        # steady state + graph replay
        net.mean_count = int(net.step_counter[0, 0].item())
        net.local_step = 0
        net.zero_grad(set_to_none=True)
        ts.capture(ro, rd, tgt, warmup=1)
        torch.manual_seed(7 + rank)
        loss_g = ts.replay(ro, rd, tgt)
        grads_g = [p.grad.detach().clone() for p in net.parameters()]
        # the exchange alone, timed
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        extra = {}
        if mode == "peer":
            from trinerflet_b200._lib import call, ptr, stream
            ex, r = ts.exch, ts.reducer

            def timed(fn, n=reps):
                fn(); torch.cuda.synchronize(); dist.barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(n):
                    fn()
                b.record(); torch.cuda.synchronize()
                return round(a.elapsed_time(b) / n, 4)

            extra["ms_two_barriers"] = timed(lambda: (ex._hdl_p.barrier(channel=0), ex._hdl_p.barrier(channel=1)))
            for mc in ([True, False] if ex._mc_p is not None else [False]):
                def tiles(mc=mc):
                    ex._hdl_p.barrier(channel=0)
                    call("tnl_tiles_allreduce", ex._mc_p if mc else None, ex._arr_p, ptr(r.list_ids), ptr(r.list_count), r.list_cap, ex.R, ex.C, ex.T,
                         ex.rank, ex.world, 1.0 / ex.world, stream())
                    ex._hdl_p.barrier(channel=1)
                extra["ms_tiles_allreduce_" + ("multimem" if mc else "p2p")] = timed(tiles)
            plain = torch.zeros_like(ex._flat_planes)
            extra["ms_rmw_1.6GB_symmetric_buffer"] = timed(lambda: ex._flat_planes.add_(1.0), 5)
            extra["ms_rmw_1.6GB_ordinary_buffer"] = timed(lambda: plain.add_(1.0), 5)
            del plain
            extra["dirty_tiles"] = r.n_tiles
            extra["dirty_MB_fp32"] = round(r.n_tiles * ex.T * ex.T * ex.C * 4 / 1e6, 1)
            e0.record()
            for _ in range(reps):
                ts.exch.exchange_()
            e1.record()
        else:
            g = torch.zeros_like(net.encoder.planes_features.new_empty(0)) if False else None
            buf = torch.zeros(3, cfg["R"], cfg["R"], cfg["C"], device=dev).permute(0, 3, 1, 2)
            ts.reducer.transport, ts.reducer.bf16 = torch.bfloat16, 1
            ts.reducer.refresh()
            e0.record()
            for _ in range(reps):
                ts.reducer.reduce_(buf)
            e1.record()
        torch.cuda.synchronize()
        res[mode] = dict(loss=float(loss), loss_graph=float(loss_g), grads=grads, grads_graph=grads_g, ms=e0.elapsed_time(e1) / reps,
                         note=ts.exchange_note, graphs=1 if ts._graphs[1] is None else 2, extra=extra)
        del ts, net
        torch.cuda.empty_cache()
    worst = max(rel_l2(a, b) for a, b in zip(res["peer"]["grads"], res["nccl"]["grads"]))
    worst_g = max(rel_l2(a, b) for a, b in zip(res["peer"]["grads_graph"], res["peer"]["grads"]))
    ok = worst <= 2e-5 and worst_g <= 2e-5 and abs(res["peer"]["loss"] - res["nccl"]["loss"]) <= 1e-6 * abs(res["nccl"]["loss"])
    t = torch.tensor([float(ok)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"what": "peer_exchange_check", "world": world, "ok_all_ranks": bool(float(t) == 1.0), "max_rel_l2_peer_vs_nccl_fp32": worst,
                          "max_rel_l2_graph_vs_eager": worst_g, "loss_peer": res["peer"]["loss"], "loss_nccl": res["nccl"]["loss"],
                          "exchange_ms_peer_fp32_in_place": round(res["peer"]["ms"], 4), "exchange_ms_nccl_bf16_pack_unpack": round(res["nccl"]["ms"], 4),
                          "peer": res["peer"]["note"], "graphs_per_step_peer": res["peer"]["graphs"], "graphs_per_step_nccl": res["nccl"]["graphs"], "detail": res["peer"]["extra"]}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

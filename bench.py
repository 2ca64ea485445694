#!/usr/bin/env python
"""Benchmark of the TriNeRFLet reconstruction hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config base_light] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): rays/s per training step (fwd+bwd): one step = encoder.get_planes() (multilevel IDWT,
outside autocast) + render of N rays (near/far, march, tri-plane sampling, sigma/color MLP, composite) + MSE +
wavelet L1 regulariser + backward down to the coefficient / MLP gradients (+ NCCL gradient all-reduce at N > 1),
in the order of reconstruction/nerf/utils.py:1138-1166.  The optimizer step and the density-grid refresh are
outside the metric and reported separately under "extras".

`--impl reference` times the reference's CPU implementation of the same path (the oracle port; pytorch_wavelets and
the CUDA-only extensions cannot run on a CPU here) on the host cores, on bounded samples of the same workload.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "rays/sec per training step (fwd+bwd)"
UNIT = "rays/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="base_light")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=0, help="rays per GPU per step (default: the config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-next-rows", action="store_true", help="skip the extras.next_rows measurements (feeder, render loops)")
    ap.add_argument("--occupancy-radius", type=float, default=0.75,
                    help="radius of the occupied ball (SURVEY.md 8d: 0.75 = the metric's scene, 0.4 = the sparse preset that exposes the plane-bound regime)")
    ap.add_argument("--tiled-sampling", action="store_true",
                    help="run the whole benchmark with the opt-in tile-binned sampling kernels (encoder.tiled_sampling)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying the captured CUDA graph")
    ap.add_argument("--cpu-budget", type=float, default=25.0, help="seconds of CPU work for the cpu_baseline leg")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed regions (NVML, every 2 ms in a thread; a timed region of a few
    steps is shorter than nvidia-smi's start-up time, so the CLI loop of the profiling recipe is only the fallback)."""
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu, self.sm, self.mask, self.max_mhz = gpu_index, [], 0, None
        self.h = self.nv = self.proc = None
        self.active = False
        self.stop_flag = False
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[gpu_index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            if self.active:
                try:
                    self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                except Exception:
                    pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self.active = True
            return
        try:
            self.rows = []
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.rows.append([c.strip() for c in ln.split(",")]) for ln in self.proc.stdout],
                             daemon=True).start()
        except Exception:
            self.proc = None

    def pause(self):
        self.active = False

    def stop(self):
        self.active = False
        self.stop_flag = True
        if self.nv is not None:
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(n for b, n in self.REASONS if self.mask & b), "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ---------------------------------------------------------------------------------------------------------
# algorithmic bytes per ABI call, from the scalar arguments the profiler hook records
# ---------------------------------------------------------------------------------------------------------
def algorithmic_bytes(name, meta, C, m_valid, plan=None):
    g = 12 * C * 4
    if name in ("tnl_idwt_level_forward", "tnl_idwt_level_backward"):
        n, c = meta[0], meta[1]
        return 2 * (3 * c * (2 * n) ** 2 * 4)            # read all coefficients of the level + write its planes (= 2 P_level)
    if name in ("tnl_idwt_level_forward_sparse", "tnl_idwt_level_backward_sparse") and plan is not None:
        # work-list mode: px = one (plane, channel) layer of n x n coefficients.  Forward reads x + 3 bands on the active
        # blocks, the 3 bands only (|yh| sum) on the clean ones, writes 4 px of planes per active coefficient position;
        # backward reads the 4 px of incoming gradient on the active blocks + the 3 bands everywhere (regulariser sign),
        # writes all 4 gradient layers.
        n, c = meta[0], meta[1]
        lvl = int(round(math.log2(n / plan.n0)))
        px = 3 * c * n * n * 4
        a = plan.stats["active_fraction_forward"][lvl]
        b = plan.stats["active_fraction_backward"][lvl]
        parts = int(meta[-1])            # bit 0: active blocks, bit 1: clean blocks (they may be separate calls)
        if name.endswith("forward_sparse"):
            # the |yh| pass covers all non-reconstructed blocks, or (regulariser value completed by the backward's clean
            # part, plan.defer_clean_abs) only those the backward treats as active
            clean = (b - a) if plan.defer_clean_abs else (1 - a)
            return px * (((parts & 1) and a * (4 + 4)) + ((parts & 2) and clean * 3))
        return px * (((parts & 1) and b * (4 + 3 + 4)) + ((parts & 2) and (1 - b) * (3 + 4)))
    if name in ("tnl_sample_planes_forward", "tnl_tsample_forward"):       # the algorithm's bytes, whatever the
        return m_valid * (12 + g)                                            # kernel's on-chip reuse makes of them
    if name in ("tnl_sample_planes_backward", "tnl_tsample_backward"):
        return m_valid * 2 * g
    if name == "tnl_mlp_forward":
        return m_valid * (2 * 3 * C + 28)            # fp16 feature row + dirs in, sigma + rgb out
    if name == "tnl_mlp_backward":
        return m_valid * (2 * 2 * 3 * C + 28)        # feature row in, feature-gradient row out (fp16), dirs, g_sigma, g_rgb
    if name == "tnl_composite_rays_train_forward":
        return m_valid * 24 + meta[1] * 32
    if name == "tnl_composite_rays_train_backward":
        return m_valid * 40 + meta[1] * 32
    if name == "tnl_march_rays_train":
        return meta[3] * 36 + m_valid * 32
    return None


def step_bytes(P, C, M, N):
    return 5 * P + M * (3 * 12 * C * 4 + 152) + N * 120


# ---------------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port)
# ---------------------------------------------------------------------------------------------------------
def cpu_reference(cfg, n_rays, budget_s, seed=0):
    import torch
    from oracle import pipeline
    from trinerflet_b200 import scene
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sc = scene.make_scene()
    grid = scene.ball_density_grid(1.5, 0.75)
    bits = scene.packbits_cpu(grid, 0.5).numpy()
    n_large = 1024 if budget_s >= 10 else 384
    n_small = n_large // 4
    batch = scene.sample_batch(sc, n_large, torch.Generator().manual_seed(seed))
    # IDWT sample: full IDWT of base-light costs ~11 s on 8 cores; 1 of 3 planes, and a halved resolution on small budgets
    P_GB = 3 * cfg["C"] * cfg["R"] ** 2 * 4 / 1e9
    est_full = 7.0 * P_GB * (8.0 / cores)
    planes_sub = 3 if est_full < 0.4 * budget_s else 1
    r_div = 1
    while est_full * planes_sub / 3 / r_div ** 2 > 0.5 * budget_s and r_div < 4:
        r_div *= 2
    r = pipeline.timed_components(cfg["C"], cfg["R"], cfg["S"], cfg["hidden"], n_rays, batch, bits, planes_sub, r_div, n_small, n_large, seed)
    r["cores"] = cores
    r["value"] = n_rays / r["step_seconds"]
    r["sample"] = (f"IDWT fwd+bwd on {planes_sub}/3 planes at R={cfg['R'] // r_div} (scaled x{3 / planes_sub * r_div ** 2:g}: planes are "
                   f"independent, cost ~ pixels) + march/sample/MLP/composite fwd+bwd on {n_small} and {n_large} rays against "
                   f"full-size planes, linear fit extrapolated to {n_rays} rays")
    return r


def run_reference(args, cfg, n_rays):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step_budget = max(2.0, min(12.0, 150.0 / max(1, args.steps + args.warmup)))
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference(cfg, n_rays, per_step_budget, seed=i)
        if i >= args.warmup:
            vals.append(r)
    step_s = sum(v["step_seconds"] for v in vals) / len(vals)
    value = n_rays / step_s
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: C={cfg['C']} R={cfg['R']} levels={cfg['S']} rays={n_rays} (CPU, bounded sample extrapolated)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": vals[-1]["cores"], "kind": "port", "sample": vals[-1]["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def next_rows(args, net, ts, sc, n_rays, dev, use_graph, devb=None):
    """Step feeder (rays + targets generated on the device into the captured step's inputs), full-frame inference with the
    host-driven / device-driven loop, and the A/B of the opt-in tile-binned sampling kernels.  Reported under extras only."""
    import torch
    from trinerflet_b200 import rays, scene, trainer
    out = {}
    try:
        # A/B: the same captured step with encoder.tiled_sampling (csrc/tsample.cu) instead of the point-ordered kernels
        if use_graph and devb is not None:
            enc = net.encoder
            was = enc.tiled_sampling
            enc.tiled_sampling = not was
            try:
                ts3 = trainer.TrainStep(net, ts.opt, optimizer=None, world_size=1)
                net.zero_grad(set_to_none=True)
                ts3.forward_backward(*devb[-1], update_grid=False)
                ts3.capture(*devb[-2], warmup=1)
                for i in range(3):
                    ts3.replay(*devb[i])
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(args.steps):
                    ts3.replay(*devb[args.warmup + i])
                e1.record()
                torch.cuda.synchronize()
                key = "point_ordered_sampling_ms_per_step" if was else "tiled_sampling_ms_per_step"
                out[key] = round(e0.elapsed_time(e1) / args.steps, 4)
                out["sampling_ab_note"] = "the same captured fwd+bwd step with the other sampling kernels (tile-binned <-> point-ordered); compare with ms_per_step"
                del ts3
            finally:
                enc.tiled_sampling = was
    except Exception as ex:  # pragma: no cover
        out["tiled_sampling_ms_per_step"] = f"failed: {ex}"
    try:
        if use_graph:
            H, W = scene.H_IMG, scene.W_IMG
            images = torch.rand(sc.poses.shape[0], H * W, 3, device=dev)          # resident targets (768 MB fp32)
            feeder = rays.RayFeeder(sc.poses.to(dev), sc.intrinsics, H, W, images, seed=0).shuffle()
            for b in range(3):
                ts.replay_from_feeder(feeder, b, n_rays).item()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for b in range(args.steps):
                ts.replay_from_feeder(feeder, 3 + b, n_rays).item()
            e1.record()
            torch.cuda.synchronize()
            out["e2e_device_feeder_rays_per_s"] = n_rays / (e0.elapsed_time(e1) / args.steps * 1e-3)
            out["e2e_device_feeder_note"] = "rays_o / rays_d / targets generated by tnl_rays_from_ids into the graph's static inputs, loss read back every step; 0 H2D bytes per step"
            del feeder, images
    except Exception as ex:  # pragma: no cover
        out["e2e_device_feeder_rays_per_s"] = f"failed: {ex}"
    try:
        net.eval()
        ro, rd = scene.full_frame(sc, 0)
        ro, rd = ro.to(dev), rd.to(dev)
        ref = None
        for chunk in (0, 8):
            net.infer_chunk = chunk

            def frame():
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                    return net.render(ro.unsqueeze(0), rd.unsqueeze(0), staged=True, bg_color=1, perturb=False, max_steps=1024)

            img = frame()["image"]
            ref = img if ref is None else ref
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                frame()
            e1.record()
            torch.cuda.synchronize()
            key = "host_loop" if chunk == 0 else f"device_loop_chunk{chunk}"
            out[f"render_800x800_ms_{key}"] = round(e0.elapsed_time(e1) / 2, 3)
            if chunk:
                out["render_max_abs_diff_between_loops"] = float((img - ref).abs().max())
                out["render_iterations"] = net.last_infer_loop.iterations_done
                out["render_state_reads"] = net.last_infer_loop.reads
    except Exception as ex:  # pragma: no cover
        out["render"] = f"failed: {ex}"
    finally:
        net.infer_chunk = 0
        net.train()
    return out


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    import torch
    from trinerflet_b200 import scene
    cfg = scene.CONFIGS[args.config]
    n_rays = args.rays or cfg["rays"]
    if args.impl == "reference":
        run_reference(args, cfg, n_rays)
        return

    import torch.distributed as dist
    from trinerflet_b200 import _lib, trainer
    from trinerflet_b200.network import NeRFNetwork

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    _lib.load()

    C, R, S = cfg["C"], cfg["R"], cfg["S"]
    net = NeRFNetwork(bound=1.5, cuda_ray=True, density_thresh=10, min_near=0.2, triplane_channels=C, triplane_resolution=R,
                      triplane_wavelet_levels=S, hidden_dim=cfg["hidden"], hidden_dim_color=cfg["hidden"]).to(dev)
    scene.init_model_(net, seed=0)                      # identical replicas on every rank
    scene.install_ball_occupancy(net, args.occupancy_radius)
    net.encoder.tiled_sampling = bool(args.tiled_sampling)
    opt = trainer.default_opt()
    ts = trainer.TrainStep(net, opt, optimizer=None, world_size=world)
    sc = scene.make_scene()
    gen = torch.Generator().manual_seed(1234 + rank)     # each rank draws its own shard of the global batch
    total = args.warmup + args.steps
    host = [tuple(t.pin_memory() for t in scene.sample_batch(sc, n_rays, gen)) for _ in range(total + 3)]
    devb = [tuple(t.to(dev) for t in b) for b in host]
    torch.manual_seed(100 + rank)

    # establish the steady state: mean_count > 0 (no D2H sync in march_rays_train), M rounded up to 128
    net.train()
    probe = ts.forward_backward(*devb[-1], update_grid=False)
    net.mean_count = int(net.step_counter[0, 0].item())
    net.local_step = 0
    m_valid = net.mean_count
    P = 3 * C * R * R * 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def eager_step(b):
        net.zero_grad(set_to_none=True)
        return ts.forward_backward(*b, update_grid=False)

    use_graph = not args.no_graph
    if use_graph:
        c0 = _lib.launch_count
        ts.capture(*devb[-2], warmup=0)
        ts._graph_kernel_count = _lib.launch_count - c0

    def one_step(b):
        return ts.replay(*b) if use_graph else eager_step(b)

    for i in range(args.warmup):
        one_step(devb[i])
    # ---- timed region: device-resident inputs ----
    clocks = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    if rank == 0:
        clocks.start()
    launches0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        one_step(devb[args.warmup + i])
    e1.record()
    barrier()
    launches = _lib.launch_count - launches0
    if use_graph:   # replays do not pass through the Python launch counter: kernels per captured step x steps
        launches = ts._graph_kernel_count * args.steps
    if rank == 0:
        clocks.pause()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    counts = net.step_counter[:, 0].float().max().reshape(1)     # samples of the (graph-captured) step slot
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        counts /= world
    ms_per_step = float(ms) / args.steps
    m_valid = float(counts)
    value = n_rays * world / (ms_per_step * 1e-3)

    # ---- e2e: host (pinned) buffers in, loss out, every step ----
    barrier()
    if rank == 0:
        clocks.start()       # the end-to-end region is a timed region under load too
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    h2d = d2h = 0
    for i in range(args.steps):
        hb = host[args.warmup + i]
        b = hb if use_graph else tuple(t.to(dev, non_blocking=True) for t in hb)   # graph mode: H2D into the static buffers
        h2d = sum(t.numel() * t.element_size() for t in hb)
        loss = one_step(b)
        _ = loss.item()
        d2h = 4
    f1.record()
    barrier()
    clk = clocks.stop() if rank == 0 else None
    ms2 = torch.tensor([f0.elapsed_time(f1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = n_rays * world / (float(ms2) / args.steps * 1e-3)

    line = None
    if rank == 0:
        # ---- per-kernel times (CUDA events around every ABI call on the launching stream) ----
        ts.world_size = 1          # rank-0-only section: no collectives from here on
        if ts.reducer is not None:
            ts.reducer.world_size = 1
        ts.prefetch_planes = False   # per-kernel events must not overlap kernels of two streams
        _lib.profile_start()
        nprof = min(args.steps, 5)
        for i in range(nprof):
            eager_step(devb[args.warmup + i])
        prof = _lib.profile_stop()
        kernels = {}
        for name, recs in prof.items():
            t = sum(r[0] for r in recs) / nprof
            by = sum((algorithmic_bytes(name, r[1], C, m_valid, ts._plan) or 0) for r in recs) / nprof
            kernels[name] = {"ms_per_step": round(t, 4), "calls_per_step": len(recs) / nprof,
                             "algorithmic_GB_per_step": round(by / 1e9, 4),
                             "achieved_GBps": round(by / 1e9 / (t * 1e-3), 1) if t > 0 and by else None}
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        dom = max((k for k in kernels if kernels[k]["achieved_GBps"]), key=lambda k: kernels[k]["ms_per_step"])
        # DRAM bytes of the dominant kernel measured by ncu (dram__bytes_read.sum + dram__bytes_write.sum, one --set full
        # capture of a base-light step, summarised under profiles/), per launch like `achieved`
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
            if tr and args.config == tr.get("config"):
                traffic = round(tr["dram_bytes_per_step"] / kernels[dom]["calls_per_step"] / 1e9, 4)
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_GBps"], "peak": peak, "unit": "GB/s",
                    "frac": round(kernels[dom]["achieved_GBps"] / peak, 4), "traffic": traffic, "traffic_unit": "GB/launch",
                    "algorithmic_GB_per_launch": round(kernels[dom]["algorithmic_GB_per_step"] / kernels[dom]["calls_per_step"], 4),
                    "peak_source": peak_src,
                    "launch_ms": round(kernels[dom]["ms_per_step"] / kernels[dom]["calls_per_step"], 4)}
        # B_step: what the reference's (dense) data flow has to move per step; the work-list IDWT moves less
        # plane data than that, so this fraction is "dense-equivalent" throughput, not DRAM utilisation
        Bstep = step_bytes(P, C, m_valid, n_rays)
        extras = {"sparse_allreduce_tile_fraction": (round(ts.reducer.fraction, 4) if ts.reducer is not None else None),
                  "idwt_worklist": (ts._plan.stats if ts._plan is not None else None),
                  "M_samples_per_step": m_valid, "B_step_GB": round(Bstep / 1e9, 3),
                  "step_achieved_GBps": round(Bstep / 1e9 / (ms_per_step * 1e-3), 1),
                  "step_frac_of_hbm_roofline": round(Bstep / 1e9 / (ms_per_step * 1e-3) / peak, 4), "kernels": kernels}
        # optimizer + density-grid refresh, outside the metric
        try:
            optim = trainer.make_optimizer(net, 1e-2)
            ts2 = trainer.TrainStep(net, opt, optimizer=optim, world_size=1)
            ts2.step(*devb[0], update_grid=False)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(3):
                ts2.step(*devb[i], update_grid=False)
            torch.cuda.synchronize()
            extras["ms_per_step_with_optimizer"] = round((time.perf_counter() - t0) / 3 * 1e3, 3)
            del optim, ts2
            torch.cuda.empty_cache()
            # the same with the fused epilogue (trinerflet_b200/optim.py: unscale + check + Adam + scale update, 2 kernels)
            optim = trainer.make_optimizer(net, 1e-2, fused=True)
            ts2 = trainer.TrainStep(net, opt, optimizer=optim, world_size=1)
            ts2.step(*devb[0], update_grid=False)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(3):
                ts2.step(*devb[i], update_grid=False)
            torch.cuda.synchronize()
            extras["ms_per_step_with_fused_optimizer"] = round((time.perf_counter() - t0) / 3 * 1e3, 3)
            del optim, ts2
        except Exception as ex:  # pragma: no cover
            extras["ms_per_step_with_optimizer"] = f"failed: {ex}"
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference(cfg, n_rays, args.cpu_budget)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
                   "step_seconds": round(r["step_seconds"], 3), "idwt_seconds": round(r["idwt_seconds"], 3),
                   "rays_seconds": round(r["rays_seconds"], 3)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic",
            "config": {"workload": f"{args.config}: C={C} R={R} wavelet_levels={S} ({int(round(__import__('math').log2(S)))} IDWT levels), "
                                   f"{n_rays} rays/GPU/step, synthetic 800x800 Blender-shaped scene, ball occupancy r={args.occupancy_radius:g}, random-init",
                       "sampling_kernels": "tile-binned (csrc/tsample.cu, --tiled-sampling)" if args.tiled_sampling else "point-ordered (csrc/sample.cu)",
                       "rays_per_gpu": n_rays, "global_rays": n_rays * world, "parallelism": (f"ray-sharded dp{world}, replicated coefficients; plane gradient exchanged as bf16 dirty tiles (NCCL all-reduce) "
                                                       "between the render backward and the IDWT backward" if world > 1 else "single GPU"),
                       "timed_region": "get_planes (work-list IDWT over the occupied tiles) + render + loss + backward (+ gradient exchange); optimizer and density-grid refresh excluded (metric definition), see extras",
                       "launch": ("one CUDA-graph replay per step" if world == 1 else "two CUDA-graph replays per step around the NCCL exchange") if use_graph else "eager Python launches",
                       "l2": "inputs (1.6 GB of coefficients/planes per pass) exceed the 126 MB L2; a different ray batch every step"},
            "clocks": clk, "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "extras": extras,
        }
        # ---- the two "next" rows, outside the metric; N = 1 only.  Everything the contract line needs
        # has been computed above: a failure in here can only turn into a "failed: ..." string inside extras ----
        if world == 1 and not args.no_next_rows:
            extras["next_rows"] = next_rows(args, net, ts, sc, n_rays, dev, use_graph, devb)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

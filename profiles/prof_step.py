"""Runs a few base-light training steps (no CPU baseline, no timing) -- the target command for ncu captures:
   ncu --set full --clock-control none --import-source on --profile-from-start off [-k regex:<kernel>] -o gpurun_out/<name> python profiles/prof_step.py
   (the last step is bracketed by cudaProfilerStart/Stop; TNL_PREFETCH=0 keeps every kernel on one stream)
   TNL_CONFIG / TNL_STEPS select the workload.
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from trinerflet_b200 import scene, trainer
from trinerflet_b200.network import NeRFNetwork

cfg = scene.CONFIGS[os.environ.get("TNL_CONFIG", "base_light")]
steps = int(os.environ.get("TNL_STEPS", "2"))
net = NeRFNetwork(bound=1.5, cuda_ray=True, density_thresh=10, min_near=0.2, triplane_channels=cfg["C"],
                  triplane_resolution=cfg["R"], triplane_wavelet_levels=cfg["S"], hidden_dim=cfg["hidden"],
                  hidden_dim_color=cfg["hidden"]).cuda()
scene.init_model_(net, 0)
scene.install_ball_occupancy(net, 0.75)
ts = trainer.TrainStep(net, trainer.default_opt(), None)
ts.prefetch_planes = os.environ.get("TNL_PREFETCH", "1") == "1"
sc = scene.make_scene()
g = torch.Generator().manual_seed(0)
batches = [tuple(t.cuda() for t in scene.sample_batch(sc, cfg["rays"], g)) for _ in range(steps + 1)]
ts.forward_backward(*batches[0], update_grid=False)
net.mean_count = int(net.step_counter[0, 0].item()); net.local_step = 0
world = 1
for i in range(steps):
    net.zero_grad(set_to_none=True)
    if i == steps - 1:      # ncu --profile-from-start off: exactly one steady-state step is captured
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    ts.forward_backward(*batches[i + 1], update_grid=False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", net.mean_count)

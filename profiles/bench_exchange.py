"""torchrun micro-benchmark of the multi-GPU gradient exchange pieces (CUDA events, max over ranks)."""
import os, sys, datetime
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from trinerflet_b200 import scene, parallel
from trinerflet_b200._lib import call, ptr, stream
from trinerflet_b200.network import NeRFNetwork
from trinerflet_b200.triplane_encoder import cl_empty_planes

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr), timeout=datetime.timedelta(seconds=120))
cfg = scene.CONFIGS["base_light"]
net = NeRFNetwork(bound=1.5, cuda_ray=True, density_thresh=10, triplane_channels=cfg["C"], triplane_resolution=cfg["R"],
                  triplane_wavelet_levels=cfg["S"]).cuda()
scene.install_ball_occupancy(net, 0.75)
red = parallel.PlaneGradReducer(net, world).refresh()
g = cl_empty_planes(cfg["C"], cfg["R"], device="cuda").normal_()
dense = parallel._dense_view(g)
T, R, C = red.tile, cfg["R"], cfg["C"]

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

res = {}
res["pack"] = timeit(lambda: call("tnl_tiles_pack", ptr(dense), ptr(red.tile_ids), red.n_tiles, R, C, T, ptr(red.compact), red.bf16, stream()))
res["unpack"] = timeit(lambda: call("tnl_tiles_unpack", ptr(red.compact), ptr(red.tile_ids), red.n_tiles, R, C, T, 0.125, red.bf16, ptr(dense), stream()))
res["allreduce_compact_fp32_%.0fMB" % (red.compact.numel() * 4 / 1e6)] = timeit(lambda: dist.all_reduce(red.compact))
half = red.compact.to(torch.bfloat16)
res["allreduce_compact_bf16"] = timeit(lambda: dist.all_reduce(half))
res["cast_fp32_to_bf16"] = timeit(lambda: red.compact.to(torch.bfloat16))
small = torch.randn(13440, device="cuda")
res["allreduce_small_13k"] = timeit(lambda: dist.all_reduce(small))
res["reduce_full"] = timeit(lambda: red.reduce_(g))
chunks = red.compact.chunk(4)
def chunked():
    for c in chunks: dist.all_reduce(c)
res["allreduce_compact_4chunks"] = timeit(chunked)
big = dense.view(-1)
res["allreduce_dense_fp32_1611MB"] = timeit(lambda: dist.all_reduce(big), iters=5)
if rank == 0:
    print("world", world, "tiles", red.n_tiles, "fraction", red.fraction)
    for k, v in res.items(): print(f"{k:40s} {v:8.3f} ms")
dist.barrier(); dist.destroy_process_group()

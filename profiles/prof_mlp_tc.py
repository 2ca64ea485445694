"""Stage-level cycle profile of the tcgen05 MLP backward (CTA 0) + kernel timings at the base-light sample count."""
import sys
import torch
sys.path.insert(0, ".")
from trinerflet_b200 import _lib
from trinerflet_b200._lib import call, ptr
from trinerflet_b200.network import _FieldMLP

M, C = int(sys.argv[1]) if len(sys.argv) > 1 else 3_800_000, 32
g = torch.Generator(device="cuda").manual_seed(0)
feat = (torch.randn(M, 3 * C, device="cuda", generator=g) * 0.1).half().requires_grad_(True)
d = torch.nn.functional.normalize(torch.randn(M, 3, device="cuda", generator=g), dim=-1)
W = [torch.randn(o, i, device="cuda", generator=g) * (1.0 / i ** 0.5) for o, i in ((64, 96), (16, 64), (64, 31), (64, 64), (3, 64))]
W = [w.requires_grad_(True) for w in W]
gs = torch.randn(M, device="cuda", generator=g)
grgb = torch.randn(M, 3, device="cuda", generator=g)


def step():
    s, rgb = _FieldMLP.apply(feat, d, None, *W)
    torch.autograd.backward([s, rgb], [gs, grgb])


for _ in range(3):
    step()
torch.cuda.synchronize()
_lib.profile_start()
for _ in range(5):
    step()
for k, v in _lib.profile_stop().items():
    print(k, "ms:", [round(ms, 4) for ms, _ in v])
dbg = torch.zeros(64, dtype=torch.int64, device="cuda")
call("tnl_mlp_tc_profile", ptr(dbg))
step()
torch.cuda.synchronize()
call("tnl_mlp_tc_profile", None)
c = dbg.cpu().tolist()
iters = (M + 255) // 256 // 148 + 1
print("iterations of CTA 0 ~", iters)
print("stage  wg_wait  wg_epi | issuer_wait(g0,g1)  issuer_issue(g0,g1)   [cycles per iteration]")
for st in range(10):
    print(f"{st:5d} {c[st]/iters:8.0f} {c[10+st]/iters:7.0f} | {c[20+2*st]/iters:8.0f} {c[21+2*st]/iters:8.0f}   {c[40+2*st]/iters:8.0f} {c[41+2*st]/iters:8.0f}")
print("sum wg:", sum(c[:20]) / iters, " sum issuer:", sum(c[20:60]) / iters)

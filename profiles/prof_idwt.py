import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from trinerflet_b200._lib import call, ptr, stream
from trinerflet_b200.triplane_encoder import cl_empty_planes, cl_empty_coefs
C, n = 32, 1024
x = cl_empty_planes(C, n, device="cuda").normal_(); yh = cl_empty_coefs(C, n, device="cuda").normal_()
out = cl_empty_planes(C, 2 * n, device="cuda"); asum = torch.zeros(1, device="cuda")
gx = cl_empty_planes(C, n, device="cuda"); gyh = cl_empty_coefs(C, n, device="cuda"); g1 = torch.ones(1, device="cuda")
for _ in range(2):
    call("tnl_idwt_level_forward", ptr(x), ptr(yh), ptr(out), n, C, ptr(asum), stream())
    call("tnl_idwt_level_backward", ptr(out), ptr(gx), ptr(gyh), n, C, ptr(yh), ptr(g1), 1.0, 0, 3, stream())
torch.cuda.synchronize()

"""Drop-in for the reference's `shencoder` package (aux_libs/shencoder/sphere_harmonics.py:14-87).
Degree <= 4 forward only: on the hot path view directions never require grad (sphere_harmonics.py:84), and
inside NeRFNetwork the SH evaluation is fused into the MLP kernel; this module serves stand-alone callers."""
import torch
import torch.nn as nn
from torch.autograd import Function
from torch.amp import custom_fwd

from ._lib import call, ptr, stream


class _sh_encoder(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, inputs, degree, calc_grad_inputs=False):
        if calc_grad_inputs:
            raise NotImplementedError("SH input gradients are not on the reconstruction hot path (dirs never require grad)")
        inputs = inputs.contiguous()
        B = inputs.shape[0]
        outputs = torch.empty(B, degree ** 2, dtype=inputs.dtype, device=inputs.device)
        call("tnl_sh_encode_forward", ptr(inputs), ptr(outputs), B, int(degree), stream())
        return outputs

    @staticmethod
    def backward(ctx, grad):
        return None, None, None


sh_encode = _sh_encoder.apply


class SHEncoder(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim = input_dim
        self.degree = degree
        self.output_dim = degree ** 2
        assert self.input_dim == 3, "SH encoder only support input dim == 3"
        assert 0 < self.degree <= 4, "trinerflet_b200 SH encoder supports degree in [1, 4] (reference hot path uses 4)"

    def __repr__(self):
        return f"SHEncoder: input_dim={self.input_dim} degree={self.degree}"

    def forward(self, inputs, size=1):
        inputs = inputs / size
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.reshape(-1, self.input_dim)
        outputs = sh_encode(inputs, self.degree, False)
        return outputs.reshape(prefix_shape + [self.output_dim])

"""Density activation of the sigma head: sigma = exp(h0) evaluated in fp32 whatever the ambient autocast state, with the
gradient truncated to g * exp(clamp(h0, -15, 15)) so that a run-away logit cannot blow up the backward pass
(contract of reconstruction/activation.py:5-17).  Inside NeRFNetwork this is fused into the MLP kernels (csrc/mlp*.cu);
the stand-alone op below serves the fp32 (non-autocast) path.

Only the forward value is kept for the backward: exp is monotonic, so exp(clamp(x, -15, 15)) == clamp(exp(x), e^-15, e^15)
with the two bounds evaluated by the same fp32 exp -- the result is bit-identical to recomputing the exponential."""
import torch

_LIMIT = 15.0


class TruncatedExp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits):
        with torch.autocast(device_type=logits.device.type, enabled=False):
            value = logits.to(torch.float32).exp()
        ctx.save_for_backward(value)
        return value

    @staticmethod
    def backward(ctx, grad_value):
        (value,) = ctx.saved_tensors
        bounds = torch.tensor([-_LIMIT, _LIMIT], dtype=value.dtype, device=value.device).exp()
        return grad_value.to(value.dtype) * torch.minimum(torch.maximum(value, bounds[0]), bounds[1])


def trunc_exp(logits):
    return TruncatedExp.apply(logits)

"""Optimizer epilogue of the reference training loop as two streaming kernels.

Reference (reconstruction/nerf/utils.py:1170-1173, reconstruction/main_nerf.py:119):

    self.scaler.step(self.optimizer)        # unscale the gradients, skip the step if any is inf/NaN, else Adam
    self.scaler.update()                    # grow / back off the loss scale

with `torch.optim.Adam(model.get_params(lr), betas=(0.9, 0.99), eps=1e-15)`.  For the 1.6 GB of wavelet coefficients the
library path makes about nine coefficient-sized passes (non-finite check + unscale read/write, then the foreach Adam
chain) and costs more than the whole forward + backward of this framework.  `FusedAdam.step(scaler)` does the same
arithmetic in one read pass over the gradients (non-finite check) and one pass that reads p, g, m, v and writes p, m, v
(tnl_grad_nonfinite / tnl_adam_prepare / tnl_adam_step); step count, bias corrections, inverse scale and the skip flag
stay on the device, so there is no host synchronisation.  The loss-scale update is torch's own `_amp_update_scale_`.

`FusedAdam` is a `torch.optim.Optimizer` (param_groups / state_dict / LR schedulers work as usual); state keys follow
torch.optim.Adam (`exp_avg`, `exp_avg_sq`), the step count lives in a per-group device tensor.
"""
import torch

from ._lib import call, ptr, stream


def _dense_storage(t):
    """True if t's elements occupy numel() consecutive slots in some dimension order (e.g. a permuted contiguous tensor)."""
    expect = 1
    for size, stride in sorted(((sz, st) for sz, st in zip(t.shape, t.stride()) if sz != 1), key=lambda x: x[1]):
        if stride != expect:
            return False
        expect *= size
    return True


def _require_cuda_f32(p):
    if not p.is_cuda or p.dtype != torch.float32:
        raise RuntimeError("trinerflet_b200.FusedAdam: parameters must be fp32 CUDA tensors (no CPU fallback)")


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))

    @staticmethod
    def _flat_like(p, t):
        """t in p's memory order (the kernels walk p, g, m, v as flat arrays)."""
        if t.dtype != torch.float32 or t.stride() != p.stride():
            t = torch.empty_like(p, dtype=torch.float32).copy_(t)
        return t

    @torch.no_grad()
    def step(self, closure=None, scaler=None):
        """One optimizer step.  scaler: the torch.amp.GradScaler that scaled the loss of this step (None or disabled: the
        gradients are used as they are).  With a scaler this call replaces BOTH scaler.step(optimizer) and scaler.update()."""
        if closure is not None:
            raise RuntimeError("FusedAdam does not support closures")
        work = []
        dev = None
        for gi, group in enumerate(self.param_groups):
            for p in group["params"]:
                if p.grad is None:
                    continue
                _require_cuda_f32(p)
                if not _dense_storage(p):
                    raise RuntimeError("FusedAdam needs dense parameter storage")
                dev = p.device
                st = self.state[p]
                if len(st) == 0:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                else:
                    # moments restored by load_state_dict (e.g. from a reference / torch.optim.Adam checkpoint) keep the SAVED
                    # strides (NCHW-contiguous) and possibly another device: bring them into p's memory order, once
                    for k in ("exp_avg", "exp_avg_sq"):
                        if st[k].stride() != p.stride() or st[k].device != p.device or st[k].dtype != torch.float32:
                            st[k] = torch.empty_like(p, dtype=torch.float32).copy_(st[k])
                    if "step" in st:        # torch.optim.Adam's per-parameter step count: the group's counter starts from it
                        loaded = float(st.pop("step"))
                        group["_tnl_loaded_step"] = max(loaded, group.get("_tnl_loaded_step", 0.0))
                work.append((gi, group, p, self._flat_like(p, p.grad), st["exp_avg"], st["exp_avg_sq"]))
        if not work:
            return None
        use_scaler = scaler is not None and scaler.is_enabled()
        inv_scale = found_inf = None
        if use_scaler:
            if scaler._scale is None:
                scaler._lazy_init_scale_growth_tracker(dev)
            inv_scale = scaler._scale.double().reciprocal().float().reshape(1)      # as GradScaler._unscale_grads_
            found_inf = torch.zeros(1, dtype=torch.float32, device=dev)
            for _, _, p, g, _, _ in work:
                call("tnl_grad_nonfinite", ptr(g), p.numel(), ptr(found_inf), stream())
        states = {}
        for gi, group, p, g, m, v in work:
            if gi not in states:
                key = "_tnl_state"
                if key not in group:
                    group[key] = torch.zeros(4, dtype=torch.float32, device=dev)
                elif group[key].device != dev:      # e.g. a checkpoint mapped to the CPU: move, do not restart the bias correction
                    group[key] = group[key].to(dev)
                if "_tnl_loaded_step" in group:
                    loaded = group.pop("_tnl_loaded_step")
                    if float(group[key][0]) < loaded:
                        group[key][0] = loaded
                b1, b2 = group["betas"]
                call("tnl_adam_prepare", ptr(group[key]), ptr(found_inf), float(b1), float(b2), stream())
                states[gi] = group[key]
            b1, b2 = group["betas"]
            call("tnl_adam_step", ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), ptr(inv_scale), ptr(found_inf), ptr(states[gi]),
                 float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]), stream())
        if use_scaler:
            torch._amp_update_scale_(scaler._scale, scaler._growth_tracker, found_inf, scaler._growth_factor,
                                     scaler._backoff_factor, scaler._growth_interval)
        return None

"""trinerflet_b200 -- B200-native (sm_100a) implementation of the TriNeRFLet reconstruction hot path.

Host side = thin PyTorch modules that mirror the reference's own interfaces
(reconstruction/triplaneencoder, reconstruction/nerf/{network,renderer}.py, aux_libs/{raymarching,shencoder})
and call hand-written CUDA kernels through the C ABI of include/trinerflet_b200.h.
"""
from . import _lib  # noqa: F401

__all__ = ["raymarching", "shencoder", "activation", "encoding", "triplane_encoder", "network", "renderer", "trainer",
           "scene", "parallel"]

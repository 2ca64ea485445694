"""python profiles/check_sr_encoder.py [--points M] [--emu]
Measurement for SURVEY.md 8 f-4: the encoder of the reference's super_resolution application at the size of its config
(super_resolution/configs/triplane-sr100_400_2.yaml: 16 channels, 1024^2 planes, wavelet scale 16, low_res_scale 4) on one B200,

  ours       trinerflet_b200.sr_encoder.TriPlaneVolume (CUDA kernels through the C ABI)
  reference  the torch library ops the reference's module reaches, on the SAME GPU: the pytorch_wavelets IDWT restated with
             the same conv_transpose2d calls + F.grid_sample, in the reference's level loop (oracle/sr_encoder.py)

for the pieces of one application step (threestudio/systems/triplane_wavelet_sr.py:364-476): plane reconstruction of the
low-resolution phase (double mode off), of the double-mode phase, and the two-render encoder step (features at both
resolutions for M positions, backward into every coefficient and into the positions).  Times are CUDA events after warm-up,
median of `reps`; the two arms are also compared value by value, so the record validates itself.  Prints one JSON line.
`--emu` runs the same script on CPU tensors over the host build of the kernels (logic check only; the times mean nothing)."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--emu", action="store_true")
    ap.add_argument("--resolution", type=int, default=1024)
    args = ap.parse_args()
    if args.emu:
        from _pytest.monkeypatch import MonkeyPatch
        from tests import emu_backend
        emu_backend.install(MonkeyPatch())
        dev = torch.device("cpu")
    else:
        dev = torch.device("cuda", 0)
    from oracle import sr_encoder as osr
    from trinerflet_b200 import _lib, sr_encoder

    C, R, scale, low, high, M = 16, args.resolution, 16, 4, 1, args.points
    gen = torch.Generator().manual_seed(0)
    enc = sr_encoder.TriPlaneVolume(number_of_features=C, plane_resolution=R, inner_multi_res_scale=scale, low_res_scale=low,
                                    high_res_scale=high)
    with torch.no_grad():
        for p in enc.parameters():
            p.copy_(0.3 * torch.randn(p.shape, generator=gen))
    enc = enc.to(dev)
    x = torch.rand(M, 3, generator=gen).to(dev)
    w_low = torch.randn(M, 3 * C, generator=gen).to(dev)
    w_high = torch.randn(M, 3 * C, generator=gen).to(dev)
    # the reference arm works on NCHW-contiguous copies of the same parameters, as the reference module stores them
    pf = enc.planes_features.detach().contiguous().clone().requires_grad_(True)
    coefs = [p.detach().contiguous().clone().requires_grad_(True) for p in enc.planes_features_wavelet_coefs]

    def sync():
        if dev.type == "cuda":
            torch.cuda.synchronize()

    def timed(fn, reps=args.reps, warm=2):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(reps):
            sync()
            if dev.type == "cuda":
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            else:
                t0 = time.perf_counter()
                fn()
                ts.append(1e3 * (time.perf_counter() - t0))
        ts.sort()
        return ts[len(ts) // 2]

    # ---- ours --------------------------------------------------------------------------------------------------------
    def ours_planes(double):
        enc.enable_cache = False
        enc.set_double_mode(double)
        enc.set_resolution_mode('high_res' if double else 'low_res')
        with torch.no_grad():
            return enc.get_planes()

    def ours_step(with_x):
        enc.enable_cache = True
        enc.reset_cahce()
        enc.set_double_mode(True)
        enc.set_resolution_mode('low_res')
        enc.get_planes()
        xs = x.detach().clone().requires_grad_(with_x)
        f_low = enc(xs)
        enc.set_resolution_mode('high_res')
        f_high = enc(xs)
        enc.set_resolution_mode('low_res')
        for p in enc.parameters():
            p.grad = None
        ((f_low * w_low).sum() + (f_high * w_high).sum()).backward()
        enc.reset_cahce()
        return f_low.detach(), f_high.detach(), xs.grad

    # ---- reference arm (torch library ops, same device) ---------------------------------------------------------------------
    def ref_planes(double):
        with torch.no_grad():
            lo, hi = osr.two_readings(pf, coefs, R, low, high, double)
            return hi if double else lo

    def ref_step(with_x):
        x_low, x_high = osr.two_readings(pf, coefs, R, low, high, True)
        xs = x.detach().clone().requires_grad_(with_x)
        f_low, f_high = osr.encode(x_low, xs), osr.encode(x_high, xs)
        pf.grad = None
        for c in coefs:
            c.grad = None
        ((f_low * w_low).sum() + (f_high * w_high).sum()).backward()
        return f_low.detach(), f_high.detach(), xs.grad

    def rel(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))

    # parity of the two arms on this device, this size
    a, b = ours_step(True), ref_step(True)
    parity = {"feat_low": rel(a[0], b[0]), "feat_high": rel(a[1], b[1]), "grad_x": rel(a[2], b[2]),
              "grad_base": rel(enc.planes_features.grad, pf.grad),
              "grad_coefs": [rel(p.grad, c.grad) for p, c in zip(enc.planes_features_wavelet_coefs, coefs)],
              "planes_low": rel(ours_planes(False), ref_planes(False)), "planes_high": rel(ours_planes(True), ref_planes(True))}

    before = _lib.launch_count
    ours_step(True)
    launches = _lib.launch_count - before
    out = {
        "what": "super_resolution encoder (SURVEY 8 f-4), one B200" if not args.emu else "EMULATED on CPU: logic check only",
        "config": {"channels": C, "plane_resolution": R, "wavelet_scale": scale, "low_res_scale": low, "high_res_scale": high,
                   "points": M, "dtype": "f32", "reps": args.reps},
        "ms": {
            "planes_low_phase": {"ours": timed(lambda: ours_planes(False)), "reference_ops": timed(lambda: ref_planes(False))},
            "planes_double_mode": {"ours": timed(lambda: ours_planes(True)), "reference_ops": timed(lambda: ref_planes(True))},
            "two_render_step": {"ours": timed(lambda: ours_step(False)), "reference_ops": timed(lambda: ref_step(False))},
            "two_render_step_with_position_grad": {"ours": timed(lambda: ours_step(True)), "reference_ops": timed(lambda: ref_step(True))},
        },
        "launches_per_step_ours": launches,
        "parity_rel_l2_ours_vs_reference_ops": parity,
    }
    # the new kernel alone: d/d(positions) at the high-resolution planes
    planes = ours_planes(True)
    planes_cl = planes.permute(0, 2, 3, 1).contiguous()
    g_xyz = torch.empty(M, 3, device=dev)
    xyz = (x * 2 - 1).contiguous()
    ms = timed(lambda: _lib.call("tnl_sample_planes_backward_coords", sr_encoder.ptr(w_high), sr_encoder.ptr(planes_cl), sr_encoder.ptr(xyz),
                                 M, R, C, 1.0, 0, sr_encoder.ptr(g_xyz), sr_encoder.stream()))
    stream_bytes = M * (3 * C * 4 + 12 + 12)                  # feature gradient in, positions in, position gradient out
    texel_bytes = M * 3 * 4 * C * 4                           # four corner texels per plane, served mostly by L2
    out["k_sample_xyz_grad"] = {"ms": ms, "streamed_GB_per_s": stream_bytes / ms / 1e6, "with_texel_reads_GB_per_s": (stream_bytes + texel_bytes) / ms / 1e6,
                                "note": "streamed = g_feat + xyz + g_xyz (the bytes that must cross HBM once); texel reads are random 64-byte runs"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Device-timed 1920x1080 render of nerf/fox (training view 0) after N training steps, at the two transmittance thresholds in use
(0.01 = the application default, 1e-4 = scripts/run.py's evaluation setting), plus the wall-clock time of Testbed.render (with D2H)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))


def main():
    import torch

    import ref_app as R

    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    skips = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
    impl = R.B200(False, "Nerf")
    tb = impl.tb
    split, _ = R.fox_split()
    impl.load_transforms(split["train"])
    impl.set_network(R.network_config("L16F2"))
    for _ in range(steps):
        tb.train(R.BATCH)
    tb.sync()
    tb.shall_train = False
    tb.snap_to_pixel_centers = True
    tb.set_camera_to_training_view(0)
    tb.render_with_lens_distortion = False
    W, H = 1920, 1080
    cam, focal, center = tb._camera.render_args(W, H)
    rgba = torch.zeros(H, W, 4, device="cuda")
    depth = torch.zeros(H, W, device="cuda")
    rec = {"steps": steps}
    for sk, mt in [(sk, mt) for sk in skips for mt in (0.01, 1e-4)]:
        tb._set("render_skips_per_tile", float(sk))
        tb.nerf.render_min_transmittance = mt
        for _ in range(3):
            tb.render_device(W, H, cam, focal, rgba.data_ptr(), depth.data_ptr(), center)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            tb.render_device(W, H, cam, focal, rgba.data_ptr(), depth.data_ptr(), center)
        e1.record()
        torch.cuda.synchronize()
        ts = []
        for _ in range(4):
            a = time.perf_counter()
            img = tb.render(W, H, 1, True)
            ts.append((time.perf_counter() - a) * 1e3)
        rec[f"skips_{sk}_min_transmittance_{mt:g}"] = {"device_ms": e0.elapsed_time(e1) / 10, "wall_ms_best": min(ts), "steps": int(tb.last_render_steps),
                                             "coverage": float((img[..., 3] > 0.5).mean()), "mean_rgb": [float(x) for x in img[..., :3].mean(axis=(0, 1))]}
    print(json.dumps(rec))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""ref_app.py — ONE script that drives either the unmodified reference application (pyngp from baseline/_ref, built by
baseline/build_ref.sh) or this repo's Testbed through the same pyngp calls, on the same scene, and writes one JSON record:

    python tools/ref_app.py --impl reference --scene fox --enc L16F2 --jit 1 --steps 1000 --out gpurun_out/ref_fox.json
    python tools/ref_app.py --impl ngp_b200  --scene fox --enc L16F2          --steps 1000 --out gpurun_out/b200_fox.json

What it measures (PSNR protocol of scripts/run.py:257-317 with scripts/scenes.py's
`test_every = 5` split):
  * train: `testbed.train(1 << 18)` x steps, wall clock per step window (the reference synchronises its stream at the end of every
    train(), src/testbed.cu:4641, so wall clock IS its step time); steady state = steps 500..1000;
  * loss curve (testbed.loss every 16th step), counters (rays_per_batch, measured_batch_size) from the object or the snapshot;
  * PSNR on held-out views: load the test transforms, set_camera_to_training_view(i) (lens distortion on), render at the view's
    resolution with snap_to_pixel_centers, min transmittance 1e-4, black background; sRGB PSNR against the decoded ground truth;
  * render: 1920x1080, spp 1, wall clock of testbed.render (D2H included), best of 5;
  * artefacts: <out>.npz with a 64x64 linear crop + an 8x-downsampled frame of test view 0, <out>.ingp snapshot.
  * --load-snapshot S --no-train: renders the same views from a snapshot written by the OTHER implementation (pixel parity on
    identical weights and occupancy grid).
Nothing here is product code; nothing under oracle/ is touched."""
from __future__ import annotations

import argparse
import copy
import gzip
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "baseline" / "_ref"
BATCH = 1 << 18


def linear_to_srgb(x):
    x = np.asarray(x, dtype=np.float32)
    return np.where(x < 0.0031308, 12.92 * x, 1.055 * np.power(np.maximum(x, 1e-12), 0.41666) - 0.055)


def srgb_to_linear(x):
    x = np.asarray(x, dtype=np.float32)
    return np.where(x <= 0.04045, x / 12.92, np.power((x + 0.055) / 1.055, 2.4))


def psnr_srgb(image_linear_rgba, gt_srgb_rgb):
    a = np.clip(linear_to_srgb(image_linear_rgba[..., :3]), 0.0, 1.0)
    r = np.clip(gt_srgb_rgb, 0.0, 1.0)
    mse = float(np.mean((a - r) ** 2))
    return 10.0 * np.log10(1.0 / max(mse, 1e-12)), mse


# ---------------------------------------------------------------------------------------------------------------------
# scenes
# ---------------------------------------------------------------------------------------------------------------------
def fox_split(test_every: int = 5):
    """transforms_train.json / transforms_test.json beside the fox images: every `test_every`-th existing frame held out"""
    d = REF / "data" / "nerf" / "fox"
    if not (d / "transforms.json").exists():
        raise FileNotFoundError(f"{d}/transforms.json missing: run baseline/build_ref.sh here (it copies data/nerf/fox)")
    j = json.loads((d / "transforms.json").read_text())
    frames = sorted((f for f in j["frames"] if (d / f["file_path"]).exists()), key=lambda f: f["file_path"])
    test = [f for i, f in enumerate(frames) if i % test_every == 0]
    train = [f for i, f in enumerate(frames) if i % test_every != 0]
    out = {}
    for name, fr in (("train", train), ("test", test)):
        jj = {k: v for k, v in j.items() if k != "frames"}
        jj["frames"] = fr
        p = d / f"transforms_{name}.json"
        p.write_text(json.dumps(jj, indent=1))
        out[name] = p
    return out, d


def load_gt_srgb(path: Path) -> np.ndarray:
    from PIL import Image

    return np.asarray(Image.open(path).convert("RGB"), dtype=np.float32) / 255.0


def ball_scene(n_train=100, n_test=10, res=800):
    sys.path.insert(0, str(ROOT))
    import importlib

    S = importlib.import_module("instant-ngp_b200.synthetic")
    imgs, cams, focal = S.make_dataset(n_images=n_train, width=res, height=res)
    timgs, tcams, _ = S.make_dataset(n_images=n_test, width=res, height=res, seed=7)
    return dict(imgs=imgs, cams=cams, focal=focal, timgs=timgs, tcams=tcams, res=res)


def network_config(enc: str) -> dict:
    base = REF / "configs" / "nerf" / "base.json"
    if base.exists():
        cfg = json.loads(base.read_text())
    else:
        import importlib

        cfg = copy.deepcopy(importlib.import_module("instant-ngp_b200.synthetic").BASE_CONFIG_L16F2)
    if enc == "L16F2":
        cfg["encoding"].update(n_levels=16, n_features_per_level=2, log2_hashmap_size=19, base_resolution=16)
    elif enc == "L8F4":
        cfg["encoding"].update(n_levels=8, n_features_per_level=4, log2_hashmap_size=19, base_resolution=16)
    else:
        raise ValueError(enc)
    return cfg


# ---------------------------------------------------------------------------------------------------------------------
# the two implementations behind the same few calls
# ---------------------------------------------------------------------------------------------------------------------
class Impl:
    name = ""

    def view_resolution(self, i):
        raise NotImplementedError

    def n_views(self):
        raise NotImplementedError


class Reference(Impl):
    name = "reference"

    def __init__(self, jit: bool, train_mode: str):
        sys.path.insert(0, str(REF))
        import pyngp as ngp  # noqa: the unmodified reference, built by baseline/build_ref.sh

        self.ngp = ngp
        self.tb = ngp.Testbed()
        self.tb.root_dir = str(REF)
        self.jit, self.train_mode = jit, train_mode

    def _apply_modes(self):
        self.tb.jit_fusion = bool(self.jit)
        self.tb.nerf.training.train_mode = getattr(self.ngp.TrainMode, self.train_mode)
        self.tb.training_batch_size = BATCH

    def load_transforms(self, path):
        self.data_dir = Path(path).parent
        self.tb.load_training_data(str(path))

    def load_arrays(self, imgs, cams, focal, aabb_scale=1):
        tb = self.tb
        n, h, w, _ = imgs.shape
        tb.create_empty_nerf_dataset(n, aabb_scale)
        nodepth = np.zeros((0, 0), dtype=np.float32)
        for i in range(n):
            tb.nerf.training.set_image(i, imgs[i], nodepth, -1.0)
            tb.nerf.training.set_camera_extrinsics(i, cams[i][:3, :4], False)
            tb.nerf.training.set_camera_intrinsics(i, fx=focal, fy=focal, cx=0.5 * w, cy=0.5 * h)
        tb.nerf.training.n_images_for_training = n

    def set_network(self, cfg):
        self.tb.reload_network_from_json(cfg)
        self._apply_modes()

    def train(self):
        self.tb.train(BATCH)

    @property
    def loss(self):
        return float(self.tb.loss)

    @property
    def step(self):
        return int(self.tb.training_step)

    def counters(self):
        return None

    def prepare_eval(self):
        tb = self.tb
        tb.shall_train = False
        tb.background_color = [0.0, 0.0, 0.0, 1.0]
        tb.snap_to_pixel_centers = True
        tb.nerf.render_min_transmittance = 1e-4

    def n_views(self):
        return int(self.tb.nerf.training.dataset.n_images)

    def view_resolution(self, i):
        r = self.tb.nerf.training.dataset.metadata[i].resolution
        return int(r[0]), int(r[1])

    def view_path(self, i):
        p = Path(str(self.tb.nerf.training.dataset.paths[i]))
        return str(p if p.is_absolute() else self.data_dir / p)

    def render_view(self, i, w, h, spp=1):
        self.tb.set_camera_to_training_view(i)
        self.tb.render_with_lens_distortion = True
        return np.asarray(self.tb.render(w, h, spp, True))

    def render_current(self, w, h, spp=1):
        return np.asarray(self.tb.render(w, h, spp, True))

    def save_snapshot(self, path):
        self.tb.save_snapshot(str(path), False, True)

    def load_snapshot(self, path):
        self.tb.load_snapshot(str(path))
        self._apply_modes()

    def sync(self):
        pass


class B200(Impl):
    name = "ngp_b200"

    def __init__(self, jit: bool, train_mode: str):
        sys.path.insert(0, str(ROOT))
        import importlib

        self.P = importlib.import_module("instant-ngp_b200")
        self.tb = self.P.Testbed(self.P.TestbedMode.Nerf)
        self.tb.root_dir = str(REF)
        self.train_mode = train_mode

    def _apply_modes(self):
        self.tb.nerf.training.train_mode = getattr(self.P.TrainMode, self.train_mode)
        self.tb.training_batch_size = BATCH

    def load_transforms(self, path):
        self.tb.load_training_data(str(path))

    def load_arrays(self, imgs, cams, focal, aabb_scale=1):
        import importlib

        importlib.import_module("instant-ngp_b200.synthetic").load_into_testbed(self.tb, imgs, cams, focal, aabb_scale=aabb_scale)

    def set_network(self, cfg):
        self.tb.reload_network_from_json(cfg)
        self._apply_modes()

    def train(self):
        self.tb.train(BATCH)

    @property
    def loss(self):
        return float(self.tb.loss)

    @property
    def step(self):
        return int(self.tb.training_step)

    def counters(self):
        return self.tb.counters()

    def prepare_eval(self):
        tb = self.tb
        tb.shall_train = False
        tb.background_color = [0.0, 0.0, 0.0, 1.0]
        tb.snap_to_pixel_centers = True
        tb.nerf.render_min_transmittance = 1e-4

    def n_views(self):
        return int(self.tb._get("nerf.training.dataset.n_images"))

    def view_resolution(self, i):
        return tuple(int(x) for x in self.tb.training_view(i)["resolution"])

    def view_path(self, i):
        return str(self.tb.dataset["images"][i]["path"])

    def render_view(self, i, w, h, spp=1):
        self.tb.set_camera_to_training_view(i)
        return self.tb.render(w, h, spp, True)

    def render_current(self, w, h, spp=1):
        return self.tb.render(w, h, spp, True)

    def save_snapshot(self, path):
        self.tb.save_snapshot(str(path), False, True)

    def load_snapshot(self, path):
        self.tb.load_snapshot(str(path))
        self._apply_modes()

    def sync(self):
        self.tb.sync()


def snapshot_counters(path: Path) -> dict:
    try:
        import msgpack

        raw = Path(path).read_bytes()
        if raw[:2] == b"\x1f\x8b":
            raw = gzip.decompress(raw)
        d = msgpack.unpackb(raw, raw=False, strict_map_key=False)
        rgb = d["snapshot"]["nerf"]["rgb"]
        return {"rays_per_batch": int(rgb["rays_per_batch"]), "measured_batch_size": int(rgb["measured_batch_size"]),
                "measured_batch_size_before_compaction": int(rgb["measured_batch_size_before_compaction"]),
                "training_step": int(d["snapshot"]["training_step"])}
    except Exception as e:  # pragma: no cover
        return {"error": repr(e)}


# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", required=True, choices=["reference", "ngp_b200"])
    ap.add_argument("--scene", default="fox", choices=["fox", "ball"])
    ap.add_argument("--enc", default="L16F2", choices=["L16F2", "L8F4"])
    ap.add_argument("--jit", type=int, default=1)
    ap.add_argument("--train-mode", default="Nerf", choices=["Nerf", "Rfl", "RflRelax"])
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--out", required=True)
    ap.add_argument("--load-snapshot", default="")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-eval", action="store_true")
    ap.add_argument("--render-repeats", type=int, default=5)
    ap.add_argument("--ball-res", type=int, default=800)
    ap.add_argument("--save-views", action="store_true", help="keep every held-out view (every 6th pixel, sRGB 8 bit) in the .npz")
    ap.add_argument("--sync-every-step", action="store_true", help="wait for the device after every train() (the reference does: src/testbed.cu:4641)")
    ap.add_argument("--option", action="append", default=[], help="ngp_b200 only: name=value passed to Testbed._set before training")
    args = ap.parse_args()
    out = Path(args.out)
    out.parent.mkdir(parents=True, exist_ok=True)

    t_import = time.perf_counter()
    impl = Reference(args.jit, args.train_mode) if args.impl == "reference" else B200(args.jit, args.train_mode)
    rec = {"impl": impl.name, "scene": args.scene, "enc": args.enc, "jit": bool(args.jit) if args.impl == "reference" else None,
           "train_mode": args.train_mode, "batch": BATCH, "steps": args.steps}

    t0 = time.perf_counter()
    if args.scene == "fox":
        split, _ = fox_split()
        impl.load_transforms(split["train"])
    else:
        ball = ball_scene(res=args.ball_res)
        impl.load_arrays(ball["imgs"], ball["cams"], ball["focal"])
    rec["load_data_s"] = time.perf_counter() - t0
    rec["n_train_views"] = impl.n_views()

    impl.set_network(network_config(args.enc))
    for kv in args.option:
        k, v = kv.split("=")
        impl.tb._set(k, float(v))
    if args.load_snapshot:
        impl.load_snapshot(args.load_snapshot)
        rec["loaded_snapshot"] = args.load_snapshot

    # ---- training
    if not args.no_train and args.steps > 0:
        marks = sorted({0, 16, 100, 250, 500, args.steps} & set(range(args.steps + 1)))
        stamps, losses = {}, {}
        impl.sync()
        t_start = time.perf_counter()
        stamps[0] = t_start
        for s in range(1, args.steps + 1):
            impl.train()
            if args.sync_every_step:
                impl.sync()
            if s in marks:
                impl.sync()
                stamps[s] = time.perf_counter()
            if s in (1, 2, 4, 8) or s % 16 == 1:
                losses[s] = impl.loss
        impl.sync()
        windows = {}
        ks = sorted(stamps)
        for a, b in zip(ks[:-1], ks[1:]):
            windows[f"{a}..{b}"] = (stamps[b] - stamps[a]) / (b - a) * 1e3
        rec["ms_per_step_windows"] = windows
        lo = 500 if args.steps > 500 else ks[len(ks) // 2]
        steady_ms = (stamps[args.steps] - stamps[lo]) / (args.steps - lo) * 1e3
        rec["steady_window"] = f"{lo}..{args.steps}"
        rec["ms_per_step"] = steady_ms
        rec["train_total_s"] = stamps[args.steps] - t_start
        rec["loss_curve"] = {str(k): v for k, v in losses.items()}
        rec["final_step"] = impl.step
        snap = out.with_suffix(".ingp")
        impl.save_snapshot(snap)
        rec["snapshot"] = str(snap)
        c = impl.counters() or snapshot_counters(snap)
        rec["counters"] = c
        mbs = c.get("measured_batch_size", BATCH) or BATCH
        rec["samples_per_sec"] = mbs / (steady_ms * 1e-3)
        rec["rays_per_sec"] = c.get("rays_per_batch", 0) / (steady_ms * 1e-3)
        rec["samples_per_sec_nominal"] = BATCH / (steady_ms * 1e-3)
        out.write_text(json.dumps(rec, indent=1))   # the training record survives a failure in the evaluation below

    # ---- evaluation on held-out views
    arte = {}
    if not args.no_eval:
        impl.prepare_eval()
        psnrs, mses = [], []
        if args.scene == "fox":
            split, _ = fox_split()
            impl.load_transforms(split["test"])
            impl.prepare_eval()
            n = impl.n_views()
            for i in range(n):
                w, h = impl.view_resolution(i)
                img = impl.render_view(i, w, h, 1)
                gt = load_gt_srgb(Path(impl.view_path(i)))
                p, m = psnr_srgb(img, gt)
                psnrs.append(p)
                mses.append(m)
                if i == 0:
                    cy, cx = h // 2, w // 2
                    arte["crop64_linear"] = img[cy - 32:cy + 32, cx - 32:cx + 32].astype(np.float32)
                    arte["view0_small_linear"] = img[::8, ::8].astype(np.float32)
                    arte["view0_alpha_mean"] = np.float32(img[..., 3].mean())
                if args.save_views:
                    arte[f"view{i}_srgb8"] = (np.clip(linear_to_srgb(img[::6, ::6, :3]), 0, 1) * 255).astype(np.uint8)
                    arte[f"view{i}_err8"] = (np.clip(np.abs(np.clip(linear_to_srgb(img[::6, ::6, :3]), 0, 1) - np.clip(gt[::6, ::6], 0, 1)).mean(-1) * 4, 0, 1) * 255).astype(np.uint8)
        else:
            impl.load_arrays(ball["timgs"], ball["tcams"], ball["focal"])
            impl.prepare_eval()
            n = impl.n_views()
            for i in range(n):
                w, h = impl.view_resolution(i)
                img = impl.render_view(i, w, h, 1)
                gt = np.clip(linear_to_srgb(ball["timgs"][i][..., :3]), 0, 1)
                p, m = psnr_srgb(img, gt)
                psnrs.append(p)
                mses.append(m)
                if i == 0:
                    cy, cx = h // 2, w // 2
                    arte["crop64_linear"] = img[cy - 32:cy + 32, cx - 32:cx + 32].astype(np.float32)
                    arte["view0_small_linear"] = img[::8, ::8].astype(np.float32)
        rec["psnr_mean"] = float(np.mean(psnrs))
        rec["psnr_of_mean_mse"] = float(10.0 * np.log10(1.0 / max(np.mean(mses), 1e-12)))
        rec["psnr_per_view"] = [float(p) for p in psnrs]
        rec["n_test_views"] = len(psnrs)

        # ---- render timing: 1920x1080 from test view 0's pose (camera stays where set_camera_to_training_view left it)
        impl.render_view(0, 1920, 1080, 1)
        ts = []
        for _ in range(args.render_repeats):
            t1 = time.perf_counter()
            frame = impl.render_current(1920, 1080, 1)
            ts.append(time.perf_counter() - t1)
        rec["render_1080p_ms_best"] = min(ts) * 1e3
        rec["render_1080p_ms_all"] = [t * 1e3 for t in ts]
        rec["render_1080p_mrays_per_sec"] = 1920 * 1080 / min(ts) / 1e6
        rec["render_1080p_coverage"] = float((frame[..., 3] > 0.5).mean())
        arte["frame1080_small_linear"] = np.asarray(frame)[::8, ::8].astype(np.float32)

    rec["total_s"] = time.perf_counter() - t_import
    out.write_text(json.dumps(rec, indent=1))
    if arte:
        np.savez_compressed(out.with_suffix(".npz"), **arte)
    print(json.dumps({k: v for k, v in rec.items() if k not in ("loss_curve", "psnr_per_view", "render_1080p_ms_all")}))


if __name__ == "__main__":
    main()

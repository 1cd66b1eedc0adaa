"""Inputs for oracle/_ref/ref_nerf (the reference's own NeRF kernels, oracle/ref/ref_nerf_harness.cu) and the packing of its outputs.

    python tools/ref_nerf_cases.py write <root> [case ...]   # <root>/<case>/case.json + *.bin (every case below by default)
    python tools/ref_nerf_cases.py pack <root> <out dir>   # <out dir>/ref_nerf_<case>.npz from the harness's out_*.bin

The same `build_case` is imported by tests/test_oracle_vs_reference_nerf.py, which feeds the CPU oracle with identical inputs and
compares against the committed outputs of the reference kernels.  Everything is seeded; nothing depends on the GPU."""
import ctypes as C
import importlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import util  # noqa: E402  (tests/util.py)
from oracle import march_oracle as M  # noqa: E402

S = importlib.import_module("instant-ngp_b200.synthetic")

TRAIN_CASES = {
    # name: scene (as tests/test_gpu_march.py SCENES), rays, seed
    "train_aabb1": dict(aabb_scale=1, lens=None, full=False, radius=1.3, n_rays=2048, seed=1337),
    "train_aabb4": dict(aabb_scale=4, lens=None, full=False, radius=1.6, n_rays=2048, seed=4242),
    "train_aabb4_lens_full": dict(aabb_scale=4, lens=(0.0578421, -0.0805099, -0.000980296, 0.00015575), full=True, radius=1.2, n_rays=256, seed=99),
}
LOSS_VARIANTS = [dict(loss_type=4, random_bg_color=1), dict(loss_type=0, random_bg_color=0), dict(loss_type=1, random_bg_color=1), dict(loss_type=5, random_bg_color=0)]
GRID_CASES = {
    "grid_aabb1": dict(aabb_scale=1, radius=1.3, seed=99),
    "grid_aabb2": dict(aabb_scale=2, radius=1.5, seed=7),
}
GRID_STEPS = [dict(mark_untrained=1, clear_visible=1, n_uniform=1 << 18, n_nonuniform=0), dict(mark_untrained=0, clear_visible=0, n_uniform=1 << 17, n_nonuniform=1 << 17),
              dict(mark_untrained=1, clear_visible=0, n_uniform=1 << 16, n_nonuniform=1 << 17)]
MAX_SAMPLES = 1 << 19
RENDER_CASES = {
    # analytic field in warped coordinates: raw density = a - b * |p - 0.5|^2, raw colour = c * (p - 0.5)
    "render_aabb1": dict(aabb_scale=1, width=96, height=72, field=(7.0, 100.0, 8.0)),
    "render_aabb4": dict(aabb_scale=4, width=96, height=72, field=(7.0, 1600.0, 32.0)),
}
RENDER_MAX_STEPS = 512


def _views_json(views, n):
    out = []
    for i in range(n):
        v = views[i]
        out.append(dict(w=v.width, h=v.height, fx=float(v.focal_x), fy=float(v.focal_y), px=float(v.principal_x), py=float(v.principal_y), lens_mode=int(v.lens_mode),
                        lens_params=[float(x) for x in v.lens_params], xform=[float(x) for x in v.xform]))
    return out


def _cfg_json(cfg):
    return dict(aabb_min=[float(x) for x in cfg.aabb_min], aabb_max=[float(x) for x in cfg.aabb_max], max_cascade=int(cfg.max_cascade),
                snap_to_pixel_centers=int(cfg.snap_to_pixel_centers), cone_angle_constant=float(cfg.march.cone_angle), near_distance=float(cfg.near_distance),
                loss_scale=float(cfg.loss_scale), linear_colors=int(cfg.linear_colors), color_space=int(cfg.color_space),
                background_color=[float(x) for x in cfg.background_color], rgb_activation=int(cfg.rgb_activation), density_activation=int(cfg.density_activation))


def net_outputs(n, seed=5):
    """synthetic network outputs (fp16 x 4 per sample): a moderately dense medium so that rays terminate at different depths"""
    rng = np.random.default_rng(seed)
    out = np.zeros((n, 4), dtype=np.float16)
    out[:, 0:3] = rng.normal(0, 1.5, size=(n, 3)).astype(np.float16)
    out[:, 3] = rng.normal(1.0, 2.5, size=n).astype(np.float16)
    return out


def grid_net_outputs(n, k, seed):
    """synthetic density-network outputs (raw, before the exponential) for density-grid step k"""
    return np.random.default_rng(1000 * seed + k).normal(-2.0, 3.0, size=n).astype(np.float16)


def make_render_cfg(w, h, cam, focal, aabb_scale, spp_index=0):
    P = util.pkg()
    lib = P.load_library()
    rc = P.RenderCfg()
    rc.width, rc.height = w, h
    rc.focal_x = rc.focal_y = focal
    rc.screen_x = rc.screen_y = 0.5
    m = np.asarray(cam, dtype=np.float32)
    for c in range(4):
        for r in range(3):
            rc.camera[c * 3 + r] = m[r, c]
    half = 0.5 * aabb_scale
    for k in range(3):
        rc.aabb_min[k] = rc.render_aabb_min[k] = 0.5 - half
        rc.aabb_max[k] = rc.render_aabb_max[k] = 0.5 + half
    mc = 0
    while (1 << mc) < aabb_scale:
        mc += 1
    rc.max_cascade = mc
    assert lib.ngp_march_consts_init(C.byref(rc.march), 0.0 if aabb_scale <= 1 else 1.0 / 256.0) == 0
    rc.rgb_activation, rc.density_activation = 2, 3
    rc.min_transmittance = 0.01
    rc.spp_index = spp_index
    rc.pixel_offset[0] = rc.pixel_offset[1] = 0.5   # snap_to_pixel_centers (ld_random_pixel_offset(0) == (0.5, 0.5))
    rc.near_distance = 0.0
    rc.render_mode = 1          # ERenderMode::Shade
    rc.depth_scale = 1.0 / 0.33
    return rc


def render_field(coords, a, b, c):
    """the harness's analytic_field_kernel in numpy: float32, round to nearest, the same operation order; fp16 [n, 4]"""
    p = np.asarray(coords, dtype=np.float32)[..., :3]
    d = p - np.float32(0.5)
    r2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    out = np.empty(p.shape[:-1] + (4,), dtype=np.float16)
    out[..., :3] = (d * np.float32(c)).astype(np.float16)
    out[..., 3] = (np.float32(a) - r2 * np.float32(b)).astype(np.float16)
    return out


def build_case(name):
    """dict(kind, cfg, views, keep, rng, json, arrays{file name: ndarray}) — arrays are what the harness reads"""
    if name in RENDER_CASES:
        rcs = RENDER_CASES[name]
        w, h = rcs["width"], rcs["height"]
        cam = S.sphere_cameras(4, radius=1.25)[1]
        focal = 0.5 * w / np.tan(0.5 * np.deg2rad(45.0))
        rc = make_render_cfg(w, h, cam, focal, rcs["aabb_scale"])
        bf = util.sphere_bitfield(radius=0.3, max_cascade=rc.max_cascade)
        j = dict(type="render", width=w, height=h, focal_x=float(rc.focal_x), focal_y=float(rc.focal_y), screen_x=0.5, screen_y=0.5, camera=[float(x) for x in rc.camera],
                 aabb_min=[float(x) for x in rc.aabb_min], aabb_max=[float(x) for x in rc.aabb_max], render_aabb_min=[float(x) for x in rc.render_aabb_min],
                 render_aabb_max=[float(x) for x in rc.render_aabb_max], max_cascade=int(rc.max_cascade), cone_angle_constant=float(rc.march.cone_angle),
                 min_transmittance=float(rc.min_transmittance), near_distance=float(rc.near_distance), spp_index=int(rc.spp_index), rgb_activation=int(rc.rgb_activation),
                 density_activation=int(rc.density_activation), field_a=rcs["field"][0], field_b=rcs["field"][1], field_c=rcs["field"][2])
        return dict(kind="render", rc=rc, bitfield=bf, json=j, arrays={"bitfield.bin": bf}, field=rcs["field"])
    if name in TRAIN_CASES:
        sc = TRAIN_CASES[name]
        imgs, cams, focal = S.make_dataset(n_images=7, width=96, height=64, radius=sc["radius"])
        cfg = util.make_train_cfg(aabb_scale=sc["aabb_scale"])
        bf = util.sphere_bitfield(radius=0.3, max_cascade=cfg.max_cascade, full=sc["full"])
        views, keep = util.make_views(imgs, cams, focal, lens=sc["lens"])
        rng = M.pcg32_seed(sc["seed"])
        j = dict(type="train", n_rays=sc["n_rays"], n_rays_total=0, max_samples=MAX_SAMPLES, batch=MAX_SAMPLES, rng_state=str(rng[0]), rng_inc=str(rng[1]),
                 mean_density=0.02, loss_variants=LOSS_VARIANTS, views=_views_json(views, len(views)), **_cfg_json(cfg))
        arrays = {f"pixels_{i}.bin": keep[i] for i in range(len(keep))}
        arrays["bitfield.bin"] = bf
        arrays["net_out.bin"] = net_outputs(MAX_SAMPLES)
        return dict(kind="train", cfg=cfg, views=views, keep=keep, rng=rng, bitfield=bf, json=j, arrays=arrays, n_rays=sc["n_rays"])
    gc = GRID_CASES[name]
    imgs, cams, focal = S.make_dataset(n_images=5, width=48, height=48, radius=gc["radius"])
    cfg = util.make_train_cfg(aabb_scale=gc["aabb_scale"])
    views, keep = util.make_views(imgs, cams, focal)
    rng = M.pcg32_seed(gc["seed"])
    j = dict(type="grid", rng_state=str(rng[0]), rng_inc=str(rng[1]), decay=0.95, steps=GRID_STEPS, views=_views_json(views, len(views)), **_cfg_json(cfg))
    arrays = {f"pixels_{i}.bin": keep[i] for i in range(len(keep))}
    for k, st in enumerate(GRID_STEPS):
        arrays[f"grid_net_{k}.bin"] = grid_net_outputs(st["n_uniform"] + st["n_nonuniform"], k, gc["seed"])
    return dict(kind="grid", cfg=cfg, views=views, keep=keep, rng=rng, json=j, arrays=arrays, seed=gc["seed"])


def write_all(root, names=None):
    root = Path(root)
    for name in (names or list(TRAIN_CASES) + list(GRID_CASES) + list(RENDER_CASES)):
        c = build_case(name)
        d = root / name
        d.mkdir(parents=True, exist_ok=True)
        (d / "case.json").write_text(json.dumps(c["json"]))
        for fn, a in c["arrays"].items():
            np.ascontiguousarray(a).tofile(d / fn)
    return names or list(TRAIN_CASES) + list(GRID_CASES) + list(RENDER_CASES)


OUT_DTYPES = {"counters": np.uint32, "counter": np.uint32, "ray_indices": np.uint32, "numsteps": np.uint32, "indices": np.uint32, "dloss": np.float16, "bitfield": np.uint8,
              "steps": np.uint32}


def pack(root, out_dir):
    root, out_dir = Path(root), Path(out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    for d in sorted(p for p in root.iterdir() if p.is_dir()):
        arrs = {}
        for f in sorted(d.glob("out_*.bin")):
            key = f.stem[4:]
            dt = next((t for suffix, t in OUT_DTYPES.items() if key.endswith(suffix)), np.float32)
            arrs[key] = np.fromfile(f, dtype=dt)
        if (d / "out_rng.txt").exists():
            s, i = (d / "out_rng.txt").read_text().split()
            arrs["rng"] = np.array([int(s), int(i)], dtype=np.uint64)
        if arrs:
            np.savez_compressed(out_dir / f"ref_nerf_{d.name}.npz", **arrs)
            print(d.name, {k: v.shape for k, v in arrs.items()}, (out_dir / f"ref_nerf_{d.name}.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    if sys.argv[1] == "write":
        print(" ".join(write_all(sys.argv[2], sys.argv[3:] or None)))
    elif sys.argv[1] == "pack":
        pack(sys.argv[2], sys.argv[3])

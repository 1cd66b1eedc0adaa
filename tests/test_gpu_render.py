"""GPU parity: the persistent render kernel against the oracle (march: bit-exact sample positions; colours: tolerance)."""
import ctypes as C
import importlib

import numpy as np
import pytest

import util
from oracle import march_oracle as M
from oracle import net_oracle as O

pytestmark = pytest.mark.gpu
S = importlib.import_module("instant-ngp_b200.synthetic")


def make_render_cfg(P, lib, w, h, cam, focal, aabb_scale, spp_index=0):
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


@pytest.mark.parametrize("aabb_scale", [1, 4])
@pytest.mark.parametrize("mode", [1, 4, 3, 0, 6])   # ERenderMode: Shade, Depth, Positions, AO, Cost
def test_render_matches_oracle(aabb_scale, mode):
    import torch

    P = util.pkg()
    lib = P.load_library()
    d, L = util.make_desc(n_levels=16, F=2, log2_T=15, aabb_scale=aabb_scale)
    rng = np.random.default_rng(4)
    params = util.random_params(L, seed=31, trained_like=True)
    # make the medium moderately dense so that rays terminate after a varying number of steps
    params = params.astype(np.float16)
    cams = S.sphere_cameras(4, radius=1.25)
    w, h = 48, 40
    focal = 0.5 * w / np.tan(0.5 * np.deg2rad(45.0))
    rc = make_render_cfg(P, lib, w, h, cams[1], focal, aabb_scale)
    rc.render_mode = mode
    bf = util.sphere_bitfield(radius=0.3, max_cascade=rc.max_cascade)
    max_steps = 1024
    n_px = w * h
    counts = np.zeros(n_px, dtype=np.uint32)
    coords = np.zeros((n_px, max_steps, 7), dtype=np.float32)
    M.lib().orc_render_march(C.byref(rc), 0, h, bf.ctypes.data, max_steps, counts.ctypes.data, coords.ctypes.data)
    assert counts.max() < max_steps and counts.max() > 10
    flat = np.concatenate([coords[q, :counts[q]] for q in range(n_px)]) if counts.sum() else np.zeros((0, 7), np.float32)
    net = O.nerf_forward(L, params, flat)
    # boost density so that early termination happens (composited alpha crosses 1 - min_transmittance)
    net_out = np.zeros((n_px, max_steps, 4), dtype=np.float16)
    o = 0
    for q in range(n_px):
        net_out[q, :counts[q]] = net[o:o + counts[q]]
        o += counts[q]
    rgba_w = np.zeros((n_px, 4), dtype=np.float32)
    depth_w = np.zeros(n_px, dtype=np.float32)
    used_w = np.zeros(n_px, dtype=np.uint32)
    M.lib().orc_render_composite(C.byref(rc), 0, h, max_steps, counts.ctypes.data, coords.ctypes.data, net_out.ctypes.data, rgba_w.ctypes.data, depth_w.ctypes.data,
                                 used_w.ctypes.data)

    t_p = torch.from_numpy(params).cuda()
    t_bf = torch.from_numpy(bf).cuda()
    t_rgba = torch.full((h, w, 4), -1.0, dtype=torch.float32, device="cuda")
    t_depth = torch.full((h, w), -1.0, dtype=torch.float32, device="cuda")
    t_scr = torch.zeros(lib.ngp_nerf_render_scratch_bytes(w, h), dtype=torch.uint8, device="cuda")
    t_steps = torch.zeros(1, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    assert lib.ngp_nerf_render(C.byref(d), stream, C.byref(rc), 0, h, t_p.data_ptr(), t_bf.data_ptr(), t_rgba.data_ptr(), t_depth.data_ptr(), t_scr.data_ptr(),
                               t_steps.data_ptr()) == 0, lib.ngp_last_error()
    torch.cuda.synchronize()
    rgba_g = t_rgba.cpu().numpy().reshape(n_px, 4)
    depth_g = t_depth.cpu().numpy().reshape(n_px)
    steps_g = int(t_steps.cpu().numpy().view(np.uint32)[0])
    assert (rgba_g >= 0).all(), "pixels left unwritten"
    # network evaluations: equal to the oracle's unless a ray terminates one step earlier/later within network tolerance
    assert abs(steps_g - int(used_w.sum())) <= 0.01 * used_w.sum() + 8, (steps_g, used_w.sum())
    err = np.abs(rgba_g - rgba_w)
    scale = max(1.0, float(np.abs(rgba_w).max()))      # Depth carries distances (x 1 / dataset scale), Cost step counts / 128
    print("mode", mode, "render max abs err", err.max(axis=0), "mean alpha", rgba_w[:, 3].mean(), "steps", steps_g)
    if mode == 6:
        # Cost: integer step counts; a ray may stop one evaluation earlier / later within the network tolerance
        assert (rgba_g[:, 3] == 1.0).all() and np.abs(rgba_g[:, 0] - rgba_w[:, 0]).max() <= 1.0 / 128 + 1e-6
        assert int(round(float(rgba_g[:, 0].sum()) * 128)) == steps_g
        return
    assert err.max() <= 2e-2 * scale
    assert np.mean(err) <= 1e-3 * scale
    hit = rgba_w[:, 3] > 0.2
    close = np.abs(depth_g[hit] - depth_w[hit]) <= 2e-2
    assert close.mean() > 0.98  # the max-weight sample can flip between near-equal weights


def test_render_empty_and_tiles():
    import torch

    P = util.pkg()
    lib = P.load_library()
    d, L = util.make_desc(n_levels=8, F=4, log2_T=14, aabb_scale=1)
    params = util.random_params(L, seed=3).astype(np.float16)
    cam = S.sphere_cameras(3, radius=1.3)[0]
    w, h = 33, 17  # ragged: not a multiple of the 128-ray tile
    rc = make_render_cfg(P, lib, w, h, cam, 40.0, 1)
    t_p = torch.from_numpy(params).cuda()
    t_bf = torch.zeros(128 ** 3, dtype=torch.uint8, device="cuda")  # nothing occupied
    t_rgba = torch.full((h, w, 4), -1.0, dtype=torch.float32, device="cuda")
    t_depth = torch.full((h, w), -1.0, dtype=torch.float32, device="cuda")
    t_scr = torch.zeros(256, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    assert lib.ngp_nerf_render(C.byref(d), stream, C.byref(rc), 0, h, t_p.data_ptr(), t_bf.data_ptr(), t_rgba.data_ptr(), t_depth.data_ptr(), t_scr.data_ptr(), None) == 0
    torch.cuda.synchronize()
    assert (t_rgba == 0).all() and (t_depth == 16384.0).all()
    assert lib.ngp_nerf_render(C.byref(d), stream, C.byref(rc), 5, 5, t_p.data_ptr(), t_bf.data_ptr(), t_rgba.data_ptr(), t_depth.data_ptr(), t_scr.data_ptr(), None) != 0


@pytest.fixture(scope="module")
def lib():
    l = util.pkg().load_library()
    assert l.ngp_device_count() > 0
    return l


@pytest.mark.parametrize("color_space", [0, 1])
def test_accumulate_is_bit_exact(lib, color_space):
    """accumulate_kernel (src/render_buffer.cu:228-262) against the C oracle"""
    import ctypes as C

    import torch

    from oracle import march_oracle as M

    w, h = 37, 23
    rng = np.random.default_rng(color_space)
    acc = np.zeros((h, w, 4), dtype=np.float32)
    t_acc = torch.zeros(h, w, 4, dtype=torch.float32, device="cuda")
    for s in range(3):
        frame = rng.uniform(0, 1.2, size=(h, w, 4)).astype(np.float32)
        M.lib().orc_accumulate(w * h, frame.ctypes.data, acc.ctypes.data, float(s), color_space)
        t_f = torch.from_numpy(frame).cuda()
        assert lib.ngp_render_accumulate(torch.cuda.current_stream().cuda_stream, w, h, t_f.data_ptr(), t_acc.data_ptr(), float(s), color_space) == 0, lib.ngp_last_error()
    torch.cuda.synchronize()
    assert t_acc.cpu().numpy().tobytes() == acc.tobytes()


@pytest.mark.parametrize("curve", [0, 1, 2, 3])
@pytest.mark.parametrize("spaces", [(0, 0), (0, 1), (1, 1)])
def test_tonemap_is_bit_exact(lib, curve, spaces):
    """tonemap_kernel (src/render_buffer.cu:264-342, 511-545) against the C oracle: background blend, exposure, curve, output space"""
    import ctypes as C

    import torch

    from oracle import march_oracle as M

    P = util.pkg()
    w, h = 41, 19
    rng = np.random.default_rng(10 * curve + spaces[0] + 2 * spaces[1])
    acc = rng.uniform(0, 1, size=(h, w, 4)).astype(np.float32)
    acc[..., :3] *= acc[..., 3:4]          # premultiplied
    cfg = P.TonemapCfg()
    cfg.exposure = 0.75
    cfg.background_color[:] = [0.2, 0.4, 0.9, 1.0]
    cfg.color_space, cfg.output_color_space = spaces
    cfg.tonemap_curve = curve
    cfg.clamp_output_color = 1
    cfg.unmultiply_alpha = curve % 2
    want = np.zeros_like(acc)
    M.lib().orc_tonemap(w * h, C.byref(cfg), acc.ctypes.data, want.ctypes.data)
    t_acc = torch.from_numpy(acc).cuda()
    t_out = torch.zeros(h, w, 4, dtype=torch.float32, device="cuda")
    assert lib.ngp_render_tonemap(torch.cuda.current_stream().cuda_stream, w, h, C.byref(cfg), t_acc.data_ptr(), t_out.data_ptr()) == 0, lib.ngp_last_error()
    torch.cuda.synchronize()
    assert t_out.cpu().numpy().tobytes() == want.tobytes()
    assert (want >= 0).all() and (want <= 1).all() and (want[..., 3] == 1.0).all()    # opaque background fills alpha

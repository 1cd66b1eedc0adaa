"""GPU parity: training-ray generation / occupancy marching, compositing + loss + compaction, roll-over padding and the
occupancy-grid maintenance kernels against the C oracle — BIT-EXACT for ray indices, per-ray sample counts, marched
coordinates, loss gradients and occupancy bits (north_star: "bit-exact ray indices and sample counts")."""
import ctypes as C
import importlib

import numpy as np
import pytest

import util
from oracle import march_oracle as M
from oracle import net_oracle as O

pytestmark = pytest.mark.gpu
S = importlib.import_module("instant-ngp_b200.synthetic")


@pytest.fixture(scope="module")
def lib():
    l = util.pkg().load_library()
    assert l.ngp_device_count() > 0
    return l


def stream():
    import torch

    return torch.cuda.current_stream().cuda_stream


def dev(x):
    import torch

    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


SCENES = [
    dict(aabb_scale=1, lens=None, full=False, radius=1.3),
    dict(aabb_scale=4, lens=None, full=False, radius=1.6),
    dict(aabb_scale=4, lens=(0.0578421, -0.0805099, -0.000980296, 0.00015575), full=True, radius=1.2),  # fox-like OpenCV lens
]


def run_generator(lib, scene, n_rays, max_samples, seed=1337, ray_offset=0, n_rays_global=None, ray_stride=0, math_mode=0, lanes=0, walk=0, speculation=0):
    import time

    import torch

    t_start = time.time()

    imgs, cams, focal = S.make_dataset(n_images=7, width=96, height=64, radius=scene["radius"])
    cfg = util.make_train_cfg(aabb_scale=scene["aabb_scale"])
    cfg.ray_stride = ray_stride
    cfg.math_mode = math_mode
    cfg.gen_lanes_per_ray = lanes   # 0: chosen from the batch size; otherwise that many lanes of a warp march one ray together
    cfg.gen_walk_empty = walk       # 0: library default; empty cells one lane crosses on its own per round
    cfg.gen_speculation = speculation  # 0: library default; samples speculated in the first round after a skip
    bf = util.sphere_bitfield(radius=0.3, max_cascade=cfg.max_cascade, full=scene["full"])
    views, keep = util.make_views(imgs, cams, focal, lens=scene["lens"])
    t_views, tens = util.views_to_device(views, keep)
    rng = M.pcg32_seed(seed)
    n_rays_global = n_rays_global or n_rays
    want = M.generate_training_samples(n_rays, ray_offset, n_rays_global, rng, cfg, views, len(views), bf, max_samples)

    t_bf = dev(bf)
    t_cnt = torch.zeros(4, dtype=torch.int32, device="cuda")
    t_ri = torch.zeros(n_rays, dtype=torch.int32, device="cuda")
    t_rays = torch.zeros(n_rays, 6, dtype=torch.float32, device="cuda")
    t_ns = torch.zeros(n_rays, 2, dtype=torch.int32, device="cuda")
    t_co = torch.zeros(max_samples, 7, dtype=torch.float32, device="cuda")
    rc = lib.ngp_nerf_generate_training_samples(stream(), n_rays, ray_offset, n_rays_global, rng[0], rng[1], C.byref(cfg), t_views.data_ptr(), len(views), t_bf.data_ptr(), max_samples,
                                                t_cnt.data_ptr(), t_ri.data_ptr(), t_rays.data_ptr(), t_ns.data_ptr(), t_co.data_ptr())
    assert rc == 0, lib.ngp_last_error()
    torch.cuda.synchronize()
    print(f"run_generator: {time.time() - t_start:.2f} s total")
    cnt = t_cnt.cpu().numpy().view(np.uint32)
    got = dict(n_kept=int(cnt[0]), n_samples=int(cnt[1]), ray_indices=t_ri.cpu().numpy().view(np.uint32), rays=t_rays.cpu().numpy(),
               numsteps=t_ns.cpu().numpy().view(np.uint32), coords=t_co.cpu().numpy())
    ctx = dict(cfg=cfg, views=views, keep=keep, t_views=t_views, tens=tens, bf=bf, t_bf=t_bf, rng=rng, dev=dict(cnt=t_cnt, ri=t_ri, rays=t_rays, ns=t_ns, co=t_co))
    want["ray_offset"] = ray_offset
    return want, got, ctx


@pytest.mark.parametrize("scene", SCENES)
@pytest.mark.parametrize("shard", [(0, None, 0, 0), (4096, 16384, 0, 1), (3, 16384, 4, 4), (0, None, 0, 32), (0, None, 0, 8), (0, None, 0, 2),
                                   (0, None, 0, 16, 1, 1), (0, None, 0, 16, 64, 16), (0, None, 0, 4, 7, 3), (0, None, 0, 1, 1024, 1)])
def test_training_samples_bit_exact(lib, scene, shard):
    """shard = (ray_offset, n_rays_global, ray_stride, lanes per ray[, empty cells walked per round, speculation after a skip]): the whole batch
    with the automatic schedule; a contiguous shard, one thread per ray; rank 3 of 4 with interleaved ids, 4 lanes per ray; the whole batch with
    a warp / 8 lanes / 2 lanes per ray; the schedule knobs at their extremes"""
    n_rays, max_samples = 4096, 4096 * 1024
    walk, speculation = (shard[4], shard[5]) if len(shard) > 4 else (0, 0)
    want, got, _ = run_generator(lib, scene, n_rays, max_samples, ray_offset=shard[0], n_rays_global=shard[1], ray_stride=shard[2], lanes=shard[3], walk=walk, speculation=speculation)
    stride = shard[2] or 1
    assert want["n_samples"] <= max_samples, "test scene overflows: slot order would decide which rays are kept"
    assert want["n_samples"] > 1000, "degenerate scene"
    assert got["n_kept"] == want["n_kept"]
    assert got["n_samples"] == want["n_samples"]
    k = got["n_kept"]
    # the reference's slot order depends on atomics; compare as a map ray id -> (count, ray, coordinates)
    gmap = {int(r): j for j, r in enumerate(got["ray_indices"][:k])}
    wmap = {int(r): j for j, r in enumerate(want["ray_indices"])}
    assert set(gmap) == set(wmap)
    for rid, wj in wmap.items():
        gj = gmap[rid]
        wn, wb = want["numsteps"][wj]
        gn, gb = got["numsteps"][gj]
        assert gn == wn == want["per_ray_numsteps"][(rid - want.get("ray_offset", 0)) // stride]
        assert got["rays"][gj].tobytes() == want["rays"][wj].tobytes()
        assert got["coords"][gb:gb + gn].tobytes() == want["coords"][wb:wb + wn].tobytes()
    # slots tile [0, n_samples) without gaps or overlaps
    order = np.argsort(got["numsteps"][:k, 1])
    ends = got["numsteps"][:k, 1][order] + got["numsteps"][:k, 0][order]
    assert got["numsteps"][:k, 1][order][0] == 0 and np.array_equal(ends[:-1], got["numsteps"][:k, 1][order][1:]) and ends[-1] == got["n_samples"]


def test_training_samples_overflow_drops_rays_like_the_reference(lib):
    want, got, _ = run_generator(lib, SCENES[0], 4096, 2048)  # far too small on purpose
    assert got["n_samples"] == want["n_samples"]          # the counter keeps counting past the limit
    k = got["n_kept"]
    assert 0 < k < 4096
    assert (got["numsteps"][:k, 0] + got["numsteps"][:k, 1] <= 2048).all()


def test_empty_and_ragged_inputs(lib):
    import torch

    # zero rays is a no-op; an empty occupancy grid yields zero samples and zero rays
    imgs, cams, focal = S.make_dataset(n_images=3, width=32, height=32)
    cfg = util.make_train_cfg(aabb_scale=1)
    views, keep = util.make_views(imgs, cams, focal)
    t_views, tens = util.views_to_device(views, keep)
    t_bf = torch.zeros(128 ** 3, dtype=torch.uint8, device="cuda")
    t_cnt = torch.zeros(4, dtype=torch.int32, device="cuda")
    buf = torch.zeros(1 << 16, dtype=torch.float32, device="cuda")
    rng = M.pcg32_seed(7)
    for n in (0, 1, 33):
        assert lib.ngp_nerf_generate_training_samples(stream(), n, 0, n, rng[0], rng[1], C.byref(cfg), t_views.data_ptr(), 3, t_bf.data_ptr(), 1024, t_cnt.data_ptr(),
                                                      buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), buf.data_ptr()) == 0
    torch.cuda.synchronize()
    assert t_cnt.cpu().numpy().tolist() == [0, 0, 0, 0]


@pytest.mark.parametrize("scene", SCENES[:2])
@pytest.mark.parametrize("loss_type,random_bg,train_mode", [(4, 1, 0), (0, 0, 0), (4, 1, 1), (0, 1, 2), (4, 0, 2), (6, 1, 1)])
def test_loss_and_compaction_bit_exact(lib, scene, loss_type, random_bg, train_mode):
    import torch

    n_rays, max_samples = 4096, 4096 * 1024
    want, got, ctx = run_generator(lib, scene, n_rays, max_samples)
    # no overflow of the compacted buffer here: which rays get clipped depends on slot order (atomics-dependent in the reference too)
    batch = 1 << int(np.ceil(np.log2(max(got["n_samples"], 2))))
    cfg = ctx["cfg"]
    cfg.loss_type, cfg.random_bg_color, cfg.train_mode = loss_type, random_bg, train_mode   # train_mode: Nerf / Rfl / RflRelax (train_nerf.cuh:391-410)
    k, ns = got["n_kept"], got["n_samples"]
    # synthetic network outputs: moderately dense medium so that rays terminate at different depths
    rng = np.random.default_rng(5)
    net_out = np.zeros((max_samples, 4), dtype=np.float16)
    net_out[:, 0:3] = rng.normal(0, 1.5, size=(max_samples, 3)).astype(np.float16)
    net_out[:, 3] = rng.normal(1.0, 2.5, size=max_samples).astype(np.float16)
    mean_density = np.float32(0.02)

    # oracle on the GPU's own slot assignment (so that outputs line up index by index)
    ns_host = got["numsteps"][:k].copy()
    co_w = np.zeros((batch, 7), dtype=np.float32)
    dl_w = np.zeros((batch, 4), dtype=np.float16)
    loss_w = np.zeros(n_rays, dtype=np.float32)
    comp_w = M.lib().orc_compute_loss(k, n_rays, ctx["rng"][0], ctx["rng"][1], C.byref(cfg), C.addressof(ctx["views"]), len(ctx["views"]),
                                      net_out.ctypes.data, batch, got["ray_indices"].ctypes.data, got["rays"].ctypes.data, ns_host.ctypes.data,
                                      got["coords"].ctypes.data, co_w.ctypes.data, dl_w.ctypes.data, loss_w.ctypes.data, mean_density)

    d = ctx["dev"]
    t_no = dev(net_out)
    t_coc = torch.zeros(batch, 7, dtype=torch.float32, device="cuda")
    t_dl = torch.zeros(batch, 4, dtype=torch.float16, device="cuda")
    t_loss = torch.zeros(n_rays, dtype=torch.float32, device="cuda")
    t_md = dev(np.array([mean_density], dtype=np.float32))
    rc = lib.ngp_nerf_compute_loss(stream(), n_rays, n_rays, ctx["rng"][0], ctx["rng"][1], C.byref(cfg), ctx["t_views"].data_ptr(), len(ctx["views"]), t_no.data_ptr(), batch,
                                   d["cnt"].data_ptr(), d["ri"].data_ptr(), d["rays"].data_ptr(), d["ns"].data_ptr(), d["co"].data_ptr(), t_coc.data_ptr(),
                                   t_dl.data_ptr(), t_loss.data_ptr(), t_md.data_ptr())
    assert rc == 0, lib.ngp_last_error()
    torch.cuda.synchronize()
    cnt = d["cnt"].cpu().numpy().view(np.uint32)
    assert int(cnt[2]) == comp_w and comp_w > 0
    ns_g = d["ns"].cpu().numpy().view(np.uint32)[:k]
    co_g, dl_g, loss_g = t_coc.cpu().numpy(), t_dl.cpu().numpy(), t_loss.cpu().numpy()
    # per-ray compacted counts are exact; the compacted base differs (warp order vs ray order) -> compare ray by ray
    assert np.array_equal(ns_g[:, 0], ns_host[:, 0])
    assert loss_g[:k].tobytes() == loss_w[:k].tobytes()
    checked = 0
    for i in range(k):
        n_i = int(ns_g[i, 0])
        if n_i == 0:
            continue
        gb, wb = int(ns_g[i, 1]), int(ns_host[i, 1])
        assert co_g[gb:gb + n_i].tobytes() == co_w[wb:wb + n_i].tobytes()
        assert dl_g[gb:gb + n_i].tobytes() == dl_w[wb:wb + n_i].tobytes(), f"ray slot {i}"
        checked += n_i
    assert checked == min(comp_w, batch) or comp_w > batch

    # roll-over padding (fill_rollover_and_rescale + fill_rollover)
    t_c3 = torch.zeros(4, dtype=torch.int32, device="cuda")
    n_comp_small = 1000
    batch = min(batch, 1 << 16)
    t_c3[2] = n_comp_small
    co_pad, dl_pad = co_g.copy(), dl_g.copy()
    M.lib().orc_fill_rollover(batch, n_comp_small, co_pad.ctypes.data, dl_pad.ctypes.data)
    assert lib.ngp_nerf_fill_rollover(stream(), batch, t_c3.data_ptr(), t_coc.data_ptr(), t_dl.data_ptr()) == 0
    torch.cuda.synchronize()
    assert t_coc.cpu().numpy().tobytes() == co_pad.tobytes()
    assert t_dl.cpu().numpy().tobytes() == dl_pad.tobytes()


def _run_loss(lib, ctx, got, net_out_dev, batch, order, n_rays):
    import torch

    cfg = ctx["cfg"]
    cfg.compaction_order = order
    d = ctx["dev"]
    t_cnt = d["cnt"].clone()
    t_ns = d["ns"].clone()
    t_coc = torch.zeros(batch, 7, dtype=torch.float32, device="cuda")
    t_dl = torch.zeros(batch, 4, dtype=torch.float16, device="cuda")
    t_loss = torch.zeros(n_rays, dtype=torch.float32, device="cuda")
    t_md = dev(np.array([0.02], dtype=np.float32))
    rc = lib.ngp_nerf_compute_loss(stream(), n_rays, n_rays, ctx["rng"][0], ctx["rng"][1], C.byref(cfg), ctx["t_views"].data_ptr(), len(ctx["views"]), net_out_dev.data_ptr(),
                                   batch, t_cnt.data_ptr(), d["ri"].data_ptr(), d["rays"].data_ptr(), t_ns.data_ptr(), d["co"].data_ptr(), t_coc.data_ptr(), t_dl.data_ptr(),
                                   t_loss.data_ptr(), t_md.data_ptr())
    assert rc == 0, lib.ngp_last_error()
    torch.cuda.synchronize()
    k = got["n_kept"]
    return dict(total=int(t_cnt.cpu().numpy().view(np.uint32)[2]), ns=t_ns.cpu().numpy().view(np.uint32)[:k], co=t_coc.cpu().numpy(), dl=t_dl.cpu().numpy(),
                loss=t_loss.cpu().numpy()[:k])


def test_per_image_exposure_matches_oracle(lib):
    """ngp_nerf_train_cfg.cam_exposure / cam_exposure_gradient (testbed_nerf.cu:979-995, 1142-1155): targets, and with them every sample
    gradient, bit-exact against the oracle; the per-view accumulators to float-atomics order"""
    import torch

    n_rays, max_samples = 2048, 2048 * 1024
    want, got, ctx = run_generator(lib, SCENES[1], n_rays, max_samples)
    k = got["n_kept"]
    n_views = len(ctx["views"])
    rng = np.random.default_rng(9)
    net_out = np.zeros((max_samples, 4), dtype=np.float16)
    net_out[:, 0:3] = rng.normal(0, 1.5, size=(max_samples, 3)).astype(np.float16)
    net_out[:, 3] = rng.normal(1.0, 2.5, size=max_samples).astype(np.float16)
    expo = rng.uniform(-1.0, 1.0, size=(n_views, 3)).astype(np.float32)
    batch = 1 << int(np.ceil(np.log2(max(got["n_samples"], 2))))
    cfg = ctx["cfg"]
    cfg.compaction_order = 1    # one kernel; the oracle hands slots out in ray-slot order, outputs are compared ray by ray
    for linear_colors in (0, 1):
        cfg.linear_colors = linear_colors
        # oracle (host pointers)
        grad_w = np.zeros((n_views, 3), dtype=np.float32)
        cfg.cam_exposure, cfg.cam_exposure_gradient = expo.ctypes.data, grad_w.ctypes.data
        ns_host = got["numsteps"][:k].copy()
        co_w = np.zeros((batch, 7), dtype=np.float32)
        dl_w = np.zeros((batch, 4), dtype=np.float16)
        loss_w = np.zeros(n_rays, dtype=np.float32)
        M.lib().orc_compute_loss(k, n_rays, ctx["rng"][0], ctx["rng"][1], C.byref(cfg), C.addressof(ctx["views"]), n_views, net_out.ctypes.data, batch,
                                 got["ray_indices"].ctypes.data, got["rays"].ctypes.data, ns_host.ctypes.data, got["coords"].ctypes.data, co_w.ctypes.data, dl_w.ctypes.data,
                                 loss_w.ctypes.data, np.float32(0.02))
        # library (device pointers)
        t_expo, t_grad = dev(expo), torch.zeros(n_views, 3, dtype=torch.float32, device="cuda")
        cfg.cam_exposure, cfg.cam_exposure_gradient = t_expo.data_ptr(), t_grad.data_ptr()
        r = _run_loss(lib, ctx, got, dev(net_out), batch, 1, n_rays)
        assert np.array_equal(r["ns"][:, 0], ns_host[:, 0])
        assert r["loss"].tobytes() == loss_w[:k].tobytes()
        for i in range(0, k, 5):
            n_i, gb, wb = int(r["ns"][i, 0]), int(r["ns"][i, 1]), int(ns_host[i, 1])
            assert r["dl"][gb:gb + n_i].tobytes() == dl_w[wb:wb + n_i].tobytes(), f"ray slot {i}"
        g = t_grad.cpu().numpy()
        assert np.abs(grad_w).max() > 0 and np.allclose(g, grad_w, rtol=2e-4, atol=1e-5 * np.abs(grad_w).max())
        # exposure without an accumulator: same gradients, accumulator untouched
        t_grad.zero_()
        cfg.cam_exposure_gradient = None
        r2 = _run_loss(lib, ctx, got, dev(net_out), batch, 1, n_rays)
        assert not t_grad.cpu().numpy().any()
        for i in range(0, k, 5):   # slot bases differ from run to run in this order: ray by ray
            n_i, a, b = int(r["ns"][i, 0]), int(r["ns"][i, 1]), int(r2["ns"][i, 1])
            assert int(r2["ns"][i, 0]) == n_i and r2["dl"][b:b + n_i].tobytes() == r["dl"][a:a + n_i].tobytes()
    cfg.cam_exposure, cfg.cam_exposure_gradient = None, None


def test_compaction_orders(lib):
    """ngp_nerf_train_cfg.compaction_order decides where a ray's samples go in the compacted batch — and with it which rays an overfull
    batch cuts — not what they are: groups of 32 consecutive rays (the default), one atomic per ray, ascending ray id."""
    n_rays, max_samples = 4096, 4096 * 1024
    want, got, ctx = run_generator(lib, SCENES[1], n_rays, max_samples)
    k = got["n_kept"]
    rng = np.random.default_rng(5)
    net_out = np.zeros((max_samples, 4), dtype=np.float16)
    net_out[:, 0:3] = rng.normal(0, 1.5, size=(max_samples, 3)).astype(np.float16)
    net_out[:, 3] = rng.normal(1.0, 2.5, size=max_samples).astype(np.float16)
    t_no = dev(net_out)
    ray_ids = got["ray_indices"][:k].astype(np.int64)
    big = 1 << int(np.ceil(np.log2(max(got["n_samples"], 2))))
    runs = {order: _run_loss(lib, ctx, got, t_no, big, order, n_rays) for order in (1, 0, 2)}
    ref = runs[1]
    assert ref["total"] > 0
    for order, r in runs.items():
        assert r["total"] == ref["total"]                                  # the counter: every ray's samples, not clamped
        assert np.array_equal(r["ns"][:, 0], ref["ns"][:, 0])              # per-ray counts
        assert r["loss"].tobytes() == ref["loss"].tobytes()
        cnt, base = r["ns"][:, 0].astype(np.int64), r["ns"][:, 1].astype(np.int64)
        o = np.argsort(base, kind="stable")
        nz = o[cnt[o] > 0]
        assert base[nz][0] == 0 and np.array_equal(base[nz][1:], (base[nz] + cnt[nz])[:-1]) and base[nz][-1] + cnt[nz][-1] == r["total"]   # slots tile [0, total)
        for i in np.flatnonzero(cnt)[::7]:
            a, b, n = int(base[i]), int(ref["ns"][i, 1]), int(cnt[i])
            assert r["co"][a:a + n].tobytes() == ref["co"][b:b + n].tobytes()
            assert r["dl"][a:a + n].tobytes() == ref["dl"][b:b + n].tobytes()
    # ascending ray id
    cnt, base = runs[2]["ns"][:, 0].astype(np.int64), runs[2]["ns"][:, 1].astype(np.int64)
    by_id = np.argsort(ray_ids)
    assert np.array_equal(base[by_id], np.concatenate([[0], np.cumsum(cnt[by_id])[:-1]]))
    # groups of 32 consecutive ray ids stay together, in ray order; the groups themselves are shuffled
    cnt, base = runs[0]["ns"][:, 0].astype(np.int64), runs[0]["ns"][:, 1].astype(np.int64)
    starts = []
    for g in range(n_rays // 32):
        m = by_id[(ray_ids[by_id] // 32) == g]
        if len(m) == 0:
            continue
        assert np.array_equal(base[m], base[m][0] + np.concatenate([[0], np.cumsum(cnt[m])[:-1]]))
        starts.append(base[m][0])
    assert len(starts) > 8 and not np.array_equal(np.sort(starts), np.array(starts))
    # an overfull batch: the rays behind the limit keep nothing, the one across it keeps its head, everything before is whole
    small = (ref["total"] // 2) // 256 * 256
    for order in (0, 2):
        whole = runs[order]["ns"]
        cut = _run_loss(lib, ctx, got, t_no, small, order, n_rays)
        assert cut["total"] == ref["total"]
        assert np.array_equal(cut["ns"][:, 1], whole[:, 1])
        room = np.clip(small - whole[:, 1].astype(np.int64), 0, None)
        assert np.array_equal(cut["ns"][:, 0], np.minimum(whole[:, 0], room))
        assert int(cut["ns"][:, 0].sum()) == small
        assert cut["co"][:small].tobytes() == runs[order]["co"][:small].tobytes()


@pytest.mark.parametrize("aabb_scale", [1, 4])
def test_density_grid_update_matches_oracle(lib, aabb_scale):
    import torch

    d, L = util.make_desc(n_levels=16, F=2, log2_T=17, aabb_scale=aabb_scale)
    params = util.random_params(L, seed=21, trained_like=True).astype(np.float16)
    imgs, cams, focal = S.make_dataset(n_images=5, width=48, height=48, radius=1.3)
    cfg = util.make_train_cfg(aabb_scale=aabb_scale)
    views, keep = util.make_views(imgs, cams, focal)
    t_views, tens = util.views_to_device(views, keep)
    n_casc = cfg.max_cascade + 1
    n_el = 128 ** 3 * n_casc
    t_p = dev(params)
    t_grid = torch.zeros(128 ** 3 * 8, dtype=torch.float32, device="cuda")
    t_bf = torch.zeros(128 ** 3, dtype=torch.uint8, device="cuda")
    t_mean = torch.zeros(4, dtype=torch.float32, device="cuda")
    scratch_bytes = lib.ngp_nerf_density_grid_scratch_bytes(cfg.max_cascade)
    t_scr = torch.zeros(scratch_bytes, dtype=torch.uint8, device="cuda")
    rng = M.pcg32_seed(99)

    grid_w = np.zeros(n_el, dtype=np.float32)
    tmp_w = np.zeros(n_el, dtype=np.float32)
    state = C.c_uint64(rng[0])
    state_w = C.c_uint64(rng[0])
    for step, ema_step in [(0, 0), (16, 1), (300, 2)]:
        rc = lib.ngp_nerf_update_density_grid(C.byref(d), stream(), C.byref(cfg), t_p.data_ptr(), C.byref(state), rng[1], step, ema_step, 0.95, t_views.data_ptr(), len(views),
                                              t_grid.data_ptr(), t_bf.data_ptr(), t_mean.data_ptr(), t_scr.data_ptr())
        assert rc == 0, lib.ngp_last_error()
        torch.cuda.synchronize()
        # ---- oracle, fed with the GPU's density-network outputs so that the integer / bit results can be compared exactly
        if step == 0:
            M.lib().orc_mark_untrained_density_grid(n_el, grid_w.ctypes.data, len(views), C.addressof(views), 1)
        n_uni = n_el if step < 256 else n_el // 4
        n_non = 0 if step < 256 else n_el // 4
        pos_w = np.zeros((n_uni + n_non, 4), dtype=np.float32)
        idx_w = np.zeros(n_uni + n_non, dtype=np.uint32)
        M.lib().orc_generate_grid_samples(n_uni, state_w.value, rng[1], ema_step, C.byref(cfg), grid_w.ctypes.data, pos_w.ctypes.data, idx_w.ctypes.data, n_casc, -0.01)
        M.lib().orc_pcg32_advance(C.byref(state_w), rng[1], 1 << 32)
        if n_non:
            M.lib().orc_generate_grid_samples(n_non, state_w.value, rng[1], ema_step, C.byref(cfg), grid_w.ctypes.data, pos_w[n_uni:].ctypes.data, idx_w[n_uni:].ctypes.data,
                                              n_casc, 0.01)
        M.lib().orc_pcg32_advance(C.byref(state_w), rng[1], 1 << 32)
        assert state.value == state_w.value
        scr = t_scr.cpu().numpy()
        n_tot = n_uni + n_non
        pos_g = scr[: n_el * 16].view(np.float32).reshape(-1, 4)[:n_tot]
        idx_g = scr[n_el * 16: n_el * 20].view(np.uint32)[:n_tot]
        mlp_g = scr[n_el * 24: n_el * 24 + n_tot * 2].view(np.float16)
        assert idx_g.tobytes() == idx_w.tobytes()
        assert pos_g.tobytes() == pos_w.tobytes()
        # network: tolerance parity on a sample
        sel = np.random.default_rng(step).choice(n_tot, 2048, replace=False)
        want_d = O.nerf_density(L, params, pos_w[sel, :3]).astype(np.float32)
        assert np.abs(mlp_g[sel].astype(np.float32) - want_d).max() <= 1e-2 * max(1.0, np.abs(want_d).max())
        M.lib().orc_splat_and_ema(n_tot, idx_w.ctypes.data, np.ascontiguousarray(mlp_g).ctypes.data, cfg.density_activation, n_el, 0.95, tmp_w.ctypes.data, grid_w.ctypes.data)
        grid_g = t_grid.cpu().numpy()[:n_el]
        assert grid_g.tobytes() == grid_w.tobytes()
        mean_w = M.lib().orc_density_mean(grid_w.ctypes.data)
        assert np.float32(mean_w).tobytes() == t_mean.cpu().numpy()[:1].tobytes()
        bf_w = np.zeros(128 ** 3, dtype=np.uint8)
        M.lib().orc_update_bitfield(cfg.max_cascade, grid_w.ctypes.data, mean_w, bf_w.ctypes.data)
        assert t_bf.cpu().numpy().tobytes() == bf_w.tobytes()
    assert (grid_w > 0).sum() > 0
    if aabb_scale > 1:
        assert (grid_w < 0).sum() > 0  # voxels no camera sees were culled at step 0


@pytest.mark.parametrize("scene,train_mode", [(SCENES[0], 0), (SCENES[1], 0), (SCENES[1], 2)])
def test_ray_ordered_inference_equals_full_inference_where_the_loss_reads(lib, scene, train_mode):
    """the early-terminating inference pass (k_nerf_forward_rays) writes, for every ray, exactly the rows the loss kernel
    consumes, bit-identical to the evaluate-everything pass of the reference schedule"""
    import torch

    n_rays, max_samples = 4096, 4096 * 1024
    want, got, ctx = run_generator(lib, scene, n_rays, max_samples)
    cfg = ctx["cfg"]
    cfg.train_mode = train_mode   # Rfl / RflRelax stop a ray on the fused train kernel's form of the transmittance (1 - accumulated weight)
    k, ns = got["n_kept"], got["n_samples"]
    d, L = util.make_desc(n_levels=16, F=2, log2_T=16, aabb_scale=scene["aabb_scale"])
    rng = np.random.default_rng(77)
    params = util.random_params(L, seed=41, trained_like=True)
    # bias the density head so that rays saturate after a varying, finite number of samples
    dens_out_off = 64 * 32
    params[dens_out_off:dens_out_off + 64] = np.abs(params[dens_out_off:dens_out_off + 64]) * 60.0
    params = params.astype(np.float16)
    t_p = dev(params)
    dv = ctx["dev"]
    full = torch.zeros(ns, 4, dtype=torch.float16, device="cuda")
    assert lib.ngp_nerf_inference(C.byref(d), stream(), ns, dv["co"].data_ptr(), t_p.data_ptr(), full.data_ptr(), 4) == 0, lib.ngp_last_error()
    rays_out = torch.full((ns, 4), float("nan"), dtype=torch.float16, device="cuda")
    queue = torch.zeros(1, dtype=torch.int32, device="cuda")
    assert lib.ngp_nerf_inference_rays(C.byref(d), stream(), n_rays, dv["cnt"].data_ptr(), queue.data_ptr(), dv["ns"].data_ptr(), dv["co"].data_ptr(), t_p.data_ptr(),
                                       cfg.density_activation, rays_out.data_ptr(), cfg.train_mode) == 0, lib.ngp_last_error()
    # run the loss kernel on the FULL outputs to learn what it consumes
    batch = 1 << int(np.ceil(np.log2(max(ns, 2))))
    t_coc = torch.zeros(batch, 7, dtype=torch.float32, device="cuda")
    t_dl = torch.zeros(batch, 4, dtype=torch.float16, device="cuda")
    t_md = dev(np.array([0.02], dtype=np.float32))
    ns_before = dv["ns"].cpu().numpy().view(np.uint32)[:k].copy()
    assert lib.ngp_nerf_compute_loss(stream(), n_rays, n_rays, ctx["rng"][0], ctx["rng"][1], C.byref(cfg), ctx["t_views"].data_ptr(), len(ctx["views"]), full.data_ptr(), batch,
                                     dv["cnt"].data_ptr(), dv["ri"].data_ptr(), dv["rays"].data_ptr(), dv["ns"].data_ptr(), dv["co"].data_ptr(), t_coc.data_ptr(),
                                     t_dl.data_ptr(), None, t_md.data_ptr()) == 0, lib.ngp_last_error()
    torch.cuda.synchronize()
    consumed = dv["ns"].cpu().numpy().view(np.uint32)[:k, 0]
    full_h, rays_h = full.cpu().numpy().view(np.uint16), rays_out.cpu().numpy().view(np.uint16)
    n_eval = int((~np.isnan(rays_out.float().cpu().numpy()[:, 0])).sum())
    assert consumed.sum() > 0 and (consumed < ns_before[:, 0]).any(), "scene does not exercise early termination"
    for i in range(k):
        b, c = int(ns_before[i, 1]), int(consumed[i])
        assert np.array_equal(rays_h[b:b + c], full_h[b:b + c]), f"ray slot {i}"
    # and it evaluates far fewer samples than the full pass: at most one 8-sample chunk beyond what is consumed
    assert n_eval <= consumed.sum() + 8 * k
    print("evaluated", n_eval, "of", ns, "consumed", int(consumed.sum()))


@pytest.mark.parametrize("scene", SCENES)
def test_reference_flavour_stays_within_rounding_of_the_deterministic_one(lib, scene):
    """NGP_MATH_REFERENCE (march_ref.cu: the reference build's fast-math arithmetic) against NGP_MATH_DETERMINISTIC on the same batch:
    the same rays, sample counts equal on all but the rays that graze a voxel face, coordinates equal to ~1e-5.  (That the reference
    flavour reproduces the reference KERNEL's counts exactly is tests/test_gpu_vs_reference_nerf.py.)"""
    n_rays, max_samples = 4096, 4096 * 1024
    _, det, _ = run_generator(lib, scene, n_rays, max_samples)
    _, ref, _ = run_generator(lib, scene, n_rays, max_samples, math_mode=1)
    kd, kr = det["n_kept"], ref["n_kept"]
    dmap = {int(r): j for j, r in enumerate(det["ray_indices"][:kd])}
    rmap = {int(r): j for j, r in enumerate(ref["ray_indices"][:kr])}
    assert len(set(dmap) ^ set(rmap)) <= max(2, 0.003 * kd)
    same, worst = 0, 0.0
    for rid in set(dmap) & set(rmap):
        dn, db = det["numsteps"][dmap[rid]]
        rn, rb = ref["numsteps"][rmap[rid]]
        assert np.allclose(det["rays"][dmap[rid]], ref["rays"][rmap[rid]], rtol=0, atol=3e-6)
        if dn == rn:
            same += 1
            worst = max(worst, float(np.abs(det["coords"][db:db + dn] - ref["coords"][rb:rb + rn]).max()))
        else:
            assert abs(int(dn) - int(rn)) <= max(2, 0.02 * int(dn))
    assert same >= 0.985 * len(set(dmap) & set(rmap)) and worst < 3e-5
    # slots tile [0, n_samples) in this flavour too
    order = np.argsort(ref["numsteps"][:kr, 1])
    ends = ref["numsteps"][:kr, 1][order] + ref["numsteps"][:kr, 0][order]
    assert ref["numsteps"][:kr, 1][order][0] == 0 and np.array_equal(ends[:-1], ref["numsteps"][:kr, 1][order][1:]) and ends[-1] == ref["n_samples"]

"""CPU, world_size 2 over gloo: the data-parallel sharding rule of SURVEY.md §8(e) / DESIGN.md §6.

Each rank generates the training rays of its shard with GLOBAL ray ids — rank r of W takes ids r, r + W, r + 2W, ... (the interleaved
partition of Testbed::train: every rank draws from every view); the CUDA generator takes the same (ray_offset, ray_stride,
n_rays_global) triple; tests/test_gpu_march.py checks it against this oracle — computes loss gradients normalised
by the global ray count and the network gradient of its samples; the ranks all-reduce (sum) the flat gradient.  The result
must equal what one process computes for the whole batch."""
import ctypes as C
import importlib
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard_work(rank, world, n_rays_global):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import util
    from oracle import march_oracle as M
    from oracle import net_oracle as O

    S = importlib.import_module("instant-ngp_b200.synthetic")
    imgs, cams, focal = S.make_dataset(n_images=5, width=48, height=48)
    cfg = util.make_train_cfg(aabb_scale=1)
    views, keep = util.make_views(imgs, cams, focal)
    bf = util.sphere_bitfield(radius=0.3, max_cascade=0)
    rng = M.pcg32_seed(1337)
    n_local = n_rays_global // world
    cfg.ray_stride = world
    g = M.generate_training_samples(n_local, rank, n_rays_global, rng, cfg, views, len(views), bf, n_local * 256)
    k, ns = g["n_kept"], g["n_samples"]
    d, L = util.make_desc(n_levels=16, F=2, log2_T=12, aabb_scale=1)
    params = util.random_params(L, seed=0, trained_like=True).astype(np.float16)
    net_out = O.nerf_forward(L, params, g["coords"][:ns])
    numsteps = g["numsteps"].copy()
    batch = 1 << 17
    co = np.zeros((batch, 7), dtype=np.float32)
    dl = np.zeros((batch, 4), dtype=np.float16)
    loss = np.zeros(n_local, dtype=np.float32)
    comp = M.lib().orc_compute_loss(k, n_rays_global, rng[0], rng[1], C.byref(cfg), C.addressof(views), len(views), np.ascontiguousarray(net_out).ctypes.data, batch,
                                    g["ray_indices"].ctypes.data, g["rays"].ctypes.data, numsteps.ctypes.data, g["coords"].ctypes.data, co.ctypes.data, dl.ctypes.data,
                                    loss.ctypes.data, 0.02)
    comp = min(comp, batch)
    grad = O.nerf_backward(L, params, co[:comp], dl[:comp])
    return dict(per_ray=g["per_ray_numsteps"], grad=grad, loss=float(loss.sum()), n_samples=ns, comp=comp)


def _worker(rank, world, port, n_rays_global, out_q):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = _shard_work(rank, world, n_rays_global)
    g = torch.from_numpy(w["grad"])
    dist.all_reduce(g)  # the one exchange step of the path: sum of the flat gradient buffer
    cnt = torch.tensor([w["n_samples"], w["comp"]], dtype=torch.int64)
    dist.all_reduce(cnt)
    per_ray = [torch.zeros(n_rays_global // world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(per_ray, torch.from_numpy(w["per_ray"].astype(np.int64)))
    loss = torch.tensor([w["loss"]], dtype=torch.float64)
    dist.all_reduce(loss)
    if rank == 0:
        # rank r holds global ids r, r + W, ...: interleave the per-rank lists back into id order
        out_q.put(dict(grad=g.numpy(), counts=cnt.numpy(), per_ray=torch.stack(per_ray, dim=1).reshape(-1).numpy(), loss=float(loss)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_reproduce_the_single_process_batch():
    import torch.multiprocessing as mp

    n_rays_global = 1024
    single = _shard_work(0, 1, n_rays_global)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_rays_global, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # bit-exact integer parity: every ray marches the same number of steps whichever rank draws it
    assert np.array_equal(res["per_ray"], single["per_ray"].astype(np.int64))
    assert res["counts"][0] == single["n_samples"] and res["counts"][1] == single["comp"]
    assert abs(res["loss"] - single["loss"]) <= 1e-6 * abs(single["loss"]) + 1e-9
    # gradient of the union == sum of shard gradients (float summation order differs)
    scale = np.abs(single["grad"]).max()
    assert scale > 0
    assert np.abs(res["grad"] - single["grad"]).max() <= 2e-3 * scale

"""GPU: this library's CUDA sample generator, through the C-ABI, against what the REFERENCE's own generate_training_samples_nerf kernel
produced for the same seeded inputs (tests/golden/ref_nerf_train_*.npz, oracle/ref/ref_nerf_harness.cu) — the direct form of the chain
CUDA == oracle (tests/test_gpu_march.py, bit for bit) and oracle ~ reference kernel (tests/test_oracle_vs_reference_nerf.py).  Same
comparison, same tolerances: the residue is the reference's fast-math arithmetic (see that file's docstring)."""
import ctypes as C

import numpy as np
import pytest

import util
from ref_nerf_compare import RC, compare_generation, golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("math_mode", [0, 1])
@pytest.mark.parametrize("name", list(RC.TRAIN_CASES))
def test_cuda_sample_generation_matches_the_reference_kernel(name, math_mode):
    import torch

    lib = util.pkg().load_library()
    assert lib.ngp_device_count() > 0
    c = RC.build_case(name)
    n_rays, max_samples = c["n_rays"], RC.MAX_SAMPLES
    views, cfg, rng = c["views"], c["cfg"], c["rng"]
    cfg.math_mode = math_mode   # 0: ngp_detmath.h (the oracle's arithmetic), 1: the reference build's arithmetic (march_ref.cu)
    t_views, tens = util.views_to_device(views, c["keep"])
    t_bf = torch.from_numpy(np.ascontiguousarray(c["bitfield"])).cuda()
    t_cnt = torch.zeros(4, dtype=torch.int32, device="cuda")
    t_ri = torch.zeros(n_rays, dtype=torch.int32, device="cuda")
    t_rays = torch.zeros(n_rays, 6, dtype=torch.float32, device="cuda")
    t_ns = torch.zeros(n_rays, 2, dtype=torch.int32, device="cuda")
    t_co = torch.zeros(max_samples, 7, dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    rc = lib.ngp_nerf_generate_training_samples(stream, n_rays, 0, n_rays, rng[0], rng[1], C.byref(cfg), t_views.data_ptr(), len(views), t_bf.data_ptr(), max_samples,
                                                t_cnt.data_ptr(), t_ri.data_ptr(), t_rays.data_ptr(), t_ns.data_ptr(), t_co.data_ptr())
    assert rc == 0, lib.ngp_last_error()
    torch.cuda.synchronize()
    cnt = t_cnt.cpu().numpy().view(np.uint32)
    k = int(cnt[0])
    got = dict(n_kept=k, n_samples=int(cnt[1]), ray_indices=t_ri.cpu().numpy().view(np.uint32)[:k], rays=t_rays.cpu().numpy()[:k],
               numsteps=t_ns.cpu().numpy().view(np.uint32)[:k], coords=t_co.cpu().numpy())
    assert 0 < got["n_samples"] <= max_samples
    compare_generation(name, got, golden(name), exact=bool(math_mode))

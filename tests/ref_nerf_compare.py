"""Shared by tests/test_oracle_vs_reference_nerf.py (CPU oracle) and tests/test_gpu_vs_reference_nerf.py (this library's CUDA path):
loading the reference-kernel goldens (tests/golden/ref_nerf_<case>.npz) and the comparison of generated training samples."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT / "tools") not in sys.path:
    sys.path.insert(0, str(ROOT / "tools"))
import ref_nerf_cases as RC  # noqa: E402

GOLD = Path(__file__).resolve().parent / "golden"


def golden(name):
    p = GOLD / f"ref_nerf_{name}.npz"
    if not p.exists():
        pytest.skip(f"{p.name} not generated yet (tools/make_ref_nerf_golden.sh on a GPU box)")
    return np.load(p)


def compare_generation(name, want, g, exact=False):
    """`want`: dict(n_samples, ray_indices, rays, numsteps, coords) from the oracle — or from this library's CUDA generator, which equals
    the oracle bit for bit (tests/test_gpu_march.py); `g`: the reference kernel's outputs"""
    k_ref, ns_ref = int(g["gen_counters"][0]), int(g["gen_counters"][1])
    assert ns_ref <= RC.MAX_SAMPLES
    ref_numsteps = g["numsteps"].reshape(-1, 2)
    ref_rays = g["rays"].reshape(-1, 6)
    ref_coords = g["coords"].reshape(-1, 7)
    rmap = {int(r): j for j, r in enumerate(g["ray_indices"])}
    wmap = {int(r): j for j, r in enumerate(want["ray_indices"])}
    assert len(rmap) == k_ref
    # rays that produce samples: the same set up to rays grazing the occupied region
    both = set(rmap) & set(wmap)
    assert len(set(rmap) ^ set(wmap)) <= max(2, 0.003 * k_ref), (len(rmap), len(wmap))
    same_count, checked, rays_identical, coords_identical = 0, 0, 0, 0
    max_pos, max_dt, max_dir = 0.0, 0.0, 0.0
    for rid in sorted(both):
        rj, wj = rmap[rid], wmap[rid]
        # the unnormalised ray: identical arithmetic apart from FMA contraction in uv_to_ray / the lens undistortion
        assert np.allclose(ref_rays[rj], want["rays"][wj], rtol=0, atol=2e-6)
        rays_identical += int(ref_rays[rj].tobytes() == want["rays"][wj].tobytes())
        rn, rb = ref_numsteps[rj]
        wn, wb = want["numsteps"][wj]
        if rn != wn:
            assert abs(int(rn) - int(wn)) <= max(2, 0.02 * int(rn))   # a sample flipped at a voxel face, not a different march
            continue
        same_count += 1
        coords_identical += int(ref_coords[rb:rb + rn].tobytes() == want["coords"][wb:wb + wn].tobytes())
        if checked < 400:                                              # coordinates of a few hundred rays
            a, b = ref_coords[rb:rb + rn], want["coords"][wb:wb + wn]
            max_pos = max(max_pos, float(np.abs(a[:, :3] - b[:, :3]).max()))
            # warped dt lives in [0, 1]; in unit-cube scenes every step is the minimum step, whose warped value is 0 up to rounding
            max_dt = max(max_dt, float(np.abs(a[:, 3] - b[:, 3]).max()))
            max_dir = max(max_dir, float(np.abs(a[:, 4:] - b[:, 4:]).max()))
            checked += 1
    print(f"{name}: {k_ref} rays, {ns_ref} samples (oracle {want['n_samples']}); identical step counts on {same_count}/{len(both)} rays; "
          f"max |pos| diff {max_pos:.2e}, max |warped dt| diff {max_dt:.2e}, max |dir| diff {max_dir:.2e}; bit-identical ray records {rays_identical}, "
          f"bit-identical coordinate blocks {coords_identical}")
    if exact:
        # the reference build's arithmetic (NGP_MATH_REFERENCE): the north star's "bit-exact ray indices and sample counts"
        assert set(rmap) == set(wmap), (len(set(rmap) ^ set(wmap)), "rays differ in whether they produce samples")
        assert same_count == len(both), f"{len(both) - same_count} of {len(both)} rays differ in their sample count"
        assert ns_ref == want["n_samples"]
    assert same_count >= 0.985 * len(both)
    assert abs(ns_ref - want["n_samples"]) <= 0.003 * ns_ref
    # Warped positions live in [0, 1].  The reference's t drifts by a few 1e-6 per ray against exact arithmetic: every empty-voxel
    # skip goes through to_stepping_space's division (nerf_device.cuh:379-395), which --use_fast_math turns into an approximate
    # reciprocal, so t picks up a relative error of ~1e-7 per skip (measured: 4e-6 at the first sample after ~40 skips, 1e-5 at the
    # far end); with cone stepping (aabb_scale > 1) the skips go through __logf / __expf instead, to the same effect, and dt = t *
    # cone_angle inherits t's relative error.  3e-5 is a fiftieth of the finest step (sqrt(3)/1024 = 1.7e-3).
    assert max_pos < 3e-5 and max_dt < 3e-5 and max_dir < 2e-6

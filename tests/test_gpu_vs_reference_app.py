"""GPU: this library against the UNMODIFIED reference APPLICATION (pyngp from baseline/_ref, built by baseline/build_ref.sh with the GUI off
for sm_100), both driven through the same pyngp calls by tools/ref_app.py on nerf/fox (BASELINE config #2) — the application-level parity
and speed statement of the north star, live on the box the test runs on:

  * identical weights, identical views  ->  rendered pixels and PSNR agree (the reference's snapshot loaded here, and ours loaded there);
  * identical protocol, own training    ->  PSNR after 1000 steps agrees within run-to-run noise, the loss curves track each other;
  * samples/s of Testbed.train          ->  not below the reference's.

Skipped where baseline/_ref is absent (it is git-ignored and travels with gpurun; /root/reference is needed to build it)."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
REF = ROOT / "baseline" / "_ref"
HAVE_REF = bool(list(REF.glob("pyngp*.so"))) and (REF / "data" / "nerf" / "fox" / "transforms.json").exists()
STEPS = 1000


def run_app(out: Path, *args) -> dict:
    cmd = [sys.executable, str(ROOT / "tools" / "ref_app.py"), *args, "--scene", "fox", "--enc", "L16F2", "--render-repeats", "2", "--out", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and out.exists(), (r.stdout[-2000:], r.stderr[-2000:])
    rec = json.loads(out.read_text())
    npz = out.with_suffix(".npz")
    rec["arrays"] = dict(np.load(npz)) if npz.exists() else {}
    return rec


@pytest.fixture(scope="module")
def runs(tmp_path_factory):
    if not HAVE_REF:
        pytest.skip("baseline/_ref (reference pyngp + data/nerf/fox) is not on this box: baseline/build_ref.sh")
    d = tmp_path_factory.mktemp("refapp")
    ref = run_app(d / "ref.json", "--impl", "reference", "--jit", "1", "--train-mode", "Nerf", "--steps", str(STEPS))
    ours_on_ref = run_app(d / "ours_on_ref.json", "--impl", "ngp_b200", "--no-train", "--load-snapshot", str(d / "ref.ingp"))
    ours = run_app(d / "ours.json", "--impl", "ngp_b200", "--train-mode", "Nerf", "--steps", str(STEPS))
    ref_on_ours = run_app(d / "ref_on_ours.json", "--impl", "reference", "--jit", "1", "--no-train", "--load-snapshot", str(d / "ours.ingp"))
    out = dict(ref=ref, ours_on_ref=ours_on_ref, ours=ours, ref_on_ours=ref_on_ours)
    summary = {k: {q: v.get(q) for q in ("ms_per_step", "samples_per_sec", "psnr_mean", "render_1080p_ms_best", "counters")} for k, v in out.items()}
    print(json.dumps(summary))
    try:   # keep the record of this run next to the other profiles of the box
        (ROOT / "gpurun_out").mkdir(exist_ok=True)
        (ROOT / "gpurun_out" / "test_vs_reference_app.json").write_text(json.dumps(summary, indent=1))
    except OSError:
        pass
    return out


def pixel_stats(a, b):
    d = np.abs(np.asarray(a, np.float32) - np.asarray(b, np.float32))[..., :3]
    return float(d.mean()), float(np.quantile(d, 0.999)), float(d.max())


def test_rendered_pixels_and_psnr_agree_on_identical_weights(runs):
    """the reference's snapshot rendered by the reference and by this library: same held-out views, same camera, same epilogue"""
    ref, ours = runs["ref"], runs["ours_on_ref"]
    assert ours["n_test_views"] == ref["n_test_views"] == 10
    rel = abs(ours["psnr_mean"] - ref["psnr_mean"]) / ref["psnr_mean"]
    per_view = np.abs(np.array(ours["psnr_per_view"]) - np.array(ref["psnr_per_view"])) / np.array(ref["psnr_per_view"])
    mean, q999, mx = pixel_stats(ours["arrays"]["view0_small_linear"], ref["arrays"]["view0_small_linear"])
    cmean, cq, cmx = pixel_stats(ours["arrays"]["crop64_linear"], ref["arrays"]["crop64_linear"])
    print(f"PSNR ref {ref['psnr_mean']:.4f} ours {ours['psnr_mean']:.4f} rel {rel:.2e} (per view max {per_view.max():.2e}); "
          f"pixels (every 8th of view 0): mean |d| {mean:.2e}, 99.9 % {q999:.2e}, max {mx:.2e}; centre crop: mean {cmean:.2e}, max {cmx:.2e}")
    assert rel < 1e-3 and per_view.max() < 2e-3          # north star: PSNR within 1e-3 relative
    assert mean < 1e-3 and cmean < 1e-3                  # rendered pixels within 1e-3 (linear colour, [0, 1])
    assert q999 < 2e-2                                   # isolated pixels at depth discontinuities: one sample more or less along the ray


def test_the_reference_renders_our_snapshot_like_we_do(runs):
    ours, ref = runs["ours"], runs["ref_on_ours"]
    rel = abs(ours["psnr_mean"] - ref["psnr_mean"]) / ref["psnr_mean"]
    mean, q999, mx = pixel_stats(ours["arrays"]["view0_small_linear"], ref["arrays"]["view0_small_linear"])
    print(f"our snapshot: PSNR rendered here {ours['psnr_mean']:.4f}, by the reference {ref['psnr_mean']:.4f} (rel {rel:.2e}); mean |d| {mean:.2e}")
    assert rel < 1e-3 and mean < 1e-3


def test_trained_psnr_and_loss_track_the_reference(runs):
    """own training, same protocol.  The two runs draw the same rays only while their rays_per_batch controllers agree, the atomics order
    differs and this library accumulates the MLP in fp32 (the reference in fp16).  Measured on a B200 (profiles/r2/psnr_ab.md): the
    reference lands on 26.03-26.11 dB over six runs, this library on 25.99-26.08 dB over four since rays take their compacted slots in
    groups of 32 consecutive rays (before: 25.1-26.1, floaters in two held-out views in most runs) — means 1.5e-3 apart, spreads alike;
    the training loss at step ~1000 is 6-19 % above the reference's three samples.  On identical weights the two renderers agree to 3e-6
    relative (the test above).  The bounds here are three times the measured spread, not the north star's 1e-3."""
    ref, ours = runs["ref"], runs["ours"]
    rel = abs(ours["psnr_mean"] - ref["psnr_mean"]) / ref["psnr_mean"]
    print(f"PSNR@{STEPS}: reference {ref['psnr_mean']:.4f} dB, this library {ours['psnr_mean']:.4f} dB, relative difference {rel:.2e}")
    assert rel < 8e-3
    for k in ("1", "97", "497", "993"):
        a, b = ref["loss_curve"].get(k), ours["loss_curve"].get(k)
        if a is not None and b is not None:
            assert abs(a - b) <= 0.25 * max(a, b), (k, a, b)
    # the controller settles on the same workload
    assert abs(ours["counters"]["rays_per_batch"] - ref["counters"]["rays_per_batch"]) <= 0.15 * ref["counters"]["rays_per_batch"]


def test_training_and_render_are_not_slower_than_the_reference(runs):
    ref, ours = runs["ref"], runs["ours"]
    print(f"ms/step {ours['ms_per_step']:.3f} vs reference {ref['ms_per_step']:.3f}; samples/s {ours['samples_per_sec'] / 1e6:.1f} M vs {ref['samples_per_sec'] / 1e6:.1f} M; "
          f"1080p render {ours['render_1080p_ms_best']:.1f} ms vs {ref['render_1080p_ms_best']:.1f} ms (wall clock incl. the frame's D2H)")
    assert ours["samples_per_sec"] >= ref["samples_per_sec"]      # north star: >= reference instant-ngp samples/sec on 1 x B200 for nerf/fox

#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the NeRF hot path.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched through torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

Metric (BASELINE.json): NeRF training samples/s on the configuration it is quoted on — BASELINE config #2, nerf/fox (the
reference's own data/nerf/fox: 50 JPEGs 1080x1920, aabb_scale 4, OpenCV lens; every 5th frame held out as scripts/scenes.py does),
hash grid L=16 F=2 T=2^19, 64-wide MLPs (density 1 hidden, rgb 2 hidden), 2^18-sample batches.  The scene travels to the GPU box
inside baseline/_ref/ (git-ignored, shipped by gpurun; baseline/build_ref.sh puts it there); when it is absent the workload
falls back to the synthetic 100-view 800x800 ball (`--scene ball`, BASELINE config #3's stand-in: NeRF-synthetic is not in the
reference tree).  One "step" = one Testbed.train(2^18): occupancy-grid maintenance on the reference's schedule, training-ray
generation + marching, inference, loss + compaction, fused forward/backward, optimizer.

value  : compacted samples trained per second, all ranks, dataset resident in HBM (device-timed with CUDA events).
e2e    : the same through the public pyngp-style API while one training image per step is streamed from pinned host
         memory (H2D inside the timed region: update_image_async uploads on a copy stream into a staging buffer and the training
         stream moves the frame into place before the loss kernel) and the step's counters + loss are read back (D2H).
timing : steady state — 700 untimed set-up steps (--pretrain) + W warm-up steps come first: the per-step workload (rays per batch,
         samples per ray) only settles once the scene has formed (SURVEY §8d asks for steps 500-1000); then exactly K timed steps.
extras : roofline (k_nerf_train), cpu_baseline (oracle port, rank 0, N = 1), clocks (NVML during the timed region), render
         (1920x1080 Mrays/s), quality (PSNR of the trained model), phase_ms_per_step (CUDA events per phase, separate pass),
         reference_gpu (the UNMODIFIED reference application, baseline/_ref/pyngp*.so, driven by tools/ref_app.py on the same
         box, same scene, same protocol, in a subprocess after this arm's timing; N = 1 only).
N > 1  : weak scaling — every rank trains its own 2^18-sample batch on its (interleaved) shard of the global ray batch; the all-reduce of the
         flat fp16 gradient buffer and of the counter block happen inside Testbed.train over NCCL (ngp_testbed_init_dp); torch.distributed only
         hands the NCCL ids to the ranks and takes the max over ranks of the timings.
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

BATCH = 1 << 18
N_VIEWS, RES = 100, 800
WORKLOADS = {
    "fox": "nerf/fox (reference data/nerf/fox: 40 training views 1080x1920 of 50, every 5th held out; aabb_scale 4, cone stepping, OpenCV lens) "
           "L16 F2 T2^19 mlp64x(1,2) batch 2^18",
    "ball": "nerf-synthetic-ball-100x800x800 L16 F2 T2^19 mlp64x(1,2) batch 2^18 aabb_scale 1",
}
FOX_DIR = ROOT / "baseline" / "_ref" / "data" / "nerf" / "fox"
REF_PYNGP = list((ROOT / "baseline" / "_ref").glob("pyngp*.so")) if (ROOT / "baseline" / "_ref").exists() else []
# the unmodified reference application on a B200, same protocol (tools/ref_app.py), as recorded in profiles/r2/ — used only to label
# the line when baseline/_ref is not on the box (then `reference_gpu.source` says "recorded")
RECORDED_REFERENCE_GPU = {"fox": {"ms_per_step": 2.556, "samples_per_sec": 109.0e6, "psnr_mean": 26.032, "render_1080p_ms_best": 33.17, "source": "recorded: profiles/r2/refapp_fox.md"}}


def pkg():
    return importlib.import_module("instant-ngp_b200")


def syn():
    return importlib.import_module("instant-ngp_b200.synthetic")


def ref_app():
    if str(ROOT / "tools") not in sys.path:
        sys.path.insert(0, str(ROOT / "tools"))
    return importlib.import_module("ref_app")


def default_scene() -> str:
    return "fox" if (FOX_DIR / "transforms.json").exists() else "ball"


# ---------------------------------------------------------------------------------------------------------------------
# dram__bytes_read.sum + dram__bytes_write.sum of one k_nerf_train launch (ncu --set full): nerf/fox profiles/r2/r2f_train_fox_ncu.md,
# the synthetic scene profiles/r1e_end_of_round.md
K_NERF_TRAIN_DRAM_BYTES = {"fox": 61.01e6 + 0.82e6, "ball": 54.52e6 + 0.20e6}
K_NERF_TRAIN_DRAM_SOURCE = {"fox": "profiles/r2/r2n_train_fox_ncu.md", "ball": "profiles/r1e_end_of_round.md"}


class ClockSampler:
    """samples SM clocks / throttle reasons of one GPU while the timed region runs (B200_PROFILING.md's clocks line).  NVML in a
    thread every 5 ms (a 150 ms timed region still gets ~30 samples); `nvidia-smi -lms` as the fallback when NVML is unusable."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}

    def __init__(self, gpu_index: int):
        self.gpu, self.sm, self.mx, self.reasons = gpu_index, [], [], set()
        self.stop, self.thread, self.proc = threading.Event(), None, None

    def _nvml_loop(self, nv, h):
        while not self.stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in self.REASONS.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop.wait(0.005)

    def _smi_loop(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                self.sm.append(float(r[0]))
                self.mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass

    def __enter__(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
        except Exception:
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            try:
                self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.thread = threading.Thread(target=self._smi_loop, daemon=True)
                self.thread.start()
            except Exception:
                self.proc = None
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        elif self.thread:
            self.thread.join(timeout=1)

    def summary(self) -> dict:
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": float(max(self.mx)) if self.mx else None, "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


def build_testbed(scene: str, n_views: int = N_VIEWS, res: int = RES, device: int = 0):
    """returns (testbed, host copies of the training images for the e2e leg, scene info)"""
    P, S = pkg(), syn()
    tb = P.Testbed(P.TestbedMode.Nerf, device=device)
    if scene == "fox":
        R = ref_app()
        split, _ = R.fox_split()
        tb.load_training_data(str(split["train"]))
        NL = importlib.import_module("instant-ngp_b200.nerf_loader")
        host = [NL.read_image_bytes_rgba(im["path"], tb.dataset["white_transparent"], tb.dataset["black_transparent"]) for im in tb.dataset["images"]]
        info = dict(n_views=len(host), test_split=str(split["test"]))
    else:
        imgs, cams, focal = S.make_dataset(n_images=n_views, width=res, height=res)
        S.load_into_testbed(tb, imgs, cams, focal, aabb_scale=1)
        host = [imgs[i] for i in range(imgs.shape[0])]
        info = dict(n_views=n_views, cams=cams, focal=focal, imgs=imgs)
    tb.reload_network_from_json(S.base_config(16, 2, 19))
    return tb, host, info


def peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "source": "fallback"}


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the CPU port of the same path (tiny-cuda-nn has no CPU implementation) on ALL host cores, on a bounded unit of
# the same workload (4 096 rays / the samples they generate): ray generation + occupancy march and compositing / loss /
# compaction = the C oracle (oracle/ngp_oracle.c) over ray chunks on a thread pool (ctypes releases the GIL); network inference over
# every generated sample, forward + backward over the compacted ones and Adam = oracle/ngp_net_cpu.c (fp32, OpenMP).
# ---------------------------------------------------------------------------------------------------------------------
class CpuSampler:
    def __init__(self, scene: str):
        sys.path.insert(0, str(ROOT / "tests"))
        import util
        from oracle import march_oracle as M
        from oracle import net_cpu

        self.util, self.M, self.net_cpu = util, M, net_cpu
        S = syn()
        fox = scene == "fox"
        aabb_scale = 4 if fox else 1
        imgs, cams, focal = S.make_dataset(n_images=8, width=200, height=200, radius=1.6 if fox else 1.3)
        self.cfg = util.make_train_cfg(aabb_scale=aabb_scale)
        self.views, self.keep = util.make_views(imgs, cams, focal, lens=(0.0578421, -0.0805099, -0.000980296, 0.00015575) if fox else None)
        self.bf = util.sphere_bitfield(radius=0.3, max_cascade=self.cfg.max_cascade)
        _, self.L = util.make_desc(n_levels=16, F=2, log2_T=19, aabb_scale=aabb_scale)
        params = util.random_params(self.L, seed=0, trained_like=True)
        self.net = net_cpu.NetCpu(self.L, params)
        self.m1 = np.zeros_like(self.net.params)
        self.m2 = np.zeros_like(self.net.params)
        self.rng = M.pcg32_seed(1337)
        self.cores = os.cpu_count() or 1
        self.scene = scene

    def _chunk(self, args):
        """march + (after the network) loss of one chunk of rays: returns the generator's outputs"""
        lo, n, n_global = args
        return self.M.generate_training_samples(n, lo, n_global, self.rng, self.cfg, self.views, len(self.views), self.bf, n * 1024)

    def step(self, n_rays: int):
        """one training step over n_rays rays: (compacted samples, seconds, description)"""
        from concurrent.futures import ThreadPoolExecutor

        util, M = self.util, self.M
        n_chunks = min(self.cores, max(1, n_rays // 32))
        per = (n_rays + n_chunks - 1) // n_chunks
        jobs = [(lo, min(per, n_rays - lo), n_rays) for lo in range(0, n_rays, per)]
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=n_chunks) as ex:
            gens = list(ex.map(self._chunk, jobs))
            t_march = time.perf_counter() - t0
            # inference over every generated sample (the reference's schedule, testbed_nerf.cu:3233-3235)
            coords = np.concatenate([g["coords"][:g["n_samples"]] for g in gens]) if gens else np.zeros((0, 7), np.float32)
            t1 = time.perf_counter()
            net_out = self.net.forward(coords).astype(np.float16)
            t_inf = time.perf_counter() - t1
            # compositing, loss, compaction per chunk
            t2 = time.perf_counter()
            offs = np.cumsum([0] + [g["n_samples"] for g in gens])

            def loss_chunk(k):
                g = gens[k]
                ns, kept = g["n_samples"], g["n_kept"]
                if kept == 0:
                    return np.zeros((0, 7), np.float32), np.zeros((0, 4), np.float16)
                numsteps = g["numsteps"].copy()
                co = np.zeros((ns, 7), dtype=np.float32)
                dl = np.zeros((ns, 4), dtype=np.float16)
                out = np.ascontiguousarray(net_out[offs[k]:offs[k] + ns])
                comp = M.lib().orc_compute_loss(kept, n_rays, self.rng[0], self.rng[1], C.byref(self.cfg), C.addressof(self.views), len(self.views), out.ctypes.data, ns,
                                                g["ray_indices"].ctypes.data, g["rays"].ctypes.data, numsteps.ctypes.data, g["coords"].ctypes.data, co.ctypes.data,
                                                dl.ctypes.data, None, 0.02)
                comp = min(comp, ns)
                return co[:comp], dl[:comp]

            parts = list(ex.map(loss_chunk, range(len(gens))))
            t_loss = time.perf_counter() - t2
        cc = np.concatenate([p[0] for p in parts])
        dd = np.concatenate([p[1] for p in parts]).astype(np.float32)
        t3 = time.perf_counter()
        _, grads = self.net.forward_backward(cc, dd)
        t_fb = time.perf_counter() - t3
        t4 = time.perf_counter()
        self.n_steps = getattr(self, "n_steps", 0) + 1
        self.net.adam_step(grads, self.m1, self.m2, self.n_steps)
        t_opt = time.perf_counter() - t4
        total = time.perf_counter() - t0
        desc = (f"CPU port on {self.cores} host threads, one step over {n_rays} rays of a {'fox-like (aabb_scale 4, cone stepping, OpenCV lens)' if self.scene == 'fox' else 'unit-cube'} "
                f"scene: C march ({t_march * 1e3:.0f} ms) + fp32 OpenMP network inference of {len(coords)} samples ({t_inf * 1e3:.0f} ms) + C loss/compaction "
                f"({t_loss * 1e3:.0f} ms) + forward/backward of {len(cc)} compacted samples ({t_fb * 1e3:.0f} ms) + OpenMP Adam over {self.net.params.shape[0]} parameters "
                f"({t_opt * 1e3:.0f} ms), L16F2T19")
        return len(cc), total, desc


def cpu_training_sample(scene: str, n_rays: int = 4096):
    s = CpuSampler(scene)
    s.step(256)   # page in, spin the pools up
    n, sec, desc = s.step(n_rays)
    return n / sec, desc, s.cores


def run_reference_arm(args, scene: str) -> None:
    """--impl reference: the CPU port on all host cores (tiny-cuda-nn has no CPU path), rank 0 only.  Every step is a bounded unit of the
    workload (4 096 rays and the samples they generate); the unit shrinks if W + K steps would not finish within ~2.5 minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sampler = CpuSampler(scene)
    n_rays = 4096
    n_steps = args.warmup + args.steps
    vals, desc, last_n = [], "", 1
    budget_s = 150.0
    t_start = time.perf_counter()
    for i in range(n_steps):
        n, spent, desc = sampler.step(n_rays)
        if i >= args.warmup:
            vals.append(n / spent)
            last_n = n
        remaining = n_steps - 1 - i
        left = budget_s - (time.perf_counter() - t_start)
        if remaining > 0 and spent * remaining > max(left, 1.0) and n_rays > 64:
            n_rays = max(64, int(n_rays * max(left, 1.0) / (spent * remaining)) // 32 * 32)
    v = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "nerf_training_samples_per_sec", "value": v, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": last_n / v * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[scene], "note": "tiny-cuda-nn has no CPU implementation (SURVEY.md §0.2); this is the CPU port of the same path on all host cores, a bounded unit per step"},
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": sampler.cores, "kind": "port", "sample": desc + " (last step's size; shrunk to keep W + K steps within ~2.5 min)"},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
def reference_gpu_arm(scene: str, steps: int = 1000) -> dict:
    """the UNMODIFIED reference application (baseline/_ref/pyngp*.so, built by baseline/build_ref.sh) on this box, same scene, same
    protocol (tools/ref_app.py: steady-state ms/step over steps 500..1000, PSNR on the held-out views, 1920x1080 render), in a
    subprocess so that nothing of it shares a process with this library"""
    if not REF_PYNGP:
        rec = dict(RECORDED_REFERENCE_GPU.get(scene, {}))
        rec.setdefault("source", "unavailable: baseline/_ref/pyngp*.so is not on this box (baseline/build_ref.sh builds it where /root/reference exists)")
        return rec
    out = ROOT / "gpurun_out" / f"bench_reference_gpu_{scene}.json"
    out.parent.mkdir(parents=True, exist_ok=True)
    cmd = [sys.executable, str(ROOT / "tools" / "ref_app.py"), "--impl", "reference", "--scene", scene, "--enc", "L16F2", "--jit", "1", "--train-mode", "Nerf",
           "--steps", str(steps), "--out", str(out)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        j = json.loads(out.read_text())
        for ext in (".ingp", ".npz"):
            try:
                out.with_suffix(ext).unlink()
            except OSError:
                pass
        keep = ("ms_per_step", "samples_per_sec", "rays_per_sec", "samples_per_sec_nominal", "steady_window", "counters", "psnr_mean", "n_test_views",
                "render_1080p_ms_best", "render_1080p_mrays_per_sec", "jit", "train_mode", "enc")
        rec = {k: j[k] for k in keep if k in j}
        rec["source"] = "measured in this run: tools/ref_app.py --impl reference (pyngp from baseline/_ref, unmodified reference, GUI off, sm_100)"
        return rec
    except Exception as e:  # the reference arm never takes this arm's line down
        return {"source": f"failed: {type(e).__name__}: {e}"[:300]}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=300)
    ap.add_argument("--pretrain", type=int, default=700, help="untimed set-up steps before the warm-up: the metric is STEADY-STATE training throughput "
                    "(SURVEY 8d: steps 500-1000), and the per-step workload (rays, samples per ray) only settles once the scene has formed")
    ap.add_argument("--impl", default="ngp_b200", choices=["ngp_b200", "reference"])
    ap.add_argument("--scene", default=None, choices=["fox", "ball"], help="default: fox when baseline/_ref/data/nerf/fox is on the box, else the synthetic ball")
    ap.add_argument("--views", type=int, default=N_VIEWS)
    ap.add_argument("--res", type=int, default=RES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fields", action="store_true", help="skip the image / SDF primitive micro-benchmark (extra key `fields`, N = 1)")
    ap.add_argument("--no-reference-gpu", action="store_true", help="skip the reference application's run on this GPU (extra key `reference_gpu`, N = 1)")
    ap.add_argument("--no-overlap", action="store_true", help="disable the side-stream prefetch of the next step's sample generation")
    ap.add_argument("--chunk", type=int, default=0, help="ray-ordered inference chunk (4 or 8)")
    ap.add_argument("--full-inference", action="store_true", help="evaluate every generated sample like the reference schedule")
    ap.add_argument("--no-render", action="store_true", help="skip the 1920x1080 render timing (reported under the extra key `render`)")
    ap.add_argument("--option", action="append", default=[], help="name=value passed to Testbed._set before training")
    args = ap.parse_args()
    scene = args.scene or default_scene()
    if args.impl == "reference":
        run_reference_arm(args, scene)
        return
    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch through torch.distributed.run"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    __import__("__graft_entry__").build() if not (ROOT / "instant-ngp_b200" / "libngp_b200.so").exists() else None
    P = pkg()
    lib = P.load_library()
    tb, host_images, info = build_testbed(scene, args.views, args.res, device=local_rank)
    n_views = info["n_views"]
    if args.no_overlap:
        tb._set("nerf.training.overlap_sample_generation", 0.0)
    if args.chunk:
        tb._set("nerf.training.inference_chunk", float(args.chunk))
    if args.full_inference:
        tb._set("nerf.training.full_inference", 1.0)
    for kv in args.option:
        k, v = kv.split("=")
        tb._set(k, float(v))
    if world > 1:
        # torch.distributed is the plumbing that hands rank 0's NCCL id to the other ranks; the step's collectives are the library's own
        ids = [tb.dp_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        tb.init_data_parallel(rank, world, ids[0])

    def step():
        tb.train(BATCH)   # N > 1: counters + gradient all-reduce inside (ngp_testbed_init_dp)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- set-up: train until the scene has formed (untimed, not part of the warm-up count), then the warm-up proper
    for _ in range(args.pretrain):
        step()
    for _ in range(args.warmup):
        step()
    sync_all()

    # ---- timed region: device-resident dataset
    launches0 = lib.ngp_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    samples = 0
    rays = 0
    pre = 0
    with ClockSampler(local_rank) as clk:
        sync_all()
        e0.record()
        for _ in range(args.steps):
            step()
            c = tb.counters()
            samples += c["measured_batch_size"]
            rays += c["rays_per_batch"]
            pre += c["measured_batch_size_before_compaction"]
        e1.record()
        sync_all()
    ms_total = e0.elapsed_time(e1)
    launches = lib.ngp_launch_count() - launches0
    clocks = clk.summary()

    # ---- per-phase device times (CUDA events around every phase, on the stream each phase runs on).  Separate, untimed
    # pass: collecting event times needs a stream synchronisation per step, which the timed region above does not pay.
    tb.set_profiling(True)
    for _ in range(max(8, min(args.steps, 32))):
        step()
    sync_all()
    phases = tb.phase_ms()
    tb.set_profiling(False)

    # ---- e2e: one training view streamed from pinned host memory per step + counters/loss read back
    pinned = [torch.from_numpy(np.ascontiguousarray(im)).pin_memory() for im in host_images]   # the image set in pinned host memory, one view re-sent per step
    h2d_bytes = pinned[0].numel() * pinned[0].element_size()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_samples = 0
    sync_all()
    f0.record()
    for i in range(args.steps):
        tb.update_image_async(i % n_views, pinned[i % n_views].data_ptr())   # H2D of this step's training view
        step()
        e2e_samples += tb.counters()["measured_batch_size"]        # D2H: counters (16 B) each step, loss every 16th
        _ = tb.loss
    f1.record()
    sync_all()
    ms_e2e = f0.elapsed_time(f1)

    # ---- render: 1920x1080 from training view 0's pose; N > 1: row tiles sharded over the ranks + exchange of tiles
    render = None
    if not args.no_render:
        W, H = 1920, 1080
        if scene == "fox":
            tb.set_camera_to_training_view(0)
            tb.render_with_lens_distortion = False     # what the reference's Python render() does (python_api.cu:212)
            cam, focal, center = tb._camera.render_args(W, H)
        else:
            cam, center = syn().sphere_cameras(8, radius=1.3)[3], (0.5, 0.5)
            focal = 0.5 * H / np.tan(0.5 * np.deg2rad(40.0))
        rgba = torch.zeros(H, W, 4, device="cuda")
        depth = torch.zeros(H, W, device="cuda")

        def frame():
            if world > 1:
                tb.render_device_sharded(W, H, cam, focal, rgba.data_ptr(), depth.data_ptr(), center)
            else:
                tb.render_device(W, H, cam, focal, rgba.data_ptr(), depth.data_ptr(), center)

        for _ in range(3):
            frame()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        r0.record()
        for _ in range(10):
            frame()
        r1.record()
        sync_all()
        render_ms = torch.tensor([r0.elapsed_time(r1) / 10], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(render_ms, op=dist.ReduceOp.MAX)
        render = {"ms_per_frame": float(render_ms), "mrays_per_sec": W * H / float(render_ms) / 1e3, "resolution": [W, H],
                  "coverage": float((rgba[..., 3] > 0.5).float().mean()), "sharding": f"{world} row tiles + ncclBroadcast of every tile" if world > 1 else "none",
                  "min_transmittance": 0.01, "pose": "training view 0" if scene == "fox" else "orbit camera 3"}

    # ---- reduce over ranks: time = max, work = sum
    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device="cuda")
    w = torch.tensor([samples, e2e_samples, rays, pre], dtype=torch.float64, device="cuda")
    ph = torch.tensor([phases[k] for k in tb.PHASES], dtype=torch.float64, device="cuda")
    ph_all = None
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w)
        gathered = [torch.zeros_like(ph) for _ in range(world)]
        dist.all_gather(gathered, ph)
        ph_all = torch.stack(gathered).cpu().numpy()
    ms_total, ms_e2e = t.tolist()
    samples, e2e_samples, rays, pre = w.tolist()

    if rank == 0:
        value = samples / (ms_total * 1e-3)
        e2e = e2e_samples / (ms_e2e * 1e-3)
        pk = peaks()
        # dominant kernel: the fused forward/backward (k_nerf_train).  Algorithmic bytes per trained sample (SURVEY.md §8d,
        # DESIGN.md): 512 B hash-table reads + 512 B fp16 gradient reductions (1024 B read-modify-write traffic) + 28 B
        # coordinates + 8 B dL/dout = 1572 B.
        n_fb = max(phases["steps"], 1)
        fb_ms = phases["forward_backward"] / n_fb
        alg_bytes = 1572.0 * BATCH
        achieved = alg_bytes / (fb_ms * 1e-3) / 1e9 if fb_ms > 0 else None
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        roofline = {"bound": "hbm", "kernel": "k_nerf_train (+k_mlp_grads_finalize)", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": (achieved / pk["hbm_gbs"]) if achieved else None, "traffic": K_NERF_TRAIN_DRAM_BYTES[scene], "peak_source": pk["source"], "ms_per_launch": fb_ms,
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "traffic_source": K_NERF_TRAIN_DRAM_SOURCE[scene] + " (per launch)",
                    # what binds the kernel: the SM's load/store path takes SCATTERED accesses (every lane another 128-byte line, L2 hits) at 1.01 cycles
                    # per lane for gathers and 1.50 for fp16x2 reductions, whatever the width (4 or 8 bytes) and the occupancy — measured on this GPU
                    # with tools/microbench/lsu_scatter.cu (profiles/r2/r2l_lsu_scatter.jsonl).  Lane accesses per sample from ncu's L1 sector counts
                    # on nerf/fox (profiles/r2/r2n_train_fox_ncu.md): 16 levels x 8 corners, x-neighbour pairs sharing one access when adjacent and
                    # aligned, lanes of a warp reading the same sector counted once, runs of samples in one cell reduced once (mlp_train.cuh AGG).
                    "lsu_floor": {"gather_cycles_per_lane": 1.01, "red_cycles_per_lane": 1.50, "lane_accesses_per_sample": {"gather": 77.6, "red": 79.9}, "sm_count": 148,
                                  "sm_mhz": sm_mhz, "source": "profiles/r2/r2l_lsu_scatter.jsonl, profiles/r2/r2n_train_fox_ncu.md",
                                  "floor_ms": (77.6 * 1.01 + 79.9 * 1.50) * BATCH / 148 / (sm_mhz * 1e3),
                                  "frac": ((77.6 * 1.01 + 79.9 * 1.50) * BATCH / 148 / (sm_mhz * 1e3)) / fb_ms if fb_ms > 0 else None,
                                  "mlp_phase_ms": 0.097, "mlp_phase_source": "profiles/r2/r2h_mlp_phase.json (tensor pipe 11.5 % while it runs)"},
                    "note": "the 26 MB fp16 table and its gradient table are L2-resident on B200: DRAM traffic is an eighth of the algorithmic bytes; the kernel runs at the SM load/store path's scattered-access rate (lsu_floor) plus the MLP phase"}
        # second kernel of the step by time: the sample generator.  Algorithmic bytes: the 28-byte coordinate record of every generated sample
        # (SURVEY 8d "1 march step": the bit test itself is L1/L2 resident); it is latency bound (a serial log/exp recurrence per ray), which
        # is what the fraction says.
        gen_ms = phases["sample_generation"] / n_fb
        gen_bytes = 28.0 * pre / args.steps / world
        roofline_gen = {"bound": "hbm", "kernel": "k_generate_training_samples", "achieved": gen_bytes / (gen_ms * 1e-3) / 1e9 if gen_ms > 0 else None, "peak": pk["hbm_gbs"],
                        "unit": "GB/s", "frac": (gen_bytes / (gen_ms * 1e-3) / 1e9 / pk["hbm_gbs"]) if gen_ms > 0 else None, "traffic": None, "ms_per_launch": gen_ms,
                        "algorithmic_bytes_per_launch": gen_bytes, "note": "latency bound: per ray a serial recurrence t -> t + dt(t) through log/exp; see gen_kernel.cuh"}
        line = {
            "metric": "nerf_training_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
            "data": "reference data/nerf/fox (real photographs shipped in baseline/_ref), random-init weights" if scene == "fox" else "synthetic",
            "config": {"workload": WORKLOADS[scene] if scene == "fox" or (args.views, args.res) == (N_VIEWS, RES) else f"synthetic ball {args.views}x{args.res}^2 L16F2T19 batch 2^18",
                       "batch_per_gpu": BATCH,
                       "l2_policy": "inputs larger than L2: %.2f GB image set + 340 MB parameter/optimizer state per step, no explicit flush" % (sum(x.numel() * x.element_size() for x in pinned) / 1e9),
                       "parallelism": f"dp{world}", "pretrain_steps": args.pretrain,
                       "timed_steps": f"{args.pretrain + args.warmup}..{args.pretrain + args.warmup + args.steps} of a from-scratch training run",
                       "march_arithmetic": "reference build (--use_fast_math expression trees, march_ref.cu)" if tb._get("nerf.training.math_mode") == 1 else "deterministic (ngp_detmath.h)",
                       "compaction_order": {0: "groups of 32 consecutive rays, shuffled per step", 1: "one atomic per ray", 2: "ascending ray id"}.get(int(tb._get("nerf.training.compaction_order")), "?"),
                       "e2e_upload": "copy stream -> staging buffer -> pixel buffer before the loss kernel (update_image_async)"},
            "rays_per_sec": rays / (ms_total * 1e-3),
            "per_step": {"rays": rays / args.steps / world, "samples_before_compaction": pre / args.steps / world, "samples_compacted": samples / args.steps / world},
            "phase_ms_per_step": {k: v / n_fb for k, v in phases.items() if k != "steps"},
            "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 20, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "roofline_generator": roofline_gen,
        }
        if ph_all is not None:
            line["phase_ms_per_step_by_rank"] = {k: [float(x) / n_fb for x in ph_all[:, i]] for i, k in enumerate(tb.PHASES)}
        if render is not None:
            line["render"] = render
        if world == 1 and not args.no_render:
            # quality of the model the timed steps produced
            if scene == "fox":
                R = ref_app()
                tb.shall_train = False
                tb.background_color = [0.0, 0.0, 0.0, 1.0]
                tb.snap_to_pixel_centers = True
                tb.nerf.render_min_transmittance = 1e-4
                steps_done = int(tb.training_step)
                tb.load_training_data(info["test_split"])
                ps = []
                for i in range(len(tb.dataset["images"])):
                    v = tb.training_view(i)
                    wv, hv = (int(x) for x in v["resolution"])
                    tb.set_camera_to_training_view(i)
                    img = tb.render(wv, hv, 1, True)       # reference-shaped call: no lens, like the reference's Python render()
                    ps.append(R.psnr_srgb(img, R.load_gt_srgb(Path(tb.dataset["images"][i]["path"])))[0])
                line["quality"] = {"psnr_db_held_out_views": float(np.mean(ps)), "n_views": len(ps), "after_steps": steps_done,
                                   "protocol": "scripts/run.py:257-317 with scripts/scenes.py test_every = 5; reference-shaped render() (its Python render ignores the lens: 26.0 dB for the reference too)"}
            else:
                got = tb.render(args.res, args.res, info["cams"][0], info["focal"])
                mse = float(np.mean((np.clip(got[..., :3], 0, 1) - info["imgs"][0][..., :3]) ** 2))
                line["quality"] = {"psnr_db_train_view_0": 10.0 * np.log10(1.0 / max(mse, 1e-12)), "after_steps": int(tb.training_step)}
        if world == 1 and not args.no_reference_gpu:
            del tb
            torch.cuda.empty_cache()
            line["reference_gpu"] = reference_gpu_arm(scene)
            rg = line["reference_gpu"]
            if rg.get("samples_per_sec"):
                line["vs_reference_gpu"] = {"samples_per_sec_ratio": value / rg["samples_per_sec"], "ms_per_step_ratio": rg["ms_per_step"] / (ms_total / args.steps),
                                            "definition": "this repo / reference application, same box, same scene, same batch; > 1 = faster than the reference"}
        if world == 1 and not args.no_fields:
            # BASELINE configs #1 (image, 64 K-sample batch) and #4 (SDF, 2^20 samples): encoding + MLP training step of the same kernels
            # (tools/bench_fields.py, CUDA-event timed), so that they are part of the driver-visible line
            try:
                r = subprocess.run([sys.executable, str(ROOT / "tools" / "bench_fields.py")], capture_output=True, text=True, timeout=300)
                rows = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
                pick = [x for x in rows if (x["case"], x["n"]) in (("image", 1 << 16), ("sdf", 1 << 20))]
                line["fields"] = [{"config": "image 64K-sample batch (configs/image/base.json shape)" if x["case"] == "image" else "sdf 2^20-sample batch (configs/sdf/base.json shape)",
                                   "train_step_ms": x["fwd_bwd_opt_ms"], "train_msamples_per_s": x["step_msamples_per_s"], "inference_msamples_per_s": x["fwd_msamples_per_s"],
                                   "fwd_bwd_algorithmic_gbs": x["fwd_bwd_algorithmic_gbs"]} for x in pick]
            except Exception as e:
                line["fields"] = {"error": f"{type(e).__name__}: {e}"[:200]}
        if world == 1 and not args.no_cpu_baseline:
            sps, desc, cores = cpu_training_sample(scene)
            line["cpu_baseline"] = {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port", "sample": desc}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

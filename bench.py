#!/usr/bin/env python
"""Benchmark of the SuperPoint extract + match hot path (BASELINE.json metric:
frames/sec extract+match @ 752x480 on 1/2/4/8 B200, and % of the conv roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path over one batch of B synthetic 752x480 frames
of one camera stream: extract every frame, then mutual-NN match every frame
against the previous frame of the stream (SPFE_MATCH_PREV).  `value` is timed
with CUDA events on the library's stream with the frames already in HBM;
`e2e` goes through the host-pointer C-ABI calls (spfe_submit / spfe_wait) with
H2D / D2H copies inside the timed region.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WEIGHTS = os.path.join(ROOT, "tests", "golden", "superpoint_v1.spw")
METRIC = "frames/sec SuperPoint extract+match @ 752x480"
FLOP_PER_PIXEL = 169608.0          # conv stack, 2*MAC, SURVEY.md §8d
SMI_FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


def workload_name(W, H, nf):
    return (f"synthetic {W}x{H} u8 camera stream per GPU (BASELINE configs[1] geometry + configs[2] matching): "
            f"extract + mutual-NN match to previous frame, nfeatures {nf}")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops_sustained=d["bf16_tflops_sustained"], tflops_burst=d["bf16_tflops"], hbm_gbs=d["hbm_gbs"], src="measured")
    return dict(tflops_sustained=1400.0, tflops_burst=1590.0, hbm_gbs=6650.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed regions."""

    def __init__(self, gpu_index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={SMI_FIELDS}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, windows):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        rows = [r for ts, r in self.rows if any(a <= ts <= b + 0.2 for a, b in windows)] or [r for _, r in self.rows]
        sm, mx, reasons, pw = [], [], set(), []
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def make_pool(H, W, B, n_batches, rank):
    """n_batches x B frames: consecutive views of drifting scenes (one camera stream per rank)."""
    from sp_orb_slam_b200 import synth
    uniq = min(48, n_batches * B)
    frames = synth.make_stream(H, W, uniq, seed=1234 + 100 * rank, n_shapes=400)
    idx = np.arange(n_batches * B) % uniq
    return frames[idx].reshape(n_batches, B, H, W)


def cpu_reference_step(weights, frames, nf, use_ref, prev=None):
    """Reference CPU path on a few frames: network (reference's own compiled SPFrontend when oracle/_ref exists,
    else the torch restatement) + post-processing + BFMatcher-equivalent matching against the previous frame.
    Returns the last frame's result so the next step can match against it (every frame is extracted once)."""
    from oracle import sp_oracle as O
    for f in frames:
        if use_ref:
            from oracle import ref_frontend as R
            fwd = R.forward(weights, f)
        else:
            fwd = O.frontend_forward(weights, f)
        out = O.postprocess(fwd, f.shape[0], f.shape[1], nf)
        if prev is not None:
            O.match_mutual_nn(out["desc"], prev["desc"])
        prev = out
    return prev


def cpu_baseline(H, W, nf, n_frames=64):   # one bench step worth of frames: 10-20 s of host work
    import torch
    from oracle import ref_frontend as R, sp_oracle as O, weights as OW
    from sp_orb_slam_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    w = OW.read_spw(WEIGHTS)
    O.build_post()
    frames = synth.make_stream(H, W, n_frames + 1, seed=1234, n_shapes=400)
    use_ref = R.available()
    prev = cpu_reference_step(w, frames[:1], nf, use_ref)   # warm-up; its result is the first "previous frame"
    t0 = time.perf_counter()
    cpu_reference_step(w, frames[1:], nf, use_ref, prev)
    dt = time.perf_counter() - t0
    return {"value": n_frames / dt, "unit": "frames/s", "cores": cores, "kind": "reference" if use_ref else "port",
            "sample": f"{n_frames} synthetic {W}x{H} frames, extract + mutual-NN match to the previous frame, fp32 libtorch CPU "
                      f"({'reference SPFrontend compiled from its own sources' if use_ref else 'torch restatement'}) + C post-processing"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path, timed on the host cores."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import torch
    from oracle import ref_frontend as R, sp_oracle as O, weights as OW
    from sp_orb_slam_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    w = OW.read_spw(WEIGHTS)
    O.build_post()
    use_ref = R.available()
    per_step = 2
    frames = synth.make_stream(args.height, args.width, per_step * 4, seed=1234, n_shapes=400)
    prev = None
    for i in range(args.warmup):
        prev = cpu_reference_step(w, frames[:per_step], args.nf, use_ref, prev)
    t0 = time.perf_counter()
    for i in range(args.steps):
        o = (i % 4) * per_step
        prev = cpu_reference_step(w, frames[o:o + per_step], args.nf, use_ref, prev)   # every frame extracted once, matched to its predecessor
    dt = time.perf_counter() - t0
    fps = args.steps * per_step / dt
    kind = "reference" if use_ref else "port"
    sample = f"{per_step} synthetic {args.width}x{args.height} frames per step (extract + match to previous), all host threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.width, args.height, args.nf), "frames_per_step": per_step,
                   "outputs": "everything SPExtractor::operator() fills (heat, NMS, computeCovariance) + the match to the previous frame"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=752)
    ap.add_argument("--nf", type=int, default=800)
    ap.add_argument("--slots", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lean", action="store_true", help="skip computeCovariance and the heat_ image (keypoints, descriptors, "
                    "occ_grid, dust maps and matches only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from sp_orb_slam_b200 import SPExtractor, sharding
    rank, local_rank, world = sharding.init_distributed()
    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0)
    numa = sharding.bind_to_gpu_numa(local_rank)                 # before any pinned allocation (first touch)
    H, W, B, K, S = args.height, args.width, args.batch, args.steps, args.slots
    n_pool = max(4, -(-(140 << 20) // (B * H * W)))             # inputs > 126 MB L2
    pool = make_pool(H, W, B, n_pool, rank)
    # default workload = everything Frame::ExtractORB (frame.cpp:296-314) reads from the extractor: keypoints with
    # response, descriptors, occ_grid_, dense_dust_ / semi_dust_, heat_, cov2 / cov2_inv (computeCovariance), plus the
    # match against the previous frame.  heat_inv_ (= 1 - heat_, read by nothing outside computeCovariance) stays on
    # the device.  --lean drops computeCovariance and the heat_ image.
    full = not args.lean
    ex = SPExtractor(args.nf, H, W, WEIGHTS, device_id=local_rank, max_batch=B, num_slots=S,
                     emit_heat=full, emit_heat_inv=False, emit_cov=full, match_prev=True)
    d_pool = torch.from_numpy(pool).cuda()
    stride = B * H * W
    sampler = ClockSampler(local_rank)
    windows = []

    # ---------------- device-resident throughput (`value`)
    def run_device(e):
        for i in range(args.warmup):
            e.submit_device(0, d_pool.data_ptr() + (i % n_pool) * stride, B)
        e.sync(0)
        sharding.barrier()
        torch.cuda.synchronize()
        l0 = e.launch_count()
        w0 = time.time()
        e.timer_start(0)
        for i in range(K):
            e.submit_device(0, d_pool.data_ptr() + (i % n_pool) * stride, B)
        ms = e.timer_stop(0)
        torch.cuda.synchronize()
        windows.append((w0, time.time()))
        n_launch = e.launch_count() - l0
        sharding.barrier()
        frames_all, ms_all = sharding.aggregate_throughput(K * B, ms)
        return frames_all / (ms_all * 1e-3), ms_all, n_launch

    value, ms_all, launches = run_device(ex)
    lean_value = None
    if full:                                                     # the same stream without computeCovariance / heat_, for comparison
        lean = SPExtractor(args.nf, H, W, WEIGHTS, device_id=local_rank, max_batch=B, num_slots=1,
                           emit_heat=False, emit_heat_inv=False, emit_cov=False, match_prev=True)
        lean_value = run_device(lean)[0]
        lean.close()

    # ---------------- per-kernel device times (roofline), CUDA events between stages on the library's stream
    # (median of 8 passes after one warm-up pass: the board is power-capped and a single pass can land in a clock dip)
    stages, samples = {}, {}
    for rep in range(9):
        for s in ex.profile_device(0, d_pool.data_ptr() + (rep % n_pool) * stride, B):
            if rep:                                               # first repetition is warm-up
                stages.setdefault(s["name"], dict(ms=0.0, flop=s["flop"], bytes=s["bytes"]))
                samples.setdefault(s["name"], []).append(s["ms"])
    for n, d in stages.items():
        d["ms"] = float(np.median(samples[n]))
    peaks = measured_peaks()
    conv_names = [n for n in stages if n.startswith("conv")]
    dom = max(conv_names, key=lambda n: stages[n]["ms"])
    dom_tf = stages[dom]["flop"] / (stages[dom]["ms"] * 1e-3) / 1e12
    conv_ms = sum(stages[n]["ms"] for n in conv_names)
    conv_flop = FLOP_PER_PIXEL * H * W * B
    step_ms = sum(d["ms"] for d in stages.values())
    # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture (profiles/r01_traffic.json,
    # same frames-per-launch x 752x480 shape); None for other geometries
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if (H, W, B) == (480, 752, tj.get("frames_per_launch")):
            want = {"conv1a+1b": "conv1ab_mma_kernel"}.get(dom, dom)
            traffic = next((v for k, v in tj.items() if k.startswith(want)), None)
    roofline = {"bound": "tensor", "kernel": dom, "achieved": dom_tf, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                "frac": dom_tf / peaks["tflops_sustained"], "traffic": traffic, "peak_source": peaks["src"] + " bf16 sustained",
                "kernel_ms": stages[dom]["ms"], "kernel_share_of_step": stages[dom]["ms"] / step_ms,
                "conv_stack": {"achieved": conv_flop / (conv_ms * 1e-3) / 1e12, "frac": conv_flop / (conv_ms * 1e-3) / 1e12 / peaks["tflops_sustained"],
                               "frac_of_burst": conv_flop / (conv_ms * 1e-3) / 1e12 / peaks["tflops_burst"], "gflop_per_frame": conv_flop / B / 1e9},
                "whole_path_frac": (value / world) * FLOP_PER_PIXEL * H * W / 1e12 / peaks["tflops_sustained"],
                "stages_ms": {n: round(d["ms"], 4) for n, d in stages.items()}}

    # ---------------- end to end through the host-pointer ABI (H2D + kernels + D2H in the timed region)
    # headline: frames wait in page-locked host memory (spfe_submit_pinned: DMA straight from the caller's buffer);
    # secondary: pageable numpy frames through spfe_submit (adds the library's pageable -> pinned staging copy)
    pinned_pool = ex.pinned_frames(n_pool * B).reshape(n_pool, B, H, W)
    pinned_pool[:] = pool
    host_batches = [[pool[p, b] for b in range(B)] for p in range(n_pool)]

    def run_e2e(submit, n_steps):
        for i in range(max(S, 3)):
            submit(i % S, i % n_pool)
            ex.wait(i % S, B, unpack=False)
        sharding.barrier()
        torch.cuda.synchronize()
        w0 = time.time()
        t0 = time.perf_counter()
        for i in range(n_steps):
            s = i % S
            if i >= S:
                ex.wait(s, B, unpack=False)                      # results of the batch submitted S steps ago are on the host
            submit(s, i % n_pool)
        for i in range(n_steps, n_steps + S):
            ex.wait(i % S, B, unpack=False)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        windows.append((w0, time.time()))
        sharding.barrier()
        return sharding.aggregate_throughput(n_steps * B, ms)

    Ke = max(K, 2 * S)
    e2e_frames, e2e_ms_all = run_e2e(lambda s, p: ex.submit_pinned(s, pinned_pool[p]), Ke)
    pg_frames, pg_ms_all = run_e2e(lambda s, p: ex.submit(s, host_batches[p]), Ke)
    cap, cells = ex.cap, ex.hc * ex.wc
    h2d = B * H * W
    d2h = B * (4 + cap * (8 + 4 + 1024 + 8 + (20 if full else 0)) + cells * (2 + 4 + 4) + (H * W * 4 if full else 0)) + 8
    clocks = sampler.stop(windows)

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": ms_all / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic",
            "config": {"workload": workload_name(W, H, args.nf),
                       "frames_per_step": B, "slots": S, "weights": "superpoint_v1 (reference weights, tests/golden)",
                       "outputs": "everything Frame::ExtractORB reads: keypoints (+response), descriptors, occ_grid_, dust maps, "
                                  "heat_, cov2/cov2_inv (computeCovariance on the device), matches to the previous frame; "
                                  "only heat_inv_ (= 1 - heat_) stays on the device" if full else
                                  "keypoints, scores, descriptors, occ_grid, dust maps, matches (--lean: no computeCovariance, no heat_)",
                       "l2": f"inputs rotate over {n_pool} batches = {n_pool * stride >> 20} MiB > 126 MB L2; activations per step {B * H * W * 128 * 2 >> 20}+ MiB",
                       "parallelism": f"{world} independent streams, one per GPU, no data-path collective",
                       "host_placement": numa},
            "e2e": {"value": e2e_frames / (e2e_ms_all * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": Ke, "ms_per_step": e2e_ms_all / Ke, "input": "page-locked host frames (spfe_submit_pinned)",
                    "pageable_input_value": pg_frames / (pg_ms_all * 1e-3)},
            "gpu_launches": int(launches),
            "lean_value": lean_value,
            "clocks": clocks,
            "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            os.sched_setaffinity(0, all_cpus)                    # the CPU baseline gets every host core again
            out["cpu_baseline"] = cpu_baseline(H, W, args.nf)
        print(json.dumps(out), flush=True)
    ex.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the NL-Kalman per-frame step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1], "C2"): a 20-frame synthetic 1920x1080 RGB sequence,
sigma = 20, automatic parameters, two filtering iterations per frame (flt1 + flt2), no
smoothing.  One step = one frame of the recursion: rgb2opp, warp of the two previous
outputs, filter 1, filter 2 (reference src/main-flt.c:340-380).  Frames are taken in
sequence order and the recursion restarts (spatial first frame) every 20 frames, so 20
steps are exactly one sequence.  Metric: denoised Mpixel/s = w*h*frames / seconds / 1e6.

* value     -- inputs (noisy frame, flow, occlusion mask of every frame) resident in HBM,
               state resident in HBM, CUDA-event timed on the context's stream (nlk_seq_submit_dev
               per frame, nlk_seq_join before the closing event).
* e2e       -- the same steps through the host-buffer C-ABI call (nlk_seq_filter_host):
               every step copies its inputs from pinned host memory and both outputs back.
* roofline  -- the kernel with the largest share of the step, algorithmic flops per launch
                over its CUDA-event duration, against the fp32 FMA
               peak measured live on the same device.
* cpu_baseline / --impl reference -- the UNMODIFIED reference numerics (oracle/_ref,
               OpenMP on all host cores) on a bounded crop of the same workload.

N > 1 (torchrun, one rank per GPU): every rank runs its own independent sequence
(weak scaling, no data-path collective); value = total pixels / max-over-ranks time.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, CH, SIGMA, SEQ_LEN = 1920, 1080, 3, 20.0, 20
WORKLOAD = "C2: 20-frame synthetic 1920x1080 RGB sequence, sigma=20, flt1+flt2 per frame, no smoothing"
METRIC = "denoised Mpixel/s (1080p RGB sigma=20, flt1+flt2 step)"


# ---- algorithmic work ---------------------------------------------------

def pass_flops(w, h, ch, prm, temporal, bsic, smooth=False):
    """(search flops, group flops) of one pass with every grid patch counted (alpha = 1)"""
    psz, step = prm.patch_sz, prm.patch_sz // 2
    G = ((w - psz) // step + 1) * ((h - psz) // step + 1)
    r = prm.search_sz_t if (temporal or smooth) else prm.search_sz_x
    k = prm.npatches_t if temporal else prm.npatches_x
    k = min(k, (2 * r + 1) ** 2)
    nagg = min(k, prm.npatches_tagg)
    s = 2 if temporal else 1
    f_search = 3 * ch * psz * psz * (2 * r + 1) ** 2
    f_dct = 4 * psz ** 3
    f_xform = k * s * ch * f_dct + nagg * ch * f_dct + (nagg * ch * f_dct if bsic else 0)
    f_stat = (15 if temporal else 6) * k * ch * psz * psz
    f_gain = 10 * ch * psz * psz + 3 * nagg * ch * psz * psz
    return G * f_search, G * (f_xform + f_stat + f_gain)


# ---- clocks --------------------------------------------------------------------------------------

class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---- data ----------------------------------------------------------------------------------------

def make_sequence(n_frames, w=W, h=H, ch=CH, sigma=SIGMA, seed_offset=0):
    from bwd_nlkalman_b200 import synth
    frames = [synth.noisy_frame(w, h, ch, t, sigma, noise_seed=synth.NOISE_SEED + 100003 * seed_offset)
              for t in range(n_frames)]
    return frames, synth.backward_flow(w, h), synth.occlusion_mask(w, h)


# ---- reference arm (CPU, unmodified reference numerics) -----------------------------------------

def reference_sample(steps, warmup, budget_s=150.0, state=None):
    """Times the reference's own CPU path (oracle/_ref: src/nlkalman.c compiled unmodified,
    OpenMP on all host cores) on a bounded crop of the C2 workload.  One step = one temporal
    frame step (rgb2opp, 2x warp_bicubic, filter 1, filter 2, opp2rgb) on the crop."""
    from oracle import oracle as O
    from bwd_nlkalman_b200 import synth
    kind = "reference"
    if os.path.exists(O.REF_SO):
        cores = min(os.cpu_count() or 1, 100)  # dct_threads_init exits above 100 (src/nlkalman.c:164-170)
        impl = O.Ref(threads=cores)
    else:
        kind, cores, impl = "port", 1, O.Port()
    f1 = impl.default_params(SIGMA, O.FLT1)
    f2 = impl.default_params(SIGMA, O.FLT2)

    def one_step(cw, chh, prev1, prev2, t):
        n = synth.noisy_frame(cw, chh, CH, t, SIGMA)
        bflo, occ = synth.backward_flow(cw, chh), synth.occlusion_mask(cw, chh)
        t0 = time.perf_counter()
        o = impl.rgb2opp(n)
        w1 = impl.warp_bicubic(prev1, bflo, occ) if prev1 is not None else None
        a = impl.filter_frame(o, w1, None, SIGMA, f1)
        w2 = impl.warp_bicubic(prev2, bflo, occ) if prev2 is not None else None
        b = impl.filter_frame(o, w2, a, SIGMA, f2)
        out = impl.opp2rgb(b.copy())
        dt = time.perf_counter() - t0
        return a, b, out, dt

    crops = [(960, 540), (480, 270), (240, 136)]
    total = steps + max(warmup, 1)
    for ci, (cw, chh) in enumerate(crops):
        # frame 0 of the crop gives the state (spatial step, untimed), then one temporal probe
        p1, p2, _, _ = one_step(cw, chh, None, None, 0)
        p1, p2, _, dt = one_step(cw, chh, p1, p2, 1)
        if dt * total <= budget_s or ci == len(crops) - 1:
            break
    times = []
    t = 2
    for i in range(max(warmup - 1, 0) + steps):
        p1, p2, out, dt = one_step(cw, chh, p1, p2, t)
        t += 1
        if i >= max(warmup - 1, 0):
            times.append(dt)
    sec = sum(times)
    mpix = cw * chh * len(times) / sec / 1e6
    return {"value": mpix, "unit": "Mpixel/s", "cores": cores, "kind": kind,
            "sample": f"{len(times)} temporal flt1+flt2 frame steps on a {cw}x{chh} RGB crop of the C2 scene "
                      f"(state from the preceding frames of the same run), {cores} OpenMP threads, "
                      "FFTW replaced by the table-driven stand-in of oracle/fftw_shim",
            "ms_per_step": 1e3 * sec / len(times), "crop": [cw, chh]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cb = reference_sample(args.steps, args.warmup)
    line = {"metric": METRIC, "value": cb["value"], "unit": "Mpixel/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": cb["sample"]},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ---- our arm -------------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import bwd_nlkalman_b200 as nlk

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    K, Wm = args.steps, max(args.warmup, 3)
    frames, bflo, occ = make_sequence(SEQ_LEN, seed_offset=rank)
    f1 = nlk.default_params(SIGMA, nlk.FLT1)
    f2 = nlk.default_params(SIGMA, nlk.FLT2)
    ctx = nlk.Context(W, H, CH, device=local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    dev = torch.device("cuda", local_rank)

    # inputs resident in HBM: a distinct buffer per frame (1 GB in total, larger than L2)
    d_noisy = [torch.from_numpy(f).to(dev) for f in frames]
    d_flo = [torch.from_numpy(bflo).to(dev) for _ in frames]
    d_occ = [torch.from_numpy(occ).to(dev) for _ in frames]
    d_o1 = torch.empty((H, W, CH), dtype=torch.float32, device=dev)
    d_o2 = torch.empty((H, W, CH), dtype=torch.float32, device=dev)
    # pinned host copies for the end-to-end leg
    h_noisy = [torch.from_numpy(f).pin_memory() for f in frames]
    h_flo, h_occ = torch.from_numpy(bflo).pin_memory(), torch.from_numpy(occ).pin_memory()
    h_o1 = [torch.empty((H, W, CH), dtype=torch.float32).pin_memory() for _ in range(3)]
    h_o2 = [torch.empty((H, W, CH), dtype=torch.float32).pin_memory() for _ in range(3)]
    torch.cuda.synchronize()

    def step_dev(i):
        t = i % SEQ_LEN
        if t == 0:
            ctx.seq_reset()
        # pipelined recursion: the second filtering of a frame overlaps the first of the next one
        ctx.seq_submit_dev(d_noisy[t], d_flo[t] if t else None, d_occ[t] if t else None, SIGMA, f1, f2, d_o1, d_o2)

    def step_host(i, pipelined=True):
        # the streaming call: frame i's inputs go up and its two outputs come back inside the
        # timed region; copies overlap the neighbouring frames' kernels (three output sets)
        t = i % SEQ_LEN
        if t == 0:
            ctx.seq_reset()
        call = ctx.seq_submit_host if pipelined else ctx.seq_filter_host
        call(h_noisy[t], h_flo if t else None, h_occ if t else None, SIGMA, f1, f2, h_o1[i % 3], h_o2[i % 3])

    def barrier():
        if dist is not None:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident leg: `value` ---------------------------------------------------------
    for i in range(Wm):
        step_dev(i)
    ctx.sync()
    sampler = ClockSampler(local_rank)
    launches0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if rank == 0:
        sampler.start()
    with torch.cuda.stream(stream):
        e0.record()
        for i in range(Wm, Wm + K):
            step_dev(i)
        ctx.seq_join()          # the context's stream waits for the second filtering of the last frame
        e1.record()
    ctx.sync()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = int(sum_over_ranks(ctx.launches - launches0))
    value = world * W * H * K / (ms_total * 1e-3) / 1e6

    # ---- per-kernel durations: the same K steps once more, IN ORDER (nlk_seq_filter_dev: one
    # stream, no overlap between the two filterings), every kernel bracketed by CUDA events on
    # its stream.  In the timed leg above kernels of the two filterings run concurrently, so an
    # event pair around one of them would also measure the other.
    ctx.profile(True)
    ctx.profile_collect()
    for i in range(Wm, Wm + K):
        t = i % SEQ_LEN
        if t == 0:
            ctx.seq_reset()
        ctx.seq_filter_dev(d_noisy[t], d_flo[t] if t else None, d_occ[t] if t else None, SIGMA, f1, f2, d_o1, d_o2)
    prof = ctx.profile_collect()
    ctx.profile(False)

    # ---- end-to-end leg: host buffers through the C ABI ----------------------------------------
    def e2e_leg(pipelined):
        for i in range(Wm):
            step_host(i, pipelined)
        ctx.seq_drain()
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(stream):
            e0.record()
        for i in range(Wm, Wm + K):
            step_host(i, pipelined)
        with torch.cuda.stream(stream):
            e1.record()
        ctx.seq_drain()          # the last frame's outputs are in host memory
        wall = time.perf_counter() - t0
        barrier()
        return max_over_ranks(max(e0.elapsed_time(e1), wall * 1e3))
    e2e_sync_ms = e2e_leg(False)
    e2e_ms = e2e_leg(True)
    e2e_value = world * W * H * K / (e2e_ms * 1e-3) / 1e6
    e2e_sync_value = world * W * H * K / (e2e_sync_ms * 1e-3) / 1e6
    h2d = W * H * (CH + 3) * 4  # noisy + 2-channel flow + mask (frame 0 of a sequence: noisy only)
    d2h = 2 * W * H * CH * 4    # both filtering outputs, as the reference driver writes both

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel ------------------------------------------------------
    fp32_peak = ctx.fp32_peak(300.0)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json") if "hbm_gbs" in peaks else (6650.0, "fallback")
    fl = {"flt1_temporal": pass_flops(W, H, CH, f1, True, False), "flt1_spatial": pass_flops(W, H, CH, f1, False, False),
          "flt2_temporal": pass_flops(W, H, CH, f2, True, True), "flt2_spatial": pass_flops(W, H, CH, f2, False, True)}
    kernels, step_ms = [], sum(v[0] for v in prof.values()) / K
    for (kn, pk), (ms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        ent = {"kernel": kn, "pass": pk, "launches": cnt, "avg_ms": ms / cnt, "share_of_step": ms / K / step_ms}
        if kn in ("search_knn", "group_filter") and pk in fl:
            gf = fl[pk][0 if kn == "search_knn" else 1]
            ent["algorithmic_gflop"] = gf / 1e9
            ent["tflops"] = gf / (ms / cnt * 1e-3) / 1e12
            ent["frac_fp32_peak"] = ent["tflops"] / fp32_peak
        kernels.append(ent)
    dom = next(k for k in kernels if "tflops" in k)
    # compulsory HBM bytes of a pass: B = 4 w h (ch (n_in + 3) + 2)
    n_in = {"flt1_temporal": 2, "flt1_spatial": 1, "flt2_temporal": 3, "flt2_spatial": 2}[dom["pass"]]
    roofline = {"kernel": f'{dom["kernel"]} ({dom["pass"]})', "bound": "fp32",
                "achieved": dom["tflops"], "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": dom["tflops"] / fp32_peak,
                "peak_source": "fp32 FMA micro-benchmark run live on this device (nlk_fp32_peak); "
                               "the path is CUDA-core fp32 work, neither HBM- nor tensor-bound (SURVEY 8(d))",
                "traffic": None,
                "hbm": {"compulsory_bytes_per_pass": 4 * W * H * (CH * (n_in + 3) + 2), "peak_gbs": hbm_peak,
                        "peak_source": hbm_src}}
    prof_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(prof_path):
        try:
            with open(prof_path) as f:
                roofline["traffic"] = json.load(f).get(dom["kernel"] + ":" + dom["pass"])
        except (OSError, ValueError):
            pass

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu = reference_sample(steps=2, warmup=1, budget_s=30.0)
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as exc:  # the checker is optional for the number itself
            cpu = {"value": None, "unit": "Mpixel/s", "cores": 0, "kind": "unavailable", "sample": repr(exc)}

    line = {"metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frame": [W, H, CH], "sigma": SIGMA,
                       "params": {"flt1": f1.as_dict(), "flt2": f2.as_dict()},
                       "parallelism": f"{world} independent sequence(s), one per GPU" if world > 1 else "1 GPU",
                       "l2": "inputs larger than L2: 20 distinct frames (noisy + flow + mask = 1.0 GB) cycled",
                       "timed_call": "nlk_seq_submit_dev per frame (two-lane pipelined recursion), nlk_seq_join before the "
                                     "closing event",
                       "per_kernel_events": "separate leg over the same K steps, in order on one stream "
                                            "(nlk_seq_filter_dev), CUDA events around every kernel; the timed leg "
                                            "carries no per-kernel events"},
            "e2e": {"value": e2e_value, "unit": "Mpixel/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / K,
                    "api": "nlk_seq_submit_host per frame + nlk_seq_drain (pinned host buffers; uploads, kernels and "
                           "downloads of neighbouring frames overlap on three streams)",
                    "synchronous": {"value": e2e_sync_value, "ms_per_step": e2e_sync_ms / K,
                                    "api": "nlk_seq_filter_host (returns with both outputs in host memory)"}},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "kernels": kernels[:12],
            "in_order_ms_per_step": step_ms,
            "fp32_peak_tflops": fp32_peak}
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: relaunch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

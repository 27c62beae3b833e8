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
* e2e       -- the same steps through the host-buffer C-ABI call (nlk_seq_submit_host): every
               step uploads its inputs from pinned host memory and downloads the frame's result
               (the second filtering; the first stays on the device as recursion state -- the
               figure with both outputs downloaded is reported next to it).
* roofline  -- the kernel with the largest share of the step, algorithmic flops per launch
               (every grid patch counted) over its CUDA-event duration,
               against the fp32 FMA peak measured live on the same device; alpha = processed /
               grid patches and the fraction on the work actually done are reported with it.
* cpu_baseline / --impl reference -- the UNMODIFIED reference numerics (oracle/_ref, OpenMP on
               all host cores) on full 1920x1080 frames of the same sequence.  Its FFTW calls go
               to the table-driven stand-in of oracle/fftw_shim (no libfftw3f in this image):
               read the ratio as an upper bound by the (unmeasured) FFTW-codelet advantage.
* cli       -- wall time of one nlkalman-flt process on config C1 files, file I/O included:
               the reference program (all cores) against ours.

N > 1 (torchrun, one rank per GPU): the headline is N independent sequences (weak scaling, no
data-path collective); value = total pixels / max-over-ranks time.  The line also carries
"strips": one 3840x2160 sequence (config C4, filter + smoother) with every pass split into N
horizontal strips -- north_star's strong-scaling case -- next to the same sequence on one GPU
measured in the same run (tools/bench_strips.py).
"""
from __future__ import annotations

import argparse
import json
import os

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # lanes, copy streams and NCCL on their own hardware queues
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

def bench_config(world):
    """the `config` object, identical in both arms"""
    return {"workload": WORKLOAD, "frame_w": W, "frame_h": H, "channels": CH, "sigma": SIGMA,
            "parallelism": f"{world} independent sequence(s), one per GPU" if world > 1 else "1 GPU",
            "l2": "inputs larger than L2: 20 distinct frames (noisy + flow + mask = 1.0 GB) cycled"}


def reference_sample(steps, warmup, budget_s=240.0, state=None):
    """Times the reference's own CPU path (oracle/_ref: src/nlkalman.c compiled unmodified,
    OpenMP on all host cores) on full frames of the C2 workload.  One step = one temporal
    frame step (rgb2opp, 2x warp_bicubic, filter 1, filter 2, opp2rgb).  state = (flt1, flt2)
    RGB frames of frame 0 to start from; without it frame 0 is filtered here first (spatial,
    untimed).  Falls back to a crop only if the full frame does not fit the time budget."""
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

    sizes = [(W, H), (960, 540), (480, 270)]
    total = steps + max(warmup, 1)
    for ci, (cw, chh) in enumerate(sizes):
        if state is not None and (cw, chh) == (W, H):
            p1, p2 = impl.rgb2opp(state[0].copy()), impl.rgb2opp(state[1].copy())
        else:
            # frame 0 gives the state (spatial step, untimed)
            p1, p2, _, _ = one_step(cw, chh, None, None, 0)
        p1, p2, _, dt = one_step(cw, chh, p1, p2, 1)      # temporal probe (also the first warm-up step)
        if dt * total <= budget_s or ci == len(sizes) - 1:
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
    what = "full 1920x1080 RGB frames" if (cw, chh) == (W, H) else f"a {cw}x{chh} RGB crop (the full frame exceeds the time budget)"
    return {"value": mpix, "unit": "Mpixel/s", "cores": cores, "kind": kind,
            "sample": f"{len(times)} temporal flt1+flt2 frame steps on {what} of the C2 sequence, {cores} OpenMP threads; "
                      "FFTW replaced by the table-driven stand-in of oracle/fftw_shim (no libfftw3f here: an upper "
                      "bound on the ratio by the FFTW-codelet advantage)",
            "ms_per_step": 1e3 * sec / len(times), "frame": [cw, chh]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cb = reference_sample(args.steps, args.warmup)
    line = {"metric": METRIC, "value": cb["value"], "unit": "Mpixel/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(max(args.gpus, 1)),
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
    h_occ8 = torch.from_numpy(occ.astype(np.uint8)).pin_memory()   # the samples of the 8-bit mask file
    h_o1 = [torch.empty((H, W, CH), dtype=torch.float32).pin_memory() for _ in range(3)]
    h_o2 = [torch.empty((H, W, CH), dtype=torch.float32).pin_memory() for _ in range(3)]
    torch.cuda.synchronize()

    def step_dev(i):
        t = i % SEQ_LEN
        if t == 0:
            ctx.seq_reset()
        # pipelined recursion: the second filtering of a frame overlaps the first of the next one
        ctx.seq_submit_dev(d_noisy[t], d_flo[t] if t else None, d_occ[t] if t else None, SIGMA, f1, f2, d_o1, d_o2)

    def step_host(i, mode):
        # the streaming call: frame i's inputs go up and its result comes back inside the timed
        # region; copies overlap the neighbouring frames' kernels (three staging sets)
        #   "result": mask as the bytes of its 8-bit file, the frame's result (second filtering) back
        #   "full":   float mask, both filterings back (what the reference driver writes to disk)
        #   "sync":   "full" through the synchronous call
        t = i % SEQ_LEN
        if t == 0:
            ctx.seq_reset()
        if mode == "result":
            ctx.seq_submit_host(h_noisy[t], h_flo if t else None, h_occ8 if t else None, SIGMA, f1, f2, None, h_o2[i % 3])
        else:
            call = ctx.seq_submit_host if mode == "full" else ctx.seq_filter_host
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
    alpha = {pk: a / g for pk, (a, g) in ctx.profile_alpha().items()}
    ctx.profile(False)
    # frame 0's outputs: the state the CPU baseline starts from
    ctx.seq_reset()
    ctx.seq_filter_dev(d_noisy[0], None, None, SIGMA, f1, f2, d_o1, d_o2)
    ctx.sync()
    state0 = (d_o1.cpu().numpy().copy(), d_o2.cpu().numpy().copy())

    # ---- end-to-end leg: host buffers through the C ABI ----------------------------------------
    def e2e_leg(mode):
        ctx.seq_set_mask_mode(ctx.MASK_U8 if mode == "result" else ctx.MASK_FLOAT)
        for i in range(Wm):
            step_host(i, mode)
        ctx.seq_drain()
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(stream):
            e0.record()
        for i in range(Wm, Wm + K):
            step_host(i, mode)
        with torch.cuda.stream(stream):
            e1.record()
        ctx.seq_drain()          # the last frame's outputs are in host memory
        wall = time.perf_counter() - t0
        barrier()
        ctx.seq_set_mask_mode(ctx.MASK_FLOAT)
        return max_over_ranks(max(e0.elapsed_time(e1), wall * 1e3))
    e2e_sync_ms = e2e_leg("sync")
    e2e_full_ms = e2e_leg("full")
    e2e_ms = e2e_leg("result")
    mpix = lambda ms: world * W * H * K / (ms * 1e-3) / 1e6
    e2e_value, e2e_full_value, e2e_sync_value = mpix(e2e_ms), mpix(e2e_full_ms), mpix(e2e_sync_ms)
    h2d = W * H * (CH * 4 + 2 * 4 + 1)  # noisy + 2-channel flow + 8-bit mask (frame 0 of a sequence: noisy only)
    d2h = W * H * CH * 4                # the frame's result: the second filtering
    h2d_full, d2h_full = W * H * (CH + 3) * 4, 2 * W * H * CH * 4

    fp32_peak = ctx.fp32_peak(300.0) if rank == 0 else 0.0
    # ---- strips (N > 1): one 4K sequence split over the N GPUs -----------------------------------
    strips_res = None
    if world > 1 and not args.no_strips:
        ctx.close()
        del d_noisy, d_flo, d_occ
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_strips
        try:
            strips_res = bench_strips.measure(rank, world, local_rank, dist, nf=6, reps=2, warmup=1, transport="peer")
        except Exception as exc:     # the headline stands on its own
            strips_res = {"error": repr(exc)}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel ------------------------------------------------------
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
            if kn == "group_filter" and pk in alpha:
                # group_filter only runs the processed patches (search_knn runs every grid patch)
                ent["alpha"] = alpha[pk]
                ent["frac_fp32_peak_active"] = ent["frac_fp32_peak"] * alpha[pk]
        kernels.append(ent)
    dom = next(k for k in kernels if "tflops" in k)
    # compulsory HBM bytes of a pass: B = 4 w h (ch (n_in + 3) + 2)
    n_in = {"flt1_temporal": 2, "flt1_spatial": 1, "flt2_temporal": 3, "flt2_spatial": 2}[dom["pass"]]
    roofline = {"kernel": f'{dom["kernel"]} ({dom["pass"]})', "bound": "fp32",
                "achieved": dom["tflops"], "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": dom["tflops"] / fp32_peak,
                "alpha": dom.get("alpha"), "frac_active": dom.get("frac_fp32_peak_active"),
                "definition": "algorithmic flops of SURVEY 8(d) with every grid patch counted (alpha = 1) and the "
                              "matrix-form transform count F_dct = 4 psz^3; alpha = processed / grid patches of the "
                              "launch, frac_active = frac * alpha (the work the kernel actually ran)",
                "peak_source": "fp32 FMA micro-benchmark run live on this device (nlk_fp32_peak); "
                               "the path is CUDA-core fp32 work, neither HBM- nor tensor-bound (SURVEY 8(d))",
                "traffic": None, "traffic_source": None,
                "whole_step": {"gflop": sum(fl[k][0] + fl[k][1] for k in ("flt1_temporal", "flt2_temporal")) / 1e9,
                               "frac": sum(fl[k][0] + fl[k][1] for k in ("flt1_temporal", "flt2_temporal"))
                                       / (ms_total / K * 1e-3) / 1e12 / fp32_peak},
                "hbm": {"compulsory_bytes_per_pass": 4 * W * H * (CH * (n_in + 3) + 2), "peak_gbs": hbm_peak,
                        "peak_source": hbm_src}}
    prof_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(prof_path):
        try:
            with open(prof_path) as f:
                tj = json.load(f)
            roofline["traffic"] = tj.get(dom["kernel"] + ":" + dom["pass"])
            roofline["traffic_source"] = ("static: dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the "
                                          "committed ncu --set full capture " + str(tj.get("_source", "profiles/ncu_traffic.json"))
                                          + ", not measured in this run")
        except (OSError, ValueError):
            pass

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu = reference_sample(steps=2, warmup=1, budget_s=40.0, state=state0)
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as exc:  # the checker is optional for the number itself
            cpu = {"value": None, "unit": "Mpixel/s", "cores": 0, "kind": "unavailable", "sample": repr(exc)}
    cli = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cli = cli_leg()
        except Exception as exc:
            cli = {"error": repr(exc)}

    # the flow estimator in front of the filter: its own small leg, outside the timed region
    tvl1 = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_tvl1
            tvl1 = bench_tvl1.measure(W, H, reps=5, device=local_rank)
        except Exception as exc:
            tvl1 = {"error": repr(exc)}

    cfg = bench_config(world)
    line = {"metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg,
            "e2e": {"value": e2e_value, "unit": "Mpixel/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / K,
                    "api": "nlk_seq_submit_host per frame + nlk_seq_drain (pinned host buffers; uploads, kernels and "
                           "downloads of neighbouring frames overlap on three streams)",
                    "returns": "the frame's result (second filtering, RGB); the first filtering stays on the device as "
                               "recursion state; mask uploaded as the bytes of its 8-bit file (NLK_MASK_U8)",
                    "both_outputs_value": e2e_full_value,
                    "both_outputs": {"value": e2e_full_value, "ms_per_step": e2e_full_ms / K,
                                     "h2d_bytes_per_step": h2d_full, "d2h_bytes_per_step": d2h_full,
                                     "what": "float mask up, both filterings down (what the reference driver writes)"},
                    "synchronous": {"value": e2e_sync_value, "ms_per_step": e2e_sync_ms / K,
                                    "api": "nlk_seq_filter_host (returns with both outputs in host memory)"}},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "kernels": kernels[:12],
            "alpha": alpha, "in_order_ms_per_step": step_ms, "fp32_peak_tflops": fp32_peak,
            "details": {"params": {"flt1": f1.as_dict(), "flt2": f2.as_dict()},
                        "timed_call": "nlk_seq_submit_dev per frame (two-lane pipelined recursion), nlk_seq_join before "
                                      "the closing event",
                        "per_kernel_events": "separate leg over the same K steps, in order on one stream "
                                             "(nlk_seq_filter_dev), CUDA events around every kernel; the timed leg "
                                             "carries no per-kernel events"}}
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if cli is not None:
        line["cli"] = cli
    if tvl1 is not None:
        line["tvl1"] = tvl1
    if strips_res is not None:
        line["strips"] = strips_res
        if "value" in strips_res:
            # scalars in the objects the driver keeps
            line["config"]["strips_mpixel_s"] = strips_res["value"]
            line["config"]["strips_single_gpu_mpixel_s"] = strips_res.get("single_gpu_value")
            line["config"]["strips_workload"] = strips_res["workload"]
    print(json.dumps(line))
    if world == 1:
        ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


# ---- CLI wall time ----------------------------------------------------

def cli_leg():
    """One nlkalman-flt process on config C1 (854x480 gray, sigma 20, frame 1 with flow, mask and the
    previous filtered frames), file I/O and process start included: the unmodified reference program
    (oracle/_ref/nlkalman-flt-ref, all host cores) against bwd_nlkalman_b200/bin/nlkalman-flt."""
    import tempfile
    from bwd_nlkalman_b200 import synth
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "nlkalman-flt-ref")
    our_exe = os.path.join(ROOT, "bwd_nlkalman_b200", "bin", "nlkalman-flt")
    if not (os.path.exists(ref_exe) and os.path.exists(our_exe)):
        return {"error": "programs not built"}
    w, h, sigma = 854, 480, 20.0

    def pfm(path, a):
        with open(path, "wb") as f:
            f.write(f"Pf\n{w} {h}\n-1\n".encode() + np.ascontiguousarray(a, np.float32).tobytes())
    with tempfile.TemporaryDirectory() as d:
        for t in range(2):
            pfm(os.path.join(d, f"n{t}.pfm"), synth.noisy_frame(w, h, 1, t, sigma))
        with open(os.path.join(d, "bflo.flo"), "wb") as f:
            f.write(b"PIEH" + np.array([w, h], np.int32).tobytes() + synth.backward_flow(w, h).tobytes())
        with open(os.path.join(d, "occ.pgm"), "wb") as f:
            f.write(f"P5\n{w} {h}\n255\n".encode() + synth.occlusion_mask(w, h).astype(np.uint8).tobytes())
        p = lambda n: os.path.join(d, n)
        # previous filtered frames (frame 0) with our program, untimed
        subprocess.run([our_exe, "-i", p("n0.pfm"), "-s", str(sigma), "--flt11", p("a1.pfm"), "--flt21", p("a2.pfm")],
                       check=True, capture_output=True)
        args = ["-i", p("n1.pfm"), "-s", str(sigma), "-o", p("bflo.flo"), "-k", p("occ.pgm"), "--flt10", p("a1.pfm"),
                "--flt20", p("a2.pfm")]
        out = {}
        env = {k: v for k, v in os.environ.items() if k != "CUDA_DEVICE_MAX_CONNECTIONS"}   # (this process asks for 32 hardware queues: slower context creation)
        for name, exe in (("reference", ref_exe), ("ours", our_exe)):
            ts = []
            for rep in range(3):
                t0 = time.perf_counter()
                subprocess.run([exe] + args + ["--flt11", p(f"{name}1.pfm"), "--flt21", p(f"{name}2.pfm")],
                               check=True, capture_output=True, env=env)
                ts.append(time.perf_counter() - t0)
            out[name + "_s"] = min(ts)
        out["workload"] = ("C1: one nlkalman-flt process, 854x480 gray, sigma 20, frame 1 (flow, mask, previous flt1/flt2 "
                           "from PFM/FLO/PGM files), both filterings written; best of 3, process start, CUDA context "
                           "creation and file I/O included")
        out["cores"] = os.cpu_count()
        try:
            out["pipeline"] = cli_pipeline(d, w, h, sigma, pfm, env)
        except Exception as exc:
            out["pipeline"] = {"error": repr(exc)}
        return out


def cli_pipeline(d, w, h, sigma, pfm, env, nf=5):
    """The forward half of the pipeline script (scripts/nlkalman-seq.sh:31-102) on a short C1-sized sequence,
    flows and masks included: the reference's own programs chained through files the way the script chains
    them (tvl1flow-ref, plambda-ref, nlkalman-flt-ref twice per frame; PFM / FLO files, the formats this
    build of iio writes) against ONE nlkalman-seq --tvl1 1 process (everything resident on the GPU)."""
    from bwd_nlkalman_b200 import synth
    R = os.path.join(ROOT, "oracle", "_ref")
    tv, pl, fl = (os.path.join(R, n) for n in ("tvl1flow-ref", "plambda-ref", "nlkalman-flt-ref"))
    seq = os.path.join(ROOT, "bwd_nlkalman_b200", "bin", "nlkalman-seq")
    if not all(os.path.exists(x) for x in (tv, pl, fl, seq)):
        return {"error": "programs not built"}
    p = lambda n: os.path.join(d, n)
    for t in range(nf):
        pfm(p(f"s{t}.pfm"), synth.noisy_frame(w, h, 1, t, sigma))
    run = lambda *a: subprocess.run([str(x) for x in a], check=True, capture_output=True, env=env)
    expr = "x(0,0)[0] x(-1,0)[0] - x(0,0)[1] x(0,-1)[1] - + fabs 0.75 > 255 *"
    t0 = time.perf_counter()
    run(fl, "-i", p("s0.pfm"), "-s", sigma, "--flt11", p("r1_0.pfm"), "--flt21", p("r2_0.pfm"))
    for t in range(1, nf):
        run(tv, p(f"s{t}.pfm"), p(f"r2_{t-1}.pfm"), p(f"rb{t}.flo"), 8, 0, 0.25, 0, 0, 1)
        run(pl, p(f"rb{t}.flo"), expr, "-o", p(f"ro{t}.pfm"))
        run(fl, "-i", p(f"s{t}.pfm"), "-s", sigma, "--f2_p", 0, "-o", p(f"rb{t}.flo"), "-k", p(f"ro{t}.pfm"),
            "--flt10", p(f"r1_{t-1}.pfm"), "--flt11", p(f"r1_{t}.pfm"))
        run(fl, "-i", p(f"s{t}.pfm"), "-s", sigma, "--f1_p", 0, "-o", p(f"rb{t}.flo"), "-k", p(f"ro{t}.pfm"),
            "--flt11", p(f"r1_{t}.pfm"), "--flt20", p(f"r2_{t-1}.pfm"), "--flt21", p(f"r2_{t}.pfm"))
    ref_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    run(seq, "-i", p("s%d.pfm"), "-f", 0, "-l", nf - 1, "-s", sigma, "--first_f2", 1, "--tvl1", 1,
        "--filt1", p("o1_%d.pfm"), "--filt2", p("o2_%d.pfm"))
    ours_s = time.perf_counter() - t0

    def rd(name):
        raw = open(p(name), "rb").read().split(b"\n", 3)
        return np.frombuffer(raw[3], np.float32)
    diff = float(np.mean(np.abs(rd(f"o2_{nf-1}.pfm") - rd(f"r2_{nf-1}.pfm"))))
    return {"reference_s": ref_s, "ours_s": ours_s, "frames": nf,
            "mean_abs_diff_last_frame": diff,
            "workload": f"{nf} frames 854x480 gray, sigma 20: backward TV-L1 flow (FSCALE 1, DW 0.25), occlusion mask "
                        "(TH 0.75), first and second filtering per frame; reference = its programs chained through "
                        "files as scripts/nlkalman-seq.sh does (4 processes per frame), ours = one nlkalman-seq "
                        "--tvl1 1 process; wall time, process starts and file I/O included"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strips", action="store_true", help="N > 1: skip the strip-sharded 4K leg")
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

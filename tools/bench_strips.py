#!/usr/bin/env python
"""Strip-sharded filter + smoother on one 4K sequence (BASELINE.json config C4), one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29540 tools/bench_strips.py --frames 6 --reps 3 [--transport peer|nccl]

Workload: synthetic 3840x2160 RGB, sigma = 10, automatic parameters; per sequence the forward
recursion (flt1 + flt2 per frame) and then the backward RTS smoother, every pass split into
N horizontal strips (bwd_nlkalman_b200/strips.py).  Strong scaling: the frame is fixed, N
grows.  Metric: denoised Mpixel/s = w*h*frames / seconds (a frame counted once, filter and
smoother both done).  Inputs (noisy frames, both flows, masks) resident in HBM on every rank.
Timing: CUDA events on each rank's stream, barrier on both sides, max over ranks.

`measure()` is what bench.py calls for its N > 1 "strips" object; the 1-GPU denominator of the
strong scaling is measured in the same run on rank 0 alone (single context, the two-lane
pipelined recursion of the 1-GPU headline: the best one GPU can do, not the strip code at N = 1).
"""
import argparse
import json
import os
import sys
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # lanes, side streams and NCCL on their own hardware queues

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

WORKLOAD = ("C4: 3840x2160 RGB sequence sigma=10, flt1+flt2 forward then RTS smoother backward, every pass in "
            "horizontal strips (one per GPU) with bitmap / accumulator / halo exchange over NVLink")


def measure(rank, world, lr, dist, w=3840, h=2160, sigma=10.0, nf=4, reps=2, warmup=1, transport="peer",
            single_gpu_baseline=True, lanes=None):
    import torch
    import bwd_nlkalman_b200 as nlk
    from bwd_nlkalman_b200 import strips, synth
    dev = torch.device("cuda", lr)
    ch = 3
    f1, f2, s1 = (nlk.default_params(sigma, m) for m in (nlk.FLT1, nlk.FLT2, nlk.SMO1))
    up = lambda x: torch.from_numpy(x).to(dev)
    # two distinct noisy frames are enough to exercise the recursion; they alternate
    base = [synth.noisy_frame_cuda(w, h, ch, t, sigma, dev) for t in range(2)]
    frames = [base[t & 1] for t in range(nf)]
    bflo, fflo, occ = up(synth.backward_flow(w, h)), up(synth.forward_flow(w, h)), up(synth.occlusion_mask(w, h))
    flt_rgb = [torch.empty_like(base[0]) for _ in range(nf)]
    out = torch.empty_like(base[0])

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[lr])
        torch.cuda.synchronize()

    def timed(stream, ctx, sequence):
        for _ in range(max(warmup, 1)):
            sequence()
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launches
        t0 = time.perf_counter()
        with torch.cuda.stream(stream):
            e0.record()
        for _ in range(reps):
            sequence()
        with torch.cuda.stream(stream):
            e1.record()
        ctx.sync()
        return e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3, ctx.launches - l0

    res = {"workload": WORKLOAD, "frame": [w, h, ch], "sequence_frames": nf, "frames_timed": nf * reps,
           "n_gpus": world, "unit": "Mpixel/s", "scaling": "strong", "transport": transport}

    # ---- 1 GPU, whole frame (rank 0 alone; the others wait at the barrier) -----------------------
    if single_gpu_baseline:
        if rank == 0:
            ctx = nlk.Context(w, h, ch, device=lr)
            st = torch.cuda.ExternalStream(ctx.stream, device=dev)

            def seq1():
                ctx.seq_reset()
                for t in range(nf):
                    ctx.seq_submit_dev(frames[t], bflo if t else None, occ if t else None, sigma, f1, f2, None, flt_rgb[t])
                ctx.seq_join()
                ctx.seq_smooth_start_dev(flt_rgb[-1])
                for t in range(nf - 2, -1, -1):
                    ctx.seq_smooth_dev(flt_rgb[t], fflo, occ, sigma, s1, out)
            ms1, _, _ = timed(st, ctx, seq1)
            res["single_gpu_value"] = w * h * nf * reps / (ms1 * 1e-3) / 1e6
            res["single_gpu_ms_per_frame"] = ms1 / (nf * reps)
            ctx.close()
        barrier()
    if world == 1:
        return res

    # ---- N strips ---------------------------------------------------------------------------------
    if lanes is None:
        lanes = 2 if transport == "peer" else 1
    res["lanes"] = lanes
    rk = strips.StripRank(w, h, ch, rank, world, lr, transport=transport, lanes=lanes)
    if transport == "peer":
        strips.bind_dist(rk)

    def run(gen):
        strips.run_dist(rk, gen)

    def sequence():
        rk.reset()
        for t in range(nf):
            run(rk.filter_step(frames[t], bflo if t else None, occ if t else None, sigma, f1, f2, None, flt_rgb[t],
                               out2_for=s1))
        run(rk.last_filtered(flt_rgb[-1]))
        run(rk.smooth_start(flt_rgb[-1]))
        for t in range(nf - 2, -1, -1):
            run(rk.smooth_step(flt_rgb[t], fflo, occ, sigma, s1, out))

    barrier()
    ms, wall_ms, launches = timed(rk.stream, rk.ctx, sequence)
    barrier()
    t = torch.tensor([ms, wall_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, wall_ms = float(t[0].item()), float(t[1].item())
    # where a frame's time goes on this rank: one more sequence with CUDA events around every kernel of the
    # context's stream (a device-side wait shows as its own duration: time spent waiting for a peer)
    rk.ctx.profile(True)
    rk.ctx.profile_collect()
    barrier()
    sequence()
    prof = rk.ctx.profile_collect()
    rk.ctx.profile(False)
    phases = {}
    for (kn, pk), (pms, cnt) in prof.items():
        phases[kn] = phases.get(kn, 0.0) + pms / nf
    ph = torch.tensor([phases.get(k, 0.0) for k in rk.ctx.KERNELS], dtype=torch.float64, device=dev)
    ph_max, ph_min = ph.clone(), ph.clone()
    dist.all_reduce(ph_max, op=dist.ReduceOp.MAX)
    dist.all_reduce(ph_min, op=dist.ReduceOp.MIN)
    res["phases_ms_per_frame_max_over_ranks"] = {k: round(float(v), 4) for k, v in zip(rk.ctx.KERNELS, ph_max.tolist())}
    res["phases_ms_per_frame_min_over_ranks"] = {k: round(float(v), 4) for k, v in zip(rk.ctx.KERNELS, ph_min.tolist())}
    err = rk.ctx.peer_error() if transport == "peer" else 0
    e = torch.tensor([err], dtype=torch.int64, device=dev)
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    res.update({"value": w * h * nf * reps / (ms * 1e-3) / 1e6, "ms_per_frame": ms / (nf * reps),
                "host_wall_ms_per_frame": wall_ms / (nf * reps), "gpu_launches_rank0": int(launches),
                "strip": rk.plans(0, f1)[rank].as_dict(), "wait_timeouts": int(e.item())})
    rk.close()
    return res


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--w", type=int, default=3840)
    ap.add_argument("--h", type=int, default=2160)
    ap.add_argument("--sigma", type=float, default=10.0)
    ap.add_argument("--frames", type=int, default=6)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--no-single", action="store_true")
    ap.add_argument("--lanes", type=int, default=0, help="0: two for the peer transport, one for NCCL")
    a = ap.parse_args()
    rank, world, lr = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    res = measure(rank, world, lr, dist, a.w, a.h, a.sigma, a.frames, a.reps, a.warmup, a.transport, not a.no_single,
                  a.lanes or None)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Strip-sharded filter + smoother on one 4K sequence (BASELINE.json config C4), one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29540 tools/bench_strips.py --frames 6 --reps 3

Workload: synthetic 3840x2160 RGB, sigma = 10, automatic parameters; per sequence the forward
recursion (flt1 + flt2 per frame) and then the backward RTS smoother, every pass split into
N horizontal strips (bwd_nlkalman_b200/strips.py).  Strong scaling: the frame is fixed, N
grows.  Metric: denoised Mpixel/s = w*h*frames / seconds (a frame counted once, filter and
smoother both done).  Inputs (noisy frames, both flows, masks) resident in HBM on every rank.
Timing: CUDA events on each rank's stream, barrier on both sides, max over ranks.  Prints one
JSON line on rank 0.  N = 1 runs the same code without exchanges (the scaling baseline).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    import bwd_nlkalman_b200 as nlk
    from bwd_nlkalman_b200 import strips, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--w", type=int, default=3840)
    ap.add_argument("--h", type=int, default=2160)
    ap.add_argument("--sigma", type=float, default=10.0)
    ap.add_argument("--frames", type=int, default=6)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    a = ap.parse_args()
    rank, world, lr = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w, h, ch, sigma, nf = a.w, a.h, 3, a.sigma, a.frames
    f1, f2, s1 = (nlk.default_params(sigma, m) for m in (nlk.FLT1, nlk.FLT2, nlk.SMO1))
    up = lambda x: torch.from_numpy(x).to(dev)
    # two distinct noisy frames are enough to exercise the recursion; they alternate
    base = [up(synth.noisy_frame(w, h, ch, t, sigma)) for t in range(2)]
    frames = [base[t & 1] for t in range(nf)]
    bflo, fflo, occ = up(synth.backward_flow(w, h)), up(synth.forward_flow(w, h)), up(synth.occlusion_mask(w, h))
    rk = strips.StripRank(w, h, ch, rank, world, lr)
    flt_rgb = [torch.empty_like(base[0]) for _ in range(nf)]
    out = torch.empty_like(base[0])

    def run(gen):
        if world > 1:
            strips.run_dist(rk, gen)
        else:
            for req in gen:
                assert req[0] == "wait", "no exchange expected on one rank"

    def sequence():
        rk.reset()
        for t in range(nf):
            run(rk.filter_step(frames[t], bflo if t else None, occ if t else None, sigma, f1, f2, None, flt_rgb[t],
                               out2_for=s1))
        run(rk.last_filtered(flt_rgb[-1]))
        run(rk.smooth_start(flt_rgb[-1]))
        for t in range(nf - 2, -1, -1):
            run(rk.smooth_step(flt_rgb[t], fflo, occ, sigma, s1, out))

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[lr])
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 1)):
        sequence()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = rk.ctx.launches
    with torch.cuda.stream(rk.stream):
        e0.record()
    for _ in range(a.reps):
        sequence()
    with torch.cuda.stream(rk.stream):
        e1.record()
    rk.ctx.sync()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    if rank == 0:
        nfr = nf * a.reps
        print(json.dumps({
            "metric": "denoised Mpixel/s (4K RGB sigma=10, filter + smoother, strips)", "value": w * h * nfr / (ms * 1e-3) / 1e6,
            "unit": "Mpixel/s", "n_gpus": world, "frames": nfr, "ms_per_frame": ms / nfr, "higher_is_better": True,
            "scaling": "strong", "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C4: 3840x2160 RGB sequence sigma=10, flt1+flt2 forward then RTS smoother backward, "
                                   "horizontal strips with halo/bitmap/accumulator exchange over NCCL",
                       "frame": [w, h, ch], "sequence_frames": nf, "strip": rk.plans(0, f1)[rank].as_dict()},
            "gpu_launches": int(rk.ctx.launches - l0)}))
    rk.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

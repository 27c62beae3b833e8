#!/usr/bin/env python
"""The flow estimator in front of the filter: nlk_tvl1_flow_dev on one GPU beside the
reference's Dual_TVL1_optic_flow_multiscale (oracle/_ref/libtvl1_ref.so, all host threads) on the
same frame pair, with the parameters of the pipeline script (scripts/nlkalman-seq.sh:45-51:
FSCALE 1, DW 0.25, everything else default).

    python tools/bench_tvl1.py [--w 1920 --h 1080 --reps 5]

Prints one JSON line: Mpixel/s of both, the per-scale iteration counts, and max |du| between them.
Timing: wall clock around the call with the context synchronised (the level solver reads its
stopping error back every 20 iterations, so the host is part of the loop); inputs resident in HBM."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def measure(nx=1920, ny=1080, reps=5, lam=0.25, fscale=1, device=0, with_reference=True):
    import numpy as np
    import torch
    import bwd_nlkalman_b200 as nlk
    from oracle import oracle as O
    I0, I1 = O.tvl1_frames(nx, ny, seed=3)
    kw = dict(lam=lam, fscale=fscale)
    res = {"workload": f"TV-L1 flow {nx}x{ny}, lambda {lam}, fscale {fscale}, 5 warpings, epsilon 0.01 "
                       "(the tvl1flow call of scripts/nlkalman-seq.sh:60-65)", "unit": "Mpixel/s"}
    dev = torch.device("cuda", device)
    with nlk.Context(nx, ny, 1, device=device) as ctx:
        d0, d1 = torch.from_numpy(I0).to(dev), torch.from_numpy(I1).to(dev)
        u = torch.empty((2, ny, nx), device=dev)
        for _ in range(2):
            ctx.tvl1_flow_dev(d0, d1, u[0], u[1], nx, ny, **kw)
        ctx.sync()
        l0 = ctx.launches
        t0 = time.perf_counter()
        for _ in range(reps):
            ctx.tvl1_flow_dev(d0, d1, u[0], u[1], nx, ny, **kw)
        ctx.sync()
        dt = (time.perf_counter() - t0) / reps
        res.update({"value": nx * ny / dt / 1e6, "ms_per_pair": dt * 1e3,
                    "gpu_launches_per_pair": (ctx.launches - l0) // reps,
                    "launch_note": "per scale: gradient + per warping step the warp kernel and ONE cooperative launch that runs all "
                                   "its iterations (NLK_TVL1_LOOP=kernel, the default); plus the pyramid kernels"})
        flow, its = ctx.tvl1_flow(I0, I1, **kw)
        res["iterations_per_scale"] = its.sum(1).tolist()
    dx, dy = O.tvl1_truth(nx, ny)
    res["median_error_vs_scene_px"] = [float(np.median(np.abs(flow[0] - dx))), float(np.median(np.abs(flow[1] - dy)))]
    if with_reference and os.path.exists(O.TVL1_SO):
        ref = O.Tvl1Ref()
        t0 = time.perf_counter()
        want, _ = ref.flow(I0, I1, **kw)
        dt = time.perf_counter() - t0
        res["cpu_reference"] = {"value": nx * ny / dt / 1e6, "ms_per_pair": dt * 1e3, "cores": os.cpu_count(),
                                "kind": "reference"}
        res["max_abs_diff_px"] = float(np.abs(flow - want).max())
        res["identical_pixels"] = float(np.mean(flow == want))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--w", type=int, default=1920)
    ap.add_argument("--h", type=int, default=1080)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--lam", type=float, default=0.25)
    ap.add_argument("--fscale", type=int, default=1)
    a = ap.parse_args()
    print(json.dumps(measure(a.w, a.h, a.reps, a.lam, a.fscale)))


if __name__ == "__main__":
    main()

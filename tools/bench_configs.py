#!/usr/bin/env python
"""Device-resident throughput of the other BASELINE.json configurations on one GPU (the headline
bench.py measures C2).  One JSON line per configuration:

  C1  854x480 gray, sigma 20, flt1+flt2 temporal step
  C3  1920x1080 RGB, sigma 40, 12x12 patches, radii 10 (temporal) / 15 (spatial), filter then smoother
  C4  3840x2160 RGB, sigma 10, filter then smoother, one GPU (the strip bench's N = 1 point)
  C5  8 independent 960x540 RGB sequences, sigma 30, filter then smoother, one context + stream each
      (the per-GPU share of the 64-sequence throughput mode)

Metric: denoised Mpixel/s = w*h*frames / seconds, every frame filtered (flt1 + flt2) and, where the
configuration says so, smoothed.  CUDA events on the contexts' streams, inputs resident in HBM."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run_config(name, w, h, ch, sigma, over, smooth, nseq, nframes, reps):
    import torch
    import bwd_nlkalman_b200 as nlk
    from bwd_nlkalman_b200 import synth
    dev = torch.device("cuda", 0)
    mk = lambda mode, keys: nlk.default_params(sigma, mode, nlk.Params.auto(**{k: v for k, v in over.items() if k in keys}))
    f1 = mk(nlk.FLT1, ("patch_sz", "search_sz_x", "search_sz_t"))
    f2 = mk(nlk.FLT2, ("patch_sz", "search_sz_x", "search_sz_t"))
    s1 = mk(nlk.SMO1, ("patch_sz", "search_sz_t"))
    up = lambda a: torch.from_numpy(a).to(dev)
    base = [up(synth.noisy_frame(w, h, ch, t, sigma)) for t in range(min(nframes, 3))]
    frames = [base[t % len(base)] for t in range(nframes)]
    bflo, fflo, occ = up(synth.backward_flow(w, h)), up(synth.forward_flow(w, h)), up(synth.occlusion_mask(w, h))
    ctxs = [nlk.Context(w, h, ch, 0) for _ in range(nseq)]
    flt = [[torch.empty_like(base[0]) for _ in range(nframes)] for _ in range(nseq)]
    out = [torch.empty_like(base[0]) for _ in range(nseq)]

    def sequence():
        for c in ctxs:
            c.seq_reset()
        for t in range(nframes):          # round-robin over the sequences: their streams overlap
            for k, c in enumerate(ctxs):
                c.seq_submit_dev(frames[t], bflo if t else None, occ if t else None, sigma, f1, f2, None, flt[k][t])
        if smooth:
            for k, c in enumerate(ctxs):
                c.seq_smooth_start_dev(flt[k][-1])
            for t in range(nframes - 2, -1, -1):
                for k, c in enumerate(ctxs):
                    c.seq_smooth_dev(flt[k][t], fflo, occ, sigma, s1, out[k])

    sequence()
    torch.cuda.synchronize()
    l0 = sum(c.launches for c in ctxs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sequence()
    for c in ctxs:
        c.sync()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = sum(c.launches for c in ctxs) - l0
    # per-kernel durations: ONE sequence, IN ORDER on one stream (nlk_seq_filter_dev), CUDA events around
    # every kernel -- in the timed leg above the lanes / sequences overlap, and an event pair around a
    # kernel would also time whatever shares the GPU with it
    c = ctxs[0]
    c.profile(True)
    c.profile_collect()
    c.seq_reset()
    for t in range(nframes):
        c.seq_filter_dev(frames[t], bflo if t else None, occ if t else None, sigma, f1, f2, None, flt[0][t])
    if smooth:
        c.seq_smooth_start_dev(flt[0][-1])
        for t in range(nframes - 2, -1, -1):
            c.seq_smooth_dev(flt[0][t], fflo, occ, sigma, s1, out[0])
    prof = {k: [t, n] for k, (t, n) in c.profile_collect().items()}
    c.profile(False)
    kern = [{"kernel": k[0], "pass": k[1], "launches": n, "avg_ms": t / n} for k, (t, n) in
            sorted(prof.items(), key=lambda kv: -kv[1][0])[:10]]
    nfr = nframes * reps * nseq
    print(json.dumps({"config": name, "metric": "denoised Mpixel/s", "value": w * h * nfr / (ms * 1e-3) / 1e6,
                      "ms_per_frame": ms / nfr, "frame": [w, h, ch], "sigma": sigma, "sequences": nseq,
                      "frames_per_sequence": nframes, "smoother": bool(smooth),
                      "params": {"flt1": f1.as_dict(), "flt2": f2.as_dict(), "smo1": s1.as_dict()},
                      "gpu_launches": launches, "kernels_in_order": kern}), flush=True)
    for c in ctxs:
        c.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    cfgs = {
        "C1": ("C1", 854, 480, 1, 20.0, {}, 0, 1, 6, 3),
        "C3": ("C3", 1920, 1080, 3, 40.0, dict(patch_sz=12, search_sz_t=10, search_sz_x=15), 1, 1, 4, 2),
        "C4": ("C4", 3840, 2160, 3, 10.0, {}, 1, 1, 4, 2),
        "C5": ("C5", 960, 540, 3, 30.0, {}, 1, 8, 4, 2),
    }
    for k, v in cfgs.items():
        if not a.only or k in a.only.split(","):
            run_config(*v)


if __name__ == "__main__":
    main()

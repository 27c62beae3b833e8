#!/usr/bin/env python
"""Multi-GPU check of the strip-sharded recursion (run under torchrun, one rank per GPU):
filter + smoother on a small synthetic sequence through strips.run_dist (peer-memory or NCCL
transport) against the
single-context recursion computed on rank 0.  Prints "strips_check OK" on success."""
import argparse
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # lanes, side streams and NCCL on their own hardware queues
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    import bwd_nlkalman_b200 as nlk
    from bwd_nlkalman_b200 import strips, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--w", type=int, default=160)
    ap.add_argument("--h", type=int, default=200)
    ap.add_argument("--ch", type=int, default=3)
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--sigma", type=float, default=20.0)
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--lanes", type=int, default=1, help="2: the two filterings of a frame on two streams (peer transport)")
    a = ap.parse_args()
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    w, h, ch, sigma = a.w, a.h, a.ch, a.sigma
    f1, f2, s1 = (nlk.default_params(sigma, m) for m in (nlk.FLT1, nlk.FLT2, nlk.SMO1))
    up = lambda x: torch.from_numpy(x).to(dev)
    frames = [up(synth.noisy_frame(w, h, ch, t, sigma)) for t in range(a.frames)]
    bflo, fflo, occ = up(synth.backward_flow(w, h)), up(synth.forward_flow(w, h)), up(synth.occlusion_mask(w, h))
    rk = strips.StripRank(w, h, ch, rank, world, lr, transport=a.transport, lanes=a.lanes)
    if a.transport == "peer":
        strips.bind_dist(rk)        # CUDA IPC handles of the slabs over torch.distributed, once
    o1, o2 = torch.zeros_like(frames[0]), torch.zeros_like(frames[0])
    o1s = [torch.zeros_like(frames[0]) for _ in frames]
    o2s = [torch.zeros_like(frames[0]) for _ in frames]
    p1, p2, ps = rk.plans(0, f1), rk.plans(0, f2), rk.plans(1, s1)

    def gathered(t, plans):
        g = t.clone()
        with torch.cuda.stream(rk.stream):
            strips.allgather_rows(g, [(p.oy0, p.oy1) for p in plans])
        rk.ctx.sync()
        return g
    got1, got2, gots = [], [], [None] * a.frames
    # all frames queued back to back (no host synchronisation in between: with two lanes the second
    # filtering of a frame runs beside the first of the next), then compared
    for t in range(a.frames):
        strips.run_dist(rk, rk.filter_step(frames[t], bflo if t else None, occ if t else None, sigma, f1, f2, o1s[t], o2s[t]))
    rk.ctx.sync()
    for t in range(a.frames):
        got1.append(gathered(o1s[t], p1))
        got2.append(gathered(o2s[t], p2))
    strips.run_dist(rk, rk.smooth_start(got2[-1]))
    gots[-1] = got2[-1]
    for t in range(a.frames - 2, -1, -1):
        strips.run_dist(rk, rk.smooth_step(got2[t], fflo, occ, sigma, s1, o1))
        rk.ctx.sync()
        gots[t] = gathered(o1, ps)
    ok = True
    if rank == 0:
        with nlk.Context(w, h, ch, lr) as ctx:
            r1, r2 = torch.empty_like(frames[0]), torch.empty_like(frames[0])
            want2 = []
            for t in range(a.frames):
                ctx.seq_filter_dev(frames[t], bflo if t else None, occ if t else None, sigma, f1, f2, r1, r2)
                ctx.sync()
                e1, e2 = float((r1 - got1[t]).abs().max()), float((r2 - got2[t]).abs().max())
                print(f"frame {t}: flt1 max abs {e1:.2e}  flt2 max abs {e2:.2e}")
                ok &= e1 <= 1e-3 and e2 <= 1e-3
                want2.append(r2.clone())
            ctx.seq_smooth_start_dev(got2[-1])
            for t in range(a.frames - 2, -1, -1):
                ctx.seq_smooth_dev(got2[t], fflo, occ, sigma, s1, r1)
                ctx.sync()
                e = float((r1 - gots[t]).abs().max())
                print(f"frame {t}: smoother max abs {e:.2e}")
                ok &= e <= 1e-3
    if a.transport == "peer":
        err = rk.ctx.peer_error()
        if err:
            print(f"rank {rank}: a device-side wait timed out (code {err:#x})")
        errs = torch.tensor([err], device=dev, dtype=torch.int64)
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        ok &= int(errs.item()) == 0
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    rk.close()
    dist.destroy_process_group()
    if rank == 0 and flag.item():
        print("strips_check OK")
    return 0 if flag.item() else 1


if __name__ == "__main__":
    sys.exit(main())

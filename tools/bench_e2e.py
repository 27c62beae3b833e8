#!/usr/bin/env python
"""The device-resident and the end-to-end leg of bench.py alone (C2, N independent sequences, one per GPU):
how the host-buffer path scales with the number of GPUs sharing the host's memory and PCIe complex.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_e2e.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    import bench
    import bwd_nlkalman_b200 as nlk
    rank, world, lr = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W, H, CH, SIGMA, L = bench.W, bench.H, bench.CH, bench.SIGMA, bench.SEQ_LEN
    K, Wm = 20, 5
    frames, bflo, occ = bench.make_sequence(L, seed_offset=rank)
    f1, f2 = nlk.default_params(SIGMA, nlk.FLT1), nlk.default_params(SIGMA, nlk.FLT2)
    ctx = nlk.Context(W, H, CH, device=lr)
    stream = torch.cuda.ExternalStream(ctx.stream, device=lr)
    d_noisy = [torch.from_numpy(f).to(dev) for f in frames]
    d_flo, d_occ = torch.from_numpy(bflo).to(dev), torch.from_numpy(occ).to(dev)
    d_o2 = torch.empty((H, W, CH), dtype=torch.float32, device=dev)
    h_noisy = [torch.from_numpy(f).pin_memory() for f in frames]
    h_flo, h_occ = torch.from_numpy(bflo).pin_memory(), torch.from_numpy(occ).pin_memory()
    h_occ8 = torch.from_numpy(occ.astype(np.uint8)).pin_memory()
    h_o1 = [torch.empty((H, W, CH)).pin_memory() for _ in range(3)]
    h_o2 = [torch.empty((H, W, CH)).pin_memory() for _ in range(3)]

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[lr])
        torch.cuda.synchronize()

    def mx(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def leg(step, drain):
        for i in range(Wm):
            step(i)
        drain()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        with torch.cuda.stream(stream):
            e0.record()
        for i in range(Wm, Wm + K):
            step(i)
        with torch.cuda.stream(stream):
            e1.record()
        drain()
        wall = (time.perf_counter() - t0) * 1e3
        barrier()
        return world * W * H * K / (mx(max(e0.elapsed_time(e1), wall)) * 1e-3) / 1e6

    def step_dev(i):
        t = i % L
        if t == 0:
            ctx.seq_reset()
        ctx.seq_submit_dev(d_noisy[t], d_flo if t else None, d_occ if t else None, SIGMA, f1, f2, None, d_o2)

    def step_host(mode):
        def f(i):
            t = i % L
            if t == 0:
                ctx.seq_reset()
            if mode == "result":
                ctx.seq_submit_host(h_noisy[t], h_flo if t else None, h_occ8 if t else None, SIGMA, f1, f2, None, h_o2[i % 3])
            else:
                ctx.seq_submit_host(h_noisy[t], h_flo if t else None, h_occ if t else None, SIGMA, f1, f2, h_o1[i % 3], h_o2[i % 3])
        return f
    res = {"n_gpus": world, "unit": "Mpixel/s", "workload": bench.WORKLOAD}
    res["value"] = leg(step_dev, lambda: (ctx.seq_join(), ctx.sync()))
    ctx.seq_set_mask_mode(ctx.MASK_U8)
    res["e2e_result_only"] = leg(step_host("result"), ctx.seq_drain)
    ctx.seq_set_mask_mode(ctx.MASK_FLOAT)
    res["e2e_both_outputs"] = leg(step_host("full"), ctx.seq_drain)
    if rank == 0:
        print(json.dumps(res))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

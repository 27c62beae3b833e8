"""Two-process check of the peer-memory primitives (run under torchrun, one rank per GPU, tight timeout):
IPC mapping of the slabs, push + flag, push_add + flag, side-stream push, latency of a signal round trip."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    import torch, torch.distributed as dist
    import bwd_nlkalman_b200 as nlk
    from bwd_nlkalman_b200 import api
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    print(f"rank {rank}: can access peers:", [torch.cuda.can_device_access_peer(lr, p) for p in range(world) if p != lr], flush=True)
    ctx = nlk.Context(64, 64, 1, lr)
    hdr = api.lib().nlk_peer_header_bytes()
    nbytes = hdr + (1 << 20)
    slab = ctx.peer_slab_alloc(nbytes)
    mine = ctx.peer_ipc_export(slab)
    handles = [None] * world
    dist.all_gather_object(handles, mine)
    slabs = [slab if r == rank else ctx.peer_ipc_import(h) for r, h in enumerate(handles)]
    print(f"rank {rank}: slabs {[hex(s) for s in slabs]}", flush=True)
    ctx.peer_bind(rank, world, slabs, nbytes)

    class M:
        pass
    m = M()
    m.__cuda_array_interface__ = {"shape": ((1 << 20) // 4,), "typestr": "<f4", "data": (slab + hdr, False), "version": 2, "strides": None}
    buf = torch.as_tensor(m, device=dev)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    allm = ((1 << world) - 1) & ~(1 << rank)
    with torch.cuda.stream(stream):
        buf[:1024] = float(rank + 1)           # own region [0, 1024): pushed to the peers
        buf[1024 * (1 + rank):1024 * (2 + rank)] = 100.0 + rank
    ctx.sync()
    dist.barrier()
    # 1. push own 4 KB at offset rank-specific to every peer, flag slot 3
    off = hdr + 4096 * (1 + rank)
    ctx.peer_push(off, 4096, allm, 3, 1, 0)
    t0 = time.perf_counter()
    ctx.peer_wait(3, 1, allm)
    ctx.sync()
    print(f"rank {rank}: push+wait {1e3 * (time.perf_counter() - t0):.2f} ms, error {ctx.peer_error():#x}", flush=True)
    got = [float(buf[1024 * (1 + r)].item()) for r in range(world)]
    print(f"rank {rank}: regions {got} (want {[100.0 + r for r in range(world)]})", flush=True)
    # 2. push_add: every rank adds its [0,1024) into rank 0's... use ring: into (rank+1)%world
    dist.barrier()
    peer = (rank + 1) % world
    ctx.peer_push_add(hdr, 4096, peer, 4, 1)
    ctx.peer_wait(4, 1, 1 << ((rank - 1) % world))
    ctx.sync()
    print(f"rank {rank}: after push_add buf[0] = {float(buf[0].item())} (want {rank + 1 + ((rank - 1) % world) + 1}), error {ctx.peer_error():#x}", flush=True)
    # 3. side-stream push
    dist.barrier()
    ctx.peer_push(off, 4096, allm, 5, 1, 1)
    ctx.peer_wait(5, 1, allm)
    ctx.sync()
    print(f"rank {rank}: side push ok, error {ctx.peer_error():#x}", flush=True)
    # 4. signal round trips
    dist.barrier()
    n = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
    for i in range(n):
        ctx.peer_signal(6, i + 1, allm)
        ctx.peer_wait(6, i + 1, allm)
    with torch.cuda.stream(stream):
        e1.record()
    ctx.sync()
    print(f"rank {rank}: {n} signal/wait rounds: {1e3 * e0.elapsed_time(e1) / n:.1f} us each, error {ctx.peer_error():#x}", flush=True)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

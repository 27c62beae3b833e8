"""Host-side logic of the strip-sharded pass, on CPU: the row plan (nlk_strip_plan, pure host
arithmetic in the C library) and the exchange primitives of bwd_nlkalman_b200/strips.py over
torch.distributed with the gloo backend, world_size 2 and 3 (the N > 1 path of SURVEY 8(e))."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _plans(nlk, w, h, smooth, prms, n):
    return [nlk.strip_plan(w, h, smooth, prms, n, r) for r in range(n)]


@pytest.mark.parametrize("shape,n", [((3840, 2160), 8), ((3840, 2160), 4), ((1920, 1080), 2), ((160, 121), 3)])
@pytest.mark.parametrize("mode", ["FLT1", "FLT2", "SMO1"])
def test_strip_plan_partitions_the_frame(nlk, shape, n, mode):
    w, h = shape
    prms = nlk.default_params(10.0, getattr(nlk, mode))
    smooth = 1 if mode == "SMO1" else 0
    ps = _plans(nlk, w, h, smooth, prms, n)
    psz, step = prms.patch_sz, prms.patch_sz // 2
    r = prms.search_sz_t if smooth else max(prms.search_sz_t, prms.search_sz_x)
    gh = (h - psz) // step + 1
    assert ps[0].gy0 == 0 and ps[-1].gy1 == gh and ps[0].oy0 == 0 and ps[-1].oy1 == h
    for a, b in zip(ps, ps[1:]):
        assert a.gy1 == b.gy0 and a.oy1 == b.oy0          # no gap, no overlap
    for k, p in enumerate(ps):
        assert p.gh == gh and p.gw == (w - psz) // step + 1
        assert p.gy0 == k * p.chunk_g and p.gy1 - p.gy0 <= p.chunk_g and p.chunk_g == -(-gh // n)
        assert p.chunk_y == p.chunk_g * step and (k == n - 1 or p.oy1 - p.oy0 == p.chunk_y)
        assert p.oy0 == k * p.chunk_y                      # a rank's rows start at rank * chunk
        # halo = search radius above, search radius + patch below the last reference patch
        assert p.ey0 == max(p.gy0 * step - r, 0) and p.ey1 == min((p.gy1 - 1) * step + r + psz, h)
        # what spills over a border lands inside the immediate neighbour's own rows
        if k > 0:
            assert ps[k - 1].oy0 <= p.ey0
        if k + 1 < n:
            assert p.ey1 <= ps[k + 1].oy1
        assert p.ey0 <= p.oy0 and p.oy1 <= p.ey1


def test_strip_plan_rejects_thin_strips(nlk):
    prms = nlk.default_params(20.0, nlk.FLT1)
    with pytest.raises(nlk.NlkError):
        nlk.strip_plan(128, 64, 0, prms, 8, 0)     # 15 grid rows over 8 ranks: thinner than the halo
    one = nlk.strip_plan(128, 64, 0, prms, 1, 0)   # a single strip is always fine
    assert (one.gy0, one.gy1, one.oy0, one.oy1, one.ey0, one.ey1) == (0, 15, 0, 64, 0, 64)
    with pytest.raises(nlk.NlkError):
        nlk.strip_plan(64, 72, 0, prms, 4, 0)      # 17 grid rows in chunks of 5: the last strip (12 pixel rows) is thinner than the halo


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, w, h, ch, plan_rows, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bwd_nlkalman_b200 import strips

        class P:  # the fields the exchange uses
            def __init__(self, t):
                self.gy0, self.gy1, self.oy0, self.oy1, self.ey0, self.ey1, self.chunk_y = t
        plans = [P(t) for t in plan_rows]
        p = plans[rank]
        # (1) rows: every rank fills its own rows with a rank-specific pattern
        full = torch.arange(h * w * ch, dtype=torch.float32).reshape(h, w, ch)
        t = torch.full((h, w, ch), -1.0)
        t[p.oy0:p.oy1] = full[p.oy0:p.oy1] * (rank + 1)
        strips.allgather_rows(t, [(x.oy0, x.oy1) for x in plans])
        want = torch.empty_like(full)
        for r, x in enumerate(plans):
            want[x.oy0:x.oy1] = full[x.oy0:x.oy1] * (r + 1)
        ok_rows = bool(torch.equal(t, want))
        # (2) borders: partial accumulators, non-zero on the extended rows only
        g = torch.Generator().manual_seed(100 + rank)
        acc = torch.zeros(h, w, ch + 1)
        acc[p.ey0:p.ey1] = torch.rand((p.ey1 - p.ey0, w, ch + 1), generator=g)
        mine = acc.clone()
        strips.add_borders(acc, plans, rank)
        parts = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        total = sum(parts)
        ok_borders = bool(torch.allclose(acc[p.oy0:p.oy1], total[p.oy0:p.oy1], rtol=0, atol=1e-6))
        # (3) chunks: the in-place all-gather of an output frame laid out in equal chunks (+ tail)
        c, n = plan_rows[0][6], world
        hp = max(h, n * c)
        fullp = torch.full((hp, w, ch), -1.0)
        fullp[p.oy0:p.oy1] = full[p.oy0:p.oy1] * (rank + 1)
        works = strips.allgather_chunks(fullp, c, n, rank, (n * c, h) if h > n * c else None, async_op=True)
        for wk in works:
            wk.wait()
        ok_rows &= bool(torch.equal(fullp[:h], want))
        # (4) halo: own rows valid -> halo rows fetched from the neighbours
        img = torch.full((h, w, ch), -1.0)
        img[p.oy0:p.oy1] = want[p.oy0:p.oy1]
        strips.fill_halo(img, plans, rank)
        ok_rows &= bool(torch.equal(img[p.ey0:p.ey1], want[p.ey0:p.ey1]))
        q.put((rank, ok_rows, ok_borders))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_over_gloo(nlk, world):
    w, h, ch = 40, 150, 3
    prms = nlk.default_params(20.0, nlk.FLT1)
    plan_rows = [(p.gy0, p.gy1, p.oy0, p.oy1, p.ey0, p.ey1, p.chunk_y) for p in _plans(nlk, w, h, 0, prms, world)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, w, h, ch, plan_rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(r, True, True) for r in range(world)]


def test_border_ranges_are_mutually_consistent(nlk):
    from bwd_nlkalman_b200 import strips
    prms = nlk.default_params(10.0, nlk.SMO1)
    plans = _plans(nlk, 3840, 2160, 1, prms, 8)
    for r in range(8):
        br = strips.border_ranges(plans, r)
        if r > 0:
            assert br["up_send"] == strips.border_ranges(plans, r - 1)["dn_recv"]
            assert br["up_recv"] == strips.border_ranges(plans, r - 1)["dn_send"]
        else:
            assert br["up_send"] is None and br["up_recv"] is None
    assert strips.border_ranges(plans, 7)["dn_send"] is None
    step, psz, rr = prms.patch_sz // 2, prms.patch_sz, prms.search_sz_t
    a, b = strips.border_ranges(plans, 3)["dn_send"]
    assert b - a == rr + psz - step                      # rows below the last reference patch row
    a, b = strips.border_ranges(plans, 3)["up_send"]
    assert b - a == rr


# ---- the peer-transport schedule itself, without a GPU ------------------------------------------
# StripRank issues its exchanges as calls of the nlk_peer_* C ABI (asynchronous on a stream, waits on
# the device).  With a stand-in context that only records the calls, the schedules of N ranks can be
# replayed on the CPU as N x lanes in-order streams whose wait entries block until the flag they name
# has been stored by the source rank: the test checks that the replay never dead-locks, that every
# wait is matched by a push / signal of the same slot and sequence number addressed to the waiter,
# and that the accumulator and frame ranges pushed are the ones the row plan prescribes.

class _DryCtx:
    KERNELS = []

    def __init__(self, rank, log):
        self.rank, self.log, self.lane = rank, log, 0

    def _op(self, kind, **kw):
        self.log.append(dict(kind=kind, rank=self.rank, lane=self.lane, **kw))

    def peer_slab_alloc(self, nbytes):
        return 0x10000

    def peer_bind(self, rank, nranks, slabs, nbytes):
        pass

    def strip_lane(self, lane, reserve_sm=0):
        self.lane = lane

    def lane_record(self, idx):
        self._op("record", idx=idx)

    def lane_wait(self, idx):
        self._op("lwait", idx=idx)

    def peer_push(self, off, nbytes, mask, slot, value, side=0):
        self._op("push", off=off, nbytes=nbytes, mask=mask, slot=slot, value=value, side=side)

    def peer_push_add(self, off, nbytes, peer, slot, value):
        self._op("push_add", off=off, nbytes=nbytes, mask=1 << peer, slot=slot, value=value)

    def peer_signal(self, slot, value, mask):
        self._op("signal", mask=mask, slot=slot, value=value)

    def peer_wait(self, slot, value, mask):
        self._op("wait", mask=mask, slot=slot, value=value)

    def __getattr__(self, name):            # kernels and copies: recorded, nothing else
        def f(*a, **k):
            self._op(name)
        return f


def _replay(logs, nranks):
    """in-order streams (rank, lane) with device-side waits; returns the number of executed ops"""
    queues = {}
    for r, log in enumerate(logs):
        for op in log:
            queues.setdefault((r, op["lane"]), []).append(op)
    flags, events, done = {}, {}, 0
    heads = {k: 0 for k in queues}
    progress = True
    while progress:
        progress = False
        for k, q in queues.items():
            while heads[k] < len(q):
                op = q[heads[k]]
                if op["kind"] == "wait":
                    if not all(flags.get((op["rank"], op["slot"], s), 0) >= op["value"]
                               for s in range(nranks) if (op["mask"] >> s) & 1):
                        break
                elif op["kind"] == "lwait":
                    want = op.get("need")          # the record this wait saw at issue time (None: never recorded)
                    if want is not None and events.get((op["rank"], op["idx"]), 0) < want:
                        break
                elif op["kind"] == "record":
                    events[(op["rank"], op["idx"])] = op["serial"]
                elif op["kind"] in ("push", "push_add", "signal") and op["slot"] >= 0:
                    for d in range(nranks):
                        if (op["mask"] >> d) & 1 and d != op["rank"]:
                            flags[(d, op["slot"], op["rank"])] = max(flags.get((d, op["slot"], op["rank"]), 0), op["value"])
                heads[k] += 1
                done += 1
                progress = True
    stuck = {k: q[heads[k]] for k, q in queues.items() if heads[k] < len(q)}
    assert not stuck, f"dead-lock: {stuck}"
    return done


@pytest.mark.parametrize("nranks,lanes", [(2, 1), (3, 2), (8, 2), (8, 1)])
def test_peer_schedule_replays_without_deadlock(nlk, nranks, lanes):
    from bwd_nlkalman_b200 import strips
    w, h, ch, sigma, nframes = 3840, 2160, 3, 10.0, 4
    f1, f2, s1 = (nlk.default_params(sigma, m) for m in (nlk.FLT1, nlk.FLT2, nlk.SMO1))
    logs = [[] for _ in range(nranks)]
    ranks = [strips.StripRank(w, h, ch, r, nranks, transport="peer", lanes=lanes, ctx=_DryCtx(r, logs[r]))
             for r in range(nranks)]
    x = torch.empty(1)

    def drive(gens):
        gens = list(gens)
        alive = [True] * len(gens)
        while any(alive):
            for i, g in enumerate(gens):
                if alive[i]:
                    try:
                        assert next(g) == ("sync",)
                    except StopIteration:
                        alive[i] = False
    for t in range(nframes):
        drive(rk.filter_step(x, x if t else None, x if t else None, sigma, f1, f2, None, x, out2_for=s1) for rk in ranks)
    drive(rk.last_filtered(x) for rk in ranks)
    drive(rk.smooth_start(x) for rk in ranks)
    for t in range(nframes - 2, -1, -1):
        drive(rk.smooth_step(x, x, x, sigma, s1, x) for rk in ranks)
    # cross-lane events: a wait refers to the latest record of that event issued before it on the rank
    for log in logs:
        serial, last = 0, {}
        for op in log:
            if op["kind"] == "record":
                serial += 1
                op["serial"] = serial
                last[op["idx"]] = serial
            elif op["kind"] == "lwait":
                op["need"] = last.get(op["idx"])
    n = _replay(logs, nranks)
    assert n == sum(len(g) for g in logs)
    # every wait has its producers: same slot, same sequence number, addressed to the waiter
    for r, log in enumerate(logs):
        for op in (o for o in log if o["kind"] == "wait"):
            for s in range(nranks):
                if (op["mask"] >> s) & 1:
                    assert any(o["kind"] in ("push", "push_add", "signal") and o["slot"] == op["slot"] and
                               o["value"] == op["value"] and (o["mask"] >> r) & 1 for o in logs[s]), (r, s, op)
    # the overlap-add pushes are exactly the halo rows of the plan, into the two neighbours
    p1 = ranks[0].plans(0, f1)
    rowb = w * (ch + 1) * 4
    for r, log in enumerate(logs):
        adds = [o for o in log if o["kind"] == "push_add" and o["lane"] == 0]
        want = set()
        if r > 0 and p1[r].ey0 < p1[r].oy0:
            want.add((1 << (r - 1), (p1[r].oy0 - p1[r].ey0) * rowb))
        if r + 1 < nranks and p1[r].ey1 > p1[r].oy1:
            want.add((1 << (r + 1), (p1[r].ey1 - p1[r].oy1) * rowb))
        assert want <= {(o["mask"], o["nbytes"]) for o in adds}, (r, want)
    # with two lanes the second filterings are queued on lane 1
    if lanes == 2:
        assert any(o["lane"] == 1 and o["kind"] == "strip_search" for o in logs[0])

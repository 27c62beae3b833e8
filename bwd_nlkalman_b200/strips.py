"""Strip-sharded NL-Kalman step: one frame split into horizontal strips, one GPU per strip
(BASELINE.json config C4).

One process per GPU; the kernels are the same as the single-GPU pass, restricted to a range
of grid-patch rows through the ``nlk_strip_*`` C ABI (include/nlkalman_b200.h).  What crosses
strips, per pass (reference loop: src/nlkalman.c:590-595 over py, px):

  1. ``rows`` of the neighbour bitmaps -- the processed-pixel mask (src/nlkalman.c:597-600,
     :930-931) is a sequential chain over the whole frame, so every rank replays it on the
     bitmaps of all grid rows (4 bytes per patch; skipped when a group is a single patch);
  2. ``borders`` of the accumulator -- groups within r + psz rows of a strip border
     aggregate (src/nlkalman.c:913-928) into the neighbour's pixel rows: overlap-add;
  3. ``rows`` of the output -- the next pass searches r + psz rows beyond the strip and the
     next frame's warp (src/nlkalman.c:66-88) reads at flow-displaced positions.

The schedule of one rank is written as a generator that yields these exchange requests;
``run_dist`` serves them with torch.distributed (NCCL on GPUs, gloo in the CPU tests) and
``run_virtual`` serves N ranks living in one process (all strips on one GPU, lock-step),
which is how the strip logic is parity-tested on a single-GPU box.
"""
from __future__ import annotations

from . import api

# ---- exchange primitives (torch tensors: CUDA over NCCL, CPU over gloo) ------------------------


def allgather_rows(t, ranges, group=None):
    """Row range ``ranges[r]`` of ``t`` is valid on rank r; make all of them valid everywhere
    (in place, one broadcast per owner: strips may differ in height)."""
    import torch.distributed as dist
    for src, (a, b) in enumerate(ranges):
        if b > a:
            dist.broadcast(t[a:b], src=src, group=group)


def border_ranges(plans, rank):
    """Accumulator rows rank ``rank`` sends to / receives from its neighbours.
    -> dict(up_send, dn_send, up_recv, dn_recv), each a (row0, row1) pair or None."""
    p = plans[rank]
    out = dict(up_send=None, dn_send=None, up_recv=None, dn_recv=None)
    if rank > 0:
        q = plans[rank - 1]
        if p.ey0 < p.oy0:
            out["up_send"] = (p.ey0, p.oy0)          # my halo above, owned by rank-1
        if q.ey1 > q.oy1:
            out["up_recv"] = (q.oy1, q.ey1)          # rank-1's halo below = my first rows
    if rank + 1 < len(plans):
        q = plans[rank + 1]
        if p.ey1 > p.oy1:
            out["dn_send"] = (p.oy1, p.ey1)
        if q.ey0 < q.oy0:
            out["dn_recv"] = (q.ey0, q.oy0)
    return out


def add_borders(acc, plans, rank, group=None):
    """Overlap-add of the accumulator rows that groups of one strip wrote into the pixel rows
    owned by a neighbouring strip (in place on the owners' rows)."""
    import torch
    import torch.distributed as dist
    br = border_ranges(plans, rank)
    ops, recvs = [], []
    for key, peer in (("up_recv", rank - 1), ("dn_recv", rank + 1)):
        if br[key]:
            a, b = br[key]
            buf = torch.empty_like(acc[a:b])
            recvs.append((a, b, buf))
            ops.append(dist.P2POp(dist.irecv, buf, peer, group))
    for key, peer in (("up_send", rank - 1), ("dn_send", rank + 1)):
        if br[key]:
            a, b = br[key]
            ops.append(dist.P2POp(dist.isend, acc[a:b], peer, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for a, b, buf in recvs:
        acc[a:b] += buf


# ---- one rank's schedule -----------------------------------------------------------------------

class StripRank:
    """State and schedule of one strip: the resident recursion of nlk_seq_filter_dev /
    nlk_seq_smooth_dev (include/nlkalman_b200.h) with every pass strip-sharded.
    Full-frame buffers are kept on every rank (4K RGB: 100 MB each); only the rows a rank
    needs are computed or exchanged."""

    def __init__(self, w, h, ch, rank, nranks, device=0):
        import torch
        self.torch = torch
        self.w, self.h, self.ch, self.rank, self.nranks = w, h, ch, rank, nranks
        self.ctx = api.Context(w, h, ch, device)
        self.dev = torch.device("cuda", device)
        self.stream = torch.cuda.ExternalStream(self.ctx.stream, device=self.dev)
        f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=self.dev)
        self.noisy, self.warp, self.tmp = f(h, w, ch), f(h, w, ch), f(h, w, ch)
        self.flt1, self.flt2 = [f(h, w, ch), f(h, w, ch)], [f(h, w, ch), f(h, w, ch)]
        self.smo = [f(h, w, ch), f(h, w, ch)]
        self.accw = f(h, w, ch + 1)
        self.nbr = None
        self.reset()

    def reset(self):
        self.cur, self.have_prev, self.have_flt2, self.smo_cur, self.have_smo = 0, False, False, 0, False

    def close(self):
        self.ctx.close()

    def plans(self, smooth, prms):
        return [api.strip_plan(self.w, self.h, smooth, prms, self.nranks, r) for r in range(self.nranks)]

    # generator: yields ("rows", tensor, ranges) / ("borders", tensor, plans)
    def strip_pass(self, smooth, out, in1, prev0, bsic1, sigma, prms, gather_out=True):
        torch = self.torch
        plans = self.plans(smooth, prms)
        p = plans[self.rank]
        words = p.gh * p.gw * p.nbw
        if self.nbr is None or self.nbr.numel() < words:
            self.nbr = torch.empty(words, dtype=torch.int32, device=self.dev)
        self.ctx.strip_search(smooth, in1, prev0, bsic1, sigma, prms, p.gy0, p.gy1, self.nbr, self.accw)
        rmax = prms.search_sz_t if smooth else max(prms.search_sz_t, prms.search_sz_x)
        if self.nranks > 1 and prms.npatches_tagg > 1 and rmax // (prms.patch_sz // 2) >= 1:
            yield ("rows", self.nbr[:words].view(p.gh, p.gw * p.nbw), [(q.gy0, q.gy1) for q in plans])
        self.ctx.strip_filter()
        if self.nranks > 1:
            yield ("borders", self.accw, plans)
        self.ctx.strip_normalize(out, p.oy0, p.oy1)
        if self.nranks > 1 and gather_out:
            yield ("rows", out, [(q.oy0, q.oy1) for q in plans])
        return plans

    def _rows_needed(self, smooth, *prms_list):
        ps = [api.strip_plan(self.w, self.h, smooth, q, self.nranks, self.rank) for q in prms_list]
        return min(q.ey0 for q in ps), max(q.ey1 for q in ps)

    def filter_step(self, d_noisy, d_bflo, d_bocc, sigma, f1, f2, d_out1=None, d_out2=None, rows2=None):
        """One frame of the forward recursion (reference src/main-flt.c:340-380).  Inputs are
        full-frame device tensors of which rows [ey0, ey1) must be valid; the RGB outputs are
        written on the rows this rank owns (second output: on ``rows2`` if given, e.g. the
        rows a later smoothing pass of this rank reads)."""
        ctx = self.ctx
        cur, prv = self.cur, self.cur ^ 1
        do2 = f2.patch_sz != 0
        e0, e1 = self._rows_needed(0, *([f1, f2] if do2 else [f1]))
        ctx.colour_rows_dev(self.noisy, d_noisy, 0, e0, e1)
        prev1 = None
        if self.have_prev:
            prev1 = self.flt1[prv]
            if d_bflo is not None:
                a, b = self._rows_needed(0, f1)
                ctx.warp_rows_dev(self.warp, prev1, d_bflo, d_bocc, a, b)
                prev1 = self.warp
        plans = yield from self.strip_pass(0, self.flt1[cur], self.noisy, prev1, None, sigma, f1)
        p = plans[self.rank]
        if d_out1 is not None:
            ctx.colour_rows_dev(d_out1, self.flt1[cur], 1, p.oy0, p.oy1)
        if do2:
            prev2 = None
            if self.have_prev and self.have_flt2:
                prev2 = self.flt2[prv]
                if d_bflo is not None:
                    a, b = self._rows_needed(0, f2)
                    ctx.warp_rows_dev(self.warp, prev2, d_bflo, d_bocc, a, b)
                    prev2 = self.warp
            plans = yield from self.strip_pass(0, self.flt2[cur], self.noisy, prev2, self.flt1[cur], sigma, f2)
            p = plans[self.rank]
            if d_out2 is not None:
                a, b = rows2 if rows2 is not None else (p.oy0, p.oy1)
                ctx.colour_rows_dev(d_out2, self.flt2[cur], 1, a, b)
        self.have_prev, self.have_flt2, self.cur = True, do2, prv

    def smooth_start(self, d_last_rgb):
        """The last frame of a sequence is its own smoothed version (scripts/nlkalman-seq.sh:122-124).
        d_last_rgb: full frame, valid everywhere (e.g. the gathered filter output)."""
        self.ctx.colour_rows_dev(self.smo[0], d_last_rgb, 0, 0, self.h)
        self.smo_cur, self.have_smo = 0, True
        return
        yield  # noqa: makes this a generator like the other steps

    def smooth_step(self, d_flt_rgb, d_fflo, d_focc, sigma, s1, d_out=None):
        """One frame of the backward recursion (reference src/main-smo.c:198-213)."""
        ctx = self.ctx
        assert self.have_smo, "smooth_start must come first"
        nxt, cur = self.smo_cur, self.smo_cur ^ 1
        a, b = self._rows_needed(1, s1)
        ctx.colour_rows_dev(self.tmp, d_flt_rgb, 0, a, b)
        smo0 = self.smo[nxt]
        if d_fflo is not None:
            ctx.warp_rows_dev(self.warp, smo0, d_fflo, d_focc, a, b)
            smo0 = self.warp
        plans = yield from self.strip_pass(1, self.smo[cur], self.tmp, smo0, None, sigma, s1)
        p = plans[self.rank]
        if d_out is not None:
            ctx.colour_rows_dev(d_out, self.smo[cur], 1, p.oy0, p.oy1)
        self.smo_cur = cur


# ---- drivers -----------------------------------------------------------------------------------

def run_dist(rank_obj, gen, group=None):
    """Serve one rank's schedule with torch.distributed on the context's stream."""
    torch = rank_obj.torch
    with torch.cuda.stream(rank_obj.stream):
        for kind, t, arg in gen:
            if kind == "rows":
                allgather_rows(t, arg, group)
            else:
                add_borders(t, arg, rank_obj.rank, group)


def run_virtual(rank_objs, gens):
    """Serve N ranks that live in this process (lock-step): the exchanges become copies and
    adds between the ranks' buffers.  All schedules yield the same request sequence."""
    torch = rank_objs[0].torch
    gens = list(gens)
    while True:
        reqs = []
        for g in gens:
            try:
                reqs.append(next(g))
            except StopIteration:
                reqs.append(None)
        if all(r is None for r in reqs):
            return
        assert all(r is not None for r in reqs) and len({r[0] for r in reqs}) == 1, "schedules diverged"
        for o in rank_objs:
            o.ctx.sync()
        kind = reqs[0][0]
        if kind == "rows":
            ranges = reqs[0][2]
            for src, (a, b) in enumerate(ranges):
                for dst in range(len(gens)):
                    if dst != src and b > a:
                        reqs[dst][1][a:b].copy_(reqs[src][1][a:b])
        else:
            plans = reqs[0][2]
            stage = []
            for r in range(len(gens)):
                br = border_ranges(plans, r)
                for key, peer in (("up_send", r - 1), ("dn_send", r + 1)):
                    if br[key]:
                        a, b = br[key]
                        stage.append((peer, a, b, reqs[r][1][a:b].clone()))
            for peer, a, b, buf in stage:
                reqs[peer][1][a:b] += buf
        torch.cuda.synchronize()

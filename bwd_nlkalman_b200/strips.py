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

Two transports move them:

  * ``transport="peer"`` (the product path on one NVSwitch box): every rank keeps its exchange
    buffers in one slab at the same offsets, the peers' slabs are mapped with CUDA IPC, and the
    kernels of the ``nlk_peer_*`` C ABI store / reduce straight into the peer's memory over NVLink
    and flag completion there; consumers wait on the flags on the device.  No collective library
    call and no host-side wait in the data path: the bitmaps are pushed to every rank, the
    accumulator halo rows are ``red.global.add``-ed into their owner (no staging copy, no separate
    add), an output strip goes to the neighbours' halo rows at once and to everybody with the copy
    engines beside the next pass.
  * ``transport="nccl"``: the same schedule on torch.distributed collectives (NCCL on GPUs, gloo
    in the CPU tests) -- the baseline the peer transport is measured against.

The schedule of one rank is a generator.  With the NCCL transport it yields the exchange
requests, which ``run_dist`` serves with torch.distributed and ``run_virtual`` serves for N ranks
living in one process (all strips on one GPU, lock-step) -- how the strip logic is parity-tested
on a single-GPU box.  With the peer transport it issues the exchanges itself and yields only
("sync",) markers: points where every rank has queued what the others are about to wait for
(ignored by ``run_dist``; ``run_virtual`` drains all contexts there, so that no wait kernel ever
spins on a stream that shares a hardware queue with its producer).
"""
from __future__ import annotations

from . import api

# ---- exchange primitives (torch tensors: CUDA over NCCL, CPU over gloo) ------------------------


def allgather_rows(t, ranges, group=None):
    """Row range ``ranges[r]`` of ``t`` is valid on rank r; make all of them valid everywhere
    (in place, one broadcast per owner: the general form, for strips of any height)."""
    import torch.distributed as dist
    for src, (a, b) in enumerate(ranges):
        if b > a:
            dist.broadcast(t[a:b], src=src, group=group)


def allgather_chunks(full, chunk, nranks, rank, tail=None, group=None, async_op=False):
    """``full`` has at least nranks*chunk rows and rank k owns rows [k*chunk, (k+1)*chunk): ONE
    in-place all-gather (nlk_strip_plan lays the strips out this way).  ``tail`` = (a, b): rows
    beyond nranks*chunk that the last rank owns, broadcast separately.  Returns the work
    handles when async_op (wait on them before the gathered rows are read)."""
    import torch.distributed as dist
    works = []
    out = full[:nranks * chunk]
    w = dist.all_gather_into_tensor(out.view(-1), out[rank * chunk:(rank + 1) * chunk].reshape(-1), group=group,
                                    async_op=async_op)
    if async_op:
        works.append(w)
    if tail is not None and tail[1] > tail[0]:
        w = dist.broadcast(full[tail[0]:tail[1]], src=nranks - 1, group=group, async_op=async_op)
        if async_op:
            works.append(w)
    return works


def border_ranges(plans, rank):
    """Rows rank ``rank`` holds for / needs from its neighbours: its halo above and below
    (rows it reads and accumulates into but does not own).
    -> dict(up_send, dn_send, up_recv, dn_recv), each a (row0, row1) pair or None, in the
    accumulator sense: *_send = my halo rows, owned by the neighbour; *_recv = the neighbour's
    halo rows, owned by me."""
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


def _neighbour_exchange(t, br, rank, add, group=None):
    """add = True: overlap-add (send my halo rows, add what the neighbours accumulated into my
    rows).  add = False: halo fill (send my own border rows, receive my halo rows)."""
    import torch
    import torch.distributed as dist
    send = ("up_send", "dn_send") if add else ("up_recv", "dn_recv")
    recv = ("up_recv", "dn_recv") if add else ("up_send", "dn_send")
    ops, recvs = [], []
    for key, peer in zip(recv, (rank - 1, rank + 1)):
        if br[key]:
            a, b = br[key]
            buf = torch.empty_like(t[a:b]) if add else t[a:b]
            recvs.append((a, b, buf))
            ops.append(dist.P2POp(dist.irecv, buf, peer, group))
    for key, peer in zip(send, (rank - 1, rank + 1)):
        if br[key]:
            a, b = br[key]
            ops.append(dist.P2POp(dist.isend, t[a:b], peer, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    if add:
        for a, b, buf in recvs:
            t[a:b] += buf


def add_borders(acc, plans, rank, group=None):
    """Overlap-add of the accumulator rows that groups of one strip wrote into the pixel rows
    owned by a neighbouring strip (in place on the owners' rows)."""
    _neighbour_exchange(acc, border_ranges(plans, rank), rank, True, group)


def fill_halo(img, plans, rank, group=None):
    """Fetch the halo rows [ey0, oy0) and [oy1, ey1) of ``img`` from the neighbours that own
    them (what the next pass of this rank reads beyond its own rows)."""
    _neighbour_exchange(img, border_ranges(plans, rank), rank, False, group)


# ---- one rank's schedule -----------------------------------------------------------------------

class StripRank:
    """State and schedule of one strip: the resident recursion of nlk_seq_filter_dev /
    nlk_seq_smooth_dev (include/nlkalman_b200.h) with every pass strip-sharded.
    Full-frame buffers are kept on every rank (4K RGB: 100 MB each); only the rows a rank
    needs are computed or exchanged.

    The schedule methods are generators yielding exchange requests:
      ("nbr", words2d, plans)            all-gather the neighbour bitmaps (blocking)
      ("borders", accw, plans)           overlap-add of the accumulator halo rows (blocking)
      ("halo", frame, plans)             fetch the halo rows of a frame from the neighbours
      ("gather", frame, plans, key)      all-gather a frame's strips; asynchronous: it overlaps
                                         what is queued next, until ("wait", key)
      ("wait", key)
    """

    # flag slots of the peer transport (include/nlkalman_b200.h: nlk_peer_wait)
    SLOT_SEARCH, SLOT_ACC, SLOT_HALO, SLOT_FRAME = 0, 1, 8, 16      # SEARCH / ACC + 2 * lane; HALO + b, FRAME + b for frame buffer b

    def __init__(self, w, h, ch, rank, nranks, device=0, transport="nccl", lanes=1, ctx=None):
        import torch
        self.torch = torch
        self.w, self.h, self.ch, self.rank, self.nranks = w, h, ch, rank, nranks
        self.transport = transport
        # lanes = 2 (peer transport): the two filterings of a frame as two pipelines on two streams, the
        # second filtering of frame t beside the first of frame t+1 -- what nlk_seq_submit_dev does on
        # one GPU; a lane that waits for a peer's flag leaves the GPU to the other lane
        self.lanes, self.lane = lanes, 0
        assert lanes == 1 or (lanes == 2 and transport == "peer"), "two lanes need the peer transport"
        # ctx: a stand-in context that only records the calls (CPU tests of the schedule itself)
        self.dry = ctx is not None
        self.ctx = ctx if self.dry else api.Context(w, h, ch, device)
        self.dev = torch.device("cpu") if self.dry else torch.device("cuda", device)
        self.stream = None if self.dry else torch.cuda.ExternalStream(self.ctx.stream, device=self.dev)
        self._full = {}
        self._buf = {}                   # data_ptr of an exchanged frame buffer -> (index, slab offset)
        self.slab = 0
        if transport == "peer":
            self._carve_slab()
        elif transport != "nccl":
            raise ValueError(transport)
        self.noisy, self.warp, self.tmp = self.frame(local=True), self.frame(local=True), self.frame(local=True)
        self.noisy_l = [self.noisy, self.frame(local=True)] if lanes == 2 else [self.noisy, self.noisy]
        self.warp_l = [self.warp, self.frame(local=True)] if lanes == 2 else [self.warp, self.warp]
        self.flt1, self.flt2 = [self.frame(), self.frame()], [self.frame(), self.frame()]
        self.smo = [self.frame(), self.frame()]
        if transport == "peer":
            self.accw_l = [self._view(o, (h, w, ch + 1), "<f4") for o in self.off_accw_l]
            self.nbr_l = [self._view(o, (self.nbr_words,), "<i4") for o in self.off_nbr_l]
            self.accw, self.nbr = self.accw_l[0], self.nbr_l[0]
        else:
            self.accw = torch.empty((h, w, ch + 1), dtype=torch.float32, device=self.dev)
            self.nbr = None
        self.pending = {}
        self.seq = {}                    # flag slot -> last sequence number used (same on every rank)
        self.waited = {}
        self.local_rows = {}             # exchanged frame buffer -> rows of it this rank holds (peer transport)
        self._chunk_prms, self._last_rows = None, None
        self.reset()

    # ---- peer transport: the slab ---------------------------------------------------------------
    def _carve_slab(self):
        """header | neighbour bitmaps | accumulator | six frame buffers, the same offsets on every rank"""
        al = lambda x: (x + 255) // 256 * 256
        self.hp = self.h + 16 * self.nranks
        self.frame_bytes = al(self.hp * self.w * self.ch * 4)
        # bitmaps: one word per grid patch for any patch side >= 4 (step >= 2), whole chunks per rank
        self.nbr_words = (self.w // 2 + 1) * (self.h // 2 + 1 + self.nranks)
        off = al(api.lib().nlk_peer_header_bytes())
        self.off_nbr_l, self.off_accw_l = [], []
        for _ in range(self.lanes):              # bitmaps and accumulator per lane
            self.off_nbr_l.append(off)
            off = al(off + self.nbr_words * 4)
            self.off_accw_l.append(off)
            off = al(off + self.h * self.w * (self.ch + 1) * 4)
        self.off_nbr, self.off_accw = self.off_nbr_l[0], self.off_accw_l[0]
        self.off_frames = off
        self.n_slab_frames = 6
        self.slab_bytes = off + self.n_slab_frames * self.frame_bytes
        self.slab = self.ctx.peer_slab_alloc(self.slab_bytes)
        self._next_frame = 0

    def _view(self, off, shape, typestr):
        """torch tensor over a range of the slab (CUDA array interface, no copy)"""
        if self.dry:
            return self.torch.empty(tuple(shape), dtype=self.torch.float32 if typestr == "<f4" else self.torch.int32)

        class _Mem:
            pass
        m = _Mem()
        m.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (self.slab + off, False),
                                      "version": 2, "strides": None}
        t = self.torch.as_tensor(m, device=self.dev)
        assert t.data_ptr() == self.slab + off
        return t

    def bind_peers(self, slabs):
        """slabs[r]: rank r's slab as mapped in this process (own included)"""
        assert self.transport == "peer" and len(slabs) == self.nranks and slabs[self.rank] == self.slab
        self.ctx.peer_bind(self.rank, self.nranks, slabs, self.slab_bytes)

    def frame(self, local=False):
        """A frame buffer [h][w][ch] whose allocation is padded to whole chunks for any patch
        size, so that its strips can be all-gathered in place.  Exchanged buffers of the peer
        transport live in the slab."""
        torch = self.torch
        hp = self.h + 16 * self.nranks
        if self.transport == "peer" and not local:
            if self._next_frame >= self.n_slab_frames:
                raise RuntimeError("the slab holds six exchanged frame buffers")
            idx, off = self._next_frame, self.off_frames + self._next_frame * self.frame_bytes
            self._next_frame += 1
            full = self._view(off, (hp, self.w, self.ch), "<f4")
            t = full[:self.h]
            self._buf[t.data_ptr()] = (idx, off)
        else:
            full = torch.zeros((hp, self.w, self.ch), dtype=torch.float32, device=self.dev)
            t = full[:self.h]
        self._full[t.data_ptr()] = full
        return t

    def reset(self):
        self.cur, self.have_prev, self.have_flt2, self.smo_cur, self.have_smo = 0, False, False, 0, False
        self.frames_done = 0
        self._join_lanes()

    def _set_lane(self, lane):
        if self.lanes == 2:
            self.lane = lane
            self.ctx.strip_lane(lane, 1)       # group_filter leaves an SM to the other lane's mask_resolve

    def _join_lanes(self):
        """everything queued on lane 1 (second filterings) before what lane 0 queues next"""
        if self.lanes == 2:
            self._set_lane(0)
            self.ctx.lane_wait(2)
            self.ctx.lane_wait(3)

    def close(self):
        if self.slab:
            self.ctx.sync()
            self.accw = self.nbr = self.accw_l = self.nbr_l = None
            self.flt1 = self.flt2 = self.smo = None
            self._full.clear()
            self.ctx.dev_free(self.slab)
            self.slab = 0
        self.ctx.close()

    # ---- peer transport: masks, sequence numbers, waits -------------------------------------------
    def _mask_all(self):
        return ((1 << self.nranks) - 1) & ~(1 << self.rank)

    def _mask_nb(self):
        m = 0
        if self.rank > 0:
            m |= 1 << (self.rank - 1)
        if self.rank + 1 < self.nranks:
            m |= 1 << (self.rank + 1)
        return m

    def _next_seq(self, slot):
        self.seq[slot] = self.seq.get(slot, 0) + 1
        return self.seq[slot]

    def _await(self, slot, mask):
        """queue a device-side wait for the latest sequence number of `slot` from `mask` (once)"""
        v = self.seq.get(slot, 0)
        if v == 0 or mask == 0 or self.waited.get((self.lane, slot, mask), 0) >= v:
            return
        yield ("sync",)
        self.ctx.peer_wait(slot, v, mask)
        self.waited[(self.lane, slot, mask)] = v

    def need_frame(self, buf, key):
        """before reading rows of `buf` that OTHER strips produced, anywhere in the frame"""
        if self.nranks == 1:
            return
        if self.transport == "peer":
            yield from self._await(self.SLOT_FRAME + self._buf[buf.data_ptr()][0], self._mask_all())
        else:
            yield ("wait", key)

    def need_halo(self, buf, same_rows=True):
        """before reading the halo rows of `buf` (the neighbours' border rows); when the reading
        pass cuts the frame into other strips than the pass that wrote `buf`, the whole frame"""
        if self.nranks > 1 and self.transport == "peer":
            if not same_rows:
                raise api.NlkError("peer transport: the passes of a frame must cut it into the same strips (one patch size)")
            if same_rows:
                yield from self._await(self.SLOT_HALO + self._buf[buf.data_ptr()][0], self._mask_nb())
            else:
                yield from self._await(self.SLOT_FRAME + self._buf[buf.data_ptr()][0], self._mask_all())

    def plans(self, smooth, prms):
        return [api.strip_plan(self.w, self.h, smooth, prms, self.nranks, r) for r in range(self.nranks)]

    @staticmethod
    def _same_rows(pa, pb):
        return all((a.oy0, a.oy1) == (b.oy0, b.oy1) for a, b in zip(pa, pb))

    def strip_pass(self, smooth, out, in1, prev0, bsic1, sigma, prms, key=None, next_plans=None):
        """One strip-sharded pass into ``out`` (a buffer from frame()).  Afterwards the rank's
        own rows of ``out`` are final; the gather of the other strips is in flight under
        ``key`` (or finished, if key is None).  next_plans: plans of a pass of this same frame
        that reads ``out`` next -- its halo rows are fetched right away."""
        torch = self.torch
        plans = self.plans(smooth, prms)
        p = plans[self.rank]
        words = self.nranks * p.chunk_g * p.gw * p.nbw
        if self.transport == "peer":
            yield from self._strip_pass_peer(smooth, out, in1, prev0, bsic1, sigma, prms, plans, next_plans)
            return plans
        if self.nbr is None or self.nbr.numel() < words:
            self.nbr = torch.empty(words, dtype=torch.int32, device=self.dev)
        self.ctx.strip_search(smooth, in1, prev0, bsic1, sigma, prms, p.gy0, p.gy1, self.nbr, self.accw)
        rmax = prms.search_sz_t if smooth else max(prms.search_sz_t, prms.search_sz_x)
        if self.nranks > 1 and prms.npatches_tagg > 1 and rmax // (prms.patch_sz // 2) >= 1:
            yield ("nbr", self.nbr[:words].view(self.nranks * p.chunk_g, p.gw * p.nbw), plans)
        self.ctx.strip_filter()
        if self.nranks > 1:
            yield ("borders", self.accw, plans)
        self.ctx.strip_normalize(out, p.oy0, p.oy1)
        if self.nranks > 1:
            if next_plans is not None and self._same_rows(plans, next_plans) and key is not None:
                yield ("halo", out, next_plans)
                yield ("gather", out, plans, key)
            else:
                yield ("gather", out, plans, key)
                if next_plans is not None and key is not None:
                    yield ("wait", key)
        return plans

    def _strip_pass_peer(self, smooth, out, in1, prev0, bsic1, sigma, prms, plans, next_plans):
        """The pass with the exchanges done by the kernels themselves over peer memory."""
        ctx, rk, n = self.ctx, self.rank, self.nranks
        p = plans[rk]
        self._chunk_prms = prms
        words = n * p.chunk_g * p.gw * p.nbw
        if words > self.nbr_words:
            raise api.NlkError(f"peer transport: {words} bitmap words exceed the slab's {self.nbr_words}")
        ln = self.lane
        off_nbr, off_accw = self.off_nbr_l[ln], self.off_accw_l[ln]
        SLOT_SEARCH, SLOT_ACC = self.SLOT_SEARCH + 2 * ln, self.SLOT_ACC + 2 * ln
        ctx.strip_search(smooth, in1, prev0, bsic1, sigma, prms, p.gy0, p.gy1, self.nbr_l[ln], self.accw_l[ln])
        if n > 1:
            # (A) every rank has searched: its bitmap rows are here (when groups can mark other grid
            # cells), its accumulator rows are zeroed, and it is done reading the frame buffers that
            # the peers overwrite after this pass
            rmax = prms.search_sz_t if smooth else max(prms.search_sz_t, prms.search_sz_x)
            v = self._next_seq(SLOT_SEARCH)
            if prms.npatches_tagg > 1 and rmax // (prms.patch_sz // 2) >= 1:
                rowb = p.gw * p.nbw * 4
                ctx.peer_push(off_nbr + p.gy0 * rowb, (p.gy1 - p.gy0) * rowb, self._mask_all(), SLOT_SEARCH, v, 0)
            else:
                ctx.peer_signal(SLOT_SEARCH, v, self._mask_all())
            yield from self._await(SLOT_SEARCH, self._mask_all())
        ctx.strip_filter()
        if n > 1:
            # (B) overlap-add: the rows this strip's groups aggregated into beyond its border go
            # straight into the owner's accumulator
            v = self._next_seq(SLOT_ACC)
            br = border_ranges(plans, rk)
            rowb = self.w * (self.ch + 1) * 4
            for key, peer in (("up_send", rk - 1), ("dn_send", rk + 1)):
                if not 0 <= peer < n:
                    continue
                if br[key]:
                    a, b = br[key]
                    ctx.peer_push_add(off_accw + a * rowb, (b - a) * rowb, peer, SLOT_ACC, v)
                else:
                    ctx.peer_signal(SLOT_ACC, v, 1 << peer)
            yield from self._await(SLOT_ACC, self._mask_nb())
        ctx.strip_normalize(out, p.oy0, p.oy1)
        if n > 1:
            # (C) publish the strip: its border rows (the reader's halo plus a margin for the warp's
            # flow-displaced taps) go to the two neighbours at once; everybody else only learns that the
            # rows are final -- the next frame's warp pulls the few rows it may need beyond that straight
            # from the owner's slab (nlk_warp_rows_peer_dev), no strip is broadcast
            idx, off = self._buf[out.data_ptr()]
            v = self._next_seq(self.SLOT_HALO + idx)
            self.seq[self.SLOT_FRAME + idx] = v
            rowb = self.w * self.ch * 4
            rp = next_plans if next_plans is not None else plans
            lo, hi = self._local_rows(plans, rp, rk)
            self.local_rows[idx] = (lo, hi)
            for peer in (rk - 1, rk + 1):
                if not 0 <= peer < n:
                    continue
                plo, phi = self._local_rows(plans, rp, peer)        # what the neighbour keeps of the frame
                a, b = max(plo, p.oy0), min(phi, p.oy1)             # ... of which these rows are mine
                if b > a:
                    ctx.peer_push(off + a * rowb, (b - a) * rowb, 1 << peer, self.SLOT_HALO + idx, v, 0)
                else:
                    ctx.peer_signal(self.SLOT_HALO + idx, v, 1 << peer)
            ctx.peer_signal(self.SLOT_FRAME + idx, v, self._mask_all())

    WARP_MARGIN = 8      # rows beyond a pass's halo that a rank keeps of a frame (the warp's reach for small flows)

    def _local_rows(self, plans, reader_plans, r):
        """rows of a frame that rank r holds after the pass cut by `plans` produced it: its own, and the
        neighbours' rows within the halo of the pass that reads it next plus the margin"""
        lo = max(reader_plans[r].ey0 - self.WARP_MARGIN, plans[r - 1].oy0) if r > 0 else 0
        hi = min(reader_plans[r].ey1 + self.WARP_MARGIN, plans[r + 1].oy1) if r + 1 < len(plans) else self.h
        return min(lo, plans[r].oy0), max(hi, plans[r].oy1)

    def _warp(self, dst, src, flo, occ, a, b):
        """warp rows [a, b) of `src` (reference src/nlkalman.c:66-88); with the peer transport the rows this
        rank does not hold come from their owners"""
        if self.transport == "peer" and self.nranks > 1 and src.data_ptr() in self._buf:
            idx, off = self._buf[src.data_ptr()]
            lo, hi = self.local_rows.get(idx, (0, self.h))
            chunk_y = api.strip_plan(self.w, self.h, 0, self._chunk_prms, self.nranks, 0).chunk_y
            self.ctx.warp_rows_peer_dev(dst, off, flo, occ, a, b, lo, hi, chunk_y)
        else:
            self.ctx.warp_rows_dev(dst, src, flo, occ, a, b)

    def _rows_needed(self, smooth, *prms_list):
        ps = [api.strip_plan(self.w, self.h, smooth, q, self.nranks, self.rank) for q in prms_list]
        return min(q.ey0 for q in ps), max(q.ey1 for q in ps)

    def filter_step(self, d_noisy, d_bflo, d_bocc, sigma, f1, f2, d_out1=None, d_out2=None, out2_for=None):
        """One frame of the forward recursion (reference src/main-flt.c:340-380).  Inputs are
        full-frame device tensors of which rows [ey0, ey1) must be valid; the RGB outputs are
        written on the rows this rank owns.  out2_for = smoother parameters: the second output
        is also written on the halo rows that this rank's smoothing pass of the frame reads."""
        ctx = self.ctx
        cur, prv = self.cur, self.cur ^ 1
        do2 = f2.patch_sz != 0
        two = self.lanes == 2
        noisy = self.noisy_l[cur if two else 0]
        self._set_lane(0)
        if two:
            # lane 1 is done with this parity's noisy frame and first filtering (frame t-2)
            ctx.lane_wait(2 + cur)
        e0, e1 = self._rows_needed(0, *([f1, f2] if do2 else [f1]))
        ctx.colour_rows_dev(noisy, d_noisy, 0, e0, e1)
        prev1 = None
        if self.have_prev:
            prev1 = self.flt1[prv]
            yield from self.need_frame(prev1, "flt1")   # the other strips of the previous frame (warp reads anywhere)
            if d_bflo is not None:
                a, b = self._rows_needed(0, f1)
                self._warp(self.warp_l[0], prev1, d_bflo, d_bocc, a, b)
                prev1 = self.warp_l[0]
        plans = yield from self.strip_pass(0, self.flt1[cur], noisy, prev1, None, sigma, f1, key="flt1",
                                           next_plans=self.plans(0, f2) if do2 else None)
        p = plans[self.rank]
        if d_out1 is not None:
            ctx.colour_rows_dev(d_out1, self.flt1[cur], 1, p.oy0, p.oy1)
        if two:
            ctx.lane_record(cur)                # first filtering of this frame queued: lane 1 may follow
        if do2:
            if two:
                self._set_lane(1)
                ctx.lane_wait(cur)
            prev2 = None
            if self.have_prev and self.have_flt2:
                prev2 = self.flt2[prv]
                yield from self.need_frame(prev2, "flt2")
                if d_bflo is not None:
                    a, b = self._rows_needed(0, f2)
                    self._warp(self.warp_l[1], prev2, d_bflo, d_bocc, a, b)
                    prev2 = self.warp_l[1]
            nxt = self.plans(1, out2_for) if (out2_for is not None and d_out2 is not None) else None
            # the basic estimate on the rows beyond the strip
            yield from self.need_halo(self.flt1[cur], self._same_rows(plans, self.plans(0, f2)))
            plans = yield from self.strip_pass(0, self.flt2[cur], noisy, prev2, self.flt1[cur], sigma, f2,
                                               key="flt2", next_plans=nxt)
            p = plans[self.rank]
            if d_out2 is not None:
                a, b = (nxt[self.rank].ey0, nxt[self.rank].ey1) if nxt is not None else (p.oy0, p.oy1)
                if nxt is not None:
                    yield from self.need_halo(self.flt2[cur], self._same_rows(plans, nxt))
                ctx.colour_rows_dev(d_out2, self.flt2[cur], 1, min(a, p.oy0), max(b, p.oy1))
        if two:
            ctx.lane_record(2 + cur)            # this parity's buffers are free again once this has run
            self._set_lane(0)
        self.have_prev, self.have_flt2, self.cur = True, do2, prv

    def last_filtered(self, d_out_rgb, second=True):
        """RGB of the most recent filtered frame on this rank: the whole frame with the NCCL transport
        (waits for its gather), the rows this rank holds with the peer transport."""
        self._join_lanes()
        src = (self.flt2 if second else self.flt1)[self.cur ^ 1]
        if self.transport == "peer" and self.nranks > 1:
            yield from self.need_halo(src)
            lo, hi = self.local_rows.get(self._buf[src.data_ptr()][0], (0, self.h))
            self._last_rows = (lo, hi)
            self.ctx.colour_rows_dev(d_out_rgb, src, 1, lo, hi)
            return
        yield from self.need_frame(src, "flt2" if second else "flt1")
        self.ctx.colour_rows_dev(d_out_rgb, src, 1, 0, self.h)

    def smooth_start(self, d_last_rgb):
        """The last frame of a sequence is its own smoothed version (scripts/nlkalman-seq.sh:122-124).
        d_last_rgb: valid on the whole frame, or (peer transport, after last_filtered) on the rows this
        rank holds."""
        self._join_lanes()
        lo, hi = self._last_rows if (self.transport == "peer" and self._last_rows) else (0, self.h)
        self._last_rows = None
        self.ctx.colour_rows_dev(self.smo[0], d_last_rgb, 0, lo, hi)
        self.smo_cur, self.have_smo = 0, True
        self.pending.pop("smo", None)
        if self.transport == "peer" and self.nranks > 1:
            # the rows are final: the smoothing pass of the frame before warps from them (and pulls what it
            # lacks from the owners)
            idx, _ = self._buf[self.smo[0].data_ptr()]
            v = self._next_seq(self.SLOT_HALO + idx)
            self.seq[self.SLOT_FRAME + idx] = v
            self.local_rows[idx] = (lo, hi)
            self.ctx.peer_signal(self.SLOT_HALO + idx, v, self._mask_nb())
            self.ctx.peer_signal(self.SLOT_FRAME + idx, v, self._mask_all())
        return
        yield  # noqa: makes this a generator like the other steps

    def smooth_step(self, d_flt_rgb, d_fflo, d_focc, sigma, s1, d_out=None):
        """One frame of the backward recursion (reference src/main-smo.c:198-213).  d_flt_rgb:
        the filtered frame, valid on this rank's rows [ey0, ey1) of the smoothing plan."""
        ctx = self.ctx
        assert self.have_smo, "smooth_start must come first"
        self._join_lanes()
        nxt, cur = self.smo_cur, self.smo_cur ^ 1
        a, b = self._rows_needed(1, s1)
        ctx.colour_rows_dev(self.tmp, d_flt_rgb, 0, a, b)
        smo0 = self.smo[nxt]
        yield from self.need_frame(smo0, "smo")
        if d_fflo is not None:
            self._warp(self.warp, smo0, d_fflo, d_focc, a, b)
            smo0 = self.warp
        plans = yield from self.strip_pass(1, self.smo[cur], self.tmp, smo0, None, sigma, s1, key="smo")
        p = plans[self.rank]
        if d_out is not None:
            ctx.colour_rows_dev(d_out, self.smo[cur], 1, p.oy0, p.oy1)
        self.smo_cur = cur


# ---- drivers -----------------------------------------------------------------------------------

def bind_virtual(rank_objs):
    """peer transport, all ranks in this process: the peers' slabs are plain pointers"""
    slabs = [o.slab for o in rank_objs]
    for o in rank_objs:
        o.bind_peers(slabs)


def bind_dist(rank_obj, group=None):
    """peer transport, one process per GPU: exchange the CUDA IPC handles of the slabs (plumbing
    over torch.distributed) and map the peers' slabs into this process"""
    import torch.distributed as dist
    mine = rank_obj.ctx.peer_ipc_export(rank_obj.slab)
    handles = [None] * rank_obj.nranks
    dist.all_gather_object(handles, mine, group=group)
    slabs = [rank_obj.slab if r == rank_obj.rank else rank_obj.ctx.peer_ipc_import(hd) for r, hd in enumerate(handles)]
    rank_obj.bind_peers(slabs)
    dist.barrier(group=group)


def _tail(plans, h):
    n, c = len(plans), plans[0].chunk_y
    return (n * c, h) if h > n * c else None


def run_dist(rank_obj, gen, group=None):
    """Serve one rank's schedule with torch.distributed on the context's stream."""
    torch = rank_obj.torch
    rk, n = rank_obj.rank, rank_obj.nranks
    with torch.cuda.stream(rank_obj.stream):
        for req in gen:
            kind = req[0]
            if kind == "sync":
                continue              # peer transport: the waits are on the device
            if kind == "nbr":
                allgather_chunks(req[1], req[2][0].chunk_g, n, rk, None, group)
            elif kind == "borders":
                add_borders(req[1], req[2], rk, group)
            elif kind == "halo":
                fill_halo(req[1], req[2], rk, group)
            elif kind == "gather":
                full = rank_obj._full[req[1].data_ptr()]
                works = allgather_chunks(full, req[2][0].chunk_y, n, rk, _tail(req[2], rank_obj.h), group,
                                         async_op=req[3] is not None)
                if req[3] is not None:
                    rank_obj.pending[req[3]] = works
            elif kind == "wait":
                for w in rank_obj.pending.pop(req[1], []):
                    w.wait()          # the context's stream waits; the host does not block
            else:
                raise ValueError(kind)


def run_virtual(rank_objs, gens):
    """Serve N ranks that live in this process (lock-step): the exchanges become copies and
    adds between the ranks' buffers.  All schedules yield the same request sequence."""
    torch = rank_objs[0].torch
    gens = list(gens)
    n = len(gens)
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
        kind = reqs[0][0]
        if kind == "wait":
            continue
        for o in rank_objs:
            o.ctx.sync()
        if kind == "sync":
            continue                  # peer transport: everything the coming waits need has run
        plans = reqs[0][2]
        if kind in ("nbr", "gather"):
            for src, p in enumerate(plans):
                a, b = (p.gy0, p.gy1) if kind == "nbr" else (p.oy0, p.oy1)
                for dst in range(n):
                    if dst != src and b > a:
                        reqs[dst][1][a:b].copy_(reqs[src][1][a:b])
        elif kind == "halo":
            for r in range(n):
                br = border_ranges(plans, r)
                for key, peer in (("up_send", r - 1), ("dn_send", r + 1)):
                    if br[key]:
                        a, b = br[key]
                        reqs[r][1][a:b].copy_(reqs[peer][1][a:b])
        elif kind == "borders":
            stage = []
            for r in range(n):
                br = border_ranges(plans, r)
                for key, peer in (("up_send", r - 1), ("dn_send", r + 1)):
                    if br[key]:
                        a, b = br[key]
                        stage.append((peer, a, b, reqs[r][1][a:b].clone()))
            for peer, a, b, buf in stage:
                reqs[peer][1][a:b] += buf
        else:
            raise ValueError(kind)
        torch.cuda.synchronize()

"""Parity at BASELINE.json's frame sizes and over a recursion, against the UNMODIFIED reference
(oracle/_ref, one OpenMP thread: the reference's output depends on its thread count through the
processed-pixel mask).  These are the slow tests: the CPU side takes one to
two minutes each.

  * C2: one temporal flt1 + flt2 step at 1920x1080x3 through nlk_seq_submit_dev, the call
    bench.py times (two-lane pipelined recursion, 147-SM group_filter launch);
  * C3: 12x12 patches, radii 10 / 15, sigma 40: first and second filtering and the smoother
    (k = tagg = 105) on 480x270x3;
  * a 20-frame 160x120 RGB sequence, forward recursion and backward smoother, against the
    chain of scripts/nlkalman-seq.sh:56-149 run with the reference library;
  * C4: 3840x2160x3 in 8 virtual strips against the single-context recursion.

Tolerances (north_star): max abs error <= 1e-3 on the 0-255 scale where both sides see
identical inputs; |dPSNR| <= 0.01 dB everywhere.  Where a pass consumes the OTHER side's
previous output (the recursion), a k-NN near-tie can flip a group, so the chain tests bound
the PSNR and the fraction of pixels beyond 1e-3 and print the per-frame maxima.
"""
import numpy as np
import pytest

from common import TOL_DPSNR, TOL_MAXABS, compare_knn, compare_with_reference, maxabs, psnr_between

pytestmark = [pytest.mark.gpu, pytest.mark.slow]


def _same_params(nlk, O, p):
    return O.Params(*[getattr(p, f) for f, _ in nlk.Params._fields_])


def _frac_above(a, b, tol=TOL_MAXABS):
    return float((np.abs(a.astype(np.float64) - b.astype(np.float64)) > tol).mean())


def _pass_three_ways(nlk, port, ref, O, name, smooth, in1, prev0, bsic1, sigma, prms, clean=None):
    """One pass on identical inputs: GPU (stage dump) against the restatement -- k-NN lists,
    distances (bit for bit), processed sets identical, output within 1e-3 -- and against the
    reference at one thread, where only documented distance near-ties may differ."""
    h, w, ch = in1.shape
    rp = _same_params(nlk, O, prms)
    with nlk.Context(w, h, ch) as ctx:
        g, gd = ctx.pass_host_debug(smooth, in1, prev0, bsic1, sigma, prms)
    p, pd = port.run_pass(O.PASS_SMOOTH if smooth else O.PASS_FILTER, in1, prev0, bsic1, sigma, rp, dump=True)
    assert np.array_equal(gd["nk"], pd["nk"]) and np.array_equal(gd["np0"], pd["np0"]), name
    G, nbad, details = compare_knn(gd, pd)
    assert nbad == 0, f"{name}: {nbad}/{G} k-NN lists differ: {details}"
    assert np.array_equal(gd["knn_d"], pd["knn_d"]), f"{name}: distances are not bit-identical"
    assert np.array_equal(gd["active"], pd["active"]), f"{name}: processed-patch sets differ"
    e = maxabs(g, p)
    assert e <= TOL_MAXABS, (name, e)
    r = (ref.smooth_frame if smooth else ref.filter_frame)(in1, prev0, bsic1, sigma, rp)
    src = bsic1 if bsic1 is not None else in1
    msg = compare_with_reference(name, g, r, src, pd, prms, smooth, clean)
    print(f"  vs restatement: {G} k-NN lists identical, {int(pd['active'].sum())} processed, max-abs {e:.2e} | vs reference {msg}")
    return g, r


def test_config2_temporal_step_1080p_rgb(nlk, port, ref):
    import torch
    from bwd_nlkalman_b200 import synth
    from oracle import oracle as O
    w, h, ch, sigma = 1920, 1080, 3, 20.0
    f1, f2 = nlk.default_params(sigma, nlk.FLT1), nlk.default_params(sigma, nlk.FLT2)
    frames = [synth.noisy_frame(w, h, ch, t, sigma) for t in range(2)]
    clean1 = synth.clean_frame(w, h, ch, 1)
    bflo, occ = synth.backward_flow(w, h), synth.occlusion_mask(w, h)
    dev = torch.device("cuda", 0)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_fr, d_flo, d_occ = [up(f) for f in frames], up(bflo), up(occ)
    o1 = [torch.empty_like(d_fr[0]) for _ in range(2)]
    o2 = [torch.empty_like(d_fr[0]) for _ in range(2)]
    with nlk.Context(w, h, ch) as ctx:
        # the benchmarked call: frame 0 (spatial), frame 1 (temporal), second filterings on lane 1
        ctx.seq_submit_dev(d_fr[0], None, None, sigma, f1, f2, o1[0], o2[0])
        ctx.seq_submit_dev(d_fr[1], d_flo, d_occ, sigma, f1, f2, o1[1], o2[1])
        ctx.seq_drain()
        g = [[t.cpu().numpy() for t in o1], [t.cpu().numpy() for t in o2]]
    # frame 1 on the CPU from the GPU's frame-0 state (what scripts/nlkalman-seq.sh passes between
    # processes: RGB frames, transformed again on load, src/main-flt.c:340-342)
    n1 = ref.rgb2opp(frames[1].copy())
    cl1 = ref.rgb2opp(clean1.copy())
    p1, p2 = ref.rgb2opp(g[0][0].copy()), ref.rgb2opp(g[1][0].copy())
    w1, w2 = ref.warp_bicubic(p1, bflo, occ), ref.warp_bicubic(p2, bflo, occ)
    print("C2 1920x1080x3, frame 1:")
    g11, r11 = _pass_three_ways(nlk, port, ref, O, "flt1 temporal", 0, n1, w1, None, sigma, f1, cl1)
    g21, r21 = _pass_three_ways(nlk, port, ref, O, "flt2 temporal", 0, n1, w2, r11, sigma, f2, cl1)
    # the pipelined recursion gave the same first filtering (its state never left the GPU: no RGB
    # round trip, 3e-5) ...
    e1 = maxabs(g[0][1], ref.opp2rgb(g11.copy()))
    # ... and, statistically, the same second filtering: it searches on ITS first filtering, which
    # differs from r11 in the last bits, so a near-tie may flip a group
    rgb21 = ref.opp2rgb(r21.copy())
    frac2 = _frac_above(g[1][1], rgb21)
    dp2 = abs(psnr_between(g[1][1], clean1) - psnr_between(rgb21, clean1))
    print(f"  nlk_seq_submit_dev chain: flt1 max-abs {e1:.2e} vs the single pass; flt2 {frac2:.1e} of pixels > 1e-3 "
          f"vs the reference chain, dPSNR {dp2:.1e}")
    assert e1 <= TOL_MAXABS
    assert frac2 <= 2e-3 and dp2 <= TOL_DPSNR
    assert psnr_between(g[1][1], clean1) > psnr_between(frames[1], clean1) + 8


def test_config3_patch12_filter_and_smoother(nlk, port, ref):
    from bwd_nlkalman_b200 import synth
    from oracle import oracle as O
    w, h, ch, sigma = 480, 270, 3, 40.0
    ov = dict(patch_sz=12, search_sz_t=10, search_sz_x=15)
    f1 = nlk.default_params(sigma, nlk.FLT1, nlk.Params.auto(**ov))
    f2 = nlk.default_params(sigma, nlk.FLT2, nlk.Params.auto(**ov))
    s1 = nlk.default_params(sigma, nlk.SMO1, nlk.Params.auto(patch_sz=12, search_sz_t=10))
    assert (f1.npatches_x, f2.npatches_t, s1.npatches_t, s1.npatches_tagg) == (60, 40, 105, 105)
    n0 = ref.rgb2opp(synth.noisy_frame(w, h, ch, 0, sigma))
    n1 = ref.rgb2opp(synth.noisy_frame(w, h, ch, 1, sigma))
    cl0, cl1 = ref.rgb2opp(synth.clean_frame(w, h, ch, 0)), ref.rgb2opp(synth.clean_frame(w, h, ch, 1))
    bflo, fflo, occ = synth.backward_flow(w, h), synth.forward_flow(w, h), synth.occlusion_mask(w, h)
    print("C3 480x270x3, 12x12 patches, radii 10 / 15, sigma 40:")
    # frame 0, spatial (radius 15: 961 candidates, 60 kept), then second filtering
    _, r10 = _pass_three_ways(nlk, port, ref, O, "flt1 spatial", 0, n0, None, None, sigma, f1, cl0)
    _, r20 = _pass_three_ways(nlk, port, ref, O, "flt2 spatial", 0, n0, None, r10, sigma, f2, cl0)
    # frame 1, temporal (radius 10)
    w1, w2 = ref.warp_bicubic(r10, bflo, occ), ref.warp_bicubic(r20, bflo, occ)
    _, r11 = _pass_three_ways(nlk, port, ref, O, "flt1 temporal", 0, n1, w1, None, sigma, f1, cl1)
    _, r21 = _pass_three_ways(nlk, port, ref, O, "flt2 temporal", 0, n1, w2, r11, sigma, f2, cl1)
    # smoother of frame 0 from frame 1 (k = tagg = 105)
    ws = ref.warp_bicubic(r21, fflo, occ)
    _pass_three_ways(nlk, port, ref, O, "smoother", 1, r20, ws, None, sigma, s1, cl0)


def test_sequence_20_frames_against_reference_chain(nlk, ref):
    """forward recursion (flt1 + flt2 per frame) and backward smoother of a 20-frame 160x120 RGB
    sequence through the resident-state API, against the reference library driven like
    scripts/nlkalman-seq.sh (state passed as RGB frames between calls)"""
    import torch
    from bwd_nlkalman_b200 import synth
    from oracle import oracle as O
    w, h, ch, sigma, nf = 160, 120, 3, 20.0, 20
    f1, f2, s1 = (nlk.default_params(sigma, m) for m in (nlk.FLT1, nlk.FLT2, nlk.SMO1))
    rf1, rf2, rs1 = (_same_params(nlk, O, p) for p in (f1, f2, s1))
    frames = [synth.noisy_frame(w, h, ch, t, sigma) for t in range(nf)]
    clean = [synth.clean_frame(w, h, ch, t) for t in range(nf)]
    bflo, fflo, occ = synth.backward_flow(w, h), synth.forward_flow(w, h), synth.occlusion_mask(w, h)

    # reference chain (scripts/nlkalman-seq.sh:39-41, :56-102, :122-149)
    r1, r2, rs = [], [], [None] * nf
    for t in range(nf):
        n = ref.rgb2opp(frames[t].copy())
        if t == 0:
            a = ref.filter_frame(n, None, None, sigma, rf1)
            b = ref.filter_frame(n, None, a, sigma, rf2)
        else:
            p1, p2 = ref.rgb2opp(r1[-1].copy()), ref.rgb2opp(r2[-1].copy())
            a = ref.filter_frame(n, ref.warp_bicubic(p1, bflo, occ), None, sigma, rf1)
            b = ref.filter_frame(n, ref.warp_bicubic(p2, bflo, occ), a, sigma, rf2)
        r1.append(ref.opp2rgb(a.copy()))
        r2.append(ref.opp2rgb(b.copy()))
    rs[-1] = r2[-1]
    for t in range(nf - 2, -1, -1):
        fl, nx = ref.rgb2opp(r2[t].copy()), ref.rgb2opp(rs[t + 1].copy())
        rs[t] = ref.opp2rgb(ref.smooth_frame(fl, ref.warp_bicubic(nx, fflo, occ), None, sigma, rs1))

    # ours: resident recursion, pipelined submits, outputs per frame
    dev = torch.device("cuda", 0)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_fr, d_bflo, d_fflo, d_occ = [up(f) for f in frames], up(bflo), up(fflo), up(occ)
    g1 = [torch.empty_like(d_fr[0]) for _ in range(nf)]
    g2 = [torch.empty_like(d_fr[0]) for _ in range(nf)]
    gs = [torch.empty_like(d_fr[0]) for _ in range(nf)]
    with nlk.Context(w, h, ch) as ctx:
        for t in range(nf):
            ctx.seq_submit_dev(d_fr[t], d_bflo if t else None, d_occ if t else None, sigma, f1, f2, g1[t], g2[t])
        ctx.seq_drain()
        ctx.seq_smooth_start_dev(g2[-1])
        gs[-1].copy_(g2[-1])
        for t in range(nf - 2, -1, -1):
            ctx.seq_smooth_dev(g2[t], d_fflo, d_occ, sigma, s1, gs[t])
        ctx.sync()
    rows = []
    for t in range(nf):
        row = []
        for ours, theirs in ((g1[t], r1[t]), (g2[t], r2[t]), (gs[t], rs[t])):
            o = ours.cpu().numpy()
            row.append((maxabs(o, theirs), _frac_above(o, theirs),
                        abs(psnr_between(o, clean[t]) - psnr_between(theirs, clean[t]))))
        rows.append(row)
        print(f"frame {t:2d}: " + "  ".join(f"{nm} max-abs {e:.1e} frac>1e-3 {fr:.1e} dPSNR {dp:.1e}"
                                              for nm, (e, fr, dp) in zip(("flt1", "flt2", "smo1"), row)))
    for t, row in enumerate(rows):
        for nm, (e, fr, dp) in zip(("flt1", "flt2", "smo1"), row):
            assert dp <= TOL_DPSNR, (t, nm, dp)
            assert fr <= 5e-2, (t, nm, fr)
    # the first frames, before any near-tie can have cascaded through the recursion
    assert rows[0][0][0] <= TOL_MAXABS and rows[0][1][0] <= TOL_MAXABS
    assert rows[1][0][0] <= TOL_MAXABS


def test_config4_eight_virtual_strips_2160p(nlk):
    """3840x2160x3, sigma 10: one temporal frame (flt1 + flt2) and one smoothing step in 8 strips
    (each its own context and slab, exchanges by the peer-memory kernels of the nlk_peer_* ABI) against the single-context recursion"""
    import torch
    from bwd_nlkalman_b200 import strips, synth
    w, h, ch, sigma, nranks = 3840, 2160, 3, 10.0, 8
    f1, f2, s1 = (nlk.default_params(sigma, m) for m in (nlk.FLT1, nlk.FLT2, nlk.SMO1))
    frames = [synth.noisy_frame(w, h, ch, t, sigma) for t in range(2)]
    bflo, fflo, occ = synth.backward_flow(w, h), synth.forward_flow(w, h), synth.occlusion_mask(w, h)
    dev = torch.device("cuda", 0)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_fr, d_bflo, d_fflo, d_occ = [up(f) for f in frames], up(bflo), up(fflo), up(occ)
    ref1, ref2 = [], []
    with nlk.Context(w, h, ch) as ctx:
        o1, o2 = torch.empty_like(d_fr[0]), torch.empty_like(d_fr[0])
        for t in range(2):
            ctx.seq_filter_dev(d_fr[t], d_bflo if t else None, d_occ if t else None, sigma, f1, f2, o1, o2)
            ctx.sync()
            ref1.append(o1.cpu().numpy().copy())
            ref2.append(o2.cpu().numpy().copy())
        ctx.seq_smooth_start_dev(up(ref2[1]))
        ctx.seq_smooth_dev(up(ref2[0]), d_fflo, d_occ, sigma, s1, o1)
        ctx.sync()
        refs0 = o1.cpu().numpy().copy()
    ranks = [strips.StripRank(w, h, ch, r, nranks, 0, transport="peer") for r in range(nranks)]
    strips.bind_virtual(ranks)
    try:
        outs1 = [torch.zeros_like(d_fr[0]) for _ in ranks]
        outs2 = [torch.zeros_like(d_fr[0]) for _ in ranks]

        def assemble(outs, plans):
            full = np.empty((h, w, ch), np.float32)
            for r, p in enumerate(plans):
                full[p.oy0:p.oy1] = outs[r][p.oy0:p.oy1].cpu().numpy()
            return full

        def near(a, b, what):
            d = np.abs(a.astype(np.float64) - b)
            frac = float((d > TOL_MAXABS).mean())
            print(f"  {what}: max-abs {d.max():.2e}, mean {d.mean():.1e}, {frac:.1e} of pixels > 1e-3")
            assert frac <= 2e-4 and d.mean() <= 2e-5, (what, d.max(), frac)
        pl1, pl2, pls = ranks[0].plans(0, f1), ranks[0].plans(0, f2), ranks[0].plans(1, s1)
        for t in range(2):
            strips.run_virtual(ranks, [rk.filter_step(d_fr[t], d_bflo if t else None, d_occ if t else None,
                                                      sigma, f1, f2, outs1[r], outs2[r]) for r, rk in enumerate(ranks)])
            for rk in ranks:
                rk.ctx.sync()
            # first filtering: searched on the noisy frame, identical lists on both sides
            assert maxabs(assemble(outs1, pl1), ref1[t]) <= TOL_MAXABS, f"flt1 frame {t}"
            # second filtering and smoother search on an OUTPUT of the run itself, whose last bits
            # depend on the order of the floating-point reductions of the aggregation: over 517,000
            # groups a distance near-tie flips in most pairs of runs (also of one context twice)
            near(assemble(outs2, pl2), ref2[t], f"flt2 frame {t}")
        last, flt = up(ref2[1]), up(ref2[0])
        strips.run_virtual(ranks, [rk.smooth_start(last) for rk in ranks])
        strips.run_virtual(ranks, [rk.smooth_step(flt, d_fflo, d_occ, sigma, s1, outs1[r]) for r, rk in enumerate(ranks)])
        for rk in ranks:
            rk.ctx.sync()
        near(assemble(outs1, pls), refs0, "smoother frame 0")
        assert all(rk.ctx.peer_error() == 0 for rk in ranks), "a device-side wait timed out"
    finally:
        for rk in ranks:
            rk.close()

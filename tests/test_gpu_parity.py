"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU
checkers on the same inputs.  Tolerance (north_star): max abs error <= 1e-3 on the 0-255
scale, |dPSNR| <= 0.01 dB; k-NN index lists identical (distances carry the reference's
bits; any mismatch is reported with its distance gap)."""
import os

import numpy as np
import pytest

from common import (TOL_DPSNR, TOL_MAXABS, compare_knn, golden_cases, load_golden, maxabs,
                    params_from_array, psnr_between)

pytestmark = pytest.mark.gpu


def _np(nlk, arr):
    return params_from_array(nlk.Params, arr)


def _same_params(nlk, O, p):
    return O.Params(*[getattr(p, f) for f, _ in nlk.Params._fields_])


# ---- unit stages -------------------------------------------------------------------------------

@pytest.mark.parametrize("psz", [8, 12, 4, 6, 7, 16])
def test_dct_unit(nlk, port, psz):
    """batched orthonormal DCT-II / DCT-III (reference src/nlkalman.c:248-360)"""
    rng = np.random.default_rng(psz)
    t = rng.normal(100, 60, (301, psz, psz)).astype(np.float32)
    with nlk.Context(64, 64, 1) as ctx:
        y = ctx.dct(t)
        want = port.dct2(t)
        # coefficients reach ~psz*255; fp32 matrix form vs double reference
        assert np.abs(y - want).max() <= 2e-3 * psz / 8
        back = ctx.dct(y, inverse=True)
        assert np.abs(back - t).max() <= 1e-3
        assert np.abs(ctx.dct(want, inverse=True) - port.dct2(want, inverse=True)).max() <= 1e-3


def test_default_params_match_oracle(nlk, port):
    from oracle import oracle as O
    for mode in (nlk.FLT1, nlk.FLT2, nlk.SMO1):
        for s in (0.0, 3.0, 10.0, 20.0, 25.5, 40.0, 57.0):
            assert nlk.default_params(s, mode).as_dict() == port.default_params(s, mode).as_dict()
    p = nlk.default_params(20, nlk.FLT2, nlk.Params.auto(patch_sz=12, npatches_t=7))
    q = port.default_params(20, O.FLT2, O.Params.auto(patch_sz=12, npatches_t=7))
    assert p.as_dict() == q.as_dict()


@pytest.mark.parametrize("shape", [(33, 21, 3), (64, 48, 1), (40, 30, 2)])
def test_colour_and_warp(nlk, port, shape):
    w, h, ch = shape
    rng = np.random.default_rng(3)
    im = rng.uniform(0, 255, (h, w, ch)).astype(np.float32)
    a = nlk.rgb2opp(im.copy())
    b = port.rgb2opp(im.copy())
    assert np.array_equal(a, b)  # same operations in the same order: bit-exact
    assert np.array_equal(nlk.opp2rgb(a.copy()), port.opp2rgb(b.copy()))
    of = rng.uniform(-3, 3, (h, w, 2)).astype(np.float32)
    msk = (rng.uniform(0, 1, (h, w)) > 0.9).astype(np.float32) * 255
    for m in (msk, None):
        wa = nlk.warp_bicubic(im, of, m)
        wb = port.warp_bicubic(im, of, m)
        assert maxabs(wa, wb) <= 3.1e-5  # one ulp at 255 (double Horner, fma contraction)
    # integer flow copies pixels exactly
    of[...] = 0
    of[..., 0] = 1
    wa = nlk.warp_bicubic(im, of, None)
    assert np.array_equal(wa[2:-2, 1:-3], im[2:-2, 2:-2])


# ---- golden vectors of the unmodified reference ----------------------------------------------

@pytest.mark.parametrize("name", golden_cases())
def test_golden_reference_vectors(nlk, name):
    g = load_golden(name)
    sigma = float(g["sigma"])
    f1, f2, s1 = (_np(nlk, g[k]) for k in ("f1", "f2", "s1"))
    o1 = nlk.rgb2opp(g["noisy1"].copy())
    assert maxabs(nlk.rgb2opp(g["noisy0"].copy()), g["opp0"]) <= 5e-5
    assert maxabs(nlk.nlkalman_filter_frame(g["opp0"], None, None, sigma, f1), g["flt1_0"]) <= TOL_MAXABS
    assert maxabs(nlk.nlkalman_filter_frame(g["opp0"], None, g["flt1_0"], sigma, f2), g["flt2_0"]) <= TOL_MAXABS
    assert maxabs(nlk.warp_bicubic(g["flt1_0"], g["bflo"], g["occ"]), g["warp1"]) <= 6.2e-5
    assert maxabs(nlk.nlkalman_filter_frame(o1, g["warp1"], None, sigma, f1), g["flt1_1"]) <= TOL_MAXABS
    assert maxabs(nlk.nlkalman_filter_frame(o1, g["warp2"], g["flt1_1"], sigma, f2), g["flt2_1"]) <= TOL_MAXABS
    assert maxabs(nlk.nlkalman_smooth_frame(g["flt2_0"], g["warps"], None, sigma, s1), g["smo_0"]) <= TOL_MAXABS
    assert maxabs(nlk.opp2rgb(g["flt2_1"].copy()), g["rgb_flt2_1"]) <= 5e-5


# ---- stage-level parity against the restatement's dumps -------------------------------------

def _stage_check(nlk, port, O, smooth, in1, prev0, bsic1, sigma, prms):
    h, w, ch = in1.shape
    with nlk.Context(w, h, ch) as ctx:
        out, gd = ctx.pass_host_debug(smooth, in1, prev0, bsic1, sigma, prms)
    cpu_out, cd = port.run_pass(O.PASS_SMOOTH if smooth else O.PASS_FILTER, in1, prev0, bsic1, sigma,
                                _same_params(nlk, O, prms), dump=True)
    assert np.array_equal(gd["prev_p"], cd["prev_p"])
    assert np.array_equal(gd["nk"], cd["nk"])
    assert np.array_equal(gd["np0"], cd["np0"])
    G, nbad, details = compare_knn(gd, cd)
    assert nbad == 0, f"{nbad}/{G} k-NN lists differ (g, rank, distance gap): {details}"
    assert np.array_equal(gd["knn_d"], cd["knn_d"]), "distances are not bit-identical"
    assert np.array_equal(gd["active"], cd["active"]), "processed-patch sets differ"
    act = cd["active"].astype(bool)
    rel = np.abs(gd["vp"][act] - cd["vp"][act]) / np.maximum(np.abs(cd["vp"][act]), 1e-3)
    assert rel.max() <= 2e-4, rel.max()
    assert maxabs(out, cpu_out) <= TOL_MAXABS
    return out, cpu_out, cd


@pytest.mark.parametrize("shape", [(122, 90, 1), (101, 77, 3)])
def test_stage_parity_filter_and_smoother(nlk, port, shape):
    from bwd_nlkalman_b200 import synth
    from oracle import oracle as O
    w, h, ch = shape
    sigma = 20.0
    f1, f2, s1 = (nlk.default_params(sigma, m) for m in (nlk.FLT1, nlk.FLT2, nlk.SMO1))
    n0 = port.rgb2opp(synth.noisy_frame(w, h, ch, 0, sigma))
    n1 = port.rgb2opp(synth.noisy_frame(w, h, ch, 1, sigma))
    bflo, fflo = synth.backward_flow(w, h), synth.forward_flow(w, h)
    occ = np.zeros((h, w), np.float32)
    occ[h // 3:h // 3 + 14, w // 2:w // 2 + 20] = 255
    # frame 0, spatial (processed mask skips ~30% of the patches)
    _, c11, d = _stage_check(nlk, port, O, 0, n0, None, None, sigma, f1)
    assert d["active"].mean() < 0.95
    _, c21, _ = _stage_check(nlk, port, O, 0, n0, None, c11, sigma, f2)
    # frame 1, temporal, with occlusion + NaN border
    w1 = port.warp_bicubic(c11, bflo, occ)
    w2 = port.warp_bicubic(c21, bflo, occ)
    _, c12, d = _stage_check(nlk, port, O, 0, n1, w1, None, sigma, f1)
    assert 0 < (d["np0"] == 0).sum() < d["np0"].size  # both Kalman and spatial-fallback groups
    _, c22, _ = _stage_check(nlk, port, O, 0, n1, w2, c12, sigma, f2)
    # smoother on frame 0
    ws = port.warp_bicubic(c22, fflo, occ)
    _stage_check(nlk, port, O, 1, c21, ws, None, sigma, s1)


def test_stage_parity_patch12_wide_windows(nlk, port):
    """config-3 style parameters: 12x12 patches, radius 10 temporal / 15 spatial"""
    from bwd_nlkalman_b200 import synth
    from oracle import oracle as O
    w, h, ch, sigma = 98, 74, 3, 40.0
    ov = dict(patch_sz=12, search_sz_t=10, search_sz_x=15)
    f1 = nlk.default_params(sigma, nlk.FLT1, nlk.Params.auto(**ov))
    s1 = nlk.default_params(sigma, nlk.SMO1, nlk.Params.auto(patch_sz=12, search_sz_t=10))
    n0 = port.rgb2opp(synth.noisy_frame(w, h, ch, 0, sigma))
    n1 = port.rgb2opp(synth.noisy_frame(w, h, ch, 1, sigma))
    _, c11, _ = _stage_check(nlk, port, O, 0, n0, None, None, sigma, f1)
    w1 = port.warp_bicubic(c11, synth.backward_flow(w, h), None)
    _, c12, _ = _stage_check(nlk, port, O, 0, n1, w1, None, sigma, f1)
    ws = port.warp_bicubic(c12, synth.forward_flow(w, h), None)
    _stage_check(nlk, port, O, 1, c11, ws, None, sigma, s1)


@pytest.mark.parametrize("psz,ch", [(6, 1), (10, 3), (7, 2), (16, 1), (4, 4)])
def test_generic_patch_sizes(nlk, port, psz, ch):
    """run-time patch size / channel count path of the kernels"""
    from bwd_nlkalman_b200 import synth
    from oracle import oracle as O
    w, h, sigma = 70, 58, 15.0
    f1 = nlk.default_params(sigma, nlk.FLT1, nlk.Params.auto(patch_sz=psz, search_sz_x=6, npatches_x=25))
    rng = np.random.default_rng(5)
    base = synth.noisy_frame(w, h, 3, 0, sigma)
    n0 = np.ascontiguousarray(np.concatenate([base, base[..., :1] + rng.normal(0, 3, (h, w, 1)).astype(np.float32)],
                                             axis=2)[..., :ch])
    _stage_check(nlk, port, O, 0, n0, None, None, sigma, f1)


@pytest.mark.parametrize("ch", [1, 3])
def test_team_kernel_variants(nlk, port, ch):
    """The 8x8 team kernel is instantiated per update mode (tile per lane / row-column per lane), per
    basic-estimate flag, and the search per selection network (32 smallest / full sort): group sizes
    and candidate counts on both sides of each switch, on a temporal frame with an occlusion."""
    from bwd_nlkalman_b200 import synth
    from oracle import oracle as O
    w, h, sigma = 93, 70, 20.0
    f1 = nlk.default_params(sigma, nlk.FLT1)
    n0 = port.rgb2opp(synth.noisy_frame(w, h, ch, 0, sigma))
    n1 = port.rgb2opp(synth.noisy_frame(w, h, ch, 1, sigma))
    bflo, fflo = synth.backward_flow(w, h), synth.forward_flow(w, h)
    occ = np.zeros((h, w), np.float32)
    occ[20:34, 40:60] = 255
    c11 = port.filter_frame(n0, None, None, sigma, _same_params(nlk, O, f1))
    w1 = port.warp_bicubic(c11, bflo, occ)
    # first filtering: one statistics round (k small), 32 / 33 candidates kept (selection network switch),
    # groups of one or two members through the tile-per-lane update
    for ov in (dict(npatches_t=8, npatches_tagg=20), dict(npatches_t=32), dict(npatches_t=33),
               dict(npatches_t=30, npatches_tagg=9)):
        _stage_check(nlk, port, O, 0, n1, w1, None, sigma, nlk.default_params(sigma, nlk.FLT1, nlk.Params.auto(**ov)))
    c12 = port.filter_frame(n1, w1, None, sigma, _same_params(nlk, O, f1))
    # second filtering: 1, 2 members (row / column update where tagg * ch <= 8, noisy patches prefetched),
    # 5 members (more than the prefetch area holds: restaged after the gains), a single statistics round
    for ov in (dict(), dict(npatches_tagg=2), dict(npatches_tagg=5), dict(npatches_t=6), dict(npatches_t=6, npatches_tagg=3)):
        _stage_check(nlk, port, O, 0, n1, w1, c12, sigma, nlk.default_params(sigma, nlk.FLT2, nlk.Params.auto(**ov)))
    # smoother: small groups (row / column update), 32 candidates (selection network), default
    ws = port.warp_bicubic(c12, fflo, occ)
    for ov in (dict(npatches_t=12, npatches_tagg=2), dict(npatches_t=32, npatches_tagg=32), dict()):
        _stage_check(nlk, port, O, 1, c11, ws, None, sigma, nlk.default_params(sigma, nlk.SMO1, nlk.Params.auto(**ov)))


# ---- edge cases the reference handles ----------------------------------------------------------

def test_edge_cases(nlk, port):
    from bwd_nlkalman_b200 import synth
    from oracle import oracle as O
    sigma = 20.0
    f1 = nlk.default_params(sigma, nlk.FLT1)
    s1 = nlk.default_params(sigma, nlk.SMO1)
    # (a) frame smaller than a patch: no patch at all, output falls back to the input
    tiny = synth.noisy_frame(7, 5, 1, 0, sigma)
    assert np.array_equal(nlk.nlkalman_filter_frame(tiny, None, None, sigma, f1), tiny)
    # (b) width/height not reachable by the grid (reference FIXMEs :587,:595), ragged sizes
    n0 = synth.noisy_frame(45, 31, 1, 0, sigma)
    _stage_check(nlk, port, O, 0, n0, None, None, sigma, f1)
    # (c) previous frame entirely invalid (all NaN): every group takes the spatial branch
    allnan = np.full_like(n0, np.nan)
    _stage_check(nlk, port, O, 0, n0, allnan, None, sigma, f1)
    # (d) smoother with no valid previous patch anywhere: copies the filtered frame
    out = nlk.nlkalman_smooth_frame(n0, allnan, None, sigma, s1)
    assert maxabs(out, port.smooth_frame(n0, allnan, None, sigma, _same_params(nlk, O, s1))) <= TOL_MAXABS
    # (e) k <= 1: the filter aggregates nothing and returns the noisy frame
    f1k = nlk.default_params(sigma, nlk.FLT1, nlk.Params.auto(npatches_x=1))
    assert np.array_equal(nlk.nlkalman_filter_frame(n0, None, None, sigma, f1k), n0)
    # (f) k larger than the number of candidates in the window is clamped (:707)
    f1big = nlk.default_params(sigma, nlk.FLT1, nlk.Params.auto(search_sz_x=2, npatches_x=60))
    _stage_check(nlk, port, O, 0, n0, None, None, sigma, f1big)
    # (g) exact distance ties (constant image): stable order by scan index
    flat = np.full((40, 36, 1), 100.0, np.float32)
    _stage_check(nlk, port, O, 0, flat, None, None, sigma, f1)


# ---- full-size config 1 against the unmodified reference -------------------------------------

def test_config1_against_reference_library(nlk, ref):
    """BASELINE config 1: 854x480 gray, sigma 20, auto params, one nlkalman-flt step on
    frame 1 with bflo/bocc and the previous filtered frames, vs the reference at 1 thread"""
    from bwd_nlkalman_b200 import synth
    from oracle import oracle as O
    w, h, ch, sigma = 854, 480, 1, 20.0
    f1, f2 = nlk.default_params(sigma, nlk.FLT1), nlk.default_params(sigma, nlk.FLT2)
    rf1, rf2 = _same_params(nlk, O, f1), _same_params(nlk, O, f2)
    n0, n1 = synth.noisy_frame(w, h, ch, 0, sigma), synth.noisy_frame(w, h, ch, 1, sigma)
    clean1 = synth.clean_frame(w, h, ch, 1)
    bflo, occ = synth.backward_flow(w, h), synth.occlusion_mask(w, h)
    # previous-frame state from the GPU path itself (frame 0 is checked in the other tests)
    p11 = nlk.nlkalman_filter_frame(n0, None, None, sigma, f1)
    p21 = nlk.nlkalman_filter_frame(n0, None, p11, sigma, f2)
    w1, w2 = nlk.warp_bicubic(p11, bflo, occ), nlk.warp_bicubic(p21, bflo, occ)
    assert maxabs(w1, ref.warp_bicubic(p11, bflo, occ)) <= 6.2e-5
    g12 = nlk.nlkalman_filter_frame(n1, w1, None, sigma, f1)
    r12 = ref.filter_frame(n1, w1, None, sigma, rf1)
    assert maxabs(g12, r12) <= TOL_MAXABS
    g22 = nlk.nlkalman_filter_frame(n1, w2, r12, sigma, f2)
    r22 = ref.filter_frame(n1, w2, r12, sigma, rf2)
    assert maxabs(g22, r22) <= TOL_MAXABS
    assert abs(psnr_between(g22, clean1) - psnr_between(r22, clean1)) <= TOL_DPSNR
    assert psnr_between(g22, clean1) > psnr_between(n1, clean1) + 8  # it does denoise


# ---- streaming (pipelined) host recursion ------------------------------------------------------

def test_pipelined_host_recursion_matches_synchronous(nlk):
    """nlk_seq_submit_host / nlk_seq_drain (uploads, kernels, downloads on their own streams, three
    frames in flight) give the frames nlk_seq_filter_host gives, in order"""
    import torch
    from bwd_nlkalman_b200 import synth
    w, h, ch, sigma, nf = 131, 94, 3, 20.0, 6
    f1, f2 = nlk.default_params(sigma, nlk.FLT1), nlk.default_params(sigma, nlk.FLT2)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    frames = [pin(synth.noisy_frame(w, h, ch, t, sigma)) for t in range(nf)]
    bflo, occ = pin(synth.backward_flow(w, h)), pin(synth.occlusion_mask(w, h))
    sync1 = [torch.empty((h, w, ch)).pin_memory() for _ in range(nf)]
    sync2 = [torch.empty((h, w, ch)).pin_memory() for _ in range(nf)]
    pipe1 = [torch.empty((h, w, ch)).pin_memory() for _ in range(nf)]
    pipe2 = [torch.empty((h, w, ch)).pin_memory() for _ in range(nf)]
    with nlk.Context(w, h, ch) as ctx:
        for t in range(nf):
            ctx.seq_filter_host(frames[t], bflo if t else None, occ if t else None, sigma, f1, f2, sync1[t], sync2[t])
        ctx.seq_reset()
        for t in range(nf):
            ctx.seq_submit_host(frames[t], bflo if t else None, occ if t else None, sigma, f1, f2, pipe1[t], pipe2[t])
        ctx.seq_drain()
    for t in range(nf):
        assert maxabs(pipe1[t].numpy(), sync1[t].numpy()) <= TOL_MAXABS, t
        assert maxabs(pipe2[t].numpy(), sync2[t].numpy()) <= TOL_MAXABS, t
    assert float(np.abs(sync2[-1].numpy() - frames[-1].numpy()).mean()) > 1.0   # it did filter


# ---- occlusion mask from the flow ---------------------------------------------

def test_occlusion_from_flow_divergence(nlk, tmp_path):
    """the plambda expression of reference scripts/nlkalman-seq.sh:70-72, bit for bit, through the
    C ABI and through the nlkalman-occ program"""
    import subprocess
    from oracle import oracle as O
    rng = np.random.default_rng(11)
    w, h = 157, 93
    of = rng.normal(0, 0.6, (h, w, 2)).astype(np.float32)
    of[20:40, 30:60] += rng.normal(0, 2.0, (20, 30, 2)).astype(np.float32)   # a disoccluded region
    with nlk.Context(w, h, 1) as ctx:
        for th in (0.25, 0.75, 2.0):
            want = O.occlusion_from_flow(of, th)
            got = ctx.occlusion(of, th)
            assert np.array_equal(got, want)
            assert 0 < int((want > 0).sum()) < want.size
    # the program: .flo in, PNG out
    import struct
    flo, png = tmp_path / "f.flo", tmp_path / "o.png"
    with open(flo, "wb") as f:
        f.write(b"PIEH" + struct.pack("<ii", w, h) + of.tobytes())
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bwd_nlkalman_b200", "bin", "nlkalman-occ")
    subprocess.run([exe, str(flo), "0.75", str(png)], check=True)
    from PIL import Image
    assert np.array_equal(np.asarray(Image.open(png)).astype(np.float32), O.occlusion_from_flow(of, 0.75))


def test_two_lane_device_recursion_matches_in_order(nlk):
    """nlk_seq_submit_dev (second filtering of frame t on a second stream beside the first
    filtering of frame t+1) against nlk_seq_filter_dev (everything in order), frame by frame,
    distinct output buffers per frame; then the forms mixed in one sequence"""
    import torch
    from bwd_nlkalman_b200 import synth
    w, h, ch, sigma, nf = 150, 110, 3, 20.0, 7
    f1, f2 = nlk.default_params(sigma, nlk.FLT1), nlk.default_params(sigma, nlk.FLT2)
    dev = torch.device("cuda", 0)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    frames = [up(synth.noisy_frame(w, h, ch, t, sigma)) for t in range(nf)]
    flo = [up(synth.backward_flow(w, h)) for _ in range(nf)]
    occ = [up(synth.occlusion_mask(w, h)) for _ in range(nf)]
    mk = lambda: [torch.empty_like(frames[0]) for _ in range(nf)]
    a1, a2, b1, b2, c2 = mk(), mk(), mk(), mk(), mk()
    with nlk.Context(w, h, ch) as ctx:
        for t in range(nf):
            ctx.seq_filter_dev(frames[t], flo[t] if t else None, occ[t] if t else None, sigma, f1, f2, a1[t], a2[t])
        ctx.sync()
        ctx.seq_reset()
        for t in range(nf):
            ctx.seq_submit_dev(frames[t], flo[t] if t else None, occ[t] if t else None, sigma, f1, f2, b1[t], b2[t])
        ctx.seq_drain()
        ctx.seq_reset()
        for t in range(nf):   # mixed: every third frame in order
            call = ctx.seq_filter_dev if t % 3 == 2 else ctx.seq_submit_dev
            call(frames[t], flo[t] if t else None, occ[t] if t else None, sigma, f1, f2, None, c2[t])
        ctx.seq_join()
        ctx.sync()
    for t in range(nf):
        assert maxabs(b1[t].cpu().numpy(), a1[t].cpu().numpy()) <= TOL_MAXABS, t
        assert maxabs(b2[t].cpu().numpy(), a2[t].cpu().numpy()) <= TOL_MAXABS, t
        assert maxabs(c2[t].cpu().numpy(), a2[t].cpu().numpy()) <= TOL_MAXABS, t


# ---- branches outside the drivers' call patterns ---------------------------------------------

def test_smoother_with_basic_estimate(nlk, port):
    """nlkalman_smooth_frame with bsic1 != NULL: search and statistics on bsic1, the updated
    members from filt1 (reference src/nlkalman.c:1669).  No reference driver passes one
    (src/main-smo.c:209), the entry point accepts it: 8x8 RGB and gray, and 12x12."""
    from bwd_nlkalman_b200 import synth
    from oracle import oracle as O
    sigma = 20.0
    for (w, h, ch, ov) in ((93, 70, 3, {}), (88, 66, 1, {}), (90, 66, 3, dict(patch_sz=12, search_sz_t=6))):
        s1 = nlk.default_params(sigma, nlk.SMO1, nlk.Params.auto(npatches_t=24, npatches_tagg=10, **ov))
        rng = np.random.default_rng(w)
        clean0 = port.rgb2opp(synth.clean_frame(w, h, ch, 0))
        flt = clean0 + rng.normal(0, 4, clean0.shape).astype(np.float32)      # "filtered" frame t
        bsic = clean0 + rng.normal(0, 2, clean0.shape).astype(np.float32)     # a better estimate of it
        nxt = port.rgb2opp(synth.clean_frame(w, h, ch, 1)) + rng.normal(0, 2, clean0.shape).astype(np.float32)
        occ = np.zeros((h, w), np.float32)
        occ[20:34, 40:60] = 255
        ws = port.warp_bicubic(nxt, synth.forward_flow(w, h), occ)
        out, cpu_out, cd = _stage_check(nlk, port, O, 1, flt, ws, bsic, sigma, s1)
        assert (cd["np0"] > 0).any() and (cd["np0"] == 0).any()
        # the basic estimate matters: without it the result differs
        plain = port.smooth_frame(flt, ws, None, sigma, _same_params(nlk, O, s1))
        assert maxabs(cpu_out, plain) > 0.05


def test_smoother_single_patch_branch_is_rejected(nlk):
    """--s1_nt <= 1 with a valid next frame: the reference's branch (src/nlkalman.c:1699-1730)
    aggregates at uninitialised coordinates -- there is no defined result, the
    library refuses the configuration instead of inventing one.  Without a next frame (plain copy,
    :1795-1804) and for the filter (k <= 1: nothing aggregated, :815-849) the call stays valid."""
    from bwd_nlkalman_b200 import synth
    sigma, w, h, ch = 20.0, 64, 48, 1
    n0 = synth.noisy_frame(w, h, ch, 0, sigma)
    prev = synth.noisy_frame(w, h, ch, 1, sigma)
    for nt in (1, 0):
        s1 = nlk.default_params(sigma, nlk.SMO1, nlk.Params.auto(npatches_t=nt, npatches_tagg=1))
        with nlk.Context(w, h, ch) as ctx:
            with pytest.raises(nlk.NlkError, match="npatches_t"):
                ctx.pass_host_debug(1, n0, prev, None, sigma, s1)
            out, _ = ctx.pass_host_debug(1, n0, None, None, sigma, s1)     # no next frame: copy
            assert maxabs(out, n0) <= TOL_MAXABS


def test_mask_modes_of_the_streaming_recursion(nlk):
    """nlk_seq_set_mask_mode: the mask as the bytes of its 8-bit file, or built on the device from the
    divergence of the uploaded flow (the script's plambda expression), against the float mask"""
    import torch
    from bwd_nlkalman_b200 import synth
    from oracle import oracle as O
    w, h, ch, sigma, nf, th = 140, 100, 3, 20.0, 4, 0.75
    f1, f2 = nlk.default_params(sigma, nlk.FLT1), nlk.default_params(sigma, nlk.FLT2)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    frames = [pin(synth.noisy_frame(w, h, ch, t, sigma)) for t in range(nf)]
    # a flow with a divergent region, so that the divergence mask is not empty
    rng = np.random.default_rng(5)
    of = synth.backward_flow(w, h).copy()
    of[30:50, 60:90] += rng.normal(0, 1.5, (20, 30, 2)).astype(np.float32)
    occ = O.occlusion_from_flow(of, th)
    assert 0 < int((occ > 0).sum()) < occ.size
    flo, occ_f, occ_8 = pin(of), pin(occ), pin(occ.astype(np.uint8))
    outs = {}
    with nlk.Context(w, h, ch) as ctx:
        for name, mode, mask in (("float", ctx.MASK_FLOAT, occ_f), ("u8", ctx.MASK_U8, occ_8), ("flow", ctx.MASK_FROM_FLOW, None)):
            ctx.seq_reset()
            ctx.seq_set_mask_mode(mode, th)
            o1 = [torch.empty((h, w, ch)).pin_memory() for _ in range(nf)]
            o2 = [torch.empty((h, w, ch)).pin_memory() for _ in range(nf)]
            for t in range(nf):
                ctx.seq_submit_host(frames[t], flo if t else None, mask if t else None, sigma, f1, f2, o1[t], o2[t])
            ctx.seq_drain()
            outs[name] = ([x.numpy().copy() for x in o1], [x.numpy().copy() for x in o2])
        ctx.seq_set_mask_mode(ctx.MASK_FLOAT)
    for t in range(nf):
        for name in ("u8", "flow"):
            # first filtering: searched on the noisy frame -- the same lists in every run, only the
            # reduction order of the aggregation differs
            assert maxabs(outs[name][0][t], outs["float"][0][t]) <= TOL_MAXABS, (name, t)
            # second filtering: searched on the run's own first filtering, whose last bits differ from run
            # to run, so a distance near-tie may flip a group (also between two runs of one mode)
            d = np.abs(outs[name][1][t].astype(np.float64) - outs["float"][1][t])
            assert (d > TOL_MAXABS).mean() <= 2e-2 and d.mean() <= 1e-3, (name, t, d.max(), (d > TOL_MAXABS).mean())
    assert float(np.abs(outs["float"][1][-1] - frames[-1].numpy()).mean()) > 1.0

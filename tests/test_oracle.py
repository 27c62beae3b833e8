"""CPU tests of the checker itself: the C restatement (oracle/nlk_port.c) against the
golden vectors produced by the unmodified reference (tests/golden/make_golden.py), and,
where oracle/_ref is present, against the reference library directly."""
import numpy as np
import pytest

from common import TOL_MAXABS, golden_cases, load_golden, maxabs, params_from_array


def _port_params(O, arr):
    return params_from_array(O.Params, arr)


@pytest.mark.parametrize("name", golden_cases())
def test_port_matches_reference_golden(port, name):
    from oracle import oracle as O
    g = load_golden(name)
    sigma = float(g["sigma"])
    f1, f2, s1 = (_port_params(O, g[k]) for k in ("f1", "f2", "s1"))
    o0 = port.rgb2opp(g["noisy0"].copy())
    o1 = port.rgb2opp(g["noisy1"].copy())
    assert maxabs(o0, g["opp0"]) <= 5e-5
    flt1_0 = port.filter_frame(g["opp0"], None, None, sigma, f1)
    assert maxabs(flt1_0, g["flt1_0"]) <= TOL_MAXABS
    flt2_0 = port.filter_frame(g["opp0"], None, g["flt1_0"], sigma, f2)
    assert maxabs(flt2_0, g["flt2_0"]) <= TOL_MAXABS
    warp1 = port.warp_bicubic(g["flt1_0"], g["bflo"], g["occ"])
    assert maxabs(warp1, g["warp1"]) <= 1e-5
    flt1_1 = port.filter_frame(o1, g["warp1"], None, sigma, f1)
    assert maxabs(flt1_1, g["flt1_1"]) <= TOL_MAXABS
    flt2_1 = port.filter_frame(o1, g["warp2"], g["flt1_1"], sigma, f2)
    assert maxabs(flt2_1, g["flt2_1"]) <= TOL_MAXABS
    smo = port.smooth_frame(g["flt2_0"], g["warps"], None, sigma, s1)
    assert maxabs(smo, g["smo_0"]) <= TOL_MAXABS
    rgb = port.opp2rgb(g["flt2_1"].copy())
    assert maxabs(rgb, g["rgb_flt2_1"]) <= 5e-5


def test_default_params_table(port):
    """SURVEY.md App. A table (reference src/nlkalman.c:456-486)"""
    from oracle import oracle as O
    want = {10: ((45, 30, 20), (15, 10, 1), 15), 20: ((50, 30, 20), (20, 20, 1), 45),
            30: ((55, 30, 20), (25, 30, 1), 75), 40: ((60, 30, 20), (30, 40, 1), 105)}
    for s, (a, b, c) in want.items():
        f1 = port.default_params(s, O.FLT1)
        f2 = port.default_params(s, O.FLT2)
        s1 = port.default_params(s, O.SMO1)
        assert (f1.npatches_x, f1.npatches_t, f1.npatches_tagg) == a
        assert (f2.npatches_x, f2.npatches_t, f2.npatches_tagg) == b
        assert s1.npatches_t == c and s1.npatches_tagg == c and s1.npatches_x == 0
        assert f1.patch_sz == 8 and f1.search_sz_x == 10 and f1.search_sz_t == 5
    assert abs(port.default_params(20, O.FLT1).beta_x - 3.11) < 1e-6
    assert abs(port.default_params(20, O.SMO1).beta_t - 5.2) < 1e-6


def test_port_params_match_reference_library(port, ref):
    from oracle import oracle as O
    for mode in (O.FLT1, O.FLT2, O.SMO1):
        for s in (0.0, 3.0, 5.0, 10.0, 20.0, 25.5, 30.0, 40.0, 57.0, 80.0):
            assert ref.default_params(s, mode).as_dict() == port.default_params(s, mode).as_dict()
    # user-set fields are kept
    p = O.Params.auto(patch_sz=12, npatches_t=7, beta_t=0.5)
    q = O.Params.auto(patch_sz=12, npatches_t=7, beta_t=0.5)
    assert ref.default_params(20, O.FLT2, p).as_dict() == port.default_params(20, O.FLT2, q).as_dict()


def test_port_dct_is_orthonormal(port):
    from scipy.fft import dctn
    rng = np.random.default_rng(0)
    for psz in (4, 8, 12, 7):
        t = rng.normal(0, 50, (9, psz, psz)).astype(np.float32)
        y = port.dct2(t)
        want = dctn(t.astype(np.float64), type=2, norm="ortho", axes=(1, 2))
        assert np.abs(y - want).max() < 1e-4
        back = port.dct2(y, inverse=True)
        assert np.abs(back - t).max() < 1e-4


def test_port_window_and_colour(port):
    w = port.window(8)
    n2 = 3.5
    w1 = np.exp(-0.5 * (((np.arange(8) - n2) / n2 / 0.4) ** 2))
    assert np.abs(w - np.outer(w1, w1)).max() < 1e-6
    rng = np.random.default_rng(1)
    im = rng.uniform(0, 255, (17, 23, 3)).astype(np.float32)
    back = port.opp2rgb(port.rgb2opp(im.copy()))
    assert np.abs(back - im).max() < 1e-3
    gray = rng.uniform(0, 255, (5, 6, 1)).astype(np.float32)
    assert np.array_equal(port.rgb2opp(gray.copy()), gray)


def test_port_warp_against_numpy(port):
    """bicubic (Keys a=-1/2) with NaN outside the image and under the mask"""
    rng = np.random.default_rng(2)
    h, w = 20, 24
    im = rng.uniform(0, 255, (h, w, 1)).astype(np.float32)
    of = np.zeros((h, w, 2), np.float32)
    of[..., 0], of[..., 1] = 2.0, -1.0  # integer shift: exact copy where taps are inside
    msk = np.zeros((h, w), np.float32)
    msk[3, 4] = 255
    out = port.warp_bicubic(im, of, msk)
    assert np.isnan(out[3, 4, 0])
    for y in range(h):
        for x in range(w):
            if msk[y, x]:
                continue
            ix, iy = x + 2 - 1, y - 1 - 1
            inside = ix >= 0 and ix + 3 < w and iy >= 0 and iy + 3 < h
            if inside:
                assert out[y, x, 0] == im[y - 1, x + 2, 0]
            else:
                assert np.isnan(out[y, x, 0])


def test_mask_separability_against_reference(port, ref):
    """the restatement splits search / mask replay / filtering; the reference runs one
    loop.  Same output within fp32 noise on a frame large enough for thousands of skips."""
    from bwd_nlkalman_b200 import synth
    from oracle import oracle as O
    w, h, ch, sigma = 200, 152, 1, 20.0
    n0 = synth.noisy_frame(w, h, ch, 0, sigma)
    f1 = ref.default_params(sigma, O.FLT1)
    a = ref.filter_frame(n0, None, None, sigma, f1)
    b, dump = port.filter_frame(n0, None, None, sigma, f1, dump=True)
    assert maxabs(a, b) <= TOL_MAXABS
    frac = dump["active"].mean()
    assert 0.5 < frac < 0.95, frac  # the mask really skips patches here


def test_occlusion_oracle_hand_example():
    """the numpy restatement of the script's plambda expression (reference
    scripts/nlkalman-seq.sh:70-72) on a case small enough to do by hand"""
    from oracle import oracle as O
    of = np.zeros((2, 3, 2), np.float32)
    of[0, :, 0] = [1.0, 1.5, 4.0]      # u, row 0: du/dx = 0 (clamped), 0.5, 2.5
    of[1, :, 0] = [0.0, 0.0, 0.0]
    of[1, :, 1] = [0.0, -3.0, 0.25]    # v, row 1: dv/dy = 0, -3, 0.25 (row 0: clamped, 0)
    got = O.occlusion_from_flow(of, 0.75)
    want = np.array([[0, 0, 255], [0, 255, 0]], np.float32)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("ch", [1, 3])
def test_port_matches_reference_on_parameter_variants(port, ref, ch):
    """The parameter combinations of tests/test_gpu_parity.py::test_team_kernel_variants (group
    sizes, candidate counts, single statistics round, second filtering with several members,
    small smoother groups): the restatement the GPU is checked against must itself agree with
    the unmodified reference there."""
    from bwd_nlkalman_b200 import synth
    from oracle import oracle as O
    w, h, sigma = 93, 70, 20.0
    f1 = ref.default_params(sigma, O.FLT1)
    n0 = port.rgb2opp(synth.noisy_frame(w, h, ch, 0, sigma))
    n1 = port.rgb2opp(synth.noisy_frame(w, h, ch, 1, sigma))
    bflo, fflo = synth.backward_flow(w, h), synth.forward_flow(w, h)
    occ = np.zeros((h, w), np.float32)
    occ[20:34, 40:60] = 255
    c11 = ref.filter_frame(n0, None, None, sigma, f1)
    w1 = ref.warp_bicubic(c11, bflo, occ)
    for ov in (dict(npatches_t=8, npatches_tagg=20), dict(npatches_t=32), dict(npatches_t=33),
               dict(npatches_t=30, npatches_tagg=9)):
        p = ref.default_params(sigma, O.FLT1, O.Params.auto(**ov))
        assert maxabs(ref.filter_frame(n1, w1, None, sigma, p), port.filter_frame(n1, w1, None, sigma, p)) <= TOL_MAXABS, ov
    c12 = ref.filter_frame(n1, w1, None, sigma, f1)
    for ov in (dict(), dict(npatches_tagg=2), dict(npatches_tagg=5), dict(npatches_t=6), dict(npatches_t=6, npatches_tagg=3)):
        p = ref.default_params(sigma, O.FLT2, O.Params.auto(**ov))
        assert maxabs(ref.filter_frame(n1, w1, c12, sigma, p), port.filter_frame(n1, w1, c12, sigma, p)) <= TOL_MAXABS, ov
    ws = ref.warp_bicubic(c12, fflo, occ)
    for ov in (dict(npatches_t=12, npatches_tagg=2), dict(npatches_t=32, npatches_tagg=32), dict()):
        p = ref.default_params(sigma, O.SMO1, O.Params.auto(**ov))
        assert maxabs(ref.smooth_frame(c11, ws, None, sigma, p), port.smooth_frame(c11, ws, None, sigma, p)) <= TOL_MAXABS, ov


def test_port_smoother_with_basic_estimate_matches_reference(port, ref):
    """nlkalman_smooth_frame with bsic1 != NULL (reference src/nlkalman.c:1669, :1734): the inputs of
    tests/test_gpu_parity.py::test_smoother_with_basic_estimate, restatement vs the reference library"""
    from bwd_nlkalman_b200 import synth
    from oracle import oracle as O
    sigma = 20.0
    for (w, h, ch, ov) in ((93, 70, 3, {}), (88, 66, 1, {}), (90, 66, 3, dict(patch_sz=12, search_sz_t=6))):
        s1 = ref.default_params(sigma, O.SMO1, O.Params.auto(npatches_t=24, npatches_tagg=10, **ov))
        rng = np.random.default_rng(w)
        clean0 = port.rgb2opp(synth.clean_frame(w, h, ch, 0))
        flt = clean0 + rng.normal(0, 4, clean0.shape).astype(np.float32)
        bsic = clean0 + rng.normal(0, 2, clean0.shape).astype(np.float32)
        nxt = port.rgb2opp(synth.clean_frame(w, h, ch, 1)) + rng.normal(0, 2, clean0.shape).astype(np.float32)
        occ = np.zeros((h, w), np.float32)
        occ[20:34, 40:60] = 255
        ws = port.warp_bicubic(nxt, synth.forward_flow(w, h), occ)
        assert maxabs(ref.smooth_frame(flt, ws, bsic, sigma, s1), port.smooth_frame(flt, ws, bsic, sigma, s1)) <= TOL_MAXABS


def test_occlusion_oracle_pinned_to_reference_plambda(tmp_path):
    """the occlusion mask is the reference's plambda tool (lib/imscript-lite/src/plambda.c,
    compiled unmodified as oracle/_ref/plambda-ref) run with the exact expression of
    scripts/nlkalman-seq.sh:70-72; the numpy restatement the GPU kernel is checked against must
    reproduce it bit for bit (boundary = nearest sample, plambda.c:2176)."""
    import os
    import struct
    import subprocess
    from oracle import oracle as O
    exe = os.path.join(os.path.dirname(O.REF_SO), "plambda-ref")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/plambda-ref not built (needs /root/reference)")
    rng = np.random.default_rng(11)
    for (w, h) in ((157, 93), (33, 2), (1, 40)):
        of = rng.normal(0, 0.6, (h, w, 2)).astype(np.float32)
        of[h // 4:h // 2, w // 5:w // 2] += rng.normal(0, 2.0, (h // 2 - h // 4, w // 2 - w // 5, 2)).astype(np.float32)
        flo, out = tmp_path / "f.flo", tmp_path / "o.pfm"
        with open(flo, "wb") as f:
            f.write(b"PIEH" + struct.pack("<ii", w, h) + of.tobytes())
        for th in (0.25, 0.75, 2.0):
            expr = f"x(0,0)[0] x(-1,0)[0] - x(0,0)[1] x(0,-1)[1] - + fabs {th} > 255 *"
            subprocess.run([exe, str(flo), expr, "-o", str(out)], check=True)
            raw = open(out, "rb").read()
            head = raw.split(b"\n", 3)
            ww, hh = (int(x) for x in head[1].split())
            got = np.frombuffer(head[3], np.float32).reshape(hh, ww)
            assert np.array_equal(got, O.occlusion_from_flow(of, th)), (w, h, th)

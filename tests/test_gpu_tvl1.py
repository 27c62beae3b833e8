"""Dual TV-L1 optical flow on the GPU: one scale (lib/tvl1flow/tvl1flow_lib.c:93-280)
and the whole pyramid (:345-477) against the reference's own library (compiled unmodified as
oracle/_ref/libtvl1_ref.so) and against golden vectors it produced.

The kernels round like the reference (no fused multiply-add; tests/test_tvl1_model.py checks the same
device functions bit for bit on the CPU), so with the same iterations the flows are expected to be
IDENTICAL; what can differ is the float sum behind the stopping rule (:164, another order here and in
the reference's own OpenMP reduction): when it lands on the other side of epsilon^2 one side runs one
more iteration, which moves the flow by about epsilon = 0.01 px.  Tolerances: 1e-4 px when both
sides run all iterations (epsilon = 0), 5e-2 px with the stopping rule; the fraction of bit-identical
pixels is printed."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tvl1", "level_96x72.npz")


def _err(a, b):
    return float(np.abs(a.astype(np.float64) - b).max())


def test_tvl1_level_against_golden(nlk):
    g = dict(np.load(GOLD))
    tau, lam, theta = (float(x) for x in g["params"])
    z = np.zeros_like(g["I0"])
    with nlk.Context(96, 72, 1) as ctx:
        u1, u2, its = ctx.tvl1_level(g["I0"], g["I1"], z, z, tau, lam, theta, warps=2, epsilon=0.0)
        assert list(its) == [300, 300]
        e = max(_err(u1, g["full_u1"]), _err(u2, g["full_u2"]))
        print(f"tvl1 96x72, 2 x 300 iterations: max |du| = {e:.2e} px")
        assert e <= 1e-4
        v1, v2, its = ctx.tvl1_level(g["I0"], g["I1"], z, z, tau, lam, theta, warps=5, epsilon=0.01)
        e = max(_err(v1, g["dflt_u1"]), _err(v2, g["dflt_u2"]))
        print(f"tvl1 96x72, default stopping rule: iterations {list(its)}, max |du| = {e:.2e} px")
        assert e <= 5e-2 and all(1 <= n <= 300 for n in its)
    # it is the flow of the scene: I1(x) = I0(x - (1.5, -0.75))
    assert abs(np.median(v1[15:-15, 15:-15]) - 1.5) < 0.05 and abs(np.median(v2[15:-15, 15:-15]) + 0.75) < 0.05


@pytest.mark.parametrize("shape", [(321, 240), (77, 53)])
def test_tvl1_level_against_reference_library(nlk, shape):
    from oracle import oracle as O
    if not os.path.exists(O.TVL1_SO):
        pytest.skip("oracle/_ref/libtvl1_ref.so not built (needs /root/reference)")
    ref = O.Tvl1Ref()
    nx, ny = shape
    I0, I1 = O.tvl1_pair(nx, ny, shift=(2.25, 1.5), seed=nx)
    rng = np.random.default_rng(1)
    # a non-zero initial flow, as a level gets from the coarser one (tvl1flow_lib.c:423-431)
    u0 = (2.0 + rng.normal(0, 0.2, (ny, nx))).astype(np.float32)
    v0 = (1.2 + rng.normal(0, 0.2, (ny, nx))).astype(np.float32)
    with nlk.Context(nx, ny, 1) as ctx:
        a1, a2, its = ctx.tvl1_level(I0, I1, u0, v0, warps=3, epsilon=0.0)
        r1, r2 = ref.level(I0, I1, u0, v0, warps=3, epsilon=0.0)
        e = max(_err(a1, r1), _err(a2, r2))
        print(f"tvl1 {nx}x{ny}, 3 x 300 iterations: max |du| = {e:.2e} px")
        assert list(its) == [300, 300, 300] and e <= 1e-4
        b1, b2, its = ctx.tvl1_level(I0, I1, u0, v0, warps=5, epsilon=0.01)
        q1, q2 = ref.level(I0, I1, u0, v0, warps=5, epsilon=0.01)
        e = max(_err(b1, q1), _err(b2, q2))
        print(f"tvl1 {nx}x{ny}, default stopping rule: iterations {list(its)}, max |du| = {e:.2e} px")
        assert e <= 5e-2


def _same(a, b):
    return float(np.mean(a == b))


def test_tvl1_flow_against_golden(nlk):
    g = np.load(os.path.join(os.path.dirname(GOLD), "flow_160x120.npz"))
    I0, I1 = g["I0"], g["I1"]
    ny, nx = I0.shape
    assert nlk.tvl1_scales(nx, ny, 0.5, 100) == int(g["nscales"][0])
    with nlk.Context(nx, ny, 1) as ctx:
        for key, kw in (("flow_default", {}), ("flow_script", dict(lam=0.4, fscale=1))):
            flow, its = ctx.tvl1_flow(I0, I1, **kw)
            e = _err(flow, g[key])
            print(f"tvl1 pyramid {nx}x{ny} {key}: max |du| = {e:.2e} px, identical pixels {_same(flow, g[key]):.4f}, "
                  f"iterations per scale {its.sum(1).tolist()}")
            assert e <= 5e-2
            assert (its[kw.get("fscale", 0):] >= 1).all() and (its[:kw.get("fscale", 0)] == 0).all()
    from oracle import oracle as O
    dx, dy = O.tvl1_truth(nx, ny)     # it is the motion of the scene
    assert np.median(np.abs(flow[0] - dx)) < 0.1 and np.median(np.abs(flow[1] - dy)) < 0.1


@pytest.mark.parametrize("nx,ny,kw", [(320, 240, {}), (200, 150, dict(lam=0.4, fscale=1)), (131, 97, dict(zfactor=0.7)),
                                      (96, 72, dict(nscales=1)), (640, 360, dict(lam=0.25, fscale=1, epsilon=0.0, warps=2))])
def test_tvl1_flow_against_reference_library(nlk, nx, ny, kw):
    from oracle import oracle as O
    if not os.path.exists(O.TVL1_SO):
        pytest.skip("oracle/_ref/libtvl1_ref.so not built (needs /root/reference)")
    ref = O.Tvl1Ref()
    I0, I1 = O.tvl1_frames(nx, ny, seed=nx)
    want, nscales = ref.flow(I0, I1, **kw)
    with nlk.Context(nx, ny, 1) as ctx:
        flow, its = ctx.tvl1_flow(I0, I1, **kw)
    assert its.shape[0] == nscales
    e = _err(flow, want)
    print(f"tvl1 pyramid {nx}x{ny} {kw}: {nscales} scales, max |du| = {e:.2e} px, identical pixels {_same(flow, want):.4f}, "
          f"iterations per scale {its.sum(1).tolist()}")
    assert np.isfinite(flow).all() and np.abs(want).max() > 2.0
    assert e <= (1e-4 if kw.get("epsilon", 0.01) == 0.0 else 5e-2)


def test_tvl1_flow_on_device_buffers(nlk):
    import torch
    from oracle import oracle as O
    nx, ny = 256, 192
    I0, I1 = O.tvl1_frames(nx, ny, seed=9)
    with nlk.Context(nx, ny, 1) as ctx:
        host, _ = ctx.tvl1_flow(I0, I1, lam=0.4, fscale=1)
        d0, d1 = torch.from_numpy(I0).cuda(), torch.from_numpy(I1).cuda()
        u = torch.full((2, ny, nx), float("nan"), device="cuda")
        ctx.tvl1_flow_dev(d0, d1, u[0], u[1], nx, ny, lam=0.4, fscale=1)
        ctx.sync()
        e = _err(u.cpu().numpy(), host)
        print(f"tvl1 pyramid on device buffers vs host entry: max |du| = {e:.2e}")
        assert e <= 5e-2
        # requests the reference would abort on or index out of its pyramid for
        with pytest.raises(nlk.NlkError):
            ctx.tvl1_flow(I0[:4, :4], I1[:4, :4], nscales=1)       # Gaussian window larger than the image
        with pytest.raises(nlk.NlkError):
            nlk.api._check(nlk.lib().nlk_tvl1_flow_dev(ctx._h, nlk.api._vp(d0), nlk.api._vp(d1), nlk.api._vp(u[0]),
                                                       nlk.api._vp(u[1]), nx, ny, 0.25, 0.15, 0.3, 3, 0, 1.5, 5, 0.01, None))


def test_flow_mask_between_resident_frames(nlk):
    """nlk_flow_mask_dev = luminance of both frames (as the reference's program reads a colour float file),
    the estimator with the script's parameters, the flow interleaved, the plambda mask"""
    import torch
    from oracle import oracle as O
    nx, ny = 224, 160
    g0, g1 = O.tvl1_frames(nx, ny, seed=13)
    rgb = lambda g: np.stack([g, 0.6 * g + 25, 180 - 0.4 * g], -1).astype(np.float32)
    a, b = rgb(g0), rgb(g1)
    lum = lambda x: (lambda d: (.299 * d[..., 0] + .587 * d[..., 1] + .114 * d[..., 2]).astype(np.float32))(x.astype(np.float64))
    with nlk.Context(nx, ny, 3) as ctx:
        want, _ = ctx.tvl1_flow(lum(a), lum(b), lam=0.25, fscale=1)
        of = torch.empty((ny, nx, 2), device="cuda")
        occ = torch.empty((ny, nx), device="cuda")
        ctx.flow_mask_dev(of, occ, torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), nlk.Tvl1Params.script(0.25, 1), 0.02)
        ctx.sync()
        got = of.cpu().numpy()
        assert np.array_equal(got[..., 0], want[0]) and np.array_equal(got[..., 1], want[1])
        m = occ.cpu().numpy()
        assert np.array_equal(m, ctx.occlusion(got, 0.02)) and 0 < (m > 0).mean() < 1
    with nlk.Context(nx, ny, 1) as ctx:    # single-channel frames go in as they are
        of = torch.empty((ny, nx, 2), device="cuda")
        ctx.flow_mask_dev(of, None, torch.from_numpy(lum(a)).cuda(), torch.from_numpy(lum(b)).cuda(), nlk.Tvl1Params.script(0.25, 1), 0.75)
        ctx.sync()
        assert np.array_equal(of.cpu().numpy()[..., 0], want[0])


def test_tvl1_three_loop_forms_agree(nlk, monkeypatch):
    """how the iterations of a warping step are driven (NLK_TVL1_LOOP): `kernel` (default) = one cooperative
    launch per warping step, grid barriers between the half iterations; `graph` = one CUDA graph per level,
    WHILE nodes around the two iteration kernels; `host` = the host queues batches and reads the error back.
    Same per-pixel functions, same stopping iteration: identical flows and iteration counts -- and the
    launch counts show which form ran (no silent fallback)."""
    from oracle import oracle as O
    nx, ny = 320, 240
    I0, I1 = O.tvl1_frames(nx, ny, seed=4)
    out = {}
    for mode in ("kernel", "graph", "host"):
        monkeypatch.setenv("NLK_TVL1_LOOP", mode)
        with nlk.Context(nx, ny, 1) as ctx:
            l0 = ctx.launches
            flow, its = ctx.tvl1_flow(I0, I1, lam=0.25)
            flow2, its2 = ctx.tvl1_flow(I0, I1, lam=0.25)      # (graph: the cached graphs, launched again)
            out[mode] = (flow, its, ctx.launches - l0)
            assert np.array_equal(flow, flow2) and np.array_equal(its, its2)
    for mode in ("graph", "host"):
        assert np.array_equal(out["kernel"][0], out[mode][0]) and np.array_equal(out["kernel"][1], out[mode][1]), mode
    nscales = out["kernel"][1].shape[0]
    print(f"launches for two flows: kernel {out['kernel'][2]}, graph {out['graph'][2]}, host {out['host'][2]}; "
          f"iterations {out['kernel'][1].sum(1).tolist()}")
    assert out["graph"][2] < out["kernel"][2] < out["host"][2] / 4
    assert out["kernel"][2] >= 2 * nscales * (1 + 2 * 5)     # gradient + (warp, iterate) x 5 per scale


def test_reference_entry_points_of_the_flow_library(nlk):
    """include/tvl1flow.h: Dual_TVL1_optic_flow and Dual_TVL1_optic_flow_multiscale with the reference's
    names and argument lists, exported by the product library -- bound here through the same ctypes
    wrapper as the reference's library and compared with it"""
    from oracle import oracle as O
    if not os.path.exists(O.TVL1_SO):
        pytest.skip("oracle/_ref/libtvl1_ref.so not built (needs /root/reference)")
    so = os.path.join(os.path.dirname(os.path.abspath(nlk.__file__)), "libnlkalman_b200.so")
    ours, ref = O.Tvl1Ref(so=so), O.Tvl1Ref()
    nx, ny = 200, 150
    I0, I1 = O.tvl1_frames(nx, ny, seed=8)
    a, ns = ours.flow(I0, I1, lam=0.4, fscale=1)
    b, _ = ref.flow(I0, I1, lam=0.4, fscale=1)
    print(f"Dual_TVL1_optic_flow_multiscale {nx}x{ny}: max |du| = {_err(a, b):.2e}, identical {_same(a, b):.4f}")
    assert _err(a, b) <= 5e-2
    J0, J1 = O.tvl1_pair(96, 72)
    z = np.zeros((72, 96), np.float32)
    p, q = ours.level(J0, J1, z, z), ref.level(J0, J1, z, z)
    assert max(_err(p[0], q[0]), _err(p[1], q[1])) <= 5e-2

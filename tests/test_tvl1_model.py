"""The TV-L1 device arithmetic, checked on the CPU: tests/models/tvl1_host_model.cpp compiles the
per-pixel device functions of csrc/nlk_tvl1.cuh for the host and runs them through the library's own
pyramid sequence (csrc/nlk_tvl1_pyramid.h); here it is compared BIT FOR BIT with the reference's
library (oracle/_ref/libtvl1_ref.so: lib/tvl1flow compiled unmodified)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
MODEL_SRC = os.path.join(HERE, "models", "tvl1_host_model.cpp")
MODEL_SO = os.path.join(HERE, "models", "libtvl1_model.so")
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)


def _p(a):
    return a.ctypes.data_as(_fp)


@pytest.fixture(scope="module")
def model():
    deps = [MODEL_SRC] + [os.path.join(HERE, "..", "bwd_nlkalman_b200", "csrc", f) for f in ("nlk_tvl1.cuh", "nlk_tvl1_pyramid.h")]
    if not os.path.exists(MODEL_SO) or any(os.path.getmtime(d) > os.path.getmtime(MODEL_SO) for d in deps):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", MODEL_SO, MODEL_SRC], check=True)
    L = C.CDLL(MODEL_SO)
    L.model_tvl1_level.argtypes = [_fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int,
                                   C.c_float, _ip]
    L.model_tvl1_gaussian.argtypes = [_fp, C.c_int, C.c_int, C.c_double]
    L.model_tvl1_zoom.argtypes = [_fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int]
    L.model_tvl1_flow.argtypes = [_fp, _fp, _fp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int,
                                  C.c_float, C.c_int, C.c_float, _ip]
    L.model_tvl1_scales.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int]
    return L


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(O.TVL1_SO):
        pytest.skip("oracle/_ref/libtvl1_ref.so not built (needs /root/reference)")
    return O.Tvl1Ref(threads=1)     # one thread: the error sum of the stopping test in pixel order


@pytest.mark.parametrize("nx,ny,sigma", [(96, 72, 0.8), (61, 47, 0.6 * np.sqrt(3.0)), (33, 40, 2.3)])
def test_gaussian_bit_exact(model, ref, nx, ny, sigma):
    I = np.random.default_rng(1).uniform(0, 255, (ny, nx)).astype(np.float32)
    sigma = float(np.float32(sigma))
    want = ref.gaussian(I, sigma)
    got = I.copy()
    assert model.model_tvl1_gaussian(_p(got), nx, ny, sigma) == 0
    assert np.array_equal(got, want)


@pytest.mark.parametrize("nx,ny,factor", [(96, 72, 0.5), (61, 47, 0.5), (80, 50, 0.7)])
def test_zoom_out_and_in_bit_exact(model, ref, nx, ny, factor):
    I = np.random.default_rng(2).uniform(0, 255, (ny, nx)).astype(np.float32)
    want = ref.zoom_out(I, factor)
    nyy, nxx = want.shape
    f = np.float32(factor)
    zsigma = float(np.float32(0.6 * np.sqrt(1.0 / float(f * f) - 1.0)))
    sm = I.copy()
    assert model.model_tvl1_gaussian(_p(sm), nx, ny, zsigma) == 0
    got = np.zeros_like(want)
    model.model_tvl1_zoom(_p(sm), _p(got), nx, ny, nxx, nyy, factor, factor, 1.0, 0)
    assert np.array_equal(got, want)
    # and back up (the flow's path: zoom_in, then times 1 / zfactor)
    up_want = ref.zoom_in(want, nx, ny) * (np.float32(1.0) / f)
    up = np.zeros((ny, nx), np.float32)
    model.model_tvl1_zoom(_p(want), _p(up), nxx, nyy, nx, ny, np.float32(nx) / np.float32(nxx), np.float32(ny) / np.float32(nyy),
                          float(np.float32(1.0) / f), 1)
    assert np.array_equal(up, up_want.astype(np.float32))


def test_level_bit_exact(model, ref):
    nx, ny = 96, 72
    I0, I1 = O.tvl1_pair(nx, ny)
    z = np.zeros((ny, nx), np.float32)
    want = ref.level(I0, I1, z, z)
    a, b = z.copy(), z.copy()
    its = np.zeros(5, np.int32)
    model.model_tvl1_level(_p(I0), _p(I1), _p(a), _p(b), nx, ny, 0.25, 0.15, 0.3, 5, 0.01, its.ctypes.data_as(_ip))
    assert its.min() >= 1
    assert np.array_equal(a, want[0]) and np.array_equal(b, want[1])


@pytest.mark.parametrize("nx,ny,kw", [(160, 120, {}), (131, 97, {"fscale": 1}), (120, 90, {"zfactor": 0.7, "lam": 0.4}),
                                      (96, 72, {"nscales": 1})])
def test_flow_bit_exact(model, ref, nx, ny, kw):
    I0, I1 = O.tvl1_frames(nx, ny)
    p = dict(tau=0.25, lam=0.15, theta=0.3, nscales=100, fscale=0, zfactor=0.5, warps=5, epsilon=0.01)
    p.update(kw)
    want, nscales = ref.flow(I0, I1, **p)
    assert model.model_tvl1_scales(nx, ny, p["zfactor"], p["nscales"]) == nscales
    got = np.zeros((2, ny, nx), np.float32)
    its = np.zeros((nscales, 5), np.int32)
    assert model.model_tvl1_flow(_p(I0), _p(I1), _p(got), nx, ny, p["tau"], p["lam"], p["theta"], nscales,
                                 min(p["fscale"], nscales), p["zfactor"], p["warps"], p["epsilon"],
                                 its.ctypes.data_as(_ip)) == 0
    assert np.isfinite(got).all()
    if p["nscales"] > 1 and p["zfactor"] == 0.5:     # it is the motion of the scene
        dx, dy = O.tvl1_truth(nx, ny)
        assert np.median(np.abs(want[0] - dx)) < 0.1 and np.median(np.abs(want[1] - dy)) < 0.1
    assert np.array_equal(got, want), float(np.abs(got - want).max())
    assert (its[min(p["fscale"], nscales - 1):] >= 1).all() and (its[:p["fscale"]] == 0).all()


def test_flow_against_golden(model):
    """the committed vectors of the reference's pyramid (tests/golden/make_golden_tvl1.py), for boxes
    without oracle/_ref: default parameters and the pipeline script's (lambda 0.4, fscale 1)"""
    g = np.load(os.path.join(HERE, "golden", "tvl1", "flow_160x120.npz"))
    I0, I1 = g["I0"], g["I1"]
    ny, nx = I0.shape
    nscales = int(g["nscales"][0])
    assert model.model_tvl1_scales(nx, ny, 0.5, 100) == nscales
    for key, lam, fscale in (("flow_default", 0.15, 0), ("flow_script", 0.4, 1)):
        got = np.zeros((2, ny, nx), np.float32)
        assert model.model_tvl1_flow(_p(I0), _p(I1), _p(got), nx, ny, 0.25, lam, 0.3, nscales, fscale, 0.5, 5, 0.01, None) == 0
        assert np.array_equal(got, g[key]), key


def test_flow_bit_exact_on_random_requests(model, ref):
    """sizes, zoom factors, weights, warpings, tolerances, scale counts and first scales drawn at random; smooth
    scenes, noise, and constant images (the normalisation's max = min branch): every flow identical to
    the reference's"""
    rng = np.random.default_rng(0)
    checked = 0
    for trial in range(36):
        nx, ny = int(rng.integers(17, 140)), int(rng.integers(17, 110))
        zf = float(rng.choice([0.3, 0.5, 0.6, 0.75, 0.9]))
        p = dict(tau=float(rng.choice([0.1, 0.25])), lam=float(rng.choice([0.05, 0.15, 0.4, 1.0])),
                 theta=float(rng.choice([0.1, 0.3, 0.6])), nscales=int(rng.choice([1, 2, 3, 100])), fscale=0, zfactor=zf,
                 warps=int(rng.integers(1, 6)), epsilon=float(rng.choice([0.0005, 0.01, 0.05])))
        if trial % 3 == 0:
            I0, I1 = O.tvl1_frames(nx, ny, seed=trial)
        elif trial % 3 == 1:
            I0 = rng.uniform(0, 255, (ny, nx)).astype(np.float32)
            I1 = np.roll(I0, 1, 1) + rng.normal(0, 3, (ny, nx)).astype(np.float32)
        else:
            I0 = np.full((ny, nx), 7.0, np.float32)
            I1 = I0.copy()
            I1[ny // 2, nx // 2] += trial % 2
        ns = model.model_tvl1_scales(nx, ny, zf, p["nscales"])
        if ns < 1:
            continue
        p["fscale"] = int(rng.integers(0, ns + 1))
        got = np.zeros((2, ny, nx), np.float32)
        rc = model.model_tvl1_flow(_p(I0), _p(I1), _p(got), nx, ny, p["tau"], p["lam"], p["theta"], ns, p["fscale"], zf,
                                   p["warps"], p["epsilon"], None)
        if rc != 0:
            continue        # a request the library refuses (the reference aborts or reads out of bounds on it)
        want, ns_ref = ref.flow(np.ascontiguousarray(I0), np.ascontiguousarray(I1), **p)
        assert ns_ref == ns and np.array_equal(got, want, equal_nan=True), (trial, nx, ny, p)
        checked += 1
    assert checked >= 30

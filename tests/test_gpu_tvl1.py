"""Dual TV-L1 optical flow at one scale on the GPU (first slice) against the
reference's own library (lib/tvl1flow/tvl1flow_lib.c:93-280 compiled unmodified as
oracle/_ref/libtvl1_ref.so) and against golden vectors it produced.

Floating point, iterative: the tolerance is 2e-3 px on the flow when both sides run the same
iterations (epsilon = 0: all 300 of every warping step), and with the data-dependent stopping rule
(:164) the iteration counts may differ by one near the threshold, where one update moves the flow
by about epsilon = 0.01 px."""
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
        assert e <= 2e-3
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
        assert list(its) == [300, 300, 300] and e <= 2e-3
        b1, b2, its = ctx.tvl1_level(I0, I1, u0, v0, warps=5, epsilon=0.01)
        q1, q2 = ref.level(I0, I1, u0, v0, warps=5, epsilon=0.01)
        e = max(_err(b1, q1), _err(b2, q2))
        print(f"tvl1 {nx}x{ny}, default stopping rule: iterations {list(its)}, max |du| = {e:.2e} px")
        assert e <= 5e-2

"""NumPy model of the packed-fp32 8x8 transform of csrc/nlk_dct.cuh: the tile as 32 pairs
(t[y][x], t[7-y][x]); rows by the even/odd/even 8-point factorisation applied to both halves
of a pair; columns from (s, d) = (lo + hi, lo - hi) against the constant pairs
(T[2j][y], T[2j+1][y]), which leaves the pairs (coef[2j][k], coef[2j+1][k]) -- no data movement
between the passes.  Checked against the orthonormal DCT-II of scipy (what the reference gets
from FFTW after its rescaling, src/nlkalman.c:276-299), forward, inverse and the fused
forward / gain / inverse of the update."""
import numpy as np
from scipy.fft import dctn, idctn

N = 8
T = np.array([[np.sqrt((2.0 if k else 1.0) / N) * np.cos(np.pi * (j + 0.5) * k / N) for j in range(N)]
              for k in range(N)])
TV = np.array([[(T[2 * j][y], T[2 * j + 1][y]) for y in range(4)] for j in range(4)])   # [j][y] -> pair


def fwd8(x):
    """dct8_fwd_x2 on a vector of 8 pairs: x[i] is (lo, hi)"""
    s = [x[i] + x[7 - i] for i in range(4)]
    d = [x[i] - x[7 - i] for i in range(4)]
    ss0, ss1, sd0, sd1 = s[0] + s[3], s[1] + s[2], s[0] - s[3], s[1] - s[2]
    out = [None] * 8
    out[0] = T[0][1] * ss1 + T[0][0] * ss0
    out[4] = T[4][1] * ss1 + T[4][0] * ss0
    out[2] = T[2][1] * sd1 + T[2][0] * sd0
    out[6] = T[6][1] * sd1 + T[6][0] * sd0
    for k in (1, 3, 5, 7):
        out[k] = T[k][3] * d[3] + T[k][2] * d[2] + T[k][1] * d[1] + T[k][0] * d[0]
    return out


def inv8(X):
    p0 = T[4][0] * X[4] + T[0][0] * X[0]
    p1 = T[4][1] * X[4] + T[0][1] * X[0]
    q0 = T[6][0] * X[6] + T[2][0] * X[2]
    q1 = T[6][1] * X[6] + T[2][1] * X[2]
    e = [p0 + q0, p1 + q1, p1 - q1, p0 - q0]
    o = [T[7][j] * X[7] + T[5][j] * X[5] + T[3][j] * X[3] + T[1][j] * X[1] for j in range(4)]
    return [e[0] + o[0], e[1] + o[1], e[2] + o[2], e[3] + o[3], e[3] - o[3], e[2] - o[2], e[1] - o[1], e[0] - o[0]]


def pack(t):
    return [[np.array([t[y][x], t[7 - y][x]]) for x in range(8)] for y in range(4)]


def cols_fwd(P):
    C = [[None] * 8 for _ in range(4)]
    for k in range(8):
        sd = [np.array([P[y][k][0] + P[y][k][1], P[y][k][0] - P[y][k][1]]) for y in range(4)]
        for j in range(4):
            C[j][k] = sum(TV[j][y] * sd[y] for y in range(4))
    return C


def cols_inv(C):
    P = [[None] * 8 for _ in range(4)]
    for k in range(8):
        for y in range(4):
            e, o = sum(TV[j][y] * C[j][k] for j in range(4))
            P[y][k] = np.array([e + o, e - o])
    return P


def unpack_coef(C):
    out = np.zeros((8, 8))
    for j in range(4):
        for k in range(8):
            out[2 * j][k], out[2 * j + 1][k] = C[j][k]
    return out


def unpack_pix(P):
    out = np.zeros((8, 8))
    for y in range(4):
        for x in range(8):
            out[y][x], out[7 - y][x] = P[y][x]
    return out


def test_forward_inverse_and_shrink_match_the_orthonormal_dct():
    rng = np.random.default_rng(3)
    for _ in range(10):
        t = rng.uniform(0, 255, (8, 8))
        P = pack(t)
        P = [fwd8(P[y]) for y in range(4)]                 # rows
        C = cols_fwd(P)                                    # columns
        want = dctn(t, type=2, norm="ortho")
        assert np.abs(unpack_coef(C) - want).max() < 1e-9
        # inverse
        back = cols_inv(C)
        back = [inv8(back[y]) for y in range(4)]
        assert np.abs(unpack_pix(back) - t).max() < 1e-9
        # the update: T^-1(a * T(t) + b), gains paired like the coefficients
        a, b = rng.uniform(0, 1, (8, 8)), rng.uniform(-5, 5, (8, 8))
        S = [[np.array([a[2 * j][k], a[2 * j + 1][k]]) * C[j][k] + np.array([b[2 * j][k], b[2 * j + 1][k]])
              for k in range(8)] for j in range(4)]
        upd = cols_inv(S)
        upd = [inv8(upd[y]) for y in range(4)]
        assert np.abs(unpack_pix(upd) - idctn(a * want + b, type=2, norm="ortho")).max() < 1e-9

"""Shared helpers of the parity tests."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PARAM_FIELDS = ["patch_sz", "search_sz_x", "search_sz_t", "npatches_x", "npatches_t",
                "npatches_tagg", "dista_lambda", "beta_x", "beta_t"]

# north_star tolerance: per-pixel max abs error <= 1e-3 on the 0-255 scale, |dPSNR| <= 0.01 dB
TOL_MAXABS = 1e-3
TOL_DPSNR = 0.01


def golden_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def params_from_array(cls, arr):
    p = cls()
    for f, v in zip(PARAM_FIELDS, arr):
        setattr(p, f, int(v) if f in PARAM_FIELDS[:6] else float(v))
    return p


def maxabs(a, b):
    """max |a-b| with NaN positions required to coincide"""
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), "NaN patterns differ"
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    d[na] = 0
    return float(d.max()) if d.size else 0.0


def psnr_between(a, ref_clean):
    mse = float(np.mean((a.astype(np.float64) - ref_clean.astype(np.float64)) ** 2))
    return 10.0 * np.log10(255.0 ** 2 / max(mse, 1e-30))


def compare_knn(gpu, cpu):
    """compare search dumps; returns (#groups, #groups whose kept index lists differ,
    list of (g, first differing rank, gap) for the first mismatches)"""
    assert np.array_equal(gpu["nk"], cpu["nk"]), "nk differs"
    G, kmax = gpu["knn_xy"].shape[:2]
    diff = np.any(gpu["knn_xy"] != cpu["knn_xy"], axis=(1, 2))
    bad = np.nonzero(diff)[0]
    details = []
    for g in bad[:20]:
        k = int(cpu["nk"][g])
        r = int(np.nonzero(np.any(gpu["knn_xy"][g, :k] != cpu["knn_xy"][g, :k], axis=1))[0][0])
        gap = float(abs(cpu["knn_d"][g, min(r + 1, k - 1)] - cpu["knn_d"][g, r]))
        details.append((int(g), r, gap))
    return G, len(bad), details

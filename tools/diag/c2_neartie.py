"""Diagnostic (GPU box): first filtering of frame 1 at 1920x1080x3 -- GPU vs our strict-IEEE
restatement (oracle/nlk_port.c) vs the unmodified reference (-ffast-math), on identical inputs.
Where the reference and the restatement differ by more than 1e-3, the candidate distances of the
groups involved are recomputed in float64 and the gap at the cut (rank k-1 / k, and at the group
boundary) is printed: the "documented near-ties" of north_star."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bwd_nlkalman_b200 as nlk
from bwd_nlkalman_b200 import synth
from oracle import oracle as O


def main():
    w, h, ch, sigma = 1920, 1080, 3, 20.0
    if len(sys.argv) > 2:
        w, h = int(sys.argv[1]), int(sys.argv[2])
    ref, port = O.Ref(threads=1), O.Port()
    f1 = nlk.default_params(sigma, nlk.FLT1)
    pf1 = O.Params(*[getattr(f1, f) for f, _ in nlk.Params._fields_])
    n0 = ref.rgb2opp(synth.noisy_frame(w, h, ch, 0, sigma))
    n1 = ref.rgb2opp(synth.noisy_frame(w, h, ch, 1, sigma))
    bflo, occ = synth.backward_flow(w, h), synth.occlusion_mask(w, h)
    with nlk.Context(w, h, ch) as ctx:
        a0, _ = ctx.pass_host_debug(0, n0, None, None, sigma, f1)
        w1 = ref.warp_bicubic(a0, bflo, occ)
        t = time.time()
        g, gd = ctx.pass_host_debug(0, n1, w1, None, sigma, f1)
    t = time.time(); p, pd = port.run_pass(O.PASS_FILTER, n1, w1, None, sigma, pf1, dump=True); tp = time.time() - t
    t = time.time(); r = ref.filter_frame(n1, w1, None, sigma, pf1); tr = time.time() - t
    print(f"port {tp:.1f}s ref(1 thread) {tr:.1f}s")
    def cmp(a, b, name):
        d = np.abs(a.astype(np.float64) - b); d[np.isnan(d)] = 0
        print(f"{name}: max-abs {d.max():.3e}, pixels > 1e-3: {(d.max(axis=2) > 1e-3).sum()}")
        return d.max(axis=2)
    cmp(g, p, "GPU vs port")
    dpr = cmp(p, r, "port vs ref")
    cmp(g, r, "GPU vs ref")
    print("knn identical GPU/port:", np.array_equal(gd["knn_xy"], pd["knn_xy"]), "dist bit-identical:",
          np.array_equal(gd["knn_d"], pd["knn_d"]), "active identical:", np.array_equal(gd["active"], pd["active"]))
    ys, xs = np.nonzero(dpr > 1e-3)
    if len(ys) == 0:
        return
    print("bbox of port/ref differences: x", xs.min(), xs.max(), "y", ys.min(), ys.max())
    # groups whose reference patch lies within r+psz of a differing pixel: float64 distances, gap at the cuts
    psz, step, gw = 8, 4, pd["gw"]
    k, tagg = f1.npatches_t, f1.npatches_tagg
    src = n1.astype(np.float64)
    seen = 0
    cand_groups = set()
    for y, x in zip(ys[::16], xs[::16]):
        for gy in range(max(0, (y - 12) // step), min(pd["gh"], (y + 6) // step + 1)):
            for gx in range(max(0, (x - 12) // step), min(gw, (x + 6) // step + 1)):
                cand_groups.add(gy * gw + gx)
    rows = []
    for gidx in sorted(cand_groups):
        if not pd["active"][gidx]:
            continue
        nk = int(pd["nk"][gidx])
        if nk < 2:
            continue
        d = pd["knn_d"][gidx, :nk].astype(np.float64)
        # all candidate distances in float64
        gy, gx = divmod(gidx, gw)
        px, py = gx * step, gy * step
        rad = f1.search_sz_t if pd["prev_p"][gidx] else f1.search_sz_x
        x0, x1 = max(px - rad, 0), min(px + rad, w - psz)
        y0, y1 = max(py - rad, 0), min(py + rad, h - psz)
        refp = src[py:py + psz, px:px + psz]
        allv = []
        for qy in range(y0, y1 + 1):
            for qx in range(x0, x1 + 1):
                allv.append(((src[qy:qy + psz, qx:qx + psz] - refp) ** 2).mean())
        allv = np.sort(np.array(allv))
        kk = min(nk, len(allv) - 1)
        gap_cut = (allv[kk] - allv[kk - 1]) / max(allv[kk], 1e-30)
        gaps_in = np.diff(allv[:kk + 1]) / np.maximum(allv[1:kk + 1], 1e-30)
        rows.append((gap_cut, float(gaps_in.min()), gidx, px, py, nk))
    rows.sort()
    print("groups near the differences, smallest relative gap at the cut (k-th / (k+1)-th distance) and inside the list:")
    for gap_cut, gmin, gidx, px, py, nk in rows[:12]:
        print(f"  g {gidx} p=({px},{py}) k={nk}: gap at cut {gap_cut:.2e}, smallest gap inside {gmin:.2e}")


if __name__ == "__main__":
    main()

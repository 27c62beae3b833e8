"""Generates tests/golden/*.npz by running the UNMODIFIED reference numerics
(oracle/_ref/libnlkalman_ref.so built by oracle/Makefile from /root/reference,
OMP_NUM_THREADS=1) on small seeded synthetic inputs.  Run in the build container:

    make -C oracle ref && python tests/golden/make_golden.py

Each file holds the inputs and every intermediate frame of a two-frame run of the
reference's per-frame sequence (src/main-flt.c:340-380, src/main-smo.c:198-213):
frame 0 spatial flt1+flt2, frame 1 temporal flt1+flt2, then the smoother on frame 0.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bwd_nlkalman_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = {
    # name: (w, h, ch, sigma, overrides for (flt1, flt2, smo1))
    "gray_96x72_s20": (96, 72, 1, 20.0, {}, {}, {}),
    "rgb_80x64_s20": (80, 64, 3, 20.0, {}, {}, {}),
    "rgb_90x66_s40_p12": (90, 66, 3, 40.0,
                          dict(patch_sz=12, search_sz_t=10, search_sz_x=15),
                          dict(patch_sz=12, search_sz_t=10, search_sz_x=15),
                          dict(patch_sz=12, search_sz_t=10)),
}


def run_case(ref, w, h, ch, sigma, o1, o2, o3):
    f1 = ref.default_params(sigma, O.FLT1, O.Params.auto(**o1))
    f2 = ref.default_params(sigma, O.FLT2, O.Params.auto(**o2))
    s1 = ref.default_params(sigma, O.SMO1, O.Params.auto(**o3))
    n0 = synth.noisy_frame(w, h, ch, 0, sigma)
    n1 = synth.noisy_frame(w, h, ch, 1, sigma)
    bflo, fflo, occ = synth.backward_flow(w, h), synth.forward_flow(w, h), synth.occlusion_mask(w, h)
    # a small occluded rectangle that fits these tiny frames
    occ[:] = 0
    occ[h // 3:h // 3 + 10, w // 2:w // 2 + 14] = 255.0
    out = dict(noisy0=n0, noisy1=n1, bflo=bflo, fflo=fflo, occ=occ, sigma=np.float32(sigma),
               f1=np.array(list(f1.as_dict().values()), np.float64),
               f2=np.array(list(f2.as_dict().values()), np.float64),
               s1=np.array(list(s1.as_dict().values()), np.float64))
    o0 = ref.rgb2opp(n0.copy())
    o1_ = ref.rgb2opp(n1.copy())
    out["opp0"] = o0
    out["flt1_0"] = ref.filter_frame(o0, None, None, sigma, f1)
    out["flt2_0"] = ref.filter_frame(o0, None, out["flt1_0"], sigma, f2)
    out["warp1"] = ref.warp_bicubic(out["flt1_0"], bflo, occ)
    out["warp2"] = ref.warp_bicubic(out["flt2_0"], bflo, occ)
    out["flt1_1"] = ref.filter_frame(o1_, out["warp1"], None, sigma, f1)
    out["flt2_1"] = ref.filter_frame(o1_, out["warp2"], out["flt1_1"], sigma, f2)
    out["warps"] = ref.warp_bicubic(out["flt2_1"], fflo, occ)
    out["smo_0"] = ref.smooth_frame(out["flt2_0"], out["warps"], None, sigma, s1)
    out["rgb_flt2_1"] = ref.opp2rgb(out["flt2_1"].copy())
    return out


def main():
    ref = O.Ref(threads=1)
    here = os.path.dirname(os.path.abspath(__file__))
    for name, (w, h, ch, sigma, o1, o2, o3) in CASES.items():
        res = run_case(ref, w, h, ch, sigma, o1, o2, o3)
        path = os.path.join(here, name + ".npz")
        np.savez_compressed(path, **res)
        print(name, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()

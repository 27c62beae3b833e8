"""TEST INFRASTRUCTURE ONLY: ctypes access to the two CPU checkers.

* ``Ref``  -- oracle/_ref/libnlkalman_ref.so, the UNMODIFIED reference numerics
  (reference src/nlkalman.c) compiled by oracle/Makefile with an FFTW stand-in.
* ``Port`` -- oracle/libnlk_port.so, our plain-C restatement (oracle/nlk_port.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product (bwd_nlkalman_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libnlkalman_ref.so")
PORT_SO = os.path.join(HERE, "libnlk_port.so")

FLT1, FLT2, SMO1 = 0, 1, 2
PASS_FILTER, PASS_SMOOTH = 0, 1


class Params(C.Structure):
    """struct nlkalman_params, reference src/nlkalman.h:22-37 (K_SIMILAR_PATCHES on)."""
    _fields_ = [("patch_sz", C.c_int), ("search_sz_x", C.c_int), ("search_sz_t", C.c_int),
                ("npatches_x", C.c_int), ("npatches_t", C.c_int), ("npatches_tagg", C.c_int),
                ("dista_lambda", C.c_float), ("beta_x", C.c_float), ("beta_t", C.c_float)]

    @classmethod
    def auto(cls, **kw):
        p = cls(-1, -1, -1, -1, -1, -1, -1.0, -1.0, -1.0)
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


_fp = C.POINTER(C.c_float)


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_fp)


def build(ref: bool = True):
    """make -C oracle (port always; _ref only when /root/reference is present)."""
    subprocess.run(["make", "-C", HERE, "port"] + (["ref"] if ref else []), check=True,
                   stdout=subprocess.DEVNULL)


TVL1_SO = os.path.join(HERE, "_ref", "libtvl1_ref.so")


class Tvl1Ref:
    """The reference's TV-L1 flow library (lib/tvl1flow/tvl1flow_lib.c compiled unmodified)."""

    def __init__(self, threads: int | None = None, so: str | None = None):
        """so: another library exporting the same two entry points (the tests bind the PRODUCT's
        include/tvl1flow.h symbols through this very wrapper to compare like with like)"""
        so = so or TVL1_SO
        if not os.path.exists(so):
            raise FileNotFoundError(f"{so} missing: run `make -C oracle ref` where /root/reference exists")
        self.lib = C.CDLL(so)
        f = self.lib.Dual_TVL1_optic_flow          # tvl1flow_lib.c:93
        f.argtypes = [_fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_bool]
        f.restype = None
        if threads is not None:
            try:
                C.CDLL("libgomp.so.1", mode=C.RTLD_GLOBAL).omp_set_num_threads(int(threads))
            except OSError:
                pass

    def level(self, I0, I1, u1, u2, tau=0.25, lam=0.15, theta=0.3, warps=5, epsilon=0.01):
        ny, nx = I0.shape
        a = np.ascontiguousarray(u1, np.float32).copy()
        b = np.ascontiguousarray(u2, np.float32).copy()
        self.lib.Dual_TVL1_optic_flow(_p(np.ascontiguousarray(I0, np.float32)), _p(np.ascontiguousarray(I1, np.float32)),
                                      _p(a), _p(b), nx, ny, tau, lam, theta, warps, epsilon, False)
        return a, b

    def flow(self, I0, I1, tau=0.25, lam=0.15, theta=0.3, nscales=100, fscale=0, zfactor=0.5, warps=5, epsilon=0.01):
        """Dual_TVL1_optic_flow_multiscale (tvl1flow_lib.c:345) with the scale cap of the reference's driver
        (lib/tvl1flow/main.c:159-163); returns the flow as (2, ny, nx): u then v"""
        ny, nx = I0.shape
        N = np.float32(1 + np.log(np.hypot(nx, ny) / 16.0) / np.log(float(np.float32(1) / np.float32(zfactor))))
        if N < nscales:
            nscales = int(N)
        fscale = min(fscale, nscales)
        f = self.lib.Dual_TVL1_optic_flow_multiscale
        f.argtypes = [_fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_float,
                      C.c_int, C.c_float, C.c_bool]
        f.restype = None
        out = np.zeros((2, ny, nx), np.float32)
        f(_p(np.ascontiguousarray(I0, np.float32)), _p(np.ascontiguousarray(I1, np.float32)), _p(out[0]), _p(out[1]),
          nx, ny, tau, lam, theta, nscales, fscale, zfactor, warps, epsilon, False)
        return out, nscales

    def gaussian(self, I, sigma):
        """gaussian() of lib/tvl1flow/mask.c:216, in place there, a copy here"""
        a = np.ascontiguousarray(I, np.float32).copy()
        f = self.lib.gaussian
        f.argtypes = [_fp, C.c_int, C.c_int, C.c_double]
        f.restype = None
        f(_p(a), a.shape[1], a.shape[0], float(sigma))
        return a

    def zoom_out(self, I, factor):
        """zoom_out of lib/tvl1flow/zoom.c:44"""
        ny, nx = I.shape
        nxx, nyy = C.c_int(), C.c_int()
        self.lib.zoom_size.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_float]
        self.lib.zoom_size(nx, ny, C.byref(nxx), C.byref(nyy), factor)
        out = np.zeros((nyy.value, nxx.value), np.float32)
        f = self.lib.zoom_out
        f.argtypes = [_fp, _fp, C.c_int, C.c_int, C.c_float]
        f.restype = None
        f(_p(np.ascontiguousarray(I, np.float32)), _p(out), nx, ny, factor)
        return out

    def zoom_in(self, I, nxx, nyy):
        """zoom_in of lib/tvl1flow/zoom.c:91"""
        ny, nx = I.shape
        out = np.zeros((nyy, nxx), np.float32)
        f = self.lib.zoom_in
        f.argtypes = [_fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int]
        f.restype = None
        f(_p(np.ascontiguousarray(I, np.float32)), _p(out), nx, ny, nxx, nyy)
        return out


def tvl1_truth(nx, ny):
    """the motion of tvl1_frames: I1(x) = I0(x - d(x)), a smooth field of a few pixels"""
    xs, ys = np.meshgrid(np.arange(nx, dtype=np.float64), np.arange(ny, dtype=np.float64))
    return 2.5 + 1.5 * np.sin(ys / ny * 3.0), -1.25 + 1.0 * np.cos(xs / nx * 2.0)


def tvl1_frames(nx, ny, seed=5, noise=4.0):
    """two frames of a moving textured scene in arbitrary units (not normalised, not smoothed): the input
    of the whole estimator.  Band-limited texture (below 0.3 rad / px) so that the coarse scales see it
    and the estimator converges to the motion (tvl1_truth) instead of aliasing."""
    rng = np.random.default_rng(seed)
    xs, ys = np.meshgrid(np.arange(nx, dtype=np.float64), np.arange(ny, dtype=np.float64))
    r = np.random.default_rng(seed + 1)
    comps = [(r.uniform(0.03, 0.3) * r.choice([-1, 1]), r.uniform(0.03, 0.3) * r.choice([-1, 1]), r.uniform(0, 6.28))
             for _ in range(24)]

    def img(x, y):
        out = np.zeros_like(x)
        for fx, fy, ph in comps:
            out += 1.0 / (abs(fx) + abs(fy)) ** 0.5 * np.sin(fx * x + fy * y + ph)
        return out
    dx, dy = tvl1_truth(nx, ny)
    I0 = img(xs, ys) + rng.normal(0, noise / 255.0, (ny, nx))
    I1 = img(xs - dx, ys - dy) + rng.normal(0, noise / 255.0, (ny, nx))
    return (I0 * 20 + 40).astype(np.float32), (I1 * 20 + 40).astype(np.float32)


def tvl1_pair(nx, ny, shift=(1.5, -0.75), seed=3):
    """a smooth textured image pair, I1(x) = I0(x - shift) up to a little noise, in 0..255 like a
    normalised, pre-smoothed pyramid level (tvl1flow_lib.c:379-384)"""
    rng = np.random.default_rng(seed)
    xs, ys = np.meshgrid(np.arange(nx, dtype=np.float64), np.arange(ny, dtype=np.float64))

    def img(dx, dy):
        r = np.random.default_rng(seed + 1)
        out = np.zeros((ny, nx))
        for _ in range(12):
            a, fx, fy, ph = r.uniform(8, 30), r.uniform(-0.25, 0.25), r.uniform(-0.25, 0.25), r.uniform(0, 6.28)
            out += a * np.sin(fx * (xs - dx) + fy * (ys - dy) + ph)
        return out
    I0 = img(0, 0) + rng.normal(0, 0.5, (ny, nx))
    I1 = img(*shift) + rng.normal(0, 0.5, (ny, nx))
    lo, hi = min(I0.min(), I1.min()), max(I0.max(), I1.max())
    sc = 255.0 / (hi - lo)
    return ((I0 - lo) * sc).astype(np.float32), ((I1 - lo) * sc).astype(np.float32)


class Ref:
    """The reference's own six entry points (reference src/nlkalman.h:14-53)."""

    def __init__(self, threads: int | None = 1):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(f"{REF_SO} missing: run `make -C oracle ref` where /root/reference exists")
        self.lib = C.CDLL(REF_SO)
        try:
            self.omp = C.CDLL("libgomp.so.1", mode=C.RTLD_GLOBAL)
        except OSError:
            self.omp = None
        L = self.lib
        L.rgb2opp.argtypes = L.opp2rgb.argtypes = [_fp, C.c_int, C.c_int, C.c_int]
        L.warp_bicubic.argtypes = [_fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int]
        L.nlkalman_default_params.argtypes = [C.POINTER(Params), C.c_float, C.c_int]
        sig = [_fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_float, Params, C.c_int]
        L.nlkalman_filter_frame.argtypes = sig
        L.nlkalman_smooth_frame.argtypes = sig
        for f in (L.rgb2opp, L.opp2rgb, L.warp_bicubic, L.nlkalman_default_params,
                  L.nlkalman_filter_frame, L.nlkalman_smooth_frame):
            f.restype = None
        if threads is not None:
            self.set_threads(threads)

    def set_threads(self, n: int):
        if self.omp is not None:
            self.omp.omp_set_num_threads(int(n))

    def max_threads(self) -> int:
        return int(self.omp.omp_get_max_threads()) if self.omp is not None else 1

    def default_params(self, sigma, mode, p: Params | None = None) -> Params:
        p = p or Params.auto()
        self.lib.nlkalman_default_params(C.byref(p), float(sigma), mode)
        return p

    def rgb2opp(self, im):
        h, w, ch = im.shape
        self.lib.rgb2opp(_p(im), w, h, ch)
        return im

    def opp2rgb(self, im):
        h, w, ch = im.shape
        self.lib.opp2rgb(_p(im), w, h, ch)
        return im

    def warp_bicubic(self, im, of, msk):
        h, w, ch = im.shape
        out = np.empty_like(im)
        self.lib.warp_bicubic(_p(out), _p(im), _p(of), _p(msk), w, h, ch)
        return out

    def filter_frame(self, nisy1, deno0, bsic1, sigma, prms: Params):
        h, w, ch = nisy1.shape
        out = np.empty_like(nisy1)
        self.lib.nlkalman_filter_frame(_p(out), _p(nisy1), _p(deno0), _p(bsic1), w, h, ch,
                                       float(sigma), prms, 0)
        return out

    def smooth_frame(self, filt1, smoo0, bsic1, sigma, prms: Params):
        h, w, ch = filt1.shape
        out = np.empty_like(filt1)
        self.lib.nlkalman_smooth_frame(_p(out), _p(filt1), _p(smoo0), _p(bsic1), w, h, ch,
                                       float(sigma), prms, 0)
        return out


class _Dump(C.Structure):
    _fields_ = [("kmax", C.c_int), ("nk", C.POINTER(C.c_int)), ("np0", C.POINTER(C.c_int)),
                ("knn_xy", C.POINTER(C.c_int)), ("knn_d", _fp),
                ("prev_p", C.POINTER(C.c_ubyte)), ("active", C.POINTER(C.c_ubyte)), ("vp", _fp)]


class Port:
    """Our C restatement (oracle/nlk_port.c)."""

    def __init__(self):
        if not os.path.exists(PORT_SO):
            build(ref=False)
        self.lib = C.CDLL(PORT_SO)
        L = self.lib
        L.port_rgb2opp.argtypes = L.port_opp2rgb.argtypes = [_fp, C.c_int, C.c_int, C.c_int]
        L.port_warp_bicubic.argtypes = [_fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int]
        L.port_default_params.argtypes = [C.POINTER(Params), C.c_float, C.c_int]
        L.port_window.argtypes = [_fp, C.c_int]
        L.port_dct2.argtypes = [_fp, C.c_int, C.c_int, C.c_int]
        L.port_pass.argtypes = [C.c_int, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_float,
                                Params, C.POINTER(_Dump)]
        for f in (L.port_rgb2opp, L.port_opp2rgb, L.port_warp_bicubic, L.port_default_params,
                  L.port_window, L.port_dct2, L.port_pass):
            f.restype = None

    def default_params(self, sigma, mode, p: Params | None = None) -> Params:
        p = p or Params.auto()
        self.lib.port_default_params(C.byref(p), float(sigma), mode)
        return p

    def rgb2opp(self, im):
        h, w, ch = im.shape
        self.lib.port_rgb2opp(_p(im), w, h, ch)
        return im

    def opp2rgb(self, im):
        h, w, ch = im.shape
        self.lib.port_opp2rgb(_p(im), w, h, ch)
        return im

    def warp_bicubic(self, im, of, msk):
        h, w, ch = im.shape
        out = np.empty_like(im)
        self.lib.port_warp_bicubic(_p(out), _p(im), _p(of), _p(msk), w, h, ch)
        return out

    def window(self, psz):
        out = np.empty((psz, psz), np.float32)
        self.lib.port_window(_p(out), psz)
        return out

    def dct2(self, tiles, inverse=False):
        t = np.ascontiguousarray(tiles, dtype=np.float32).copy()
        n, psz, _ = t.shape
        self.lib.port_dct2(_p(t), psz, n, 1 if inverse else 0)
        return t

    @staticmethod
    def grid(w, h, psz):
        step = psz // 2
        return (w - psz) // step + 1, (h - psz) // step + 1

    def run_pass(self, mode, in1, prev0, bsic1, sigma, prms: Params, dump: bool = False):
        h, w, ch = in1.shape
        out = np.empty_like(in1)
        d = None
        res = None
        if dump:
            gw, gh = self.grid(w, h, prms.patch_sz)
            G = gw * gh
            kmax = max(prms.npatches_x, prms.npatches_t, 1)
            res = dict(kmax=kmax, gw=gw, gh=gh,
                       nk=np.zeros(G, np.int32), np0=np.zeros(G, np.int32),
                       knn_xy=np.full((G, kmax, 2), -1, np.int32),
                       knn_d=np.zeros((G, kmax), np.float32),
                       prev_p=np.zeros(G, np.uint8), active=np.zeros(G, np.uint8),
                       vp=np.zeros(G, np.float32))
            d = _Dump(kmax, res["nk"].ctypes.data_as(C.POINTER(C.c_int)),
                      res["np0"].ctypes.data_as(C.POINTER(C.c_int)),
                      res["knn_xy"].ctypes.data_as(C.POINTER(C.c_int)), _p(res["knn_d"]),
                      res["prev_p"].ctypes.data_as(C.POINTER(C.c_ubyte)),
                      res["active"].ctypes.data_as(C.POINTER(C.c_ubyte)), _p(res["vp"]))
        self.lib.port_pass(mode, _p(out), _p(in1), _p(prev0), _p(bsic1), w, h, ch, float(sigma),
                           prms, C.byref(d) if d is not None else None)
        return (out, res) if dump else out

    def filter_frame(self, nisy1, deno0, bsic1, sigma, prms, dump=False):
        return self.run_pass(PASS_FILTER, nisy1, deno0, bsic1, sigma, prms, dump)

    def smooth_frame(self, filt1, smoo0, bsic1, sigma, prms, dump=False):
        return self.run_pass(PASS_SMOOTH, filt1, smoo0, bsic1, sigma, prms, dump)


def occlusion_from_flow(of, th):
    """TEST INFRASTRUCTURE: the plambda expression the pipeline script evaluates on a flow field
    (reference scripts/nlkalman-seq.sh:70-72 and :95-97)
        x(0,0)[0] x(-1,0)[0] - x(0,0)[1] x(0,-1)[1] - + fabs TH > 255 *
    with plambda's default nearest-sample boundary (reference lib/imscript-lite/src/getpixel.c:18-29),
    in float32 like plambda's stack machine.  of: (h, w, 2) float32 -> (h, w) float32 of 0 / 255."""
    import numpy as np
    of = np.asarray(of, np.float32)
    u, v = of[..., 0], of[..., 1]
    ul = np.concatenate([u[:, :1], u[:, :-1]], axis=1)
    vu = np.concatenate([v[:1, :], v[:-1, :]], axis=0)
    d = np.abs((u - ul).astype(np.float32) + (v - vu).astype(np.float32)).astype(np.float32)
    return np.where(d > np.float32(th), np.float32(255), np.float32(0)).astype(np.float32)

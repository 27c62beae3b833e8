"""ctypes binding of libnlkalman_b200.so (include/nlkalman.h + include/nlkalman_b200.h).

Host-side mirror of the reference interface for Python callers (tests, bench): the six
drop-in entry points keep the reference's names and argument meaning
(reference src/nlkalman.h:14-53); ``Context`` wraps the additive resident-state API.
There is no fallback of any kind: a missing library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnlkalman_b200.so")

FLT1, FLT2, SMO1 = 0, 1, 2


class NlkError(RuntimeError):
    pass


class Params(C.Structure):
    """struct nlkalman_params (include/nlkalman.h; reference src/nlkalman.h:22-37)."""
    _fields_ = [("patch_sz", C.c_int), ("search_sz_x", C.c_int), ("search_sz_t", C.c_int),
                ("npatches_x", C.c_int), ("npatches_t", C.c_int), ("npatches_tagg", C.c_int),
                ("dista_lambda", C.c_float), ("beta_x", C.c_float), ("beta_t", C.c_float)]

    @classmethod
    def auto(cls, **kw):
        """All fields -1 ("automatic", reference src/main-flt.c:39-52), then overrides."""
        p = cls(-1, -1, -1, -1, -1, -1, -1.0, -1.0, -1.0)
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class Tvl1Params(C.Structure):
    """struct nlk_tvl1_params (include/nlkalman_b200.h): the arguments of the reference's tvl1flow program
    after nproc (lib/tvl1flow/main.c:95-104); zeros mean "default" like on its command line."""
    _fields_ = [("tau", C.c_float), ("lam", C.c_float), ("theta", C.c_float), ("nscales", C.c_int),
                ("fscale", C.c_int), ("zfactor", C.c_float), ("warps", C.c_int), ("epsilon", C.c_float)]

    @classmethod
    def script(cls, dw=0.25, fscale=1):
        """what scripts/nlkalman-seq.sh:51 passes: "NPROC 0 DW 0 0 FSCALE" """
        return cls(0.0, dw, 0.0, 0, fscale, 0.0, 0, 0.0)


class StripPlan(C.Structure):
    """struct nlk_strip_plan (include/nlkalman_b200.h): rows of one rank in a strip-sharded pass."""
    _fields_ = [("gw", C.c_int), ("gh", C.c_int), ("nbw", C.c_int), ("gy0", C.c_int), ("gy1", C.c_int),
                ("oy0", C.c_int), ("oy1", C.c_int), ("ey0", C.c_int), ("ey1", C.c_int),
                ("chunk_g", C.c_int), ("chunk_y", C.c_int)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_bp = C.POINTER(C.c_ubyte)
_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NlkError(f"{LIB_PATH} is missing: build it with `make -C {HERE}` "
                       "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.nlk_last_error.restype = C.c_char_p
    L.nlk_device_count.restype = C.c_int
    L.nlk_ctx_create.restype = vp
    L.nlk_ctx_create.argtypes = [C.c_int] * 4
    L.nlk_ctx_destroy.argtypes = [vp]
    L.nlk_ctx_destroy.restype = None
    L.nlk_ctx_sync.argtypes = [vp]
    L.nlk_ctx_launch_count.argtypes = [vp]
    L.nlk_ctx_launch_count.restype = C.c_longlong
    L.nlk_ctx_stream.argtypes = [vp]
    L.nlk_ctx_stream.restype = vp
    L.nlk_ctx_profile.argtypes = [vp, C.c_int]
    L.nlk_ctx_profile_collect.argtypes = [vp, C.POINTER(C.c_double), _ip]
    L.nlk_fp32_peak.argtypes = [vp, C.c_float, C.POINTER(C.c_double)]
    L.nlk_ctx_profile_alpha.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.nlk_seq_set_mask_mode.argtypes = [vp, C.c_int, C.c_float]
    L.nlk_host_alloc.argtypes = [C.c_size_t]
    L.nlk_host_alloc.restype = vp
    L.nlk_host_free.argtypes = [vp]
    L.nlk_host_free.restype = None
    L.nlk_rgb2opp_dev.argtypes = [vp, vp, vp]
    L.nlk_opp2rgb_dev.argtypes = [vp, vp, vp]
    L.nlk_warp_dev.argtypes = [vp, vp, vp, vp, vp]
    L.nlk_pass_dev.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_float, Params]
    L.nlk_occlusion_dev.argtypes = [vp, vp, vp, C.c_float]
    L.nlk_occlusion_host.argtypes = [vp, vp, vp, C.c_float]
    L.nlk_strip_plan.argtypes = [C.c_int, C.c_int, C.c_int, Params, C.c_int, C.c_int, C.POINTER(StripPlan)]
    L.nlk_colour_rows_dev.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int]
    L.nlk_warp_rows_dev.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int]
    L.nlk_strip_search.argtypes = [vp, C.c_int, vp, vp, vp, C.c_float, Params, C.c_int, C.c_int, vp, vp]
    L.nlk_strip_filter.argtypes = [vp]
    L.nlk_strip_lane.argtypes = [vp, C.c_int, C.c_int]
    L.nlk_lane_record.argtypes = [vp, C.c_int]
    L.nlk_lane_wait.argtypes = [vp, C.c_int]
    L.nlk_strip_normalize.argtypes = [vp, vp, C.c_int, C.c_int]
    L.nlk_peer_header_bytes.restype = C.c_size_t
    L.nlk_peer_slab_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.nlk_peer_ipc_export.argtypes = [vp, vp, C.c_char_p]
    L.nlk_peer_ipc_import.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    L.nlk_peer_bind.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp), C.c_size_t]
    L.nlk_peer_push.argtypes = [vp, C.c_size_t, C.c_size_t, C.c_uint, C.c_int, C.c_uint, C.c_int]
    L.nlk_peer_push_add.argtypes = [vp, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_uint]
    L.nlk_peer_signal.argtypes = [vp, C.c_int, C.c_uint, C.c_uint]
    L.nlk_peer_wait.argtypes = [vp, C.c_int, C.c_uint, C.c_uint]
    L.nlk_peer_error.argtypes = [vp, C.POINTER(C.c_uint)]
    L.nlk_warp_rows_peer_dev.argtypes = [vp, vp, C.c_size_t, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.nlk_dev_free.argtypes = [vp, vp]
    L.nlk_dev_free.restype = None
    L.nlk_seq_reset.argtypes = [vp]
    L.nlk_seq_filter_dev.argtypes = [vp, vp, vp, vp, C.c_float, Params, Params, vp, vp]
    L.nlk_seq_filter_host.argtypes = [vp, vp, vp, vp, C.c_float, Params, Params, vp, vp]
    L.nlk_seq_submit_host.argtypes = [vp, vp, vp, vp, C.c_float, Params, Params, vp, vp]
    L.nlk_seq_drain.argtypes = [vp]
    L.nlk_seq_submit_dev.argtypes = [vp, vp, vp, vp, C.c_float, Params, Params, vp, vp]
    L.nlk_seq_join.argtypes = [vp]
    L.nlk_seq_smooth_start_dev.argtypes = [vp, vp]
    L.nlk_seq_smooth_dev.argtypes = [vp, vp, vp, vp, C.c_float, Params, vp]
    L.nlk_seq_smooth_start_host.argtypes = [vp, vp]
    L.nlk_seq_smooth_host.argtypes = [vp, vp, vp, vp, C.c_float, Params, vp]
    L.nlk_pass_host_debug.argtypes = [vp, C.c_int, _fp, _fp, _fp, _fp, C.c_float, Params, C.c_int,
                                      _ip, _ip, _ip, _fp, _bp, _bp, _fp]
    L.nlk_dct_host.argtypes = [vp, _fp, C.c_int, C.c_int, C.c_int]
    L.nlk_tvl1_level_host.argtypes = [vp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                      C.c_int, C.c_float, _ip]
    L.nlk_tvl1_level_dev.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                     C.c_int, C.c_float, _ip]
    L.nlk_flow_mask_dev.argtypes = [vp, vp, vp, vp, vp, Tvl1Params, C.c_float]
    L.nlk_tvl1_default_params.argtypes = [C.POINTER(Tvl1Params)]
    L.nlk_tvl1_default_params.restype = None
    L.nlk_tvl1_scales.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int]
    L.nlk_tvl1_flow_host.argtypes = [vp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                     C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, _ip]
    L.nlk_tvl1_flow_dev.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                    C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, _ip]
    # drop-in entry points
    L.rgb2opp.argtypes = L.opp2rgb.argtypes = [_fp, C.c_int, C.c_int, C.c_int]
    L.warp_bicubic.argtypes = [_fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int]
    L.nlkalman_default_params.argtypes = [C.POINTER(Params), C.c_float, C.c_int]
    sig = [_fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_float, Params, C.c_int]
    L.nlkalman_filter_frame.argtypes = sig
    L.nlkalman_smooth_frame.argtypes = sig
    for f in (L.rgb2opp, L.opp2rgb, L.warp_bicubic, L.nlkalman_default_params,
              L.nlkalman_filter_frame, L.nlkalman_smooth_frame):
        f.restype = None
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise NlkError(f"libnlkalman_b200 error {rc}: {lib().nlk_last_error().decode()}")


def _p(a):
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.dtype == np.float32 and a.flags["C_CONTIGUOUS"], \
        "float32 C-contiguous HWC arrays only"
    return a.ctypes.data_as(_fp)


def _vp(a):
    """host ndarray, torch tensor (host or device) or raw int -> void*"""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        assert a.dtype in (np.float32, np.uint8) and a.flags["C_CONTIGUOUS"]
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous()
        return C.c_void_p(a.data_ptr())
    if hasattr(a, "ctypes"):  # other ndarray dtypes (bitmaps)
        return C.c_void_p(a.ctypes.data)
    raise TypeError(type(a))


# ---- the six reference entry points, same names and argument meaning ------------------------

def default_params(sigma: float, mode: int, p: Params | None = None) -> Params:
    p = p if p is not None else Params.auto()
    lib().nlkalman_default_params(C.byref(p), float(sigma), int(mode))
    return p


def rgb2opp(im: np.ndarray) -> np.ndarray:
    h, w, ch = im.shape
    lib().rgb2opp(_p(im), w, h, ch)
    return im


def opp2rgb(im: np.ndarray) -> np.ndarray:
    h, w, ch = im.shape
    lib().opp2rgb(_p(im), w, h, ch)
    return im


def warp_bicubic(im: np.ndarray, of: np.ndarray, msk: np.ndarray | None) -> np.ndarray:
    h, w, ch = im.shape
    out = np.empty_like(im)
    lib().warp_bicubic(_p(out), _p(im), _p(of), _p(msk), w, h, ch)
    return out


def nlkalman_filter_frame(nisy1, deno0, bsic1, sigma, prms: Params) -> np.ndarray:
    h, w, ch = nisy1.shape
    out = np.empty_like(nisy1)
    lib().nlkalman_filter_frame(_p(out), _p(nisy1), _p(deno0), _p(bsic1), w, h, ch, float(sigma), prms, 0)
    return out


def nlkalman_smooth_frame(filt1, smoo0, bsic1, sigma, prms: Params) -> np.ndarray:
    h, w, ch = filt1.shape
    out = np.empty_like(filt1)
    lib().nlkalman_smooth_frame(_p(out), _p(filt1), _p(smoo0), _p(bsic1), w, h, ch, float(sigma), prms, 0)
    return out


# ---- resident-state context --------------------------------------------------------------------

class Context:
    """nlk_ctx: buffers and the recursion state of one sequence resident on one GPU."""

    def __init__(self, w: int, h: int, ch: int, device: int = 0):
        L = lib()
        self.w, self.h, self.ch, self.device = w, h, ch, device
        self._h = L.nlk_ctx_create(w, h, ch, device)
        if not self._h:
            raise NlkError(f"nlk_ctx_create failed: {L.nlk_last_error().decode()}")

    def close(self):
        if getattr(self, "_h", None):
            lib().nlk_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def sync(self):
        _check(lib().nlk_ctx_sync(self._h))

    @property
    def launches(self) -> int:
        return int(lib().nlk_ctx_launch_count(self._h))

    @property
    def stream(self) -> int:
        return int(lib().nlk_ctx_stream(self._h) or 0)

    # per-kernel CUDA-event timing (include/nlkalman_b200.h)
    KERNELS = ["colour", "warp_bicubic", "valid_map", "search_knn", "mask_resolve", "group_filter",
               "normalize", "memset", "peer_wait", "peer_push"]
    PASS_KINDS = ["flt1_temporal", "flt1_spatial", "flt2_temporal", "flt2_spatial", "smoother", "other"]

    def profile(self, enable: bool):
        _check(lib().nlk_ctx_profile(self._h, 1 if enable else 0))

    def profile_collect(self):
        """{(kernel, pass_kind): (ms_sum, count)} since the last collection"""
        nk, npk = len(self.KERNELS), len(self.PASS_KINDS)
        ms = (C.c_double * (nk * npk))()
        cnt = (C.c_int * (nk * npk))()
        _check(lib().nlk_ctx_profile_collect(self._h, ms, cnt))
        out = {}
        for i, kn in enumerate(self.KERNELS):
            for j, pk in enumerate(self.PASS_KINDS):
                if cnt[i * npk + j]:
                    out[(kn, pk)] = (ms[i * npk + j], cnt[i * npk + j])
        return out

    def profile_alpha(self):
        """{pass_kind: (processed patches, grid patches)} summed over the profiled passes"""
        npk = len(self.PASS_KINDS)
        a, g = (C.c_double * npk)(), (C.c_double * npk)()
        _check(lib().nlk_ctx_profile_alpha(self._h, a, g))
        return {pk: (a[j], g[j]) for j, pk in enumerate(self.PASS_KINDS) if g[j] > 0}

    MASK_FLOAT, MASK_U8, MASK_FROM_FLOW = 0, 1, 2

    def seq_set_mask_mode(self, mode: int, th: float = 0.0):
        _check(lib().nlk_seq_set_mask_mode(self._h, int(mode), float(th)))

    def fp32_peak(self, ms: float = 200.0) -> float:
        v = C.c_double(0)
        _check(lib().nlk_fp32_peak(self._h, float(ms), C.byref(v)))
        return v.value

    # device-pointer operations (torch CUDA tensors or raw pointers)
    def rgb2opp_dev(self, dst, src):
        _check(lib().nlk_rgb2opp_dev(self._h, _vp(dst), _vp(src)))

    def opp2rgb_dev(self, dst, src):
        _check(lib().nlk_opp2rgb_dev(self._h, _vp(dst), _vp(src)))

    def warp_dev(self, imw, im, of, msk):
        _check(lib().nlk_warp_dev(self._h, _vp(imw), _vp(im), _vp(of), _vp(msk)))

    def pass_dev(self, smooth, out, in1, prev0, bsic1, sigma, prms: Params):
        _check(lib().nlk_pass_dev(self._h, int(smooth), _vp(out), _vp(in1), _vp(prev0), _vp(bsic1),
                                  float(sigma), prms))

    def occlusion_dev(self, occ, of, th):
        _check(lib().nlk_occlusion_dev(self._h, _vp(occ), _vp(of), float(th)))

    def flow_mask_dev(self, of, occ, from_rgb, to_rgb, prms: "Tvl1Params", th: float):
        """flow (interleaved, like a .flo) and occlusion mask between two resident RGB frames: the
        tvl1flow + plambda step of the pipeline script (reference scripts/nlkalman-seq.sh:60-72)"""
        _check(lib().nlk_flow_mask_dev(self._h, _vp(of), _vp(occ), _vp(from_rgb), _vp(to_rgb), prms, float(th)))

    def occlusion(self, of: np.ndarray, th: float) -> np.ndarray:
        """0 / 255 mask from the divergence of the flow (reference scripts/nlkalman-seq.sh:70-72)"""
        h, w, two = of.shape
        assert (w, h, two) == (self.w, self.h, 2)
        occ = np.empty((h, w), np.float32)
        _check(lib().nlk_occlusion_host(self._h, _vp(occ), _vp(np.ascontiguousarray(of, np.float32)), float(th)))
        return occ

    # row ranges and the strip-sharded pass (full-frame device buffers, rows [row0, row1))
    def colour_rows_dev(self, dst, src, inverse, row0, row1):
        _check(lib().nlk_colour_rows_dev(self._h, _vp(dst), _vp(src), int(inverse), int(row0), int(row1)))

    def warp_rows_dev(self, imw, im, of, msk, row0, row1):
        _check(lib().nlk_warp_rows_dev(self._h, _vp(imw), _vp(im), _vp(of), _vp(msk), int(row0), int(row1)))

    def strip_search(self, smooth, in1, prev0, bsic1, sigma, prms: Params, gy0, gy1, nbr, accw):
        _check(lib().nlk_strip_search(self._h, int(smooth), _vp(in1), _vp(prev0), _vp(bsic1), float(sigma), prms,
                                      int(gy0), int(gy1), _vp(nbr), _vp(accw)))

    def strip_filter(self):
        _check(lib().nlk_strip_filter(self._h))

    def strip_lane(self, lane: int, reserve_sm: int = 0):
        _check(lib().nlk_strip_lane(self._h, int(lane), int(reserve_sm)))

    def lane_record(self, idx: int):
        _check(lib().nlk_lane_record(self._h, int(idx)))

    def lane_wait(self, idx: int):
        _check(lib().nlk_lane_wait(self._h, int(idx)))

    def strip_normalize(self, out, row0, row1):
        _check(lib().nlk_strip_normalize(self._h, _vp(out), int(row0), int(row1)))

    # peer-memory exchanges between the strips' GPUs (include/nlkalman_b200.h)
    def peer_slab_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        _check(lib().nlk_peer_slab_alloc(self._h, int(nbytes), C.byref(p)))
        return int(p.value)

    def peer_ipc_export(self, slab: int) -> bytes:
        buf = C.create_string_buffer(64)
        _check(lib().nlk_peer_ipc_export(self._h, C.c_void_p(slab), buf))
        return buf.raw

    def peer_ipc_import(self, handle: bytes) -> int:
        p = C.c_void_p()
        _check(lib().nlk_peer_ipc_import(self._h, handle, C.byref(p)))
        return int(p.value)

    def peer_bind(self, rank: int, nranks: int, slabs, slab_bytes: int):
        arr = (C.c_void_p * nranks)(*[C.c_void_p(int(x)) for x in slabs])
        _check(lib().nlk_peer_bind(self._h, int(rank), int(nranks), arr, int(slab_bytes)))

    def peer_push(self, off, nbytes, peer_mask, slot, value, side=0):
        _check(lib().nlk_peer_push(self._h, int(off), int(nbytes), int(peer_mask), int(slot), int(value) & 0xffffffff, int(side)))

    def peer_push_add(self, off, nbytes, peer, slot, value):
        _check(lib().nlk_peer_push_add(self._h, int(off), int(nbytes), int(peer), int(slot), int(value) & 0xffffffff))

    def peer_signal(self, slot, value, peer_mask):
        _check(lib().nlk_peer_signal(self._h, int(slot), int(value) & 0xffffffff, int(peer_mask)))

    def peer_wait(self, slot, value, src_mask):
        _check(lib().nlk_peer_wait(self._h, int(slot), int(value) & 0xffffffff, int(src_mask)))

    def warp_rows_peer_dev(self, imw, frame_off, of, msk, row0, row1, lo, hi, chunk_y):
        _check(lib().nlk_warp_rows_peer_dev(self._h, _vp(imw), int(frame_off), _vp(of), _vp(msk), int(row0), int(row1),
                                            int(lo), int(hi), int(chunk_y)))

    def peer_error(self) -> int:
        v = C.c_uint(0)
        _check(lib().nlk_peer_error(self._h, C.byref(v)))
        return int(v.value)

    def dev_free(self, ptr: int):
        lib().nlk_dev_free(self._h, C.c_void_p(int(ptr)))

    # resident sequence recursion
    def seq_reset(self):
        _check(lib().nlk_seq_reset(self._h))

    def seq_filter_dev(self, noisy, bflo, bocc, sigma, f1: Params, f2: Params, flt1_out, flt2_out):
        _check(lib().nlk_seq_filter_dev(self._h, _vp(noisy), _vp(bflo), _vp(bocc), float(sigma), f1, f2,
                                        _vp(flt1_out), _vp(flt2_out)))

    def seq_filter_host(self, noisy, bflo, bocc, sigma, f1: Params, f2: Params, flt1_out, flt2_out):
        _check(lib().nlk_seq_filter_host(self._h, _vp(noisy), _vp(bflo), _vp(bocc), float(sigma), f1, f2,
                                         _vp(flt1_out), _vp(flt2_out)))

    def seq_submit_host(self, noisy, bflo, bocc, sigma, f1: Params, f2: Params, flt1_out, flt2_out):
        _check(lib().nlk_seq_submit_host(self._h, _vp(noisy), _vp(bflo), _vp(bocc), float(sigma), f1, f2,
                                         _vp(flt1_out), _vp(flt2_out)))

    def seq_submit_dev(self, noisy, bflo, bocc, sigma, f1: Params, f2: Params, flt1_out, flt2_out):
        _check(lib().nlk_seq_submit_dev(self._h, _vp(noisy), _vp(bflo), _vp(bocc), float(sigma), f1, f2,
                                        _vp(flt1_out), _vp(flt2_out)))

    def seq_join(self):
        _check(lib().nlk_seq_join(self._h))

    def seq_drain(self):
        _check(lib().nlk_seq_drain(self._h))

    def seq_smooth_start_dev(self, last_rgb):
        _check(lib().nlk_seq_smooth_start_dev(self._h, _vp(last_rgb)))

    def seq_smooth_dev(self, flt_rgb, fflo, focc, sigma, s1: Params, smo_out):
        _check(lib().nlk_seq_smooth_dev(self._h, _vp(flt_rgb), _vp(fflo), _vp(focc), float(sigma), s1,
                                        _vp(smo_out)))

    def seq_smooth_start_host(self, last_rgb):
        _check(lib().nlk_seq_smooth_start_host(self._h, _vp(last_rgb)))

    def seq_smooth_host(self, flt_rgb, fflo, focc, sigma, s1: Params, smo_out):
        _check(lib().nlk_seq_smooth_host(self._h, _vp(flt_rgb), _vp(fflo), _vp(focc), float(sigma), s1,
                                         _vp(smo_out)))

    # parity-test helpers
    def pass_host_debug(self, smooth, in1, prev0, bsic1, sigma, prms: Params):
        h, w, ch = in1.shape
        assert (w, h, ch) == (self.w, self.h, self.ch)
        psz, step = prms.patch_sz, prms.patch_sz // 2
        gw, gh = (w - psz) // step + 1, (h - psz) // step + 1
        G = gw * gh
        kmax = max(prms.npatches_x, prms.npatches_t, 1)
        out = np.empty_like(in1)
        res = dict(kmax=kmax, gw=gw, gh=gh,
                   nk=np.zeros(G, np.int32), np0=np.zeros(G, np.int32),
                   knn_xy=np.full((G, kmax, 2), -1, np.int32), knn_d=np.zeros((G, kmax), np.float32),
                   prev_p=np.zeros(G, np.uint8), active=np.zeros(G, np.uint8), vp=np.zeros(G, np.float32))
        _check(lib().nlk_pass_host_debug(
            self._h, int(smooth), _p(out), _p(in1), _p(prev0), _p(bsic1), float(sigma), prms, kmax,
            res["nk"].ctypes.data_as(_ip), res["np0"].ctypes.data_as(_ip),
            res["knn_xy"].ctypes.data_as(_ip), _p(res["knn_d"]),
            res["prev_p"].ctypes.data_as(_bp), res["active"].ctypes.data_as(_bp), _p(res["vp"])))
        return out, res

    def tvl1_level(self, I0, I1, u1, u2, tau=0.25, lam=0.15, theta=0.3, warps=5, epsilon=0.01):
        """Dual TV-L1 flow at one scale (reference lib/tvl1flow/tvl1flow_lib.c:93-280): host arrays
        (ny, nx); returns (u1, u2, iterations per warping step)"""
        ny, nx = I0.shape
        a = np.ascontiguousarray(u1, np.float32).copy()
        b = np.ascontiguousarray(u2, np.float32).copy()
        its = np.zeros(max(warps, 1), np.int32)
        _check(lib().nlk_tvl1_level_host(self._h, _p(np.ascontiguousarray(I0, np.float32)),
                                         _p(np.ascontiguousarray(I1, np.float32)), _p(a), _p(b), nx, ny, float(tau),
                                         float(lam), float(theta), int(warps), float(epsilon), its.ctypes.data_as(_ip)))
        return a, b, its[:warps]

    def tvl1_flow(self, I0, I1, tau=0.25, lam=0.15, theta=0.3, nscales=100, fscale=0, zfactor=0.5, warps=5,
                  epsilon=0.01):
        """Dual TV-L1 flow, pyramid and all (reference lib/tvl1flow/tvl1flow_lib.c:345-477 with the scale
        cap of its driver, main.c:159-163): host images (ny, nx); returns (flow (2, ny, nx): u then v,
        iterations (nscales, warps))"""
        ny, nx = I0.shape
        nscales = tvl1_scales(nx, ny, zfactor, nscales)
        fscale = min(fscale, nscales)
        flow = np.empty((2, ny, nx), np.float32)
        its = np.zeros((nscales, max(warps, 1)), np.int32)
        _check(lib().nlk_tvl1_flow_host(self._h, _p(np.ascontiguousarray(I0, np.float32)),
                                        _p(np.ascontiguousarray(I1, np.float32)), _p(flow), nx, ny, float(tau), float(lam),
                                        float(theta), int(nscales), int(fscale), float(zfactor), int(warps),
                                        float(epsilon), its.ctypes.data_as(_ip) if warps > 0 else None))
        return flow, its[:, :warps]

    def tvl1_flow_dev(self, I0, I1, u1, u2, nx, ny, tau=0.25, lam=0.15, theta=0.3, nscales=100, fscale=0, zfactor=0.5,
                      warps=5, epsilon=0.01):
        """the same on device buffers (pointers or objects with data_ptr()), queued on the context's stream"""
        nscales = tvl1_scales(nx, ny, zfactor, nscales)
        _check(lib().nlk_tvl1_flow_dev(self._h, _vp(I0), _vp(I1), _vp(u1), _vp(u2), int(nx), int(ny), float(tau),
                                       float(lam), float(theta), int(nscales), int(min(fscale, nscales)), float(zfactor),
                                       int(warps), float(epsilon), None))

    def dct(self, tiles: np.ndarray, inverse: bool = False) -> np.ndarray:
        t = np.ascontiguousarray(tiles, dtype=np.float32).copy()
        n, psz, _ = t.shape
        _check(lib().nlk_dct_host(self._h, _p(t), psz, n, 1 if inverse else 0))
        return t


def tvl1_scales(nx: int, ny: int, zfactor: float = 0.5, nscales: int = 100) -> int:
    """nscales as the reference's tvl1flow driver caps it (lib/tvl1flow/main.c:159-161); host arithmetic"""
    return int(lib().nlk_tvl1_scales(int(nx), int(ny), float(zfactor), int(nscales)))


def strip_plan(w: int, h: int, smooth: int, prms: Params, nranks: int, rank: int) -> StripPlan:
    """nlk_strip_plan: host arithmetic only (works without a GPU)."""
    out = StripPlan()
    _check(lib().nlk_strip_plan(int(w), int(h), int(smooth), prms, int(nranks), int(rank), C.byref(out)))
    return out


def device_count() -> int:
    n = lib().nlk_device_count()
    if n < 0:
        raise NlkError(lib().nlk_last_error().decode())
    return n

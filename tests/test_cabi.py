"""The drop-in boundary without a GPU: the shared library loads, exports every function that
include/nlkalman.h (the reference's six entry points, reference src/nlkalman.h:14-53),
include/tvl1flow.h (the two of its flow library, reference lib/tvl1flow/tvl1flow_lib.c:93, :345) and
include/nlkalman_b200.h declare, the pure-host entry points work, and the ones that need a
device fail with an error code and a message instead of falling back to anything."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADERS = [os.path.join(ROOT, "include", n) for n in ("nlkalman.h", "nlkalman_b200.h", "tvl1flow.h")]
IO_HEADER = os.path.join(ROOT, "bwd_nlkalman_b200", "host", "nlk_image_io.h")


def declared_functions(path):
    """names of the function declarations of a C header (comments and macros stripped)"""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    src = re.sub(r"^\s*#.*?$", " ", src, flags=re.M)
    src = src.replace('extern "C" {', " ")
    src = re.sub(r"\{[^{}]*\}", " ", src)          # struct / enum bodies
    names = []
    for decl in src.split(";"):
        m = re.search(r"\b([A-Za-z_]\w*)\s*\([^()]*\)\s*$", decl.strip(), flags=re.S)
        if m and m.group(1) not in ("defined", "__attribute__"):
            names.append(m.group(1))
    return names


def test_headers_declare_the_reference_entry_points():
    names = set(declared_functions(HEADERS[0]))
    assert {"rgb2opp", "opp2rgb", "warp_bicubic", "nlkalman_default_params", "nlkalman_filter_frame",
            "nlkalman_smooth_frame"} <= names


@pytest.mark.parametrize("header", HEADERS)
def test_library_exports_every_declared_function(nlk, header):
    lib = nlk.lib()
    names = declared_functions(header)
    assert len(names) >= (2 if header.endswith("tvl1flow.h") else 6), names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in {os.path.basename(header)} but not exported: {missing}"


def test_image_io_library_exports_its_header():
    so = os.path.join(ROOT, "bwd_nlkalman_b200", "libnlk_image_io.so")
    lib = C.CDLL(so)
    names = declared_functions(IO_HEADER)
    assert names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_params_struct_layout_matches_the_reference(nlk):
    """struct nlkalman_params is passed BY VALUE across the boundary (reference
    src/nlkalman.h:22-37): 6 ints + 3 floats, 36 bytes, this order"""
    P = nlk.Params
    assert C.sizeof(P) == 36
    want = ["patch_sz", "search_sz_x", "search_sz_t", "npatches_x", "npatches_t", "npatches_tagg",
            "dista_lambda", "beta_x", "beta_t"]
    assert [f[0] for f in P._fields_] == want
    assert [getattr(P, n).offset for n in want] == [4 * i for i in range(9)]


def test_pure_host_entry_points_work_without_a_gpu(nlk):
    lib = nlk.lib()
    # nlkalman_default_params: reference src/nlkalman.c:426-487 (sigma 20 row of SURVEY App. A)
    p = nlk.default_params(20.0, nlk.FLT1)
    assert (p.patch_sz, p.search_sz_x, p.search_sz_t, p.npatches_x, p.npatches_t, p.npatches_tagg) == (8, 10, 5, 50, 30, 20)
    assert abs(p.beta_x - 3.11) < 1e-6 and abs(p.beta_t - 1.95) < 1e-6
    lib.nlk_last_error.restype = C.c_char_p
    assert isinstance(lib.nlk_last_error(), bytes)


def test_no_cpu_fallback_without_a_device(nlk):
    """without a CUDA device the context cannot be created: an error and a message, never a
    CPU path (the oracle is test infrastructure only)"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = nlk.lib()
    lib.nlk_ctx_create.restype = C.c_void_p
    lib.nlk_ctx_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    h = lib.nlk_ctx_create(64, 48, 3, 0)
    assert not h
    lib.nlk_last_error.restype = C.c_char_p
    assert lib.nlk_last_error()   # a message says why


def test_product_does_not_import_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import, link or run anything
    under oracle/ (comments may cite it)"""
    pkg = os.path.join(ROOT, "bwd_nlkalman_b200")
    uses = re.compile(r"^\s*(from|import)\s+oracle\b|#\s*include\s*[\"<][^\n]*(oracle|nlk_port|fftw)"
                      r"|libnlk_port|libnlkalman_ref|nlkalman-(flt|smo)-ref|-lnlk_port", re.M)
    for dirpath, dirs, files in os.walk(pkg):
        dirs[:] = [d for d in dirs if d not in ("__pycache__", "bin")]
        for f in files:
            if f.endswith((".py", ".c", ".h", ".cu", ".cuh", ".sh")) or f == "Makefile":
                path = os.path.join(dirpath, f)
                m = uses.search(open(path, errors="replace").read())
                assert m is None, f"{path}: {m.group(0)!r}"

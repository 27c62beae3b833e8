"""GPU tests of the host drivers: nlkalman-flt / nlkalman-smo against the UNMODIFIED reference
programs (oracle/_ref/nlkalman-*-ref, built from reference src/main-flt.c and src/main-smo.c,
one OpenMP thread) on the same input files with the same command lines, and nlkalman-seq
(state resident in HBM) against the chain of per-frame invocations."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from common import TOL_MAXABS, maxabs

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bwd_nlkalman_b200", "bin")
REF = os.path.join(ROOT, "oracle", "_ref")


def _write_pfm(path, a):
    a = np.ascontiguousarray(a, np.float32)
    h, w = a.shape[:2]
    ch = 1 if a.ndim == 2 else a.shape[2]
    with open(path, "wb") as f:
        f.write(f"P{'F' if ch == 3 else 'f'}\n{w} {h}\n-1\n".encode())
        f.write(a.tobytes())


def _read_pfm(path):
    raw = open(path, "rb").read()
    head = raw.split(b"\n", 3)
    ch = 3 if head[0] == b"PF" else 1
    w, h = (int(x) for x in head[1].split())
    return np.frombuffer(head[3], np.float32).reshape(h, w, ch)


def _write_flo(path, a):
    h, w = a.shape[:2]
    with open(path, "wb") as f:
        f.write(b"PIEH" + np.array([w, h], np.int32).tobytes() + np.ascontiguousarray(a, np.float32).tobytes())


def _write_pgm(path, a):
    h, w = a.shape
    with open(path, "wb") as f:
        f.write(f"P5\n{w} {h}\n255\n".encode() + a.astype(np.uint8).tobytes())


def _run(exe, *args, ok=(0,), env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([exe, *[str(a) for a in args]], capture_output=True, text=True, env=e)
    assert r.returncode in ok, (exe, args, r.returncode, r.stderr[-2000:])
    return r


@pytest.fixture(scope="module")
def scene(tmp_path_factory):
    from bwd_nlkalman_b200 import synth
    d = tmp_path_factory.mktemp("cli")
    w, h, ch, sigma = 160, 120, 3, 20.0
    for t in range(3):
        _write_pfm(d / f"n{t}.pfm", synth.noisy_frame(w, h, ch, t, sigma))
    _write_flo(d / "bflo.flo", synth.backward_flow(w, h))
    _write_flo(d / "fflo.flo", synth.forward_flow(w, h))
    occ = np.zeros((h, w), np.uint8)
    occ[40:60, 70:100] = 255
    _write_pgm(d / "occ.pgm", occ)
    return d, sigma


def test_flt_and_smo_against_reference_programs(scene):
    d, sigma = scene
    ours_flt, ours_smo = os.path.join(BIN, "nlkalman-flt"), os.path.join(BIN, "nlkalman-smo")
    ref_flt, ref_smo = os.path.join(REF, "nlkalman-flt-ref"), os.path.join(REF, "nlkalman-smo-ref")
    for p in (ours_flt, ours_smo):
        assert os.path.exists(p), f"{p} missing: run __graft_entry__.build()"
    if not os.path.exists(ref_flt):
        pytest.skip("oracle/_ref programs not built (needs /root/reference)")
    # one thread: the reference's output depends on the thread count (processed-pixel mask), and its
    # smoother program pins two threads (src/main-smo.c:23) -- OMP_THREAD_LIMIT caps that to one
    one = {"OMP_NUM_THREADS": "1", "OMP_THREAD_LIMIT": "1"}
    # frame 0: both filterings, spatial (reference scripts/nlkalman-seq.sh:39-41)
    _run(ref_flt, "-i", d / "n0.pfm", "-s", sigma, "--flt11", d / "r_a1.pfm", "--flt21", d / "r_a2.pfm", env=one)
    _run(ours_flt, "-i", d / "n0.pfm", "-s", sigma, "--flt11", d / "g_a1.pfm", "--flt21", d / "g_a2.pfm")
    assert maxabs(_read_pfm(d / "g_a1.pfm"), _read_pfm(d / "r_a1.pfm")) <= TOL_MAXABS
    assert maxabs(_read_pfm(d / "g_a2.pfm"), _read_pfm(d / "r_a2.pfm")) <= TOL_MAXABS
    # frame 1, the two invocations of the script (:80-81, :100-102), previous state = the reference's
    for exe, tag, env in ((ref_flt, "r", one), (ours_flt, "g", None)):
        _run(exe, "-i", d / "n1.pfm", "-s", sigma, "--f2_p", 0, "-o", d / "bflo.flo", "-k", d / "occ.pgm",
             "--flt10", d / "r_a1.pfm", "--flt11", d / f"{tag}_b1.pfm", env=env)
    assert maxabs(_read_pfm(d / "g_b1.pfm"), _read_pfm(d / "r_b1.pfm")) <= TOL_MAXABS
    for exe, tag, env in ((ref_flt, "r", one), (ours_flt, "g", None)):
        _run(exe, "-i", d / "n1.pfm", "-s", sigma, "--f1_p", 0, "-o", d / "bflo.flo", "-k", d / "occ.pgm",
             "--flt11", d / "r_b1.pfm", "--flt20", d / "r_a2.pfm", "--flt21", d / f"{tag}_b2.pfm", env=env)
    assert maxabs(_read_pfm(d / "g_b2.pfm"), _read_pfm(d / "r_b2.pfm")) <= TOL_MAXABS
    # smoother on frame 0 from frame 1 (:147-149); both programs exit with status 1 on success
    for exe, tag, env in ((ref_smo, "r", one), (ours_smo, "g", None)):
        _run(exe, "--flt1", d / "r_a2.pfm", "--smo0", d / "r_b2.pfm", "-s", sigma, "-o", d / "fflo.flo",
             "-k", d / "occ.pgm", "--smo1", d / f"{tag}_s0.pfm", ok=(1,), env=env)
    assert maxabs(_read_pfm(d / "g_s0.pfm"), _read_pfm(d / "r_s0.pfm")) <= TOL_MAXABS
    # float TIFF output carries the same samples as PFM
    import ctypes as C
    _run(ours_flt, "-i", d / "n0.pfm", "-s", sigma, "--f2_p", 0, "--flt11", d / "g_a1.tif")
    from PIL import Image  # multi-channel float TIFF: read back with our own codec instead
    so = C.CDLL(os.path.join(ROOT, "bwd_nlkalman_b200", "libnlk_image_io.so"))
    so.nlk_read_image.restype = C.POINTER(C.c_float)
    w, h, c = C.c_int(), C.c_int(), C.c_int()
    p = so.nlk_read_image(str(d / "g_a1.tif").encode(), C.byref(w), C.byref(h), C.byref(c))
    tif = np.ctypeslib.as_array(p, shape=(h.value, w.value, c.value)).copy()
    # two separate runs: the aggregation's floating-point atomics commute only up to rounding
    assert maxabs(tif, _read_pfm(d / "g_a1.pfm")) <= TOL_MAXABS


def test_reference_drivers_linked_against_the_library(scene):
    """The drop-in claim itself: the reference's OWN drivers (src/main-flt.c, src/main-smo.c, unmodified,
    with its iio and argparse; oracle/Makefile target `dropin`) linked against libnlkalman_b200.so
    instead of src/nlkalman.c, against the all-reference programs on the same files and command lines
    (scripts/nlkalman-seq.sh:39-41, :80-81, :100-102, :147-149)."""
    d, sigma = scene
    di_flt, di_smo = os.path.join(REF, "nlkalman-flt-dropin"), os.path.join(REF, "nlkalman-smo-dropin")
    ref_flt, ref_smo = os.path.join(REF, "nlkalman-flt-ref"), os.path.join(REF, "nlkalman-smo-ref")
    if not (os.path.exists(di_flt) and os.path.exists(ref_flt)):
        pytest.skip("oracle/_ref drop-in drivers not built (needs /root/reference)")
    one = {"OMP_NUM_THREADS": "1", "OMP_THREAD_LIMIT": "1"}
    _run(ref_flt, "-i", d / "n0.pfm", "-s", sigma, "--flt11", d / "R_a1.pfm", "--flt21", d / "R_a2.pfm", env=one)
    _run(di_flt, "-i", d / "n0.pfm", "-s", sigma, "--flt11", d / "D_a1.pfm", "--flt21", d / "D_a2.pfm")
    assert maxabs(_read_pfm(d / "D_a1.pfm"), _read_pfm(d / "R_a1.pfm")) <= TOL_MAXABS
    assert maxabs(_read_pfm(d / "D_a2.pfm"), _read_pfm(d / "R_a2.pfm")) <= TOL_MAXABS
    for exe, tag, env in ((ref_flt, "R", one), (di_flt, "D", None)):
        _run(exe, "-i", d / "n1.pfm", "-s", sigma, "--f2_p", 0, "-o", d / "bflo.flo", "-k", d / "occ.pgm",
             "--flt10", d / "R_a1.pfm", "--flt11", d / f"{tag}_b1.pfm", env=env)
    assert maxabs(_read_pfm(d / "D_b1.pfm"), _read_pfm(d / "R_b1.pfm")) <= TOL_MAXABS
    for exe, tag, env in ((ref_flt, "R", one), (di_flt, "D", None)):
        _run(exe, "-i", d / "n1.pfm", "-s", sigma, "--f1_p", 0, "-o", d / "bflo.flo", "-k", d / "occ.pgm",
             "--flt11", d / "R_b1.pfm", "--flt20", d / "R_a2.pfm", "--flt21", d / f"{tag}_b2.pfm", env=env)
    assert maxabs(_read_pfm(d / "D_b2.pfm"), _read_pfm(d / "R_b2.pfm")) <= TOL_MAXABS
    for exe, tag, env in ((ref_smo, "R", one), (di_smo, "D", None)):
        _run(exe, "--flt1", d / "R_a2.pfm", "--smo0", d / "R_b2.pfm", "-s", sigma, "-o", d / "fflo.flo",
             "-k", d / "occ.pgm", "--smo1", d / f"{tag}_s0.pfm", ok=(1,), env=env)
    assert maxabs(_read_pfm(d / "D_s0.pfm"), _read_pfm(d / "R_s0.pfm")) <= TOL_MAXABS
    # the configuration the library refuses (reference UB, src/nlkalman.c:1699-1730): message + exit 1
    r = _run(di_smo, "--flt1", d / "R_a2.pfm", "--smo0", d / "R_b2.pfm", "-s", sigma, "-o", d / "fflo.flo",
             "--s1_nt", 1, "--smo1", d / "D_bad.pfm", ok=(1,))
    assert "npatches_t" in r.stderr and not os.path.exists(d / "D_bad.pfm")


def test_seq_driver_matches_per_frame_chain(scene):
    """nlkalman-seq keeps every frame in HBM in opponent space; the per-frame chain passes RGB
    files between processes -- same recursion (with --first_f2 1), results equal up to the
    colour round trip"""
    d, sigma = scene
    flt, smo, seq = (os.path.join(BIN, n) for n in ("nlkalman-flt", "nlkalman-smo", "nlkalman-seq"))
    for t in range(3):
        if t:
            for name in ("bflo", "fflo"):
                if not os.path.exists(d / f"{name}{t}.flo"):
                    os.symlink(d / f"{name}.flo", d / f"{name}{t}.flo")
            if not os.path.exists(d / f"occ{t}.pgm"):
                os.symlink(d / "occ.pgm", d / f"occ{t}.pgm")
    for name in ("fflo0.flo",):
        if not os.path.exists(d / name):
            os.symlink(d / "fflo.flo", d / name)
    if not os.path.exists(d / "occ0.pgm"):
        os.symlink(d / "occ.pgm", d / "occ0.pgm")
    # per-frame chain (our own binaries)
    _run(flt, "-i", d / "n0.pfm", "-s", sigma, "--flt11", d / "c1_0.pfm", "--flt21", d / "c2_0.pfm")
    for t in (1, 2):
        _run(flt, "-i", d / f"n{t}.pfm", "-s", sigma, "--f2_p", 0, "-o", d / "bflo.flo", "-k", d / "occ.pgm",
             "--flt10", d / f"c1_{t-1}.pfm", "--flt11", d / f"c1_{t}.pfm")
        _run(flt, "-i", d / f"n{t}.pfm", "-s", sigma, "--f1_p", 0, "-o", d / "bflo.flo", "-k", d / "occ.pgm",
             "--flt11", d / f"c1_{t}.pfm", "--flt20", d / f"c2_{t-1}.pfm", "--flt21", d / f"c2_{t}.pfm")
    shutil.copy(d / "c2_2.pfm", d / "cs_2.pfm")   # (reference scripts/nlkalman-seq.sh:122)
    for t in (1, 0):
        _run(smo, "--flt1", d / f"c2_{t}.pfm", "--smo0", d / f"cs_{t+1}.pfm", "-s", sigma, "-o", d / "fflo.flo",
             "-k", d / "occ.pgm", "--smo1", d / f"cs_{t}.pfm", ok=(1,))
    # the resident driver
    _run(seq, "-i", d / "n%d.pfm", "-f", 0, "-l", 2, "-s", sigma, "--first_f2", 1, "--s1_p", 8,
         "-o", d / "bflo%d.flo", "-k", d / "occ%d.pgm", "--fflow", d / "fflo%d.flo", "--foccl", d / "occ%d.pgm",
         "--filt1", d / "q1_%d.pfm", "--filt2", d / "q2_%d.pfm", "--smoo1", d / "qs_%d.pfm")
    # A 1e-5 perturbation of the state (the RGB round trip) can flip a k-NN near-tie and with it
    # one group, so the comparison is statistical: tiny on average, rare outliers
    def close(a, b):
        diff = np.abs(_read_pfm(a).astype(np.float64) - _read_pfm(b))
        assert diff.mean() <= 2e-4 and (diff > 1e-2).mean() <= 2e-3, (a, diff.mean(), diff.max(), (diff > 1e-2).mean())
    for t in range(3):
        close(d / f"q1_{t}.pfm", d / f"c1_{t}.pfm")
        close(d / f"q2_{t}.pfm", d / f"c2_{t}.pfm")
    for t in (0, 1):
        close(d / f"qs_{t}.pfm", d / f"cs_{t}.pfm")


def _read_flo(path):
    raw = open(path, "rb").read()
    assert raw[:4] == b"PIEH"
    w, h = np.frombuffer(raw[4:12], np.int32)
    return np.frombuffer(raw[12:], np.float32).reshape(h, w, 2)


def test_tvl1flow_against_reference_program(tmp_path):
    """bin/tvl1flow against the reference's own program (oracle/_ref/tvl1flow-ref: lib/tvl1flow/main.c
    unmodified) with the command lines of the pipeline script (scripts/nlkalman-seq.sh:51, :111:
    "NPROC 0 DW 0 0 FSCALE": zeros mean defaults) and with every argument given."""
    from oracle import oracle as O
    ref = os.path.join(REF, "tvl1flow-ref")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/tvl1flow-ref not built (needs /root/reference)")
    nx, ny = 240, 176
    I0, I1 = O.tvl1_frames(nx, ny, seed=21)
    # a colour pair as well: both programs read it as luminance
    rgb = lambda g: np.stack([g, 0.7 * g + 30, 200 - 0.5 * g], -1).astype(np.float32)
    _write_pfm(tmp_path / "a.pfm", I0); _write_pfm(tmp_path / "b.pfm", I1)
    _write_pfm(tmp_path / "ca.pfm", rgb(I0)); _write_pfm(tmp_path / "cb.pfm", rgb(I1))
    cases = [("script", ("a.pfm", "b.pfm"), ("8", "0", "0.25", "0", "0", "1")),
             ("colour", ("ca.pfm", "cb.pfm"), ("2", "0", "0.40", "0", "0", "1")),
             ("explicit", ("a.pfm", "b.pfm"), ("1", "0.2", "0.3", "0.25", "4", "0", "0.6", "3", "0.02", "1")),
             ("defaults", ("a.pfm", "b.pfm"), ())]
    for name, (f0, f1), args in cases:
        ours, theirs = tmp_path / f"ours-{name}.flo", tmp_path / f"ref-{name}.flo"
        if args:
            _run(os.path.join(BIN, "tvl1flow"), tmp_path / f0, tmp_path / f1, ours, *args)
            _run(ref, tmp_path / f0, tmp_path / f1, theirs, *args)
        else:   # default output name, in the working directory
            for exe, dst in ((os.path.join(BIN, "tvl1flow"), ours), (ref, theirs)):
                r = subprocess.run([exe, str(tmp_path / f0), str(tmp_path / f1)], cwd=tmp_path, capture_output=True, text=True)
                assert r.returncode == 0, r.stderr
                os.replace(tmp_path / "flow.flo", dst)
        a, b = _read_flo(ours), _read_flo(theirs)
        e = maxabs(a, b)
        print(f"tvl1flow {name}: max |du| = {e:.2e} px, identical pixels {float(np.mean(a == b)):.4f}")
        # and the reference's OWN program (main.c unmodified) on top of the product library (include/tvl1flow.h)
        dropin = os.path.join(REF, "tvl1flow-dropin")
        if args and os.path.exists(dropin):
            mine = tmp_path / f"dropin-{name}.flo"
            _run(dropin, tmp_path / f0, tmp_path / f1, mine, *args)
            d = _read_flo(mine)
            print(f"tvl1flow-dropin {name}: max |du| = {maxabs(d, b):.2e} px, identical pixels {float(np.mean(d == b)):.4f}")
            assert maxabs(d, b) <= 5e-2
        assert a.shape == (ny, nx, 2) and np.abs(b).max() > 2.0
        assert e <= 5e-2
    r = _run(os.path.join(BIN, "tvl1flow"), ok=(1,))
    assert "Usage" in r.stderr
    _write_pfm(tmp_path / "small.pfm", I0[:50, :60])
    r = _run(os.path.join(BIN, "tvl1flow"), tmp_path / "a.pfm", tmp_path / "small.pfm", tmp_path / "x.flo", ok=(1,))
    assert "size mismatch" in r.stderr


def _read_image(path):
    """any format the drivers write, through the drivers' own codec (libnlk_image_io.so)"""
    import ctypes as C
    L = C.CDLL(os.path.join(ROOT, "bwd_nlkalman_b200", "libnlk_image_io.so"))
    L.nlk_read_image.restype = C.POINTER(C.c_float)
    L.nlk_read_image.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    w, h, c = C.c_int(), C.c_int(), C.c_int()
    p = L.nlk_read_image(str(path).encode(), C.byref(w), C.byref(h), C.byref(c))
    assert p, path
    a = np.ctypeslib.as_array(p, shape=(h.value, w.value, c.value)).copy()
    C.CDLL(None).free(C.cast(p, C.c_void_p))
    return a


def test_seq_driver_whole_pipeline_matches_the_script(tmp_path):
    """nlkalman-seq --tvl1 1 is the pipeline script in one process: flows (TV-L1) and occlusion masks
    computed on the GPU between resident frames.  Against bin/nlkalman-seq.sh driving the per-frame
    programs (tvl1flow, nlkalman-occ, nlkalman-flt, nlkalman-smo) through files: same flows, same
    masks, same frames, up to the effects of the RGB round trip of the state (see above)."""
    from bwd_nlkalman_b200 import synth
    w, h, ch, sigma, nf = 192, 144, 3, 10.0, 4
    for t in range(nf):
        _write_pfm(tmp_path / f"n{t}.pfm", synth.noisy_frame(w, h, ch, t, sigma))
    a, b = tmp_path / "script", tmp_path / "seq"
    os.makedirs(b)
    r = subprocess.run(["bash", os.path.join(BIN, "nlkalman-seq.sh"), str(tmp_path / "n%d.pfm"), "0", str(nf - 1), str(sigma), str(a)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    _run(os.path.join(BIN, "nlkalman-seq"), "-i", tmp_path / "n%d.pfm", "-f", 0, "-l", nf - 1, "-s", sigma, "--first_f2", 1,
         "--s1_p", 8, "--tvl1", 1, "-o", b / "bflo1-%03d.flo", "-k", b / "bocc1-%03d.png", "--fflow", b / "fflo-%03d.flo",
         "--foccl", b / "focc-%03d.png", "--filt1", b / "flt1-%03d.tif", "--filt2", b / "flt2-%03d.tif",
         "--smoo1", b / "smo1-%03d.tif")
    mot = np.array(synth.MOTION, np.float64)
    for t in range(1, nf):
        fa, fb = _read_flo(a / f"bflo1-{t:03d}.flo"), _read_flo(b / f"bflo1-{t:03d}.flo")
        d = np.abs(fa.astype(np.float64) - fb)
        print(f"backward flow {t}: median |d| {np.median(d):.2e}, max {d.max():.2e}; median flow {np.median(fb, (0, 1))}")
        assert np.median(d) <= 1e-3 and (d > 0.05).mean() <= 0.02
        # it is the motion of the scene (frame t to frame t-1)
        assert np.abs(np.median(fb, (0, 1)) + mot).max() < 0.4
        ma, mb = _read_image(a / f"bocc1-{t:03d}.png"), _read_image(b / f"bocc1-{t:03d}.png")
        assert (ma != mb).mean() <= 0.01
    for t in range(nf - 1):
        fa, fb = _read_flo(a / f"fflo-{t:03d}.flo"), _read_flo(b / f"fflo-{t:03d}.flo")
        d = np.abs(fa.astype(np.float64) - fb)
        assert np.median(d) <= 1e-3 and (d > 0.05).mean() <= 0.02
    for t in range(nf):
        for name in ("flt1", "flt2", "smo1"):
            if name == "smo1" and t == nf - 1:
                continue     # (the script copies flt2 there, scripts/nlkalman-seq.sh:122; the program writes what it computes)
            x, y = _read_image(a / f"{name}-{t:03d}.tif"), _read_image(b / f"{name}-{t:03d}.tif")
            diff = np.abs(x.astype(np.float64) - y)
            print(f"{name} {t}: mean |d| {diff.mean():.2e}, max {diff.max():.2e}, > 1e-2: {(diff > 1e-2).mean():.2e}")
            assert diff.mean() <= 2e-3 and (diff > 0.1).mean() <= 1e-2, (name, t)

"""CPU tests of the host side: the self-contained file codecs of the drivers
(bwd_nlkalman_b200/host/nlk_image_io.c, standing where the reference calls iio) against
Pillow / numpy, and the command-line surface of the drivers on the paths that need no
GPU (option table, mode rules and messages of reference src/main-flt.c:129-149,
src/main-smo.c:87-92)."""
import ctypes as C
import os
import re
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "bwd_nlkalman_b200")
BIN = os.path.join(PKG, "bin")


@pytest.fixture(scope="module")
def io():
    so = os.path.join(PKG, "libnlk_image_io.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", os.path.join(PKG, "host"), "../libnlk_image_io.so"], check=True)
    L = C.CDLL(so)
    L.nlk_read_image.restype = C.POINTER(C.c_float)
    L.nlk_read_image.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.nlk_write_image.argtypes = [C.c_char_p, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int]
    L.nlk_io_error.restype = C.c_char_p
    L.nlk_read_image_gray.restype = C.POINTER(C.c_float)
    L.nlk_read_image_gray.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]

    class IO:
        def read(self, path):
            w, h, c = C.c_int(), C.c_int(), C.c_int()
            p = L.nlk_read_image(str(path).encode(), C.byref(w), C.byref(h), C.byref(c))
            if not p:
                raise IOError(L.nlk_io_error().decode())
            a = np.ctypeslib.as_array(p, shape=(h.value, w.value, c.value)).copy()
            libc.free(p)
            return a

        def read_gray(self, path):
            w, h = C.c_int(), C.c_int()
            p = L.nlk_read_image_gray(str(path).encode(), C.byref(w), C.byref(h))
            if not p:
                raise IOError(L.nlk_io_error().decode())
            a = np.ctypeslib.as_array(p, shape=(h.value, w.value)).copy()
            libc.free(p)
            return a

        def write(self, path, a):
            a = np.ascontiguousarray(a, np.float32)
            if a.ndim == 2:
                a = a[..., None]
            rc = L.nlk_write_image(str(path).encode(), a.ctypes.data_as(C.POINTER(C.c_float)),
                                   a.shape[1], a.shape[0], a.shape[2])
            if rc:
                raise IOError(L.nlk_io_error().decode())
    return IO()


@pytest.mark.parametrize("ch", [1, 2, 3, 4])
def test_tiff_float_roundtrip_and_pillow_reads_it(io, tmp_path, ch):
    rng = np.random.default_rng(ch)
    a = rng.normal(100, 80, (13, 17, ch)).astype(np.float32)
    a[0, 0, 0] = np.nan
    p = tmp_path / "x.tif"
    io.write(p, a)
    b = io.read(p)
    assert b.shape == a.shape and np.array_equal(np.isnan(a), np.isnan(b))
    assert np.array_equal(np.nan_to_num(a), np.nan_to_num(b))
    if ch == 1:  # Pillow understands single-channel float TIFF
        from PIL import Image
        q = np.asarray(Image.open(p))
        assert np.array_equal(np.nan_to_num(q), np.nan_to_num(a[..., 0]))


@pytest.mark.parametrize("compression", [None, "tiff_lzw", "tiff_adobe_deflate", "packbits"])
def test_tiff_written_by_libtiff(io, tmp_path, compression):
    """float TIFFs as the reference writes them through libtiff: LZW below 2000x2000 pixels,
    uncompressed above (reference lib/iio/iio.c:3022-3026), several strips"""
    from PIL import Image
    rng = np.random.default_rng(7)
    a = rng.normal(120, 60, (301, 211)).astype(np.float32)
    a[100:140] = 7.0  # long runs
    p = tmp_path / "f.tif"
    kw = {"compression": compression} if compression else {}
    Image.fromarray(a, mode="F").save(p, **kw)
    assert np.array_equal(io.read(p)[..., 0], a)
    # 8-bit RGB and 16-bit gray, with the horizontal predictor where libtiff applies it
    rgb = rng.integers(0, 256, (64, 50, 3), dtype=np.uint8)
    Image.fromarray(rgb).save(p, **kw)
    assert np.array_equal(io.read(p), rgb.astype(np.float32))
    g16 = rng.integers(0, 65536, (40, 33), dtype=np.uint16)
    Image.fromarray(g16).save(p, **kw)
    assert np.array_equal(io.read(p)[..., 0], g16.astype(np.float32))
    if compression == "tiff_lzw":
        Image.fromarray(rgb).save(p, compression="tiff_lzw", tiffinfo={317: 2})
        assert np.array_equal(io.read(p), rgb.astype(np.float32))


def test_png_masks(io, tmp_path):
    """occlusion masks are 8-bit PNGs of 0 / 255 (reference scripts/nlkalman-seq.sh:68-73)"""
    from PIL import Image
    rng = np.random.default_rng(1)
    m = ((rng.uniform(0, 1, (97, 131)) > 0.8) * 255).astype(np.uint8)
    p = tmp_path / "m.png"
    Image.fromarray(m).save(p)
    assert np.array_equal(io.read(p)[..., 0], m.astype(np.float32))
    rgb = rng.integers(0, 256, (31, 45, 3), dtype=np.uint8)
    Image.fromarray(rgb).save(p)
    assert np.array_equal(io.read(p), rgb.astype(np.float32))
    Image.fromarray(m > 0).save(p)  # 1-bit
    assert np.array_equal(io.read(p)[..., 0], (m > 0).astype(np.float32))
    g16 = rng.integers(0, 65536, (20, 21), dtype=np.uint16)
    Image.fromarray(g16).save(p)
    assert np.array_equal(io.read(p)[..., 0], g16.astype(np.float32))
    # our writer: rounded, clamped to 0..255, readable by Pillow
    f = rng.normal(128, 100, (33, 29, 3)).astype(np.float32)
    io.write(p, f)
    assert np.array_equal(np.asarray(Image.open(p)), np.clip(np.rint(f), 0, 255).astype(np.uint8))


def test_pfm_flo_pnm_follow_iio(io, tmp_path):
    rng = np.random.default_rng(2)
    a = rng.normal(0, 50, (9, 11, 3)).astype(np.float32)
    p = tmp_path / "a.pfm"
    io.write(p, a)
    raw = open(p, "rb").read()
    # iio layout: "PF\n11 9\n-1\n" then rows top to bottom (reference lib/iio/iio.c:3124-3138)
    assert raw.startswith(b"PF\n11 9\n-1\n") and raw[len(b"PF\n11 9\n-1\n"):] == a.tobytes()
    assert np.array_equal(io.read(p), a)
    fl = rng.normal(0, 3, (9, 11, 2)).astype(np.float32)
    p = tmp_path / "f.flo"
    io.write(p, fl)
    raw = open(p, "rb").read()
    assert raw[:4] == b"PIEH" and struct.unpack("<ii", raw[4:12]) == (11, 9) and raw[12:] == fl.tobytes()
    assert np.array_equal(io.read(p), fl)
    g = rng.integers(0, 256, (9, 11, 1)).astype(np.float32)
    p = tmp_path / "g.pgm"
    io.write(p, g)
    assert np.array_equal(io.read(p), g)
    open(tmp_path / "t.pgm", "w").write("P2\n# comment\n3 2\n255\n1 2 3\n4 5 6\n")
    assert np.array_equal(io.read(tmp_path / "t.pgm")[..., 0], np.array([[1, 2, 3], [4, 5, 6]], np.float32))


def test_io_errors(io, tmp_path):
    with pytest.raises(IOError):
        io.read(tmp_path / "missing.tif")
    open(tmp_path / "junk.tif", "wb").write(b"hello world, not an image")
    with pytest.raises(IOError):
        io.read(tmp_path / "junk.tif")
    with pytest.raises(IOError):
        io.write(tmp_path / "x.unknown", np.zeros((2, 2, 1), np.float32))


# ---- command-line surface (no GPU needed on these paths) -----------------------------------------

def _run(tool, *args):
    exe = os.path.join(BIN, tool)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built")
    return subprocess.run([exe, *args], capture_output=True, text=True)


def test_flt_option_table_and_mode_rules():
    r = _run("nlkalman-flt", "-h")
    assert r.returncode == 0
    for opt in ["-i, --nisy", "-o, --bflo", "-k, --bocc", "--flt10", "--flt20", "--flt11", "--flt21", "-s, --sigma",
                "--f1_p", "--f1_sx", "--f1_st", "--f1_nx", "--f1_nt", "--f1_nt_agg", "--f1_bx", "--f1_bt", "--f1_l",
                "--f2_p", "--f2_sx", "--f2_st", "--f2_nx", "--f2_nt", "--f2_nt_agg", "--f2_bx", "--f2_bt", "--f2_l",
                "-v, --verbose"]:
        assert opt in r.stdout, opt
    r = _run("nlkalman-flt", "--f1_p", "0", "--f2_p", "0")
    assert r.returncode == 1 and "nothing to do" in r.stderr
    r = _run("nlkalman-flt", "--f1_p", "0", "--flt21", "o.tif")
    assert r.returncode == 1 and "f1_p == 0 and no input path given" in r.stderr
    r = _run("nlkalman-flt", "-s", "20")
    assert r.returncode == 1 and "no output path given" in r.stderr
    # the last occurrence of an option wins (the scripts append --f2_p 0 after $FPM)
    r = _run("nlkalman-flt", "--f1_p", "8", "--f1_p", "0", "--f2_p=0")
    assert r.returncode == 1 and "nothing to do" in r.stderr
    r = _run("nlkalman-flt", "--nope")
    assert r.returncode == 1 and "unknown option" in r.stderr
    r = _run("nlkalman-flt", "--f1_p", "eight")
    assert r.returncode == 1 and "expects an integer value" in r.stderr
    r = _run("nlkalman-flt", "-s")
    assert r.returncode == 1 and "requires a value" in r.stderr
    r = _run("nlkalman-flt", "-i", "/nonexistent.tif", "--flt11", "o.tif", "-s", "10")
    assert r.returncode == 1 and "Error while openning" in r.stderr


def test_smo_and_seq_option_tables():
    r = _run("nlkalman-smo", "-h")
    assert r.returncode == 0
    for opt in ["--flt1", "--smo0", "-o, --fflo", "-k, --focc", "--smo1", "-s, --sigma", "--s1_p", "--s1_st",
                "--s1_nt", "--s1_nt_agg", "--s1_bt", "--s1_l"]:
        assert opt in r.stdout, opt
    r = _run("nlkalman-smo", "-s", "20")
    assert r.returncode == 1 and "no output path given" in r.stderr
    r = _run("nlkalman-smo", "--smo1", "o.tif", "--s1_p", "0")
    assert r.returncode == 1 and "s1_p == 0" in r.stderr
    r = _run("nlkalman-seq", "-h")
    assert r.returncode == 0
    for opt in ["-i, --nisy", "-o, --bflow", "-k, --boccl", "--fflow", "--foccl", "--filt1", "--filt2", "--smoo1",
                "-f, --first", "-l, --last", "--s1_full"]:
        assert opt in r.stdout, opt
    r = _run("nlkalman-seq", "--f1_p", "0")
    assert r.returncode == 1 and "f1_p == 0" in r.stderr
    # the whole pipeline in one process: flows computed on the GPU, the script's OPM values
    assert "--tvl1" in _run("nlkalman-seq", "-h").stdout and "--of_prms" in _run("nlkalman-seq", "-h").stdout
    r = _run("nlkalman-seq", "-i", "n%d.pfm", "-f", "0", "-l", "1", "-s", "20", "--filt1", "o%d.pfm", "--tvl1", "1",
             "--of_prms", "1 0.25 0.75")
    assert r.returncode == 1 and "six values" in r.stderr


def test_decoders_reject_malformed_headers(io, tmp_path):
    """header fields are not trusted: zero rows per strip, float samples narrower than 32 bits, PNG bit
    depths outside the specification (a 16-bit "palette" image would index past the palette)"""
    import zlib
    from PIL import Image
    a = (np.arange(12 * 9, dtype=np.uint8).reshape(9, 12) * 2)
    tif = tmp_path / "ok.tif"
    Image.fromarray(a).save(tif)
    assert np.array_equal(io.read(tif)[..., 0], a.astype(np.float32))
    raw = bytearray(open(tif, "rb").read())
    assert raw[:2] == b"II"
    ifd = struct.unpack_from("<I", raw, 4)[0]
    n = struct.unpack_from("<H", raw, ifd)[0]

    def patched(tag, value):
        out = bytearray(raw)
        for k in range(n):
            off = ifd + 2 + 12 * k
            if struct.unpack_from("<H", out, off)[0] == tag:
                typ = struct.unpack_from("<H", out, off + 2)[0]
                struct.pack_into("<H" if typ == 3 else "<I", out, off + 8, value)
                return out
        # tag absent: overwrite the last entry (a resolution tag nobody reads)
        off = ifd + 2 + 12 * (n - 1)
        struct.pack_into("<HHII", out, off, tag, 3, 1, value)
        return out
    for tag, value in ((278, 0), (339, 3)):          # RowsPerStrip = 0; SampleFormat = float with 8-bit samples
        bad = tmp_path / f"bad{tag}.tif"
        open(bad, "wb").write(patched(tag, value))
        with pytest.raises(IOError):
            io.read(bad)

    def png(depth, ctype, w=4, h=2):
        def chunk(t, d):
            return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))
        rowb = (w * depth * {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype] + 7) // 8
        body = zlib.compress(b"".join(b"\x00" + bytes(rowb) for _ in range(h)))
        return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0)) +
                (chunk(b"PLTE", bytes(6)) if ctype == 3 else b"") + chunk(b"IDAT", body) + chunk(b"IEND", b""))
    good = tmp_path / "g.png"
    open(good, "wb").write(png(8, 0))
    assert io.read(good).shape == (2, 4, 1)
    for depth, ctype in ((16, 3), (3, 0), (4, 2), (32, 0)):
        bad = tmp_path / f"bad_{depth}_{ctype}.png"
        open(bad, "wb").write(png(depth, ctype))
        with pytest.raises(IOError):
            io.read(bad)


def test_scalar_read_follows_iio(io, tmp_path):
    """the flow estimator reads its images as one channel the way iio_read_image_float does (reference
    lib/iio/iio.c:3984-4003, :1021-1060): luminance computed in the file's sample type.  Checked through
    the reference's own program where it is built: tvl1flow-ref gives the same flow for a colour pair as
    for the grey pair our reader makes of it."""
    rng = np.random.default_rng(5)
    ys, xs = np.mgrid[0:48, 0:64]
    def frame(dx):
        base = 120 + 60 * np.sin((xs - dx) * 0.3) * np.cos(ys * 0.23) + 30 * np.sin((xs - dx) * 0.11 + ys * 0.17)
        rgb = np.stack([base, 0.8 * base + 20, 255 - base], -1) + rng.normal(0, 2, (48, 64, 3))
        return np.clip(rgb, 0, 255)
    a, b = frame(0), frame(1.5)
    lum = lambda x: .299 * x[..., 0] + .587 * x[..., 1] + .114 * x[..., 2]
    for ext, conv in ((".ppm", lambda x: np.floor(x)), (".pfm", lambda x: x.astype(np.float32)), (".png", np.floor)):
        for name, img in (("a", a), ("b", b)):
            io.write(tmp_path / (name + ext), conv(img).astype(np.float32))
    a8 = np.floor(a)
    # PNM and float files: iio holds float samples, the luminance is rounded to float;
    # 8 bit PNG (and 8 / 16 bit TIFF): iio holds integers, the luminance is truncated
    assert np.array_equal(io.read_gray(tmp_path / "a.ppm"), lum(a8).astype(np.float32))
    assert np.array_equal(io.read_gray(tmp_path / "a.png"), np.floor(lum(a8)).astype(np.float32))
    af = a.astype(np.float32).astype(np.float64)
    assert np.array_equal(io.read_gray(tmp_path / "a.pfm"), lum(af).astype(np.float32))
    one = tmp_path / "one.pgm"
    io.write(one, np.floor(lum(a8)).astype(np.float32))
    assert np.array_equal(io.read_gray(one), np.floor(lum(a8)).astype(np.float32))
    two = tmp_path / "two.flo"
    io.write(two, np.zeros((4, 5, 2), np.float32))
    with pytest.raises(IOError):
        io.read_gray(two)
    ref = os.path.join(ROOT, "oracle", "_ref", "tvl1flow-ref")
    if not os.path.exists(ref):
        return
    env = dict(os.environ, OMP_NUM_THREADS="1")
    for ext in (".ppm", ".pfm"):
        for name in ("a", "b"):
            io.write(tmp_path / ("g" + name + ".pfm"), io.read_gray(tmp_path / (name + ext)))
        for src, dst in (((f"a{ext}", f"b{ext}"), "colour.flo"), (("ga.pfm", "gb.pfm"), "grey.flo")):
            r = subprocess.run([ref, tmp_path / src[0], tmp_path / src[1], tmp_path / dst, "1"], capture_output=True,
                               text=True, env=env)
            assert r.returncode == 0, r.stderr
        assert np.array_equal(io.read(tmp_path / "colour.flo"), io.read(tmp_path / "grey.flo")), ext
        assert np.abs(io.read(tmp_path / "colour.flo")).max() > 0.5


def test_tvl1flow_arguments_follow_the_reference_program(io, tmp_path):
    """bin/tvl1flow takes the positional command line of the reference's program (lib/tvl1flow/main.c:73-148):
    missing and out-of-range values become the defaults (the pipeline script passes 0 for tau, theta and
    nscales), nscales is capped by the image size.  With verbose = 1 both programs print what they will
    run with before computing anything: compared here (the computation itself needs the GPU)."""
    rng = np.random.default_rng(3)
    for name in ("a", "b"):
        io.write(tmp_path / f"{name}.pfm", rng.uniform(0, 255, (60, 80)).astype(np.float32))
    ours = os.path.join(BIN, "tvl1flow")
    ref = os.path.join(ROOT, "oracle", "_ref", "tvl1flow-ref")
    cases = [("8", "0", "0.40", "0", "0", "1", "0", "0", "0", "1"),            # the script's, then zeros, verbose
             ("1", "0.2", "0.3", "0.25", "2", "0", "0.6", "3", "0.02", "1"),   # everything given
             ("0", "0.3", "-1", "-2", "-5", "7", "1.0", "-1", "-0.5", "1")]    # everything out of range
    want_values = [dict(tau=0.25, lam=0.4, theta=0.3, zfactor=0.5, nwarps=5, epsilon=0.01),
                   dict(tau=0.2, lam=0.3, theta=0.25, zfactor=0.6, nwarps=3, epsilon=0.02),
                   dict(tau=0.25, lam=0.15, theta=0.3, zfactor=0.5, nwarps=5, epsilon=0.01)]
    pat = re.compile(r"tau=(\S+) lambda=(\S+) theta=(\S+) nscales=(\d+) zfactor=(\S+) nwarps=(\d+) epsilon=(\S+)")
    for args, want in zip(cases, want_values):
        r = subprocess.run([ours, str(tmp_path / "a.pfm"), str(tmp_path / "b.pfm"), str(tmp_path / "o.flo"), *args],
                           capture_output=True, text=True)
        assert r.returncode in (0, 2), r.stderr      # 2: no CUDA device here
        m = pat.search(r.stderr)
        assert m, r.stderr
        tau, lam, theta, nscales, zf, nw, eps = m.groups()
        got = dict(tau=float(tau), lam=float(lam), theta=float(theta), zfactor=float(zf), nwarps=int(nw), epsilon=float(eps))
        assert all(abs(got[k] - want[k]) < 1e-6 for k in want), (got, want)
        warned = sorted(re.findall(r"warning: (\w+) changed", r.stderr))
        if os.path.exists(ref):
            q = subprocess.run([ref, str(tmp_path / "a.pfm"), str(tmp_path / "b.pfm"), str(tmp_path / "r.flo"), *args],
                               capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS="1"))
            assert q.returncode == 0, q.stderr
            assert pat.search(q.stderr).groups() == m.groups(), (q.stderr, r.stderr)
            assert sorted(w for w in re.findall(r"warning: (\w+) changed", q.stderr) if w != "nproc") == warned
    r = subprocess.run([ours], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.startswith("Usage:") and "zfactor nwarps epsilon verbose" in r.stderr

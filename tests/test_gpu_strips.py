"""Strip-sharded pass on the GPU: N strips of one frame, each through its own
context and the nlk_strip_* C ABI, exchanges served in-process (strips.run_virtual), against
the single-context recursion on the same inputs and against the CPU restatement.  The
multi-process NCCL driver (strips.run_dist) is covered when more than one GPU is visible."""
import os
import subprocess
import sys

import numpy as np
import pytest

from common import TOL_MAXABS, maxabs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _scene(w, h, ch, sigma, nframes):
    from bwd_nlkalman_b200 import synth
    frames = [synth.noisy_frame(w, h, ch, t, sigma) for t in range(nframes)]
    occ = np.zeros((h, w), np.float32)
    occ[h // 3:h // 3 + 14, w // 2:w // 2 + 20] = 255
    return frames, synth.backward_flow(w, h), synth.forward_flow(w, h), occ


@pytest.mark.parametrize("transport", ["peer2", "peer", "nccl", "peer2-bigflow"])
@pytest.mark.parametrize("shape,nranks", [((101, 150, 3), 3), ((122, 96, 1), 2), ((90, 200, 3), 4)])
def test_virtual_strips_match_single_context(nlk, shape, nranks, transport):
    """peer2 = peer transport with the two filterings of a frame on two lanes (streams); bigflow = a flow
    that reaches far beyond a strip's halo, so that the warp pulls rows from other strips' slabs"""
    bigflow = transport.endswith("-bigflow")
    transport = transport.replace("-bigflow", "")
    import torch
    from bwd_nlkalman_b200 import strips
    w, h, ch = shape
    lanes = 2 if transport == "peer2" else 1
    transport = "peer" if transport == "peer2" else transport
    sigma, nframes = 20.0, 5 if lanes == 2 else 3
    frames, bflo, fflo, occ = _scene(w, h, ch, sigma, nframes)
    if bigflow:
        bflo, fflo = bflo.copy(), fflo.copy()
        bflo[h // 4:h // 2, :, 1] += 0.45 * h        # these rows look two strips further down
        fflo[h // 2:3 * h // 4, :, 1] -= 0.4 * h     # ... and these two strips up
    f1, f2, s1 = (nlk.default_params(sigma, m) for m in (nlk.FLT1, nlk.FLT2, nlk.SMO1))
    dev = torch.device("cuda", 0)
    up = lambda a: torch.from_numpy(a).to(dev)
    d_frames, d_bflo, d_fflo, d_occ = [up(f) for f in frames], up(bflo), up(fflo), up(occ)

    # single context: forward recursion, then smoother backwards
    ref1, ref2, refs = [], [], [None] * nframes
    with nlk.Context(w, h, ch) as ctx:
        o1, o2 = torch.empty_like(d_frames[0]), torch.empty_like(d_frames[0])
        for t in range(nframes):
            ctx.seq_filter_dev(d_frames[t], d_bflo if t else None, d_occ if t else None, sigma, f1, f2, o1, o2)
            ctx.sync()
            ref1.append(o1.cpu().numpy().copy())
            ref2.append(o2.cpu().numpy().copy())
        ctx.seq_smooth_start_dev(up(ref2[-1]))
        refs[-1] = ref2[-1]
        for t in range(nframes - 2, -1, -1):
            ctx.seq_smooth_dev(up(ref2[t]), d_fflo, d_occ, sigma, s1, o1)
            ctx.sync()
            refs[t] = o1.cpu().numpy().copy()

    ranks = [strips.StripRank(w, h, ch, r, nranks, 0, transport=transport, lanes=lanes) for r in range(nranks)]
    if transport == "peer":
        strips.bind_virtual(ranks)
    try:
        # (distinct output buffers per frame: with two lanes the second filtering of a frame is still
        # running when the next frame is queued)
        outs1 = [torch.zeros_like(d_frames[0]) for _ in ranks]
        outs2 = [torch.zeros_like(d_frames[0]) for _ in ranks]

        def assemble(outs, plans):
            full = np.empty((h, w, ch), np.float32)
            for r, p in enumerate(plans):
                full[p.oy0:p.oy1] = outs[r][p.oy0:p.oy1].cpu().numpy()
            return full
        pl1, pl2, pls = (ranks[0].plans(0, f1), ranks[0].plans(0, f2), ranks[0].plans(1, s1))
        for t in range(nframes):
            strips.run_virtual(ranks, [rk.filter_step(d_frames[t], d_bflo if t else None, d_occ if t else None,
                                                      sigma, f1, f2, outs1[r], outs2[r])
                                       for r, rk in enumerate(ranks)])
            for rk in ranks:
                rk.ctx.sync()
            assert maxabs(assemble(outs1, pl1), ref1[t]) <= TOL_MAXABS, f"flt1 frame {t}"
            assert maxabs(assemble(outs2, pl2), ref2[t]) <= TOL_MAXABS, f"flt2 frame {t}"
        last = up(ref2[-1])
        strips.run_virtual(ranks, [rk.smooth_start(last) for rk in ranks])
        for t in range(nframes - 2, -1, -1):
            flt = up(ref2[t])
            strips.run_virtual(ranks, [rk.smooth_step(flt, d_fflo, d_occ, sigma, s1, outs1[r])
                                       for r, rk in enumerate(ranks)])
            for rk in ranks:
                rk.ctx.sync()
            assert maxabs(assemble(outs1, pls), refs[t]) <= TOL_MAXABS, f"smoother frame {t}"
        if transport == "peer":
            assert all(rk.ctx.peer_error() == 0 for rk in ranks), "a device-side wait timed out"
    finally:
        for rk in ranks:
            rk.close()


def test_virtual_strips_match_oracle(nlk, port):
    """one temporal first-filtering pass, 3 strips, against the CPU restatement"""
    import torch
    from bwd_nlkalman_b200 import strips
    from oracle import oracle as O
    w, h, ch, sigma, nranks = 96, 140, 3, 20.0, 3
    frames, bflo, _, occ = _scene(w, h, ch, sigma, 2)
    f1 = nlk.default_params(sigma, nlk.FLT1)
    pf1 = O.Params(*[getattr(f1, f) for f, _ in nlk.Params._fields_])
    n0, n1 = port.rgb2opp(frames[0].copy()), port.rgb2opp(frames[1].copy())
    a = port.filter_frame(n0, None, None, sigma, pf1)
    wa = port.warp_bicubic(a, bflo, occ)
    want = port.filter_frame(n1, wa, None, sigma, pf1)
    dev = torch.device("cuda", 0)
    ranks = [strips.StripRank(w, h, ch, r, nranks, 0) for r in range(nranks)]
    try:
        d_n1, d_wa = torch.from_numpy(n1).to(dev), torch.from_numpy(wa).to(dev)
        outs = [rk.frame() for rk in ranks]
        strips.run_virtual(ranks, [rk.strip_pass(0, outs[r], d_n1, d_wa, None, sigma, f1) for r, rk in enumerate(ranks)])
        for r, rk in enumerate(ranks):
            rk.ctx.sync()
            # after the final gather every rank holds the whole frame
            assert maxabs(outs[r].cpu().numpy(), want) <= TOL_MAXABS
    finally:
        for rk in ranks:
            rk.close()


@pytest.mark.parametrize("transport", ["peer", "nccl"])
def test_strips_two_gpus(nlk, transport):
    """one process per GPU: slabs mapped across processes with CUDA IPC and the exchanges done by the
    kernels over NVLink (peer), or torch.distributed over NCCL; needs two GPUs on the box"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU visible")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tools", "strips_check.py"), "--w", "160", "--h", "200", "--frames", "3",
           "--transport", transport]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "strips_check OK" in res.stdout

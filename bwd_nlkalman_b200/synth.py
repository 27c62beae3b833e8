"""Seeded synthetic video used by the tests and by bench.py.

Clean frame t = sum of 24 random 2-D sinusoids + a 20-grey-level checkerboard
(37x29 px cells) + 128, clipped to [0, 255], translated by (+1.5, -0.75) px per
frame.  Noise is sigma * N(0, 1), i.i.d., not clipped, fp32.  The backward flow is
the constant (-1.5, +0.75), the forward flow (+1.5, -0.75); the occlusion mask is
0 (valid, reference src/nlkalman.c:77) except a 60x40 rectangle of 255 at
(w/2, h/3).  Images are float32 HWC, the layout of the reference boundary
(reference lib/iio/iio.h:36-37).
"""
from __future__ import annotations

import numpy as np

SCENE_SEED = 7
NOISE_SEED = 1234
MOTION = (1.5, -0.75)  # px / frame (x, y)


def _scene_params(ch: int, seed: int = SCENE_SEED):
    rng = np.random.default_rng(seed)
    n = 24
    amp = rng.uniform(5.0, 25.0, n)
    fx = rng.uniform(-0.15, 0.15, n)
    fy = rng.uniform(-0.15, 0.15, n)
    ph = rng.uniform(0.0, 2.0 * np.pi, n)
    gain = rng.uniform(0.3, 1.0, (n, 3))
    return amp, fx, fy, ph, gain[:, :ch]


def clean_frame(w: int, h: int, ch: int, t: int, seed: int = SCENE_SEED) -> np.ndarray:
    """Clean frame t, float32 (h, w, ch)."""
    amp, fx, fy, ph, gain = _scene_params(ch, seed)
    # content moves by MOTION per frame: frame_t(x) = frame_0(x - t*MOTION)
    xs = np.arange(w, dtype=np.float64)[None, :] - t * MOTION[0]
    ys = np.arange(h, dtype=np.float64)[:, None] - t * MOTION[1]
    out = np.zeros((h, w, ch), dtype=np.float64)
    for i in range(len(amp)):
        s = amp[i] * np.sin(fx[i] * xs + fy[i] * ys + ph[i])
        out += s[:, :, None] * gain[i][None, None, :]
    cx = np.floor(xs / 37.0).astype(np.int64)
    cy = np.floor(ys / 29.0).astype(np.int64)
    out += (20.0 * ((cx + cy) & 1))[:, :, None]
    out += 128.0
    return np.clip(out, 0.0, 255.0).astype(np.float32)


def noisy_frame(w: int, h: int, ch: int, t: int, sigma: float,
                seed: int = SCENE_SEED, noise_seed: int = NOISE_SEED) -> np.ndarray:
    rng = np.random.default_rng(noise_seed + 7919 * t)
    noise = rng.standard_normal((h, w, ch), dtype=np.float32)
    return clean_frame(w, h, ch, t, seed) + np.float32(sigma) * noise


def noisy_frame_cuda(w: int, h: int, ch: int, t: int, sigma: float, device, seed: int = SCENE_SEED,
                     noise_seed: int = NOISE_SEED):
    """noisy_frame evaluated with torch on `device` (float64 scene, the same NumPy noise stream): the
    4K frames of the strip benchmark without seconds of host trigonometry.  Returns a float32 tensor."""
    import torch
    amp, fx, fy, ph, gain = _scene_params(ch, seed)
    xs = torch.arange(w, dtype=torch.float64, device=device)[None, :] - t * MOTION[0]
    ys = torch.arange(h, dtype=torch.float64, device=device)[:, None] - t * MOTION[1]
    out = torch.zeros((h, w, ch), dtype=torch.float64, device=device)
    g = torch.from_numpy(np.ascontiguousarray(gain)).to(device)
    for i in range(len(amp)):
        s = amp[i] * torch.sin(fx[i] * xs + fy[i] * ys + ph[i])
        out += s[:, :, None] * g[i][None, None, :]
    cx = torch.floor(xs / 37.0).to(torch.int64)
    cy = torch.floor(ys / 29.0).to(torch.int64)
    out += (20.0 * ((cx + cy) & 1))[:, :, None]
    out += 128.0
    clean = torch.clamp(out, 0.0, 255.0).to(torch.float32)
    rng = np.random.default_rng(noise_seed + 7919 * t)
    noise = torch.from_numpy(rng.standard_normal((h, w, ch), dtype=np.float32)).to(device)
    return clean + np.float32(sigma) * noise


def backward_flow(w: int, h: int) -> np.ndarray:
    """Flow from frame t to frame t-1, float32 (h, w, 2), [..., 0] = dx."""
    f = np.empty((h, w, 2), dtype=np.float32)
    f[..., 0] = -MOTION[0]
    f[..., 1] = -MOTION[1]
    return f


def forward_flow(w: int, h: int) -> np.ndarray:
    f = np.empty((h, w, 2), dtype=np.float32)
    f[..., 0] = MOTION[0]
    f[..., 1] = MOTION[1]
    return f


def occlusion_mask(w: int, h: int) -> np.ndarray:
    """float32 (h, w): 0 = valid, 255 = occluded."""
    m = np.zeros((h, w), dtype=np.float32)
    x0, y0 = w // 2, h // 3
    m[y0:min(y0 + 40, h), x0:min(x0 + 60, w)] = 255.0
    return m


def psnr(a: np.ndarray, b: np.ndarray) -> float:
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return float("inf") if mse == 0 else 10.0 * np.log10(255.0 ** 2 / mse)

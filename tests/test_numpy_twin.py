"""The C oracle against a second, independently written restatement (oracle/numpy_twin.py) of the part of the path
the reference's own tests pin nothing of: spec_to_grey, image-0.23 Lanczos3 resize WITH its per-pass clamp, and the
colour map (display.rs:24-61).  Content is chosen so that the clamp bites (white noise: negative lobes), so that a
misreading of the clamp order, the tap windows or the normalisation would show.  Plus an outside check of the mel
filterbank against torchaudio (Slaney scale, no area norm) re-normalised by the sum as mel.rs:80-82 does."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy_twin as twin  # noqa: E402


@pytest.mark.parametrize("shape,new", [((37, 61), (90, 50)), ((64, 200), (200, 64)), ((120, 33), (40, 100)), ((347, 257), (256, 500)),
                                       ((5, 7), (3, 2)), ((1, 9), (5, 1)), ((50, 400), (117, 23))])
def test_resize_with_clamp_matches_the_twin(orc, shape, new):
    h, w = shape
    nw, nh = new
    rng = np.random.default_rng(h * 1000 + w)
    grey = rng.random((h, w), dtype=np.float32)          # white noise in [0, 1): Lanczos lobes undershoot -> the clamp bites
    grey[rng.random((h, w)) < 0.3] = 0.0                 # ... especially next to zeros
    a = orc.resize_lanczos3(grey, nw, nh)
    b = twin.resize_lanczos3(grey, nw, nh)
    assert a.shape == b.shape == (nh, nw)
    # without the clamp after the first pass a good part of the image would differ by > 1e-2 (checked below)
    assert np.abs(a - b).max() <= 2e-6, np.abs(a - b).max()
    if nh > h and nw > w and min(h, w) >= 8:
        assert (a == 0).mean() > 0.005   # magnification of noise with zeros: the clamp really was active
    # negative control: one clamp only, at the end, is a different function
    lin = twin._sample_axis(grey, nh)
    unclamped_mid = np.maximum(_no_clamp_pass(_no_clamp_pass(grey, nh).T.copy(), nw).T, 0)
    if min(h, w) >= 30 and h != nh and w != nw:   # both passes really interpolate
        assert np.abs(unclamped_mid - a).max() > 1e-3
    assert (lin >= 0).all()


def _no_clamp_pass(img, n_out):
    """the twin's pass with the clamp removed (negative control only)"""
    F = np.float32
    img = np.asarray(img, F)
    n_in = img.shape[0]
    ratio = F(n_in) / F(n_out); sratio = max(ratio, F(1)); support = F(3) * sratio
    o = np.arange(n_out, dtype=F)
    x = (o + F(0.5)) * ratio
    left = np.clip(np.floor(x - support).astype(np.int64), 0, n_in - 1)
    right = np.clip(np.ceil(x + support).astype(np.int64), left + 1, n_in)
    x = x - F(0.5)
    out = np.zeros((n_out, img.shape[1]), F); ws = np.zeros(n_out, F)
    for j in range(int((right - left).max())):
        i = left + j
        wgt = np.where(i < right, twin._lanczos3((i.astype(F) - x) / sratio), F(0)).astype(F)
        out += wgt[:, None] * img[np.minimum(i, n_in - 1)]; ws += wgt
    return out / ws[:, None]


def test_colour_map_matches_the_twin(orc):
    x = np.concatenate([np.linspace(0, 1.2, 6001, dtype=np.float32), np.float32([0.0, 0.1, 0.899999, 0.9, 1.0, 5.0]),
                        (np.arange(0, 10, dtype=np.float32) + np.float32(0.5)) / np.float32(10)])
    a = np.stack([orc.convert_grey_to_color(float(v)) for v in x])
    b = twin.convert_grey_to_color(x)
    assert np.array_equal(a, b)


def test_spec_to_grey_and_pixels_match_the_twin(orc):
    rng = np.random.default_rng(7)
    spec = (rng.standard_normal((150, 90)).astype(np.float32) * 25 - 70)
    for up in (1.0, 1.37, 2.6):
        a = orc.spec_to_grey(spec, up, -20.0, -140.0)
        b = twin.spec_to_grey(spec, up, -20.0, -140.0)
        assert a.shape == b.shape and np.array_equal(a, b)
    grey = twin.spec_to_grey(spec, 1.37, -20.0, -140.0)
    pa = orc.grey_to_rgb(grey, 333, 200)
    pb = twin.grey_to_rgb(grey, 333, 200)
    d = np.abs(pa.astype(int) - pb.astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3   # a 2e-6 difference in the resampled grey can flip a rounding


def test_mel_bank_against_torchaudio(orc):
    torchaudio = pytest.importorskip("torchaudio")
    for sr, n_fft, n_mel in ((24000, 2048, 80), (48000, 2048, 347), (44100, 1024, 128)):
        ta = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, 0.0, sr / 2.0, n_mel, sr, norm=None, mel_scale="slaney").numpy()
        ta = ta / np.maximum(ta.sum(axis=0, keepdims=True), 1e-30)   # mel.rs:80-82: every filter divided by its sum
        ours = orc.calc_mel_fb(sr, n_fft, n_mel)
        assert ours.shape == ta.shape
        assert np.abs(ours - ta).max() <= 2e-5, np.abs(ours - ta).max()
        assert np.array_equal(ours > 1e-6, ta > 1e-6) or np.abs(ours - ta)[(ours > 1e-6) != (ta > 1e-6)].max() < 1e-5

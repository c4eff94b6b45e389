"""Pins the oracle to the committed golden files generated from the reference's sample audio
(tools/make_golden.py), and -- where /root/reference is mounted -- to the full fixtures."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ["8k", "16k", "22k05", "24k", "44k1"]


@pytest.fixture(scope="module")
def golden():
    clips = np.load(os.path.join(HERE, "golden", "clips.npz"))
    exp = np.load(os.path.join(HERE, "golden", "expected.npz"))
    return clips, exp


def test_oracle_reproduces_golden_specs(orc, golden):
    clips, exp = golden
    for name in NAMES:
        x = clips[f"pcm_{name}"].astype(np.float32) / np.float32(32768.0)
        sr = int(clips[f"sr_{name}"])
        win, hop, n_fft = orc.track_params(sr)
        fb = orc.calc_mel_fb_default(sr, n_fft)
        spec = orc.calc_spec(x, win, hop, n_fft, None, fb)
        assert tuple(exp[f"spec_shape_{name}"]) == spec.shape
        # libm differences between hosts are the only slack allowed
        assert np.max(np.abs(spec[::7, ::5] - exp[f"spec_sub_{name}"])) < 2e-3
        assert abs(float(spec.max()) - float(exp[f"spec_max_{name}"])) < 1e-4
        lin = orc.calc_spec(x, win, hop, n_fft, None, None)
        d = np.abs(lin[::11, ::13] - exp[f"lin_sub_{name}"])
        assert np.max(d[exp[f"lin_sub_{name}"] > -150]) < 2e-3


def test_oracle_reproduces_golden_images(orc, golden):
    clips, exp = golden
    wavs, srs, params, windows, fbs = [], [], [], [], []
    for name in NAMES:
        x = clips[f"pcm_{name}"].astype(np.float32) / np.float32(32768.0)
        sr = int(clips[f"sr_{name}"])
        p = orc.track_params(sr)
        wavs.append(x); srs.append(sr); params.append(p)
        windows.append(orc.calc_window(p[0], p[2])); fbs.append(orc.calc_mel_fb_default(sr, p[2]))
    imgs, mx, mn = orc.pipeline(wavs, srs, params, windows, fbs, px_per_sec=100.0, nheight=120, channels=3)
    assert abs(mx - float(exp["max_db"])) < 1e-4 and abs(mn - float(exp["min_db"])) < 1e-4
    for name, im in zip(NAMES, imgs):
        want = exp[f"img_{name}"]
        assert im.shape == want.shape == (120, 200, 3)
        assert np.max(np.abs(im.astype(int) - want.astype(int))) <= 1


def test_fixture_facts_match_survey():
    """tests/golden/fixtures.json (oracle on the full reference fixtures) vs SURVEY appendix B."""
    with open(os.path.join(HERE, "golden", "fixtures.json")) as f:
        facts = json.load(f)
    assert {k: (v["frames"], v["n_mel"]) for k, v in facts.items() if k != "global"} == {
        "8k": (4404, 257), "16k": (4404, 385), "22k05": (4394, 308), "24k": (4404, 289), "44k1": (4404, 370)}
    assert abs(facts["global"]["max_db"] + 31.6947) < 1e-3
    assert abs(facts["global"]["min_db"] + 151.6947) < 1e-3


@pytest.mark.skipif(not os.path.isdir("/root/reference/samples"), reason="reference fixtures not mounted")
def test_full_fixture_lengths(orc):
    import wave

    with open(os.path.join(HERE, "golden", "fixtures.json")) as f:
        facts = json.load(f)
    for name in NAMES:
        with wave.open(f"/root/reference/samples/sample_{name}.wav") as w:
            assert w.getnframes() == facts[name]["n"] and w.getframerate() == facts[name]["sr"]
            assert orc.stft_n_frames(w.getnframes(), facts[name]["win"], facts[name]["hop"]) == facts[name]["frames"]

"""Host-side tables of libsgx (window, mel filterbank, parameter derivation, frame counts) against the
oracle, bit for bit.  No GPU needed: these functions run on the CPU in the reference too
(lib.rs:143-158) and their outputs are inputs of the kernels."""
import numpy as np
import pytest


def test_known_answers(msv):
    assert np.array_equal(msv.hann(4, False), np.array([0, 0.5, 1, 0.5], np.float32))  # windows.rs:35-38
    assert abs(msv.hz_to_mel(100.0) - 1.5) < 1e-6 and abs(msv.hz_to_mel(1100.0) - 16.38629404765444) < 2e-6  # mel.rs:107-113
    assert abs(msv.mel_to_hz(1.0) - 66.66666666666667) < 1e-5 and abs(msv.mel_to_hz(16.0) - 1071.1702874944676) < 2e-4
    assert msv.calc_proper_n_fft(1920) == 2048 and msv.calc_proper_n_fft(2048) == 2048 and msv.calc_proper_n_fft(2049) == 4096
    assert list(msv.get_colormap()[:6]) == [0, 0, 4, 27, 12, 65]


@pytest.mark.parametrize("n", [1, 2, 4, 5, 320, 884, 1764, 1920, 4096])
def test_hann_and_window_bit_exact(msv, orc, n):
    for sym in (False, True):
        if sym and n == 1:
            continue
        assert np.array_equal(msv.hann(n, sym), orc.hann(n, sym))
    n_fft = msv.calc_proper_n_fft(n)
    assert np.array_equal(msv.calc_window(n, n_fft), orc.calc_window(n, n_fft))


@pytest.mark.parametrize("sr", [8000, 16000, 22050, 24000, 44100, 48000, 96000, 11025])
def test_params_and_default_mel_bit_exact(msv, orc, sr):
    assert msv.track_params(sr) == orc.track_params(sr)
    win, hop, n_fft = msv.track_params(sr)
    a, b = msv.calc_mel_fb_default(sr, n_fft), orc.calc_mel_fb_default(sr, n_fft)
    assert a.shape == b.shape and np.array_equal(a, b)


@pytest.mark.parametrize("sr,n_fft,n_mel", [(24000, 2048, 80), (48000, 4096, 128), (44100, 2048, 128), (8000, 512, 128), (44100, 512, 128), (48000, 64, 10)])
def test_mel_fb_bit_exact(msv, orc, sr, n_fft, n_mel):
    for norm in (True, False):
        assert np.array_equal(msv.calc_mel_fb(sr, n_fft, n_mel, 0.0, None, norm), orc.calc_mel_fb(sr, n_fft, n_mel, 0.0, None, norm))
    assert np.array_equal(msv.calc_mel_fb(sr, n_fft, n_mel, 50.0, sr / 4.0, True), orc.calc_mel_fb(sr, n_fft, n_mel, 50.0, sr / 4.0, True))


def test_settings_overrides(msv):
    s = msv.Settings.default(n_fft=4096, hop_length=256, win_length=4096)
    assert msv.track_params(48000, s) == (4096, 256, 4096)
    s = msv.Settings.default(f_overlap=2)
    assert msv.track_params(48000, s) == (1920, 480, 4096)
    d = msv.Settings.default()
    assert (d.win_ms, d.t_overlap, d.f_overlap, d.freq_scale, d.db_range) == (40.0, 4, 1, msv.FREQ_MEL, 120.0)  # lib.rs:93-99


def test_num_frames_matches_oracle(msv, orc):
    rng = np.random.default_rng(5)
    cases = [(4, 4, 2), (64, 64, 16), (65, 64, 64), (352255, 320, 80), (970902, 884, 221), (10, 4, 1), (3, 4, 2), (100, 7, 3), (100, 9, 2)]
    for _ in range(300):
        win = int(rng.integers(2, 400))
        cases.append((int(rng.integers(1, 3000)), win, int(rng.integers(1, 2 * win))))
    for n, win, hop in cases:
        assert msv.stft_num_frames(n, win, hop) == orc.stft_n_frames(n, win, hop), (n, win, hop)
    # even windows with enough samples: T = n // hop + 1
    assert msv.stft_num_frames(1941805, 1764, 441) == 1941805 // 441 + 1

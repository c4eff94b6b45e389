"""The reference's own known-answer tests, restated against the CPU oracle.
Every test names the Rust test it mirrors.  These pin the oracle; the GPU path is then
compared with the oracle in the -m gpu tests."""
import numpy as np
import pytest


def impulse(n, loc):  # utils.rs:21-34 Impulse
    x = np.zeros(n, np.float32)
    x[loc] = 1.0
    return x


def test_stft_works(orc):
    """lib.rs:491-514"""
    got = orc.perform_stft(impulse(4, 2), 4, 2, 4, None, parallel=False)
    want = np.array([[0, 0, 0], [0.25, -0.25, 0.25], [0.25, -0.25, 0.25]], np.complex64)
    assert got.shape == (3, 3)
    assert np.array_equal(got, want)
    assert np.array_equal(orc.perform_stft(impulse(4, 2), 4, 2, 4, None, parallel=True), want)


def test_hann_window_works(orc):
    """windows.rs:35-38"""
    assert np.array_equal(orc.hann(4, False), np.array([0, 0.5, 1, 0.5], np.float32))


def test_pad_works(orc):
    """utils.rs:125-140"""
    assert np.array_equal(orc.pad_constant([1, 2, 3], 1, 2, 10.0), np.array([10, 1, 2, 3, 10, 10], np.float32))
    assert np.array_equal(orc.pad_reflect([1, 2, 3], 1, 2), np.array([2, 1, 2, 3, 2, 1], np.float32))


def test_rfft_wrapper_works(orc):
    """utils.rs:117-123"""
    assert np.array_equal(orc.rfft_f32(impulse(4, 0)), np.array([1, 1, 1], np.complex64))


def test_real_to_complex(orc):
    """realfft.rs:253-272: RealFFT::<f64>(256) vs the planner's complex FFT, abs 1e-15... the
    reference compares the first 129 bins of a 256-point complex FFT of the real input."""
    x = np.zeros(256, np.float64)
    x[1] = 1.0
    x[3] = 0.5
    got = orc.rfft_f64(x)
    want = orc.cfft_f64(x.astype(np.complex128))[:129]
    assert np.max(np.abs(got - want)) < 1e-13  # f64 twiddles via libm; the Rust test allows 1e-15 per component
    assert np.max(np.abs(got - np.fft.rfft(x))) < 1e-13


def test_mel_hz_convert(orc):
    """mel.rs:107-113 (the reference runs these in f64)"""
    L = orc.lib
    assert abs(L.orc_hz_to_mel_f64(100.0) - 1.5) < 1e-14
    assert abs(L.orc_hz_to_mel_f64(1100.0) - 16.38629404765444) < 1e-14
    assert abs(L.orc_mel_to_hz_f64(1.0) - 66.66666666666667) < 1e-14
    assert abs(L.orc_mel_to_hz_f64(16.0) - 1071.1702874944676) < 1e-12
    # f32 twins agree to f32 precision
    assert abs(L.orc_hz_to_mel(1100.0) - 16.38629404765444) < 2e-6
    assert abs(L.orc_mel_to_hz(16.0) - 1071.1702874944676) < 2e-4


def test_mel_works_is_stale(orc):
    """mel.rs:115-133 pins librosa's Slaney AREA normalisation, but the code divides each filter by
    its SUM (mel.rs:80-82).  The oracle follows the code, so the reference's expected values must NOT
    match (they would need 2/(f[m+2]-f[m])); the un-normalised shape is what both agree on."""
    answer = np.array([0.0, 6.613916251808404922e-03, 1.322783250361680984e-02, 1.984174735844135284e-02,
                       2.105801925063133240e-02, 1.444410253316164017e-02, 7.830185815691947937e-03,
                       1.216269447468221188e-03])
    fb = orc.calc_mel_fb(24000, 2048, 80, 0.0, None, True)
    assert not np.allclose(fb[:8, 0], answer, atol=1e-8)
    raw = orc.calc_mel_fb(24000, 2048, 80, 0.0, None, False)
    ratio = raw[1:8, 0] / answer[1:]
    assert np.allclose(ratio, ratio[0], rtol=2e-4)          # same triangle, different scale
    assert abs(fb[:, 0].sum() - 1.0) < 1e-6                  # sum-normalised, as the code says


@pytest.mark.parametrize("sr", [400, 800, 1000, 2000, 4000, 8000, 16000, 24000, 44100, 48000, 88200, 96000])
def test_mel_default_works(orc, sr):
    """mel.rs:135-165: no empty filter, and one more filter would leave one empty."""
    for e in range(5, 13):
        n_fft = 1 << e
        fb = orc.calc_mel_fb_default(sr, n_fft)
        n_freq, n_mel = fb.shape
        assert n_freq == n_fft // 2 + 1
        assert (fb.sum(axis=0) > 0).all(), (sr, n_fft, n_mel)
        if n_mel == n_freq:
            continue
        more = orc.calc_mel_fb(sr, n_fft, n_mel + 1, 0.0, None, True)
        assert not (more.sum(axis=0) > 0).all(), (sr, n_fft, n_mel)


def test_colormap_and_color(orc):
    """display.rs:10-42: palette bytes; len (not len-1) scaling saturates grey >= 0.9."""
    cm = orc.colormap().reshape(10, 3)
    assert tuple(cm[0]) == (0, 0, 4) and tuple(cm[9]) == (252, 255, 164) and tuple(cm[5]) == (207, 68, 70)
    assert tuple(orc.convert_grey_to_color(0.0)) == (0, 0, 4)
    assert tuple(orc.convert_grey_to_color(0.9)) == (252, 255, 164)
    assert tuple(orc.convert_grey_to_color(1.0)) == (252, 255, 164)
    assert tuple(orc.convert_grey_to_color(5.0)) == (252, 255, 164)
    # half way between stop 0 and 1: round half away from zero
    assert tuple(orc.convert_grey_to_color(0.05)) == (14, 6, 35)
    with pytest.raises(ValueError):
        orc.convert_grey_to_color(-0.1)


def test_track_params(orc):
    """lib.rs:43-46 for the six sample rates of multitrack_works."""
    want = {8000: (320, 80, 512), 16000: (640, 160, 1024), 22050: (884, 221, 1024), 24000: (960, 240, 1024),
            44100: (1764, 441, 2048), 48000: (1920, 480, 2048)}
    for sr, p in want.items():
        assert orc.track_params(sr) == p


def test_framing_closed_form(orc):
    """The three-list construction of lib.rs:412-433 equals frames of reflect-padded input at stride hop."""
    rng = np.random.default_rng(0)
    for n, win, hop, n_fft in [(100, 16, 4, 16), (1000, 64, 16, 64), (997, 884 // 4, 55, 256), (5000, 320, 80, 512),
                               (333, 30, 7, 32), (64, 64, 16, 64), (65, 64, 64, 128)]:
        x = rng.standard_normal(n).astype(np.float32)
        got = orc.perform_stft(x, win, hop, n_fft)
        T = n // hop + 1 if win % 2 == 0 else got.shape[0]
        assert got.shape[0] == T
        w = orc.calc_window(win, n_fft).astype(np.float64)
        p = np.pad(x.astype(np.float64), win // 2, mode="reflect")
        pl = (n_fft - win) // 2
        for t in [0, 1, T // 2, T - 2, T - 1]:
            g = np.zeros(n_fft)
            g[pl:pl + win] = p[t * hop:t * hop + win] * w
            ref = np.fft.rfft(g)
            assert np.max(np.abs(got[t] - ref)) <= 2e-6 * max(1e-6, np.abs(ref).max()) + 1e-9


def test_db_floor_and_range(orc):
    """decibel.rs:79-88 and lib.rs:208-209."""
    x = np.array([0.0, 1e-19, 1e-18, 1.0, 10.0, 0.5], np.float32)
    db = orc.amp_to_db_default(x)
    assert db[0] == -360.0 and db[1] == -360.0 and db[3] == 0.0 and db[4] == 20.0
    assert abs(db[5] - 20 * np.log10(0.5)) < 1e-5
    with pytest.raises(ValueError):
        orc.amp_to_db_default(np.array([-1.0], np.float32))
    assert orc.clamp_range(-31.5, -200.0) == (-31.5, -151.5)
    assert orc.clamp_range(3.0, -50.0) == (0.0, -50.0)


def test_resize_matches_pillow_linear_part(orc):
    """image 0.23's Lanczos3 resize is third-party and not vendored; its linear part (no clamp active)
    must agree with Pillow's F-mode LANCZOS on data that stays positive."""
    PIL = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(3)
    g = (0.4 + 0.2 * rng.random((37, 53))).astype(np.float32)  # no negative lobes large enough to clamp
    for nw, nh in [(80, 50), (30, 20), (53, 37)]:
        got = orc.resize_lanczos3(g, nw, nh)
        ref = np.asarray(PIL.fromarray(g, mode="F").resize((nw, nh), PIL.LANCZOS))
        assert np.max(np.abs(got - ref)) < 2e-4

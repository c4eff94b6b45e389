"""Parity of the CUDA path (through the C ABI) with the CPU oracle.  Tolerances are north_star's:
magnitudes 1e-4 relative (to the frame's peak), dB 1e-3 (above the display floor), pixels +-1 LSB."""
import os

import numpy as np
import pytest

import synth
from parity import (DB_ATOL, MAG_RTOL, PX_LSB, assert_db_close, assert_db_vs_truth, assert_min_db_vs_truth, assert_pixels_close,
                    assert_range_close, db_report)

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ["8k", "16k", "22k05", "24k", "44k1"]


def impulse(n, loc):
    x = np.zeros(n, np.float32)
    x[loc] = 1.0
    return x


def frame_rel_err(got, ref):
    peak = np.maximum(np.abs(ref).max(axis=1, keepdims=True), 1e-30)
    return float((np.abs(got - ref) / peak).max())


def db_err(got, ref):
    return db_report(got, ref)[0]


# ---- reference KATs through the GPU ---------------------------------------------------------------
def test_stft_works_kat(msv):
    """lib.rs:491-514"""
    got = msv.perform_stft(impulse(4, 2), 4, 2, 4)
    want = np.array([[0, 0, 0], [0.25, -0.25, 0.25], [0.25, -0.25, 0.25]], np.complex64)
    assert got.shape == (3, 3)
    assert np.max(np.abs(got - want)) <= 1e-7


def test_rfft_impulse_kat(msv):
    """utils.rs:117-123 rfft(impulse) == all ones: frame 0 of an impulse placed where the left reflect
    puts it at sample 0 of the frame, rectangular window."""
    x = impulse(64, 32)
    got = msv.perform_stft(x, 64, 64, 64, window=np.ones(64, np.float32))
    assert got.shape == (2, 33)
    assert np.max(np.abs(got[0] - 1.0)) <= 1e-6


def test_real_to_complex_kat(msv, orc):
    """realfft.rs:253-272 shape (n_fft=256) against the f64 truth."""
    x = np.zeros(256 * 3, np.float32)
    x[1 + 256] = 1.0
    x[3 + 256] = 0.5
    got = msv.perform_stft(x, 256, 128, 256, window=np.ones(256, np.float32))
    truth = orc.stft_mag_f64(x, 256, 128, 256, window=np.ones(256, np.float32))
    assert np.max(np.abs(np.abs(got) - truth)) <= 2e-6


# ---- STFT / magnitude parity over sizes -------------------------------------------------------------
STFT_CASES = [
    # (n, win, hop, n_fft)
    (4000, 4, 2, 4), (3000, 6, 3, 8), (5000, 30, 7, 32), (9000, 128, 32, 128), (9000, 200, 50, 256),
    (20000, 320, 80, 512), (20000, 512, 128, 512), (30000, 640, 160, 1024), (30000, 884, 221, 1024),
    (60000, 1764, 441, 2048), (60000, 1920, 480, 2048), (60000, 2048, 512, 2048), (90000, 4096, 256, 4096),
    (90000, 3000, 1000, 4096), (150000, 8192, 2048, 8192), (200000, 16384, 4096, 16384), (200000, 10000, 3333, 16384),
    (2048, 2048, 512, 2048),      # shortest legal input: n == win
    (2050, 2048, 2048, 2048),     # hop == win
    (50000, 1024, 5000, 1024),    # hop > n_fft (tile staging cannot be used)
]


@pytest.mark.parametrize("n,win,hop,n_fft", STFT_CASES)
def test_perform_stft_parity(msv, orc, n, win, hop, n_fft):
    x = synth.base_clip(n, 16000, seed=n + win)
    ref = orc.perform_stft(x, win, hop, n_fft)
    got = msv.perform_stft(x, win, hop, n_fft)
    assert got.shape == ref.shape
    err = frame_rel_err(got, ref)
    truth = orc.stft_mag_f64(x, win, hop, n_fft)
    e_gpu = frame_rel_err(np.abs(got).astype(np.float64), truth)
    e_ref = frame_rel_err(np.abs(ref).astype(np.float64), truth)
    print(f"stft n_fft={n_fft} win={win} hop={hop}: gpu-vs-oracle {err:.2e}; vs f64 truth gpu {e_gpu:.2e} oracle {e_ref:.2e}")
    assert err <= MAG_RTOL
    mag = msv.stft_magnitude(x, win, hop, n_fft)
    assert frame_rel_err(mag, np.abs(ref)) <= MAG_RTOL


def test_custom_window_and_mismatch(msv, orc):
    x = synth.base_clip(30000, 16000, seed=9)
    w = (np.hamming(640) / 1024).astype(np.float32)
    assert frame_rel_err(msv.perform_stft(x, 640, 160, 1024, window=w), orc.perform_stft(x, 640, 160, 1024, window=w)) <= MAG_RTOL
    with pytest.raises(msv.SgxError) as e:   # assert_eq!(w.len(), win_length) lib.rs:404
        msv.perform_stft(x, 640, 160, 1024, window=w[:-1])
    assert e.value.code == msv.SGX_ERR_BAD_ARG
    with pytest.raises(msv.SgxError):        # input shorter than the window: the reference panics on its slices
        msv.perform_stft(x[:100], 640, 160, 1024)
    with pytest.raises(msv.SgxError):        # n_fft not a power of two (Radix4)
        msv.perform_stft(x, 600, 150, 1000)


# ---- dB spectrograms ---------------------------------------------------------------------------------
@pytest.mark.parametrize("sr", [8000, 16000, 22050, 24000, 44100, 48000])
def test_default_mel_db_parity(msv, orc, sr):
    """MultiTrack defaults per sample rate (lib.rs:43-46, mel.rs:87-99)."""
    win, hop, n_fft = msv.track_params(sr)
    x = synth.base_clip(3 * sr + 17, sr, seed=sr)
    fb = msv.calc_mel_fb_default(sr, n_fft)
    ref = orc.calc_spec(x, win, hop, n_fft, None, fb)
    got = msv.melspectrogram_db(x, win, hop, n_fft, None, fb)
    truth = orc.calc_spec_f64(x, win, hop, n_fft, None, fb)
    e, m = assert_db_close(got, ref, f"mel sr={sr}")
    assert_db_vs_truth(got, ref, truth, f"mel sr={sr}")
    print(f"mel dB sr={sr} M={fb.shape[1]}: gpu-vs-oracle {e:.2e} dB / {m:.2e} mag; vs truth gpu {db_err(got, truth):.2e} oracle {db_err(ref, truth):.2e}")


@pytest.mark.parametrize("n_fft,hop,n_mel,sr", [(2048, 512, 128, 44100), (4096, 256, 128, 48000), (512, 128, 128, 44100), (1024, 256, 64, 16000),
                                                (8192, 2048, 128, 48000), (16384, 4096, 128, 48000), (256, 64, 40, 8000), (64, 16, 8, 8000)])
def test_fixed_mel_db_parity(msv, orc, n_fft, hop, n_mel, sr):
    x = synth.base_clip(max(4 * n_fft, 2 * sr), sr, seed=n_fft + n_mel)
    fb = msv.calc_mel_fb(sr, n_fft, n_mel)
    ref = orc.calc_spec(x, n_fft, hop, n_fft, None, fb)
    got = msv.melspectrogram_db(x, n_fft, hop, n_fft, None, fb)
    e, m = assert_db_close(got, ref, f"mel-{n_mel} n_fft={n_fft}")
    if not (ref <= -359.0).any():   # C3 shape and friends: anchored on the f64 truth over the whole display range
        assert_db_vs_truth(got, ref, orc.calc_spec_f64(x, n_fft, hop, n_fft, None, fb), f"mel-{n_mel} n_fft={n_fft}")
    print(f"mel-{n_mel} n_fft={n_fft}: {e:.2e} dB / {m:.2e} mag")
    # empty filters (no bin inside) sit on the -360 dB floor in both
    assert np.array_equal(got <= -359.0, ref <= -359.0)


@pytest.mark.parametrize("n_fft", [512, 1024, 2048, 4096, 8192, 16384])
def test_linear_db_parity(msv, orc, n_fft):
    """C4 shape: linear-frequency dB, hop = n_fft/4."""
    x = synth.base_clip(6 * n_fft + 123, 44100, seed=n_fft)
    ref = orc.calc_spec(x, n_fft, n_fft // 4, n_fft, None, None)
    got = msv.melspectrogram_db(x, n_fft, n_fft // 4, n_fft, None, None)
    e, m = assert_db_close(got, ref, f"linear n_fft={n_fft}")
    truth = orc.calc_spec_f64(x, n_fft, n_fft // 4, n_fft, None, None)
    assert_db_vs_truth(got, ref, truth, f"linear n_fft={n_fft}")
    print(f"linear dB n_fft={n_fft}: {e:.2e} dB / {m:.2e} mag; vs truth gpu {db_err(got, truth):.2e} oracle {db_err(ref, truth):.2e}")


def test_silence_hits_the_floor(msv, orc):
    x = np.zeros(20000, np.float32)
    got = msv.melspectrogram_db(x, 640, 160, 1024, None, msv.calc_mel_fb_default(16000, 1024))
    assert np.all(got == -360.0)                       # decibel.rs:49-53 with amin = 1e-18
    assert np.array_equal(msv.amp_to_db_default([0.0, 1e-19, 1.0, 10.0]), orc.amp_to_db_default([0.0, 1e-19, 1.0, 10.0]))
    with pytest.raises(msv.SgxError):                  # decibel.rs:34
        msv.amp_to_db_default([1.0, -1.0])


def test_stereo_is_channel_sum(msv, orc):
    """lib.rs:42 sums channels (no 1/2)."""
    sr = 16000
    l = synth.base_clip(2 * sr, sr, 1)
    r = np.roll(l, 1234) * np.float32(0.75)
    mt = msv.MultiTrack()
    mt.add_tracks_pcm([0], [np.stack([l, r], axis=1)], [sr])
    win, hop, n_fft = msv.track_params(sr)
    ref = orc.calc_spec(l + r, win, hop, n_fft, None, msv.calc_mel_fb_default(sr, n_fft))
    assert_db_close(mt.get_spec_db(0), ref, "stereo sum")
    mt.close()


# ---- display stages -----------------------------------------------------------------------------------
def test_spec_to_grey_parity(msv, orc):
    rng = np.random.default_rng(0)
    spec = (-150 + 130 * rng.random((300, 77))).astype(np.float32)
    for up in (1.0, 1.3259, 1.70609):
        got, ref = msv.spec_to_grey(spec, up, -20.0, -140.0), orc.spec_to_grey(spec, up, -20.0, -140.0)
        assert got.shape == ref.shape and np.array_equal(got, ref)


@pytest.mark.parametrize("shape,new", [((370, 440), (1000, 500)), ((438, 440), (1000, 500)), ((257, 2000), (300, 120)), ((64, 64), (64, 64)),
                                       ((1025, 300), (300, 500)), ((40, 3000), (50, 10)), ((500, 100), (1, 1)), ((3, 5), (40, 30)),
                                       ((8193, 64), (64, 500))])
def test_grey_to_rgb_parity(msv, orc, shape, new):
    """display.rs:56-61 incl. the per-pass clamp of image 0.23's resize; noisy content makes the clamp bite."""
    rng = np.random.default_rng(shape[0] * 7 + new[0])
    g = rng.random(shape).astype(np.float32) ** 3
    g[rng.random(shape) < 0.3] = 0.0
    for ch in (3, 4):
        got, ref = msv.grey_to_rgb(g, new[0], new[1], ch), orc.grey_to_rgb(g, new[0], new[1], ch)
        mx, frac = assert_pixels_close(got, ref, f"grey_to_rgb {shape}->{new}")
        print(f"grey_to_rgb {shape}->{new} ch={ch}: max diff {mx}, mismatching bytes {frac:.2e}")
        if ch == 4:
            assert np.all(got[..., 3] == 255)


# ---- MultiTrack: the public path ----------------------------------------------------------------------
def _clips():
    z = np.load(os.path.join(HERE, "golden", "clips.npz"))
    return {k: (z[f"pcm_{k}"], int(z[f"sr_{k}"])) for k in NAMES}


def test_multitrack_golden_clips(msv, orc):
    """The reference's multitrack_works flow (lib.rs:516-546) on its own sample audio (first 2 s of each
    fixture, int16 ingest), against the committed oracle outputs and the live oracle."""
    clips = _clips()
    exp = np.load(os.path.join(HERE, "golden", "expected.npz"))
    mt = msv.MultiTrack()
    assert mt.add_tracks_pcm(list(range(5)), [clips[k][0] for k in NAMES], [clips[k][1] for k in NAMES]) is True
    assert_range_close((mt.get_max_db(), mt.get_min_db()), (float(exp["max_db"]), float(exp["min_db"])), what="golden clips")
    for i, k in enumerate(NAMES):
        spec = mt.get_spec_db(i)
        assert tuple(exp[f"spec_shape_{k}"]) == spec.shape
        # the committed file keeps every 7th frame / 5th band; the live oracle must reproduce it, then the
        # full spectrogram is compared with the live oracle
        x = clips[k][0].astype(np.float32) / np.float32(32768.0)
        win, hop, n_fft = msv.track_params(clips[k][1])
        live = orc.calc_spec(x, win, hop, n_fft, None, msv.calc_mel_fb_default(clips[k][1], n_fft))
        assert np.max(np.abs(live[::7, ::5] - exp[f"spec_sub_{k}"])) < 2e-3
        assert_db_close(spec, live, f"golden clip {k}")
        img = mt.get_spec_image(i, 100.0, 120).reshape(120, -1, 3)
        mx, frac = assert_pixels_close(img, exp[f"img_{k}"], f"golden image {k}")
        print(f"golden image {k}: max diff {mx}, mismatching bytes {frac:.2e}")
        assert mt.get_sr(i) == clips[k][1] and abs(mt.get_sec(i) - 2.0) < 1e-6
    assert mt.get_max_sec() == pytest.approx(2.0)
    mt.close()


def _oracle_batch(orc, msv, wavs, srs, settings=None, nheight=500, px=100.0, channels=3):
    params = [msv.track_params(sr, settings) for sr in srs]
    mel = settings is None or settings.freq_scale == msv.FREQ_MEL
    windows = [orc.calc_window(p[0], p[2]) for p in params]
    if not mel:
        fbs = [None] * len(wavs)
    elif settings is not None and settings.n_mel:
        fbs = [orc.calc_mel_fb(sr, p[2], settings.n_mel) for sr, p in zip(srs, params)]
    else:
        fbs = [orc.calc_mel_fb_default(sr, p[2]) for sr, p in zip(srs, params)]
    return orc.pipeline(wavs, srs, params, windows, fbs, mel_scale=mel, px_per_sec=px, nheight=nheight, channels=channels)


def test_multitrack_six_rates_end_to_end(msv, orc):
    """C2 shape: six sample rates, default settings, 100 px/s x 500 (bench.rs:57), RGB and RGBA."""
    srs = [8000, 16000, 22050, 24000, 44100, 48000]
    wavs = [synth.derive_track(synth.base_clip(int(3.3 * sr), sr, seed=sr), i) for i, sr in enumerate(srs)]
    imgs, mx, mn = _oracle_batch(orc, msv, wavs, srs)
    mt = msv.MultiTrack()
    assert mt.add_tracks_pcm(list(range(6)), wavs, srs)
    assert_range_close((mt.get_max_db(), mt.get_min_db()), (mx, mn), what="six rates")
    # the committed range against the f64 truth: the engine's min_db may not be further off than the f32 reference's
    truths = []
    for w, sr in zip(wavs, srs):
        win, hop, n_fft = msv.track_params(sr)
        truths.append(orc.calc_spec_f64(w, win, hop, n_fft, None, msv.calc_mel_fb_default(sr, n_fft)))
    tmx = min(max(float(t.max()) for t in truths), 0.0)
    tmn = max(min(float(t.min()) for t in truths), tmx - 120.0)   # lib.rs:208-209 in f64
    assert_min_db_vs_truth(mt.get_min_db(), mn, tmn, "six rates")
    for i, t in enumerate(truths):
        win, hop, n_fft = msv.track_params(srs[i])
        ref_i = orc.calc_spec(wavs[i], win, hop, n_fft, None, msv.calc_mel_fb_default(srs[i], n_fft))
        assert_db_vs_truth(mt.get_spec_db(i), ref_i, t, f"six rates track {i}")
    for i in range(6):
        rgb = mt.get_spec_image(i, 100.0, 500).reshape(500, -1, 3)
        dmx, frac = assert_pixels_close(rgb, imgs[i], f"track {i}")
        print(f"track {i} sr={srs[i]}: image {rgb.shape}, max diff {dmx}, mismatching bytes {frac:.2e}")
        rgba = mt.get_spec_image_rgba(i, 100.0, 500).reshape(500, -1, 4)
        assert np.array_equal(rgba[..., :3], rgb) and np.all(rgba[..., 3] == 255)
        assert mt.image_width(i, 100.0) == orc.calc_nwidth(100.0, wavs[i].size, srs[i])
    mt.close()


def test_multitrack_linear_and_fixed_mel(msv, orc):
    srs = [44100, 22050]
    wavs = [synth.base_clip(2 * sr, sr, seed=3 + sr) for sr in srs]
    for settings in (msv.Settings.default(freq_scale=msv.FREQ_LINEAR, win_length=2048, hop_length=512, n_fft=2048),
                     msv.Settings.default(n_mel=128, win_length=1024, hop_length=256, n_fft=1024)):
        imgs, mx, mn = _oracle_batch(orc, msv, wavs, srs, settings, nheight=200, px=80.0, channels=4)
        mt = msv.MultiTrack(settings)
        mt.add_tracks_pcm([10, 20], wavs, srs)
        assert_range_close((mt.get_max_db(), mt.get_min_db()), (mx, mn), what="linear / fixed mel")
        for i, tid in enumerate([10, 20]):
            got = mt.get_spec_image_rgba(tid, 80.0, 200).reshape(200, -1, 4)
            assert_pixels_close(got, imgs[i], f"track {tid}")
        mt.close()


def test_multitrack_semantics(msv, orc):
    """changed flags, re-normalisation on remove (lib.rs:265-292), getters, unknown ids."""
    sr = 16000
    base = synth.base_clip(2 * sr, sr, 5)
    quiet, loud = base * np.float32(0.125), base
    mt = msv.MultiTrack()
    assert mt.add_tracks_pcm([1], [quiet], [sr]) is True
    r1 = (mt.get_max_db(), mt.get_min_db())
    img_before = mt.get_spec_image(1, 50.0, 64)
    assert mt.add_tracks_pcm([2], [loud], [sr]) is True            # louder track moves the global max
    assert mt.get_max_db() > r1[0] + 17.0
    img_mid = mt.get_spec_image(1, 50.0, 64)
    assert not np.array_equal(img_before, img_mid)                  # track 1 re-normalised
    assert mt.add_tracks_pcm([3], [quiet], [sr]) is False           # range and max_sr unchanged (lib.rs:211-229)
    assert mt.remove_track(3) is False
    assert mt.remove_track(2) is True                                # back to the quiet range
    assert (mt.get_max_db(), mt.get_min_db()) == pytest.approx(r1, abs=1e-6)
    assert np.array_equal(mt.get_spec_image(1, 50.0, 64), img_before)
    assert mt.get_frequency_hz(1, 1.0) == pytest.approx(sr / 2, rel=1e-5)
    assert mt.get_frequency_hz(1, 0.5) == pytest.approx(msv.mel_to_hz(msv.hz_to_mel(sr / 2) * 0.5), rel=1e-6)
    for call in (lambda: mt.get_spec_image(99, 50.0, 64), lambda: mt.remove_track(99), lambda: mt.get_sr(99)):
        with pytest.raises(msv.SgxError) as e:
            call()
        assert e.value.code == msv.SGX_ERR_UNKNOWN_ID
    with pytest.raises(msv.SgxError):                               # shorter than one window
        mt.add_tracks_pcm([7], [base[:100]], [sr])
    mt.close()


def test_add_tracks_from_wav_files(msv, orc, tmp_path):
    import wave

    clips = _clips()
    paths = []
    for k in NAMES[:3]:
        p = tmp_path / f"sample_{k}.wav"
        with wave.open(str(p), "wb") as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(clips[k][1]); w.writeframes(clips[k][0].tobytes())
        paths.append(str(p))
    mt, mt2 = msv.MultiTrack(), msv.MultiTrack()
    assert mt.add_tracks([0, 1, 2], "\n".join(paths))
    mt2.add_tracks_pcm([0, 1, 2], [clips[k][0].astype(np.float32) / np.float32(32768) for k in NAMES[:3]], [clips[k][1] for k in NAMES[:3]])
    for i in range(3):
        assert np.array_equal(mt.get_spec_db(i), mt2.get_spec_db(i))          # int16 ingest == f32 ingest, bit for bit
        assert np.array_equal(mt.get_spec_image(i, 100.0, 100), mt2.get_spec_image(i, 100.0, 100))
    assert mt.get_filename(1) == "sample_16k.wav" and mt.get_path(1) == paths[1]
    with pytest.raises(msv.SgxError) as e:
        mt.add_tracks([5, 6], paths[0] + "\n" + str(tmp_path / "nope.wav"))
    assert e.value.code == msv.SGX_ERR_IO
    with pytest.raises(msv.SgxError):
        mt.get_sr(5)                                                            # atomic: nothing inserted
    mt.close(); mt2.close()


def test_wav_image_parity(msv, orc):
    """display.rs:63-115 (get_wav_image)."""
    sr = 8000
    x = synth.base_clip(3 * sr, sr, 11) * np.float32(3.0)
    for nw, nh, rng in [(300, 100, (-1.0, 1.0)), (24000 * 2, 64, (-1.0, 1.0)), (50, 40, (-0.5, 0.5))]:
        assert np.array_equal(msv.wav_to_image(x, nw, nh, rng), orc.wav_to_image(x, nw, nh, *rng)), (nw, nh)
    mt = msv.MultiTrack()
    mt.add_tracks_pcm([0], [x], [sr])
    got = mt.get_wav_image(0, 100.0, 100, -1.0, 1.0).reshape(100, 300, 4)
    assert np.array_equal(got, orc.wav_to_image(x, 300, 100, -1.0, 1.0))
    mt.close()


# ---- edge cases: ragged batches, maximum sizes, buffer re-use, extreme render geometry ----------------------
def test_ragged_batch_and_id_reuse(msv, orc):
    """Tracks of very different lengths in ONE add_tracks call (tile prefix logic), then the same ids re-added with
    other lengths / rates (device buffers and range slots are re-used, lib.rs HashMap::insert semantics)."""
    sr = 22050
    lens = [884, 885, 1500, 22050, 3 * 22050 + 7, 16 * 221 + 884, 40000]        # from the shortest legal input (n == win) up
    wavs = [synth.base_clip(n, sr, seed=n) for n in lens]
    imgs, mx, mn = _oracle_batch(orc, msv, wavs, [sr] * len(wavs), nheight=64, px=300.0)
    mt = msv.MultiTrack()
    mt.add_tracks_pcm(list(range(len(wavs))), wavs, [sr] * len(wavs))
    assert_range_close((mt.get_max_db(), mt.get_min_db()), (mx, mn), what="ragged")
    for i, im in enumerate(imgs):
        got = mt.get_spec_image(i, 300.0, 64).reshape(64, -1, 3)
        assert_pixels_close(got, im, f"ragged track {i} (n={lens[i]})")
    # re-use ids with different content: longer, shorter, other sample rate
    srs2 = [48000, 22050, 8000]
    wavs2 = [synth.base_clip(n, r, seed=n + r) for n, r in zip([100000, 900, 30000], srs2)]
    mt.add_tracks_pcm([4, 0, 2], wavs2, srs2)
    all_w = [wavs2[1], wavs[1], wavs2[2], wavs[3], wavs2[0], wavs[5], wavs[6]]
    all_sr = [22050, sr, 8000, sr, 48000, sr, sr]
    imgs, mx, mn = _oracle_batch(orc, msv, all_w, all_sr, nheight=64, px=300.0)
    assert_range_close((mt.get_max_db(), mt.get_min_db()), (mx, mn), what="after id re-use")
    for i, im in enumerate(imgs):
        assert_pixels_close(mt.get_spec_image(i, 300.0, 64).reshape(64, -1, 3), im, f"re-used track {i}")
    mt.close()


def test_maximum_fft_size_default_mel(msv, orc):
    """n_fft = 16384 (the largest supported), 96 kHz, default mel bank (mel.rs:87-99 search), int16 stereo ingest."""
    sr = 96000
    st = msv.Settings.default(win_length=16384, hop_length=4096, n_fft=16384)
    l = synth.base_clip_i16(3 * sr, sr, 12)
    pcm = np.stack([l, np.roll(l, 77)], axis=1)
    mt = msv.MultiTrack(st)
    mt.add_tracks_pcm([0], [pcm], [sr])
    mono = pcm[:, 0].astype(np.float32) / np.float32(32768) + pcm[:, 1].astype(np.float32) / np.float32(32768)
    fb = msv.calc_mel_fb_default(sr, 16384)
    ref = orc.calc_spec(mono, 16384, 4096, 16384, None, fb)
    assert mt.spec_shape(0) == ref.shape
    assert_db_close(mt.get_spec_db(0), ref, "n_fft 16384 default mel")
    mt.close()


def test_many_tracks_grow_range_slots(msv):
    """More tracks than the initial 1024 range slots: the slot array grows without losing extrema."""
    sr = 8000
    base = synth.base_clip(400, sr, 3)
    n_tr = 1100
    wavs = [base * np.float32(0.5 + 0.5 * (i % 7) / 7.0) for i in range(n_tr)]
    mt = msv.MultiTrack()
    mt.add_tracks_pcm(list(range(n_tr)), wavs, [sr] * n_tr)
    loud = msv.MultiTrack()
    loud.add_tracks_pcm([0], [base * np.float32(0.5 + 0.5 * 6 / 7.0)], [sr])
    assert mt.get_max_db() == loud.get_max_db()
    assert np.array_equal(mt.get_spec_db(6), loud.get_spec_db(0))
    assert mt.remove_track(6) in (True, False) and mt.get_sr(1099) == sr
    mt.close(); loud.close()


@pytest.mark.parametrize("px,nh", [(0.5, 1), (3.0, 7), (1000.0, 33), (100.0, 2000), (0.01, 10)])
def test_extreme_render_geometry(msv, orc, px, nh):
    """Strong minification (chunked general path), strong magnification, 1-pixel and empty images."""
    sr = 16000
    x = synth.base_clip(12 * sr, sr, 21)
    imgs, mx, mn = _oracle_batch(orc, msv, [x], [sr], nheight=nh, px=px, channels=4)
    mt = msv.MultiTrack()
    mt.add_tracks_pcm([0], [x], [sr])
    got = mt.get_spec_image_rgba(0, px, nh)
    assert got.size == imgs[0].size
    if got.size:
        assert_pixels_close(got.reshape(imgs[0].shape), imgs[0], f"px={px} nh={nh}", max_mismatch_frac=0.02)
    mt.close()


@pytest.mark.parametrize("n_fft,hop,px,nh,seconds", [
    (512, 128, 100.0, 500, 8),      # C4 short window: 3.4x horizontal minification (21 taps per column)
    (8192, 2048, 100.0, 500, 20),   # C4 long window: 8x vertical minification, horizontal magnification (256-column tiles)
    (16384, 4096, 100.0, 500, 30),  # 16x vertical minification (~100 taps per row)
    (4096, 1024, 1000.0, 33, 6),    # a handful of frames under ~1,300 rows per tile: row-major fill of the grey tile
])
def test_wide_render_paths(msv, orc, n_fft, hop, px, nh, seconds):
    """The wide K3 path (runtime tap counts, source window of a tile in shared memory) through the public
    MultiTrack call, linear frequency scale as in the C4 sweep, against the oracle's resize."""
    sr = 44100
    x = synth.base_clip(seconds * sr, sr, 404 + n_fft)
    settings = msv.Settings.default(freq_scale=msv.FREQ_LINEAR, win_length=n_fft, hop_length=hop, n_fft=n_fft)
    imgs, mx, mn = _oracle_batch(orc, msv, [x], [sr], settings, nheight=nh, px=px, channels=4)
    mt = msv.MultiTrack(settings)
    mt.add_tracks_pcm([0], [x], [sr])
    assert_range_close((mt.get_max_db(), mt.get_min_db()), (mx, mn), what=f"n_fft={n_fft}")
    got = mt.get_spec_image_rgba(0, px, nh).reshape(imgs[0].shape)
    dmx, frac = assert_pixels_close(got, imgs[0], f"wide render n_fft={n_fft} px={px} nh={nh}")
    print(f"wide render n_fft={n_fft}: image {got.shape}, max diff {dmx}, mismatching bytes {frac:.2e}")
    rgb = mt.get_spec_image(0, px, nh).reshape(nh, -1, 3)
    assert np.array_equal(rgb, got[..., :3])
    mt.close()


@pytest.mark.parametrize("kind,n_fft,n_mel", [("dense", 2048, 6), ("dense", 2048, 64), ("dense", 4096, 10), ("ragged", 2048, 96),
                                              ("ragged", 512, 40), ("ragged", 1024, 33), ("ragged", 16384, 50)])
def test_arbitrary_filterbank_parity(msv, orc, kind, n_fft, n_mel):
    """calc_spec takes ANY [bins x n_mel] matrix (lib.rs:131 is a dense dot): dense banks (every filter spans all
    bins: 32 lanes per filter; the large one does not fit the shared-memory region and takes the banded path) and
    ragged banks (unordered, overlapping, empty and single-bin filters) through the block-padded mel path."""
    sr = 44100
    rng = np.random.default_rng(n_fft + n_mel)
    bins = n_fft // 2 + 1
    if kind == "dense":
        fb = (rng.random((bins, n_mel)) * 2e-3 + 1e-5).astype(np.float32)
    else:
        fb = np.zeros((bins, n_mel), dtype=np.float32)
        for m in range(n_mel):
            if m % 11 == 5:
                continue                                   # an empty filter
            lo = int(rng.integers(0, bins))
            cnt = 1 if m % 7 == 3 else int(rng.integers(1, max(2, bins // 6)))
            hi = min(bins, lo + cnt)
            fb[lo:hi, m] = (rng.random(hi - lo) * 1e-2 + 1e-4).astype(np.float32)
    x = synth.base_clip(max(5 * n_fft, sr), sr, seed=n_fft * 3 + n_mel)
    ref = orc.calc_spec(x, n_fft, n_fft // 4, n_fft, None, fb)
    got = msv.melspectrogram_db(x, n_fft, n_fft // 4, n_fft, None, fb)
    e, m_ = assert_db_close(got, ref, f"{kind} bank n_fft={n_fft} n_mel={n_mel}")
    print(f"{kind} bank n_fft={n_fft} n_mel={n_mel}: {e:.2e} dB / {m_:.2e} mag")
    assert np.array_equal(got <= -359.0, ref <= -359.0)

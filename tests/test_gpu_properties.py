"""Size-independent properties at the full BASELINE sizes (where the CPU oracle would take minutes)."""
import numpy as np
import pytest

import synth
from parity import assert_db_close

pytestmark = pytest.mark.gpu


def _torch():
    import torch

    return torch


def test_c5_track_properties(msv, orc):
    """One C5 track (10 min, 48 kHz, defaults): frame count, gain linearity (+20log10 g dB exactly up to
    rounding), time-shift covariance, image geometry; a 20 s window is checked against the oracle."""
    sr, n = 48000, 600 * 48000
    base = synth.base_clip(n, sr, 5005)
    mt = msv.MultiTrack()
    g = np.float32(0.5)
    mt.add_tracks_pcm([0, 1], [base, base * g], [sr, sr])
    T, M = mt.spec_shape(0)
    assert (T, M) == (n // 480 + 1, 347)
    a, b = mt.get_spec_db(0), mt.get_spec_db(1)
    assert np.abs((a - b) - np.float32(20 * np.log10(2.0))).max() <= 1e-3          # linearity
    win, hop, n_fft = msv.track_params(sr)
    fb = msv.calc_mel_fb_default(sr, n_fft)
    f0 = 10417
    seg = slice(f0 * hop, f0 * hop + 20 * sr)                                        # starts on a frame boundary
    ref = orc.calc_spec(base[seg], win, hop, n_fft, None, fb)
    inner = slice(8, ref.shape[0] - 8)                                               # frames not touching the segment's reflect edges
    assert_db_close(a[f0:f0 + ref.shape[0]][inner], ref[inner], "C5 window")
    assert mt.image_width(0, 100.0) == 60000
    img = mt.get_spec_image_rgba(0, 100.0, 500).reshape(500, 60000, 4)
    assert img[..., 3].min() == 255 and img[..., :3].std() > 10
    mt.close()


def test_c3_stereo_properties(msv, orc):
    """C3 shape (48 kHz stereo, n_fft 4096, hop 256, mel-128) on 2 minutes: L+R sum, frame count, a window vs the oracle."""
    sr, n = 48000, 120 * 48000
    l = synth.base_clip(n, sr, 3003)
    r = np.roll(l, 1234) * np.float32(0.75)
    s = msv.Settings.default(win_length=4096, hop_length=256, n_fft=4096, n_mel=128)
    mt = msv.MultiTrack(s)
    mt.add_tracks_pcm([0], [np.stack([l, r], axis=1)], [sr])
    T, M = mt.spec_shape(0)
    assert (T, M) == (n // 256 + 1, 128)
    a = mt.get_spec_db(0)
    fb = msv.calc_mel_fb(sr, 4096, 128)
    start = 256 * 9000
    ref = orc.calc_spec((l + r)[start:start + 5 * sr], 4096, 256, 4096, None, fb)
    inner = slice(16, ref.shape[0] - 16)
    assert_db_close(a[9000:9000 + ref.shape[0]][inner], ref[inner], "C3 window")
    w = mt.image_width(0, 100.0)
    img = mt.get_spec_image(0, 100.0, 500)
    assert img.size == w * 500 * 3 and w == 12000
    mt.close()


def test_device_resident_batch_matches_host_path(msv):
    """The bench's `value` path (PCM resident in HBM, deferred range, batched device render) produces the same
    bytes as the public host-buffer path."""
    torch = _torch()
    sr = 48000
    base = synth.base_clip(20 * sr, sr, 42)
    tracks = [synth.derive_track(base, t) for t in range(4)]
    mt = msv.MultiTrack()
    mt.add_tracks_pcm(list(range(4)), tracks, [sr] * 4)
    want = [mt.get_spec_image_rgba(i, 100.0, 500) for i in range(4)]
    rng_want = (mt.get_max_db(), mt.get_min_db())
    mt.close()

    dev = [torch.from_numpy(t).cuda() for t in tracks]
    sm = msv.ShardedMultiTrack()
    sm.add_tracks_device(list(range(4)), [d.data_ptr() for d in dev], [d.numel() for d in dev], [sr] * 4, keepalive=dev)
    w = sm.mt.image_width(0, 100.0)
    outs = [torch.empty(w * 500 * 4, dtype=torch.uint8, device="cuda") for _ in range(4)]
    sm.render_device(list(range(4)), 100.0, 500, 4, [o.data_ptr() for o in outs], [o.numel() for o in outs])
    assert sm.synchronize() is True
    assert (sm.get_max_db(), sm.get_min_db()) == rng_want
    for o, wv in zip(outs, want):
        assert np.array_equal(o.cpu().numpy(), wv)
    sm.close()


def test_bitwise_determinism(msv):
    """Race detector of last resort (compute-sanitizer is closed on this pool): the analysis and render
    kernels use named barriers and alias shared buffers, so two runs on the same input must agree bit for bit."""
    sr = 48000
    x = synth.base_clip(30 * sr, sr, 99)
    outs = []
    for _ in range(3):
        mt = msv.MultiTrack()
        mt.add_tracks_pcm([0, 1], [x, x[::-1].copy()], [sr, sr])
        outs.append((mt.get_spec_db(0), mt.get_spec_db(1), mt.get_spec_image_rgba(0, 100.0, 500), mt.get_spec_image(1, 37.5, 123)))
        mt.close()
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert np.array_equal(a, b)
    lin = [msv.melspectrogram_db(x[: 5 * sr], 4096, 256, 4096) for _ in range(2)]
    assert np.array_equal(lin[0], lin[1])


def test_many_zoom_levels_do_not_invalidate_axis_tables(msv):
    """The per-geometry resampling tables live in a bounded cache (128 entries).  A viewer zooming through more
    than 128 distinct px_per_sec values must keep getting the image a fresh handle renders: the cache may only be
    emptied between render calls, never while one is assembling its descriptors (round-1 advisor finding)."""
    sr = 16000
    x = synth.base_clip(6 * sr, sr, 17)
    y = synth.base_clip(4 * sr + 321, sr, 18)
    mt = msv.MultiTrack()
    mt.add_tracks_pcm([0, 1], [x, y], [sr, sr])
    zooms = [20.0 + 0.75 * i for i in range(150)]
    got = [(mt.get_spec_image(0, z, 64), mt.get_spec_image(1, z, 48)) for z in zooms]
    mt.close()
    for i in (0, 63, 64, 65, 127, 128, 129, 149):   # around the points where the cache is emptied
        fresh = msv.MultiTrack()
        fresh.add_tracks_pcm([0, 1], [x, y], [sr, sr])
        assert np.array_equal(fresh.get_spec_image(0, zooms[i], 64), got[i][0]), f"zoom {i}"
        assert np.array_equal(fresh.get_spec_image(1, zooms[i], 48), got[i][1]), f"zoom {i}"
        fresh.close()


def test_tensor_core_render_path_parity(orc):
    """The tcgen05 render kernel (SGX_K3_TC=1, csrc/render_tc_kernel.cu: both Lanczos passes as 3xTF32 MMAs with TMEM
    accumulators) must meet the pixel tolerance of the FP32 path: +-1 LSB on < 1 % of the bytes against the oracle, for
    RGB and RGBA, several sample rates (tile shapes) and a column window that is not a multiple of the tile width."""
    import os
    import subprocess
    import sys

    code = r'''
import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import msv_b200 as msv, oracle_binding, synth
from parity import assert_pixels_close, assert_range_close
orc = oracle_binding.load()
srs = [8000, 16000, 22050, 44100, 48000]
wavs = [synth.derive_track(synth.base_clip(int(3.7 * sr) + 13, sr, seed=sr), i) for i, sr in enumerate(srs)]
params = [orc.track_params(sr) for sr in srs]
fbs = [orc.calc_mel_fb_default(sr, p[2]) for sr, p in zip(srs, params)]
wins = [orc.calc_window(p[0], p[2]) for p in params]
for ch in (3, 4):
    imgs, mx, mn = orc.pipeline(wavs, srs, params, wins, fbs, channels=ch, px_per_sec=100.0, nheight=500)
    n0 = msv.kernel_launch_count()
    mt = msv.MultiTrack()
    mt.add_tracks_pcm(list(range(len(srs))), wavs, srs)
    assert_range_close((mt.get_max_db(), mt.get_min_db()), (mx, mn), what="tc range")
    got = mt.get_spec_images(list(range(len(srs))), 100.0, 500, ch)
    for i, g in enumerate(got):
        d, frac = assert_pixels_close(g.reshape(imgs[i].shape), imgs[i], f"tc track {i} ch {ch}")
        print(f"tc sr={srs[i]} ch={ch}: max diff {d}, mismatching bytes {frac:.2e}")
    mt.close()
print("K3 TC OK")
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SGX_K3_TC="1")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "K3 TC OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_sliding_window_render_kernel_matches_the_tile_kernel_bit_for_bit():
    """The 8-tap geometry class has two FP32 render kernels: the sliding-window one (default,
    csrc/render_slide_kernel.cu) and render_fast_kernel (SGX_K3_SLIDE=0).  They apply the same operations in the same
    order, so every pixel must be IDENTICAL -- over RGB and RGBA, widths that are and are not multiples of 4 (128-bit
    and scalar store paths), horizontal ratios at and below 1 (unit-step and general column blocks), heights that
    are not a multiple of the tile, several sample rates in one call (one launch over mixed geometries), and one
    10-minute track at the bench geometry (compared by digest).  The kernel is chosen per process, so both run in
    children."""
    import os
    import subprocess
    import sys
    import tempfile

    code = r'''
import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import msv_b200 as msv, synth
srs = [8000, 16000, 22050, 44100, 48000]
wavs = [synth.derive_track(synth.base_clip(int(3.3 * sr) + 17 * i + 5, sr, seed=sr), i) for i, sr in enumerate(srs)]
out = {}
mt = msv.MultiTrack()
mt.add_tracks_pcm(list(range(len(srs))), wavs, srs)
for ch in (3, 4):
    for pps, nh in ((100.0, 500), (100.0, 333), (173.0, 257), (250.0, 64), (97.0, 1000)):
        for i, g in enumerate(mt.get_spec_images(list(range(len(srs))), pps, nh, ch)):
            out[f"{ch}_{pps}_{nh}_{i}"] = np.asarray(g).copy()
mt.close()
# the bench geometry at FULL size (BASELINE C5: 10 minutes at 48 kHz -> 60,001 frames -> 60,000 x 500 RGBA): the tile
# capacity of the sliding-window kernel is exact for it (127 or 128 source frames per 120 columns, never 129)
import hashlib
x = synth.base_clip(600 * 48000, 48000, seed=99)
mt = msv.MultiTrack()
mt.add_tracks_pcm([0], [x], [48000])
img = np.asarray(mt.get_spec_images([0], 100.0, 500, 4)[0])
assert img.size == 60000 * 500 * 4
out["c5_full_sha256"] = np.frombuffer(hashlib.sha256(img.tobytes()).digest(), np.uint8).copy()
out["c5_full_mean"] = np.array([img.mean()])
mt.close()
np.savez(sys.argv[1], **out)
print("K3 AB OK", len(out))
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    got = {}
    with tempfile.TemporaryDirectory() as td:
        for name, flag in (("slide", "1"), ("tile", "0")):
            path = os.path.join(td, name + ".npz")
            env = dict(os.environ, SGX_K3_SLIDE=flag)
            r = subprocess.run([sys.executable, "-c", code, path], cwd=root, env=env, capture_output=True, text=True, timeout=600)
            assert r.returncode == 0 and "K3 AB OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
            with np.load(path) as z:
                got[name] = {k: z[k] for k in z.files}
    assert got["slide"].keys() == got["tile"].keys() and len(got["slide"]) == 52
    for k in got["slide"]:
        a, b = got["slide"][k], got["tile"][k]
        assert a.shape == b.shape and np.array_equal(a, b), f"{k}: {int((a != b).sum())} of {a.size} bytes differ"


@pytest.mark.parametrize("flag", ["SGX_K1W2", "SGX_K1W1", "SGX_K1BLOCK"])
def test_every_kernel_for_n_fft_2048_meets_parity(orc, flag):
    """n_fft = 2048 has three kernels: the block kernel (csrc/stft_kernel.cu), the warp-per-frame-pair one (SGX_K1W2=1,
    csrc/stft_warp2_kernel.cu) and the warp-per-frame one on packed complex values (SGX_K1W1=1, csrc/stft_warp1_kernel.cu).
    Whichever is the default is exercised by every other test; each must meet the same tolerances.  The choice is made
    per process, so they run in child interpreters (SGX_K1BLOCK=1 forces the block kernel)."""
    import subprocess
    import sys
    import os

    code = r'''
import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import msv_b200 as msv, oracle_binding, synth
from parity import assert_db_close
orc = oracle_binding.load()
x = synth.base_clip(3 * 48000 + 77, 48000, 5)
ref = orc.perform_stft(x, 1920, 480, 2048)
got = msv.perform_stft(x, 1920, 480, 2048)
peak = np.abs(ref).max(axis=1, keepdims=True)
assert (np.abs(got - ref) / peak).max() <= 1e-4
fb = msv.calc_mel_fb_default(48000, 2048)
assert_db_close(msv.melspectrogram_db(x, 1920, 480, 2048, None, fb), orc.calc_spec(x, 1920, 480, 2048, None, fb), "K1 W2 mel")
assert_db_close(msv.melspectrogram_db(x, 2048, 512, 2048), orc.calc_spec(x, 2048, 512, 2048), "K1 W2 linear")
y = synth.base_clip(2 * 44100, 44100, 6)  # odd hop 441: unaligned frame starts
fb = msv.calc_mel_fb_default(44100, 2048)
assert_db_close(msv.melspectrogram_db(y, 1764, 441, 2048, None, fb), orc.calc_spec(y, 1764, 441, 2048, None, fb), "K1 W2 44.1k")
# the whole path on a stereo int16 track and a mono f32 one (gathered and TMA-staged tiles), pixels against the oracle
from parity import assert_pixels_close, assert_range_close
srs = [48000, 48000]
wavs = [synth.derive_track(synth.base_clip(5 * 48000 + 123, 48000, seed=11), 0), synth.derive_track(synth.base_clip(4 * 48000 + 7, 48000, seed=12), 1)]
params = [orc.track_params(sr) for sr in srs]
fbs = [orc.calc_mel_fb_default(sr, p[2]) for sr, p in zip(srs, params)]
wins = [orc.calc_window(p[0], p[2]) for p in params]
imgs, mx, mn = orc.pipeline(wavs, srs, params, wins, fbs, channels=4, px_per_sec=100.0, nheight=500)
mt = msv.MultiTrack()
mt.add_tracks_pcm([0, 1], wavs, srs)
assert_range_close((mt.get_max_db(), mt.get_min_db()), (mx, mn), what="range")
for i, g in enumerate(mt.get_spec_images([0, 1], 100.0, 500, 4)):
    assert_pixels_close(g.reshape(imgs[i].shape), imgs[i], f"track {i}")
mt.close()
print("K1 W2 OK")
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SGX_K1W2="0", SGX_K1W1="0")
    env[flag] = "1"
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "K1 W2 OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("sr,seconds,px,settings_kw", [
    (48000, 40, 100.0, {}),                                                    # defaults: n_fft 2048, ratio ~1 (8-tap path)
    (44100, 21, 100.0, {}),                                                    # odd hop 441
    (48000, 30, 100.0, dict(win_length=4096, hop_length=256, n_fft=4096, n_mel=128)),  # C3 shape, 16-tap horizontal path
    (16000, 25, 7.5, {}),                                                      # zoomed out: general render path
    (22050, 9, 400.0, dict(freq_scale=0)),                                     # zoomed in, linear scale
    (44100, 30, 100.0, dict(freq_scale=0, win_length=8192, hop_length=2048, n_fft=8192)),  # C4 long window: wide path, 256-column tiles
    (44100, 12, 100.0, dict(freq_scale=0, win_length=512, hop_length=128, n_fft=512)),     # C4 short window: wide path, 21 horizontal taps
])
def test_time_sliced_track_equals_whole_track(msv, sr, seconds, px, settings_kw):
    """n3: one track cut into time slices -- each holding only the samples its strip of columns needs --
    must give, strip by strip, exactly the pixels of the whole-track image, and the same dB range."""
    torch = _torch()
    st = msv.Settings.default(**settings_kw)
    n = seconds * sr + 311
    x = synth.base_clip(n, sr, seed=sr + seconds)
    whole = msv.MultiTrack(st)
    whole.add_tracks_pcm([0], [x], [sr])
    nw = whole.image_width(0, px)
    ref = whole.get_spec_image_rgba(0, px, 200).reshape(200, nw, 4)
    ref_range = (whole.get_max_db(), whole.get_min_db())
    whole.close()

    parts = 5
    mt = msv.MultiTrack(st)           # one handle stands in for `parts` ranks: the slots of all slices are reduced together
    keep, strips = [], []
    for r in range(parts):
        ob = nw * r // parts
        oc = nw * (r + 1) // parts - ob
        fb, fc, sb, sc = msv.slice_plan(n, sr, px, ob, oc, st)
        assert sc < n or parts == 1                                   # a slice really is a part of the track
        chunk = torch.from_numpy(x[sb:sb + sc].copy()).cuda()
        keep.append(chunk)
        mt.add_track_slice_device(r, chunk.data_ptr(), sb, sc, n, sr, 1, fb, fc)
        strips.append((ob, oc))
    mt.commit_range_device()
    mt.set_global_max_sr(sr)
    outs = []
    for r, (ob, oc) in enumerate(strips):
        o = torch.empty(200 * oc * 4, dtype=torch.uint8, device="cuda")
        assert mt.render_slice_device(r, px, 200, 4, ob, oc, o.data_ptr(), o.numel()) == o.numel()
        outs.append(o)
    mt.synchronize()
    assert (mt.get_max_db(), mt.get_min_db()) == ref_range
    got = np.concatenate([o.cpu().numpy().reshape(200, oc, 4) for o, (ob, oc) in zip(outs, strips)], axis=1)
    assert np.array_equal(got, ref)
    with pytest.raises(msv.SgxError):                                  # a strip this slice does not cover
        mt.render_slice_device(0, px, 200, 4, strips[2][0], strips[2][1], outs[2].data_ptr(), outs[2].numel())
    mt.close()


def test_bench_rs_entry_points_run(msv):
    """benches/bench.cpp registers the reference's four criterion benches by name (benches/bench.rs:35,55,68,87)
    and drives them through the C ABI only."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "benches", "bench")
    if not os.path.exists(exe):
        pytest.skip("benches/bench not built (run __graft_entry__.build())")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    for name in ("get mel spectrogram", "draw spectrogram", "add track", "multitrack get spec image"):
        assert name in r.stdout
    print(r.stdout)

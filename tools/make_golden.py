"""Generates tests/golden/*.npz from the reference's own fixtures (run in the build container,
where /root/reference exists; the GPU box only sees the committed outputs).

  clips.npz      the first 2 s of each samples/sample_*.wav as int16 (the reference's test audio,
                 src_rust/lib.rs:516-546), with the sample rate of each clip
  expected.npz   what the CPU oracle produces for those clips through the MultiTrack default path
                 (per-clip dB max/min, strided dB samples, global range, image checksums) -- pins
                 the oracle against drift and gives the GPU tests a file-based target
  fixtures.json  facts about the full-length fixtures (frames, default n_mel, own dB extrema,
                 global range) as computed by the oracle here
"""
import json
import os
import sys
import wave
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding  # noqa: E402

REF = "/root/reference/samples"
NAMES = ["8k", "16k", "22k05", "24k", "44k1"]


def read_wav(path):
    with wave.open(path, "rb") as w:
        assert w.getsampwidth() == 2
        sr, ch, n = w.getframerate(), w.getnchannels(), w.getnframes()
        data = np.frombuffer(w.readframes(n), dtype="<i2").reshape(n, ch)
    return data, sr


def main():
    orc = oracle_binding.load()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    clips, srs = {}, {}
    full = {}
    for name in NAMES:
        data, sr = read_wav(os.path.join(REF, f"sample_{name}.wav"))
        assert data.shape[1] == 1
        clips[name] = data[: 2 * sr, 0].copy()
        srs[name] = sr
        full[name] = (data[:, 0].astype(np.float32) / np.float32(32768.0), sr)
    np.savez_compressed(os.path.join(out_dir, "clips.npz"), **{f"pcm_{k}": v for k, v in clips.items()},
                        **{f"sr_{k}": np.int64(v) for k, v in srs.items()})

    # ---- expected outputs for the clips (MultiTrack defaults: mel, 40 ms, 4x overlap, 120 dB) ----
    exp = {}
    wavs, params, windows, fbs, srl = [], [], [], [], []
    for name in NAMES:
        x = clips[name].astype(np.float32) / np.float32(32768.0)
        sr = srs[name]
        win, hop, n_fft = orc.track_params(sr)
        window = orc.calc_window(win, n_fft)
        fb = orc.calc_mel_fb_default(sr, n_fft)
        spec = orc.calc_spec(x, win, hop, n_fft, window, fb)
        exp[f"spec_shape_{name}"] = np.array(spec.shape)
        exp[f"spec_max_{name}"] = np.float32(spec.max())
        exp[f"spec_min_{name}"] = np.float32(spec.min())
        exp[f"spec_sub_{name}"] = spec[::7, ::5].copy()
        lin = orc.calc_spec(x, win, hop, n_fft, window, None)
        exp[f"lin_sub_{name}"] = lin[::11, ::13].copy()
        wavs.append(x); params.append((win, hop, n_fft)); windows.append(window); fbs.append(fb); srl.append(sr)
    imgs, mx, mn = orc.pipeline(wavs, srl, params, windows, fbs, mel_scale=True, px_per_sec=100.0, nheight=120, channels=3)
    exp["max_db"] = np.float32(mx)
    exp["min_db"] = np.float32(mn)
    for name, im in zip(NAMES, imgs):
        exp[f"img_{name}"] = im
    np.savez_compressed(os.path.join(out_dir, "expected.npz"), **exp)

    # ---- facts about the full fixtures ----
    facts = {}
    specs = []
    for name in NAMES:
        x, sr = full[name]
        win, hop, n_fft = orc.track_params(sr)
        fb = orc.calc_mel_fb_default(sr, n_fft)
        spec = orc.calc_spec(x, win, hop, n_fft, None, fb)
        specs.append(spec)
        facts[name] = {"sr": sr, "n": int(x.size), "win": win, "hop": hop, "n_fft": n_fft, "frames": int(spec.shape[0]),
                       "n_mel": int(spec.shape[1]), "db_max": float(spec.max()), "db_min": float(spec.min()),
                       "pcm_crc32": int(zlib.crc32((x * 32768).astype("<i2").tobytes()))}
    gmax = max(float(s.max()) for s in specs)
    gmin = min(float(s.min()) for s in specs)
    mx, mn = orc.clamp_range(gmax, gmin, 120.0)
    facts["global"] = {"max_db": mx, "min_db": mn, "max_sr": 44100}
    with open(os.path.join(out_dir, "fixtures.json"), "w") as f:
        json.dump(facts, f, indent=1, sort_keys=True)
    print(json.dumps(facts, indent=1))


if __name__ == "__main__":
    main()

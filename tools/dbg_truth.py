"""Where does the engine sit further from the f64 truth than the f32 oracle?  (diagnostic, GPU)"""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import msv_b200 as msv, oracle_binding, synth
orc = oracle_binding.load()
for sr in (24000, 48000, 8000):
    win, hop, n_fft = msv.track_params(sr)
    x = synth.base_clip(3 * sr + 17, sr, seed=sr)
    fb = msv.calc_mel_fb_default(sr, n_fft)
    # linear magnitudes against the f64 truth, relative to the frame peak
    truth = orc.stft_mag_f64(x, win, hop, n_fft)
    g = np.abs(msv.perform_stft(x, win, hop, n_fft)).astype(np.float64)
    r = np.abs(orc.perform_stft(x, win, hop, n_fft)).astype(np.float64)
    pk = truth.max(axis=1, keepdims=True)
    eg, er = np.abs(g - truth) / pk, np.abs(r - truth) / pk
    print(f"sr={sr} n_fft={n_fft}: |mag - truth| / frame peak: gpu rms {np.sqrt((eg**2).mean()):.2e} max {eg.max():.2e} | oracle rms {np.sqrt((er**2).mean()):.2e} max {er.max():.2e}")
    mg = msv.stft_magnitude(x, win, hop, n_fft).astype(np.float64)
    em = np.abs(mg - truth) / pk
    print(f"      stft_magnitude (sqrt.approx): rms {np.sqrt((em**2).mean()):.2e} max {em.max():.2e}")
    got = msv.melspectrogram_db(x, win, hop, n_fft, None, fb).astype(np.float64)
    ref = orc.calc_spec(x, win, hop, n_fft, None, fb).astype(np.float64)
    tr = orc.calc_spec_f64(x, win, hop, n_fft, None, fb)
    depth = tr.max(axis=1, keepdims=True) - tr
    for lo in (0, 30, 50, 60):
        band = depth >= lo
        print(f"      mel dB, bins >= {lo} dB below frame peak ({band.sum()}): gpu rms {np.sqrt(((got-tr)[band]**2).mean()):.2e} max {np.abs(got-tr)[band].max():.2e} | oracle rms {np.sqrt(((ref-tr)[band]**2).mean()):.2e} max {np.abs(ref-tr)[band].max():.2e}")
    w = np.argsort(np.abs(got - tr).ravel())[-3:]
    for i in w:
        t, m = divmod(int(i), tr.shape[1])
        print(f"      worst: frame {t} band {m}: truth {tr[t,m]:.5f} gpu {got[t,m]:.5f} oracle {ref[t,m]:.5f} depth {depth[t,m]:.1f} dB")

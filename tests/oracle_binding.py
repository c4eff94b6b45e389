"""ctypes access to oracle/liboracle.so -- the CPU restatement of the reference (TEST
INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")

_vp, _sz, _u32, _f = C.c_void_p, C.c_size_t, C.c_uint32, C.c_float


def build():
    src = os.path.join(ORACLE_DIR, "thesia_oracle.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])
    return LIB


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        L = lib
        L.orc_calc_proper_n_fft.restype = _sz; L.orc_calc_proper_n_fft.argtypes = [_sz]
        L.orc_hz_to_mel.restype = _f; L.orc_hz_to_mel.argtypes = [_f]
        L.orc_mel_to_hz.restype = _f; L.orc_mel_to_hz.argtypes = [_f]
        L.orc_hz_to_mel_f64.restype = C.c_double; L.orc_hz_to_mel_f64.argtypes = [C.c_double]
        L.orc_mel_to_hz_f64.restype = C.c_double; L.orc_mel_to_hz_f64.argtypes = [C.c_double]
        L.orc_stft_n_frames.restype = C.c_long; L.orc_stft_n_frames.argtypes = [_sz, _sz, _sz]
        L.orc_perform_stft.restype = C.c_long
        L.orc_perform_stft.argtypes = [_vp, _sz, _sz, _sz, _sz, _vp, _vp, C.c_int]
        L.orc_stft_mag_f64.restype = C.c_long
        L.orc_stft_mag_f64.argtypes = [_vp, _sz, _sz, _sz, _sz, _vp, _vp]
        L.orc_calc_mel_fb_default.restype = _sz; L.orc_calc_mel_fb_default.argtypes = [_u32, _sz, _vp, _sz]
        L.orc_calc_mel_fb.restype = None; L.orc_calc_mel_fb.argtypes = [_u32, _sz, _sz, _f, _f, C.c_int, _vp]
        L.orc_calc_spec.restype = C.c_long
        L.orc_calc_spec.argtypes = [_vp, _sz, _sz, _sz, _sz, _vp, _vp, _sz, _vp, C.c_int, C.c_int]
        L.orc_calc_spec_f64.restype = C.c_long
        L.orc_calc_spec_f64.argtypes = [_vp, _sz, _sz, _sz, _sz, _vp, _vp, _sz, _vp]
        L.orc_amp_to_db_default.restype = C.c_int; L.orc_amp_to_db_default.argtypes = [_vp, _sz]
        L.orc_grey_height.restype = _u32; L.orc_grey_height.argtypes = [_sz, _f]
        L.orc_spec_to_grey.restype = _u32; L.orc_spec_to_grey.argtypes = [_vp, _sz, _sz, _f, _f, _f, _vp]
        L.orc_resize_lanczos3.restype = C.c_int
        L.orc_resize_lanczos3.argtypes = [_vp, _u32, _u32, _u32, _u32, _vp, C.c_int]
        L.orc_grey_to_rgb.restype = C.c_int
        L.orc_grey_to_rgb.argtypes = [_vp, _u32, _u32, _u32, _u32, C.c_int, _vp, C.c_int]
        L.orc_convert_grey_to_color.restype = C.c_int; L.orc_convert_grey_to_color.argtypes = [_f, _vp]
        L.orc_calc_nwidth.restype = _u32; L.orc_calc_nwidth.argtypes = [_f, _sz, _u32]
        L.orc_up_ratio.restype = _f; L.orc_up_ratio.argtypes = [_u32, _u32, C.c_int]
        L.orc_wav_to_image.restype = C.c_int; L.orc_wav_to_image.argtypes = [_vp, _sz, _u32, _u32, _f, _f, _vp]
        L.orc_track_params.restype = None
        L.orc_track_params.argtypes = [_u32, _f, _sz, _sz, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz)]
        L.orc_clamp_range.restype = None
        L.orc_clamp_range.argtypes = [_f, _f, _f, C.POINTER(_f), C.POINTER(_f)]
        L.orc_spec_max_min.restype = None
        L.orc_spec_max_min.argtypes = [_vp, _sz, C.POINTER(_f), C.POINTER(_f)]
        L.orc_rfft_f64.restype = C.c_int; L.orc_rfft_f64.argtypes = [_vp, _sz, _vp]
        L.orc_cfft_f64.restype = C.c_int; L.orc_cfft_f64.argtypes = [_vp, _sz]
        L.orc_rfft_plan_new.restype = _vp; L.orc_rfft_plan_new.argtypes = [_sz]
        L.orc_rfft_plan_free.restype = None; L.orc_rfft_plan_free.argtypes = [_vp]
        L.orc_rfft_process.restype = None; L.orc_rfft_process.argtypes = [_vp, _vp, _vp]
        L.orc_pad_reflect.restype = C.c_int; L.orc_pad_reflect.argtypes = [_vp, _sz, _sz, _sz, _vp]
        L.orc_pad_constant.restype = None; L.orc_pad_constant.argtypes = [_vp, _sz, _sz, _sz, _f, _vp]
        L.orc_hann.restype = None; L.orc_hann.argtypes = [_sz, C.c_int, _vp]
        L.orc_calc_window.restype = None; L.orc_calc_window.argtypes = [_sz, _sz, _vp]
        L.orc_get_colormap.restype = None; L.orc_get_colormap.argtypes = [_vp]
        L.orc_num_threads.restype = C.c_int
        L.orc_pipeline.restype = C.c_int

    # ---- tables ---------------------------------------------------------------------------
    def hann(self, n, symmetric=False):
        out = np.empty(n, np.float32)
        self.lib.orc_hann(n, int(symmetric), out.ctypes.data)
        return out

    def calc_window(self, win, n_fft):
        out = np.empty(win, np.float32)
        self.lib.orc_calc_window(win, n_fft, out.ctypes.data)
        return out

    def track_params(self, sr, win_ms=40.0, t_overlap=4, f_overlap=1):
        w, h, f = _sz(), _sz(), _sz()
        self.lib.orc_track_params(sr, win_ms, t_overlap, f_overlap, C.byref(w), C.byref(h), C.byref(f))
        return w.value, h.value, f.value

    def calc_mel_fb(self, sr, n_fft, n_mel, fmin=0.0, fmax=None, do_norm=True):
        out = np.empty((n_fft // 2 + 1, n_mel), np.float32)
        self.lib.orc_calc_mel_fb(sr, n_fft, n_mel, fmin, -1.0 if fmax is None else fmax, int(do_norm), out.ctypes.data)
        return out

    def calc_mel_fb_default(self, sr, n_fft):
        n_mel = self.lib.orc_calc_mel_fb_default(sr, n_fft, None, 0)
        out = np.empty((n_fft // 2 + 1, n_mel), np.float32)
        self.lib.orc_calc_mel_fb_default(sr, n_fft, out.ctypes.data, out.size)
        return out

    def pad_reflect(self, x, left, right):
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty(x.size + left + right, np.float32)
        rc = self.lib.orc_pad_reflect(x.ctypes.data, x.size, left, right, out.ctypes.data)
        assert rc == 0
        return out

    def pad_constant(self, x, left, right, c=0.0):
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty(x.size + left + right, np.float32)
        self.lib.orc_pad_constant(x.ctypes.data, x.size, left, right, c, out.ctypes.data)
        return out

    # ---- FFT / STFT -------------------------------------------------------------------------
    def rfft_f32(self, x):
        x = np.ascontiguousarray(x, np.float32).copy()
        p = self.lib.orc_rfft_plan_new(x.size)
        assert p
        out = np.empty((x.size // 2 + 1, 2), np.float32)
        self.lib.orc_rfft_process(p, x.ctypes.data, out.ctypes.data)
        self.lib.orc_rfft_plan_free(p)
        return out.view(np.complex64)[:, 0]

    def rfft_f64(self, x):
        x = np.ascontiguousarray(x, np.float64)
        out = np.empty((x.size // 2 + 1, 2), np.float64)
        assert self.lib.orc_rfft_f64(x.ctypes.data, x.size, out.ctypes.data) == 0
        return out.view(np.complex128)[:, 0]

    def cfft_f64(self, z):
        z = np.ascontiguousarray(z, np.complex128).copy()
        assert self.lib.orc_cfft_f64(z.ctypes.data, z.size) == 0
        return z

    def stft_n_frames(self, n, win, hop):
        return self.lib.orc_stft_n_frames(n, win, hop)

    def perform_stft(self, x, win, hop, n_fft, window=None, parallel=True):
        x = np.ascontiguousarray(x, np.float32)
        T = self.lib.orc_stft_n_frames(x.size, win, hop)
        if T < 0:
            raise ValueError("reference would panic")
        w = None if window is None else np.ascontiguousarray(window, np.float32)
        out = np.empty((T, n_fft // 2 + 1, 2), np.float32)
        rc = self.lib.orc_perform_stft(x.ctypes.data, x.size, win, hop, n_fft, None if w is None else w.ctypes.data,
                                       out.ctypes.data, int(parallel))
        assert rc == T
        return out.view(np.complex64)[..., 0]

    def stft_mag_f64(self, x, win, hop, n_fft, window=None):
        x = np.ascontiguousarray(x, np.float32)
        T = self.lib.orc_stft_n_frames(x.size, win, hop)
        w = None if window is None else np.ascontiguousarray(window, np.float32)
        out = np.empty((T, n_fft // 2 + 1), np.float64)
        rc = self.lib.orc_stft_mag_f64(x.ctypes.data, x.size, win, hop, n_fft, None if w is None else w.ctypes.data, out.ctypes.data)
        assert rc == T
        return out

    def calc_spec(self, x, win, hop, n_fft, window=None, mel_fb=None, parallel=True, dense=False):
        """lib.rs:112-136: dB spectrogram [T, n_out] (mel_fb None -> linear)."""
        x = np.ascontiguousarray(x, np.float32)
        T = self.lib.orc_stft_n_frames(x.size, win, hop)
        if T < 0:
            raise ValueError("reference would panic")
        w = self.calc_window(win, n_fft) if window is None else np.ascontiguousarray(window, np.float32)
        fb = None if mel_fb is None else np.ascontiguousarray(mel_fb, np.float32)
        n_out = n_fft // 2 + 1 if fb is None else fb.shape[1]
        out = np.empty((T, n_out), np.float32)
        rc = self.lib.orc_calc_spec(x.ctypes.data, x.size, win, hop, n_fft, w.ctypes.data, None if fb is None else fb.ctypes.data,
                                    0 if fb is None else fb.shape[1], out.ctypes.data, int(parallel), int(dense))
        assert rc == T, rc
        return out

    def calc_spec_f64(self, x, win, hop, n_fft, window=None, mel_fb=None):
        x = np.ascontiguousarray(x, np.float32)
        T = self.lib.orc_stft_n_frames(x.size, win, hop)
        w = self.calc_window(win, n_fft) if window is None else np.ascontiguousarray(window, np.float32)
        fb = None if mel_fb is None else np.ascontiguousarray(mel_fb, np.float32)
        n_out = n_fft // 2 + 1 if fb is None else fb.shape[1]
        out = np.empty((T, n_out), np.float64)
        rc = self.lib.orc_calc_spec_f64(x.ctypes.data, x.size, win, hop, n_fft, w.ctypes.data, None if fb is None else fb.ctypes.data,
                                        0 if fb is None else fb.shape[1], out.ctypes.data)
        assert rc == T
        return out

    def amp_to_db_default(self, x):
        a = np.ascontiguousarray(x, np.float32).copy()
        rc = self.lib.orc_amp_to_db_default(a.ctypes.data, a.size)
        if rc:
            raise ValueError("negative or NaN input")
        return a

    # ---- range / display --------------------------------------------------------------------
    def clamp_range(self, gmax, gmin, db_range=120.0):
        a, b = _f(), _f()
        self.lib.orc_clamp_range(gmax, gmin, db_range, C.byref(a), C.byref(b))
        return a.value, b.value

    def spec_max_min(self, spec):
        s = np.ascontiguousarray(spec, np.float32)
        a, b = _f(), _f()
        self.lib.orc_spec_max_min(s.ctypes.data, s.size, C.byref(a), C.byref(b))
        return a.value, b.value

    def up_ratio(self, max_sr, sr, mel=True):
        return self.lib.orc_up_ratio(max_sr, sr, int(mel))

    def calc_nwidth(self, px_per_sec, n, sr):
        return self.lib.orc_calc_nwidth(px_per_sec, n, sr)

    def spec_to_grey(self, spec, up_ratio, mx, mn):
        s = np.ascontiguousarray(spec, np.float32)
        T, n_out = s.shape
        h = self.lib.orc_grey_height(n_out, up_ratio)
        out = np.empty((h, T), np.float32)
        self.lib.orc_spec_to_grey(s.ctypes.data, T, n_out, up_ratio, mx, mn, out.ctypes.data)
        return out

    def resize_lanczos3(self, grey, nwidth, nheight, parallel=True):
        g = np.ascontiguousarray(grey, np.float32)
        h, w = g.shape
        out = np.empty((nheight, nwidth), np.float32)
        assert self.lib.orc_resize_lanczos3(g.ctypes.data, w, h, nwidth, nheight, out.ctypes.data, int(parallel)) == 0
        return out

    def grey_to_rgb(self, grey, nwidth, nheight, channels=3, parallel=True):
        g = np.ascontiguousarray(grey, np.float32)
        h, w = g.shape
        out = np.empty((nheight, nwidth, channels), np.uint8)
        rc = self.lib.orc_grey_to_rgb(g.ctypes.data, w, h, nwidth, nheight, channels, out.ctypes.data, int(parallel))
        assert rc == 0
        return out

    def convert_grey_to_color(self, x):
        out = np.empty(3, np.uint8)
        rc = self.lib.orc_convert_grey_to_color(float(x), out.ctypes.data)
        if rc:
            raise ValueError("assert x >= 0")
        return out

    def colormap(self):
        out = np.empty(30, np.uint8)
        self.lib.orc_get_colormap(out.ctypes.data)
        return out

    def wav_to_image(self, wav, nwidth, nheight, amp_min=-1.0, amp_max=1.0):
        w = np.ascontiguousarray(wav, np.float32)
        out = np.empty((nheight, nwidth, 4), np.uint8)
        rc = self.lib.orc_wav_to_image(w.ctypes.data, w.size, nwidth, nheight, amp_min, amp_max, out.ctypes.data)
        if rc:
            raise ValueError("reference would panic")
        return out

    def num_threads(self):
        return self.lib.orc_num_threads()

    def set_num_threads(self, n):
        self.lib.orc_set_num_threads(int(n))

    # ---- whole pipeline (MultiTrack::add_tracks + get_spec_image for all tracks) ---------------
    def pipeline(self, wavs, srs, params, windows, mel_fbs, mel_scale=True, db_range=120.0, px_per_sec=100.0, nheight=500,
                 channels=3, dense_mel=False, parallel_render=True, render=True, out_imgs=None):
        """wavs: list of mono f32 arrays; params: list of (win, hop, n_fft); returns (images, max_db, min_db).
        out_imgs: images of a previous call to write into (a timed loop then allocates nothing)."""
        n = len(wavs)
        wavs = [np.ascontiguousarray(w, np.float32) for w in wavs]
        windows = [np.ascontiguousarray(w, np.float32) for w in windows]
        fbs = [None if f is None else np.ascontiguousarray(f, np.float32) for f in mel_fbs]
        PP = _vp * n
        SZ = _sz * n
        U32 = _u32 * n
        imgs = [] if out_imgs is None else out_imgs
        if out_imgs is None:
            for w, sr in zip(wavs, srs):
                nw = self.calc_nwidth(px_per_sec, w.size, sr)
                imgs.append(np.zeros((nheight, nw, channels), np.uint8))
        a, b = _f(), _f()
        use_mel = any(f is not None for f in fbs)
        self.lib.orc_pipeline.argtypes = [_sz, PP, SZ, U32, SZ, SZ, SZ, PP, _vp, _vp, C.c_int, _f, _f, _u32, C.c_int, C.c_int,
                                          C.c_int, _vp, C.POINTER(_f), C.POINTER(_f)]
        fb_ptrs = PP(*[None if f is None else f.ctypes.data for f in fbs])
        nm = SZ(*[0 if f is None else f.shape[1] for f in fbs])
        img_ptrs = PP(*[im.ctypes.data for im in imgs])
        rc = self.lib.orc_pipeline(n, PP(*[w.ctypes.data for w in wavs]), SZ(*[w.size for w in wavs]), U32(*srs),
                                   SZ(*[p[0] for p in params]), SZ(*[p[1] for p in params]), SZ(*[p[2] for p in params]),
                                   PP(*[w.ctypes.data for w in windows]),
                                   C.cast(fb_ptrs, _vp) if use_mel else None, C.cast(nm, _vp) if use_mel else None,
                                   int(mel_scale), db_range, px_per_sec, nheight, channels, int(dense_mel), int(parallel_render),
                                   C.cast(img_ptrs, _vp) if render else None, C.byref(a), C.byref(b))
        if rc:
            raise ValueError("oracle pipeline failed")
        return imgs, a.value, b.value


_cached = None


def load():
    global _cached
    if _cached is None:
        _cached = Oracle(C.CDLL(build()))
    return _cached

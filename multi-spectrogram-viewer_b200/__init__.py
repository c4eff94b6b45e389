"""Host-side mirror of the reference's public surface on top of libsgx.so (ctypes over the C ABI).

The directory name contains a hyphen, so import it through the repo-root shim::

    import msv_b200 as msv
    mt = msv.MultiTrack()
    mt.add_tracks([0, 1], "a.wav\\nb.wav")
    rgb = mt.get_spec_image(0, 100.0, 500)

Names, argument order and error behaviour follow ``src_rust/lib.rs`` (``MultiTrack``) and the
rlib functions ``benches/bench.rs`` imports (``perform_stft``, ``mel::calc_mel_fb_default``,
``windows::hann``, ``decibel::amp_to_db_default``, ``display::{spec_to_grey, grey_to_rgb}``).
Where the reference panics, :class:`SgxError` is raised.  There is no CPU fallback: importing
this module without the built CUDA library fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsgx.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(nvcc, sm_100a).  This engine has no CPU or PyTorch fallback."
    )
_lib = C.CDLL(LIB_PATH)

SGX_OK, SGX_ERR_IO, SGX_ERR_UNKNOWN_ID, SGX_ERR_BAD_ARG, SGX_ERR_CUDA, SGX_ERR_STATE, SGX_ERR_NOMEM, SGX_ERR_BUFFER = range(8)
FREQ_LINEAR, FREQ_MEL = 0, 1


class SgxError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[sgx {code}] {msg}")
        self.code = code


class Settings(C.Structure):
    """struct SpecSetting (lib.rs:64-70) plus the explicit overrides of include/sgx.h."""

    _fields_ = [
        ("win_ms", C.c_float), ("t_overlap", C.c_size_t), ("f_overlap", C.c_size_t),
        ("freq_scale", C.c_int), ("db_range", C.c_float), ("win_length", C.c_size_t),
        ("hop_length", C.c_size_t), ("n_fft", C.c_size_t), ("n_mel", C.c_size_t),
    ]

    @classmethod
    def default(cls, **over) -> "Settings":
        s = cls()
        _lib.sgx_settings_default(C.byref(s))
        for k, v in over.items():
            setattr(s, k, v)
        return s


_vp, _sz, _u32, _f = C.c_void_p, C.c_size_t, C.c_uint32, C.c_float
_pf = C.POINTER(C.c_float)
_psz = C.POINTER(C.c_size_t)
_pi = C.POINTER(C.c_int)
_pu32 = C.POINTER(C.c_uint32)
_pu8 = C.POINTER(C.c_uint8)

# every exported symbol of include/sgx.h with its prototype (tests check this list against the header)
PROTOTYPES = {
    "sgx_last_error": (C.c_char_p, []),
    "sgx_device_info": (C.c_int, [C.c_int, _pi, _pi, _pi, _psz]),
    "sgx_kernel_launch_count": (C.c_uint64, []),
    "sgx_host_pin": (C.c_int, [_vp, _sz]),
    "sgx_host_unpin": (C.c_int, [_vp]),
    "sgx_settings_default": (None, [C.POINTER(Settings)]),
    "sgx_mt_new": (C.c_int, [C.POINTER(_vp)]),
    "sgx_mt_new_ex": (C.c_int, [C.POINTER(Settings), C.c_int, _vp, C.POINTER(_vp)]),
    "sgx_mt_free": (None, [_vp]),
    "sgx_mt_new_sharded": (C.c_int, [C.POINTER(Settings), _pi, _sz, C.POINTER(_vp)]),
    "sgx_nccl_unique_id": (C.c_int, [_pu8]),
    "sgx_mt_attach_nccl": (C.c_int, [_vp, _pu8, C.c_int, C.c_int]),
    "sgx_mt_get_device_count": (C.c_int, [_vp, _pi, _pi, _pi]),
    "sgx_mt_add_tracks": (C.c_int, [_vp, _psz, _sz, C.c_char_p, _pi]),
    "sgx_mt_add_tracks_pcm": (C.c_int, [_vp, _psz, _sz, C.POINTER(_vp), _psz, _pu32, _pu32, _pi]),
    "sgx_mt_add_tracks_pcm_i16": (C.c_int, [_vp, _psz, _sz, C.POINTER(_vp), _psz, _pu32, _pu32, _pi]),
    "sgx_mt_add_tracks_pcm_device": (C.c_int, [_vp, _psz, _sz, C.POINTER(_vp), _psz, _pu32, _pu32, _pi]),
    "sgx_mt_remove_track": (C.c_int, [_vp, _sz, _pi]),
    "sgx_mt_get_spec_image": (C.c_int, [_vp, _sz, _f, _u32, _vp, _sz, _psz]),
    "sgx_mt_get_spec_image_rgba": (C.c_int, [_vp, _sz, _f, _u32, _vp, _sz, _psz]),
    "sgx_mt_get_spec_image_device": (C.c_int, [_vp, _sz, _f, _u32, C.c_int, _vp, _sz, _psz]),
    "sgx_mt_get_spec_images_device": (C.c_int, [_vp, _psz, _sz, _f, _u32, C.c_int, C.POINTER(_vp), _psz, _psz]),
    "sgx_mt_get_spec_images": (C.c_int, [_vp, _psz, _sz, _f, _u32, C.c_int, C.POINTER(_vp), _psz, _psz]),
    "sgx_mt_get_spec_images_async": (C.c_int, [_vp, _psz, _sz, _f, _u32, C.c_int, C.POINTER(_vp), _psz, _psz]),
    "sgx_mt_wait_images": (C.c_int, [_vp]),
    "sgx_mt_get_wav_image": (C.c_int, [_vp, _sz, _f, _u32, _f, _f, _vp, _sz, _psz]),
    "sgx_mt_get_frequency_hz": (C.c_int, [_vp, _sz, _f, _pf]),
    "sgx_mt_get_max_db": (C.c_int, [_vp, _pf]),
    "sgx_mt_get_min_db": (C.c_int, [_vp, _pf]),
    "sgx_mt_get_max_sec": (C.c_int, [_vp, _pf]),
    "sgx_mt_get_sec": (C.c_int, [_vp, _sz, _pf]),
    "sgx_mt_get_sr": (C.c_int, [_vp, _sz, _pu32]),
    "sgx_mt_get_path": (C.c_int, [_vp, _sz, C.c_char_p, _sz, _psz]),
    "sgx_mt_get_filename": (C.c_int, [_vp, _sz, C.c_char_p, _sz, _psz]),
    "sgx_get_colormap": (C.c_int, [_pu8]),
    "sgx_mt_get_spec_shape": (C.c_int, [_vp, _sz, _psz, _psz]),
    "sgx_mt_get_spec_db": (C.c_int, [_vp, _sz, _vp, _sz, _psz]),
    "sgx_mt_get_image_width": (C.c_int, [_vp, _sz, _f, _pu32]),
    "sgx_mt_range_device_ptr": (C.c_int, [_vp, C.POINTER(_vp)]),
    "sgx_mt_commit_range_device": (C.c_int, [_vp]),
    "sgx_mt_set_global_max_sr": (C.c_int, [_vp, _u32]),
    "sgx_slice_plan": (C.c_int, [_sz, _u32, C.POINTER(Settings), _f, _u32, _u32, _psz, _psz, _psz, _psz]),
    "sgx_mt_add_track_slice_device": (C.c_int, [_vp, _sz, _vp, _sz, _sz, _sz, _u32, _u32, _sz, _sz]),
    "sgx_mt_get_spec_image_slice_device": (C.c_int, [_vp, _sz, _f, _u32, C.c_int, _u32, _u32, _vp, _sz, _psz]),
    "sgx_mt_set_profiling": (C.c_int, [_vp, C.c_int]),
    "sgx_mt_get_stage_times": (C.c_int, [_vp, _pf, _pf]),
    "sgx_mt_synchronize": (C.c_int, [_vp, _pi]),
    "sgx_calc_proper_n_fft": (_sz, [_sz]),
    "sgx_track_params": (C.c_int, [_u32, C.POINTER(Settings), _psz, _psz, _psz]),
    "sgx_hann": (C.c_int, [_sz, C.c_int, _vp]),
    "sgx_calc_window": (C.c_int, [_sz, _sz, _vp]),
    "sgx_hz_to_mel": (_f, [_f]),
    "sgx_mel_to_hz": (_f, [_f]),
    "sgx_calc_mel_fb": (C.c_int, [_u32, _sz, _sz, _f, _f, C.c_int, _vp]),
    "sgx_calc_mel_fb_default": (C.c_int, [_u32, _sz, _vp, _sz, _psz]),
    "sgx_stft_num_frames": (C.c_long, [_sz, _sz, _sz]),
    "sgx_perform_stft": (C.c_int, [_vp, _sz, _sz, _sz, _sz, _vp, _vp, _sz, _psz]),
    "sgx_stft_magnitude": (C.c_int, [_vp, _sz, _sz, _sz, _sz, _vp, _vp, _sz, _psz]),
    "sgx_amp_to_db_default": (C.c_int, [_vp, _sz]),
    "sgx_melspectrogram_db": (C.c_int, [_vp, _sz, _sz, _sz, _sz, _vp, _vp, _sz, _vp, _sz, _psz]),
    "sgx_spec_to_grey": (C.c_int, [_vp, _sz, _sz, _f, _f, _f, _vp, _sz, _pu32]),
    "sgx_grey_to_rgb": (C.c_int, [_vp, _u32, _u32, _u32, _u32, C.c_int, _vp, _sz]),
    "sgx_wav_to_image": (C.c_int, [_vp, _sz, _u32, _u32, _f, _f, _vp, _sz]),
    "sgx_open_wav": (C.c_int, [C.c_char_p, _vp, _sz, _psz, _pu32, _pu32]),
}
for _name, (_res, _args) in PROTOTYPES.items():
    _fn = getattr(_lib, _name)  # AttributeError here == symbol missing from the build
    _fn.restype = _res
    _fn.argtypes = _args


def _check(code: int) -> None:
    if code != SGX_OK:
        raise SgxError(code, _lib.sgx_last_error().decode("utf-8", "replace"))


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_vp)


def kernel_launch_count() -> int:
    return int(_lib.sgx_kernel_launch_count())


def device_info(device: int = 0) -> dict:
    sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
    _check(_lib.sgx_device_info(device, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)))
    return {"sm_count": sm.value, "cc": (ma.value, mi.value), "total_mem": mem.value}


# --------------------------------------------------------------------------------------------
# surface 2: stage functions (bench.rs:5)
# --------------------------------------------------------------------------------------------
def calc_proper_n_fft(win_length: int) -> int:
    """utils::calc_proper_n_fft (utils.rs:17-19)."""
    return int(_lib.sgx_calc_proper_n_fft(win_length))


def track_params(sr: int, settings: Optional[Settings] = None):
    """AudioTrack::new parameter derivation (lib.rs:43-46) -> (win_length, hop_length, n_fft)."""
    w, h, f = C.c_size_t(), C.c_size_t(), C.c_size_t()
    _check(_lib.sgx_track_params(sr, C.byref(settings) if settings is not None else None, C.byref(w), C.byref(h), C.byref(f)))
    return w.value, h.value, f.value


def hann(size: int, symmetric: bool = False) -> np.ndarray:
    """windows::hann (windows.rs:21-30)."""
    out = np.empty(size, np.float32)
    _check(_lib.sgx_hann(size, int(symmetric), _ptr(out)))
    return out


def calc_window(win_length: int, n_fft: int) -> np.ndarray:
    """MultiTrack::calc_window (lib.rs:138-140)."""
    out = np.empty(win_length, np.float32)
    _check(_lib.sgx_calc_window(win_length, n_fft, _ptr(out)))
    return out


def hz_to_mel(hz: float) -> float:
    return float(_lib.sgx_hz_to_mel(hz))


def mel_to_hz(mel: float) -> float:
    return float(_lib.sgx_mel_to_hz(mel))


def calc_mel_fb(sr: int, n_fft: int, n_mel: int, fmin: float = 0.0, fmax: Optional[float] = None, do_norm: bool = True) -> np.ndarray:
    """mel::calc_mel_fb::<f32> (mel.rs:33-85) -> [n_fft/2+1, n_mel]."""
    out = np.empty((n_fft // 2 + 1, max(n_mel, 1)), np.float32)
    _check(_lib.sgx_calc_mel_fb(sr, n_fft, n_mel, fmin, -1.0 if fmax is None else fmax, int(do_norm), _ptr(out)))
    return out


def calc_mel_fb_default(sr: int, n_fft: int) -> np.ndarray:
    """mel::calc_mel_fb_default (mel.rs:87-99)."""
    n_mel = C.c_size_t()
    _check(_lib.sgx_calc_mel_fb_default(sr, n_fft, None, 0, C.byref(n_mel)))
    out = np.empty((n_fft // 2 + 1, n_mel.value), np.float32)
    _check(_lib.sgx_calc_mel_fb_default(sr, n_fft, _ptr(out), out.size, C.byref(n_mel)))
    return out


def stft_num_frames(n: int, win_length: int, hop_length: int) -> int:
    return int(_lib.sgx_stft_num_frames(n, win_length, hop_length))


def _stft_like(fn, wav, win_length, hop_length, n_fft, window, complex_out: bool):
    wav = _f32(wav)
    w = None if window is None else _f32(window)
    if w is not None and w.size != win_length:
        raise SgxError(SGX_ERR_BAD_ARG, "window length != win_length (assert_eq at lib.rs:404)")
    T = C.c_size_t()
    _check(fn(_ptr(wav), wav.size, win_length, hop_length, n_fft, _ptr(w), None, 0, C.byref(T)))
    B = n_fft // 2 + 1
    out = np.empty((T.value, B, 2) if complex_out else (T.value, B), np.float32)
    _check(fn(_ptr(wav), wav.size, win_length, hop_length, n_fft, _ptr(w), _ptr(out), out.size, C.byref(T)))
    return out


def perform_stft(wav, win_length: int, hop_length: int, n_fft: int, window=None, fft_module=None, parallel: bool = False) -> np.ndarray:
    """perform_stft (lib.rs:388-471) -> complex64 [T, n_fft/2+1].  fft_module / parallel are accepted
    for signature compatibility and ignored (the GPU has one plan per size and is always parallel)."""
    out = _stft_like(_lib.sgx_perform_stft, wav, win_length, hop_length, n_fft, window, True)
    return out.view(np.complex64)[..., 0]


def stft_magnitude(wav, win_length: int, hop_length: int, n_fft: int, window=None) -> np.ndarray:
    """stft.mapv(|x| x.norm()) (lib.rs:124)."""
    return _stft_like(_lib.sgx_stft_magnitude, wav, win_length, hop_length, n_fft, window, False)


def amp_to_db_default(x) -> np.ndarray:
    """DeciBelInplace::amp_to_db_default (decibel.rs:79-88); returns a new array."""
    a = _f32(x).copy()
    _check(_lib.sgx_amp_to_db_default(_ptr(a), a.size))
    return a


def melspectrogram_db(wav, win_length: int, hop_length: int, n_fft: int, window=None, mel_fb=None) -> np.ndarray:
    """get_melspectrogram of bench.rs:7-25 (mel_fb=None: the Linear branch of lib.rs:126-129)."""
    wav = _f32(wav)
    w = None if window is None else _f32(window)
    if w is not None and w.size != win_length:
        raise SgxError(SGX_ERR_BAD_ARG, "window length != win_length (assert_eq at lib.rs:404)")
    fb = None if mel_fb is None else _f32(mel_fb)
    n_mel = 0 if fb is None else fb.shape[1]
    if fb is not None and fb.shape[0] != n_fft // 2 + 1:
        raise SgxError(SGX_ERR_BAD_ARG, "mel_fb must be [n_fft/2+1, n_mel]")
    T = C.c_size_t()
    _check(_lib.sgx_melspectrogram_db(_ptr(wav), wav.size, win_length, hop_length, n_fft, _ptr(w), _ptr(fb), n_mel, None, 0, C.byref(T)))
    out = np.empty((T.value, n_mel if fb is not None else n_fft // 2 + 1), np.float32)
    _check(_lib.sgx_melspectrogram_db(_ptr(wav), wav.size, win_length, hop_length, n_fft, _ptr(w), _ptr(fb), n_mel, _ptr(out), out.size, C.byref(T)))
    return out


def spec_to_grey(spec, up_ratio: float, max_db: float, min_db: float) -> np.ndarray:
    """display::spec_to_grey (display.rs:44-54) -> grey [height, T] (row-major image)."""
    spec = _f32(spec)
    T, n_out = spec.shape
    h = C.c_uint32()
    _check(_lib.sgx_spec_to_grey(_ptr(spec), T, n_out, up_ratio, max_db, min_db, None, 0, C.byref(h)))
    out = np.empty((h.value, T), np.float32)
    _check(_lib.sgx_spec_to_grey(_ptr(spec), T, n_out, up_ratio, max_db, min_db, _ptr(out), out.size, C.byref(h)))
    return out


def grey_to_rgb(grey, nwidth: int, nheight: int, channels: int = 3) -> np.ndarray:
    """display::grey_to_rgb (display.rs:56-61) -> uint8 [nheight, nwidth, channels]."""
    grey = _f32(grey)
    height, width = grey.shape
    out = np.empty((nheight, nwidth, channels), np.uint8)
    _check(_lib.sgx_grey_to_rgb(_ptr(grey), width, height, nwidth, nheight, channels, _ptr(out), out.size))
    return out


def wav_to_image(wav, nwidth: int, nheight: int, amp_range=(-1.0, 1.0)) -> np.ndarray:
    """display::wav_to_image (display.rs:63-115) -> uint8 RGBA [nheight, nwidth, 4]."""
    wav = _f32(wav)
    out = np.empty((nheight, nwidth, 4), np.uint8)
    _check(_lib.sgx_wav_to_image(_ptr(wav), wav.size, nwidth, nheight, amp_range[0], amp_range[1], _ptr(out), out.size))
    return out


def open_audio_file(path: str):
    """audio::open_audio_file (audio.rs:9-37, WAV only) -> (float32 [ch, n], sr)."""
    n, ch, sr = C.c_size_t(), C.c_uint32(), C.c_uint32()
    _check(_lib.sgx_open_wav(path.encode(), None, 0, C.byref(n), C.byref(ch), C.byref(sr)))
    buf = np.empty((n.value, ch.value), np.float32)
    _check(_lib.sgx_open_wav(path.encode(), _ptr(buf), buf.size, C.byref(n), C.byref(ch), C.byref(sr)))
    return buf.T, sr.value  # the [ch, n] view over interleaved memory of audio.rs:33-35


def calc_nwidth_like(px_per_sec: float, n: int, sr: int) -> int:
    """nwidth of lib.rs:296: (px_per_sec * len as f32 / sr as f32) as u32, in f32 arithmetic."""
    v = np.float32(px_per_sec) * np.float32(n) / np.float32(sr)
    return int(v) if v > 0 else 0


def slice_plan(n_total: int, sr: int, px_per_sec: float, ox_begin: int, ox_count: int, settings: Optional[Settings] = None):
    """Frames and samples a strip of output columns needs -> (frame_begin, frame_count, sample_begin, sample_count)."""
    fb, fc, sb, sc = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_size_t()
    _check(_lib.sgx_slice_plan(n_total, sr, C.byref(settings) if settings is not None else None, px_per_sec, ox_begin, ox_count,
                               C.byref(fb), C.byref(fc), C.byref(sb), C.byref(sc)))
    return fb.value, fc.value, sb.value, sc.value


def get_colormap() -> np.ndarray:
    """get_colormap (lib.rs:473-480): 30 bytes."""
    out = np.empty(30, np.uint8)
    _check(_lib.sgx_get_colormap(out.ctypes.data_as(_pu8)))
    return out


# --------------------------------------------------------------------------------------------
# surface 1: MultiTrack (lib.rs:72-365)
# --------------------------------------------------------------------------------------------
def _size_array(v: Iterable[int]):
    v = list(v)
    return (C.c_size_t * len(v))(*v), len(v)


def nccl_unique_id() -> bytes:
    """128 bytes drawn on rank 0 and shipped to every rank before MultiTrack.attach_nccl (sgx_nccl_unique_id)."""
    buf = (C.c_uint8 * 128)()
    _check(_lib.sgx_nccl_unique_id(buf))
    return bytes(buf)


class MultiTrack:
    """Mirror of the reference's ``MultiTrack`` class.  ``device=d``: one GPU + stream.  ``devices=[...]`` (or
    ``devices="all"``): one handle over several GPUs of this process, track t on devices[t mod G]
    (sgx_mt_new_sharded).  ``attach_nccl`` joins the single-device handles of several processes."""

    def __init__(self, settings: Optional[Settings] = None, device: int = 0, stream: Optional[int] = None, devices=None):
        self._h = _vp()
        self.device = device
        self.settings = settings if settings is not None else Settings.default()
        if devices is not None:
            devs = [] if devices == "all" else [int(d) for d in devices]
            arr = (C.c_int * max(1, len(devs)))(*devs)
            _check(_lib.sgx_mt_new_sharded(C.byref(self.settings), arr if devs else None, len(devs), C.byref(self._h)))
        else:
            _check(_lib.sgx_mt_new_ex(C.byref(self.settings), device, _vp(stream) if stream else None, C.byref(self._h)))
        self._keep = {}  # id -> objects that must outlive the track (borrowed device tensors)
        self._retired = []  # keepalives replaced while work that may still read them was only enqueued (sync=False)

    def attach_nccl(self, unique_id: bytes, rank: int, world: int) -> None:
        """Collective over all ranks: this handle becomes shard `rank` of `world` (sgx_mt_attach_nccl)."""
        buf = (C.c_uint8 * 128)(*unique_id)
        _check(_lib.sgx_mt_attach_nccl(self._h, buf, rank, world))

    def device_count(self):
        """(engines inside the handle, rank, world)."""
        n, r, w = C.c_int(), C.c_int(), C.c_int()
        _check(_lib.sgx_mt_get_device_count(self._h, C.byref(n), C.byref(r), C.byref(w)))
        return n.value, r.value, w.value

    def _swap_keepalive(self, id, new, sync):
        old = self._keep.pop(id, None)
        if new is not None:
            self._keep[id] = new
        # enqueued kernels may still read the old buffer: it is released at the next synchronising call
        if old is not None and not sync:
            self._retired.append(old)
        if sync:
            self._retired.clear()

    def close(self) -> None:
        if getattr(self, "_h", None):
            _lib.sgx_mt_free(self._h)
            self._h = None
            self._keep = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- mutators ---------------------------------------------------------------------------
    def add_tracks(self, id_list: Sequence[int], path_list: str) -> bool:
        """add_tracks(&mut self, id_list, path_list: '\\n'-joined) -> changed (lib.rs:171-191)."""
        ids, n = _size_array(id_list)
        ch = C.c_int()
        _check(_lib.sgx_mt_add_tracks(self._h, ids, n, path_list.encode(), C.byref(ch)))
        return bool(ch.value)

    def add_tracks_pcm(self, id_list: Sequence[int], pcm: Sequence[np.ndarray], sr: Sequence[int]) -> bool:
        """Same as add_tracks with decoded audio: pcm[i] is [n] or interleaved [n, ch]; float32 or int16."""
        arrs = []
        for a in pcm:
            if a is None:  # a track another rank owns (attached handles take the whole id list)
                arrs.append(None)
                continue
            a = np.asarray(a)
            if a.dtype != np.int16:
                a = a.astype(np.float32, copy=False)
            arrs.append(np.ascontiguousarray(a))
        kinds = {a.dtype == np.int16 for a in arrs if a is not None}
        if len(kinds) > 1:
            raise SgxError(SGX_ERR_BAD_ARG, "mix of int16 and float32 tracks in one call")
        fn = _lib.sgx_mt_add_tracks_pcm_i16 if kinds == {True} else _lib.sgx_mt_add_tracks_pcm
        ids, n = _size_array(id_list)
        ptrs = (_vp * n)(*[None if a is None else a.ctypes.data for a in arrs])
        ns, _ = _size_array(0 if a is None else a.shape[0] for a in arrs)
        srs = (C.c_uint32 * n)(*[int(s) for s in sr])
        chs = (C.c_uint32 * n)(*[0 if a is None else (1 if a.ndim == 1 else a.shape[1]) for a in arrs])
        ch = C.c_int()
        _check(fn(self._h, ids, n, ptrs, ns, srs, chs, C.byref(ch)))
        return bool(ch.value)

    def add_tracks_device(self, id_list: Sequence[int], ptrs: Sequence[int], n_samples: Sequence[int], sr: Sequence[int],
                          channels: Optional[Sequence[int]] = None, keepalive=None, sync: bool = True) -> Optional[bool]:
        """PCM already resident in HBM (raw device pointers, float32 interleaved).  sync=False enqueues only."""
        ids, n = _size_array(id_list)
        p = (_vp * n)(*[int(x) for x in ptrs])
        ns, _ = _size_array(n_samples)
        srs = (C.c_uint32 * n)(*[int(s) for s in sr])
        chs = (C.c_uint32 * n)(*([1] * n if channels is None else [int(c) for c in channels]))
        ch = C.c_int()
        _check(_lib.sgx_mt_add_tracks_pcm_device(self._h, ids, n, p, ns, srs, chs, C.byref(ch) if sync else None))
        for i in id_list:
            self._swap_keepalive(i, keepalive, sync)
        return bool(ch.value) if sync else None

    def add_track_slice_device(self, id: int, ptr: int, chunk_offset: int, chunk_len: int, n_total: int, sr: int, channels: int,
                               frame_begin: int, frame_count: int, keepalive=None) -> None:
        """One time slice of a long track (device PCM, deferred; see sgx_mt_add_track_slice_device)."""
        _check(_lib.sgx_mt_add_track_slice_device(self._h, id, _vp(int(ptr)), chunk_offset, chunk_len, n_total, sr, channels,
                                                  frame_begin, frame_count))
        self._swap_keepalive(id, keepalive, False)

    def render_slice_device(self, id: int, px_per_sec: float, nheight: int, channels: int, ox_begin: int, ox_count: int,
                            out_ptr: int, cap: int) -> int:
        wr = C.c_size_t()
        _check(_lib.sgx_mt_get_spec_image_slice_device(self._h, id, px_per_sec, nheight, channels, ox_begin, ox_count,
                                                       _vp(int(out_ptr)), cap, C.byref(wr)))
        return wr.value

    def remove_track(self, id: int, sync: bool = True) -> Optional[bool]:
        ch = C.c_int()
        _check(_lib.sgx_mt_remove_track(self._h, id, C.byref(ch) if sync else None))
        self._swap_keepalive(id, None, sync)
        return bool(ch.value) if sync else None

    # -- images -------------------------------------------------------------------------------
    def _image(self, fn, id, px_per_sec, nheight, channels, *extra):
        need = C.c_size_t()
        _check(fn(self._h, id, px_per_sec, nheight, *extra, None, 0, C.byref(need)))
        out = np.empty(need.value, np.uint8)
        if need.value:
            _check(fn(self._h, id, px_per_sec, nheight, *extra, _ptr(out), out.size, C.byref(need)))
        return out

    def get_spec_image(self, id: int, px_per_sec: float, nheight: int) -> np.ndarray:
        """get_spec_image -> flat uint8 RGB, nheight*nwidth*3 (lib.rs:294-298)."""
        return self._image(_lib.sgx_mt_get_spec_image, id, px_per_sec, nheight, 3)

    def get_spec_image_rgba(self, id: int, px_per_sec: float, nheight: int) -> np.ndarray:
        return self._image(_lib.sgx_mt_get_spec_image_rgba, id, px_per_sec, nheight, 4)

    def get_wav_image(self, id: int, px_per_sec: float, nheight: int, amp_min: float, amp_max: float) -> np.ndarray:
        """get_wav_image -> flat uint8 RGBA (lib.rs:300-313)."""
        return self._image(_lib.sgx_mt_get_wav_image, id, px_per_sec, nheight, 4, amp_min, amp_max)

    def get_spec_images(self, id_list: Sequence[int], px_per_sec: float, nheight: int, channels: int = 3, out=None, wait: bool = True):
        """get_spec_image for a list of tracks in one call (sgx_mt_get_spec_images): renders and device->host copies
        are pipelined inside the library.  `out`: optional list of writable uint8 buffers (numpy arrays or objects with
        a data_ptr(): pinned torch tensors make the copies asynchronous); returns the list of flat images.
        wait=False returns after everything is enqueued -- call wait_images() before reading the buffers."""
        ids, n = _size_array(id_list)
        need = (C.c_size_t * n)()
        _check(_lib.sgx_mt_get_spec_images(self._h, ids, n, px_per_sec, nheight, channels, None, None, need))
        if out is None:
            out = [np.empty(need[i], np.uint8) for i in range(n)]
        ptrs = (_vp * n)(*[int(o.data_ptr()) if hasattr(o, "data_ptr") else o.ctypes.data for o in out])
        caps, _ = _size_array(int(o.numel()) if hasattr(o, "numel") else o.size for o in out)
        fn = _lib.sgx_mt_get_spec_images if wait else _lib.sgx_mt_get_spec_images_async
        _check(fn(self._h, ids, n, px_per_sec, nheight, channels, ptrs, caps, need))
        if not wait:
            self._out_keep = out
        return out

    def wait_images(self) -> None:
        _check(_lib.sgx_mt_wait_images(self._h))
        self._out_keep = None

    def image_width(self, id: int, px_per_sec: float) -> int:
        w = C.c_uint32()
        _check(_lib.sgx_mt_get_image_width(self._h, id, px_per_sec, C.byref(w)))
        return w.value

    def render_device(self, id_list: Sequence[int], px_per_sec: float, nheight: int, channels: int, out_ptrs: Sequence[int], caps: Sequence[int]) -> None:
        """Batched, asynchronous render into device buffers (raw pointers)."""
        ids, n = _size_array(id_list)
        p = (_vp * n)(*[int(x) for x in out_ptrs])
        cp, _ = _size_array(caps)
        wr = (C.c_size_t * n)()
        _check(_lib.sgx_mt_get_spec_images_device(self._h, ids, n, px_per_sec, nheight, channels, p, cp, wr))

    # -- getters ------------------------------------------------------------------------------
    def _getf(self, fn, *a) -> float:
        v = C.c_float()
        _check(fn(self._h, *a, C.byref(v)))
        return v.value

    def get_frequency_hz(self, id: int, relative_freq: float) -> float:
        return self._getf(_lib.sgx_mt_get_frequency_hz, id, relative_freq)

    def get_max_db(self) -> float:
        return self._getf(_lib.sgx_mt_get_max_db)

    def get_min_db(self) -> float:
        return self._getf(_lib.sgx_mt_get_min_db)

    def get_max_sec(self) -> float:
        return self._getf(_lib.sgx_mt_get_max_sec)

    def get_sec(self, id: int) -> float:
        return self._getf(_lib.sgx_mt_get_sec, id)

    def get_sr(self, id: int) -> int:
        v = C.c_uint32()
        _check(_lib.sgx_mt_get_sr(self._h, id, C.byref(v)))
        return v.value

    def _gets(self, fn, id) -> str:
        need = C.c_size_t()
        _check(fn(self._h, id, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        _check(fn(self._h, id, buf, need.value, C.byref(need)))
        return buf.value.decode("utf-8", "replace")

    def get_path(self, id: int) -> str:
        return self._gets(_lib.sgx_mt_get_path, id)

    def get_filename(self, id: int) -> str:
        return self._gets(_lib.sgx_mt_get_filename, id)

    # -- engine-side extensions ---------------------------------------------------------------
    def spec_shape(self, id: int):
        t, m = C.c_size_t(), C.c_size_t()
        _check(_lib.sgx_mt_get_spec_shape(self._h, id, C.byref(t), C.byref(m)))
        return t.value, m.value

    def get_spec_db(self, id: int) -> np.ndarray:
        """The cached dB spectrogram [T, n_out] (== MultiTrack.specs[id], lib.rs:78)."""
        t, m = self.spec_shape(id)
        out = np.empty((t, m), np.float32)
        wr = C.c_size_t()
        _check(_lib.sgx_mt_get_spec_db(self._h, id, _ptr(out), out.size, C.byref(wr)))
        return out

    def range_device_ptr(self) -> int:
        p = _vp()
        _check(_lib.sgx_mt_range_device_ptr(self._h, C.byref(p)))
        return int(p.value)

    def commit_range_device(self) -> None:
        _check(_lib.sgx_mt_commit_range_device(self._h))

    def set_global_max_sr(self, max_sr: int) -> None:
        _check(_lib.sgx_mt_set_global_max_sr(self._h, max_sr))

    def set_profiling(self, on: bool) -> None:
        _check(_lib.sgx_mt_set_profiling(self._h, int(on)))

    def stage_times(self):
        """(analysis_ms, render_ms) of the most recent add / render calls (CUDA events on the stream)."""
        a, r = C.c_float(), C.c_float()
        _check(_lib.sgx_mt_get_stage_times(self._h, C.byref(a), C.byref(r)))
        return a.value, r.value

    def synchronize(self) -> bool:
        ch = C.c_int()
        _check(_lib.sgx_mt_synchronize(self._h, C.byref(ch)))
        self._retired.clear()
        return bool(ch.value)


from .sharded import ShardedMultiTrack, shard_ids  # noqa: E402  (multi-GPU driver)

__all__ = [n for n in dir() if not n.startswith("_")]

"""The C-ABI library loads without a GPU and exports exactly what include/sgx.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sgx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"SGX_API\s+[\w\s\*]+?\b(sgx_\w+)\s*\(", src)))


def test_header_declares_the_two_surfaces():
    syms = header_symbols()
    # surface 1: every #[wasm_bindgen] method of lib.rs:87-365 + get_colormap
    for name in ["mt_new", "mt_add_tracks", "mt_remove_track", "mt_get_spec_image", "mt_get_wav_image", "mt_get_frequency_hz",
                 "mt_get_max_db", "mt_get_min_db", "mt_get_max_sec", "mt_get_sec", "mt_get_sr", "mt_get_path", "mt_get_filename",
                 "get_colormap"]:
        assert f"sgx_{name}" in syms
    # surface 2: what benches/bench.rs:5 imports
    for name in ["perform_stft", "amp_to_db_default", "calc_mel_fb_default", "hann", "spec_to_grey", "grey_to_rgb", "open_wav",
                 "melspectrogram_db"]:
        assert f"sgx_{name}" in syms


def test_library_exports_every_declared_symbol(msv):
    lib = ctypes.CDLL(msv.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 45
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/sgx.h but not exported"
    assert sorted(msv.PROTOTYPES) == syms  # the ctypes mirror binds the same set
    out = subprocess.run(["nm", "-D", "--defined-only", msv.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (sgx_\w+)", out)))
    assert exported == syms, "library exports symbols that the header does not declare (or vice versa)"


def test_library_contains_sm100a_code(msv):
    exe = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-lelf", msv.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_gpu_means_loud_failure_not_fallback(msv):
    """Without a CUDA device every compute entry point must fail with SGX_ERR_CUDA (no CPU path)."""
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(msv.SgxError) as e:
        msv.perform_stft([0, 0, 1, 0], 4, 2, 4)
    assert e.value.code == msv.SGX_ERR_CUDA
    with pytest.raises(msv.SgxError) as e:
        msv.MultiTrack()
    assert e.value.code == msv.SGX_ERR_CUDA
    with pytest.raises(msv.SgxError):
        msv.amp_to_db_default([1.0, 2.0])


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may import, link or load it."""
    pkg = os.path.join(ROOT, "multi-spectrogram-viewer_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h", ".cuh")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in txt and "oracle_binding" not in txt and "thesia_oracle" not in txt, f
    out = subprocess.run(["ldd", os.path.join(pkg, "libsgx.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_wav_reader_and_errors(msv, tmp_path):
    import wave

    import numpy as np

    p = tmp_path / "t.wav"
    data = (np.arange(-50, 50, dtype=np.int16) * 300).reshape(50, 2)
    with wave.open(str(p), "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(22050); w.writeframes(data.tobytes())
    x, sr = msv.open_audio_file(str(p))
    assert sr == 22050 and x.shape == (2, 50)
    assert np.array_equal(x.T, data.astype(np.float32) / np.float32(32768.0))  # audio.rs:16-19
    with pytest.raises(msv.SgxError) as e:
        msv.open_audio_file(str(tmp_path / "missing.wav"))
    assert e.value.code == msv.SGX_ERR_IO
    (tmp_path / "bad.wav").write_bytes(b"not a wav file at all")
    with pytest.raises(msv.SgxError) as e:
        msv.open_audio_file(str(tmp_path / "bad.wav"))
    assert e.value.code == msv.SGX_ERR_IO


def test_wav_reader_survives_hostile_headers(msv, tmp_path):
    """Chunk sizes come from the file: a data / fmt / unknown chunk that announces 4 GiB must neither allocate that
    much nor loop; a truncated data chunk yields the samples that are there (round-1 advisor finding)."""
    import struct

    import numpy as np

    fmt = struct.pack("<HHIIHH", 1, 1, 8000, 16000, 2, 16)
    body = np.arange(100, dtype=np.int16).tobytes()

    def wav(chunks):
        payload = b"WAVE" + b"".join(tag + struct.pack("<I", size) + data for tag, size, data in chunks)
        return b"RIFF" + struct.pack("<I", len(payload)) + payload

    p = tmp_path / "huge_data.wav"     # data chunk claims 0xFFFFFFFF bytes, holds 200
    p.write_bytes(wav([(b"fmt ", 16, fmt), (b"data", 0xFFFFFFFF, body)]))
    x, sr = msv.open_audio_file(str(p))
    assert sr == 8000 and x.shape == (1, 100) and np.array_equal(x[0], np.arange(100, dtype=np.float32) / 32768)
    for name, chunks in (("huge_fmt.wav", [(b"fmt ", 0xFFFFFFF0, fmt), (b"data", 200, body)]),
                         ("huge_junk.wav", [(b"JUNK", 0xFFFFFFFF, b"xx"), (b"fmt ", 16, fmt), (b"data", 200, body)])):
        q = tmp_path / name
        q.write_bytes(wav(chunks))
        with pytest.raises(msv.SgxError) as e:
            msv.open_audio_file(str(q))
        assert e.value.code == msv.SGX_ERR_IO


def test_reference_arm_does_not_import_the_gpu_package():
    """bench.py --impl reference times the CPU restatement only: neither msv_b200 nor libsgx.so may be loaded by it
    (round-1 review: the arm imported the package for three integers)."""
    import subprocess
    import sys
    code = ("import runpy, sys; sys.argv = ['bench.py', '--impl', 'reference', '--workload', 'c1', '--steps', '1', '--warmup', '0'];\n"
            "try:\n    runpy.run_path('bench.py', run_name='__main__')\nexcept SystemExit as e:\n    assert not e.code, e.code\n"
            "bad = [m for m in sys.modules if 'msv_b200' in m or 'multi-spectrogram-viewer_b200' in m]\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert not bad and 'libsgx' not in maps, (bad, 'libsgx' in maps)\nprint('REFERENCE ARM CLEAN')")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "REFERENCE ARM CLEAN" in r.stdout and '"impl": "reference"' in r.stdout, r.stdout[-800:] + r.stderr[-800:]


def test_rust_sys_crate_is_in_step_with_the_header():
    """bindings/rust/sgx-sys/src/lib.rs (uncompiled: no Rust toolchain here) is generated from include/sgx.h and
    declares every entry point the header does."""
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call([sys.executable, os.path.join(root, "tools", "gen_rust_sys.py"), "--check"])
    hdr = open(os.path.join(root, "include", "sgx.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    declared = set(re.findall(r"\b(sgx_[a-z0-9_]+)\s*\(", hdr))
    rs = open(os.path.join(root, "bindings", "rust", "sgx-sys", "src", "lib.rs")).read()
    bound = set(re.findall(r"pub fn (sgx_[a-z0-9_]+)\(", rs))
    assert declared == bound, (sorted(declared - bound), sorted(bound - declared))

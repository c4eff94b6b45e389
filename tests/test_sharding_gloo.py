"""The N>1 path on CPU: two gloo ranks shard a batch per file, exchange {max, -min} with one
all_reduce(MAX) and must arrive at the range a single process computes (lib.rs:194-209)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, n_tracks, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch
    import torch.distributed as dist

    import msv_b200 as msv
    import oracle_binding
    import synth

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    # the host-side plumbing of the sharded path: rank 0 draws the NCCL id (libsgx.so loads libnccl.so.2 at run
    # time), every rank receives the same 128 bytes
    uid = msv.sharded.broadcast_unique_id()
    # what the library does with them needs GPUs; the SEMANTICS of the exchange are replayed here on the CPU:
    # every rank analyses the tracks it owns (t mod world == rank), {max, -min, max_sr, max_sec} are max-reduced,
    # lib.rs:208-209 is applied to the result
    orc = oracle_binding.load()
    srs = [8000, 4000, 8000, 2000, 8000][:n_tracks]
    base = synth.base_clip(4 * 8000, 8000, 77)
    mine = msv.shard_ids(n_tracks, world, rank)
    lmax, lmin, lsr, lsec = -np.inf, np.inf, 0.0, 0.0
    for t in mine:
        sr = srs[t]
        x = synth.derive_track(base[: 4 * sr], t)
        win, hop, n_fft = msv.track_params(sr)
        spec = orc.calc_spec(x, win, hop, n_fft, None, msv.calc_mel_fb_default(sr, n_fft))
        lmax, lmin = max(lmax, float(spec.max())), min(lmin, float(spec.min()))
        lsr, lsec = max(lsr, float(sr)), max(lsec, len(x) / sr)
    v = torch.tensor([lmax, -lmin, lsr, lsec], dtype=torch.float32)
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
    mx, mn = msv.sharded.clamp_range(float(v[0]), -float(v[1]), 120.0)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([mx, mn, float(v[2]), float(v[3]), len(mine)], np.float64))
    with open(os.path.join(out_dir, f"uid{rank}.bin"), "wb") as f:
        f.write(uid)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ids(msv):
    assert msv.shard_ids(10, 4, 1) == [1, 5, 9]
    assert sorted(sum((msv.shard_ids(256, 8, r) for r in range(8)), [])) == list(range(256))
    assert [len(msv.shard_ids(256, 8, r)) for r in range(8)] == [32] * 8
    assert msv.shard_ids(3, 8, 5) == []
    with pytest.raises(ValueError):
        msv.shard_ids(4, 2, 2)


def test_two_rank_range_exchange(orc, msv, tmp_path):
    import torch.multiprocessing as mp

    import synth

    n_tracks, world = 5, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n_tracks, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert np.array_equal(r0[:4], r1[:4]) and r0[4] + r1[4] == n_tracks
    u0, u1 = (tmp_path / "uid0.bin").read_bytes(), (tmp_path / "uid1.bin").read_bytes()
    assert len(u0) == 128 and u0 == u1 and any(u0)
    # single-process answer
    srs = [8000, 4000, 8000, 2000, 8000]
    base = synth.base_clip(4 * 8000, 8000, 77)
    specs = []
    for t, sr in enumerate(srs):
        win, hop, n_fft = msv.track_params(sr)
        specs.append(orc.calc_spec(synth.derive_track(base[: 4 * sr], t), win, hop, n_fft, None, msv.calc_mel_fb_default(sr, n_fft)))
    want = orc.clamp_range(max(float(s.max()) for s in specs), min(float(s.min()) for s in specs), 120.0)
    assert (r0[0], r0[1]) == pytest.approx(want, abs=0)
    assert r0[2] == 8000 and r0[3] == 4.0   # max_sr and max_sec travel with the range
    # rank 1 holds only the 4 kHz and 2 kHz tracks: without the exchange its images would be laid out for 4 kHz
    assert max(srs[t] for t in msv.shard_ids(n_tracks, world, 1)) == 4000

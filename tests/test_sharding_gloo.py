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
    orc = oracle_binding.load()
    sr = 8000
    base = synth.base_clip(4 * sr, sr, 77)
    win, hop, n_fft = msv.track_params(sr)
    fb = msv.calc_mel_fb_default(sr, n_fft)
    mine = msv.shard_ids(n_tracks, world, rank)
    lmax, lmin = -np.inf, np.inf
    for t in mine:
        spec = orc.calc_spec(synth.derive_track(base, t), win, hop, n_fft, None, fb)
        lmax, lmin = max(lmax, float(spec.max())), min(lmin, float(spec.min()))
    rng = torch.tensor([lmax, -lmin], dtype=torch.float32)
    msv.sharded.all_reduce_range(rng)
    max_sr = msv.sharded.all_reduce_max_sr(sr if rank == 0 else sr // 2)
    mx, mn = msv.sharded.clamp_range(float(rng[0]), -float(rng[1]), 120.0)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([mx, mn, max_sr, len(mine)], np.float64))
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
    assert np.array_equal(r0[:3], r1[:3]) and r0[3] + r1[3] == n_tracks
    # single-process answer
    sr = 8000
    base = synth.base_clip(4 * sr, sr, 77)
    win, hop, n_fft = msv.track_params(sr)
    fb = msv.calc_mel_fb_default(sr, n_fft)
    specs = [orc.calc_spec(synth.derive_track(base, t), win, hop, n_fft, None, fb) for t in range(n_tracks)]
    want = orc.clamp_range(max(float(s.max()) for s in specs), min(float(s.min()) for s in specs), 120.0)
    assert (r0[0], r0[1]) == pytest.approx(want, abs=0) and r0[2] == sr
    # the loudest track (gain) is not on every rank: the exchange mattered
    per_track_max = [float(s.max()) for s in specs]
    assert int(np.argmax(per_track_max)) % world in (0, 1)

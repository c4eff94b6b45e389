"""Per-file sharding of a multi-track batch across the GPUs of one box.

One process per GPU (torch.distributed, NCCL).  Track ``t`` of the global batch lives on rank
``t mod world_size``; PCM, the dB spectrogram and the pixels of a track never leave their GPU.
The only exchange step of the path is the global dB range of ``update_spec_greys``
(lib.rs:194-209): every rank holds ``{max, -min}`` of its own tracks in device memory, one
``all_reduce(MAX)`` over those 8 bytes runs on the engine's stream, and the clamp of
lib.rs:208-209 is then applied on the device -- no host round trip between analysis and render.
``max_sr`` (lib.rs:220-224) is metadata and is max-reduced on the host.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np


def shard_ids(n_tracks: int, world_size: int, rank: int) -> List[int]:
    """Global track ids owned by ``rank``: t -> GPU t mod G."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    return list(range(rank, n_tracks, world_size))


def all_reduce_range(max_negmin, group=None):
    """In-place all-reduce(MAX) of a 2-element tensor {max, -min} (any device / backend)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(max_negmin, op=dist.ReduceOp.MAX, group=group)
    return max_negmin


def clamp_range(gmax: float, gmin: float, db_range: float):
    """lib.rs:208-209 on the host (used by tests and by CPU-side drivers)."""
    mx = np.float32(min(np.float32(gmax), np.float32(0.0)))
    mn = np.float32(max(np.float32(gmin), np.float32(mx - np.float32(db_range))))
    return float(mx), float(mn)


def all_reduce_max_sr(local_max_sr: int, group=None, device=None) -> int:
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return int(local_max_sr)
    t = torch.tensor([int(local_max_sr)], dtype=torch.int64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


class _DevicePtrView:
    """Zero-copy view of engine-owned device memory for torch (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {
            "shape": (n,), "typestr": "<f4", "data": (int(ptr), False), "version": 3, "strides": None,
        }


class ShardedMultiTrack:
    """The reference's MultiTrack over a rank-sharded batch.  Ids are GLOBAL ids; each rank only
    passes the tracks it owns.  Every rank must call add_tracks_device / remove_track collectively."""

    def __init__(self, settings=None, device: Optional[int] = None, group=None):
        import torch

        from . import MultiTrack

        self.torch = torch
        self.group = group
        self.device = torch.cuda.current_device() if device is None else device
        # one torch-owned stream carries our kernels AND the NCCL collective, so they are ordered on the
        # device without any host synchronisation
        self.stream = torch.cuda.Stream(device=self.device)
        self.mt = MultiTrack(settings, device=self.device, stream=self.stream.cuda_stream)
        self._range = torch.as_tensor(_DevicePtrView(self.mt.range_device_ptr(), 2), device=f"cuda:{self.device}")

    def add_tracks_device(self, id_list: Sequence[int], ptrs: Sequence[int], n_samples: Sequence[int], sr: Sequence[int],
                          channels: Optional[Sequence[int]] = None, keepalive=None, exchange_max_sr: bool = True) -> None:
        """Analysis of the local shard + the global range exchange; asynchronous."""
        self.stream.wait_stream(self.torch.cuda.current_stream(self.device))  # inputs produced on the caller's stream
        self.mt.add_tracks_device(id_list, ptrs, n_samples, sr, channels, keepalive=keepalive, sync=False)
        if exchange_max_sr:
            local = max([int(s) for s in sr], default=0)
            self.mt.set_global_max_sr(all_reduce_max_sr(local, self.group, device=f"cuda:{self.device}"))
        with self.torch.cuda.stream(self.stream):
            all_reduce_range(self._range, self.group)  # NCCL, 8 bytes, ordered on the engine's stream
        self.mt.commit_range_device()

    def add_track_time_sharded(self, id: int, pcm_host: np.ndarray, sr: int, px_per_sec: float, rank: int, world: int):
        """n3: ONE long track over `world` GPUs.  This rank takes a strip of output columns, uploads only the samples
        that strip's frames read, analyses them and joins the global range exchange.  pcm_host is the whole track
        ([n] or interleaved [n, ch] float32) -- only this rank's chunk is copied.  Returns (ox_begin, ox_count)."""
        from . import calc_nwidth_like, slice_plan

        torch = self.torch
        n_total = pcm_host.shape[0]
        ch = 1 if pcm_host.ndim == 1 else pcm_host.shape[1]
        nwidth = calc_nwidth_like(px_per_sec, n_total, sr)
        ob = nwidth * rank // world
        oc = nwidth * (rank + 1) // world - ob
        fb, fc, sb, sc = slice_plan(n_total, sr, px_per_sec, ob, oc, self.mt.settings)
        chunk = torch.from_numpy(np.ascontiguousarray(pcm_host[sb:sb + sc], dtype=np.float32)).to(f"cuda:{self.device}")
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        self.mt.add_track_slice_device(id, chunk.data_ptr(), sb, sc, n_total, sr, ch, fb, fc, keepalive=chunk)
        self.mt.set_global_max_sr(sr)
        with torch.cuda.stream(self.stream):
            all_reduce_range(self._range, self.group)
        self.mt.commit_range_device()
        return ob, oc

    def remove_track(self, id: int, owned: bool) -> None:
        if owned:
            self.mt.remove_track(id, sync=False)
        with self.torch.cuda.stream(self.stream):
            all_reduce_range(self._range, self.group)
        self.mt.commit_range_device()

    def render_device(self, id_list, px_per_sec, nheight, channels, out_ptrs, caps) -> None:
        self.mt.render_device(id_list, px_per_sec, nheight, channels, out_ptrs, caps)

    def synchronize(self) -> bool:
        return self.mt.synchronize()

    def get_max_db(self) -> float:
        return self.mt.get_max_db()

    def get_min_db(self) -> float:
        return self.mt.get_min_db()

    def close(self) -> None:
        self.mt.close()

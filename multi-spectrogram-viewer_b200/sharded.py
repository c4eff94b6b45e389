"""Per-file sharding of a multi-track batch across the GPUs of one box -- a thin caller.

The sharding itself lives behind the C ABI (include/sgx.h, "multi-GPU inside the library"): track ``t`` of the
global batch lives on shard ``t mod G``; PCM, the dB spectrogram and the pixels of a track never leave their GPU;
the one exchange of the path -- ``{max, -min, max_sr, max_sec}`` of ``update_spec_greys`` (lib.rs:194-209, 220-224) --
is an ``ncclAllReduce(MAX)`` that libsgx.so enqueues on the stream carrying the analysis and the render.

  * one process, several GPUs:  ``MultiTrack(devices=[0, 1, ...])``            (sgx_mt_new_sharded)
  * one process per GPU:        ``ShardedMultiTrack()`` below                   (sgx_mt_attach_nccl)

This module only ships the 128-byte NCCL id from rank 0 to the other ranks through ``torch.distributed`` (any host
channel would do) and keeps borrowed device tensors alive.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np


def shard_ids(n_tracks: int, world_size: int, rank: int) -> List[int]:
    """Global track ids owned by ``rank``: t -> GPU t mod G (the rule the library applies)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    return list(range(rank, n_tracks, world_size))


def clamp_range(gmax: float, gmin: float, db_range: float):
    """lib.rs:208-209 on the host (used by tests)."""
    mx = np.float32(min(np.float32(gmax), np.float32(0.0)))
    mn = np.float32(max(np.float32(gmin), np.float32(mx - np.float32(db_range))))
    return float(mx), float(mn)


def broadcast_unique_id(group=None, device=None) -> Optional[bytes]:
    """Rank 0 draws the NCCL id (sgx_nccl_unique_id), every rank receives it.  None when not distributed."""
    import torch
    import torch.distributed as dist

    from . import nccl_unique_id

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    on_gpu = dist.get_backend(group) == "nccl"
    dev = (device if device is not None else f"cuda:{torch.cuda.current_device()}") if on_gpu else "cpu"
    buf = torch.zeros(128, dtype=torch.uint8, device=dev)
    if dist.get_rank(group) == 0:
        buf.copy_(torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return bytes(buf.cpu().numpy().tobytes())


class ShardedMultiTrack:
    """The reference's MultiTrack over a rank-sharded batch, one process per GPU.  Ids are GLOBAL ids: a rank may
    pass the whole batch (entries of other ranks are ignored) or only the tracks it owns (t mod world == rank).
    add_tracks_device / remove_track are collective: every rank must call them."""

    def __init__(self, settings=None, device: Optional[int] = None, group=None):
        import torch
        import torch.distributed as dist

        from . import MultiTrack

        self.torch = torch
        self.group = group
        self.device = torch.cuda.current_device() if device is None else device
        # a torch-owned stream, so that tensors produced on torch's current stream can be ordered before our kernels
        self.stream = torch.cuda.Stream(device=self.device)
        self.mt = MultiTrack(settings, device=self.device, stream=self.stream.cuda_stream)
        self.rank, self.world = 0, 1
        uid = broadcast_unique_id(group, device=f"cuda:{self.device}")
        if uid is not None:
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
            self.mt.attach_nccl(uid, self.rank, self.world)

    def add_tracks_device(self, id_list: Sequence[int], ptrs: Sequence[int], n_samples: Sequence[int], sr: Sequence[int],
                          channels: Optional[Sequence[int]] = None, keepalive=None) -> None:
        """Analysis of this rank's share + the global range exchange (inside the library); asynchronous."""
        self.stream.wait_stream(self.torch.cuda.current_stream(self.device))  # inputs produced on the caller's stream
        self.mt.add_tracks_device(id_list, ptrs, n_samples, sr, channels, keepalive=keepalive, sync=False)
        if self.world == 1:
            self.mt.commit_range_device()  # a lone deferred handle commits its own range

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
        if self.world == 1:
            self.mt.commit_range_device()
        return ob, oc

    def remove_track(self, id: int) -> None:
        """Collective: the owner drops the track, every rank joins the exchange."""
        self.mt.remove_track(id, sync=False)
        if self.world == 1:
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

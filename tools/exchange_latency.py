"""Latency of the in-library range exchange (reduce + 16-byte ncclAllReduce + commit) on the engine's stream.
torchrun --nproc-per-node N tools/exchange_latency.py   (empty add_tracks calls: nothing but the exchange runs)"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import msv_b200 as msv  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sm = msv.ShardedMultiTrack(device=local)
    x = torch.randn(48000 * 5, device=f"cuda:{local}")
    sm.add_tracks_device(list(range(world)), [x.data_ptr()] * world, [x.numel()] * world, [48000] * world, [1] * world)
    sm.synchronize()
    for reps in (1, 200):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(sm.stream)
        for _ in range(reps):
            sm.add_tracks_device([], [], [], [], [])
        e1.record(sm.stream)
        sm.synchronize(); torch.cuda.synchronize()
        if reps > 1:
            print(f"rank {rank}: exchange (reduce + all-reduce + commit) {e0.elapsed_time(e1) / reps * 1e3:.1f} us per call over {reps} calls", flush=True)
    # the same with a torch.distributed all_reduce of 4 floats, for scale
    t = torch.zeros(4, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e1.record(); torch.cuda.synchronize()
    print(f"rank {rank}: torch.distributed all_reduce(4 floats) {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per call", flush=True)
    sm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

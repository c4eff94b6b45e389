"""Run under torchrun on N GPUs: every rank analyses its shard of a small batch, the ranks exchange the
dB range over NCCL in-stream, and each rank's pixels must equal what ONE process computes for the whole
batch (bit for bit: same kernels, same global range).  Not a pytest file (needs N GPUs)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def main():
    import torch
    import torch.distributed as dist

    import msv_b200 as msv
    import synth

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    srs = [48000, 44100, 22050, 48000, 16000, 8000, 48000]          # 7 tracks, mixed rates -> max_sr must be exchanged too
    n_tracks = len(srs)
    tracks = [synth.derive_track(synth.base_clip(4 * sr + 321, sr, seed=100 + sr), t) for t, sr in enumerate(srs)]
    mine = msv.shard_ids(n_tracks, world, rank)
    dev = [torch.from_numpy(tracks[t]).cuda() for t in mine]
    sm = msv.ShardedMultiTrack(device=local)   # attaches this rank's handle to the library's NCCL communicator
    assert sm.mt.device_count() == (1, rank, world)
    sm.add_tracks_device(mine, [d.data_ptr() for d in dev], [d.numel() for d in dev], [srs[t] for t in mine], keepalive=dev)
    outs = []
    for t in mine:
        w = sm.mt.image_width(t, 100.0)
        outs.append(torch.empty(w * 300 * 4, dtype=torch.uint8, device="cuda"))
    sm.render_device(mine, 100.0, 300, 4, [o.data_ptr() for o in outs], [o.numel() for o in outs])
    sm.synchronize()
    got_range = (sm.get_max_db(), sm.get_min_db())
    # single-process answer for the whole batch on this rank's GPU
    ref = msv.MultiTrack(device=local)
    ref.add_tracks_pcm(list(range(n_tracks)), tracks, srs)
    want_range = (ref.get_max_db(), ref.get_min_db())
    ok = got_range == want_range
    for t, o in zip(mine, outs):
        want = ref.get_spec_image_rgba(t, 100.0, 300)
        same = np.array_equal(o.cpu().numpy(), want)
        ok = ok and same
        print(f"rank {rank}: track {t} sr={srs[t]} pixels identical to single-process: {same}", flush=True)
    print(f"rank {rank}: range sharded {got_range} single {want_range}", flush=True)
    # metadata of the other shards arrived with the range: max_sec over ALL tracks, not just this rank's
    ok = ok and abs(sm.mt.get_max_sec() - ref.get_max_sec()) < 1e-6
    # the whole-batch form of the call: every rank passes all ids, the library keeps its share; host PCM, host images
    hm = msv.MultiTrack(device=local)
    hm.attach_nccl(msv.sharded.broadcast_unique_id(device=f"cuda:{local}"), rank, world)
    hm.add_tracks_pcm(list(range(n_tracks)), [tracks[t] if t % world == rank else None for t in range(n_tracks)], srs)
    imgs = hm.get_spec_images(mine, 100.0, 300, 4)
    for t, im in zip(mine, imgs):
        same = np.array_equal(im, ref.get_spec_image_rgba(t, 100.0, 300))
        ok = ok and same
        print(f"rank {rank}: track {t} whole-batch call + batched host images identical: {same}", flush=True)
    # removing the loudest track re-normalises every rank (lib.rs:265-292): collective remove
    loud = int(np.argmax([float(np.abs(x).max()) for x in tracks]))
    hm.remove_track(loud)
    ref2 = msv.MultiTrack(device=local)
    keep = [t for t in range(n_tracks) if t != loud]
    ref2.add_tracks_pcm(keep, [tracks[t] for t in keep], [srs[t] for t in keep])
    same = (hm.get_max_db(), hm.get_min_db()) == (ref2.get_max_db(), ref2.get_min_db())
    for t in mine:
        if t != loud:
            same = same and np.array_equal(hm.get_spec_image(t, 100.0, 300), ref2.get_spec_image(t, 100.0, 300))
    print(f"rank {rank}: after the collective remove of track {loud}: identical to single-process: {same}", flush=True)
    ok = ok and same
    hm.close(); ref2.close()
    # ---- n3: ONE long track time-sharded over all ranks (strips of columns) -----------------------------------
    sr2 = 48000
    long_track = synth.base_clip(90 * sr2 + 123, sr2, seed=4242)
    st2 = msv.ShardedMultiTrack(device=local)
    ob, oc = st2.add_track_time_sharded(0, long_track, sr2, 100.0, rank, world)
    strip = torch.empty(300 * oc * 4, dtype=torch.uint8, device="cuda")
    st2.mt.render_slice_device(0, 100.0, 300, 4, ob, oc, strip.data_ptr(), strip.numel())
    st2.synchronize()
    one = msv.MultiTrack(device=local)
    one.add_tracks_pcm([0], [long_track], [sr2])
    full = one.get_spec_image_rgba(0, 100.0, 300).reshape(300, -1, 4)
    same = np.array_equal(strip.cpu().numpy().reshape(300, oc, 4), full[:, ob:ob + oc]) and \
        (st2.get_max_db(), st2.get_min_db()) == (one.get_max_db(), one.get_min_db())
    print(f"rank {rank}: time-sharded strip [{ob}, {ob + oc}) identical to the single-GPU image: {same}", flush=True)
    ok = ok and same
    st2.close(); one.close()
    # ---- ONE process, several GPUs: a single handle over all devices (sgx_mt_new_sharded) -- rank 0 only ------------
    dist.barrier()
    if rank == 0 and world > 1:
        allg = msv.MultiTrack(devices=list(range(world)))
        assert allg.device_count()[0] == world
        changed = allg.add_tracks_pcm(list(range(n_tracks)), tracks, srs)
        same = changed and (allg.get_max_db(), allg.get_min_db()) == want_range and abs(allg.get_max_sec() - ref.get_max_sec()) < 1e-6
        imgs = allg.get_spec_images(list(range(n_tracks)), 100.0, 300, 4)
        for t, im in enumerate(imgs):
            same = same and np.array_equal(im, ref.get_spec_image_rgba(t, 100.0, 300)) and allg.get_sr(t) == srs[t]
        allg.remove_track(3)
        ref.remove_track(3)
        same = same and (allg.get_max_db(), allg.get_min_db()) == (ref.get_max_db(), ref.get_min_db())
        same = same and np.array_equal(allg.get_spec_image(5, 100.0, 300), ref.get_spec_image(5, 100.0, 300))
        print(f"rank 0: one handle over {world} GPUs identical to the single-GPU handle: {same}", flush=True)
        ok = ok and same
        allg.close()
        assert torch.cuda.current_device() == local, "a call into libsgx.so left another device current"
    dist.barrier()
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    sm.close(); ref.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI-GPU CHECK", "PASSED" if int(flag.item()) == 1 else "FAILED", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()

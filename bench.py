#!/usr/bin/env python
"""bench.py -- headline benchmark of the spectrogram->RGBA path (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c5|c3|c3s|c4 --n-fft F|c2|c1]

A "step" is one pass of the hot path over one batch of synthetic PCM:
    MultiTrack.add_tracks (K1 fused analysis of every track -> dB, K2 global range incl. the
    all-reduce across GPUs, inside libsgx.so) followed by get_spec_image for every track (K3 -> RGBA, 100 px/s x 500).
Default workload (`c5`, BASELINE.json configs[4], weak scaling): every GPU owns 32 synthetic
10-minute 48 kHz mono tracks (256 tracks at 8 GPUs) analysed with the reference's MultiTrack defaults
(W=1920, hop=480, n_fft=2048, default mel bank of 347 bands, 120 dB range).
`value` times the step with PCM and pixels resident in HBM (CUDA events on the engine's stream, max over
ranks); `e2e` times a stream of such batches through the host-buffer C ABI (pinned host PCM in, host RGBA out;
every batch's upload and download inside the timed region, the download of one batch overlapping the upload of
the next; `e2e.one_batch_at_a_time` is the same without that overlap, `e2e.int16_pcm_pipelined` with 16-bit PCM).
The default single-GPU run appends `configs`: the other BASELINE configs (C3 full, three points of the C4 sweep,
C2, C1) measured in the same process, device-resident, compact.
`--impl reference` times the CPU restatement of the reference (oracle/) on the box's host cores and imports
nothing of the GPU package.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "audio-seconds/sec (spectrogram->RGBA)"
UNIT = "audio-s/s"
PX_PER_SEC, NHEIGHT = 100.0, 500  # benches/bench.rs:57
FREQ_LINEAR, FREQ_MEL = 0, 1      # enum FreqScale, lib.rs:25-28 (== SGX_FREQ_*)


def workload(name, n_fft=2048):
    """Returns dict(sr, seconds, channels, tracks_per_gpu, settings kwargs, description)."""
    if name == "c5":
        return dict(sr=48000, seconds=600, channels=1, tracks=32, settings={}, seed=5005,
                    desc="C5 shard: 32 tracks x 10 min x 48 kHz mono per GPU (256 tracks on 8), MultiTrack defaults "
                         "W=1920 hop=480 n_fft=2048, default mel (347 bands), 120 dB range, RGBA 100 px/s x 500")
    if name == "c3":
        return dict(sr=48000, seconds=3600, channels=2, tracks=1, seed=3003,
                    settings=dict(win_length=4096, hop_length=256, n_fft=4096, n_mel=128),
                    desc="C3: 1 h x 48 kHz stereo, n_fft=4096 hop=256 Hann, mel-128, dB + RGBA 100 px/s x 500")
    if name == "c3s":  # n3: ONE track time-sharded over the GPUs (strong scaling)
        w = workload("c3")
        w["desc"] = "C3 time-sharded: ONE 1 h x 48 kHz stereo track, every GPU analyses and renders a strip of columns " \
                    "(n_fft=4096 hop=256, mel-128, RGBA 100 px/s x 500); strong scaling"
        w["sliced"] = True
        return w
    if name == "c2":
        return dict(sr=48000, srs=[8000, 16000, 22050, 24000, 44100, 48000], seconds=44.032, channels=1, tracks=6, seed=2002,
                    settings=dict(n_mel=128),
                    desc="C2: six tracks of 44.032 s at 8/16/22.05/24/44.1/48 kHz (the reference's fixture rates, synthetic PCM), "
                         "per-rate W/hop/n_fft of lib.rs:43-46, 128-band mel, dB + RGBA 100 px/s x 500")
    if name == "c4":  # one point of the long-window sweep; --n-fft picks F
        return dict(sr=44100, seconds=600, channels=1, tracks=1, seed=4004,
                    settings=dict(win_length=n_fft, hop_length=n_fft // 4, n_fft=n_fft, freq_scale=FREQ_LINEAR),
                    desc=f"C4: 10 min x 44.1 kHz mono, n_fft=W={n_fft} hop={n_fft // 4} Hann, linear-frequency dB + RGBA 100 px/s x 500")
    if name == "c1":
        return dict(sr=48000, seconds=44.031854, channels=1, tracks=1, seed=1001,
                    settings=dict(win_length=2048, hop_length=512, n_fft=2048, freq_scale=FREQ_LINEAR),
                    desc="C1: 48 kHz mono 44 s (N=2113529), n_fft=2048 hop=512 Hann, linear-frequency dB + RGBA 100 px/s x 500")
    raise SystemExit(f"unknown workload {name}")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "samples": len(sm)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (a C restatement of the reference; no Rust toolchain exists to build the reference).
# Nothing of the GPU package is imported here: parameters come from the oracle's own lib.rs:43-46 restatement.
# ---------------------------------------------------------------------------------------------------------
def cpu_params(orc, sr, settings):
    win, hop, n_fft = orc.track_params(sr)
    hop = settings.get("hop_length") or hop
    win = settings.get("win_length") or (hop * 4 if settings.get("hop_length") else win)
    n_fft = settings.get("n_fft") or (n_fft if not settings.get("win_length") else 1 << int(np.ceil(np.log2(win))))
    return int(win), int(hop), int(n_fft)


def cpu_sample(wl, cores):
    """A bounded sample of the workload: `cores` tracks (so the reference's per-track rayon parallelism,
    lib.rs:161-166, can use every core) of at most 120 s each."""
    import synth

    sr = wl["sr"]
    secs = min(120, int(wl["seconds"]))
    ntr = max(2, min(cores, 64)) if wl["tracks"] > 1 else 1
    if wl["tracks"] == 1:
        secs = min(int(wl["seconds"]), 60 if wl["channels"] == 2 else 44)
    base = synth.base_clip(secs * sr, sr, wl["seed"])
    wavs = [synth.derive_track(base, t) for t in range(ntr)]
    if wl["channels"] == 2:  # the reference sums channels while loading (lib.rs:42)
        wavs = [w + np.roll(w, 1234) * np.float32(0.75) for w in wavs]
    return wavs, secs, ntr


def run_cpu(wl, steps, warmup):
    import oracle_binding

    orc = oracle_binding.load()
    orc.set_num_threads(len(os.sched_getaffinity(0)))  # all host cores (torchrun pins OMP_NUM_THREADS=1)
    cores = orc.num_threads()
    wavs, secs, ntr = cpu_sample(wl, cores)
    st = wl["settings"]
    sr = wl["sr"]
    win, hop, n_fft = cpu_params(orc, sr, st)
    mel = st.get("freq_scale", FREQ_MEL) == FREQ_MEL
    window = orc.calc_window(win, n_fft)
    fb = None if not mel else (orc.calc_mel_fb(sr, n_fft, st["n_mel"]) if st.get("n_mel") else orc.calc_mel_fb_default(sr, n_fft))
    times, imgs = [], None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        # faithful variant: dense mel GEMM (lib.rs:131), FFT plan per frame in single-track mode (lib.rs:455);
        # rendering runs across tracks on all threads (generous: display.rs is single-threaded per call);
        # the output images are allocated once, by the first (warm-up) call
        imgs, _, _ = orc.pipeline(wavs, [sr] * ntr, [(win, hop, n_fft)] * ntr, [window] * ntr, [fb] * ntr, mel_scale=mel,
                                  px_per_sec=PX_PER_SEC, nheight=NHEIGHT, channels=4, dense_mel=True, parallel_render=True,
                                  out_imgs=imgs)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec_per_step = float(np.median(times)) if times else float("nan")
    return {"value": ntr * secs / sec_per_step, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{ntr} tracks x {secs} s of the same synthetic workload per step, median of {len(times)} steps after {warmup} warm-up; "
                      "C restatement of the reference (oracle/, gcc -O3, OpenMP across tracks like rayon); "
                      "the Rust reference itself cannot be built here (no cargo/rustc)",
            "ms_per_step": sec_per_step * 1e3}


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def make_batch(wl, rank, world, dev, torch, synth):
    """Synthetic batch derived on the device from one uploaded base clip.  Returns
    (global ids, device tensors, samples per track, rates, channels)."""
    sr, ch, ntr = wl["sr"], wl["channels"], wl["tracks"]
    n = int(round(wl["seconds"] * sr)) if wl["seconds"] != 44.031854 else 2113529
    gids = [rank + i * world for i in range(ntr)]  # track t -> GPU t mod G: what the library's sharding keeps on this rank
    tracks = []
    if "srs" in wl:  # mixed sample rates: one clip per rate (every GPU holds the same six tracks)
        srs = list(wl["srs"])
        ns = [int(round(wl["seconds"] * r)) for r in srs]
        for i, (r, m) in enumerate(zip(srs, ns)):
            tracks.append(torch.from_numpy(synth.derive_track(synth.base_clip(m, r, wl["seed"] + r), i)).to(dev))
    else:
        base = torch.from_numpy(synth.base_clip(n, sr, wl["seed"])).to(dev)
        for t in gids:
            gain, shift = synth.track_gain_shift(t, n)
            x = torch.roll(base, -shift) * float(gain)
            if ch == 2:
                x = torch.stack([x, torch.roll(x, 1234) * 0.75], dim=1).contiguous()
            tracks.append(x)
        del base
        ns, srs = [n] * ntr, [sr] * ntr
    return gids, tracks, ns, srs, [ch] * ntr


def measure_device(msv, torch, dist, wl, st, rank, world, local_rank, steps, warmup, clocks=True):
    """`value`: the step with PCM and pixels resident in HBM.  Returns a dict (rank 0) or None."""
    import synth

    dev = torch.device("cuda", local_rank)
    gids, tracks, ns, srs, chs = make_batch(wl, rank, world, dev, torch, synth)
    ntr, ch = len(gids), wl["channels"]
    sm = msv.ShardedMultiTrack(st, device=local_rank)   # attaches the library's NCCL communicator when world > 1
    sm.mt.set_profiling(True)
    sm.mt.set_global_max_sr(max(srs))                    # known up front: keeps the step free of host synchronisation
    ptrs = [x.data_ptr() for x in tracks]
    nwidths = [int(np.float32(PX_PER_SEC) * np.float32(m) / np.float32(r)) for m, r in zip(ns, srs)]
    caps = [w * NHEIGHT * 4 for w in nwidths]
    outs = [torch.empty(c, dtype=torch.uint8, device=dev) for c in caps]
    optrs = [o.data_ptr() for o in outs]

    def step():
        sm.add_tracks_device(gids, ptrs, ns, srs, chs)
        sm.render_device(gids, PX_PER_SEC, NHEIGHT, 4, optrs, caps)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0 and clocks:
        sampler.start()
    l0 = msv.kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(sm.stream)
    for _ in range(steps):
        step()
    ev1.record(sm.stream)
    barrier()
    launches = msv.kernel_launch_count() - l0
    clk = sampler.stop() if (rank == 0 and clocks) else None
    ms_total = ev0.elapsed_time(ev1)
    k1_ms, k3_ms = [], []
    for _ in range(3):  # per-kernel durations (CUDA events on the engine's stream around K1 / K3), steady state
        for _ in range(3):
            step()
        a, r = sm.mt.stage_times()
        k1_ms.append(a); k3_ms.append(r)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / steps
    audio_s_per_gpu = float(sum(m / r for m, r in zip(ns, srs)))
    sm.synchronize()
    rng = (sm.get_max_db(), sm.get_min_db())
    shapes = {i: sm.mt.spec_shape(i) for i in gids}
    sm.close()
    res = {"tracks": tracks, "gids": gids, "ns": ns, "srs": srs, "chs": chs, "caps": caps}
    if rank != 0:
        return res
    peak, peak_src = measured_peak()
    alg_step = float(sum(4 * ch * m + c for m, c in zip(ns, caps)))  # per GPU per step: PCM in + RGBA out
    k1, k3 = float(np.median(k1_ms)), float(np.median(k3_ms))
    db_bytes = sum(4 * shapes[i][0] * shapes[i][1] for i in gids)
    res.update({
        "value": audio_s_per_gpu * world / (ms_step * 1e-3), "ms_per_step": ms_step, "launches": int(launches), "clocks": clk,
        "db_range": rng, "k1_ms": k1, "k3_ms": k3, "k1_ms_samples": k1_ms, "k3_ms_samples": k3_ms, "alg_step": alg_step,
        "k1_own": float(sum(4 * ch * m for m in ns) + db_bytes), "k3_own": float(db_bytes + sum(caps)),
        "audio_s_per_gpu": audio_s_per_gpu, "peak": peak, "peak_src": peak_src})
    return res


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_near_gpu(torch, local_rank):
    """Run this rank (and first-touch its pinned buffers) on the cores of the NUMA node its GPU hangs off, so that host
    copies do not cross the socket interconnect.  Returns what was done, for the JSON line; never fails the run."""
    try:
        p = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        with open(base + "/numa_node") as f:
            node = int(f.read().strip())
        with open(base + "/local_cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if node < 0 or not use or use == allowed:
            return {"pci": bdf, "numa_node": node, "bound": False, "cpus": len(allowed)}
        os.sched_setaffinity(0, use)
        return {"pci": bdf, "numa_node": node, "bound": True, "cpus": len(use)}
    except Exception as ex:  # no sysfs, no NUMA, a container without the attribute
        return {"bound": False, "why": str(ex)[:80]}


def measure_e2e(msv, torch, dist, wl, st, rank, world, local_rank, batch, steps):
    """The same step through the host-buffer C ABI: pinned host PCM -> sgx_mt_add_tracks_pcm -> sgx_mt_get_spec_images
    (batched, renders and downloads pipelined inside the library) -> pinned host RGBA."""
    import synth

    dev = torch.device("cuda", local_rank)
    affinity0 = os.sched_getaffinity(0)
    binding = bind_near_gpu(torch, local_rank)  # before any pinned buffer exists
    tracks, gids, ns, srs, caps = batch["tracks"], batch["gids"], batch["ns"], batch["srs"], batch["caps"]
    ntr, ch, n, sr = len(gids), wl["channels"], batch["ns"][0], wl["sr"]
    host_in = [torch.empty(x.shape, dtype=torch.float32).pin_memory() for x in tracks]
    for h, x in zip(host_in, tracks):
        h.copy_(x)
    np_in = [h.numpy() for h in host_in]
    outs = [[torch.empty(c, dtype=torch.uint8).pin_memory() for c in caps] for _ in range(2)]
    mt = msv.MultiTrack(st, device=local_rank)
    if world > 1:
        mt.attach_nccl(msv.sharded.broadcast_unique_id(device=f"cuda:{local_rank}"), rank, world)
    audio_s = float(sum(m / r for m, r in zip(ns, srs)))
    h2d = int(sum(h.numel() * 4 for h in host_in)); d2h = int(sum(caps))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, reps):
        fn(0)
        barrier()
        t0 = time.perf_counter()
        for k in range(reps):
            fn(k + 1)
        mt.wait_images()
        torch.cuda.synchronize(dev)
        dt = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device=dev)
        per_rank = [float(dt.item())]
        if world > 1:
            parts = [torch.zeros_like(dt) for _ in range(world)]
            dist.all_gather(parts, dt)
            per_rank = [float(p.item()) for p in parts]
        return max(per_rank), per_rank

    reps = max(2, min(steps, 4))

    def sync_step(k):   # one batch at a time: upload, analyse, render, download, done
        mt.add_tracks_pcm(gids, np_in, srs)
        mt.get_spec_images(gids, PX_PER_SEC, NHEIGHT, 4, out=outs[0])

    dt, mine = timed(sync_step, reps)
    gbs = lambda nbytes, secs: [round(nbytes / t / 1e9, 2) for t in secs]  # achieved link rate of every rank over its own step

    def piped_step(k):  # the downloads of batch k run under the upload + analysis of batch k+1 (one handle)
        mt.add_tracks_pcm(gids, np_in, srs)
        mt.get_spec_images(gids, PX_PER_SEC, NHEIGHT, 4, out=outs[k & 1], wait=False)

    dtp, minep = timed(piped_step, reps + 1)
    same = bool(torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][-1], outs[1][-1]))
    # Headline: a stream of batches through ONE handle, as a viewer loading a file list does.  Every batch's upload from
    # pinned host memory and the download of all its images lie inside the timed region (the clock stops after the last
    # image has landed); what overlaps is the download of batch k with the upload + analysis of batch k+1.
    e2e = {"value": audio_s * world / dtp, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": dtp * 1e3,
           "steps": reps + 1, "mode": "consecutive batches, uploads and downloads of neighbouring batches overlap",
           "api": "per batch: sgx_mt_add_tracks_pcm (f32 host PCM) + sgx_mt_get_spec_images_async (batched host RGBA; renders and "
                  "downloads pipelined inside libsgx.so), pinned host buffers (two sets of output buffers), sgx_mt_wait_images at the end",
           "outputs_identical": same, "cpu_binding_rank0": binding,
           "per_rank_link_gbs": {"h2d": gbs(h2d, minep), "d2h": gbs(d2h, minep)},
           "one_batch_at_a_time": {"value": audio_s * world / dt, "ms_per_step": dt * 1e3, "steps": reps,
                                   "api": "sgx_mt_add_tracks_pcm + sgx_mt_get_spec_images (blocking): upload, analyse, render, download, "
                                          "and only then the next batch",
                                   "per_rank_link_gbs": {"h2d_plus_d2h_over_step": gbs(h2d + d2h, mine)}}}
    # the same batch shape with 16-bit host PCM (what WAV files hold and add_tracks(paths) uploads; sgx_mt_add_tracks_pcm_i16)
    del host_in, np_in
    base16 = torch.from_numpy(synth.base_clip_i16(n, sr, wl["seed"]))
    host16 = []
    for t in gids:
        x = torch.roll(base16, -synth.track_gain_shift(t, n)[1])
        if ch == 2:
            x = torch.stack([x, torch.roll(x, 1234)], dim=1).contiguous()
        host16.append(x.pin_memory())
    np16 = [h.numpy() for h in host16]
    h2d16 = int(sum(h.numel() * 2 for h in host16))

    def i16_step(k):
        mt.add_tracks_pcm(gids, np16, srs)
        mt.get_spec_images(gids, PX_PER_SEC, NHEIGHT, 4, out=outs[k & 1], wait=False)

    dt16, mine16 = timed(i16_step, reps + 1)
    e2e["int16_pcm_pipelined"] = {"value": audio_s * world / dt16, "ms_per_step": dt16 * 1e3, "h2d_bytes_per_step": h2d16,
                                  "per_rank_link_gbs": {"h2d": gbs(h2d16, mine16), "d2h": gbs(d2h, mine16)},
                                  "api": "sgx_mt_add_tracks_pcm_i16 (int16 host PCM, scaled on the GPU) + sgx_mt_get_spec_images_async"}
    mt.close()
    os.sched_setaffinity(0, affinity0)  # the CPU baseline that follows uses every core
    return e2e


def run_gpu_sliced(args, wl, rank, world, local_rank, st, msv, torch, dist):
    """One track, time-sharded: rank r owns the columns [nw*r/G, nw*(r+1)/G) (sgx_slice_plan)."""
    import synth

    dev = torch.device("cuda", local_rank)
    sr, ch = wl["sr"], wl["channels"]
    n = int(round(wl["seconds"] * sr))
    base = synth.base_clip(n, sr, wl["seed"])
    nw = msv.calc_nwidth_like(PX_PER_SEC, n, sr)
    ob = nw * rank // world
    oc = nw * (rank + 1) // world - ob
    fb, fc, sb, sc = msv.slice_plan(n, sr, PX_PER_SEC, ob, oc, st)
    x = torch.from_numpy(base[sb:sb + sc]).to(dev)
    if ch == 2:
        full_r = np.roll(base, 1234) * np.float32(0.75)
        x = torch.stack([x, torch.from_numpy(full_r[sb:sb + sc]).to(dev)], dim=1).contiguous()
    del base
    sm = msv.ShardedMultiTrack(st, device=local_rank)
    sm.mt.set_profiling(True)
    sm.mt.set_global_max_sr(sr)
    out = torch.empty(oc * NHEIGHT * 4, dtype=torch.uint8, device=dev)

    def step():
        sm.stream.wait_stream(torch.cuda.current_stream(dev))
        sm.mt.add_track_slice_device(0, x.data_ptr(), sb, sc, n, sr, ch, fb, fc)  # the range exchange happens inside
        if world == 1:
            sm.mt.commit_range_device()
        sm.mt.render_slice_device(0, PX_PER_SEC, NHEIGHT, 4, ob, oc, out.data_ptr(), out.numel())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = msv.kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(sm.stream)
    for _ in range(args.steps):
        step()
    ev1.record(sm.stream)
    barrier()
    launches = msv.kernel_launch_count() - l0
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    k1s, k3s = [], []
    for _ in range(3):
        for _ in range(3):
            step()
        a, r = sm.mt.stage_times()
        k1s.append(a); k3s.append(r)
    sm.synchronize()
    rng = (sm.get_max_db(), sm.get_min_db())
    sm.close()
    if rank != 0:
        return None
    peak, peak_src = measured_peak()
    alg_total = float(4 * ch * n + nw * NHEIGHT * 4)        # whole job, all GPUs
    alg_rank = float(4 * ch * sc + oc * NHEIGHT * 4)        # this rank's launch
    k1, k3 = float(np.median(k1s)), float(np.median(k3s))
    roofline = {"bound": "hbm", "kernel": "stft_db_kernel (K1) on this rank's slice", "achieved": alg_rank / (k1 * 1e-3) / 1e9,
                "peak": peak, "unit": "GB/s", "frac": alg_rank / (k1 * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_rank, "kernel_ms": k1}
    step_roof = {"achieved": alg_total / (ms_step * 1e-3) / 1e9, "frac": alg_total / (ms_step * 1e-3) / 1e9 / (peak * world),
                 "k1_ms": k1, "k3_ms": k3, "step_ms": ms_step, "note": "whole job over all GPUs against N x the measured peak"}
    return {"value": (n / sr) / (ms_step * 1e-3), "ms_per_step": ms_step, "roofline": roofline, "roofline_step": step_roof, "e2e": None,
            "gpu_launches": int(launches), "clocks": clk, "db_range": rng, "scaling": "strong"}


def roofline_of(args, d, workload_name):
    traffic = None
    try:  # DRAM bytes of the dominant kernel from the committed ncu capture, scaled to this launch
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tr = json.load(f).get(workload_name)
        if tr and not (args.tracks or args.seconds):
            traffic = tr["k1_stft_db_bytes"] / tr["audio_seconds_in_capture"] * d["audio_s_per_gpu"]
    except Exception:
        traffic = None
    k1, k3, alg, peak = d["k1_ms"], d["k3_ms"], d["alg_step"], d["peak"]
    roofline = {"bound": "hbm", "kernel": "stft_db_kernel (K1, fused frame/window/rFFT/|X|/mel/dB)",
                "achieved": alg / (k1 * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": alg / (k1 * 1e-3) / 1e9 / peak, "traffic": traffic, "peak_source": d["peak_src"],
                "algorithmic_bytes_per_launch": alg, "kernel_ms": k1,
                "note": "algorithmic bytes = SURVEY 8(d) per-unit figure (f32 PCM in + RGBA out) x audio seconds per launch"}
    step = {"achieved": alg / (d["ms_per_step"] * 1e-3) / 1e9, "frac": alg / (d["ms_per_step"] * 1e-3) / 1e9 / peak,
            "k1_ms": k1, "k3_ms": k3, "step_ms": d["ms_per_step"], "k1_ms_samples": d["k1_ms_samples"], "k3_ms_samples": d["k3_ms_samples"],
            "k1_own_bytes_gbs": d["k1_own"] / (k1 * 1e-3) / 1e9, "k3_own_bytes_gbs": d["k3_own"] / (k3 * 1e-3) / 1e9,
            "note": "whole step (K1+K2+K3) against the same algorithmic bytes; *_own_bytes = each kernel's own minimal HBM traffic incl. the dB intermediate"}
    return roofline, step


def other_configs(msv, torch, dist, args):
    """BASELINE.json configs 0-3 next to the headline, device-resident, one GPU, compact."""
    out = []
    plan = [("c3", 2048, 0), ("c4", 512, 4), ("c4", 2048, 4), ("c4", 16384, 4), ("c2", 2048, 0), ("c1", 2048, 0)]
    for name, n_fft, tracks in plan:
        wl = workload(name, n_fft)
        if tracks:
            wl["tracks"] = tracks  # the C4 config is ONE 10-minute track per FFT size; four are batched so the persistent grid is filled
        st = msv.Settings.default(**wl["settings"])
        try:
            d = measure_device(msv, torch, dist, wl, st, 0, 1, 0, steps=5, warmup=3, clocks=False)
            out.append({"workload": name + (f"_{n_fft}" if name == "c4" else ""), "tracks": wl["tracks"], "value": d["value"],
                        "ms_per_step": d["ms_per_step"], "k1_ms": d["k1_ms"], "k3_ms": d["k3_ms"],
                        "frac_k1": d["alg_step"] / (d["k1_ms"] * 1e-3) / 1e9 / d["peak"],
                        "frac_step": d["alg_step"] / (d["ms_per_step"] * 1e-3) / 1e9 / d["peak"], "gpu_launches_per_step": d["launches"] / 5,
                        "db_range": d["db_range"]})
        except Exception as ex:  # a secondary config must not take the headline down
            out.append({"workload": name, "error": str(ex)[:200]})
        torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c5", "c3", "c3s", "c4", "c2", "c1"])
    ap.add_argument("--n-fft", type=int, default=2048, help="FFT size of the c4 sweep point (512 ... 16384)")
    ap.add_argument("--tracks", type=int, default=0, help="override tracks per GPU (profiling runs only)")
    ap.add_argument("--seconds", type=float, default=0, help="override track length (profiling runs only)")
    ap.add_argument("--channels", type=int, default=0, help="override the channel count of the workload (experiments only)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the compact C1-C4 entries of the default run")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = workload(args.workload, args.n_fft)
    if args.channels:
        wl["channels"] = args.channels
        wl["desc"] += f" [channels overridden: {args.channels}]"
    if args.tracks:
        wl["tracks"] = args.tracks; wl["desc"] += f" [OVERRIDE tracks={args.tracks}: not a headline run]"
    if args.seconds:
        wl["seconds"] = args.seconds; wl["desc"] += f" [OVERRIDE seconds={args.seconds}: not a headline run]"
    cfg = {"workload": wl["desc"], "px_per_sec": PX_PER_SEC, "nheight": NHEIGHT,
           "l2": "inputs larger than L2: every step streams the whole PCM batch and writes every pixel (GBs per step vs 126 MB L2)"
                 if args.workload != "c1" else "C1 fits in L2; not a headline number",
           "sharding": "track t -> GPU t mod G inside libsgx.so; one 16-byte ncclAllReduce(MAX) of {max,-min,max_sr,max_sec} per step on the engine's stream"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps, warmup = max(1, args.steps), max(0, args.warmup)
        cb = run_cpu(wl, steps, warmup)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
                "warmup": warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": cfg, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist

    import msv_b200 as msv

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    st = msv.Settings.default(**wl["settings"])
    if wl.get("sliced"):
        res = run_gpu_sliced(args, wl, rank, world, local_rank, st, msv, torch, dist)
        roofline, step_roof = (res["roofline"], res["roofline_step"]) if res else (None, None)
        e2e, extra = None, None
    else:
        d = measure_device(msv, torch, dist, wl, st, rank, world, local_rank, args.steps, args.warmup)
        e2e = None
        if not args.no_e2e and "srs" not in wl:
            e2e = measure_e2e(msv, torch, dist, wl, st, rank, world, local_rank, d, args.steps)
        res = None
        if rank == 0:
            roofline, step_roof = roofline_of(args, d, args.workload)
            res = {"value": d["value"], "ms_per_step": d["ms_per_step"], "gpu_launches": d["launches"], "clocks": d["clocks"],
                   "db_range": d["db_range"]}
        del d
        torch.cuda.empty_cache()
        extra = None
        if rank == 0 and world == 1 and args.workload == "c5" and not args.no_configs and not (args.tracks or args.seconds):
            extra = other_configs(msv, torch, dist, args)
    cb = None
    if rank == 0 and world == 1 and not args.no_cpu and not wl.get("sliced"):
        cb = run_cpu(wl, 3, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": res.get("scaling", "weak"), "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg, "roofline": roofline, "roofline_step": step_roof,
            "cpu_baseline": cb, "e2e": e2e, "gpu_launches": res["gpu_launches"], "clocks": res["clocks"],
            "db_range": res["db_range"]}
    if extra is not None:
        line["configs"] = extra
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""bench.py -- headline benchmark of the spectrogram->RGBA path (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c5|c3|c3s|c4 --n-fft F|c2|c1]

A "step" is one pass of the hot path over one batch of synthetic PCM:
    MultiTrack.add_tracks (K1 fused analysis of every track -> dB, K2 global range incl. the
    all-reduce across GPUs) followed by get_spec_image for every track (K3 -> RGBA, 100 px/s x 500).
Default workload (`c5`, BASELINE.json configs[4], weak scaling): every GPU owns 32 synthetic
10-minute 48 kHz mono tracks (256 tracks at 8 GPUs) analysed with the reference's MultiTrack defaults
(W=1920, hop=480, n_fft=2048, default mel bank of 347 bands, 120 dB range).
`value` times the step with PCM and pixels resident in HBM (CUDA events on the engine's stream, max over
ranks); `e2e` times the same step through the host-buffer C ABI (pinned host PCM in, host RGBA out).
`--impl reference` times the CPU restatement of the reference (oracle/) on the box's host cores.
Prints ONE JSON line on rank 0.
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
                    settings=dict(win_length=n_fft, hop_length=n_fft // 4, n_fft=n_fft, freq_scale=0),
                    desc=f"C4: 10 min x 44.1 kHz mono, n_fft=W={n_fft} hop={n_fft // 4} Hann, linear-frequency dB + RGBA 100 px/s x 500")
    if name == "c1":
        return dict(sr=48000, seconds=44.031854, channels=1, tracks=1, seed=1001,
                    settings=dict(win_length=2048, hop_length=512, n_fft=2048, freq_scale=0),
                    desc="C1: 48 kHz mono 44 s (N=2113529), n_fft=2048 hop=512 Hann, linear-frequency dB + RGBA 100 px/s x 500")
    raise SystemExit(f"unknown workload {name}")


def alg_bytes_per_audio_second(sr, channels):
    """f32 PCM read once + RGBA written once = 4*ch*sr + 100*500*4 bytes per audio second."""
    return 4 * channels * sr + int(PX_PER_SEC) * NHEIGHT * 4


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
# CPU arm: the oracle (a C restatement of the reference; no Rust toolchain exists to build the reference)
# ---------------------------------------------------------------------------------------------------------
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
    if wl["channels"] == 2:  # the reference sums channels while loading (lib.rs:42); do it inside the timed call
        wavs = [w + np.roll(w, 1234) * np.float32(0.75) for w in wavs]
    return wavs, secs, ntr


def run_cpu(wl, steps, warmup):
    import oracle_binding

    import msv_b200 as msv

    orc = oracle_binding.load()
    orc.set_num_threads(len(os.sched_getaffinity(0)))  # all host cores (torchrun pins OMP_NUM_THREADS=1)
    cores = orc.num_threads()
    wavs, secs, ntr = cpu_sample(wl, cores)
    st = msv.Settings.default(**wl["settings"])
    sr = wl["sr"]
    win, hop, n_fft = msv.track_params(sr, st)
    mel = st.freq_scale == msv.FREQ_MEL
    window = orc.calc_window(win, n_fft)
    fb = None if not mel else (orc.calc_mel_fb(sr, n_fft, st.n_mel) if st.n_mel else orc.calc_mel_fb_default(sr, n_fft))
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        # faithful variant: dense mel GEMM (lib.rs:131), FFT plan per frame in single-track mode (lib.rs:455);
        # rendering runs across tracks on all threads (generous: display.rs is single-threaded per call)
        orc.pipeline(wavs, [sr] * ntr, [(win, hop, n_fft)] * ntr, [window] * ntr, [fb] * ntr, mel_scale=mel,
                     px_per_sec=PX_PER_SEC, nheight=NHEIGHT, channels=4, dense_mel=True, parallel_render=True)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec_per_step = float(np.median(times)) if times else float("nan")
    return {"value": ntr * secs / sec_per_step, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{ntr} tracks x {secs} s of the same synthetic workload per step, median of {len(times)} steps; "
                      "C restatement of the reference (oracle/, gcc -O3, OpenMP across tracks like rayon); "
                      "the Rust reference itself cannot be built here (no cargo/rustc)",
            "ms_per_step": sec_per_step * 1e3}


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def run_gpu(args, wl, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import msv_b200 as msv
    import synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    sr, ch, ntr = wl["sr"], wl["channels"], wl["tracks"]
    n = int(round(wl["seconds"] * sr)) if wl["seconds"] != 44.031854 else 2113529
    st = msv.Settings.default(**wl["settings"])

    # ---- synthetic batch, derived on the device from one uploaded base clip ----
    gids = [rank + i * world for i in range(ntr)]  # track t -> GPU t mod G
    tracks = []
    if wl.get("sliced"):
        return run_gpu_sliced(args, wl, rank, world, local_rank, st, n, sr, ch, dev)
    if "srs" in wl:  # mixed sample rates: one clip per rate (every GPU holds the same six tracks)
        srs = list(wl["srs"])
        ns = [int(round(wl["seconds"] * r)) for r in srs]
        for i, (r, m) in enumerate(zip(srs, ns)):
            tracks.append(torch.from_numpy(synth.derive_track(synth.base_clip(m, r, wl["seed"] + r), i)).to(dev))
    else:
        base_h = synth.base_clip(n, sr, wl["seed"])
        base = torch.from_numpy(base_h).to(dev)
        for t in gids:
            gain, shift = synth.track_gain_shift(t, n)
            x = torch.roll(base, -shift) * float(gain)
            if ch == 2:
                x = torch.stack([x, torch.roll(x, 1234) * 0.75], dim=1).contiguous()
            tracks.append(x)
        del base
        ns = [n] * ntr
        srs = [sr] * ntr
    sm = msv.ShardedMultiTrack(st, device=local_rank)
    sm.mt.set_profiling(True)
    ids = list(range(ntr))
    ptrs = [x.data_ptr() for x in tracks]
    chs = [ch] * ntr
    nwidths = [int(np.float32(PX_PER_SEC) * np.float32(m) / np.float32(r)) for m, r in zip(ns, srs)]
    caps = [w * NHEIGHT * 4 for w in nwidths]
    img_bytes = caps[0]
    outs = [torch.empty(c, dtype=torch.uint8, device=dev) for c in caps]
    optrs = [o.data_ptr() for o in outs]

    def step():
        sm.add_tracks_device(ids, ptrs, ns, srs, chs, exchange_max_sr=False)
        sm.render_device(ids, PX_PER_SEC, NHEIGHT, 4, optrs, caps)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sm.mt.set_global_max_sr(max(srs))
    for _ in range(args.warmup):
        step()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = msv.kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k1_ms, k3_ms = [], []
    barrier()
    ev0.record(sm.stream)
    for _ in range(args.steps):
        step()
    ev1.record(sm.stream)
    barrier()
    launches = msv.kernel_launch_count() - l0
    clk = clocks.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    # per-kernel durations (CUDA events on the engine's stream around K1 / K3), a few extra steps
    for _ in range(3):
        for _ in range(3):  # back-to-back steps: the events of the last one are read in steady state
            step()
        a, r = sm.mt.stage_times()
        k1_ms.append(a); k3_ms.append(r)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    audio_s_per_gpu = float(sum(m / r for m, r in zip(ns, srs)))
    value = audio_s_per_gpu * world / (ms_step * 1e-3)

    # ---- e2e: the host-buffer C ABI (pinned host PCM -> add_tracks_pcm -> get_spec_image_rgba -> host) ----
    e2e = None
    if not args.no_e2e and "srs" not in wl:
        e2e_tracks = ntr
        host_in = [torch.empty(x.shape, dtype=torch.float32).pin_memory() for x in tracks[:e2e_tracks]]
        for h, x in zip(host_in, tracks):
            h.copy_(x)
        host_out = [torch.empty(img_bytes, dtype=torch.uint8).pin_memory() for _ in range(e2e_tracks)]
        np_in = [h.numpy() for h in host_in]
        mt2 = msv.MultiTrack(st, device=local_rank)
        import ctypes as C

        def e2e_step():
            mt2.add_tracks_pcm(list(range(e2e_tracks)), np_in, srs[:e2e_tracks])
            for i in range(e2e_tracks):
                need = C.c_size_t()
                msv._check(msv._lib.sgx_mt_get_spec_image_rgba(mt2._h, i, PX_PER_SEC, NHEIGHT, host_out[i].data_ptr(), img_bytes, C.byref(need)))

        e2e_step()
        barrier()
        reps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(reps):
            e2e_step()
        torch.cuda.synchronize(dev)
        dt = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": e2e_tracks * n / sr * world / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(sum(h.numel() * 4 for h in host_in)),
               "d2h_bytes_per_step": int(img_bytes * e2e_tracks), "ms_per_step": float(dt.item()) * 1e3,
               "api": "sgx_mt_add_tracks_pcm (f32 host PCM) + sgx_mt_get_spec_image_rgba per track, pinned host buffers"}
        # Two batches in flight: a second handle (own stream, own output buffers) runs the same steps on a second
        # host thread, half a step out of phase.  Inside ONE step the global dB range forces upload -> analysis ->
        # render -> download in sequence, so a single handle uses one PCIe direction at a time; two handles let the
        # upload of one batch run under the download of the other.  Reported next to the headline, not instead of it.
        if not args.no_e2e2 and world == 1:  # one rank only: it pins a second set of output buffers on the host
            import threading
            mt3 = msv.MultiTrack(st, device=local_rank)
            host_out2 = [torch.empty(img_bytes, dtype=torch.uint8).pin_memory() for _ in range(e2e_tracks)]

            def steps_on(mt, outs, count, delay):
                time.sleep(delay)
                for _ in range(count):
                    mt.add_tracks_pcm(list(range(e2e_tracks)), np_in, srs[:e2e_tracks])
                    for i in range(e2e_tracks):
                        need = C.c_size_t()
                        msv._check(msv._lib.sgx_mt_get_spec_image_rgba(mt._h, i, PX_PER_SEC, NHEIGHT, outs[i].data_ptr(), img_bytes, C.byref(need)))

            steps_on(mt3, host_out2, 1, 0.0)  # warm-up of the second handle
            torch.cuda.synchronize(dev)
            barrier()
            reps2 = max(2, reps)
            step_s = float(dt.item())
            th = [threading.Thread(target=steps_on, args=(mt2, host_out, reps2, 0.0)),
                  threading.Thread(target=steps_on, args=(mt3, host_out2, reps2, 0.5 * step_s))]
            t0 = time.perf_counter()
            for t_ in th: t_.start()
            for t_ in th: t_.join()
            torch.cuda.synchronize(dev)
            dt2 = torch.tensor([(time.perf_counter() - t0) / (2 * reps2)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt2, op=dist.ReduceOp.MAX)
            same = bool(torch.equal(host_out[0], host_out2[0]) and torch.equal(host_out[-1], host_out2[-1]))
            e2e["two_batches_in_flight"] = {"value": e2e_tracks * n / sr * world / float(dt2.item()), "ms_per_step": float(dt2.item()) * 1e3,
                                            "steps": 2 * reps2, "outputs_identical": same,
                                            "note": "two MultiTrack handles on two host threads, half a step out of phase: upload of one batch under the download of the other"}
            mt3.close()
            del host_out2
        # the same batch shape with 16-bit host PCM (what the WAV fixtures hold; sgx_mt_add_tracks_pcm_i16)
        del host_in, np_in
        base16 = torch.from_numpy(synth.base_clip_i16(n, sr, wl["seed"]))
        host16 = []
        for t in gids[:e2e_tracks]:
            x = torch.roll(base16, -synth.track_gain_shift(t, n)[1])
            if ch == 2:
                x = torch.stack([x, torch.roll(x, 1234)], dim=1).contiguous()
            host16.append(x.pin_memory())
        np16 = [h.numpy() for h in host16]

        def e2e16_step():
            mt2.add_tracks_pcm(list(range(e2e_tracks)), np16, srs[:e2e_tracks])
            for i in range(e2e_tracks):
                need = C.c_size_t()
                msv._check(msv._lib.sgx_mt_get_spec_image_rgba(mt2._h, i, PX_PER_SEC, NHEIGHT, host_out[i].data_ptr(), img_bytes, C.byref(need)))

        e2e16_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            e2e16_step()
        torch.cuda.synchronize(dev)
        dt = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e["int16_pcm"] = {"value": e2e_tracks * n / sr * world / float(dt.item()), "ms_per_step": float(dt.item()) * 1e3,
                            "h2d_bytes_per_step": int(sum(h.numel() * 2 for h in host16)),
                            "api": "sgx_mt_add_tracks_pcm_i16 (int16 host PCM, scaled on the GPU) + sgx_mt_get_spec_image_rgba"}
        mt2.close()

    # parity spot check of what was timed (smoke-level; the real gate is tests/ -m gpu)
    sm.synchronize()
    rng = (sm.get_max_db(), sm.get_min_db())
    shapes = {i: sm.mt.spec_shape(i) for i in ids}
    sm.close()
    if rank != 0:
        return None
    peak, peak_src = measured_peak()
    alg_step = float(sum(4 * ch * m + c for m, c in zip(ns, caps)))  # per GPU per step: PCM in + RGBA out
    k1 = float(np.median(k1_ms)); k3 = float(np.median(k3_ms))
    db_bytes = 0
    for i in ids:
        T_i, n_out_i = shapes[i]
        db_bytes += 4 * T_i * n_out_i
    k1_own = float(sum(4 * ch * m for m in ns) + db_bytes)   # PCM read + dB written
    k3_own = float(db_bytes + sum(caps))                      # dB read + pixels written
    traffic = None
    try:  # DRAM bytes of the dominant kernel from the committed ncu capture, scaled to this launch
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tr = json.load(f).get(args.workload)
        if tr and not (args.tracks or args.seconds):
            traffic = tr["k1_stft_db_bytes"] / tr["audio_seconds_in_capture"] * audio_s_per_gpu
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "kernel": "stft_db_kernel (K1, fused frame/window/rFFT/|X|/mel/dB)",
                "achieved": alg_step / (k1 * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": alg_step / (k1 * 1e-3) / 1e9 / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_step, "kernel_ms": k1,
                "note": "algorithmic bytes = SURVEY 8(d) per-unit figure (f32 PCM in + RGBA out) x audio seconds per launch"}
    step_roof = {"achieved": alg_step / (ms_step * 1e-3) / 1e9, "frac": alg_step / (ms_step * 1e-3) / 1e9 / peak,
                 "k1_ms": k1, "k3_ms": k3, "step_ms": ms_step, "k1_ms_samples": k1_ms, "k3_ms_samples": k3_ms,
                 "k1_own_bytes_gbs": k1_own / (k1 * 1e-3) / 1e9, "k3_own_bytes_gbs": k3_own / (k3 * 1e-3) / 1e9,
                 "note": "whole step (K1+K2+K3) against the same algorithmic bytes; *_own_bytes = each kernel's own minimal HBM traffic incl. the dB intermediate"}
    return {"value": value, "ms_per_step": ms_step, "roofline": roofline, "roofline_step": step_roof, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clk, "db_range": rng}


def run_gpu_sliced(args, wl, rank, world, local_rank, st, n, sr, ch, dev):
    """One track, time-sharded: rank r owns the columns [nw*r/G, nw*(r+1)/G) (sgx_slice_plan)."""
    import torch
    import torch.distributed as dist

    import msv_b200 as msv
    import synth

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
        sm.mt.add_track_slice_device(0, x.data_ptr(), sb, sc, n, sr, ch, fb, fc)
        with torch.cuda.stream(sm.stream):
            msv.sharded.all_reduce_range(sm._range, sm.group)
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
    ap.add_argument("--no-e2e2", action="store_true", help="skip the two-batches-in-flight variant of the e2e measurement")
    ap.add_argument("--no-cpu", action="store_true")
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
           "sharding": "track t -> GPU t mod G; one 8-byte all-reduce(MAX) of {max,-min} per step"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = max(1, min(args.steps, 5))
        cb = run_cpu(wl, steps, max(1, min(args.warmup, 1)))
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
                "warmup": 1, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": cfg, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    res = run_gpu(args, wl, rank, world, local_rank)
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
            "data": "synthetic", "config": cfg, "roofline": res["roofline"], "roofline_step": res["roofline_step"],
            "cpu_baseline": cb, "e2e": res["e2e"], "gpu_launches": res["gpu_launches"], "clocks": res["clocks"],
            "db_range": res["db_range"]}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())

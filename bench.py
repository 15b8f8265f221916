#!/usr/bin/env python
"""Benchmark of the batched phase-vocoder hot path (BASELINE.json metric: audio-seconds per wall-second,
2048-point pitch shift, batched streams).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--streams S] [--secs T]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference      # the reference's own CPU path on the host cores

Workload (config 4 of BASELINE.json): S = 4096 synthetic mono 44.1 kHz streams of 10 s PER GPU, +7 semitones,
coremode 1 (phase locked), FFT 2048.  A step is one pass of the whole path over that batch.  `value` is measured with
the inputs resident in HBM (7.2 GB of float32 PCM per GPU, far larger than L2); `e2e` goes through the host-buffer
entry point (pinned host memory -> H2D -> kernels -> D2H) every step.  Streams are independent, so multi-GPU is one
process per GPU with its own batch and no collective on the data path (weak scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR, SEMITONES, FFT, COREMODE, MODE = 44100, 7.0, 2048, 1, 0
CHANNELS, TIMERATIO = 1, 1.0
# The judged line is cfg4 (the default).  The other BASELINE.json configurations can be timed with --workload for the
# notes in profiles/; they are parity-test cases, not bench lines.
WORKLOADS = {   # name: (sr, channels, timeratio, semitones, mode, coremode, fftsize, default streams)
    "cfg4": (44100, 1, 1.0, 7.0, 0, 1, 2048, 4096),
    "cfg1": (44100, 2, 1.0, 4.0, 0, 1, 2048, 2048),
    "cfg2": (48000, 2, 1.5, 0.0, 5, 1, 4096, 1024),
    "cfg3-formant": (44100, 1, 1.0, 4.0, 2, 1, 2048, 4096),
    "cfg3-gender": (44100, 1, 1.0, -4.0, 1, 1, 2048, 4096),
    "cfg5-robotic-2048": (44100, 2, 1.0, 0.0, 6, 1, 2048, 1024),
    "cfg5-whisper-2048": (44100, 2, 1.0, 0.0, 7, 1, 2048, 1024),
    "cfg5-vocoder-2048": (44100, 2, 1.0, 0.0, 3, 1, 2048, 1024),
    "cfg5-robotic-512": (44100, 2, 1.0, 0.0, 6, 1, 512, 1024),
    "cfg5-robotic-8192": (44100, 2, 1.0, 0.0, 6, 1, 8192, 1024),
}
METRIC = "audio-sec/sec, 2048-pt PV pitch-shift, batched streams"
UNIT = "audio-s/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--streams", type=int, default=0, help="streams per GPU (default: the workload's)")
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--secs", type=float, default=10.0)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames-per-chunk", type=int, default=0)
    ap.add_argument("--rows-per-group", type=int, default=0)
    ap.add_argument("--contexts", type=int, default=0)
    ap.add_argument("--cpu-streams-per-core", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    global SR, CHANNELS, TIMERATIO, SEMITONES, MODE, COREMODE, FFT
    SR, CHANNELS, TIMERATIO, SEMITONES, MODE, COREMODE, FFT, default_streams = WORKLOADS[args.workload]
    if args.streams <= 0:
        args.streams = default_streams
    return args


def workload_config(args, n_gpus):
    return {"workload": f"{args.workload}: {args.streams} synthetic {'mono' if CHANNELS == 1 else 'stereo'} {SR / 1000:g} kHz {args.secs:g} s streams per GPU, "
                        f"time ratio {TIMERATIO:g}, {SEMITONES:+g} semitones, mode {MODE}, coremode {COREMODE}, FFT {FFT}"
                        + (" (BASELINE.json configs[3])" if args.workload == "cfg4" else ""),
            "streams_per_gpu": args.streams, "stream_seconds": args.secs, "sample_rate": SR, "semitones": SEMITONES, "fftsize": FFT,
            "coremode": COREMODE, "channels": CHANNELS, "time_ratio": TIMERATIO, "mode": MODE, "n_gpus": n_gpus, "sharding": "streams split across GPUs, no collective",
            "l2": "inputs (7.2 GB/GPU at the default size) exceed the 126 MB L2; no explicit flush"}


# ---------------------------------------------------------------------------------------------------------------
# CPU side: the unmodified reference (oracle/_ref/pvref_drv, one OS process per stream because the reference keeps
# process-global state) or, if it was not built, the C restatement.  Only used for cpu_baseline / --impl reference.
# ---------------------------------------------------------------------------------------------------------------
def cpu_run(n_streams: int, secs: float, cores: int):
    """Process n_streams synthetic streams on `cores` host cores; returns (audio-s/s, kind, seconds)."""
    from audiomod_b200.synth import synth
    from oracle import pv_oracle as O
    xs = [synth(4000 + i, SR, secs, CHANNELS) for i in range(n_streams)]
    if O.have_ref():
        kind = "reference"
        with tempfile.TemporaryDirectory(prefix="pvbench_") as d:
            for i, x in enumerate(xs):
                x.tofile(os.path.join(d, f"i{i}.f32"))
            cmds = [[O.REF_DRV, str(SR), str(CHANNELS), repr(TIMERATIO), repr(SEMITONES), str(MODE), str(COREMODE), str(FFT),
                     os.path.join(d, f"i{i}.f32"), os.path.join(d, f"o{i}.f32")] for i in range(n_streams)]
            t0 = time.perf_counter()
            running, nxt = [], 0
            while nxt < len(cmds) or running:
                while nxt < len(cmds) and len(running) < cores:
                    running.append(subprocess.Popen(cmds[nxt]))
                    nxt += 1
                p = running.pop(0)
                if p.wait() != 0:
                    raise RuntimeError("reference driver failed")
            dt = time.perf_counter() - t0
    else:
        kind = "port"
        import multiprocessing as mp
        O.lib()
        t0 = time.perf_counter()
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_port_one, xs)
        dt = time.perf_counter() - t0
    return n_streams * secs / dt, kind, dt


def _port_one(x):
    from oracle import pv_oracle as O
    return O.run_offline(x, SR, timeratio=TIMERATIO, semitones=SEMITONES, mode=MODE, coremode=COREMODE, fftsize=FFT).shape[1]


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = max(cores, cores * args.cpu_streams_per_core)
    for _ in range(min(args.warmup, 1)):
        cpu_run(cores, min(args.secs, 2.0), cores)
    vals, times, kind = [], [], "port"
    for _ in range(args.steps):
        v, kind, dt = cpu_run(n, args.secs, cores)
        vals.append(v)
        times.append(dt)
    value = float(np.mean(vals))
    sample = f"{n} of the workload's streams ({args.secs:g} s each) per step, one OS process per stream on {cores} cores"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML; nvidia-smi as a fallback)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.max_sm, self.reasons, self.stop_flag = index, [], [], set(), threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        self.max_sm.append(float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)))
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        c = [v.strip() for v in out.split(",")]
        self.sm.append(float(c[0]))
        self.max_sm.append(float(c[1]))
        for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[2:]):
            if v.lower().startswith("active"):
                self.reasons.add(name)

    def run(self):
        while not self.stop_flag.is_set():
            try:
                self._sample_nvml() if self.nvml else self._sample_smi()
            except Exception:
                pass
            self.stop_flag.wait(0.02 if self.nvml else 0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.max_sm) if self.max_sm else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def make_inputs(torch, dev, streams, n, seed):
    """Seeded synthetic PCM generated on the GPU (same recipe family as audiomod_b200/synth.py): 8 harmonics with
    vibrato and tremolo plus noise, quantised to int16 and presented as int16/32768 like the reference WAV reader."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    out = torch.empty((streams, n), dtype=torch.float32, device=dev)
    t = torch.arange(n, dtype=torch.float64, device=dev) / SR
    step = 64
    for s0 in range(0, streams, step):
        m = min(step, streams - s0)
        u = torch.rand((m, 1), generator=g, device=dev, dtype=torch.float64) * 12 - 6
        f0 = 110.0 * 2.0 ** (u / 12.0)
        vph = torch.rand((m, 1), generator=g, device=dev, dtype=torch.float64) * 2 * np.pi
        tph = torch.rand((m, 1), generator=g, device=dev, dtype=torch.float64) * 2 * np.pi
        cyc = f0 * (t[None, :] - 0.01 / (2 * np.pi * 0.7) * torch.cos(2 * np.pi * 0.7 * t[None, :] + vph))
        cyc = (cyc - torch.floor(cyc)).to(torch.float32) * (2 * np.pi)
        x = torch.zeros((m, n), dtype=torch.float32, device=dev)
        for k in range(8):
            ph0 = torch.rand((m, 1), generator=g, device=dev, dtype=torch.float32) * (2 * np.pi)
            x += (0.25 / (k + 1)) * torch.sin((k + 1) * cyc + ph0)
        trem = (0.8 + 0.2 * torch.sin(2 * np.pi * 1.3 * t[None, :] + tph)).to(torch.float32)
        x = x * trem + 0.01 * torch.randn((m, n), generator=g, device=dev, dtype=torch.float32)
        q = torch.round(torch.clamp(x, -1.0, 1.0) * 32767.0)
        out[s0:s0 + m] = (q.to(torch.float64) * (1.0 / 32768.0)).to(torch.float32)
        del cyc, x, trem, q
    return out


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
        return
    import torch
    import torch.distributed as dist
    import audiomod_b200 as A

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the phase vocoder has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = int(round(SR * args.secs))
    n_streams = args.streams
    S = n_streams * CHANNELS          # channel rows
    stride = (n + 3) & ~3
    d_in = torch.zeros((S, stride), dtype=torch.float32, device=dev)
    d_in[:, :n] = make_inputs(torch, dev, S, n, 1234 + rank)
    batch = A.PhaseVocoderBatch(n_streams, n, SR, CHANNELS, TIMERATIO, SEMITONES, MODE, COREMODE, FFT, device=local)
    if args.frames_per_chunk or args.rows_per_group or args.contexts:
        batch.tune(args.frames_per_chunk, args.rows_per_group, args.contexts)
    n_out = batch.plan(n)
    out_stride = (int(n_out.max()) + 3) & ~3
    d_out = torch.zeros((S, out_stride), dtype=torch.float32, device=dev)
    info = batch.info()
    stream = torch.cuda.current_stream()

    def step_device():
        batch.run_device(d_in.data_ptr(), stride, d_out.data_ptr(), out_stride, stream.cuda_stream)

    # ---- device-resident throughput ----
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    stats = batch.stats()
    clocks = sampler.summary()
    # per-kernel CUDA-event times: the same steps once more with one group in flight at a time, so that a kernel's
    # events bracket only that kernel (with several groups in flight kernels of different groups share the SMs)
    batch.tune(contexts=1)
    step_device()
    torch.cuda.synchronize()
    batch.profile(True)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record(stream)
    for _ in range(args.steps):
        step_device()
    pe1.record(stream)
    torch.cuda.synchronize()
    ms_serial_step = pe0.elapsed_time(pe1) / args.steps
    ktimes = batch.kernel_times()
    batch.profile(False)
    batch.tune(contexts=args.contexts or 3)
    ms_step = ms_total / args.steps
    audio_sec_per_step = world * n_streams * args.secs
    value = audio_sec_per_step / (ms_step / 1e3)
    checksum = float(d_out[:, :int(n_out.min())].double().abs().mean().item())

    # ---- roofline of the dominant kernel (staged-traffic model, DESIGN.md "Algorithmic bytes") ----
    N, hop, H, half = info["fftsize"], info["hop"], info["bins"], info["fftsize"] // 2
    slices = stats["slices"]
    shift = hop * info["hs_ratio"]
    per_frame = {"analyse": 4 * (hop + 2 * H), "phase_core": 4 * (2 * H + half), "synthesise": 4 * (2 * H + N),
                 "ola_resample": 4 * (N + shift / info["pitch_scale"])}
    per_frame["fixed_phase"] = 4 * H
    # phase-locked core on Cartesian spectra (pv_lock.cuh): P peaks per frame is data dependent; a noise-like spectrum has a
    # strict +-2 local maximum at one bin in five, which is what the synthetic workload measures (profiles/traffic.json)
    P = half / 5.0
    per_frame["lock_peaks"] = 4 * 2 * H * (1 + 1.0 / 32) + 16 * P + 2 * half + 8   # spectrum (+ warm-up frame per run) -> records, map, header
    per_frame["lock_chain"] = 16 * P + 8 + 8 * P                                  # records, header -> (cos, sin) per region
    if ktimes.get("lock_chain", (0, 0))[1] > 0:
        per_frame["synthesise"] = 4 * (2 * H + N) + 2 * half + 8 * P + 8          # + bin->region map, rotations, header
    dom = max((k for k in per_frame if ktimes.get(k, (0, 0))[1] > 0), key=lambda k: ktimes[k][0])
    dom_ms, dom_launches = ktimes[dom]
    bytes_total = per_frame[dom] * slices * S * args.steps
    achieved = bytes_total / (dom_ms / 1e3) / 1e9
    peak, peak_src = 6650.0, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        pass
    traffic = None
    try:   # measured DRAM bytes per frame of that kernel from the committed ncu --set full capture, scaled to one launch
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f)[dom]["dram_bytes_per_frame"] * slices * S * args.steps / max(dom_launches, 1)
    except Exception:
        pass
    kernel_ms_sum = sum(v[0] for v in ktimes.values())
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "bytes_per_launch": bytes_total / max(dom_launches, 1),
                "avg_launch_ms": dom_ms / max(dom_launches, 1),
                "kernel_share_of_step": dom_ms / max(kernel_ms_sum, 1e-9),
                "kernel_ms_per_step": {k: v[0] / args.steps for k, v in ktimes.items() if v[1]},
                "serialised_ms_per_step": ms_serial_step,
                "compulsory_io_frac": (4.0 * (n + float(n_out.max())) * S) / (ms_step / 1e3) / 1e9 / peak}

    # ---- end to end through the host-buffer entry point ----
    e2e = None
    if not args.no_e2e:
        h_in = torch.empty((S, stride), dtype=torch.float32, pin_memory=True)
        h_in.copy_(d_in)
        h_out = torch.empty((S, out_stride), dtype=torch.float32, pin_memory=True)
        in_rows = [h_in.data_ptr() + 4 * stride * r for r in range(S)]
        out_rows = [h_out.data_ptr() + 4 * out_stride * r for r in range(S)]
        batch.run_host_rows(in_rows, out_rows)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            batch.run_host_rows(in_rows, out_rows)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0) / args.steps
        st = batch.stats()
        e2e = {"value": audio_sec_per_step / dt, "unit": UNIT, "h2d_bytes_per_step": st["h2d_bytes"] * world,
               "d2h_bytes_per_step": st["d2h_bytes"] * world, "ms_per_step": dt * 1e3, "format": "f32 in, f32 out, pinned host memory",
               "result_checksum": float(h_out[:, :int(n_out.min())].double().abs().mean().item())}

        # the same through int16 PCM rows (the reference CLI's own I/O format: 16-bit WAV in, 16-bit WAV out), reported beside it
        del h_in, h_out
        q_in = torch.empty((S, stride), dtype=torch.int16, pin_memory=True)
        q_in.copy_(torch.round(d_in * 32768.0).clamp_(-32768, 32767).to(torch.int16))
        q_out = torch.empty((S, out_stride), dtype=torch.int16, pin_memory=True)
        qi = [q_in.data_ptr() + 2 * stride * r for r in range(S)]
        qo = [q_out.data_ptr() + 2 * out_stride * r for r in range(S)]
        batch.run_host_rows(qi, qo, A.S16)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            batch.run_host_rows(qi, qo, A.S16)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0) / args.steps
        st = batch.stats()
        e2e["s16"] = {"value": audio_sec_per_step / dt, "unit": UNIT, "h2d_bytes_per_step": st["h2d_bytes"] * world,
                      "d2h_bytes_per_step": st["d2h_bytes"] * world, "ms_per_step": dt * 1e3,
                      "format": "int16 PCM in, int16 PCM out, pinned host memory"}
        del q_in, q_out

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        ncpu = max(cores, cores * args.cpu_streams_per_core)
        v, kind, dt = cpu_run(ncpu, args.secs, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{ncpu} streams of the same workload ({args.secs:g} s each), one OS process per stream, {dt:.1f} s wall"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": workload_config(args, world), "clocks": clocks, "e2e": e2e,
                "gpu_launches": stats["kernel_launches"] * args.steps, "roofline": roofline, "cpu_baseline": cpu,
                "result_checksum": checksum}
        print(json.dumps(line), flush=True)
    batch.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

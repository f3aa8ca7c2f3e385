#!/usr/bin/env python
"""bench.py -- IQ Msamples/s through the spectrum + FM chain (BASELINE.json metric).

Workload (BASELINE.json configs[4]): 256 independent
synthetic dongle streams, every stream pushed through the FULL chain each step:
  * a 1024-point power spectrum of EVERY frame -> float32 dB, display order (K = 1), and
  * CIC /10 -> FM discriminator -> two half-band decimators -> float32 audio (fs/40),
from ONE pass over the device-resident u8 IQ (b200_chain_exec), plus the 6-frame averaged
u8 dB spectrum per stream (what the reference ships to its UI client) written straight into
the buffer that NCCL gathers to rank 0.  Streams are sharded across ranks (stream s -> rank
s mod G); there is no collective on the data path, only that gather of 1 KiB per stream.
Total work is fixed at 256 streams => "scaling": "strong".

  python bench.py --gpus N --steps K --warmup W          (N > 1: launched under torchrun)
  python bench.py --impl reference ...                    the reference's own CPU path

One JSON line on stdout from rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import __graft_entry__ as graft  # noqa: E402

N_STREAMS = 256
FS = 2_048_000
TILE = 5120
BYTES_PER_SAMPLE = 2.0 + 4.0 + 4.0 / 40.0        # u8 IQ in, f32 dB out (K = 1), f32 audio out / 40
METRIC = "IQ Msamples/s (spectrum+FM chain)"
UNIT = "Msamples/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=N_STREAMS)
    ap.add_argument("--samples-per-stream", type=int, default=TILE * 1600)     # 8 192 000 = 4 s of IQ per stream
    ap.add_argument("--e2e-samples-per-stream", type=int, default=TILE * 100)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_config(args, world):
    return {
        "workload": "configs[4]: 256 independent synthetic dongle streams, full spectrum+FM chain "
                    "(1024-pt FFT of every frame -> f32 dB, K=1; CIC/10 -> FM discriminator -> 2x half-band -> f32 audio)",
        "streams": args.streams,
        "samples_per_stream_per_step": args.samples_per_stream,
        "fft_points": 1024, "frames_averaged": 1, "down_factor": 10, "sample_rate_hz": FS,
        "sharding": f"stream s -> rank s mod {world}; NCCL gather of 6-frame averaged u8 dB spectra to rank 0",
        "l2": "inputs (>= 1 GiB per rank per step) and outputs exceed the 126 MB L2; no flush needed",
        "algorithmic_bytes_per_sample": BYTES_PER_SAMPLE,
    }


# --------------------------------------------------------------------------------------
# synthetic IQ
# --------------------------------------------------------------------------------------

def synth_fm_device(torch, stream_ids, n, device):
    """S3 generated on the device: per-stream seed 1000 + stream id, FM with two
    message tones, 25 kHz peak deviation, amplitude 100, noise sigma 2, quantised to offset-binary u8."""
    out = torch.empty((len(stream_ids), n, 2), dtype=torch.uint8, device=device)
    t = torch.arange(n, dtype=torch.float64, device=device) / FS
    for i, sid in enumerate(stream_ids):
        g = torch.Generator(device=device).manual_seed(1000 + int(sid))
        f1 = 1000.0 + 37.0 * (sid % 16)
        msg = 0.5 * (torch.sin(2 * np.pi * f1 * t) + torch.sin(2 * np.pi * 5000.0 * t))
        phase = (2 * np.pi * 25_000.0 / FS) * torch.cumsum(msg, dim=0)
        noise = 2.0 * torch.randn((n, 2), dtype=torch.float32, device=device, generator=g)
        re = 127.5 + 100.0 * torch.cos(phase).float() + noise[:, 0]
        im = 127.5 + 100.0 * torch.sin(phase).float() + noise[:, 1]
        out[i, :, 0] = re.round().clamp(0, 255).to(torch.uint8)
        out[i, :, 1] = im.round().clamp(0, 255).to(torch.uint8)
        del msg, phase, noise, re, im
    return out


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------

class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []          # (host time, csv line)
        self.window = None

    def start(self):
        """Start sampling (nvidia-smi needs ~100 ms to come up, so start before the warm-up)."""
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark(self, t0: float, t1: float):
        """Host-clock bounds of the timed region; only samples inside it are reported."""
        self.window = (t0, t1)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = self.lines
        if self.window is not None:
            inside = [(t, l) for t, l in lines if self.window[0] <= t <= self.window[1]]
            if not inside and lines:      # region shorter than the sampling period: the closest sample
                mid = 0.5 * (self.window[0] + self.window[1])
                inside = [min(lines, key=lambda tl: abs(tl[0] - mid))]
            lines = inside
        for _, line in lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------
# CPU reference arm (oracle/_ref): the reference's own code on the host cores
# --------------------------------------------------------------------------------------

_REF_DATA = {}       # samples_per_stream -> (path, n_unique): the synthetic captures, written once per process


def _ref_data_file(samples_per_stream: int):
    import atexit
    if samples_per_stream in _REF_DATA:
        return _REF_DATA[samples_per_stream]
    pkg = graft.load_package()
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    path = os.path.join(shm, f"b200sdr_refbench_{os.getpid()}_{samples_per_stream}.bin")
    n_unique = 16          # 16 distinct captures, reused round robin (timing is data-independent)
    with open(path, "wb") as f:
        for s in range(n_unique):
            pkg.synth.s3_fm(samples_per_stream, seed=1000 + s).tofile(f)
    atexit.register(lambda: os.path.exists(path) and os.unlink(path))
    _REF_DATA[samples_per_stream] = (path, n_unique)
    return _REF_DATA[samples_per_stream]


def run_ref_bench(n_streams: int, samples_per_stream: int, workers: int, mode: str = "chain"):
    from oracle import pyoracle as po
    if not os.path.exists(po.REF_BENCH):
        po.build("all")
    if not os.path.exists(po.REF_BENCH):
        return None
    path, n_unique = _ref_data_file(samples_per_stream)
    res = subprocess.run([po.REF_BENCH, path, str(n_streams), str(samples_per_stream), str(workers), mode,
                          str(min(n_unique, n_streams))], capture_output=True, text=True, timeout=900)
    if res.returncode != 0:
        return None
    return json.loads(res.stdout.strip().splitlines()[-1])


def fft_brackets():
    """the CPU figure's FFT is the f64 stand-in, not FFTW (absent).  Two library f64 FFTs on the
    same host, one thread each, 1024-point frames, as sanity brackets for it (Msamples/s, FFT only)."""
    out = {}
    n_frames = 16384
    rng = np.random.default_rng(0)
    x = rng.standard_normal((n_frames, 1024)) + 1j * rng.standard_normal((n_frames, 1024))
    try:
        np.fft.fft(x[:256], axis=1)
        t0 = time.perf_counter()
        np.fft.fft(x, axis=1)
        out["numpy_pocketfft_f64_1thread_msamples_per_s"] = n_frames * 1024 / (time.perf_counter() - t0) / 1e6
    except Exception:
        pass
    try:
        import torch
        old = torch.get_num_threads()
        torch.set_num_threads(1)
        xt = torch.from_numpy(x)
        torch.fft.fft(xt[:256], dim=1)
        t0 = time.perf_counter()
        torch.fft.fft(xt, dim=1)
        out["torch_f64_1thread_msamples_per_s"] = n_frames * 1024 / (time.perf_counter() - t0) / 1e6
        torch.set_num_threads(old)
    except Exception:
        pass
    return out


def cpu_sample_shape(cpu_seconds: float, workers: int):
    """A bounded sample of the workload: whole streams of 1 s (2 048 000 samples), enough of them for
    ~cpu_seconds of CPU work per worker at ~20 Msamples/s/core."""
    per_stream = FS
    streams_per_worker = max(1, int(round(cpu_seconds * 45e6 / per_stream)))      # ~45-55 Msamples/s/core measured
    return workers * streams_per_worker, per_stream


def reference_arm(args):
    """The reference's own CPU implementation of the path on the host cores (oracle/_ref), all cores:
    W untimed + exactly K timed steps, each step a bounded sample of the workload (whole streams of
    1 s of IQ, one process per core) sized so the run ends within a couple of minutes."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    workers = os.cpu_count() or 1
    per_stream = FS
    cal = run_ref_bench(workers, per_stream, workers)                 # calibration, untimed
    if cal is None:
        emit_line({"impl": "reference", "unavailable": "oracle/_ref/ref_bench could not be built or run"})
        return 0
    total_steps = max(1, args.steps) + max(0, args.warmup)
    budget_s = min(10.0, max(0.5, 90.0 / total_steps))                  # CPU wall time per step
    rounds = max(1, int(round(cal["msamples_per_s"] * 1e6 * budget_s / (per_stream * workers))))
    n_streams = workers * rounds
    secs = []
    for i in range(total_steps):
        r = run_ref_bench(n_streams, per_stream, workers)
        if r is None:
            emit_line({"impl": "reference", "unavailable": "oracle/_ref/ref_bench failed mid-run"})
            return 0
        if i >= args.warmup:
            secs.append(r["seconds"])
    value = n_streams * per_stream * len(secs) / sum(secs) / 1e6
    sample = (f"{n_streams} streams x {per_stream} samples per step through the unmodified spectrum.c/rf_decimator.c/"
              f"resample.c/audio_main.c (one process per core, {workers} cores); FFT inside spectrum.c is the f64 "
              f"stand-in (FFTW3 absent); {len(secs)} timed steps after {args.warmup} warm-up")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(secs), "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)
    return 0


# --------------------------------------------------------------------------------------
# the B200 arm
# --------------------------------------------------------------------------------------

_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, torchrun) also
    write there, so everything else is pointed at stderr and the line goes to the saved fd."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_line(obj) -> None:
    data = (json.dumps(obj) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    args = parse_args()
    claim_stdout()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    pkg = graft.load_package()
    pkg.init(local_rank)

    L = (args.samples_per_stream // TILE) * TILE
    sharding = pkg.sharding
    my_streams = sharding.streams_for_rank(args.streams, world, rank)
    n_local = len(my_streams)

    ring = pkg.StreamRing(n_local, L)
    # generate in slices to bound temporaries
    for i in range(0, n_local, 8):
        ring.batch[i:i + 8].copy_(synth_fm_device(torch, my_streams[i:i + 8], L, device))
    db = torch.empty((n_local, L // 1024, 1024), dtype=torch.float32, device=device)
    audio = torch.empty((n_local, L // 40), dtype=torch.float32, device=device)
    # the averaged-spectrum kernel writes straight into the collective's send buffer
    gatherer = sharding.SpectraGatherer(args.streams, world, rank, (1024,), torch.uint8, device, dist=dist)
    avg_u8 = gatherer.local
    avg_out = {"db_u8": avg_u8.view(n_local, 1, 1024)}
    avg_plan = pkg.SpectrumPlan(1024, K=6)
    stream = torch.cuda.current_stream()
    # The UI-side products (6-frame averaged u8 spectra and their NCCL gather) only read the IQ, so
    # they ride a second CUDA stream and overlap the tail of the chain kernel instead of sitting
    # between two chain kernels; the timed region ends only after both streams have drained.
    side = torch.cuda.Stream(device=device)
    ready = torch.cuda.Event()

    kern_events = []

    def step(timed: bool):
        stream.wait_stream(side)      # a new batch may not land before the previous step's readers are done
        ready.record(stream)          # this step's IQ is in place
        side.wait_event(ready)
        with torch.cuda.stream(side):
            avg_plan.exec(ring.batch, n_rows=1, db=False, db_u8=True, out=avg_out, stream=side)
            if world > 1:  # the one exchange: averaged u8 spectra of all 256 streams to rank 0 (NCCL)
                gatherer.gather()
        if timed:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        pkg.chain_exec(ring, db=db, audio=audio)                         # the dominant kernel
        if timed:
            e1.record(stream)
            kern_events.append((e0, e1))
        ring.carry()                  # stream state for the next batch (history bytes only)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step(False)
    sync_all()

    launches_before = pkg.launch_count()
    t_start = torch.cuda.Event(enable_timing=True)
    t_stop = torch.cuda.Event(enable_timing=True)
    sync_all()
    h0 = time.perf_counter()
    t_start.record(stream)
    for _ in range(args.steps):
        step(True)
    stream.wait_stream(side)
    t_stop.record(stream)
    sync_all()
    h1 = time.perf_counter()
    sampler.mark(h0, h1)
    clocks = sampler.stop() if rank == 0 else None
    launches = pkg.launch_count() - launches_before
    elapsed_ms = t_start.elapsed_time(t_stop)
    kern_ms = [a.elapsed_time(b) for a, b in kern_events]
    t = torch.tensor([elapsed_ms, statistics.mean(kern_ms)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, kern_ms_mean = t.tolist()

    total_samples_per_step = args.streams * L
    value = total_samples_per_step * args.steps / (elapsed_ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (per launch, this rank's share) ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    alg_bytes = BYTES_PER_SAMPLE * n_local * L
    achieved = alg_bytes / (kern_ms_mean * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "chain_fused_kernel (b200_chain_exec: 1024-pt spectra + FM branch, one pass over the IQ)",
                "kernel_ms": kern_ms_mean, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "frac_of_nominal_8TBs": achieved / 8000.0}
    # DRAM bytes per launch from the committed ncu capture of this kernel (profiles/traffic.json): the capture
    # is of the N = 1 default workload; other shapes get its measured bytes-per-sample times their samples
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            cap = json.load(open(prof))
            if cap.get("samples_per_launch") == n_local * L:
                roofline["traffic"] = cap.get("chain_traffic_bytes_per_launch")
            elif cap.get("dram_bytes_per_sample"):
                roofline["traffic"] = int(cap["dram_bytes_per_sample"] * n_local * L)
                roofline["traffic_note"] = "scaled from the N = 1 capture by samples per launch"
        except Exception:
            pass

    # ---- e2e: the host-buffer C-ABI call, PCIe copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        Le = (args.e2e_samples_per_stream // TILE) * TILE
        sess = pkg.Session(n_local, Le)
        h_iq = torch.empty((n_local, Le, 2), dtype=torch.uint8).pin_memory()
        h_db = torch.empty((n_local, Le // 1024, 1024), dtype=torch.float32).pin_memory()
        h_audio = torch.empty((n_local, Le // 40), dtype=torch.float32).pin_memory()
        h_iq.copy_(ring.batch[:, :Le].cpu())
        for _ in range(2):
            sess.chain(h_iq, Le, h_db, h_audio)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            sess.chain(h_iq, Le, h_db, h_audio)          # synchronous: returns when the results are on the host
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = tt.item()
        e2e = {"value": args.streams * Le * args.e2e_steps / dt / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(h_iq.numel()) * world,
               "d2h_bytes_per_step": int(h_db.numel() * 4 + h_audio.numel() * 4) * world,
               "samples_per_stream_per_step": Le, "steps": args.e2e_steps,
               "api": "b200_session_chain (host buffers, pinned; H2D + kernels + D2H inside)"}
        sess.close()
        del h_iq, h_db, h_audio

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        workers = os.cpu_count() or 1
        n_s, per = cpu_sample_shape(args.cpu_seconds, workers)
        r = run_ref_bench(n_s, per, workers)
        if r is not None:
            cpu_baseline = {"value": r["msamples_per_s"], "unit": UNIT, "cores": workers, "kind": "reference",
                            "sample": f"{n_s} streams x {per} samples, one process per core, unmodified reference C "
                                      f"(oracle/_ref) with the f64 stand-in FFT (FFTW3 absent); {r['seconds']:.2f} s"}
            r1 = run_ref_bench(max(1, n_s // workers), per, 1)
            if r1 is not None:
                cpu_baseline["value_1core"] = r1["msamples_per_s"]
            cpu_baseline["fft_only_brackets"] = fft_brackets()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, world), "roofline": roofline,
            "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        emit_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

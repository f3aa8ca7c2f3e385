#!/usr/bin/env python
"""bench.py -- IQ Msamples/s through the spectrum + FM chain (BASELINE.json metric).

Workload (BASELINE.json configs[4]): 256 independent synthetic dongle
streams, every stream pushed through the FULL chain:
  * a 1024-point power spectrum of EVERY frame -> float32 dB, display order (K = 1), and
  * CIC /10 -> FM discriminator -> two half-band decimators -> float32 audio (fs/40),
from ONE pass over the device-resident u8 IQ (b200_chain_exec), plus per launch the 6-frame averaged u8
dB spectrum of every stream (what the reference ships to its UI client), written straight into the
buffer that b200_comm_gather_rows (NCCL underneath) gathers to rank 0.  Streams are sharded across ranks
(stream s -> rank s mod G); there is no collective on the data path, only that gather of 1 KiB per stream.

A STEP is long on purpose: 256 streams x 2 097 152 000 samples (1024 s of IQ per stream), so that the
timed region of the driver's `--steps 20` lasts ~16 s on one GPU and >= 2 s on eight -- the headline
`value` is a SUSTAINED number (power cap and clocks included), not a burst.  Total work per step is the
same for every N ("scaling": "strong"); a rank holds 256/N streams and runs the step as 256/N launches
of the same 2.1 G samples (4.2 GB of IQ) each, so a launch is the same size on every rank.

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
LAUNCH_SAMPLES_PER_STREAM = TILE * 1600                  # 8 192 000 at N = 1 (x N at N ranks: same bytes per launch)
LAUNCHES_PER_STEP_N1 = 256
BYTES_PER_SAMPLE = 2.0 + 4.0 + 4.0 / 40.0                # u8 IQ in, f32 dB out (K = 1), f32 audio out / 40
PRODUCT_PERIOD = 512_000                                 # 250 ms of IQ: the reference's spectrum cadence (cbb_main.c:16)
METRIC = "IQ Msamples/s (spectrum+FM chain)"
UNIT = "Msamples/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=N_STREAMS)
    ap.add_argument("--launches-per-step", type=int, default=LAUNCHES_PER_STEP_N1,
                    help="launches of 2.1 G samples per step at N = 1 (a rank runs this / N)")
    ap.add_argument("--e2e-seconds", type=float, default=1.0, help="minimum length of each end-to-end timed region")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[1..3] kernels (N = 1 only)")
    return ap.parse_args()


def step_shape(args, world):
    """(samples per stream per launch on a rank, launches per step on a rank, samples per stream per step)."""
    per_launch = LAUNCH_SAMPLES_PER_STREAM * world
    launches = max(1, args.launches_per_step // world)
    return per_launch, launches, per_launch * launches


def workload_config(args, world):
    per_launch, launches, per_step = step_shape(args, world)
    return {
        "workload": "configs[4]: 256 independent synthetic dongle streams, full spectrum+FM chain "
                    "(1024-pt FFT of every frame -> f32 dB, K=1; CIC/10 -> FM discriminator -> 2x half-band -> f32 audio)",
        "streams": args.streams,
        "samples_per_stream_per_step": per_step,
        "step": f"{launches} launches per rank of {args.streams // world if world else args.streams} streams x {per_launch} samples "
                f"(the same device-resident 4.2 GB batch of IQ re-run; per launch also the 6-frame averaged u8 spectra + their gather)",
        "fft_points": 1024, "frames_averaged": 1, "down_factor": 10, "sample_rate_hz": FS,
        "sharding": f"stream s -> rank s mod {world}; b200_comm_gather_rows (NCCL) of 6-frame averaged u8 dB spectra to rank 0",
        "l2": "inputs (4.2 GB per launch) and outputs (8.6 GB) exceed the 126 MB L2; no flush needed",
        "algorithmic_bytes_per_sample": BYTES_PER_SAMPLE,
        "timed_region": "sustained: the whole region runs back to back (>= 2 s at every N with the driver's 20 steps)",
    }


# --------------------------------------------------------------------------------------
# synthetic IQ
# --------------------------------------------------------------------------------------

HEAD = 8192          # the first samples of a stream come from their own generator, so any rank can regenerate them


def _fm_piece(torch, sid, n, device, seed, t0, phase0):
    g = torch.Generator(device=device).manual_seed(seed)
    t = (torch.arange(n, dtype=torch.float64, device=device) + t0) / FS
    f1 = 1000.0 + 37.0 * (sid % 16)
    msg = 0.5 * (torch.sin(2 * np.pi * f1 * t) + torch.sin(2 * np.pi * 5000.0 * t))
    phase = phase0 + (2 * np.pi * 25_000.0 / FS) * torch.cumsum(msg, dim=0)
    noise = 2.0 * torch.randn((n, 2), dtype=torch.float32, device=device, generator=g)
    re = 127.5 + 100.0 * torch.cos(phase).float() + noise[:, 0]
    im = 127.5 + 100.0 * torch.sin(phase).float() + noise[:, 1]
    out = torch.stack([re.round().clamp(0, 255).to(torch.uint8), im.round().clamp(0, 255).to(torch.uint8)], dim=1)
    return out, float(phase[-1])


def synth_fm_head(torch, sid, device):
    return _fm_piece(torch, sid, HEAD, device, 1000 + int(sid), 0, 0.0)


def synth_fm_device(torch, stream_ids, n, device):
    """S3 generated on the device: per-stream seeds, FM with two message tones, 25 kHz peak
    deviation, amplitude 100, noise sigma 2, quantised to offset-binary u8.  Long streams repeat a 4 s period
    (phase-continuous up to the wrap), which bounds the f64 temporaries."""
    out = torch.empty((len(stream_ids), n, 2), dtype=torch.uint8, device=device)
    period = min(n, LAUNCH_SAMPLES_PER_STREAM)
    for i, sid in enumerate(stream_ids):
        head, ph = synth_fm_head(torch, sid, device)
        out[i, :HEAD] = head
        tail, _ = _fm_piece(torch, sid, period - HEAD, device, 5000 + int(sid), HEAD, ph)
        out[i, HEAD:period] = tail
        for pos in range(period, n, period):
            m = min(period, n - pos)
            out[i, pos:pos + m] = out[i, :m]
        del head, tail
    return out


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------

class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []          # (host time, csv line)

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

    def window(self, t0: float, t1: float):
        """Clock statistics of the samples taken between two host timestamps."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        lines = list(self.lines)
        inside = [(t, l) for t, l in lines if t0 <= t <= t1]
        if not inside and lines:      # region shorter than the sampling period: the closest sample
            mid = 0.5 * (t0 + t1)
            inside = [min(lines, key=lambda tl: abs(tl[0] - mid))]
        sm, smax, power, reasons, capped = [], [], [], set(), 0
        for _, line in inside:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(self.NAMES, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
                    capped += name == "sw_power_cap"
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "sm_mhz_min": min(sm) if sm else None, "power_w_max": max(power) if power else None,
                "power_w_median": statistics.median(power) if power else None, "samples": len(sm),
                "sw_power_cap_samples": capped, "reasons": sorted(reasons)}

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()


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
    ~cpu_seconds of CPU work per worker at ~50 Msamples/s/core."""
    per_stream = FS
    streams_per_worker = max(1, int(round(cpu_seconds * 45e6 / per_stream)))
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
    # the reference's own products (FM audio + one 6-frame averaged payload per 250 ms) on the same cores: the
    # counterpart of the B200 arm's e2e_products
    prod = run_ref_bench(n_streams * 4, per_stream, workers, mode="products")
    sample = (f"{n_streams} streams x {per_stream} samples per step through the unmodified spectrum.c/rf_decimator.c/"
              f"resample.c/audio_main.c (one process per core, {workers} cores); FFT inside spectrum.c is the f64 "
              f"stand-in (FFTW3 absent); {len(secs)} timed steps after {args.warmup} warm-up")
    config = workload_config(args, world)
    # this arm's step is a bounded SAMPLE of that workload: say what was actually run
    config["reference_arm_sample"] = {
        "streams_per_step": n_streams, "samples_per_stream_per_step": per_stream, "workers": workers,
        "note": "throughput of the reference's per-stream loop does not depend on stream length; the B200 arm's step "
                "(256 streams x 2.1 G samples) would take the CPU about ten minutes",
    }
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(secs), "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if prod is not None:
        line["e2e_products"] = {"value": prod["msamples_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0,
                                "d2h_bytes_per_step": 0,
                                "what": "FM audio + one 6-frame averaged payload per 512000 samples per stream "
                                        "(cbb_main.c:40-70,106-135), same cores"}
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


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def timed_launches(torch, fn, reps):
    """Mean milliseconds per call of `fn` over `reps` back-to-back calls (CUDA events on the current stream)."""
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run_other_configs(torch, pkg, ring, sampler, peak):
    """BASELINE.json configs[1..3] on one GPU, device-resident, each kernel alone: value (Msamples/s), launch time
    and HBM roofline fraction from the algorithmic bytes of SURVEY.md section 8d."""
    out = []
    dev = ring.buf.device

    def entry(name, workload, n_samples, bytes_per_sample, fn, reps):
        for _ in range(3):
            fn()
        h0 = time.perf_counter()
        ms = timed_launches(torch, fn, reps)
        h1 = time.perf_counter()
        gbs = n_samples * bytes_per_sample / (ms * 1e-3) / 1e9
        out.append({"name": name, "workload": workload, "value": n_samples / (ms * 1e-3) / 1e6, "unit": UNIT,
                    "kernel_ms": ms, "launches_timed": reps, "samples_per_launch": n_samples,
                    "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                 "algorithmic_bytes_per_sample": bytes_per_sample},
                    "clocks": sampler.window(h0, h1)})

    g = torch.Generator(device=dev).manual_seed(7)
    # configs[1]: batched spectrum only, 1 048 576 frames of 1024 / 4096 points, rectangular and Hann, f32 dB rows
    for N in (1024, 4096):
        frames = 1 << 20
        cap = torch.randint(0, 256, (256, frames // 256 * N, 2), dtype=torch.uint8, device=dev, generator=g)
        db = torch.empty((256, frames // 256, N), dtype=torch.float32, device=dev)
        for wname, win in (("rect", pkg.WINDOW_RECT), ("hann", pkg.WINDOW_HANN)):
            plan = pkg.SpectrumPlan(N, window=win)
            entry(f"spectrum{N}_{wname}", f"configs[1]: {frames} frames of {N}-point FFT ({wname}) -> f32 dB, "
                  f"256 captures x {frames // 256} frames, device-resident", frames * N, 6.0,
                  lambda: plan.exec(cap, db=True, out={"db": db}), 10)
            plan.close()
        del cap, db
    # configs[2]: FM chain only (CIC/10 -> discriminator -> 2 half-bands), the bench's own 2.1 G samples (>= 2^30)
    audio = torch.empty((ring.n_streams, ring.n_samples // 40), dtype=torch.float32, device=dev)
    entry("fm_chain_r10", f"configs[2]: FM broadcast chain only, {ring.n_streams} streams x {ring.n_samples} samples "
          "(>= 2^30), u8 IQ -> 51.2 kS/s f32 audio", ring.n_streams * ring.n_samples, 2.1,
          lambda: pkg.fm_exec(ring, audio=audio), 20)
    del audio
    # configs[3]: wideband spectrogram, 65536-point Hann, 50 % overlap, one 2^31-sample capture
    n_cap = 1 << 31
    cap = torch.randint(0, 256, (1, n_cap, 2), dtype=torch.uint8, device=dev, generator=g)
    plan = pkg.SpectrumPlan(65536, hop=32768, window=pkg.WINDOW_HANN)
    rows = plan.rows(n_cap)
    db = torch.empty((1, rows, 65536), dtype=torch.float32, device=dev)
    entry("spectrogram65536_hann_50pct", f"configs[3]: 65536-point Hann spectrogram, hop 32768, one capture of 2^31 samples "
          f"({rows} rows) -> f32 dB", rows * 32768, 10.0, lambda: plan.exec(cap, db=True, out={"db": db}), 5)
    plan.close()
    del cap, db
    torch.cuda.empty_cache()
    return out


def compat_latency(pkg):
    """What one call of the zero-change drop-in costs next to the code it replaces (correctness path, latency-bound)."""
    from oracle import pyoracle as po
    out = {}
    iq = pkg.synth.s2_tones(1024, N=1024, seed=1)
    ps = np.zeros(1024, np.float64)
    s = pkg.spectrum_alloc(1024)
    for _ in range(20):
        pkg.spectrum_add_cmplx_u8(s, iq, ps, 1024)
    t0 = time.perf_counter()
    n = 300
    for _ in range(n):
        pkg.spectrum_add_cmplx_u8(s, iq, ps, 1024)
    b200_us = (time.perf_counter() - t0) / n * 1e6
    pkg.spectrum_free(s)
    ref_us = None
    blk = pkg.synth.s3_fm(204800, seed=3)
    rd = pkg.RfDecimator()
    rd.set_parameters(2048000.0, 10)
    for _ in range(3):
        rd.decimate_cmplx_u8(blk)
    t0 = time.perf_counter()
    for _ in range(20):
        rd.decimate_cmplx_u8(blk)
    b200_blk_us = (time.perf_counter() - t0) / 20 * 1e6
    rd.free()
    ref_blk_us = None
    if po.have_ref():
        ref = po.Ref()
        h = ref.lib.spectrum_alloc(1024)
        flat = np.ascontiguousarray(iq).reshape(-1)
        for _ in range(20):
            ref.lib.spectrum_add_cmplx_u8(h, flat, ps, 1024)
        t0 = time.perf_counter()
        n = 2000
        for _ in range(n):
            ref.lib.spectrum_add_cmplx_u8(h, flat, ps, 1024)
        ref_us = (time.perf_counter() - t0) / n * 1e6
        ref.lib.spectrum_free(h)
        d = ref.lib.rf_decimator_alloc()
        ref.lib.rf_decimator_set_parameters(d, 2048000.0, 10)
        fb = np.ascontiguousarray(blk).reshape(-1)
        for _ in range(3):
            ref.lib.rf_decimator_decimate_cmplx_u8(d, fb, 204800)
        t0 = time.perf_counter()
        for _ in range(50):
            ref.lib.rf_decimator_decimate_cmplx_u8(d, fb, 204800)
        ref_blk_us = (time.perf_counter() - t0) / 50 * 1e6
        ref.lib.rf_decimator_free(d)
    out["spectrum_add_cmplx_u8_1024"] = {"b200_us": b200_us, "reference_cpu_us": ref_us}
    out["rf_decimator_block_204800_no_callbacks"] = {"b200_us": b200_blk_us, "reference_cpu_us": ref_blk_us}
    out["note"] = ("the reference-named calls are synchronous 2 KB / 400 KB round trips over PCIe: correct, and slower than "
                   "the CPU code they replace; throughput goes through the batched entry points (value, e2e)")
    return out


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

    L, launches_per_step, per_step = step_shape(args, world)
    my_streams = [pkg.shard_stream(args.streams, world, rank, i) for i in range(pkg.shard_count(args.streams, world, rank))]
    n_local = len(my_streams)

    ring = pkg.StreamRing(n_local, L)
    for i in range(0, n_local, 4):      # generate in slices to bound temporaries
        ring.batch[i:i + 4].copy_(synth_fm_device(torch, my_streams[i:i + 4], L, device))
    db = torch.empty((n_local, L // 1024, 1024), dtype=torch.float32, device=device)
    audio = torch.empty((n_local, L // 40), dtype=torch.float32, device=device)
    # the one exchange of the path, through the C ABI: the averaged-spectrum kernel writes straight into the
    # buffer b200_comm_gather_rows sends from
    comm = None
    if world > 1:
        ids = [pkg.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = pkg.Comm(ids[0], world, rank)
    avg_u8 = torch.zeros((max(n_local, 1), 1024), dtype=torch.uint8, device=device)
    gathered = torch.zeros((args.streams, 1024), dtype=torch.uint8, device=device) if rank == 0 else None
    avg_out = {"db_u8": avg_u8[:n_local].view(n_local, 1, 1024)}
    avg_plan = pkg.SpectrumPlan(1024, K=6)
    stream = torch.cuda.current_stream()
    # The UI-side products (6-frame averaged u8 spectra and their gather) only read the IQ, so they ride a
    # second CUDA stream next to the chain kernel; the timed region ends only after both streams have drained.
    side = torch.cuda.Stream(device=device)
    ready = torch.cuda.Event()

    kern_events = []

    def launch(timed: bool):
        ready.record(stream)          # this launch's IQ (and the carried history) is in place
        side.wait_event(ready)
        with torch.cuda.stream(side):
            avg_plan.exec(ring.batch, n_rows=1, db=False, db_u8=True, out=avg_out, stream=side)
            if comm is not None:
                comm.gather_rows(avg_u8[:n_local], args.streams, gathered, root=0, stream=side)
        if timed:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        pkg.chain_exec(ring, db=db, audio=audio)                         # the dominant kernel
        if timed:
            e1.record(stream)
            kern_events.append((e0, e1))
        # stream state for the next batch: only the history bytes in FRONT of the batch move, which the side
        # stream's kernels never read -- the gather is not on the critical path of the next launch
        ring.carry()

    def step(timed: bool):
        for _ in range(launches_per_step):
            launch(timed)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    peak, peak_src = hbm_peak()
    alg_bytes = BYTES_PER_SAMPLE * n_local * L

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- burst: ten launches on a cool chip, before anything long has run ----
    for _ in range(3):
        launch(False)
    sync_all()
    burst_ms = timed_launches(torch, lambda: pkg.chain_exec(ring, db=db, audio=audio), 10)
    tb = torch.tensor([burst_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
    burst_ms = tb.item()

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
    clocks = sampler.window(h0, h1) if rank == 0 else None
    launches = pkg.launch_count() - launches_before
    elapsed_ms = t_start.elapsed_time(t_stop)
    kern_ms = [a.elapsed_time(b) for a, b in kern_events]
    t = torch.tensor([elapsed_ms, statistics.mean(kern_ms)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, kern_ms_mean = t.tolist()

    total_samples_per_step = args.streams * per_step
    value = total_samples_per_step * args.steps / (elapsed_ms * 1e-3) / 1e6

    # ---- the gather, checked: every row rank 0 received against a local recompute of that stream's first frames ----
    gather_check = None
    if rank == 0:
        heads = torch.stack([synth_fm_head(torch, s, device)[0] for s in range(args.streams)])
        want = avg_plan.exec(heads, n_rows=1, db=False, db_u8=True)["db_u8"][:, 0]
        got = gathered if world > 1 else None
        if world == 1:
            got = avg_u8[:n_local]
        torch.cuda.synchronize()
        bad = int((got != want).any(dim=1).sum())
        gather_check = "ok" if bad == 0 else f"{bad} of {args.streams} rows differ"

    # ---- roofline of the dominant kernel (per launch, this rank's share), sustained and burst ----
    achieved = alg_bytes / (kern_ms_mean * 1e-3) / 1e9
    burst = alg_bytes / (burst_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "chain_jobs_kernel (b200_chain_exec: 1024-pt spectra + FM branch, one pass over the IQ)",
                "kernel_ms": kern_ms_mean, "launches_timed": len(kern_ms), "algorithmic_bytes_per_launch": alg_bytes,
                "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0,
                "measured_over": "every launch of the timed region (sustained: seconds of back-to-back launches, power cap included)",
                "burst": {"launches": 10, "kernel_ms": burst_ms, "gsamples_per_s": n_local * L * world / (burst_ms * 1e-3) / 1e9,
                          "achieved": burst, "frac": burst / peak, "when": "ten launches on a cool chip before the warm-up"}}
    # DRAM bytes per launch from the committed ncu capture of this kernel (profiles/traffic.json)
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
    # the issue ceiling of the instruction stream (profiles/issue_ceiling.json: measured by tools/ubench_fft.cu)
    ceil_path = os.path.join(ROOT, "profiles", "issue_ceiling.json")
    if os.path.exists(ceil_path):
        try:
            c = json.load(open(ceil_path))
            roofline["issue_ceiling_gsamples"] = c.get("chain_gsamples")
            roofline["issue_ceiling_derivation"] = c.get("derivation")
        except Exception:
            pass

    # ---- e2e: the host-buffer C-ABI calls, PCIe copies inside the timed region, >= 1 s each ----
    e2e = e2e_products = None
    if not args.no_e2e:
        Le = LAUNCH_SAMPLES_PER_STREAM                     # the headline batch: 8 192 000 samples per stream per call
        sess = pkg.Session(n_local, Le)
        h_iq = torch.empty((n_local, Le, 2), dtype=torch.uint8).pin_memory()
        h_db = torch.empty((n_local, Le // 1024, 1024), dtype=torch.float32).pin_memory()
        h_audio = torch.empty((n_local, Le // 40), dtype=torch.float32).pin_memory()
        h_iq.copy_(ring.batch[:, :Le].cpu())
        sess.chain(h_iq, Le, h_db, h_audio)

        def timed_calls(fn, min_seconds):
            sync_all()
            t0 = time.perf_counter()
            calls = 0
            while True:
                fn()                                      # synchronous: returns when the results are on the host
                calls += 1
                done = torch.tensor([1.0 if time.perf_counter() - t0 >= min_seconds and calls >= 3 else 0.0], device=device)
                if world > 1:
                    dist.all_reduce(done, op=dist.ReduceOp.MIN)     # every rank makes the same number of calls
                if done.item() > 0:
                    break
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return calls, dt.item()

        calls, dt = timed_calls(lambda: sess.chain(h_iq, Le, h_db, h_audio), args.e2e_seconds)
        e2e = {"value": args.streams * Le * calls / dt / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(h_iq.numel()) * world,
               "d2h_bytes_per_step": int(h_db.numel() * 4 + h_audio.numel() * 4) * world,
               "samples_per_stream_per_step": Le, "steps": calls, "seconds": dt,
               "api": "b200_session_chain (host buffers, pinned; H2D + kernels + D2H of per-frame f32 dB rows + audio inside)"}
        sess.close()
        del h_db
        # the reference's own products: audio + one 6-frame averaged payload per stream per 250 ms of IQ
        Lp = PRODUCT_PERIOD // TILE * TILE
        sess = pkg.Session(n_local, Lp)
        h_avg = torch.empty((n_local, 1024), dtype=torch.uint8).pin_memory()
        h_aud = h_audio[:, :Lp // 40].contiguous().pin_memory()
        h_iqp = h_iq[:, :Lp].contiguous().pin_memory()
        sess.products(h_iqp, Lp, h_aud, h_avg)
        calls, dt = timed_calls(lambda: sess.products(h_iqp, Lp, h_aud, h_avg), args.e2e_seconds)
        e2e_products = {"value": args.streams * Lp * calls / dt / 1e6, "unit": UNIT,
                        "h2d_bytes_per_step": int(h_iqp.numel()) * world,
                        "d2h_bytes_per_step": int(h_aud.numel() * 4 + h_avg.numel()) * world,
                        "samples_per_stream_per_step": Lp, "steps": calls, "seconds": dt,
                        "api": "b200_session_products (host buffers in; FM audio + the 6-frame averaged payload bytes of "
                               "cbb_main.c:106-135 back: 0.1 B per sample over PCIe instead of 4.1)"}
        sess.close()
        del h_iq, h_audio, h_avg, h_aud, h_iqp

    # ---- configs[1..3], the compat path's latency and the CPU baseline (rank 0, N = 1 only) ----
    other = None
    compat = None
    cpu_baseline = None
    if rank == 0 and world == 1:
        if not args.no_configs:
            other = run_other_configs(torch, pkg, ring, sampler, peak)
        try:
            compat = compat_latency(pkg)
        except Exception as e:      # never lose the line over a side measurement
            compat = {"error": repr(e)}
        if not args.no_cpu_baseline:
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
                rp = run_ref_bench(n_s * 4, per, workers, mode="products")
                if rp is not None:
                    cpu_baseline["products_value"] = rp["msamples_per_s"]
                cpu_baseline["fft_only_brackets"] = fft_brackets()
    if rank == 0:
        sampler.stop()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, world), "roofline": roofline,
            "cpu_baseline": cpu_baseline, "e2e": e2e, "e2e_products": e2e_products, "gpu_launches": int(launches),
            "clocks": clocks, "gather_check": gather_check, "timed_region_s": elapsed_ms * 1e-3,
            "configs": other, "compat_latency_us": compat,
        }
        emit_line(line)
    if comm is not None:
        comm.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

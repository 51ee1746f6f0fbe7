#!/usr/bin/env python3
"""bench.py - 48 kHz samples/s through encode -> quantize -> lookup -> decode (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload symad|v1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch of synthetic utterances (BASELINE configs[1]:
symAD_vctk_48000_hop300, 64 x 48000 samples, fp32, per GPU).  Utterances are independent, so N GPUs
each run their own 64-utterance shard with no data-path collective (weak scaling);
NCCL is used only for the timing barrier and the max-over-ranks of the device time.

`value`   : whole-job samples/s, inputs resident in HBM, CUDA events on the launching stream.
`e2e`     : same metric through the reference-facing call with HOST buffers (adec_codec_host: H2D of the
            waveforms + the four calls + D2H of indices and waveforms inside the timed region).
`roofline`: HBM roofline under SURVEY.md 8(d)'s per-conv-layer algorithmic byte model
            (9,323.2 B/sample for symAD fp32) against MEASURED_PEAKS.json's copy bandwidth.
`--impl reference`: the reference's own CPU implementation of the path.  The reference is pure Python on
            torch CPU ops and cannot travel to the GPU box, so this leg times the oracle port
            (oracle/audiodec_oracle.py: the same torch CPU ops in the reference's order) on all host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SAMPLE_RATE = 48000
T_SAMPLES = 48000
BATCH_PER_GPU = 64
# algorithmic bytes / FLOPs per input sample, fp32 activations, per-conv-layer model
ALG_BYTES_PER_SAMPLE = {"symad": 9323.2, "v1": (1398192 + 576 + 4183472) / 300.0}
ALG_FLOP_PER_SAMPLE = {"symad": 549432.0, "v1": 2265247.0}
ALG_BYTES_PER_SAMPLE["stream_v1"] = ALG_BYTES_PER_SAMPLE["v1"]
ALG_FLOP_PER_SAMPLE["stream_v1"] = ALG_FLOP_PER_SAMPLE["v1"]
FFMA_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12     # not measured; informational


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock, power and throttle reasons DURING the timed region (B200_PROFILING.md recipe).  Default source is NVML in a
    thread of this process (the library nvidia-smi itself reads; no subprocess, 20 ms period); ADEC_BENCH_SAMPLER=smi runs the
    recipe's `nvidia-smi -lms 100` loop instead, =off disables sampling."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    # nvmlClocksEventReasons bits (nvml.h)
    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.mode = os.environ.get("ADEC_BENCH_SAMPLER", "nvml")
        self._stop = threading.Event()
        self._thread = None

    def start(self):
        if self.mode == "off":
            return
        if self.mode == "nvml":
            try:
                import pynvml
                pynvml.nvmlInit()
                # CUDA_VISIBLE_DEVICES may renumber: address the device by its PCI bus id
                import torch
                pr = torch.cuda.get_device_properties(self.index)
                try:
                    h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0")
                except Exception:
                    vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                    phys = int(vis.split(",")[self.index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.index
                    h = pynvml.nvmlDeviceGetHandleByIndex(phys)
                self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

                def loop():
                    while not self._stop.is_set():
                        try:
                            sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                            pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1e3
                            try:
                                rs = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                            except Exception:
                                rs = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                            self.rows.append((time.time(), (sm, pw, rs)))
                        except Exception:
                            pass
                        self._stop.wait(0.02)
                self._thread = threading.Thread(target=loop, daemon=True)
                self._thread.start()
                return
            except Exception:
                self.mode = "smi"          # NVML binding unavailable: fall back to the nvidia-smi loop
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.mode == "off":
            return {"sampler": "off"}
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=1.0)
            rows = [r for (t, r) in self.rows if t0 <= t <= t1] or [r for (_, r) in self.rows]
            if not rows:
                return None
            reasons = sorted({name for (_, _, rs) in rows for name, bit in self.REASONS if rs & bit})
            return {"sm_mhz": statistics.median(r[0] for r in rows), "sm_max_mhz": self._max, "reasons": reasons,
                    "samples": len(rows), "power_w_max": max(r[1] for r in rows), "sampler": "nvml, 20 ms period, in-process thread"}
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.2 and len(r) >= 9] or [r for (_, r) in self.rows if len(r) >= 9]
        if not rows:
            return None
        try:
            sm = [float(r[1]) for r in rows]
            reasons = set()
            for r in rows:
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons),
                    "samples": len(rows), "power_w_max": max(float(r[3]) for r in rows), "sampler": "nvidia-smi -lms 100"}
        except Exception:
            return None


def build_codec(workload, device):
    import torch
    from audiodec_b200 import synthetic as S
    from audiodec_b200.codec import HiFiGANStreamGenerator, SymADStreamGenerator
    sd = S.symad_state_dict(seed=0)
    objs = []
    for _ in range(2):
        g = SymADStreamGenerator(**S.SYMAD_PARAMS)
        g.load_state_dict(sd)
        objs.append(g.eval().to(device))
    if workload == "v1":
        d = HiFiGANStreamGenerator(**S.HIFIGAN_V1_PARAMS)
        d.load_state_dict(S.hifigan_state_dict(seed=1))
    else:
        d = SymADStreamGenerator(**S.SYMAD_PARAMS)
        d.load_state_dict(sd)
    d = d.eval().to(device)
    tx, rx = objs
    tx.initial_encoder(8192, device)                       # bin/stream.py:61
    d.initial_decoder(rx.initial_encoder(8192, device))    # bin/stream.py:70,76
    torch.cuda.synchronize(device)
    return tx, rx, d


def run_ours(args):
    import torch
    import torch.distributed as dist
    from audiodec_b200.codec import codec_host

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, T = BATCH_PER_GPU, T_SAMPLES
    if args.workload == "stream_v1":
        B, T = 256, 1500          # demoStream.py:28 default frame size, 256 concurrent streams
    tx, rx, dec = build_codec("v1" if args.workload == "stream_v1" else args.workload, dev)

    # synthetic inputs: 0.1*randn, seed 1337 (+rank); several distinct resident batches
    gen = torch.Generator().manual_seed(1337 + rank)
    n_in = 4
    x_host = [(0.1 * torch.randn(B, 1, T, generator=gen)).pin_memory() for _ in range(n_in)]
    x_dev = [x.to(dev) for x in x_host]

    def step(i):
        z = tx.encode(x_dev[i % n_in])
        idx = tx.quantize(z)
        zq = rx.lookup(idx)
        return dec.decode(zq)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # pre-warm: clocks / power state settle over the first ~second of load; these steps are not counted in W
    t_pre = time.time()
    while time.time() - t_pre < 1.5:
        step(0)
        torch.cuda.synchronize(dev)
    for i in range(args.warmup):
        step(i)
    barrier()
    l0 = tx.launch_count + rx.launch_count + dec.launch_count
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record()
    for i in range(args.steps):
        y = step(i)
    e1.record()
    barrier()
    w1 = time.time()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop(w0, w1) if rank == 0 else None
    launches = (tx.launch_count + rx.launch_count + dec.launch_count - l0)
    assert torch.isfinite(y).all()

    # ---- e2e: host buffers through adec_codec_host (H2D + 4 calls + D2H inside the timed region)
    for i in range(min(args.warmup, 2)):
        codec_host(tx, dec, x_host[i % n_in], reuse_buffers=True)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for i in range(args.steps):
        idx_h, y_h = codec_host(tx, dec, x_host[i % n_in], reuse_buffers=True)
    e3.record()
    barrier()
    ms_e2e = max_over_ranks(e2.elapsed_time(e3))
    F = idx_h.shape[-1]

    # ---- per-launch CUDA-event timing of one extra step (same inputs, same stream): which kernel dominates, and its
    #      achieved algorithmic GB/s.  Outside the timed region so the events do not perturb `value`.
    prof = None
    if rank == 0:
        tx.profile(True), dec.profile(True)
        for i in range(2):
            step(i)
        torch.cuda.synchronize(dev)
        rows = tx.profile_report() + dec.profile_report()
        tx.profile(False), dec.profile(False)
        agg = {}
        for name, ms, nbytes in rows:
            a = agg.setdefault(name, [0, 0.0, nbytes])
            a[0] += 1
            a[1] += ms
        tot = sum(v[1] for v in agg.values())
        top = sorted(agg.items(), key=lambda kv: -kv[1][1])
        if args.breakdown:
            for k, v in agg.items():
                print(f"  {k:44s} {v[1] / v[0]:8.3f} ms  {v[2] / (v[1] / v[0]) / 1e6:8.1f} GB/s(alg)", file=sys.stderr)
            print(f"  sum of launches per step: {tot / 2:.3f} ms", file=sys.stderr)
        dname, (dn, dms, dbytes) = top[0]
        step_ms = ms_total / args.steps                 # shares are of the TIMED step (which also holds the RVQ / lookup launches)
        prof = {"kernel": dname, "launch_ms": dms / dn, "alg_bytes_per_launch": dbytes, "share_of_step": (dms / dn) / step_ms,
                "conv_kernels_share_of_step": sum(v[1] / 2 for k, v in agg.items() if "res_units" in k or ".conv" in k or "project" in k
                                                  or "blocks" in k or "upsamples" in k) / step_ms,
                "top5": [{"op": k, "ms": v[1] / v[0], "GBps": v[2] / (v[1] / v[0]) / 1e6} for k, v in top[:5]]}

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    samples_per_step = world * B * T
    value = samples_per_step * args.steps / (ms_total / 1e3)
    e2e_value = samples_per_step * args.steps / (ms_e2e / 1e3)
    peak, peak_src = measured_peaks()
    per_gpu = value / world
    alg_b = ALG_BYTES_PER_SAMPLE[args.workload]
    achieved = alg_b * per_gpu / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if prof and os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        for key, val in tj.items():
            if key != "_comment" and key in prof["kernel"]:
                traffic = val
    k_achieved = prof["alg_bytes_per_launch"] / (prof["launch_ms"] * 1e-3) / 1e9 if prof else achieved
    conv_path = os.environ.get("ADEC_CONV_PATH", "tc")
    line = {
        "metric": "48 kHz audio samples/s, encode+quantize+lookup+decode (% HBM roofline in `roofline`)",
        "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (0.1*randn waveforms, seeded synthetic checkpoint; the reference ships no weights)",
        "config": {"workload": ({"symad": "symAD_vctk_48000_hop300", "v1": "AudioDec_v1 (symAD enc + HiFi-GAN v1)",
                                 "stream_v1": "libritts_v1 streaming: 256 streams x 1500-sample chunks @ 24 kHz (one chunk per step)"}[args.workload])
                   + f" batch={B}x{T} per GPU, fp32 (BASELINE configs[{'1' if args.workload == 'symad' else '2, fp32' if args.workload == 'v1' else '3'}])",
                   "utterances_per_gpu": B, "samples_per_utterance": T, "parallelism": f"independent utterance shards x{world}, no collective",
                   "l2": "per-step activation working set ~3 GB per GPU >> 126 MB L2; inputs rotate over 4 distinct resident batches",
                   "realtime_factor_per_gpu": per_gpu / (24000 if args.workload == "stream_v1" else SAMPLE_RATE)},
        "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": B * T * 4,
                "d2h_bytes_per_step": B * F * 300 * 4 + 8 * B * F * 8, "ms_per_step": ms_e2e / args.steps,
                "api": "audiodec_b200.codec.codec_host -> adec_codec_host (pinned host buffers, per GPU)"},
        "roofline": {"bound": "hbm", "achieved": k_achieved, "peak": peak, "unit": "GB/s", "frac": k_achieved / peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "kernel": (("tc_conv_persist_kernel (tcgen05 3xTF32)" if conv_path != "ffma" else "conv_gemm_kernel (fp32 FFMA)")
                                + " launch of " + (prof["kernel"] if prof else "?")),
                     "kernel_launch_ms": prof["launch_ms"] if prof else None,
                     "kernel_alg_bytes_per_launch": prof["alg_bytes_per_launch"] if prof else None,
                     "kernel_share_of_step": prof["share_of_step"] if prof else None,
                     "conv_kernels_share_of_step": prof["conv_kernels_share_of_step"] if prof else None,
                     "top5_launches": prof["top5"] if prof else None,
                     "step": {"achieved": achieved, "frac": achieved / peak,
                              "model": f"whole step: {alg_b:.1f} algorithmic B/sample (SURVEY.md 8(d) per-conv-layer model) x samples/s per GPU"},
                     "useful_tflops": ALG_FLOP_PER_SAMPLE[args.workload] * per_gpu / 1e12,
                     "fp32_ffma_peak_tflops_nominal": FFMA_PEAK_TFLOPS},
        "conv_path": conv_path,
        "clocks": clocks,
    }
    if args.cpu_baseline and world == 1:          # reported at N=1 only (rank 0); the reference arm covers every N
        line["cpu_baseline"] = cpu_baseline(args.workload, n_utt=args.cpu_utts, threads=best_cpu_threads(args.workload))
    print(json.dumps(line), flush=True)


def cpu_baseline(workload, n_utt=4, seconds=1.0, threads=None, budget_s=15.0):
    """The reference's CPU path (oracle port: same torch CPU ops) on a bounded sample: up to `n_utt` utterances of
    `seconds` s, one after another (the reference's streaming path is batch-1 only, conv_layer.py:144-146), cut short
    after `budget_s` seconds of host work (never below 2 utterances) so a slow host cannot stretch the run."""
    import torch
    from audiodec_b200 import synthetic as S
    from oracle import audiodec_oracle as O
    if threads:
        torch.set_num_threads(threads)
    cores = torch.get_num_threads()
    sd = S.symad_state_dict(seed=0)
    if workload == "v1":
        codec = O.CodecOracle(S.SYMAD_PARAMS, sd, S.HIFIGAN_V1_PARAMS, S.hifigan_state_dict(seed=1))
    else:
        codec = O.CodecOracle(S.SYMAD_PARAMS, sd)
    torch.manual_seed(1337)
    T = int(seconds * SAMPLE_RATE)
    xs = [0.1 * torch.randn(1, 1, T) for _ in range(min(n_utt, 8))]     # distinct inputs, cycled
    with torch.no_grad():
        codec.run(xs[0][:, :, :6000])          # warm the thread pool / oneDNN primitive cache
        t0 = time.perf_counter()
        done = 0
        while done < n_utt:
            codec.run(xs[done % len(xs)])
            done += 1
            if done >= 2 and time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
    n_utt = done
    return {"value": n_utt * T / dt, "unit": "samples/s", "cores": cores, "kind": "port", "utterances": n_utt,
            "sample": f"{n_utt} utterances x {seconds:g} s @ 48 kHz, per-utterance loop (reference streaming path is batch-1), "
                      f"torch {torch.__version__} CPU fp32, {cores} threads; {dt:.2f} s wall",
            "realtime_factor": n_utt * T / dt / SAMPLE_RATE}


def best_cpu_threads(workload):
    """The reference's demo default is 4 threads (demoFile.py:28); more threads help up to a point and then hurt (small
    convs, oversubscription).  Pick the fastest of a few counts on a 0.25 s clip so the CPU arm is not handicapped."""
    import torch
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (4, 8, 16, 32, 64, ncpu) if c <= ncpu})
    best, best_v = cands[0], 0.0
    for c in cands:
        v = cpu_baseline(workload, n_utt=1, seconds=0.25, threads=c)["value"]
        if v > best_v:
            best, best_v = c, v
    torch.set_num_threads(best)
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import torch
    world = int(os.environ.get("WORLD_SIZE", 1))
    best_cpu_threads(args.workload)
    per = []
    n_utt = args.ref_utts                       # 16: ~1 s of host work per step on the box's cores, K=10 steps stay well under a minute
    for _ in range(args.warmup):
        cpu_baseline(args.workload, n_utt=min(2, n_utt))
    t_all0 = time.perf_counter()
    last = None
    for _ in range(args.steps):
        last = cpu_baseline(args.workload, n_utt=n_utt)
        per.append(last["value"])
    dt = time.perf_counter() - t_all0
    value = statistics.median(per)
    line = {
        "impl": "reference",
        "metric": "48 kHz audio samples/s, encode+quantize+lookup+decode (% HBM roofline in `roofline`)",
        "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * last["utterances"] * T_SAMPLES / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (same seeded checkpoint and waveform distribution as the CUDA arm)",
        "config": {"workload": ("symAD_vctk_48000_hop300" if args.workload == "symad" else "AudioDec_v1") +
                   f" batch={BATCH_PER_GPU}x{T_SAMPLES} per GPU, fp32 (BASELINE configs[1]); each step a bounded sample of it",
                   "note": "reference = pure-Python torch-CPU path; timed via the oracle port (identical torch ops/order) because "
                           "/root/reference does not exist on the GPU box; rank 0 only"},
        "cpu_baseline": dict(last, value=value),
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": dt,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="symad", choices=["symad", "v1", "stream_v1"],
                    help="symad = BASELINE configs[1] (default); v1 = configs[2] shape in fp32; stream_v1 = configs[3]: 256 streams x 1500-sample chunks @ 24 kHz")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--cpu-utts", type=int, default=192,
                    help="utterances of the bounded CPU sample (192 x 1 s = three steps' worth of audio, 10-15 s of host work)")
    ap.add_argument("--ref-utts", type=int, default=16, help="--impl reference: utterances per step (each step time-bounded at 15 s)")
    ap.add_argument("--breakdown", action="store_true", help="print per-launch CUDA-event times to stderr")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

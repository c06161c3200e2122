#!/usr/bin/env python
"""Headline benchmark: audio-seconds coded per second (encode + decode, 22.05 kHz, 3 kbps).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision 0|1]

Workload = BASELINE.json configs[1]: var-bitrate codec, batch 256 x 10 s synthetic 22.05 kHz
utterances per GPU at 3 kbps (35 bits/frame), synthetic random-init weights with the reference's
checkpoint schema (the shipped checkpoints are git-LFS pointers).  One step = encode(x) followed
by decode(codes, L) on one batch (the two recurrences are NOT shared).

  value  device-resident: inputs already in HBM, CUDA-event timed, max over ranks
  e2e    through the public facade with pinned HOST tensors: H2D of x, encode, D2H of codes,
         H2D of codes, decode, D2H of audio, all inside the timed region
  roofline  algorithmic FLOPs (10.478 GFLOP per audio-second) / device time against the
            measured sustained bf16 tensor peak of MEASURED_PEAKS.json; per-stage breakdown attached
  cpu_baseline / --impl reference  the CPU oracle (port of the reference, PyTorch CPU, all host
            threads) on a bounded sample of the same workload, rank 0 only

N > 1: launched by torchrun, one rank per GPU, every rank codes its own 256-utterance shard
(weak scaling); the only collective is the final all-gather of codes and audio, inside the timed
region.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS, HOP, Z = 22050, 256, 64
GFLOP_PER_AUDIO_S = 10.478            # BASELINE.md section 2 (encode 4.051 + decode 6.427)
MFLOP_FRAME = dict(logmel=0.135332, encode=2 * 23.445504, decode_mel=2 * 18.055168, vocode=2 * 19.255296)
# dominant kernel = the persistent recurrent kernel; its algorithmic work is the state-dependent part of a frame
# (20 217 856 MAC encode, 11 698 176 MAC decode; the hoisted layers run in separate GEMM kernels)
MFLOP_FRAME_RECURRENT = dict(encode=2 * 20.217856, decode_mel=2 * 11.698176)
METRIC = "audio-sec coded/sec (encode+decode, 22.05 kHz, 3 kbps)"
UNIT = "audio-s/s"


def workload_config(args, extra=None):
    cfg = {
        "workload": "configs[1]: var-bitrate codec, batch 256 x 10 s synthetic 22.05 kHz utterances at 3 kbps",
        "batch_per_gpu": args.batch, "seconds": args.seconds, "bitrate_bps": 3000, "bits_per_frame": 35,
        "weights": "synthetic random-init, reference checkpoint schema",
        "l2": "inputs and intermediates are larger than L2 (226 MB of audio per step)",
        "parallelism": f"utterance-sharded x{args.gpus}",
    }
    cfg.update(extra or {})
    return cfg


def synth_batch(B, L, seed):
    g = torch.Generator().manual_seed(seed)
    return (0.1 * torch.randn(B, L, generator=g)).clamp_(-1, 1)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, device_index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop_evt = threading.Event()
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # NVML missing: report that instead of inventing clocks
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_oracle_run(args, steps, warmup, budget_s):
    """Times the CPU oracle (PyTorch CPU port of the reference) on a bounded sample; returns dict."""
    from bernoulli_var_speech_codec_b200.synth import write_synthetic_checkpoints
    from oracle.codec_oracle import OracleCodec
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ck = write_synthetic_checkpoints(os.environ.get("BVC_CKPT_DIR", "/tmp/bvc_ckpts"), seed=1, sharpen=30.0)
    oracle = OracleCodec(os.path.join(ROOT, "configs", "config_varBitRate.toml"), *ck)
    L = int(args.seconds * FS)
    # probe one short utterance to size the sample so that (warmup + steps) passes fit the budget
    xp = synth_batch(1, FS, 99)
    t0 = time.perf_counter()
    oracle.forward(xp, 3000)
    per_utt_s = (time.perf_counter() - t0) * args.seconds      # ~linear in duration at B=1
    Bs = int(max(1, min(8, budget_s / max(1e-3, per_utt_s * 0.6 * (steps + warmup)))))
    x = synth_batch(Bs, L, 1234)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        codes = oracle.encode(x, 3000)
        wav = oracle.decode(codes, L)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    assert wav.shape == x.shape
    total = sum(times)
    return {"value": Bs * args.seconds * len(times) / total, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle (PyTorch-CPU port of the reference) on B={Bs} x {args.seconds:g} s of the same workload, "
                      f"{len(times)} timed passes after {warmup} warm-up",
            "ms_per_step": 1e3 * total / len(times), "batch": Bs}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    warm = min(args.warmup, 1) if args.warmup > 0 else 0       # CPU path has no clocks/caches to warm beyond one pass
    r = cpu_oracle_run(args, args.steps, warm, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, {"cpu_sample_batch": r["batch"]}),
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", type=int, default=int(os.environ.get("BVC_PRECISION", "1")))
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU path (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    n_gpus = world

    from bernoulli_var_speech_codec_b200 import BVRNNCodecModel, SCALING
    from bernoulli_var_speech_codec_b200.sharding import gather_shards_async
    from bernoulli_var_speech_codec_b200.synth import write_synthetic_checkpoints
    ck_dir = os.environ.get("BVC_CKPT_DIR", "/tmp/bvc_ckpts")
    if rank == 0:
        ck = write_synthetic_checkpoints(ck_dir, seed=1, sharpen=30.0)
    if world > 1:
        dist.barrier()
    ck = write_synthetic_checkpoints(ck_dir, seed=1, sharpen=30.0)
    model = BVRNNCodecModel(os.path.join(ROOT, "configs", "config_varBitRate.toml"), *ck, device=dev).eval()
    eng = model._engine
    eng.set_precision(args.precision)

    B, L = args.batch, int(args.seconds * FS)
    T = L // HOP
    x_host = synth_batch(B, L, 1234 + rank).pin_memory()
    x = x_host.to(dev)
    bits = model.bits_per_frame(3000)

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    rec_ms = {"encode": [], "decode_mel": []}   # device time of the persistent recurrent kernel per launch (CUDA events)

    def step_device(events=None):
        def mark(name):
            if events is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                events.append((name, e))
        mark("start")
        mel = eng.logmel(x, SCALING)
        mark("logmel")
        codes, _, _, _, _ = eng.encode(mel, None, bits, None, want_all_h=False)
        mark("encode")
        if events is not None:
            rec_ms["encode"].append(eng.last_recurrent_ms())
        # the path's only collective is the final gather of codes and audio: the codes travel while the decoder runs, the
        # first half of the audio while the vocoder works on the second half (NCCL over NVLink, async, same process group)
        g_codes = gather_shards_async(codes, B * world) if world > 1 else None
        dmel, _ = eng.decode_mel(codes, None)
        mark("decode_mel")
        if events is not None:
            rec_ms["decode_mel"].append(eng.last_recurrent_ms())
        if world > 1:
            half = B // 2
            wav_a = eng.vocode(dmel[:half].contiguous(), L, SCALING)
            g_a = gather_shards_async(wav_a, half * world)
            wav_b = eng.vocode(dmel[half:].contiguous(), L, SCALING)
            g_b = gather_shards_async(wav_b, (B - half) * world)
            mark("vocode")
            all_codes, all_a, all_b = g_codes.result(), g_a.result(), g_b.result()
            mark("gather")
            return all_codes, (all_a, all_b)
        wav = eng.vocode(dmel, L, SCALING)
        mark("vocode")
        return codes, wav

    def step_e2e():
        codes = model.encode(x_host, 3000)          # H2D + logmel + encode + D2H inside the C ABI call
        wav = model.decode(codes, L)                # H2D + decode + vocoder + D2H
        return codes, wav

    # ---- device-resident timing ----
    for _ in range(args.warmup):
        step_device()
    sync()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = eng.kernel_launches()
    stage_events = []
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for _ in range(args.steps):
        ev = []
        step_device(ev)
        stage_events.append(ev)
    t_end.record()
    sync()
    clocks = sampler.finish()
    launches = eng.kernel_launches() - launches0
    elapsed_ms = torch.tensor([t_start.elapsed_time(t_end)], device=dev)
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
    elapsed_s = float(elapsed_ms.item()) / 1e3
    audio_s = n_gpus * B * args.seconds * args.steps
    value = audio_s / elapsed_s

    stage_ms = {}
    for ev in stage_events:
        for (n0, e0), (n1, e1) in zip(ev, ev[1:]):
            stage_ms[n1] = stage_ms.get(n1, 0.0) + e0.elapsed_time(e1) / args.steps

    # ---- end-to-end timing through the facade with pinned host buffers ----
    # warm-up with the same ownership pattern as the timed loop (the previous step's results are still referenced
    # while the next step runs), so that the facade's pool of page-locked result buffers is in steady state
    for _ in range(max(2, min(args.warmup, 3))):
        c_h, w_h = step_e2e()
    sync()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    t0.record()
    for _ in range(args.steps):
        c_h, w_h = step_e2e()
    t1.record()
    sync()
    e2e_s = torch.tensor([max(time.perf_counter() - wall0, t0.elapsed_time(t1) / 1e3)], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = audio_s / float(e2e_s.item())
    codes_bytes = B * T * Z * 4
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * L * 4 + codes_bytes,
           "d2h_bytes_per_step": codes_bytes + B * L * 4,
           "api": "BVRNNCodecModel.encode(x_cpu_pinned, 3000) -> codes_cpu; .decode(codes_cpu, L) -> wav_cpu"}

    # ---- fused forward (a separate figure with its own flop count, never mixed into `value`) ----
    # model(x, bitrate) = decode(encode(x)) with ONE recurrence: the encoder's internal decoder output feeds the vocoder
    for _ in range(2):
        w_f = model(x, 3000)
    sync()
    f0 = torch.cuda.Event(enable_timing=True)
    f1 = torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        w_f = model(x, 3000)
    f1.record()
    sync()
    fwd_ms = torch.tensor([f0.elapsed_time(f1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(fwd_ms, op=dist.ReduceOp.MAX)
    fwd_ms = float(fwd_ms.item())
    del w_f

    if rank == 0:
        peak_tf, hbm_gbs, peak_src = peaks()
        frames = B * T
        flops_step = GFLOP_PER_AUDIO_S * 1e9 * B * args.seconds
        step_ms = 1e3 * elapsed_s / args.steps
        achieved = n_gpus * flops_step / (step_ms * 1e-3) / 1e12
        stages = []
        for name in ("logmel", "encode", "decode_mel", "vocode", "gather"):
            if name in stage_ms:
                fl = MFLOP_FRAME.get(name, 0.0) * 1e6 * frames
                tf = fl / (stage_ms[name] * 1e-3) / 1e12 if stage_ms[name] > 0 else 0.0
                stages.append({"stage": name, "ms": round(stage_ms[name], 3), "tflops": round(tf, 2),
                               "frac": round(tf / peak_tf, 5)})
        # dominant kernel: recurrent_cluster_kernel of the encode call, timed inside the library with CUDA events on
        # the launching stream; algorithmic FLOPs = state-dependent MACs per frame x frames of the launch
        k_ms = sum(rec_ms["encode"]) / max(1, len(rec_ms["encode"]))
        k_flop = MFLOP_FRAME_RECURRENT["encode"] * 1e6 * frames
        k_tf = k_flop / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
        d_ms = sum(rec_ms["decode_mel"]) / max(1, len(rec_ms["decode_mel"]))
        d_tf = MFLOP_FRAME_RECURRENT["decode_mel"] * 1e6 * frames / (d_ms * 1e-3) / 1e12 if d_ms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r01_recurrent_ncu.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch_encode")
            except Exception:
                traffic = None
        roofline = {"bound": "tensor", "kernel": "recurrent_cluster_kernel (persistent BVRNN.encode time loop, one launch per step)",
                    "achieved": round(k_tf, 2), "peak": peak_tf, "unit": "TFLOP/s", "frac": round(k_tf / peak_tf, 5),
                    "traffic": traffic, "peak_source": peak_src,
                    "ms_per_launch": round(k_ms, 3), "algorithmic_flop_per_launch": k_flop,
                    "share_of_step": round((k_ms + d_ms) / step_ms, 3),
                    "decode_launch": {"ms_per_launch": round(d_ms, 3), "achieved": round(d_tf, 2), "frac": round(d_tf / peak_tf, 5)},
                    "step": {"achieved": round(achieved / n_gpus, 2), "frac": round(achieved / n_gpus / peak_tf, 5),
                             "gflop_per_audio_s": GFLOP_PER_AUDIO_S},
                    "stages": stages}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16x3" if args.precision == 1 else "f32", "data": "synthetic",
            "config": workload_config(args, {"precision_mode": args.precision}),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        }
        fwd_mflop_frame = MFLOP_FRAME["logmel"] + MFLOP_FRAME["encode"] + MFLOP_FRAME["vocode"]
        line["forward_fused"] = {
            "value": n_gpus * B * args.seconds / (fwd_ms * 1e-3), "unit": UNIT, "ms_per_step": round(fwd_ms, 3),
            "mflop_per_frame": round(fwd_mflop_frame, 3),
            "api": "BVRNNCodecModel.forward(x_cuda, 3000): one recurrence (bvc_encode_mel), device-resident; not comparable "
                   "with `value`, which runs encode and decode separately (121.65 MFLOP/frame)"}
        if n_gpus == 1 and not args.no_cpu_baseline:
            r = cpu_oracle_run(args, steps=1, warmup=0, budget_s=25.0)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

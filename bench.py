#!/usr/bin/env python
"""Headline benchmark: audio-seconds coded per second (encode + decode, 22.05 kHz, 3 kbps).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 1|2|3|4] [--precision 0|1]

Workloads (BASELINE.json `configs`, 0-based index = --config):
  1 (default, the headline)  var-bitrate codec, batch 256 x 10 s synthetic 22.05 kHz utterances per GPU at 3 kbps
                             (35 bits/frame); weak scaling: every rank codes its own 256-utterance shard
  2  fixed 64-bit/frame coder (config_64bit.toml), batch 256 x 10 s per GPU, bit-exact code check on sampled rows
  3  bitrate sweep {0, 1, 8, 16, 24, 35, 48, 64 bits/frame}, batch 1024 x 10 s in TOTAL, utterance-sharded over the N
     ranks (strong scaling); one step = the whole sweep; per-bitrate parity against the CPU oracle on a sampled row
  4  streaming: 512 concurrent real-time streams per GPU fed hop by hop (256 samples = 11.61 ms); one step = one hop of
     every stream through the stateful encoder and decoder; reports hop latency p50 / p99 and the real-time margin
Synthetic random-init weights with the reference's checkpoint schema (the shipped checkpoints are git-LFS pointers).
One step of configs 1-3 = encode(x) followed by decode(codes, L) on one batch (two recurrences).

  value     device-resident: inputs already in HBM, CUDA-event timed, max over ranks
  e2e       through the public API with pinned HOST tensors, every step's copies inside the timed region:
            pipeline.HostPipeline (model.encode + model.decode per batch; the H2D of batch k+1 and the D2H of codes and
            audio of batch k-1 run on the copy engines while batch k computes); `e2e.blocking` = the blocking facade calls
            model.encode(x_cpu) -> codes_cpu -> model.decode(codes_cpu) with every copy exposed
  roofline  dominant kernel (recurrent_cluster_kernel, encode launch): algorithmic FLOPs / device time (CUDA events
            recorded inside the library around the launch) against the measured sustained bf16 peak; per-stage breakdown
  cpu_baseline / --impl reference   the reference's CPU path on the box's host cores, bounded sample, rank 0 only: the
            UNMODIFIED reference modules from baseline/_ref (copied there by __graft_entry__.build()) when present
            (kind "reference"), else the oracle port (kind "port")
  torch_eager   the same arithmetic through PyTorch eager on the B200 (cuBLAS / cuDNN / cuFFT; the oracle restatement
            moved to CUDA), TF32 off and on: the "bar on B200" of SURVEY.md F1 / 8d (N = 1 only)

N > 1: launched by torchrun, one rank per GPU; the only collective is the final all-gather of the PACKED codes (one
uint64 per frame) and of the audio, inside the timed region and overlapped with compute; outside the timed region the
gathered tensors are checked against every rank's local shard ("gather_check").
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
SWEEP_BITS = [0, 1, 8, 16, 24, 35, 48, 64]
WORKLOADS = {
    1: "configs[1]: var-bitrate codec, batch 256 x 10 s synthetic 22.05 kHz utterances at 3 kbps",
    2: "configs[2]: fixed 64-bit/frame coder (config_64bit.toml), batch 256 x 10 s, bit-exact code check",
    3: "configs[3]: bitrate sweep over 0..64 bits/frame, batch 1024 x 10 s, utterance-sharded across the GPUs",
    4: "configs[4]: streaming frame-by-frame causal mode, 512 concurrent real-time streams per GPU, 34.8 ms algorithmic latency",
}


def workload_config(args, extra=None):
    cfg = {
        "workload": WORKLOADS[args.config], "config_index": args.config,
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


def checkpoints():
    from bernoulli_var_speech_codec_b200.synth import write_synthetic_checkpoints
    return write_synthetic_checkpoints(os.environ.get("BVC_CKPT_DIR", "/tmp/bvc_ckpts"), seed=1, sharpen=30.0)


def config_path(args):
    return os.path.join(ROOT, "configs", "config_64bit.toml" if args.config == 2 else "config_varBitRate.toml")


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


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own CPU implementation of the path on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(args, steps, warmup, budget_s):
    """Times the reference's CPU path on a bounded sample of the workload; returns dict.
    kind "reference": the unmodified reference modules (baseline/_ref or /root/reference through oracle/ref_shim.py);
    kind "port": the oracle restatement, when the reference modules are not available."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ck = checkpoints()
    cfg = config_path(args)
    codec, kind = None, "port"
    try:
        from oracle import ref_shim
        if ref_shim.available():
            ref = ref_shim.import_reference()
            codec, kind = ref.BVRNNCodecModel(cfg, *ck).eval(), "reference"
    except Exception as e:   # fall back to the port, say why
        print("reference modules unavailable, timing the oracle port:", repr(e), file=sys.stderr)
    if codec is None:
        from oracle.codec_oracle import OracleCodec
        codec = OracleCodec(cfg, *ck)
    L = int(args.seconds * FS)
    with torch.no_grad():
        # probe one short utterance to size the sample so that (warmup + steps) passes fit the budget
        xp = synth_batch(1, FS, 99)
        t0 = time.perf_counter()
        codec.decode(codec.encode(xp, 3000), FS)
        per_utt_s = (time.perf_counter() - t0) * args.seconds      # ~linear in duration at B=1
        Bs = int(max(1, min(8, budget_s / max(1e-3, per_utt_s * 0.6 * (steps + warmup)))))
        x = synth_batch(Bs, L, 1234)
        times = []
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            codes = codec.encode(x, 3000)
            wav = codec.decode(codes, L)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    assert wav.shape == x.shape
    total = sum(times)
    what = "unmodified reference modules (PyTorch CPU)" if kind == "reference" else "oracle (PyTorch-CPU port of the reference)"
    return {"value": Bs * args.seconds * len(times) / total, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{what} on B={Bs} x {args.seconds:g} s of the same workload, "
                      f"{len(times)} timed passes after {warmup} warm-up",
            "ms_per_step": 1e3 * total / len(times), "batch": Bs}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    warm = min(args.warmup, 1) if args.warmup > 0 else 0       # CPU path has no clocks/caches to warm beyond one pass
    r = cpu_reference_run(args, args.steps, warm, budget_s=150.0)
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


def torch_eager_run(args, dev):
    """The reference's arithmetic through PyTorch eager on this GPU (the oracle restatement with its tensors on CUDA:
    cuBLAS sgemm per Linear, cuDNN per convolution, cuFFT, ATen elementwise), same batch as the timed workload."""
    from oracle.codec_oracle import OracleCodec
    o = OracleCodec(config_path(args), *checkpoints(), device=dev)
    B, L = min(args.batch, 256), int(args.seconds * FS)
    x = synth_batch(B, L, 1234).to(dev)
    out = {"sample": f"OracleCodec(device=cuda) encode + decode, B={B} x {args.seconds:g} s, 1 warm-up + 1 timed pass, "
                     "vocoder in slices of 32 utterances"}

    def one_pass():
        codes = o.encode(x, 3000)
        wav = torch.cat([o.decode(codes[i:i + 32], L) for i in range(0, B, 32)], 0)
        return wav

    for name, tf32 in (("tf32_off", False), ("tf32_on", True)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        one_pass()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        wav = one_pass()
        e1.record()
        torch.cuda.synchronize(dev)
        wall = time.perf_counter() - t0
        assert wav.shape == (B, L)
        out[name] = {"value": B * args.seconds / wall, "unit": UNIT, "ms_per_step": round(1e3 * wall, 1),
                     "device_ms": round(e0.elapsed_time(e1), 1)}
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    del o
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=[1, 2, 3, 4])
    ap.add_argument("--precision", type=int, default=int(os.environ.get("BVC_PRECISION", "1")))
    ap.add_argument("--batch", type=int, default=None, help="utterances (streams) per GPU; default from --config")
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.batch is None:
        args.batch = {1: 256, 2: 256, 3: max(1, 1024 // max(1, world if args.impl == "b200" else 1)), 4: 512}[args.config]
    if args.impl == "reference":
        run_reference(args)
        return

    import torch.distributed as dist
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
    if rank == 0:
        checkpoints()
    if world > 1:
        dist.barrier()
    ck = checkpoints()
    model = BVRNNCodecModel(config_path(args), *ck, device=dev).eval()
    eng = model._engine
    eng.set_precision(args.precision)

    if args.config == 4:
        from bernoulli_var_speech_codec_b200.streaming import bench_streams
        line = bench_streams(model, args, n_gpus, rank, dev, ClockSampler(local_rank), METRIC, UNIT, workload_config)
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    B, L = args.batch, int(args.seconds * FS)
    T = L // HOP
    x_host = synth_batch(B, L, 1234 + rank).pin_memory()
    x = x_host.to(dev)
    sweep = SWEEP_BITS if args.config == 3 else [35]
    bitrates = [b * FS / HOP for b in sweep]          # encode() rounds bitrate * hop / fs back to these budgets

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def step_device(events=None, keep=None):
        def mark(name):
            if events is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                events.append((name, e))
        mark("start")
        out = None
        for bits in sweep:
            mel = eng.logmel(x, SCALING)
            mark("logmel")
            codes, _, _, _, packed = eng.encode(mel, None, float(bits), None, want_all_h=False, want_packed=True)
            mark("encode")
            # the path's only collective is the final gather of codes and audio: the PACKED codes (8 bytes per frame
            # instead of 256) travel while the decoder runs, the first half of the audio while the vocoder works on the
            # second half (NCCL over NVLink, async, same process group)
            g_codes = gather_shards_async(packed, B * world) if world > 1 else None
            dmel, _ = eng.decode_mel(codes, None)
            mark("decode_mel")
            if world > 1:
                half = B // 2
                wav_a = eng.vocode(dmel[:half].contiguous(), L, SCALING)
                g_a = gather_shards_async(wav_a, half * world)
                wav_b = eng.vocode(dmel[half:].contiguous(), L, SCALING)
                g_b = gather_shards_async(wav_b, (B - half) * world)
                mark("vocode")
                out = (g_codes.result(), g_a.result(), g_b.result(), packed, wav_a, wav_b)
                mark("gather")
            else:
                wav = eng.vocode(dmel, L, SCALING)
                mark("vocode")
                out = (packed, wav)
            if keep is not None:
                keep.append((bits, codes))
        return out

    def step_e2e():
        outs = None
        for br in bitrates:
            codes = model.encode(x_host, br)            # H2D + logmel + encode + D2H inside the C ABI call
            wav = model.decode(codes, L)                # H2D + decode + vocoder + D2H
            outs = (codes, wav)
        return outs

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
    # device time of the persistent recurrent kernel of the launches still in the library's ring (the last timed steps):
    # CUDA events recorded inside the library around the launch, read after the timed region (nothing blocks inside it)
    rec_ms = {"encode": [], "decode_mel": []}
    for kind, name in ((0, "encode"), (1, "decode_mel")):
        for age in range(min(args.steps * len(sweep), 4)):
            ms = eng.recurrent_ms(kind, age)
            if ms > 0:
                rec_ms[name].append(ms)
    elapsed_ms = torch.tensor([t_start.elapsed_time(t_end)], device=dev)
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
    elapsed_s = float(elapsed_ms.item()) / 1e3
    audio_s = n_gpus * B * args.seconds * args.steps * len(sweep)
    value = audio_s / elapsed_s

    stage_ms = {}
    for ev in stage_events:
        for (n0, e0), (n1, e1) in zip(ev, ev[1:]):
            stage_ms[n1] = stage_ms.get(n1, 0.0) + e0.elapsed_time(e1) / args.steps

    # ---- correctness of what was timed (outside the timed region) ----
    checks = {}
    keep = []
    res = step_device(keep=keep)
    torch.cuda.synchronize(dev)
    if world > 1:
        all_packed, all_a, all_b, packed, wav_a, wav_b = res
        half = B // 2
        ok = bool(torch.equal(all_packed[rank * B:(rank + 1) * B], packed)
                  and torch.equal(all_a[rank * half:(rank + 1) * half], wav_a)
                  and torch.equal(all_b[rank * (B - half):(rank + 1) * (B - half)], wav_b)
                  and all_packed.shape[0] == B * world and all_a.shape[0] + all_b.shape[0] == B * world)
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        checks["gather_check"] = "ok" if int(flag.item()) == 1 else "MISMATCH"
    if args.config in (2, 3) and rank == 0:
        # per-bitrate parity against the CPU oracle: one utterance, first ~2 s (the recurrence is causal, so the codes of
        # a prefix only depend on the prefix; the last frames of the prefix see its reflect padding and are excluded)
        from oracle.codec_oracle import OracleCodec
        oracle = OracleCodec(config_path(args), *ck)
        Lp = min(L, 2 * FS)
        n_cmp = max(1, Lp // HOP - 4)
        rows = sorted({0, B - 1})
        par = []
        for bits, codes in keep:
            row_ok, mism, masks_ok = True, 0, True
            for r in rows:
                oc = oracle.encode(x_host[r:r + 1, :Lp], bits * FS / HOP)[0, :n_cmp]
                gc = codes[r, :n_cmp].cpu()
                masks_ok = masks_ok and bool(((gc == 0.5) == (oc == 0.5)).all())
                mism += int((gc != oc).sum())
            nb = 64 if args.config == 2 else int(bits)
            c = codes[:, :, :]
            layout_ok = bool((c[:, :, nb:] == 0.5).all()) and bool(((c[:, :, :nb] == 0.0) | (c[:, :, :nb] == 1.0)).all())
            par.append({"bits_per_frame": nb, "mask_layout_ok": layout_ok, "mask_vs_oracle_ok": masks_ok,
                        "code_mismatches_vs_oracle": mism, "compared_bits": len(rows) * n_cmp * nb})
        checks["parity"] = par
    del keep, res

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
    e2e_blocking = {"value": e2e_value, "h2d_bytes_per_step": (B * L * 4 + codes_bytes) * len(sweep),
                    "d2h_bytes_per_step": (codes_bytes + B * L * 4) * len(sweep),
                    "api": "BVRNNCodecModel.encode(x_cpu_pinned, bitrate) -> codes_cpu; .decode(codes_cpu, L) -> wav_cpu (blocking calls)"}
    # ---- the same through the host pipeline: every step still copies its audio in and its codes + audio out (pinned host
    # buffers), but the copies of batch k + 1 / k - 1 run on the copy engines while batch k computes (pipeline.py) ----
    from bernoulli_var_speech_codec_b200.pipeline import HostPipeline
    pipe = HostPipeline(model, depth=2)
    n_sub = args.steps * len(bitrates)

    def run_pipeline(n):
        tickets, last = [], None
        for k in range(n):
            tickets.append(pipe.submit(x_host, bitrates[k % len(bitrates)]))
            if k >= 1:
                last = pipe.result(tickets[k - 1])
        last = pipe.result(tickets[n - 1])
        return last

    run_pipeline(max(2, min(args.warmup, 3)))
    sync()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    t0.record()
    pc_h, pw_h = run_pipeline(n_sub)
    t1.record()
    sync()
    pipe_s = torch.tensor([max(time.perf_counter() - wall0, t0.elapsed_time(t1) / 1e3)], device=dev)
    if world > 1:
        dist.all_reduce(pipe_s, op=dist.ReduceOp.MAX)
    pipe.close()
    pipe_ok = bool(torch.equal(pc_h, c_h) and torch.equal(pw_h, w_h))        # same outputs as the blocking calls
    e2e = {"value": audio_s / float(pipe_s.item()), "unit": UNIT, "h2d_bytes_per_step": B * L * 4 * len(sweep),
           "d2h_bytes_per_step": (codes_bytes + B * L * 4) * len(sweep),
           "api": "pipeline.HostPipeline(model).submit(x_cpu_pinned, bitrate) / .result() -> (codes_cpu, wav_cpu): "
                  "BVRNNCodecModel.encode + .decode per batch, H2D of batch k+1 and D2H of batch k-1 overlapped with batch k",
           "matches_blocking_calls": pipe_ok, "blocking": e2e_blocking}

    # ---- fused forward (a separate figure with its own flop count, never mixed into `value`) ----
    # model(x, bitrate) = decode(encode(x)) with ONE recurrence: the encoder's internal decoder output feeds the vocoder
    fwd_ms = None
    if args.config == 1:
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
        flops_step = GFLOP_PER_AUDIO_S * 1e9 * B * args.seconds * len(sweep)
        step_ms = 1e3 * elapsed_s / args.steps
        achieved = n_gpus * flops_step / (step_ms * 1e-3) / 1e12
        stages = []
        for name in ("logmel", "encode", "decode_mel", "vocode", "gather"):
            if name in stage_ms:
                fl = MFLOP_FRAME.get(name, 0.0) * 1e6 * frames * len(sweep)
                tf = fl / (stage_ms[name] * 1e-3) / 1e12 if stage_ms[name] > 0 else 0.0
                stages.append({"stage": name, "ms": round(stage_ms[name], 3), "tflops": round(tf, 2),
                               "frac": round(tf / peak_tf, 5)})
        # dominant kernel: recurrent_cluster_kernel of the encode call; algorithmic FLOPs = state-dependent MACs per frame
        # x frames of the launch
        k_ms = sum(rec_ms["encode"]) / max(1, len(rec_ms["encode"]))
        k_flop = MFLOP_FRAME_RECURRENT["encode"] * 1e6 * frames
        k_tf = k_flop / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
        d_ms = sum(rec_ms["decode_mel"]) / max(1, len(rec_ms["decode_mel"]))
        d_tf = MFLOP_FRAME_RECURRENT["decode_mel"] * 1e6 * frames / (d_ms * 1e-3) / 1e12 if d_ms > 0 else 0.0
        traffic, traffic_src = None, None
        for cand in ("r02_recurrent_ncu.json", "r01_recurrent_ncu.json"):
            tp = os.path.join(ROOT, "profiles", cand)
            if os.path.exists(tp):
                try:
                    traffic = json.load(open(tp)).get("dram_bytes_per_launch_encode")
                    traffic_src = "profiles/" + cand + " (ncu --set full of this kernel at B=256 x 10 s)"
                    break
                except Exception:
                    traffic = None
        roofline = {"bound": "tensor", "kernel": "recurrent_cluster_kernel (persistent BVRNN.encode time loop, one launch per step)",
                    "achieved": round(k_tf, 2), "peak": peak_tf, "unit": "TFLOP/s", "frac": round(k_tf / peak_tf, 5),
                    "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "ms_per_launch": round(k_ms, 3), "algorithmic_flop_per_launch": k_flop,
                    "share_of_step": round((k_ms + d_ms) * len(sweep) / step_ms, 3),
                    "decode_launch": {"ms_per_launch": round(d_ms, 3), "achieved": round(d_tf, 2), "frac": round(d_tf / peak_tf, 5)},
                    "step": {"achieved": round(achieved / n_gpus, 2), "frac": round(achieved / n_gpus / peak_tf, 5),
                             "gflop_per_audio_s": GFLOP_PER_AUDIO_S},
                    "stages": stages}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "strong" if args.config == 3 else "weak",
            "vs_baseline": None, "dtype": "bf16x3" if args.precision == 1 else "f32", "data": "synthetic",
            "config": workload_config(args, {"precision_mode": args.precision,
                                             "bits_per_frame_sweep": sweep if args.config == 3 else None}),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        }
        line.update(checks)
        if fwd_ms is not None:
            fwd_mflop_frame = MFLOP_FRAME["logmel"] + MFLOP_FRAME["encode"] + MFLOP_FRAME["vocode"]
            line["forward_fused"] = {
                "value": n_gpus * B * args.seconds / (fwd_ms * 1e-3), "unit": UNIT, "ms_per_step": round(fwd_ms, 3),
                "mflop_per_frame": round(fwd_mflop_frame, 3),
                "api": "BVRNNCodecModel.forward(x_cuda, 3000): one recurrence (bvc_encode_mel), device-resident; not comparable "
                       "with `value`, which runs encode and decode separately (121.65 MFLOP/frame)"}
        if n_gpus == 1 and not args.no_eager:
            try:
                line["torch_eager"] = torch_eager_run(args, dev)
            except Exception as e:       # an extra key must never cost the headline line
                line["torch_eager"] = {"unavailable": repr(e)[:200]}
        if n_gpus == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(args, steps=1, warmup=0, budget_s=25.0)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Streaming use of the codec with carried state (BASELINE configs[4]; not in the reference).

Two layers:

  StreamSession                      the real-time path: S independent streams advance hop by hop (256 samples = 11.61 ms)
                                     through the library's stateful step kernels (include/bvc.h, bvc_stream_*): per-stream
                                     window FIFO, BVRNN encoder / decoder states and the vocoder's stage-input rings live on
                                     the device; a hop costs two MMA tiles per vocoder stage and stream, nothing is
                                     re-synthesised.  Streams may be idle in any hop (masks).  Edge policy: zero-state start of
                                     the vocoder (the first 27 frames differ from the offline decode), no right reflect padding.
  StreamingEncoder / StreamingDecoder  chunks of ANY length through the offline kernels with carried BVRNN state; bit-identical
                                     to the offline calls from the first sample to the last (the decoder re-synthesises 28
                                     frames of history per chunk, so it is the reference check for StreamSession, not the fast path).


The reference only has the ingredients of a real-time mode: the BVRNN takes and returns its hidden state
(bvrnn.py:163,211) and every vocoder convolution is causal (models.py:19-20,110,117).  This module feeds the same CUDA
kernels chunk by chunk and carries the state between calls, so that the concatenated outputs are IDENTICAL to the
offline `BVRNNCodecModel.encode` / `.decode` of the whole signal:

  encoder   log-mel frame t covers samples [256 t - 256, 256 t + 768): 768 samples = 34.8 ms of look-ahead, the codec's
            algorithmic latency.  A frame is emitted as soon as its window is complete; frame 0 uses the reference's
            reflect padding on the left, `flush()` applies the reflect padding on the right, exactly like the offline
            call (meldataset.py:76-80).  The BVRNN encoder state h is carried (bvc_encode's h0 / h_final).
  decoder   the BVRNN decoder state h is carried; the vocoder is causal with a receptive field of 6 719 samples
            (26.25 frames), so each chunk is synthesised from the last 28 decoded-mel frames of history plus the new
            frames and only the samples of the new frames are kept.

All streams of a batch advance together (same chunk lengths).  Tensors live on the model's device.
"""
from __future__ import annotations

import ctypes as C
import time

import torch

from . import _lib
from .codec import SCALING, _ptr, _stream

HOP = 256
TAIL = 768                   # samples of a frame's window to the right of its first hop: n_fft - pad_left
HISTORY_FRAMES = 28          # >= ceil(6719 / 256): the vocoder's receptive field in mel frames


def _check_streamable(model):
    """The chunked paths rely on hop 256 / window 1024 / pad_left 256 and on a strictly causal vocoder."""
    conf = model.conf
    if conf["hopsize"] != HOP or conf["winsize"] - conf["mel_pad_left"] != TAIL:
        raise NotImplementedError("streaming: hopsize / winsize / mel_pad_left differ from the shipped 256 / 1024 / 256")
    v = conf["vocoder_config"]
    if any(v.get("layers_antialias", [])) or v.get("antialias_post", False):
        raise NotImplementedError(
            "streaming: the anti-aliased Activation1d (symmetric 12-tap FIRs with replicate padding) looks ahead, so "
            "chunked decoding is not identical to the offline call; use the offline decode for such configs")


class StreamingEncoder:
    def __init__(self, model, bitrate):
        _check_streamable(model)
        self.m, self.eng = model, model._engine
        self.bits = model.bits_per_frame(bitrate)
        self.buf = None          # received samples from global index self.start on
        self.start = 0
        self.n_total = 0         # samples received
        self.t_done = 0          # frames emitted
        self.h = None            # BVRNN encoder state [B, H]

    def _encode_frames(self, t1, seg_end):
        """Codes of the frames [t_done, t1), computed from the samples [a, seg_end) with a one frame before t_done:
        the library's log-mel reflects at both ends of what it is given; on the left that only touches the extra frame
        (dropped), unless t_done = 0, where the reflection IS the reference's padding."""
        t0 = self.t_done
        a = HOP * max(t0 - 1, 0)
        j0 = t0 - a // HOP
        seg = self.buf[:, a - self.start: seg_end - self.start].contiguous()
        mel = self.eng.logmel(seg, SCALING)[:, j0: j0 + (t1 - t0)].contiguous()
        assert mel.shape[1] == t1 - t0
        codes, _, self.h, _, _ = self.eng.encode(mel, None, self.bits, self.h, want_all_h=False)
        self.t_done = t1
        keep_from = HOP * max(t1 - 1, 0)               # the next segment starts one frame before frame t1
        if keep_from > self.start:
            self.buf = self.buf[:, keep_from - self.start:].contiguous()
            self.start = keep_from
        return codes

    def push(self, x):
        """x: (B, n) new samples -> codes (B, k, 64) of the frames whose windows they complete, or None if k = 0."""
        x = x.to(self.m.device, torch.float32)
        self.buf = x if self.buf is None else torch.cat([self.buf, x], 1)
        self.n_total += x.shape[1]
        # frame t is complete once sample 256 t + 768 - 1 has arrived
        t1 = (self.n_total - TAIL) // HOP + 1 if self.n_total >= TAIL else 0
        if t1 <= self.t_done:
            return None
        # cut the segment exactly at the end of frame t1-1's window: that frame is then the last one whose window lies
        # inside the segment, so the library's right reflection touches none of the frames that are kept
        return self._encode_frames(t1, HOP * (t1 - 1) + TAIL)

    def flush(self):
        """End of the signal: the remaining frames, with the reference's right reflect padding (n // 256 frames in total)."""
        t_end = self.n_total // HOP
        if t_end <= self.t_done:
            return None
        return self._encode_frames(t_end, self.n_total)


class StreamingDecoder:
    def __init__(self, model):
        _check_streamable(model)
        self.m, self.eng = model, model._engine
        self.h = None            # BVRNN decoder state [B, H]
        self.mel_hist = None     # last HISTORY_FRAMES decoded mel frames [B, <=28, X]

    def push(self, codes):
        """codes: (B, k, 64) -> the 256 k waveform samples of these frames, identical to the offline decode."""
        codes = codes.to(self.m.device, torch.float32).contiguous()
        k = codes.shape[1]
        mel, self.h = self.eng.decode_mel(codes, self.h)
        ctx = mel if self.mel_hist is None else torch.cat([self.mel_hist, mel], 1).contiguous()
        n_hist = ctx.shape[1] - k
        wav = self.eng.vocode(ctx, HOP * ctx.shape[1], SCALING)
        self.mel_hist = ctx[:, -HISTORY_FRAMES:].contiguous()
        return wav[:, HOP * n_hist:].contiguous()

    def flush(self, n_extra):
        """The n_extra (< 294) samples the offline decode returns beyond the last whole frame (transposed-conv tails)."""
        if n_extra <= 0 or self.mel_hist is None:
            return None
        n = HOP * self.mel_hist.shape[1]
        return self.eng.vocode(self.mel_hist, n + n_extra, SCALING)[:, n:].contiguous()


class StreamSession:
    """S concurrent real-time streams on one GPU (bvc_stream_create / _encode_step / _decode_step / _reset / _destroy)."""

    def __init__(self, model, n_streams, bitrate):
        _check_streamable(model)
        self.m, self.eng, self.S = model, model._engine, int(n_streams)
        self.bits = model.bits_per_frame(bitrate)
        self.dev = model.device
        self.handle = C.c_void_p()
        with torch.cuda.device(self.dev):
            _lib.check(self.eng.lib.bvc_stream_create(self.eng.handle, self.S, C.byref(self.handle)), "bvc_stream_create")

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.eng.lib.bvc_stream_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _mask(self, m):
        if m is None:
            return None
        return m.to(self.dev, torch.uint8).contiguous()

    def reset(self, which=None):
        """Start new streams in the slots marked by `which` (bool / uint8 [S]; None = all)."""
        w = self._mask(which)
        with torch.cuda.device(self.dev):
            _lib.check(self.eng.lib.bvc_stream_reset(self.handle, _ptr(w), _stream(self.dev)), "bvc_stream_reset")

    def encode_step(self, x_new, active=None, bits=None):
        """x_new (S, 256) new samples per stream -> (words int64 (S,), valid uint8 (S,)): the wire word of the frame each
        stream completed in this hop (valid = 0: none yet -- the first frame needs 768 samples -- or the stream is idle)."""
        x_new = x_new.to(self.dev, torch.float32).contiguous()
        assert tuple(x_new.shape) == (self.S, HOP)
        a = self._mask(active)
        b = bits.to(self.dev, torch.float32).contiguous() if bits is not None else None
        words = torch.empty(self.S, device=self.dev, dtype=torch.int64)
        valid = torch.empty(self.S, device=self.dev, dtype=torch.uint8)
        with torch.cuda.device(self.dev):
            _lib.check(self.eng.lib.bvc_stream_encode_step(self.handle, _ptr(x_new), _ptr(a), _ptr(b), float(self.bits), SCALING,
                                                           _ptr(words), _ptr(valid), _stream(self.dev)), "bvc_stream_encode_step")
        return words, valid

    def decode_step(self, words, valid=None, bits=None):
        """words (S,) + valid (S,) -> the 256 samples of this hop per stream (S, 256); zeros for idle streams."""
        words = words.to(self.dev, torch.int64).contiguous()
        v = self._mask(valid)
        b = bits.to(self.dev, torch.float32).contiguous() if bits is not None else None
        wav = torch.empty(self.S, HOP, device=self.dev, dtype=torch.float32)
        with torch.cuda.device(self.dev):
            _lib.check(self.eng.lib.bvc_stream_decode_step(self.handle, _ptr(words), _ptr(v), _ptr(b), float(self.bits), SCALING,
                                                           _ptr(wav), _stream(self.dev)), "bvc_stream_decode_step")
        return wav


def bench_streams(model, args, n_gpus, rank, dev, sampler, metric, unit, workload_config):
    """bench.py --config 4 (BASELINE configs[4]): args.batch real-time streams per GPU, one step = one 256-sample hop of
    every stream through encode_step + decode_step.  Returns the JSON line (rank 0) or None.
      value   device-resident: hops enqueued back to back, CUDA-event timed, max over ranks
      e2e     host audio in / out: per hop H2D of the new samples (pinned), both steps, D2H of the decoded hop, host sync --
              the per-hop latency distribution comes from this pass"""
    import numpy as np
    import torch.distributed as dist
    world = n_gpus
    S = args.batch
    hops = max(100, args.steps * 100)
    warm = max(30, args.warmup * 10)                       # past the 27-frame start-up of the vocoder rings
    sess = StreamSession(model, S, 3000)
    g = torch.Generator().manual_seed(4321 + rank)
    x_host = (0.1 * torch.randn(hops + warm, S, HOP, generator=g)).clamp_(-1, 1).pin_memory()
    x_dev = x_host.to(dev)

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for i in range(warm):
        w_, v_ = sess.encode_step(x_dev[i])
        sess.decode_step(w_, v_)
    sync()
    sampler.start()
    launches0 = model._engine.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(warm, warm + hops):
        w_, v_ = sess.encode_step(x_dev[i])
        wav = sess.decode_step(w_, v_)
    e1.record()
    sync()
    clocks = sampler.finish()
    launches = model._engine.kernel_launches() - launches0
    el = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    elapsed_s = float(el.item()) / 1e3
    audio_s = world * S * hops * HOP / 22050.0
    # host in / out, one hop at a time
    sess.reset()
    out_host = torch.empty(S, HOP).pin_memory()
    lat = []
    for i in range(warm + hops):
        t0 = time.perf_counter()
        xd = x_host[i].to(dev, non_blocking=True)
        w_, v_ = sess.encode_step(xd)
        wav = sess.decode_step(w_, v_)
        out_host.copy_(wav, non_blocking=True)
        torch.cuda.synchronize(dev)
        if i >= warm:
            lat.append(time.perf_counter() - t0)
    lat = np.asarray(lat)
    e2e_s = torch.tensor([float(lat.sum())], device=dev)
    p99 = torch.tensor([float(np.percentile(lat, 99))], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        dist.all_reduce(p99, op=dist.ReduceOp.MAX)
    sess.close()
    if rank != 0:
        return None
    budget_ms = 1e3 * HOP / 22050.0
    return {
        "metric": metric, "value": audio_s / elapsed_s, "unit": unit, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * elapsed_s / hops, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16x3", "data": "synthetic",
        "config": workload_config(args, {"streams_per_gpu": S, "streams_total": S * world, "hops_timed": hops,
                                         "step": "one 256-sample hop (11.61 ms) of every stream: encode_step + decode_step",
                                         "l2": "per-hop state (30 MB of rings + 8 MB of BVRNN state) and the 94 MB of weights exceed nothing: "
                                               "the hop is latency-bound, not bandwidth-bound"}),
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": audio_s / float(e2e_s.item()), "unit": unit, "h2d_bytes_per_step": S * HOP * 4, "d2h_bytes_per_step": S * HOP * 4,
                "api": "StreamSession.encode_step(x_hop_cpu_pinned -> cuda) + decode_step -> wav_hop_cpu, host sync per hop"},
        "streaming": {"hop_budget_ms": round(budget_ms, 3), "hop_ms_p50": round(1e3 * float(np.median(lat)), 3),
                      "hop_ms_p99_max_over_ranks": round(1e3 * float(p99.item()), 3),
                      "realtime_margin_p99": round(budget_ms / (1e3 * float(p99.item())), 2),
                      "device_ms_per_hop": round(1e3 * elapsed_s / hops, 3),
                      "algorithmic_latency_ms": 34.83,
                      "edge_policy": "zero-state start of the vocoder rings; no right reflect padding (a stream just ends)"},
    }

"""Chunked (streaming) use of the codec with carried state (not in the reference).

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

import torch

from .codec import SCALING

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

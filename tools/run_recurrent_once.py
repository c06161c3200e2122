"""Runs one BVRNN encode + decode_mel on synthetic input (for ncu captures of the recurrent kernels)."""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bernoulli_var_speech_codec_b200 import BVRNNCodecModel, SCALING
from bernoulli_var_speech_codec_b200.synth import write_synthetic_checkpoints

ap = argparse.ArgumentParser()
ap.add_argument("--precision", type=int, default=1)
ap.add_argument("--B", type=int, default=256)
ap.add_argument("--seconds", type=float, default=0.25)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
ck = write_synthetic_checkpoints(os.environ.get("BVC_CKPT_DIR", "/tmp/bvc_ckpts"), seed=1, sharpen=30.0)
m = BVRNNCodecModel(os.path.join(ROOT, "configs", "config_varBitRate.toml"), *ck).eval()
m._engine.set_precision(a.precision)
g = torch.Generator().manual_seed(1)
x = (0.1 * torch.randn(a.B, int(a.seconds * 22050), generator=g)).clamp(-1, 1).cuda()
for _ in range(a.reps):
    mel = m._engine.logmel(x, SCALING)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
    e0.record()
    codes = m._engine.encode(mel, None, 35.0, None, want_all_h=False)[0]
    e1.record()
    enc_ms = m._engine.last_recurrent_ms()
    dmel, _ = m._engine.decode_mel(codes, None)
    e2.record()
    dec_ms = m._engine.last_recurrent_ms()
torch.cuda.synchronize()
T = codes.shape[1]
print("ok", tuple(codes.shape), "encode %.2f ms (%.1f us/frame)  decode_mel %.2f ms (%.1f us/frame)  flags=%s" % (
    e0.elapsed_time(e1), 1e3 * e0.elapsed_time(e1) / T, e1.elapsed_time(e2), 1e3 * e1.elapsed_time(e2) / T,
    os.environ.get("BVC_REC_DEBUG", "0")), " recurrent kernels: encode %.1f us/frame, decode %.1f us/frame" % (1e3 * enc_ms / T, 1e3 * dec_ms / T))

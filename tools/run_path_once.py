"""Runs the whole encode -> decode path once on synthetic input (for ncu captures of every kernel on the path)."""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bernoulli_var_speech_codec_b200 import BVRNNCodecModel
from bernoulli_var_speech_codec_b200.synth import write_synthetic_checkpoints
ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=256)
ap.add_argument("--seconds", type=float, default=10.0)
ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()
ck = write_synthetic_checkpoints(os.environ.get("BVC_CKPT_DIR", "/tmp/bvc_ckpts"), seed=1, sharpen=30.0)
m = BVRNNCodecModel(os.path.join(ROOT, "configs", "config_varBitRate.toml"), *ck).eval()
g = torch.Generator().manual_seed(1)
L = int(a.seconds * 22050)
x = (0.1 * torch.randn(a.B, L, generator=g)).clamp(-1, 1).cuda()
for _ in range(a.reps):
    codes = m.encode(x, 3000)
    wav = m.decode(codes, L)
torch.cuda.synchronize()
print("ok", tuple(codes.shape), tuple(wav.shape))

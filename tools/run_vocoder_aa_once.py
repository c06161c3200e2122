"""Runs the vocoder with the anti-aliased activations on (configs/config_varBitRate_antialias.toml) on synthetic mel:
timing of the layer-by-layer path, and the launch list / ncu target for aa_act_kernel (HBM-bound: 8 B per element)."""
import argparse, os, sys
import toml, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bernoulli_var_speech_codec_b200 import BVRNNCodecModel, SCALING
from bernoulli_var_speech_codec_b200.synth import write_synthetic_checkpoints
ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=64)
ap.add_argument("--T", type=int, default=344)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
cfg = os.path.join(ROOT, "configs", "config_varBitRate_antialias.toml")
ck = write_synthetic_checkpoints(os.environ.get("BVC_CKPT_DIR", "/tmp/bvc_ckpts"), seed=1, sharpen=30.0,
                                 vcfg=toml.load(cfg)["vocoder_config"])
m = BVRNNCodecModel(cfg, *ck).eval()
g = torch.Generator().manual_seed(1)
mel = (-5.0 + 2.0 * torch.randn(a.B, a.T, 80, generator=g)).cuda()
for _ in range(a.reps):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    wav = m._engine.vocode(mel, a.T * 256, SCALING)
    e1.record()
torch.cuda.synchronize()
print("ok", tuple(wav.shape), "anti-aliased vocode %.2f ms (%.0f audio-s/s)" % (e0.elapsed_time(e1), a.B * a.T * 256 / 22050 / e0.elapsed_time(e1) * 1e3))

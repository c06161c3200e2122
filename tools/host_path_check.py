import os, sys, torch
sys.path.insert(0, "/root/repo")
from bernoulli_var_speech_codec_b200 import BVRNNCodecModel
from bernoulli_var_speech_codec_b200.synth import write_synthetic_checkpoints
ck = write_synthetic_checkpoints(os.environ.get("BVC_CKPT_DIR", "/tmp/bvc_ckpts"), seed=1, sharpen=30.0)
m = BVRNNCodecModel("/root/repo/configs/config_varBitRate.toml", *ck).eval()
g = torch.Generator().manual_seed(3)
for B, L in ((64, 220500), (17, 220500), (256, 220500)):
    x = (0.1 * torch.randn(B, L, generator=g)).clamp(-1, 1)
    codes = m.encode(x, 3000)                 # host path
    wav_h = m.decode(codes, L)                # host path (split in halves when large)
    wav_d = m.decode(codes.to(m.device), L).cpu()
    print(B, L, "host == device:", bool(torch.equal(wav_h, wav_d)), float((wav_h - wav_d).abs().max()))

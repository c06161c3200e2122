"""BASELINE config #4 probe on one GPU: throughput and mask correctness over the var-bitrate model's range of bit budgets
(cost is bitrate-independent: the budget only masks code positions), B x 10 s utterances, device-resident."""
import argparse, json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bernoulli_var_speech_codec_b200 import BVRNNCodecModel
from bernoulli_var_speech_codec_b200.synth import write_synthetic_checkpoints

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=256)
ap.add_argument("--seconds", type=float, default=10.0)
a = ap.parse_args()
ck = write_synthetic_checkpoints(os.environ.get("BVC_CKPT_DIR", "/tmp/bvc_ckpts"), seed=1, sharpen=30.0)
m = BVRNNCodecModel(os.path.join(ROOT, "configs", "config_varBitRate.toml"), *ck).eval()
L = int(a.seconds * 22050)
g = torch.Generator().manual_seed(1234)
x = (0.1 * torch.randn(a.B, L, generator=g)).clamp(-1, 1).cuda()
rows = []
for bits in (0, 1, 8, 16, 24, 35, 48, 64):
    bitrate = bits * 22050 / 256
    for rep in range(2):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        codes = m.encode(x, bitrate)
        wav = m.decode(codes, L)
        e1.record()
        torch.cuda.synchronize()
    c = codes[:8].cpu().numpy()
    ok = bool((c[:, :, bits:] == 0.5).all() and np.isin(c[:, :, :bits], (0.0, 1.0)).all() and torch.isfinite(wav).all())
    rows.append({"bits_per_frame": bits, "bitrate_bps": round(bitrate, 1), "ms_per_step": round(e0.elapsed_time(e1), 2),
                 "audio_s_per_s": round(a.B * a.seconds / e0.elapsed_time(e1) * 1e3, 1), "mask_ok": ok})
    print(json.dumps(rows[-1]), flush=True)
print(json.dumps({"workload": "configs[3]: bitrate sweep, B=%d x %.0f s, one B200, encode + decode device-resident" % (a.B, a.seconds),
                  "rows": rows}))

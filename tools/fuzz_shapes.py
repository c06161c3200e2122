"""Shape fuzz on the GPU box: random (B, L, bitrate) cases through the facade against the CPU oracle (codes with the
eps-band protocol of tests/parity.py, waveform SNR), both arithmetic modes.  Bring-up aid; prints one line per case."""
import argparse, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bernoulli_var_speech_codec_b200 import BVRNNCodecModel
from bernoulli_var_speech_codec_b200.synth import write_synthetic_checkpoints
from oracle.codec_oracle import OracleCodec
from parity import compare_codes, snr_db

ap = argparse.ArgumentParser()
ap.add_argument("--cases", type=int, default=10)
ap.add_argument("--seed", type=int, default=0)
a = ap.parse_args()
ck = write_synthetic_checkpoints(os.environ.get("BVC_CKPT_DIR", "/tmp/bvc_ckpts"), seed=1, sharpen=30.0)
cfg = os.path.join(ROOT, "configs", "config_varBitRate.toml")
m = BVRNNCodecModel(cfg, *ck).eval()
o = OracleCodec(cfg, *ck)
rng = np.random.default_rng(a.seed)
bad = 0
for i in range(a.cases):
    B = int(rng.choice([1, 2, 3, 5, 17, 129, 130]))
    L = int(rng.integers(600, 30000 if B < 17 else 4000))
    bitrate = float(rng.choice([0, 500, 1500, 3000, 5512.5, 8000]))
    g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
    x = (0.1 * torch.randn(B, L, generator=g)).clamp(-1, 1)
    for prec in (1, 0):
        m._engine.set_precision(prec)
        taps = {}
        oc = o.encode(x, bitrate, taps)
        r = m.encode_with_taps(x.to(m.device), bitrate)
        rep = compare_codes(m._engine, r["mel"], m.bits_per_frame(bitrate), oc, taps["logits"], taps["all_h"])
        ow = o.decode(oc, L)
        w = m.decode(oc.to(m.device), L).cpu()
        fw = m(x.to(m.device), bitrate).cpu()
        snr = snr_db(ow.numpy(), w.numpy())
        ok = rep["hard_mismatches"] == 0 and rep["mask_errors"] == 0 and w.shape == ow.shape and snr >= 60.0 and fw.shape == ow.shape
        bad += 0 if ok else 1
        print(f"case {i} B={B} L={L} bitrate={bitrate} precision={prec}: eps_bits={rep['eps_bits']} hard={rep['hard_mismatches']} "
              f"snr={snr:.1f} dB {'ok' if ok else 'FAIL'}", flush=True)
print("failures:", bad)
sys.exit(1 if bad else 0)

"""BASELINE config #5 probe: B concurrent real-time streams fed hop by hop (256 new samples = 11.61 ms per step).

Reports the wall time of one encode step and one decode step (p50 / p99 over the run) against the 11.61 ms budget.
"""
import argparse, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bernoulli_var_speech_codec_b200 import BVRNNCodecModel
from bernoulli_var_speech_codec_b200.streaming import StreamingDecoder, StreamingEncoder
from bernoulli_var_speech_codec_b200.synth import write_synthetic_checkpoints

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=512)
ap.add_argument("--hops", type=int, default=120)
ap.add_argument("--frames-per-step", type=int, default=1)
a = ap.parse_args()
ck = write_synthetic_checkpoints(os.environ.get("BVC_CKPT_DIR", "/tmp/bvc_ckpts"), seed=1, sharpen=30.0)
m = BVRNNCodecModel(os.path.join(ROOT, "configs", "config_varBitRate.toml"), *ck).eval()
n = 256 * a.frames_per_step
x = (0.1 * torch.randn(a.streams, n * (a.hops + 8))).clamp(-1, 1).cuda()
enc, dec = StreamingEncoder(m, 3000), StreamingDecoder(m)
te, td = [], []
for i in range(a.hops + 8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c = enc.push(x[:, i * n:(i + 1) * n])
    torch.cuda.synchronize(); t1 = time.perf_counter()
    if c is not None:
        w = dec.push(c)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        if i >= 8:
            te.append((t1 - t0) * 1e3); td.append((t2 - t1) * 1e3)
te, td = np.array(te), np.array(td)
budget = 1e3 * n / 22050
print("streams %d, %d frame(s) per step (budget %.2f ms): encode p50 %.2f p99 %.2f ms, decode p50 %.2f p99 %.2f ms, total p50 %.2f ms -> %.2fx real time" % (
    a.streams, a.frames_per_step, budget, np.median(te), np.percentile(te, 99), np.median(td), np.percentile(td, 99),
    np.median(te + td), budget / np.median(te + td)))

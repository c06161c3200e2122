"""Wall-clock breakdown of the host-buffer API (encode_host / decode_host) at the bench workload."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bernoulli_var_speech_codec_b200 import BVRNNCodecModel, SCALING
from bernoulli_var_speech_codec_b200.synth import write_synthetic_checkpoints
ck = write_synthetic_checkpoints(os.environ.get("BVC_CKPT_DIR", "/tmp/bvc_ckpts"), seed=1, sharpen=30.0)
m = BVRNNCodecModel(os.path.join(ROOT, "configs", "config_varBitRate.toml"), *ck).eval()
B, L = 256, 220500
x = (0.1 * torch.randn(B, L)).clamp(-1, 1).pin_memory()
xd = x.cuda()
def wall(f, n=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): r = f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, r
t, _ = wall(lambda: xd.copy_(x, non_blocking=True)); print("H2D 226 MB pinned: %.1f ms (%.1f GB/s)" % (t, 0.2258 / t * 1e3))
hb = torch.empty(B, L, pin_memory=True)
t, _ = wall(lambda: hb.copy_(xd, non_blocking=True)); print("D2H 226 MB pinned: %.1f ms (%.1f GB/s)" % (t, 0.2258 / t * 1e3))
t, _ = wall(lambda: torch.empty(B, L, pin_memory=True)); print("torch.empty pinned 226 MB: %.1f ms" % t)
t, codes_d = wall(lambda: m.encode(xd, 3000)); print("encode device: %.1f ms" % t)
t, codes_h = wall(lambda: m.encode(x, 3000)); print("encode host  : %.1f ms" % t)
t, wav_d = wall(lambda: m.decode(codes_d, L)); print("decode device: %.1f ms" % t)
t, wav_h = wall(lambda: m.decode(codes_h, L)); print("decode host  : %.1f ms" % t)

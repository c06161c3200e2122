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
# ---- the same work done by hand with torch copies, to locate hidden costs of the host-buffer entry points ----
wav_pin = torch.empty(B, L, pin_memory=True)
def manual_decode():
    cd = codes_h.cuda(non_blocking=True)
    wd = m.decode(cd, L)
    wav_pin.copy_(wd, non_blocking=True)
    torch.cuda.synchronize()
    return wav_pin
t, _ = wall(manual_decode); print("decode manual (H2D + device API + D2H): %.1f ms" % t)
for i in range(4):
    t0 = time.perf_counter(); w = m.decode(codes_h, L); t1 = time.perf_counter()
    print("  decode host call %d: %.1f ms" % (i, (t1 - t0) * 1e3))
    del w
for i in range(3):
    t0 = time.perf_counter(); w = m._engine.pinned.empty((B, L)); t1 = time.perf_counter()
    print("  pool.empty %d: %.2f ms" % (i, (t1 - t0) * 1e3)); del w

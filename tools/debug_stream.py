import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bernoulli_var_speech_codec_b200 import BVRNNCodecModel
from bernoulli_var_speech_codec_b200.streaming import StreamSession
from bernoulli_var_speech_codec_b200.synth import write_synthetic_checkpoints
ck = write_synthetic_checkpoints(os.environ.get("BVC_CKPT_DIR", "/tmp/bvc_ckpts"), seed=1, sharpen=30.0)
m = BVRNNCodecModel(os.path.join(ROOT, "configs", "config_varBitRate.toml"), *ck).eval()
S = 3
x = (0.1 * torch.randn(S, 256 * 8)).clamp(-1, 1).cuda()
sess = StreamSession(m, S, 3000)
for k in range(6):
    w, v = sess.encode_step(x[:, 256 * k:256 * (k + 1)])
    torch.cuda.synchronize()
    print(k, "valid", v.tolist(), "words", [hex(int(a) & (2**64 - 1)) for a in w.tolist()])
    wav = sess.decode_step(w, v)
    torch.cuda.synchronize()
    print("   wav absmax", float(wav.abs().max()))

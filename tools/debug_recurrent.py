"""Bring-up aid: tensor-core recurrent path (precision 1) against the fp32 layer path (precision 0), per frame.

    python tools/debug_recurrent.py [--B 3] [--T 6]
"""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bernoulli_var_speech_codec_b200 import BVRNNCodecModel, SCALING
from bernoulli_var_speech_codec_b200.synth import write_synthetic_checkpoints

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=3)
ap.add_argument("--T", type=int, default=6)
a = ap.parse_args()
ck = write_synthetic_checkpoints(os.environ.get("BVC_CKPT_DIR", "/tmp/bvc_ckpts"), seed=1, sharpen=30.0)
m = BVRNNCodecModel(os.path.join(ROOT, "configs", "config_varBitRate.toml"), *ck).eval()
eng = m._engine
g = torch.Generator().manual_seed(1)
x = (0.1 * torch.randn(a.B, 256 * a.T + 600, generator=g)).clamp(-1, 1).cuda()
mel = eng.logmel(x, SCALING)[:, :a.T].contiguous()
h0 = (0.3 * torch.randn(a.B, 1024, generator=g)).cuda()
out = {}
for prec in (0, 1):
    eng.set_precision(prec)
    codes, all_h, h_fin, logits, packed = eng.encode(mel, None, 35.0, h0, want_logits=True, want_all_h=True, want_packed=True)
    dmel, dh = eng.decode_mel(codes, h0)
    torch.cuda.synchronize()
    out[prec] = dict(codes=codes.cpu(), all_h=all_h.cpu(), h_fin=h_fin.cpu(), logits=logits.cpu(), packed=packed.cpu(),
                     dmel=dmel.cpu(), dh=dh.cpu())
r, t = out[0], out[1]
print("ref |logits| max %.3f  |all_h| max %.3f" % (r["logits"].abs().max(), r["all_h"].abs().max()))
for f in range(a.T):
    print("frame %d: logits err %.3e  all_h err %.3e  dec_mel err %.3e  code diffs %d" % (
        f, (r["logits"][:, f] - t["logits"][:, f]).abs().max(), (r["all_h"][:, f] - t["all_h"][:, f]).abs().max(),
        (r["dmel"][:, f] - t["dmel"][:, f]).abs().max(), int((r["codes"][:, f] != t["codes"][:, f]).sum())))
print("h_final err %.3e  decode h_final err %.3e  packed equal %s" % (
    (r["h_fin"] - t["h_fin"]).abs().max(), (r["dh"] - t["dh"]).abs().max(), bool((r["packed"] == t["packed"]).all())))
e = (r["logits"][:, 0] - t["logits"][:, 0]).abs()
print("frame0 logits err per column block of 16:", [float(e[:, i:i + 16].max()) for i in range(0, 64, 16)])
e = (r["all_h"][:, 1] - t["all_h"][:, 1]).abs() if a.T > 1 else None
if e is not None:
    print("frame1 all_h err per 64 units:", ["%.1e" % float(e[:, i:i + 64].max()) for i in range(0, 1024, 64)])
    print("frame1 all_h err per unit%16 :", ["%.1e" % float(e[:, i::16].max()) for i in range(16)])
e = (r["dmel"][:, 0] - t["dmel"][:, 0]).abs()
print("frame0 dec_mel err per 16 cols:", ["%.1e" % float(e[:, i:i + 16].max()) for i in range(0, 80, 16)])

# ---- raw linear tap: dec.0[:, H:] . h0 after a one-frame encode (no bias, no activation) ----
sd = torch.load(ck[0], map_location="cpu")["vrnn"]
Wd0h = sd["dec.0.weight"][:, 1024:].double()
eng.set_precision(1)
eng.encode(mel[:, :1].contiguous(), None, 35.0, h0, want_logits=False, want_all_h=False)
dh = eng.debug_read("rec_dh", (a.B, 1024)).double()
ref = h0.cpu().double() @ Wd0h.T
print("rec_dh err %.3e (ref max %.3f)" % ((dh - ref).abs().max(), ref.abs().max()))
parts = [h0.cpu().double()[:, q * 256:(q + 1) * 256] @ Wd0h[:, q * 256:(q + 1) * 256].T for q in range(4)]
for q in range(4):
    print("  vs K-quarter %d alone: %.3e" % (q, (dh - parts[q]).abs().max()))
e = (dh - ref).abs()
print("  err by column%64 block of 16:", ["%.1e" % float(torch.stack([e[:, c:c + 16] for c in range(o, 1024, 64)]).max()) for o in (0, 16, 32, 48)])
print("  err by row:", ["%.1e" % float(e[r].max()) for r in range(a.B)])
print("  row0 first 8 got:", [round(float(v), 4) for v in dh[0, :8]], "ref:", [round(float(v), 4) for v in ref[0, :8]])

# ---- stale-read probe: identical calls; a missing barrier / visibility bug shows up as run-to-run change ----
for rep in range(3):
    codes, all_h, h_fin, logits, packed = eng.encode(mel, None, 35.0, h0, want_logits=True, want_all_h=True)
    torch.cuda.synchronize()
    print("repeat %d: frame0 logits err %.3e  frame1 all_h err %.3e" % (
        rep, (r["logits"][:, 0] - logits.cpu()[:, 0]).abs().max(), (r["all_h"][:, 1] - all_h.cpu()[:, 1]).abs().max()))
gh = eng.debug_read("rec_gh", (a.B, 3072)).double()
Whh = sd["rnn.weight_hh_l0"].double(); bhh = sd["rnn.bias_hh_l0"].double()
hl = all_h.cpu()[:, -1].double()      # state entering the last frame
ref = hl @ Whh.T + bhh                # natural gate order [r | z | n]
j = torch.arange(1024)
perm = torch.stack([12 * (j // 4) + 4 * g + (j % 4) for g in range(3)])   # interleaved column of (gate, unit)
got = torch.stack([gh[:, perm[g]] for g in range(3)], 1).reshape(a.B, 3072)
print("rec_gh (bias path, last frame) err %.3e" % (got - ref).abs().max())
# ---- giz = W_ih[:, H:] . phi_z(z) + b_ih of the last frame: checks e4 -> z image -> phi_z chain -> out_f ----
giz = eng.debug_read("rec_giz", (a.B, 3072)).double()
z = codes.cpu()[:, -1].double()
def lin(x, n): return x @ sd[n + ".weight"].double().T + sd[n + ".bias"].double()
elu = torch.nn.functional.elu
pz = elu(lin(elu(lin(elu(lin(z, "phi_z.0")), "phi_z.2")), "phi_z.4"))
ref = pz @ sd["rnn.weight_ih_l0"][:, 1024:].double().T + sd["rnn.bias_ih_l0"].double()
got = torch.stack([giz[:, perm[g]] for g in range(3)], 1).reshape(a.B, 3072)
print("rec_giz (phi_z chain on our own codes, last frame) err %.3e" % (got - ref).abs().max())
dhl = eng.debug_read("rec_dh", (a.B, 1024)).double()
d1 = elu(pz @ sd["dec.0.weight"][:, :1024].double().T + sd["dec.0.bias"].double() + dhl)
d3 = elu(lin(elu(lin(d1, "dec.2")), "dec.4"))
melr = lin(d3, "dec.6")
mn = (melr - sd["mean_mel"].double()) / sd["std_mel"].double()
px = elu(lin(elu(lin(elu(lin(mn, "phi_x.0")), "phi_x.2")), "phi_x.4"))
gi = px @ sd["rnn.weight_ih_l0"][:, :1024].double().T + ref
ghn = hl @ Whh.T + bhh
H = 1024
rg = torch.sigmoid(gi[:, :H] + ghn[:, :H]); zg = torch.sigmoid(gi[:, H:2 * H] + ghn[:, H:2 * H])
ng = torch.tanh(gi[:, 2 * H:] + rg * ghn[:, 2 * H:])
hn = (1 - zg) * ng + zg * hl
print("h_final vs float64 restatement from our own (codes, h_last): err %.3e" % (hn - h_fin.cpu().double()).abs().max())

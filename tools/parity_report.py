"""Stage-by-stage parity of the CUDA path against the CPU oracle (run on the GPU box).

    python tools/parity_report.py [--precision 0|1] [--B 4] [--seconds 1.0] [--golden-only]

Prints one JSON object per case.  Used for bring-up and for the numbers quoted in DESIGN.md.
"""
from __future__ import annotations

import os as _os
_os.environ.setdefault("BVC_VOC_FUSE_POST", "0")   # keep the last stage's partial tensors: the stage taps are compared below

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from bernoulli_var_speech_codec_b200 import BVRNNCodecModel, SCALING  # noqa: E402
from bernoulli_var_speech_codec_b200.synth import write_synthetic_checkpoints  # noqa: E402
from oracle.codec_oracle import OracleCodec  # noqa: E402
from parity import compare_codes, snr_db  # noqa: E402


def stage_report(model, oracle, x, bitrate, vocoder_taps=True):
    dev = model.device
    out = {}
    taps = {}
    t0 = time.time()
    o_codes = oracle.encode(x, bitrate, taps)
    o_taps2 = {}
    o_wav = oracle.decode(o_codes, x.shape[1], o_taps2)
    out["oracle_s"] = round(time.time() - t0, 2)
    xd = x.to(dev)
    mel = model._engine.logmel(xd, SCALING)
    out["mel_maxabs"] = float((mel.cpu() - taps["mel"]).abs().max())
    bits = model.bits_per_frame(bitrate)
    # recurrence parity on the oracle's mel (isolates the coder), then on the product's own mel
    out["codes_on_oracle_mel"] = compare_codes(model._engine, taps["mel"].to(dev), bits, o_codes, taps["logits"], taps["all_h"])
    out["codes_full_path"] = compare_codes(model._engine, mel, bits, o_codes, taps["logits"], taps["all_h"])
    r = model.encode_with_taps(xd, bitrate, mel=taps["mel"].to(dev))
    out["all_h_maxabs_first8"] = float((r["all_h"][:, :8].cpu() - taps["all_h"][:, :8]).abs().max())
    # decoder side on the ORACLE's codes
    codes_d = o_codes.to(dev)
    dmel, _ = model._engine.decode_mel(codes_d, None)
    out["dec_mel_maxabs"] = float((dmel.cpu() - o_taps2["dec_mel"]).abs().max())
    wav_from_oracle_mel = model._engine.vocode(o_taps2["dec_mel"].to(dev), x.shape[1], SCALING).cpu()
    if vocoder_taps:
        B, T = x.shape[0], o_codes.shape[1]
        pre = model._engine.debug_read("voc_pre", (B, T + 6, 128))[:, :T].permute(0, 2, 1)
        out["voc_pre_maxabs"] = float((pre - o_taps2["pre"]).abs().max())
        for i in range(4):
            ref = o_taps2[f"stage{i}"]
            shp = (ref.shape[0], ref.shape[2], ref.shape[1])          # partials are channel-last [B, n, C]
            parts = [model._engine.debug_read(f"voc_stage{i}_{j}", shp) for j in range(3)]
            got = (((parts[0] + parts[1]) + parts[2]) / 3.0).permute(0, 2, 1)
            out[f"voc_stage{i}_maxabs"] = float((got - ref).abs().max())
            out[f"voc_stage{i}_refmax"] = float(ref.abs().max())
    out["voc_wav_maxabs"] = float((wav_from_oracle_mel - o_wav).abs().max())
    out["voc_wav_snr_db"] = round(snr_db(o_wav.numpy(), wav_from_oracle_mel.numpy()), 2)
    wav = model.decode(codes_d, x.shape[1]).cpu()
    out["decode_wav_maxabs"] = float((wav - o_wav).abs().max())
    out["decode_wav_snr_db"] = round(snr_db(o_wav.numpy(), wav.numpy()), 2)
    # host-buffer API
    c_host = model.encode(x, bitrate)
    out["host_encode_equals_device"] = bool(torch.equal(c_host, model.encode(xd, bitrate).cpu()))
    w_host = model.decode(o_codes, x.shape[1])
    out["host_decode_maxabs_vs_device"] = float((w_host - wav).abs().max())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", type=int, default=0)
    ap.add_argument("--B", type=int, default=4)
    ap.add_argument("--seconds", type=float, default=1.0)
    ap.add_argument("--golden-only", action="store_true")
    a = ap.parse_args()
    ck = write_synthetic_checkpoints(os.environ.get("BVC_CKPT_DIR", "/tmp/bvc_ckpts"), seed=1, sharpen=30.0)
    cfg = os.path.join(ROOT, "configs", "config_varBitRate.toml")
    model = BVRNNCodecModel(cfg, *ck).eval()
    model._engine.set_precision(a.precision)
    oracle = OracleCodec(cfg, *ck)
    g = np.load(os.path.join(ROOT, "tests", "golden", "synth_var_small.npz"))
    rep = stage_report(model, oracle, torch.from_numpy(g["x"]), float(g["bitrate"]))
    print(json.dumps({"case": "golden synth_var_small", "precision": a.precision, **rep}), flush=True)
    if not a.golden_only:
        gen = torch.Generator().manual_seed(1234)
        L = int(a.seconds * 22050)
        x = (0.1 * torch.randn(a.B, L, generator=gen)).clamp(-1, 1)
        rep = stage_report(model, oracle, x, 3000)
        print(json.dumps({"case": f"noise B={a.B} L={L}", "precision": a.precision, **rep}), flush=True)
    print(json.dumps({"kernel_launches": model._engine.kernel_launches()}))


if __name__ == "__main__":
    main()

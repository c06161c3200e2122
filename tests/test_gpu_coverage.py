"""GPU parity tests added in round 2: less-contractive weight regimes, the B > 256 schedules (8 m-tiles, forced batch
chunks), BASELINE configs[2] at full size, the asynchronous C ABI, two devices in one process, and the real (git-LFS)
checkpoints when they are present.  Same tolerances as test_gpu_parity.py unless stated."""
import hashlib
import json
import os
import time

import numpy as np
import pytest
import torch

from parity import compare_codes, snr_db

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT_DIR = os.path.join(ROOT, "gpurun_out")


def _noise(B, L, seed):
    g = torch.Generator().manual_seed(seed)
    return (0.1 * torch.randn(B, L, generator=g)).clamp(-1, 1)


def _pair(cfg, gain, sharpen, precision):
    from bernoulli_var_speech_codec_b200 import BVRNNCodecModel
    from bernoulli_var_speech_codec_b200.synth import write_synthetic_checkpoints
    from oracle.codec_oracle import OracleCodec
    ck = write_synthetic_checkpoints(os.environ.get("BVC_CKPT_DIR", "/tmp/bvc_ckpts"), seed=1, sharpen=sharpen, gain=gain)
    m = BVRNNCodecModel(cfg, *ck).eval()
    m._engine.set_precision(precision)
    return m, OracleCodec(cfg, *ck)


def _code_report(model, oracle, x, bitrate):
    taps = {}
    o_codes = oracle.encode(x, bitrate, taps)
    mel = model._engine.logmel(x.to(model.device), 10 ** (-10 / 20))
    rep = compare_codes(model._engine, mel, model.bits_per_frame(bitrate), o_codes, taps["logits"], taps["all_h"],
                        max_resync=400)
    rep["median_abs_logit"] = float(taps["logits"].abs().median())
    return rep, o_codes


# (gain on every weight matrix, sharpening of enc.4).  gain > 1 makes the recurrence less
# contractive (closer to a trained net): rounding differences in h then persist instead of dying out.
REGIMES = [(1.0, 30.0), (1.5, 30.0), (2.0, 30.0), (3.0, 10.0)]


@pytest.mark.parametrize("gain,sharpen", REGIMES)
@pytest.mark.parametrize("precision", [1, 0], ids=["tensorcore", "fp32"])
def test_weight_regimes(gain, sharpen, precision, cfg_var):
    """2 x 4 s, 64 bits per frame (every bit active, App. F's set-up) in each regime and arithmetic mode.  Reported per
    regime: eps-bits and hard mismatches (written to gpurun_out/r02_regimes.json for DESIGN.md section 2).
    Bar: the fp32 mode is exact in every regime; the split-bf16 mode (16 mantissa bits per product) is exact in the
    contractive-to-moderate regimes and may show hard mismatches only in the gain-3 row, where App. F measured 1.4 % for a
    two-term bf16 split -- those are counted and reported, and fp32 is the documented fallback."""
    model, oracle = _pair(cfg_var, gain, sharpen, precision)
    x = _noise(2, 4 * 22050, 77)
    rep, o_codes = _code_report(model, oracle, x, 5513)
    rep.update(gain=gain, sharpen=sharpen, precision=precision)
    os.makedirs(OUT_DIR, exist_ok=True)
    path = os.path.join(OUT_DIR, "r02_regimes.json")
    rows = json.load(open(path)) if os.path.exists(path) else []
    rows = [r for r in rows if not (r["gain"] == gain and r["sharpen"] == sharpen and r["precision"] == precision)] + [rep]
    json.dump(rows, open(path, "w"), indent=1)
    print("regime", rep)
    assert rep["mask_errors"] == 0
    if precision == 0 or gain <= 2.0:
        assert rep["hard_mismatches"] == 0, rep
        assert rep["eps_bits"] <= max(2, rep["total_bits"] // 2000), rep
    else:
        assert rep["hard_mismatches"] <= 0.05 * rep["total_bits"], rep
    # decoder side in the same regime: decode the ORACLE's codes
    o_taps = {}
    o_wav = oracle.decode(o_codes, x.shape[1], o_taps)
    dmel, _ = model._engine.decode_mel(o_codes.to(model.device), None)
    err = (dmel.cpu() - o_taps["dec_mel"]).abs().max().item()
    scale = max(1.0, o_taps["dec_mel"].abs().max().item())
    assert err <= 5e-4 * scale, (err, scale)
    wav = model.decode(o_codes.to(model.device), x.shape[1]).cpu()
    assert snr_db(o_wav.numpy(), wav.numpy()) >= 60.0


def test_eight_mtiles_b1024(oracle_var, ckpts, cfg_var):
    """B = 1024 x 2 s (BASELINE configs[3] batch): 8 m-tiles share the 32 clusters.  Oracle parity on one row of every
    m-tile (first / last / interior positions) plus batch-size invariance against a small batch."""
    from bernoulli_var_speech_codec_b200 import BVRNNCodecModel
    m = BVRNNCodecModel(cfg_var, *ckpts).eval()
    B, L = 1024, 2 * 22050
    x = _noise(B, L, 2024)
    xd = x.to(m.device)
    codes = m.encode(xd, 3000)
    assert codes.shape == (B, L // 256, 64)
    rows = torch.tensor([0, 127 + 128, 2 * 128 + 5, 3 * 128 + 64, 4 * 128 + 127, 5 * 128, 6 * 128 + 77, 1023])
    o_codes = oracle_var.encode(x[rows], 3000)
    got = codes[rows.to(codes.device)].cpu()
    assert ((got == 0.5) == (o_codes == 0.5)).all()
    mism = int((got != o_codes).sum())
    assert mism == 0, f"{mism} differing code entries on the sampled rows"
    assert torch.equal(m.encode(xd[rows].contiguous(), 3000), codes[rows.to(codes.device)])
    wav = m.decode(codes, L)
    o_wav = oracle_var.decode(o_codes, L)
    assert snr_db(o_wav.numpy(), wav[rows.to(wav.device)].cpu().numpy()) >= 60.0


def test_forced_batch_chunks(oracle_var, ckpts, cfg_var):
    """BVC_REC_CLUSTERS=2 caps the persistent kernel at 2 clusters -> 128 rows per launch, so B = 300 runs the chunk loop
    of bvrnn_encode / bvrnn_decode (3 launches: 128 + 128 + 44 rows).  Rows of every chunk are checked against the oracle
    and against the unchunked run."""
    from bernoulli_var_speech_codec_b200 import BVRNNCodecModel
    m = BVRNNCodecModel(cfg_var, *ckpts).eval()
    B, L = 300, 22050
    x = _noise(B, L, 808)
    xd = x.to(m.device)
    ref_codes = m.encode(xd, 3000)
    ref_wav = m.decode(ref_codes, L)
    old = os.environ.get("BVC_REC_CLUSTERS")
    os.environ["BVC_REC_CLUSTERS"] = "2"
    try:
        codes = m.encode(xd, 3000)
        wav = m.decode(codes, L)
        ms = m._engine.recurrent_ms(1)
        assert ms > 0
    finally:
        if old is None:
            del os.environ["BVC_REC_CLUSTERS"]
        else:
            os.environ["BVC_REC_CLUSTERS"] = old
    assert torch.equal(codes, ref_codes)
    assert snr_db(ref_wav.cpu().numpy(), wav.cpu().numpy()) >= 100.0
    rows = torch.tensor([0, 127, 128, 255, 256, 299])
    o_codes = oracle_var.encode(x[rows], 3000)
    assert torch.equal(codes[rows.to(codes.device)].cpu(), o_codes)
    o_wav = oracle_var.decode(o_codes, L)
    assert snr_db(o_wav.numpy(), wav[rows.to(wav.device)].cpu().numpy()) >= 60.0


def test_config3_fixed64_full_size(oracle_fix, ckpts, cfg_fix):
    """BASELINE configs[2]: config_64bit.toml + fixed-rate weights, B = 256 x 10 s, all 64 bits active, bitrate ignored;
    bit-exact code check against the oracle on two utterances (first and last m-tile)."""
    from bernoulli_var_speech_codec_b200 import BVRNNCodecModel
    m = BVRNNCodecModel(cfg_fix, *ckpts).eval()
    B, L = 256, 220500
    x = _noise(B, L, 64)
    xd = x.to(m.device)
    codes = m.encode(xd, 3000)
    assert codes.shape == (B, L // 256, 64)
    assert bool(((codes == 0.0) | (codes == 1.0)).all())              # no masked positions in the fixed-rate coder
    assert torch.equal(codes[:4], m.encode(xd[:4].contiguous(), 123.0))   # bitrate is ignored (bvrnn.py:180-184,193)
    rows = torch.tensor([3, 200])
    taps = {}
    o_codes = oracle_fix.encode(x[rows], 3000, taps)
    mel = m._engine.logmel(xd[rows.to(xd.device)].contiguous(), 10 ** (-10 / 20))
    rep = compare_codes(m._engine, mel, m.bits_per_frame(3000), o_codes, taps["logits"], taps["all_h"])
    assert rep["hard_mismatches"] == 0 and rep["mask_errors"] == 0, rep
    got = codes[rows.to(codes.device)].cpu()
    assert rep["eps_bits"] > 0 or torch.equal(got, o_codes)
    wav = m.decode(codes, L)
    assert wav.shape == (B, L) and bool(torch.isfinite(wav).all())
    o_wav = oracle_fix.decode(got, L)
    assert snr_db(o_wav.numpy(), wav[rows.to(wav.device)].cpu().numpy()) >= 60.0


def test_device_entry_points_do_not_block_the_host(ckpts, cfg_var):
    """ABI 3 (bvc.h conventions): bvc_logmel / bvc_encode / bvc_decode_mel / bvc_vocode only enqueue.  The host returns
    from enqueueing a whole encode -> decode step long before the device finishes it, and results are unchanged."""
    from bernoulli_var_speech_codec_b200 import BVRNNCodecModel
    m = BVRNNCodecModel(cfg_var, *ckpts).eval()
    dev = m.device
    B, L = 128, 4 * 22050
    xd = _noise(B, L, 5).to(dev)
    for _ in range(2):                                   # warm-up: workspace growth and first-launch set-up do synchronise
        ref = m.decode(m.encode(xd, 3000), L)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    codes = m.encode(xd, 3000)
    wav = m.decode(codes, L)
    e1.record()
    t_enqueue = time.perf_counter() - t0
    torch.cuda.synchronize(dev)
    t_device = e0.elapsed_time(e1) / 1e3
    print(f"enqueue {1e3 * t_enqueue:.2f} ms, device {1e3 * t_device:.2f} ms")
    assert t_enqueue < 0.25 * t_device, (t_enqueue, t_device)
    assert torch.equal(wav, ref)
    m._engine.check()                                     # no launch aborted
    assert m._engine.recurrent_ms(0) > 0 and m._engine.recurrent_ms(1) > 0


def test_packed_output_needs_z64():
    """bvc_encode refuses a packed output for z_dim != 64 instead of leaving the caller's buffer untouched."""
    import tomllib
    from bernoulli_var_speech_codec_b200 import BVRNNCodecModel
    from bernoulli_var_speech_codec_b200.synth import synth_bvrnn_state_dict, write_synthetic_checkpoints
    d = os.environ.get("BVC_CKPT_DIR", "/tmp/bvc_ckpts")
    _, voc = write_synthetic_checkpoints(d, seed=1, sharpen=30.0)
    p = os.path.join(d, "bvrnn_synth_z32")
    torch.save({"vrnn": synth_bvrnn_state_dict(seed=3, z_dim=32)}, p)
    cfg_src = open(os.path.join(ROOT, "configs", "config_varBitRate.toml")).read()
    assert tomllib.loads(cfg_src)["z_dim"] == 64
    cfg32 = os.path.join(d, "config_z32.toml")
    open(cfg32, "w").write(cfg_src.replace("z_dim = 64", "z_dim = 32"))
    assert tomllib.load(open(cfg32, "rb"))["z_dim"] == 32
    m = BVRNNCodecModel(cfg32, p, voc).eval()
    x = _noise(2, 6000, 1).to(m.device)
    codes = m.encode(x, 1000)                             # layer-by-layer path (the persistent kernel needs z_dim 64)
    assert codes.shape == (2, 6000 // 256, 32)
    with pytest.raises(RuntimeError, match="z_dim == 64"):
        m.encode_packed(x, 1000)


def test_two_devices_in_one_process(ckpts, cfg_var):
    """One handle per device (bvc.h): a second model on cuda:1 in the same process gets its own opt-in shared-memory
    attributes / occupancy figures (per-device caches) and produces the same codes."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from bernoulli_var_speech_codec_b200 import BVRNNCodecModel
    x = _noise(3, 30000, 12)
    outs = []
    for d in ("cuda:0", "cuda:1"):
        m = BVRNNCodecModel(cfg_var, *ckpts, device=d).eval()
        c = m.encode(x.to(d), 3000)
        w = m.decode(c, x.shape[1])
        assert c.device == torch.device(d)
        outs.append((c.cpu(), w.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


# git-LFS object ids of the shipped checkpoints (reference chkpts/*: pointer files)
REAL_CKPTS = {
    "bvrnn_var_bitrate_step200000": "22734e7b",
    "bvrnn_fixed_bitrate64_step200000": "566af36e",
    "bigvgan_causal_tiny_ftbvrnn_g_step3500000": "9d6efb4d",
}


def _real_ckpt(name):
    for d in (os.environ.get("BVC_REAL_CKPT_DIR", ""), os.path.join(ROOT, "chkpts")):
        p = os.path.join(d, name) if d else ""
        if p and os.path.isfile(p) and os.path.getsize(p) > 1 << 20:       # an LFS pointer file is ~130 bytes
            h = hashlib.sha256()
            with open(p, "rb") as fh:
                for blk in iter(lambda: fh.read(1 << 22), b""):
                    h.update(blk)
            if h.hexdigest().startswith(REAL_CKPTS[name]):
                return p
    return None


@pytest.mark.parametrize("which", ["var", "fix"])
def test_real_checkpoints_when_present(which, cfg_var, cfg_fix):
    """The shipped weights (sha256-checked against the LFS oids) load unchanged and meet the same parity bars as the
    synthetic ones on stim_01.  Skipped when the blobs are absent (they are LFS pointers in the reference tree)."""
    b = _real_ckpt("bvrnn_var_bitrate_step200000" if which == "var" else "bvrnn_fixed_bitrate64_step200000")
    v = _real_ckpt("bigvgan_causal_tiny_ftbvrnn_g_step3500000")
    if not (b and v):
        pytest.skip("real checkpoints not present (git-LFS blobs are not in the reference tree)")
    from bernoulli_var_speech_codec_b200 import BVRNNCodecModel
    from conftest import golden
    from oracle.codec_oracle import OracleCodec
    cfg = cfg_var if which == "var" else cfg_fix
    m, o = BVRNNCodecModel(cfg, b, v).eval(), OracleCodec(cfg, b, v)
    x = torch.from_numpy(golden("stim01_var.npz")["x"])
    rep, o_codes = _code_report(m, o, x, 3000)
    assert rep["hard_mismatches"] == 0 and rep["mask_errors"] == 0, rep
    wav = m.decode(o_codes.to(m.device), x.shape[1]).cpu()
    assert snr_db(o.decode(o_codes, x.shape[1]).numpy(), wav.numpy()) >= 60.0

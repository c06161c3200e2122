"""GPU parity tests: the CUDA path (through the facade -> C ABI) against the CPU oracle and the
golden vectors made by the unmodified reference.  Tolerances are stated per check.

  codes   : bit-exact except bits within EPS_PROB=1e-4 of the decision threshold (counted, re-synced)
  log-mel : max-abs <= 2e-4 (fp32 FFT / mel accumulation order)
  dec mel : max-abs <= 5e-4
  wave    : SNR >= 60 dB (output range +-3.16)

Every test that takes model_var / model_fix runs twice (conftest.PRECISIONS): on the split-bf16 tensor-core kernels
(library default, the path bench.py measures) and on the fp32 FFMA kernels, with the SAME tolerances.
"""
import numpy as np
import pytest
import torch

from conftest import golden
from parity import compare_codes, snr_db

pytestmark = pytest.mark.gpu

MEL_TOL = 2e-4


def _noise(B, L, seed):
    g = torch.Generator().manual_seed(seed)
    return (0.1 * torch.randn(B, L, generator=g)).clamp(-1, 1)


def _check_case(model, oracle, x, bitrate, dec_tol=5e-4, snr_min=60.0):
    dev = model.device
    taps = {}
    o_codes = oracle.encode(x, bitrate, taps)
    xd = x.to(dev)
    r = model.encode_with_taps(xd, bitrate)
    assert (r["mel"].cpu() - taps["mel"]).abs().max() <= MEL_TOL
    rep = compare_codes(model._engine, r["mel"], model.bits_per_frame(bitrate), o_codes, taps["logits"], taps["all_h"])
    assert rep["hard_mismatches"] == 0 and rep["mask_errors"] == 0, rep
    assert rep["eps_bits"] <= max(2, rep["total_bits"] // 2000), rep
    o_taps = {}
    o_wav = oracle.decode(o_codes, x.shape[1], o_taps)
    dmel, _ = model._engine.decode_mel(o_codes.to(dev), None)
    assert (dmel.cpu() - o_taps["dec_mel"]).abs().max() <= dec_tol
    wav = model.decode(o_codes.to(dev), x.shape[1]).cpu()
    assert wav.shape == o_wav.shape
    assert snr_db(o_wav.numpy(), wav.numpy()) >= snr_min
    return rep


@pytest.mark.parametrize("name,which", [("synth_var_small.npz", "var"), ("synth_fix_small.npz", "fix")])
def test_golden_small(name, which, model_var, model_fix, oracle_var, oracle_fix):
    g = golden(name)
    model, oracle = (model_var, oracle_var) if which == "var" else (model_fix, oracle_fix)
    x = torch.from_numpy(g["x"])
    _check_case(model, oracle, x, float(g["bitrate"]))
    # and directly against the reference's own outputs
    codes = model.encode(x.to(model.device), float(g["bitrate"])).cpu().numpy()
    assert ((codes == 0.5) == (g["codes"] == 0.5)).all()
    wav = model.decode(torch.from_numpy(g["codes"]).to(model.device), x.shape[1]).cpu().numpy()
    assert snr_db(g["wav"], wav) >= 60.0


def test_golden_config1_stim01(model_var, oracle_var):
    """BASELINE config #1: MUSHRA stim_01 ref.wav at 3000 bps, batch 1."""
    g = golden("stim01_var.npz")
    x = torch.from_numpy(g["x"])
    _check_case(model_var, oracle_var, x, 3000)
    y = model_var(x.to(model_var.device), 3000)
    assert y.shape == x.shape


@pytest.mark.parametrize("tag,nbits", [("0", 0), ("1", 1), ("64", 64), ("gt64", 64)])
def test_bit_budget_edges(tag, nbits, model_var, oracle_var):
    g = golden(f"synth_var_bits_{tag}.npz")
    x = torch.from_numpy(g["x"])
    codes = model_var.encode(x.to(model_var.device), float(g["bitrate"])).cpu().numpy()
    assert (codes[:, :, nbits:] == 0.5).all() and np.isin(codes[:, :, :nbits], (0.0, 1.0)).all()
    _check_case(model_var, oracle_var, x, float(g["bitrate"]))


def test_fixed_rate_ignores_bitrate(model_fix):
    x = _noise(2, 4000, 3).to(model_fix.device)
    a, b = model_fix.encode(x, 100), model_fix.encode(x, 5000)
    assert torch.equal(a, b) and np.isin(a.cpu().numpy(), (0.0, 1.0)).all()


def test_ragged_lengths_and_batch(model_var, oracle_var):
    for L in (513, 1000, 256 * 9, 256 * 9 + 255):
        x = _noise(3, L, L)
        _check_case(model_var, oracle_var, x, 3000)
        y = model_var(x.to(model_var.device), 3000)
        assert y.shape == (3, L)


def test_short_input_rejected(model_var):
    with pytest.raises(RuntimeError):
        model_var.encode(torch.zeros(1, 512, device=model_var.device), 3000)


def test_batch_invariance_and_packed(model_var):
    x = _noise(5, 6000, 7).to(model_var.device)
    r = model_var.encode_with_taps(x, 3000)
    r1 = model_var.encode_with_taps(x[2:3].contiguous(), 3000)
    assert torch.equal(r["codes"][2:3], r1["codes"])
    codes = r["codes"].cpu().numpy()
    packed = r["packed"].cpu().numpy().astype(np.uint64)
    bits = ((packed[..., None] >> np.arange(64, dtype=np.uint64)) & np.uint64(1)).astype(np.float32)
    assert np.array_equal(bits == 1.0, codes == 1.0)


def test_operator_level_api(model_var, oracle_var):
    """BVRNN.encode/decode with per-frame bit budgets and carried state; BigVGAN.forward; mel_spectrogram."""
    from bernoulli_var_speech_codec_b200 import mel_spectrogram, SCALING
    from oracle.codec_oracle import bvrnn_encode
    dev = model_var.device
    x = _noise(2, 5000, 21)
    mel_o = oracle_var.logmel(x)
    mel = mel_spectrogram(x.to(dev) * SCALING, 1024, 80, 22050, 256, 1024, 0, 8000, 256)
    assert mel.shape == (2, 80, mel_o.shape[1])
    assert (mel.permute(0, 2, 1).cpu() - mel_o).abs().max() <= MEL_TOL
    T = mel_o.shape[1]
    bits = torch.randint(0, 70, (2, T)).float()
    h0 = 0.1 * torch.randn(1, 2, 1024)
    z_o, all_h_o, _ = bvrnn_encode(oracle_var.sd, mel_o, bits, h0[0], True)
    z, all_h = model_var.bvrnn.encode(mel_o.to(dev), bits.to(dev), h0.to(dev))
    assert z.shape == z_o.shape and all_h.shape == all_h_o.shape
    assert ((z.cpu() == 0.5) == (z_o == 0.5)).all()
    assert (z.cpu() != z_o).float().mean() < 1e-3
    assert (all_h[:, 0].cpu() - h0[0]).abs().max() == 0
    # decode in two chunks with carried h equals one pass (bit-identical)
    m_full, h_full = model_var.bvrnn.decode(z_o.to(dev), torch.zeros(1, 2, 1024, device=dev))
    m_a, h_a = model_var.bvrnn.decode(z_o[:, :7].to(dev), torch.zeros(1, 2, 1024, device=dev))
    m_b, h_b = model_var.bvrnn.decode(z_o[:, 7:].contiguous().to(dev), h_a)
    assert h_full.shape == (1, 2, 1024)
    assert torch.equal(torch.cat([m_a, m_b], 1), m_full) and torch.equal(h_b, h_full)
    # bare vocoder: (B,80,T) -> (B,1,length), no /SCALING
    m_o, _ = oracle_var.decode_mel(z_o)
    w_o = oracle_var.vocoder(m_o.permute(0, 2, 1), 5000)
    w = model_var.vocoder(m_o.permute(0, 2, 1).to(dev), 5000)
    assert w.shape == w_o.shape and snr_db(w_o.numpy(), w.cpu().numpy()) >= 60.0
    assert model_var.vocoder(m_o.permute(0, 2, 1).to(dev), 10 ** 6).shape[-1] == 256 * T + 294


def test_host_buffer_api_matches_device_api(model_var):
    x = _noise(2, 7000, 9)
    c_dev = model_var.encode(x.to(model_var.device), 3000)
    c_host = model_var.encode(x, 3000)
    assert c_host.device.type == "cpu" and torch.equal(c_host, c_dev.cpu())
    w_dev = model_var.decode(c_dev, 7000)
    w_host = model_var.decode(c_host, 7000)
    assert w_host.device.type == "cpu" and torch.equal(w_host, w_dev.cpu())


def test_strict_checkpoint_schema(ckpts, cfg_var, tmp_path):
    from bernoulli_var_speech_codec_b200 import BVRNNCodecModel
    ck = torch.load(ckpts[0], map_location="cpu", weights_only=True)
    bad = dict(ck["vrnn"])
    bad.pop("enc.2.bias")
    p = tmp_path / "bad"
    torch.save({"vrnn": bad}, p)
    with pytest.raises(RuntimeError, match="missing key"):
        BVRNNCodecModel(cfg_var, str(p), ckpts[1])
    bad = dict(ck["vrnn"])
    bad["enc.4.weight"] = bad["enc.4.weight"][:32]
    torch.save({"vrnn": bad}, p)
    with pytest.raises(RuntimeError, match="size mismatch"):
        BVRNNCodecModel(cfg_var, str(p), ckpts[1])


def test_default_mode_is_tensor_core(ckpts, cfg_var, oracle_var):
    """A freshly constructed model runs the tensor-core kernels (mode 1) without any set_precision call."""
    from bernoulli_var_speech_codec_b200 import BVRNNCodecModel
    m = BVRNNCodecModel(cfg_var, *ckpts).eval()
    x = _noise(4, 8000, 33)
    _check_case(m, oracle_var, x, 3000)
    m2 = BVRNNCodecModel(cfg_var, *ckpts).eval()
    m2._engine.set_precision(1)
    assert torch.equal(m.encode(x.to(m.device), 3000), m2.encode(x.to(m2.device), 3000))


def test_larger_batch_properties(model_var, oracle_var):
    """B=16 x 2 s: oracle parity on a subset plus size-independent properties on the whole batch."""
    B, L = 16, 44100
    x = _noise(B, L, 1234)
    xd = x.to(model_var.device)
    codes = model_var.encode(xd, 3000)
    c = codes.cpu().numpy()
    assert c.shape == (B, L // 256, 64)
    assert (c[:, :, 35:] == 0.5).all() and np.isin(c[:, :, :35], (0.0, 1.0)).all()
    wav = model_var.decode(codes, L)
    assert wav.shape == (B, L) and torch.isfinite(wav).all() and wav.abs().max() <= 1.0 / (10 ** (-10 / 20)) + 1e-4
    # forward == decode(encode(x)); it runs ONE recurrence (the encoder's internal decoder output feeds the vocoder,
    # SURVEY.md F7), so the two differ by the rounding of differently grouped fp32 sums only
    fwd = model_var(xd, 3000)
    assert fwd.shape == wav.shape and _snr_db(wav, fwd) >= 90.0
    # permutation equivariance over utterances (rows are independent)
    perm = torch.randperm(B)
    assert torch.equal(model_var.encode(xd[perm].contiguous(), 3000), codes[perm])
    _check_case(model_var, oracle_var, x[:2], 3000)


def test_full_size_batch_properties(oracle_var, ckpts, cfg_var):
    """BASELINE configs[1] size (B=256 x 10 s, 3 kbps) on the default tensor-core path: properties that do not need the
    oracle at that size (mask layout, value sets, batch-size invariance, forward == decode(encode)) plus oracle parity on
    two of the utterances."""
    from bernoulli_var_speech_codec_b200 import BVRNNCodecModel
    m = BVRNNCodecModel(cfg_var, *ckpts).eval()
    B, L = 256, 220500
    x = _noise(B, L, 4321)
    xd = x.to(m.device)
    codes = m.encode(xd, 3000)
    assert codes.shape == (B, L // 256, 64)
    assert bool((codes[:, :, 35:] == 0.5).all()) and bool(((codes[:, :, :35] == 0.0) | (codes[:, :, :35] == 1.0)).all())
    frac_ones = codes[:, :, :35].mean().item()
    assert 0.2 < frac_ones < 0.8                                     # the synthetic encoder is not stuck
    wav = m.decode(codes, L)
    assert wav.shape == (B, L) and bool(torch.isfinite(wav).all())
    # rows are independent: a row coded inside the full batch equals the same row coded in a small batch, bit for bit
    sub = torch.tensor([0, 1, 127, 128, 255])
    codes_sub = m.encode(xd[sub].contiguous(), 3000)
    assert torch.equal(codes_sub, codes[sub.to(codes.device)])
    wav_sub = m.decode(codes_sub, L)
    assert _snr_db(wav[sub.to(wav.device)], wav_sub) >= 100.0         # vocoder tiling depends on B: equal to fp32 rounding
    assert _snr_db(wav[:8], m(xd[:8].contiguous(), 3000)) >= 90.0     # one-recurrence forward
    _check_case(m, oracle_var, x[126:128], 3000)


def _snr_db(ref, test):
    return snr_db(ref.detach().cpu().numpy(), test.detach().cpu().numpy())


def test_fused_forward(model_var, model_fix, oracle_var):
    """forward() takes the decoder's mel from the encode kernel (bvc_encode_mel) instead of decoding again: the side output
    equals BVRNN.decode of the same codes, the codes are those of encode(), and the waveform matches the oracle's forward."""
    x = _noise(3, 12000, 91)
    for m in (model_var, model_fix):
        xd = x.to(m.device)
        eng = m._engine
        from bernoulli_var_speech_codec_b200 import SCALING
        mel = eng.logmel(xd, SCALING)
        bits = m.bits_per_frame(3000)
        out = eng.encode(mel, None, bits, None, want_all_h=False, want_mel=True)
        codes, mel_hat = out[0], out[5]
        assert torch.equal(codes, m.encode(xd, 3000))
        dec_mel, _ = eng.decode_mel(codes, None)
        assert (mel_hat - dec_mel).abs().max().item() <= 1e-4        # log-mel units; both are fp32-grade evaluations of dec_t
        assert _snr_db(m.decode(codes, x.shape[1]), m(xd, 3000)) >= 90.0
        assert m(x, 3000).device.type == "cpu"                        # host tensors in -> host tensor out, like encode/decode
    ref = oracle_var.forward(x, 3000)
    assert _snr_db(ref, model_var(x.to(model_var.device), 3000).cpu()) >= 60.0


def test_antialiased_activation_path(model_aa, oracle_aa):
    """vocoder_config.layers_antialias / antialias_post = true (Activation1d: FIR up -> SnakeBeta -> FIR down, fused in one
    shared-memory kernel; stages run layer by layer): golden vector of the unmodified reference, then the oracle on a
    longer ragged batch (several tiles per stage, replicate padding at both utterance edges)."""
    g = golden("synth_var_aa_small.npz")
    dev = model_aa.device
    x = torch.from_numpy(g["x"]).to(dev)
    codes = model_aa.encode(x, float(g["bitrate"]))
    assert np.array_equal(codes.cpu().numpy(), g["codes"])
    wav = model_aa.decode(torch.from_numpy(g["codes"]).to(dev), x.shape[1]).cpu().numpy()
    assert wav.shape == g["wav"].shape and snr_db(g["wav"], wav) >= 60.0
    # bare vocoder against the oracle: (B, 80, T) -> (B, 1, length)
    gen = torch.Generator().manual_seed(21)
    mel = -5.0 + 2.0 * torch.randn(3, 80, 37, generator=gen)
    w_o = oracle_aa.vocoder(mel, 37 * 256 + 100)
    w = model_aa.vocoder(mel.to(dev), 37 * 256 + 100).cpu()
    assert w.shape == w_o.shape and snr_db(w_o.numpy(), w.numpy()) >= 60.0
    assert np.abs(w.numpy() - w_o.numpy()).max() <= 2e-3
    # forward through the fused path too
    xs = _noise(2, 9000, 5)
    assert snr_db(oracle_aa.forward(xs, 3000).numpy(), model_aa(xs.to(dev), 3000).cpu().numpy()) >= 60.0


def test_packed_wire_format_round_trip(model_var, model_fix, oracle_var):
    """encode_packed / decode_packed (uint64 per frame) reproduce the float-code path bit for bit, for every budget edge."""
    x = _noise(3, 9000, 77).to(model_var.device)
    for bitrate in (0, 86.2, 3000, 5512.5, 9000):          # 0, 1, 35, 64, >64 bits per frame
        codes = model_var.encode(x, bitrate)
        packed, bits = model_var.encode_packed(x, bitrate)
        assert packed.dtype == torch.int64 and packed.shape == codes.shape[:2]
        assert torch.equal(model_var._engine.unpack_codes(packed, None, bits), codes)
        nb = int(min(max(bits, 0), 64))
        if nb < 64:
            assert (packed >> nb == 0).all()                  # masked bits are never set in the word
        assert torch.equal(model_var.decode_packed(packed, bits, x.shape[1]), model_var.decode(codes, x.shape[1]))
    # per-frame budgets through the packed path (reference bvrnn.py:180-182 takes varBitrate[B, T])
    T = x.shape[1] // 256
    vb = torch.randint(0, 65, (3, T), generator=torch.Generator().manual_seed(5)).float().to(model_var.device)
    mel = model_var._engine.logmel(x, 10 ** (-10 / 20))
    codes, _, _, _, packed = model_var._engine.encode(mel, vb, 0.0, None, want_all_h=False, want_packed=True)
    assert torch.equal(model_var._engine.unpack_codes(packed, vb, 0.0), codes)
    from oracle.codec_oracle import bvrnn_encode as oracle_bvrnn_encode
    with torch.no_grad():
        o_codes = oracle_bvrnn_encode(oracle_var.sd, mel.cpu(), vb.cpu(), torch.zeros(3, 1024), True)[0]
    assert ((codes.cpu() == 0.5) == (o_codes == 0.5)).all()             # masks follow the per-frame budgets exactly
    assert (codes.cpu() == o_codes).float().mean() > 0.999              # bits: up to threshold-band flips (and their wake)
    # fixed-rate model: every bit is active whatever the budget says
    pk, bits = model_fix.encode_packed(x, 3000)
    assert torch.equal(model_fix._engine.unpack_codes(pk, None, bits), model_fix.encode(x, 3000))


def test_streaming_equals_offline(model_var):
    """Chunk-by-chunk encode / decode with carried state (streaming.py) reproduces the offline calls exactly."""
    from bernoulli_var_speech_codec_b200.streaming import StreamingDecoder, StreamingEncoder
    B, L = 3, 256 * 61 + 100
    x = _noise(B, L, 91).to(model_var.device)
    codes_off = model_var.encode(x, 3000)
    wav_off = model_var.decode(codes_off, L)
    g = torch.Generator().manual_seed(3)
    for chunk_sizes in ([256] * 70, [1000, 37, 4096, 5, 700, 2048] * 4):      # hop-sized real-time feed; ragged chunks
        enc, dec = StreamingEncoder(model_var, 3000), StreamingDecoder(model_var)
        codes_s, wav_s, pos = [], [], 0
        for n in chunk_sizes:
            if pos >= L:
                break
            c = enc.push(x[:, pos:pos + n])
            pos = min(pos + n, L)
            if c is not None:
                codes_s.append(c)
                wav_s.append(dec.push(c))
        c = enc.flush()
        if c is not None:
            codes_s.append(c)
            wav_s.append(dec.push(c))
        codes_s = torch.cat(codes_s, 1)
        assert codes_s.shape == codes_off.shape and torch.equal(codes_s, codes_off)
        tail = dec.flush(L - 256 * codes_off.shape[1])
        wav_cat = torch.cat(wav_s + ([tail] if tail is not None else []), 1)
        assert wav_cat.shape == wav_off.shape
        assert (wav_cat - wav_off).abs().max() <= 1e-5

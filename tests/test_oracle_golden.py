"""CPU: the oracle restatement reproduces the golden vectors made by the unmodified reference."""
import numpy as np
import pytest
import torch

from conftest import golden


@pytest.mark.parametrize("name,which", [("synth_var_small.npz", "var"), ("synth_fix_small.npz", "fix"),
                                        ("stim01_var.npz", "var")])
def test_oracle_matches_reference_golden(name, which, oracle_var, oracle_fix):
    g = golden(name)
    o = oracle_var if which == "var" else oracle_fix
    x = torch.from_numpy(g["x"])
    taps = {}
    codes = o.encode(x, float(g["bitrate"]), taps)
    assert np.abs(taps["mel"].numpy() - g["mel"]).max() < 5e-6
    assert np.array_equal(codes.numpy(), g["codes"])                 # bit-exact
    assert np.abs(taps["logits"].numpy() - g["logits"]).max() < 2e-5
    stride = int(g["all_h_stride"]) if "all_h_stride" in g else 1
    assert np.abs(taps["all_h"].numpy()[:, ::stride] - g["all_h"]).max() < 5e-6
    taps2 = {}
    wav = o.decode(torch.from_numpy(g["codes"]), x.shape[1], taps2)
    assert np.abs(taps2["dec_mel"].numpy() - g["dec_mel"]).max() < 5e-6
    assert wav.shape == g["wav"].shape
    assert np.abs(wav.numpy() - g["wav"]).max() < 2e-5


@pytest.mark.parametrize("tag,nbits", [("0", 0), ("1", 1), ("64", 64), ("gt64", 64)])
def test_oracle_bit_budget_edges(tag, nbits, oracle_var):
    g = golden(f"synth_var_bits_{tag}.npz")
    x = torch.from_numpy(g["x"])
    codes = oracle_var.encode(x, float(g["bitrate"])).numpy()
    assert np.array_equal(codes, g["codes"])
    assert (codes[:, :, nbits:] == 0.5).all() and (codes[:, :, :nbits] != 0.5).all()
    wav = oracle_var.decode(torch.from_numpy(g["codes"]), x.shape[1]).numpy()
    assert np.abs(wav - g["wav"]).max() < 2e-5


def test_oracle_antialiased_vocoder_matches_reference_golden(oracle_aa):
    """Anti-aliased activations (Activation1d in every resblock and before conv_post): oracle == unmodified reference."""
    g = golden("synth_var_aa_small.npz")
    x = torch.from_numpy(g["x"])
    assert np.array_equal(oracle_aa.encode(x, float(g["bitrate"])).numpy(), g["codes"])
    taps = {}
    wav = oracle_aa.decode(torch.from_numpy(g["codes"]), x.shape[1], taps)
    assert np.abs(taps["dec_mel"].numpy() - g["dec_mel"]).max() < 5e-6
    assert wav.shape == g["wav"].shape and np.abs(wav.numpy() - g["wav"]).max() < 2e-5
    plain = golden("synth_var_small.npz")["wav"]                 # same codes path, different activations: must differ
    assert np.abs(plain - g["wav"]).max() > 1e-3


def test_antialias_filter_restatement():
    """synth.kaiser_sinc_filter (written into synthetic checkpoints) == the reference's kaiser_sinc_filter1d values."""
    from bernoulli_var_speech_codec_b200.synth import kaiser_sinc_filter
    g = golden("aa_filter.npz")
    f = kaiser_sinc_filter(float(g["cutoff"]), float(g["half_width"]), int(g["kernel_size"])).flatten().numpy()
    assert np.array_equal(f, g["filter"])


def test_oracle_rejects_short_input(oracle_var):
    with pytest.raises(RuntimeError):
        oracle_var.encode(torch.zeros(1, 512), 3000)


def test_reference_cross_check_when_available(ckpts, cfg_var):
    """Only in the build container: run the unmodified reference next to the oracle on a fresh input."""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("/root/reference not present (GPU box)")
    from oracle.codec_oracle import OracleCodec
    ref = ref_shim.import_reference()
    m = ref.BVRNNCodecModel(cfg_var, *ckpts).eval()
    o = OracleCodec(cfg_var, *ckpts)
    g = torch.Generator().manual_seed(5)
    x = (0.2 * torch.randn(1, 3000, generator=g)).clamp(-1, 1)
    with torch.no_grad():
        rc = m.encode(x, 1500)
        rw = m.decode(rc, x.shape[1])
    oc = o.encode(x, 1500)
    assert torch.equal(rc, oc)
    assert (o.decode(oc, x.shape[1]) - rw).abs().max() < 2e-5


def test_oracle_stochastic_mode_and_prior_match_reference(oracle_var):
    """Sampled bits z = round(u - 0.5 + p) with supplied uniforms, per-frame budgets, the prior head and the KL term against
    the unmodified reference's training forward (p_use_gen = 1, greedy = False; torch.rand_like patched to our uniforms)."""
    from oracle.codec_oracle import bvrnn_decode, bvrnn_encode, bvrnn_kld, bvrnn_prior
    g = golden("synth_var_stochastic.npz")
    mel, u, bits = (torch.from_numpy(g[k]) for k in ("mel", "uniforms", "bits"))
    with torch.no_grad():
        codes, all_h, _, logits = bvrnn_encode(oracle_var.sd, mel, bits, torch.zeros(2, 1024), True, want_taps=True, uniforms=u)
        assert np.array_equal(codes.numpy(), g["codes"])                              # bit-exact, masks included
        p = torch.sigmoid(logits)
        assert np.abs(p.numpy() - g["enc_p"]).max() < 2e-6
        prior = bvrnn_prior(oracle_var.sd, all_h)
        assert np.abs(prior.numpy() - g["prior_p"]).max() < 2e-6
        assert abs(float(bvrnn_kld(p, prior, bits, True)) - float(g["kld"][0])) < 1e-4
        dec, _ = bvrnn_decode(oracle_var.sd, codes, torch.zeros(2, 1024))
        assert np.abs(dec.numpy() - g["dec_mel"]).max() < 5e-6
        greedy = bvrnn_encode(oracle_var.sd, mel, bits, torch.zeros(2, 1024), True)[0]
        assert not np.array_equal(greedy.numpy(), g["codes"])                         # the uniforms matter

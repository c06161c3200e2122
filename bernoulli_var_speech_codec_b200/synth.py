"""Deterministic synthetic checkpoints with the reference's exact state-dict schema.

The shipped checkpoints of the reference are git-LFS pointers (no weights
offline), so tests, ``smoke()`` and ``bench.py`` use random-init weights of the
same architecture.  The files written here have the layout
``torch.load`` + strict ``load_state_dict`` expect in the reference
(bvrnn_codec_model.py:38-42): ``{'vrnn': sd}`` for the BVRNN and
``{'generator': sd}`` for the vocoder (old-style weight-norm triplets
``weight_g`` / ``weight_v`` / ``bias``).

Regime: Linear/GRU/Conv use PyTorch's default uniform
init; the last encoder layer is scaled (``sharpen``) so logits have a
realistic spread instead of sitting at the decision threshold; mel statistics
are set to typical log-mel values; snake alpha/beta ~ N(0, 0.5); weight_g is
perturbed away from ||v|| so the weight-norm fold is exercised.
"""
from __future__ import annotations

import math
import os

import torch


def _uniform(gen, shape, bound):
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2.0 - 1.0) * bound


def _linear(gen, sd, name, n_out, n_in, scale=1.0):
    bound = 1.0 / math.sqrt(n_in)
    sd[f"{name}.weight"] = _uniform(gen, (n_out, n_in), bound) * scale
    sd[f"{name}.bias"] = _uniform(gen, (n_out,), bound) * scale


def synth_bvrnn_state_dict(seed: int = 1, x_dim: int = 80, h_dim: int = 1024, z_dim: int = 64,
                           sharpen: float = 30.0, gain: float = 1.0):
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    sd["mean_mel"] = -5.7 + 0.8 * torch.randn(x_dim, generator=gen)
    sd["std_mel"] = 2.2 * (1.0 + 0.15 * (torch.rand(x_dim, generator=gen) * 2 - 1))
    sd["log_sigma"] = torch.tensor([-1.0])
    H = h_dim
    for i, (o, k) in zip((0, 2, 4), ((H, x_dim), (H, H), (H, H))):
        _linear(gen, sd, f"phi_x.{i}", o, k, gain)
    for i, (o, k) in zip((0, 2, 4), ((H, z_dim), (H, H), (H, H))):
        _linear(gen, sd, f"phi_z.{i}", o, k, gain)
    for i, (o, k) in zip((0, 2, 4), ((H, 2 * H), (H, H), (z_dim, H))):
        _linear(gen, sd, f"enc.{i}", o, k, gain * (sharpen if i == 4 else 1.0))
    for i, (o, k) in zip((0, 2, 4), ((H, H), (H, H), (z_dim, H))):
        _linear(gen, sd, f"prior.{i}", o, k, gain)
    for i, (o, k) in zip((0, 2, 4, 6), ((H, 2 * H), (H, H), (H, H), (x_dim, H))):
        _linear(gen, sd, f"dec.{i}", o, k, gain)
    b = 1.0 / math.sqrt(H)
    sd["rnn.weight_ih_l0"] = _uniform(gen, (3 * H, 2 * H), b) * gain
    sd["rnn.weight_hh_l0"] = _uniform(gen, (3 * H, H), b) * gain
    sd["rnn.bias_ih_l0"] = _uniform(gen, (3 * H,), b) * gain
    sd["rnn.bias_hh_l0"] = _uniform(gen, (3 * H,), b) * gain
    return sd


def _wn_conv(gen, sd, name, d0, d1, k, fan_in):
    bound = 1.0 / math.sqrt(fan_in)
    v = _uniform(gen, (d0, d1, k), bound)
    norm = v.reshape(d0, -1).norm(dim=1).reshape(d0, 1, 1)
    sd[f"{name}.weight_g"] = norm * (1.0 + 0.1 * torch.randn(d0, 1, 1, generator=gen))
    sd[f"{name}.weight_v"] = v


def kaiser_sinc_filter(cutoff: float, half_width: float, kernel_size: int) -> torch.Tensor:
    """The low-pass FIR of the anti-aliased activation, [1, 1, kernel_size] (restates
    third_party/BigVGAN/alias_free_torch/filter.py:28-59: Kaiser-windowed sinc, normalised to unit DC gain).
    The reference registers it as a buffer, so it travels in the checkpoint; tests pin this restatement against the
    reference's own values (tests/golden/aa_filter.npz)."""
    half = kernel_size // 2
    delta_f = 4 * half_width
    A = 2.285 * (half - 1) * math.pi * delta_f + 7.95
    beta = 0.1102 * (A - 8.7) if A > 50.0 else (0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21.0) if A >= 21.0 else 0.0)
    window = torch.kaiser_window(kernel_size, beta=beta, periodic=False)
    time = (torch.arange(-half, half) + 0.5) if kernel_size % 2 == 0 else (torch.arange(kernel_size) - half)
    f = 2 * cutoff * window * torch.sinc(2 * cutoff * time)
    return (f / f.sum()).view(1, 1, kernel_size)


def synth_vocoder_state_dict(seed: int = 2, vcfg: dict | None = None):
    """vcfg keys layers_antialias / antialias_post switch the activation entries to the reference's Activation1d schema
    (`.act.alpha`, `.act.beta` + the `upsample.filter` / `downsample.lowpass.filter` buffers; models.py:66-88,180-190)."""
    vcfg = vcfg or {}
    aa_layers = list(vcfg.get("layers_antialias", [False] * 4))
    aa_post = bool(vcfg.get("antialias_post", False))
    aa_filter = kaiser_sinc_filter(0.25, 0.3, 12)       # Activation1d defaults: ratio 2, 12 taps (act.py:10-15, resample.py:17-19)

    def act_entries(sd, name, ch, gen, aa):
        inner = name + (".act" if aa else "")
        sd[inner + ".alpha"] = 0.5 * torch.randn(ch, generator=gen)
        sd[inner + ".beta"] = 0.5 * torch.randn(ch, generator=gen)
        if aa:
            sd[name + ".upsample.filter"] = aa_filter.clone()
            sd[name + ".downsample.lowpass.filter"] = aa_filter.clone()

    num_mels = vcfg.get("num_mels", 80)
    c0 = vcfg.get("upsample_initial_channel", 128)
    rates = vcfg.get("upsample_rates", [8, 8, 2, 2])
    ksz = vcfg.get("upsample_kernel_sizes", [16, 16, 4, 4])
    rks = vcfg.get("resblock_kernel_sizes", [3, 7, 11])
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    _wn_conv(gen, sd, "conv_pre", c0, num_mels, 7, num_mels * 7)
    sd["conv_pre.bias"] = _uniform(gen, (c0,), 1.0 / math.sqrt(num_mels * 7))
    ch = c0
    for i, (u, k) in enumerate(zip(rates, ksz)):
        cin, cout = c0 >> i, c0 >> (i + 1)
        # ConvTranspose1d weight is [C_in, C_out, k]; torch's fan_in uses dim 1.
        _wn_conv(gen, sd, f"ups.{i}.1", cin, cout, k, cout * k)
        sd[f"ups.{i}.1.bias"] = _uniform(gen, (cout,), 1.0 / math.sqrt(cout * k))
        ch = cout
        for j, rk in enumerate(rks):
            n = i * len(rks) + j
            for grp in ("convs1", "convs2"):
                for l in range(3):
                    _wn_conv(gen, sd, f"resblocks.{n}.{grp}.{l}", ch, ch, rk, ch * rk)
                    sd[f"resblocks.{n}.{grp}.{l}.bias"] = _uniform(gen, (ch,), 1.0 / math.sqrt(ch * rk))
            for a in range(6):
                act_entries(sd, f"resblocks.{n}.activations.{a}", ch, gen, aa_layers[i])
    act_entries(sd, "activation_post", ch, gen, aa_post)
    _wn_conv(gen, sd, "conv_post", 1, ch, 7, ch * 7)
    sd["conv_post.bias"] = _uniform(gen, (1,), 1.0 / math.sqrt(ch * 7))
    return sd


def write_synthetic_checkpoints(out_dir: str, seed: int = 1, sharpen: float = 30.0, gain: float = 1.0,
                                vcfg: dict | None = None, force: bool = False):
    """Writes ``bvrnn_synth_s{seed}`` and ``vocoder_synth_s{seed}`` into out_dir; returns their paths."""
    os.makedirs(out_dir, exist_ok=True)
    tag = f"s{seed}_k{sharpen:g}_g{gain:g}"
    vtag = ""
    if vcfg and (any(vcfg.get("layers_antialias", [])) or vcfg.get("antialias_post", False)):
        vtag = "_aa" + "".join("1" if a else "0" for a in vcfg.get("layers_antialias", [False] * 4)) + \
               ("p" if vcfg.get("antialias_post", False) else "")
    p1 = os.path.join(out_dir, f"bvrnn_synth_{tag}")
    p2 = os.path.join(out_dir, f"vocoder_synth_{tag}{vtag}")
    if force or not os.path.exists(p1):
        tmp = p1 + f".tmp{os.getpid()}"
        torch.save({"vrnn": synth_bvrnn_state_dict(seed, sharpen=sharpen, gain=gain)}, tmp)
        os.replace(tmp, p1)
    if force or not os.path.exists(p2):
        tmp = p2 + f".tmp{os.getpid()}"
        torch.save({"generator": synth_vocoder_state_dict(seed + 1000, vcfg)}, tmp)
        os.replace(tmp, p2)
    return p1, p2

"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY.  Usage (from the repo root, where /root/reference exists):

    python -m oracle.gen_golden

The reference modules are imported from /root/reference through
``oracle/ref_shim.py`` and loaded (strict ``load_state_dict``, so the schema is
checked) with the deterministic synthetic checkpoints of
``bernoulli_var_speech_codec_b200/synth.py`` (the shipped checkpoints are LFS
pointers).  Stage taps are taken with forward hooks so the
reference source is not modified: pre-sigmoid logits from ``bvrnn.enc[4]``.

Fixtures (all float32 unless noted):
  synth_var_small.npz   var-bitrate, B=2, L=5000+odd, bitrates 3000 -> 35 bits
  synth_fix_small.npz   fixed 64-bit config
  synth_var_bits.npz    bit budget edge cases 0 / 1 / 64 / >64 on one short clip
  synth_var_aa_small.npz  configs/config_varBitRate_antialias.toml (anti-aliased activations in every stage and
                        before conv_post; checkpoint with the Activation1d schema): codes, dec_mel, wav
  aa_filter.npz         the reference's 12-tap Kaiser-sinc filter (alias_free_torch/filter.py:28-59), pins synth.py's
  synth_var_stochastic.npz  sampled bits with supplied uniforms + prior head + KL (`python -m oracle.gen_golden stochastic`)
  stim01_var.npz        BASELINE config #1 input (MUSHRA stim_01 ref.wav, CC BY 4.0,
                        resampled 24k->22.05k, peak-normalised) at 3000 bps
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bernoulli_var_speech_codec_b200.synth import write_synthetic_checkpoints  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CKPT_DIR = os.environ.get("BVC_CKPT_DIR", "/tmp/bvc_ckpts")
SEED, SHARPEN = 1, 30.0


def synth_audio(B, L, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(L, dtype=torch.float32) / 22050.0
    x = 0.1 * torch.randn(B, L, generator=g)
    for b in range(B):  # a few harmonics so the mel is not flat
        f0 = 110.0 * (b + 1)
        x[b] += 0.3 * torch.sin(2 * np.pi * f0 * t) * torch.sin(2 * np.pi * 1.5 * t) ** 2
        x[b] += 0.1 * torch.sin(2 * np.pi * 3.1 * f0 * t)
    return x.clamp(-1, 1)


def run_reference(ref, model, x, bitrate):
    taps = {}
    hook = model.bvrnn.enc[4].register_forward_hook(lambda m, i, o: taps.setdefault("logits", []).append(o.detach()))
    with torch.no_grad():
        xmel = ref.mel_spectrogram(x * ref.SCALING, n_fft=1024, num_mels=80, sampling_rate=22050, hop_size=256,
                                   win_size=1024, fmin=0, fmax=8000, padding_left=256).permute(0, 2, 1)
        codes = model.encode(x, bitrate)
        hook.remove()
        bits = np.round(bitrate * 256 / 22050) * torch.ones(xmel.shape[0], xmel.shape[1])
        _, all_h = model.bvrnn.encode(xmel, bits, torch.zeros(1, x.shape[0], 1024))
        dec_mel, h_fin = model.bvrnn.decode(codes, torch.zeros(1, x.shape[0], 1024))
        wav = model.decode(codes, x.shape[1])
        fwd = model(x, bitrate)
    assert torch.equal(wav, fwd)
    return dict(x=x, mel=xmel, codes=codes, logits=torch.stack(taps["logits"], 1), all_h=all_h,
                dec_mel=dec_mel, h_dec_final=h_fin[0], wav=wav)


def save(name, d, **extra):
    out = {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
    out.update(extra)
    np.savez_compressed(os.path.join(GOLD, name), **out)
    print("wrote", name, {k: getattr(v, "shape", v) for k, v in out.items()})


def gen_stochastic():
    """synth_var_stochastic.npz: the reference's sampled-bit path (bvrnn.py:86-160 with p_use_gen = 1, greedy = False), i.e.
    z = round(u - 0.5 + p) with the uniforms u supplied by us (torch.rand_like is patched for the duration of the call to hand
    out the fixture's uniforms frame by frame), per-frame bit budgets, the `prior` head and the KL term.  Taps by hooks:
    z_t = input of phi_z, p_t = output of enc, prior_t = output of prior."""
    ref = ref_shim.import_reference()
    p_b, p_v = write_synthetic_checkpoints(CKPT_DIR, seed=SEED, sharpen=SHARPEN)
    cfg_var = os.path.join(ROOT, "configs", "config_varBitRate.toml")
    m = ref.BVRNNCodecModel(cfg_var, p_b, p_v).eval()
    x = synth_audio(2, 4000 + 33, seed=21)
    g = torch.Generator().manual_seed(99)
    with torch.no_grad():
        mel = ref.mel_spectrogram(x * ref.SCALING, n_fft=1024, num_mels=80, sampling_rate=22050, hop_size=256,
                                  win_size=1024, fmin=0, fmax=8000, padding_left=256).permute(0, 2, 1).contiguous()
        B, T, _ = mel.shape
        u = torch.rand(B, T, 64, generator=g)
        bits = torch.randint(0, 66, (B, T), generator=g).float()
        taps = {"z": [], "p": [], "prior": []}
        hooks = [m.bvrnn.phi_z.register_forward_pre_hook(lambda mod, inp: taps["z"].append(inp[0].detach().clone())),
                 m.bvrnn.enc.register_forward_hook(lambda mod, i, o: taps["p"].append(o.detach().clone())),
                 m.bvrnn.prior.register_forward_hook(lambda mod, i, o: taps["prior"].append(o.detach().clone()))]
        frame = {"t": 0}
        orig = torch.rand_like

        def fake_rand_like(t, *a, **k):
            out = u[:, frame["t"]].to(t.dtype)
            frame["t"] += 1
            assert out.shape == t.shape
            return out

        torch.rand_like = fake_rand_like
        try:
            m.bvrnn.device = torch.device("cpu")
            dec, kld = m.bvrnn(mel, 1.0, False, bits)
        finally:
            torch.rand_like = orig
            for h in hooks:
                h.remove()
    assert frame["t"] == T
    save("synth_var_stochastic.npz",
         dict(x=x, mel=mel, uniforms=u, bits=bits, codes=torch.stack(taps["z"], 1), enc_p=torch.stack(taps["p"], 1),
              prior_p=torch.stack(taps["prior"], 1), dec_mel=dec, kld=kld.reshape(1)), ckpt_seed=SEED, ckpt_sharpen=SHARPEN)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "stochastic":
        gen_stochastic()
        return
    torch.manual_seed(0)
    ref = ref_shim.import_reference()
    os.makedirs(GOLD, exist_ok=True)
    p_b, p_v = write_synthetic_checkpoints(CKPT_DIR, seed=SEED, sharpen=SHARPEN)
    cfg_var = os.path.join(ROOT, "configs", "config_varBitRate.toml")
    cfg_fix = os.path.join(ROOT, "configs", "config_64bit.toml")
    meta = dict(ckpt_seed=SEED, ckpt_sharpen=SHARPEN)

    m_var = ref.BVRNNCodecModel(cfg_var, p_b, p_v).eval()
    x = synth_audio(2, 5000 + 77, seed=11)
    save("synth_var_small.npz", run_reference(ref, m_var, x, 3000), bitrate=3000, **meta)

    xb = synth_audio(1, 2600, seed=12)
    for br, tag in ((0, "0"), (86.2, "1"), (5512.5, "64"), (9000, "gt64")):
        r = run_reference(ref, m_var, xb, br)
        save(f"synth_var_bits_{tag}.npz", dict(x=r["x"], codes=r["codes"], wav=r["wav"]), bitrate=br, **meta)

    m_fix = ref.BVRNNCodecModel(cfg_fix, p_b, p_v).eval()
    save("synth_fix_small.npz", run_reference(ref, m_fix, x, 3000), bitrate=3000, **meta)

    # anti-aliased activations (dead under the shipped configs; switchable, models.py:82-88,189-190)
    import toml
    cfg_aa = os.path.join(ROOT, "configs", "config_varBitRate_antialias.toml")
    _, p_v_aa = write_synthetic_checkpoints(CKPT_DIR, seed=SEED, sharpen=SHARPEN, vcfg=toml.load(cfg_aa)["vocoder_config"])
    m_aa = ref.BVRNNCodecModel(cfg_aa, p_b, p_v_aa).eval()
    r = run_reference(ref, m_aa, x, 3000)
    save("synth_var_aa_small.npz", dict(x=r["x"], codes=r["codes"], dec_mel=r["dec_mel"], wav=r["wav"]), bitrate=3000, **meta)
    filt = sys.modules["third_party.BigVGAN.alias_free_torch.filter"].kaiser_sinc_filter1d(0.25, 0.3, 12)
    save("aa_filter.npz", dict(filter=filt.flatten()), cutoff=0.25, half_width=0.3, kernel_size=12)

    # BASELINE config #1 input (reference example.py:12-17; soundfile absent -> scipy.io.wavfile)
    import scipy.io.wavfile
    import scipy.signal
    fs, wav = scipy.io.wavfile.read(os.path.join(ref_shim.REFERENCE_ROOT,
                                                 "mushra_results_dataset/audio/stim_01/ref.wav"))
    sp = wav[:, 0].astype(np.float64) / 32768.0
    sp = scipy.signal.resample_poly(sp, 22050, fs)
    sp = sp / np.max(np.abs(sp))
    xs = torch.from_numpy(sp).float()[None, :]
    r = run_reference(ref, m_var, xs, 3000)
    r["all_h"] = r["all_h"][:, ::16].contiguous()   # keep the fixture small: every 16th frame
    save("stim01_var.npz", r, bitrate=3000, all_h_stride=16, **meta)


if __name__ == "__main__":
    main()

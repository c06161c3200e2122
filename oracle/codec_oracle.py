"""CPU oracle: fp32 restatement of the reference encode->decode path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may use it, as the checker or the timed CPU arm.

Parity status: the reference ships no tests, golden vectors or KATs for this
path -> "parity unpinned" by the reference itself.  This
restatement is pinned instead against outputs of the *unmodified reference
modules* run in the build container (``oracle/gen_golden.py`` imports them from
/root/reference through ``oracle/ref_shim.py`` and commits the vectors under
``tests/golden/``); ``tests/test_oracle_golden.py`` checks this file against
those vectors.  The one piece that cannot be pinned that way is the mel
filterbank, which the reference takes from librosa (absent): it is restated
from the published Slaney construction and is therefore "parity unpinned".

It is a restatement, not an import: weight-norm is folded once, the GRU cell
and the concatenated Linear layers are written out explicitly, and the log-mel
front end is built from an explicit reflect index + rFFT.  Each function cites
the reference lines it follows.  Plain PyTorch CPU ops in fp32.
"""
from __future__ import annotations

import math
import tomllib

import numpy as np
import torch
import torch.nn.functional as F

SCALING = 10 ** (-10 / 20)  # reference bvrnn_codec_model.py:17


# --------------------------------------------------------------------------
# mel filterbank: librosa.filters.mel(htk=False, norm='slaney') restated
# (reference call site third_party/BigVGAN/meldataset.py:68)
# --------------------------------------------------------------------------
def slaney_mel(sr, n_fft, n_mels, fmin, fmax):
    f_sp = 200.0 / 3.0
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0

    def h2m(f):
        return f / f_sp if f < min_log_hz else min_log_mel + math.log(f / min_log_hz) / logstep

    def m2h(m):
        return f_sp * m if m < min_log_mel else min_log_hz * math.exp(logstep * (m - min_log_mel))

    mels = np.linspace(h2m(float(fmin)), h2m(float(fmax)), n_mels + 2)
    hz = np.array([m2h(m) for m in mels], dtype=np.float64)
    fft_hz = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    out = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    for i in range(n_mels):
        lo = (fft_hz - hz[i]) / (hz[i + 1] - hz[i])
        hi = (hz[i + 2] - fft_hz) / (hz[i + 2] - hz[i + 1])
        out[i] = np.maximum(0.0, np.minimum(lo, hi))
    out *= (2.0 / (hz[2:] - hz[:-2]))[:, None]
    return out


# --------------------------------------------------------------------------
# log-mel front end (reference third_party/BigVGAN/meldataset.py:60-95, :38-39)
# --------------------------------------------------------------------------
def logmel(x: torch.Tensor, conf: dict) -> torch.Tensor:
    """x [B, L] (already scaled) -> log-mel [B, T, num_mels], T = L // hop."""
    n_fft, hop, win = conf["winsize"], conf["hopsize"], conf["winsize"]
    pl = conf["mel_pad_left"]
    pr = win - pl - hop                                   # meldataset.py:76-78
    B, L = x.shape
    if L <= max(pl, pr):
        raise RuntimeError("reflect padding needs L > %d" % max(pl, pr))
    idx = torch.arange(-pl, L + pr, device=x.device)
    idx = torch.where(idx < 0, -idx, idx)
    idx = torch.where(idx >= L, 2 * (L - 1) - idx, idx)    # reflect, no edge repeat (:80)
    xp = x[:, idx]
    T = 1 + (xp.shape[1] - n_fft) // hop
    frames = xp.unfold(1, n_fft, hop)[:, :T]               # center=False (:84-85)
    window = torch.hann_window(win, periodic=True, dtype=torch.float32, device=x.device)  # (:70)
    spec = torch.fft.rfft(frames * window, n=n_fft, dim=-1)
    mag = torch.sqrt(spec.real ** 2 + spec.imag ** 2 + 1e-9)             # (:86-87)
    basis = torch.from_numpy(slaney_mel(conf["fs"], n_fft, conf["num_mels"], conf["fmin"], conf["fmax"])).to(x.device)
    mel = torch.matmul(mag, basis.t())                                    # (:89)
    return torch.log(torch.clamp(mel, min=1e-5))                          # (:38-39)


# --------------------------------------------------------------------------
# BVRNN coder (reference bvrnn.py:44-83 modules, :163-209 encode, :211-229 decode)
# --------------------------------------------------------------------------
def _mlp(sd, name, idxs, v, last_act=True):
    for n, i in enumerate(idxs):
        v = F.linear(v, sd[f"{name}.{i}.weight"], sd[f"{name}.{i}.bias"])
        if last_act or n < len(idxs) - 1:
            v = F.elu(v)
    return v


def _gru_cell(sd, xin, h):
    """PyTorch GRU cell, gate order r,z,n (reference bvrnn.py:83,206)."""
    H = h.shape[1]
    gi = F.linear(xin, sd["rnn.weight_ih_l0"], sd["rnn.bias_ih_l0"])
    gh = F.linear(h, sd["rnn.weight_hh_l0"], sd["rnn.bias_hh_l0"])
    r = torch.sigmoid(gi[:, :H] + gh[:, :H])
    z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    return (1.0 - z) * n + z * h


def bvrnn_encode(sd, y, bits, h0, var_bit, want_taps=False, uniforms=None):
    """y [B,T,80], bits [B,T] (bits per frame), h0 [B,H].

    Returns codes [B,T,Z] in {0,1,0.5}, all_h [B,T,H] (state entering frame t),
    final h [B,H], and (optionally) pre-sigmoid logits [B,T,Z].
    uniforms [B,T,Z] (optional): sampled bits z = round(u - 0.5 + p) instead of round(p)
    (reference bvrnn.py:123-126 with the uniforms supplied by the caller; encoder and decoder state stay in lock-step
    exactly as in BVRNN.encode, i.e. the reference's forward with p_use_gen = 1).
    """
    B, T, _ = y.shape
    Z = sd["enc.4.weight"].shape[0]
    mean, std = sd["mean_mel"], sd["std_mel"]
    yn = (y - mean) / std                                            # bvrnn.py:173
    phi_x_all = _mlp(sd, "phi_x", (0, 2, 4), yn)                     # :178
    bit_idx = torch.arange(Z, device=y.device)
    h = h0.clone()
    codes, hs, logits = [], [], []
    for t in range(T):
        u = torch.cat([phi_x_all[:, t], h], 1)                       # :189
        e = F.elu(F.linear(u, sd["enc.0.weight"], sd["enc.0.bias"]))
        e = F.elu(F.linear(e, sd["enc.2.weight"], sd["enc.2.bias"]))
        logit = F.linear(e, sd["enc.4.weight"], sd["enc.4.bias"])
        if uniforms is None:
            z = torch.round(torch.sigmoid(logit))                    # :191 (round half to even)
        else:
            z = torch.round(uniforms[:, t] - 0.5 + torch.sigmoid(logit))   # :126
        if var_bit:                                                  # :193-194
            m = (bits[:, t, None] > bit_idx[None, :]).float()
            z = z * m + 0.5 * (1.0 - m)
        codes.append(z)
        if want_taps:
            logits.append(logit)
        pz = _mlp(sd, "phi_z", (0, 2, 4), z)                          # :198
        d = _mlp(sd, "dec", (0, 2, 4, 6), torch.cat([pz, h], 1), last_act=False)   # :202
        px = _mlp(sd, "phi_x", (0, 2, 4), (d - mean) / std)          # :204
        hs.append(h)                                                 # :205 (before the update)
        h = _gru_cell(sd, torch.cat([px, pz], 1), h)                 # :206
    out = (torch.stack(codes, 1), torch.stack(hs, 1), h)
    if want_taps:
        out = out + (torch.stack(logits, 1),)
    return out


def bvrnn_prior(sd, all_h):
    """prior(h_t) for every frame: all_h [B,T,H] (state entering frame t) -> Bernoulli probabilities [B,T,Z]
    (reference bvrnn.py:68-73, evaluated at :115-120)."""
    return torch.sigmoid(_mlp(sd, "prior", (0, 2, 4), all_h, last_act=False))


def bvrnn_kld(enc_p, prior_p, bits, var_bit):
    """KL(enc || prior) per the reference's training forward (bvrnn.py:148-158): per-frame sum over the active bits, mean
    over the batch, then mean over frames.  enc_p / prior_p [B,T,Z] probabilities, bits [B,T]."""
    e = enc_p * (torch.log(torch.clip(enc_p, 1e-3)) - torch.log(torch.clip(prior_p, 1e-3))) + \
        (1 - enc_p) * (torch.log(torch.clip(1 - enc_p, 1e-3)) - torch.log(torch.clip(1 - prior_p, 1e-3)))
    if var_bit:
        mask = (bits[:, :, None] > torch.arange(enc_p.shape[-1], device=enc_p.device)[None, None, :]).float()
        e = e * mask
    return e.sum(-1).mean(0).mean()


def bvrnn_decode(sd, z, h0):
    """z [B,T,Z] (any floats), h0 [B,H] -> mel [B,T,80], final h [B,H]  (bvrnn.py:211-229)."""
    mean, std = sd["mean_mel"], sd["std_mel"]
    h = h0.clone()
    mels = []
    for t in range(z.shape[1]):
        pz = _mlp(sd, "phi_z", (0, 2, 4), z[:, t])
        d = _mlp(sd, "dec", (0, 2, 4, 6), torch.cat([pz, h], 1), last_act=False)
        mels.append(d)
        px = _mlp(sd, "phi_x", (0, 2, 4), (d - mean) / std)
        h = _gru_cell(sd, torch.cat([px, pz], 1), h)
    return torch.stack(mels, 1), h


# --------------------------------------------------------------------------
# causal BigVGAN-tiny (reference third_party/BigVGAN/models.py:103-121, :207-238;
# SnakeBeta activations.py:107-120; weight_norm folded, models.py:47-62,140,164,200)
# --------------------------------------------------------------------------
def fold_weight_norm(g, v):
    """w = g * v / ||v||, norm over all dims but 0 (old-style torch weight_norm, dim=0)."""
    n = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1)))
    return v * (g / n)


def _snakebeta(x, alpha, beta):
    a = torch.exp(alpha)[None, :, None]
    b = torch.exp(beta)[None, :, None]
    return x + (1.0 / (b + 1e-9)) * torch.sin(x * a) ** 2


def _aa_snakebeta(x, alpha, beta, f_up, f_down):
    """Anti-aliased activation (reference alias_free_torch/act.py:23-28): 2x FIR up-sampling (resample.py:24-32:
    replicate pad 5|5, transposed conv with the 12-tap Kaiser-sinc filter, x2 gain, crop 15|15), SnakeBeta at the doubled
    rate, 2x FIR down-sampling (resample.py:46-48 -> filter.py:88-96: replicate pad 5|6, strided depthwise conv).
    The filters are the checkpoint's registered buffers, as in the reference."""
    C = x.shape[1]
    k = f_up.shape[-1]
    pad = k // 2 - 1
    pad_l = pad * 2 + (k - 2) // 2
    pad_r = pad * 2 + (k - 2 + 1) // 2
    u = F.pad(x, (pad, pad), mode="replicate")
    u = 2 * F.conv_transpose1d(u, f_up.expand(C, -1, -1), stride=2, groups=C)
    u = u[..., pad_l:-pad_r]
    a = _snakebeta(u, alpha, beta)
    kd = f_down.shape[-1]
    a = F.pad(a, (kd // 2 - int(kd % 2 == 0), kd // 2), mode="replicate")
    return F.conv1d(a, f_down.expand(C, -1, -1), stride=2, groups=C)


def _cconv(x, w, b, dilation=1):
    k = w.shape[2]
    return F.conv1d(F.pad(x, ((k - 1) * dilation, 0)), w, b, dilation=dilation)


class VocoderOracle:
    def __init__(self, sd, vcfg):
        if vcfg.get("activation", "snakebeta") != "snakebeta" or any(vcfg.get("layers_sym", [])) \
                or vcfg.get("pre_sym", False) or vcfg.get("post_sym", False):
            raise NotImplementedError("oracle covers the causal snakebeta configurations only")
        self.aa = [bool(a) for a in vcfg.get("layers_antialias", [False] * len(vcfg["upsample_rates"]))]
        self.aa_post = bool(vcfg.get("antialias_post", False))
        self.rates = list(vcfg["upsample_rates"])
        self.rks = list(vcfg["resblock_kernel_sizes"])
        self.dil = [list(d) for d in vcfg["resblock_dilation_sizes"]]
        f = lambda n: fold_weight_norm(sd[n + ".weight_g"], sd[n + ".weight_v"])
        self.pre = (f("conv_pre"), sd["conv_pre.bias"])
        self.ups = [(f(f"ups.{i}.1"), sd[f"ups.{i}.1.bias"]) for i in range(len(self.rates))]
        self.blocks = []
        for n in range(len(self.rates) * len(self.rks)):
            c1 = [(f(f"resblocks.{n}.convs1.{l}"), sd[f"resblocks.{n}.convs1.{l}.bias"]) for l in range(3)]
            c2 = [(f(f"resblocks.{n}.convs2.{l}"), sd[f"resblocks.{n}.convs2.{l}.bias"]) for l in range(3)]
            act = [self._act(sd, f"resblocks.{n}.activations.{a}", self.aa[n // len(self.rks)]) for a in range(6)]
            self.blocks.append((c1, c2, act))
        self.act_post = self._act(sd, "activation_post", self.aa_post)
        self.post = (f("conv_post"), sd["conv_post.bias"])

    @staticmethod
    def _act(sd, name, aa):
        """-> callable activation; Activation1d wraps SnakeBeta as `.act` and owns two filter buffers (models.py:66-88)."""
        if not aa:
            alpha, beta = sd[name + ".alpha"], sd[name + ".beta"]
            return lambda x: _snakebeta(x, alpha, beta)
        alpha, beta = sd[name + ".act.alpha"], sd[name + ".act.beta"]
        f_up, f_down = sd[name + ".upsample.filter"], sd[name + ".downsample.lowpass.filter"]
        return lambda x: _aa_snakebeta(x, alpha, beta, f_up, f_down)

    def _amp(self, n, x):
        c1, c2, act = self.blocks[n]
        d = self.dil[n % len(self.rks)]
        for l in range(3):                                            # models.py:103-121
            xt = _cconv(act[2 * l](x), *c1[l], dilation=d[l])
            xt = _cconv(act[2 * l + 1](xt), *c2[l])
            x = xt + x
        return x

    def __call__(self, mel, length, taps=None):
        """mel [B,80,T] -> [B,1,min(length, 256T+294)]."""
        x = _cconv(mel, *self.pre)                                    # models.py:209-213
        if taps is not None:
            taps["pre"] = x
        nk = len(self.rks)
        for i, u in enumerate(self.rates):
            w, b = self.ups[i]
            x = F.conv_transpose1d(x, w, b, stride=u)                 # :216-217
            if taps is not None:
                taps[f"up{i}"] = x
            acc = self._amp(i * nk, x)
            for j in range(1, nk):
                acc = acc + self._amp(i * nk + j, x)
            x = acc / nk                                              # :219-225
            if taps is not None:
                taps[f"stage{i}"] = x
        x = self.act_post(x)
        x = torch.tanh(_cconv(x, *self.post))                         # :228-236
        return x[:, :, :length]                                       # :238


# --------------------------------------------------------------------------
# facade (reference bvrnn_codec_model.py:19-76)
# --------------------------------------------------------------------------
class OracleCodec:
    def __init__(self, config_path, bvrnn_chkpt_path, vocoder_chkpt_path, device="cpu"):
        """device: "cpu" (the oracle proper) or a CUDA device -- the same restatement run through PyTorch eager on the GPU
        (cuBLAS / cuDNN / cuFFT library kernels), which bench.py times as the "PyTorch eager on B200" bar of SURVEY.md 8d."""
        with open(config_path, "rb") as fh:
            self.conf = tomllib.load(fh)
        self.device = torch.device(device)
        self.sd = {k: v.float().to(self.device) for k, v in
                   torch.load(bvrnn_chkpt_path, map_location="cpu", weights_only=True)["vrnn"].items()}
        gsd = {k: v.float().to(self.device) for k, v in
               torch.load(vocoder_chkpt_path, map_location="cpu", weights_only=True)["generator"].items()}
        self.vocoder = VocoderOracle(gsd, self.conf["vocoder_config"])
        self.h_dim = self.conf["h_dim"]
        self.var_bit = bool(self.conf["var_bit"])

    def bits_per_frame(self, bitrate):
        return float(np.round(bitrate * self.conf["hopsize"] / self.conf["fs"]))   # codec_model.py:58

    @torch.no_grad()
    def logmel(self, x):
        return logmel(x.float() * SCALING, self.conf)

    @torch.no_grad()
    def encode(self, x, bitrate, taps=None):
        mel = self.logmel(x)
        B, T, _ = mel.shape
        bits = torch.full((B, T), self.bits_per_frame(bitrate), device=mel.device)
        h0 = torch.zeros(B, self.h_dim, device=mel.device)
        out = bvrnn_encode(self.sd, mel, bits, h0, self.var_bit, want_taps=taps is not None)
        if taps is not None:
            taps["mel"], taps["all_h"], taps["h_final"], taps["logits"] = mel, out[1], out[2], out[3]
        return out[0]

    @torch.no_grad()
    def decode_mel(self, codes, h0=None):
        h0 = torch.zeros(codes.shape[0], self.h_dim, device=codes.device) if h0 is None else h0
        return bvrnn_decode(self.sd, codes.float(), h0)

    @torch.no_grad()
    def vocode(self, mel, length, taps=None):
        """mel [B,T,80] -> wav [B, length] (includes the final /SCALING, codec_model.py:71)."""
        return self.vocoder(mel.permute(0, 2, 1), length, taps).squeeze(1) / SCALING

    @torch.no_grad()
    def decode(self, codes, length, taps=None):
        mel, _ = self.decode_mel(codes)
        if taps is not None:
            taps["dec_mel"] = mel
        return self.vocode(mel, length, taps)

    @torch.no_grad()
    def forward(self, x, bitrate):
        return self.decode(self.encode(x, bitrate), x.shape[1])

    __call__ = forward

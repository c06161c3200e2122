"""Host-side mirror of the reference's codec facade, running on libbvc (CUDA, sm_100a).

Mirrors reference ``bvrnn_codec_model.py:19-76`` (``BVRNNCodecModel``: same constructor
arguments, ``encode`` / ``decode`` / ``forward`` signatures, tensor conventions and error
behaviour) and keeps the inner operators callable with the reference's signatures:
``model.bvrnn.encode(y, varBitrate, h)`` / ``.decode(z, h)`` (bvrnn.py:163,211),
``model.vocoder(mel, length)`` (third_party/BigVGAN/models.py:207) and
``mel_spectrogram(...)`` (third_party/BigVGAN/meldataset.py:60).

PyTorch is used for checkpoint unpickling, device memory and streams only; all arithmetic
runs in libbvc.so through the C ABI of ``include/bvc.h``.  There is no CPU implementation:
CPU tensors are copied to the handle's GPU, processed there and copied back
(``bvc_encode_host`` / ``bvc_decode_host``).
"""
from __future__ import annotations

import ctypes as C
import os
import tomllib
import weakref

import numpy as np
import torch
from torch import nn

from . import _lib
from .melbasis import slaney_mel_basis

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

default_config = os.path.join(root, "configs", "config_varBitRate.toml")
default_chkpt_bvrnn = os.path.join(root, "chkpts", "bvrnn_var_bitrate_step200000")
default_chkpt_vocoder = os.path.join(root, "chkpts", "bigvgan_causal_tiny_ftbvrnn_g_step3500000")

SCALING = 10 ** (-10 / 20)   # reference bvrnn_codec_model.py:17


class AttrDict(dict):
    """dict with attribute access (reference third_party/BigVGAN/env.py:8-11)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _PinnedPool:
    """Page-locked result buffers for the host-tensor path.

    Pinning is expensive (37 ms for the 226 MB waveform batch of the benchmark), so blocks are recycled.  A block is
    handed out as a tensor made with ``torch.frombuffer`` over a ctypes array at the block's address; the storage keeps
    the array alive, so the finalizer of the array fires exactly when the LAST tensor or view sharing the memory has
    died, and only then does the block go back to the free list.  Callers therefore get fresh, exclusively owned
    tensors, like from the reference.
    """

    def __init__(self, lib, max_free_per_size=4):
        self.lib, self.free, self.max_free = lib, {}, max_free_per_size

    def _release(self, ptr, nbytes):
        lst = self.free.setdefault(nbytes, [])
        if len(lst) < self.max_free:
            lst.append(ptr)
        else:
            self.lib.bvc_host_free(C.c_void_p(ptr))

    def empty(self, shape):
        n = int(np.prod(shape))
        if n == 0:
            return torch.empty(*shape, dtype=torch.float32)
        nbytes = ((n * 4 + (1 << 20) - 1) >> 20) << 20          # 1 MiB size classes
        lst = self.free.get(nbytes)
        if lst:
            ptr = lst.pop()
        else:
            p = C.c_void_p()
            _lib.check(self.lib.bvc_host_alloc(C.byref(p), nbytes), "host_alloc")
            ptr = p.value
        arr = (C.c_char * nbytes).from_address(ptr)
        weakref.finalize(arr, self._release, ptr, nbytes)
        return torch.frombuffer(arr, dtype=torch.float32, count=n).view(*shape)


class _Engine:
    """Owns one bvc_handle (one per model instance and device)."""

    def __init__(self, conf: dict, device: torch.device):
        self.lib = _lib.load()
        self.pinned = _PinnedPool(self.lib)
        if self.lib.bvc_abi_version() != 3:
            raise RuntimeError("libbvc ABI version mismatch")
        v = conf["vocoder_config"]
        if v.get("activation", "snakebeta") != "snakebeta" or not v.get("snake_logscale", True):
            raise NotImplementedError("only the shipped log-scale snakebeta activation is implemented")
        if any(v.get("layers_sym", [])):
            raise NotImplementedError("vocoder_config.layers_sym=true is not implemented (shipped configs use false)")
        for key in ("pre_sym", "post_sym"):
            if v.get(key, False):
                raise NotImplementedError(f"vocoder_config.{key}=true is not implemented (shipped configs use false)")
        if str(v.get("resblock", "1")) != "1":
            raise ValueError("Wrong resblock")
        dil = v["resblock_dilation_sizes"]
        if any(list(d) != list(dil[0]) for d in dil) or len(dil[0]) != 3:
            raise NotImplementedError("resblock_dilation_sizes must be the same 3 dilations for every kernel size")
        if conf["winsize"] != 1024:
            raise NotImplementedError("winsize must be 1024")
        cfg = _lib.BvcConfig()
        cfg.device = device.index
        cfg.x_dim, cfg.h_dim, cfg.z_dim = conf["num_mels"], conf["h_dim"], conf["z_dim"]
        cfg.var_bit = int(bool(conf["var_bit"]))
        cfg.n_fft, cfg.hop, cfg.pad_left = conf["winsize"], conf["hopsize"], conf["mel_pad_left"]
        cfg.voc_initial_channel = v["upsample_initial_channel"]
        cfg.voc_num_stages = len(v["upsample_rates"])
        if cfg.voc_num_stages != 4 or len(v["resblock_kernel_sizes"]) != 3:
            raise NotImplementedError("vocoder must have 4 upsampling stages and 3 resblock kernel sizes")
        cfg.voc_up_rates = (C.c_int32 * 4)(*v["upsample_rates"])
        cfg.voc_up_kernels = (C.c_int32 * 4)(*v["upsample_kernel_sizes"])
        cfg.voc_num_kernels = 3
        cfg.voc_res_kernels = (C.c_int32 * 3)(*v["resblock_kernel_sizes"])
        cfg.voc_res_dilations = (C.c_int32 * 3)(*dil[0])
        cfg.voc_antialias = (C.c_int32 * 4)(*[int(bool(a)) for a in v.get("layers_antialias", [False] * 4)])
        cfg.voc_antialias_post = int(bool(v.get("antialias_post", False)))
        self.handle = C.c_void_p()
        _lib.check(self.lib.bvc_create(C.byref(self.handle), C.byref(cfg)), "bvc_create")
        self.device = device
        self.conf = conf
        self.X, self.H, self.Z = cfg.x_dim, cfg.h_dim, cfg.z_dim
        self.hop = cfg.hop
        # front-end constant tables (meldataset.py:67-70)
        window = torch.hann_window(conf["winsize"], periodic=True, dtype=torch.float32).contiguous()
        basis = np.ascontiguousarray(slaney_mel_basis(conf["fs"], conf["winsize"], conf["num_mels"],
                                                      conf["fmin"], conf["fmax"]), dtype=np.float32)
        _lib.check(self.lib.bvc_set_frontend(self.handle, _ptr(window), C.c_void_p(basis.ctypes.data)),
                   "bvc_set_frontend")

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.bvc_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _load(self, fn, state_dict, what):
        keep, arr = [], (_lib.BvcTensor * len(state_dict))()
        for i, (name, t) in enumerate(state_dict.items()):
            if not isinstance(t, torch.Tensor):
                raise RuntimeError(f"{what}: state_dict entry {name!r} is not a tensor")
            t = t.detach().to("cpu", torch.float32).contiguous()
            if t.dim() == 0:
                t = t.reshape(1)
            if t.dim() > 4:
                raise RuntimeError(f"{what}: unexpected rank for {name}")
            keep.append(t)
            arr[i].name = name.encode()
            arr[i].data = t.data_ptr()
            arr[i].ndim = t.dim()
            for d in range(t.dim()):
                arr[i].shape[d] = t.shape[d]
        _lib.check(fn(self.handle, arr, len(state_dict)), what)

    def load_bvrnn(self, sd):
        self._load(self.lib.bvc_load_bvrnn, sd, "load_state_dict(vrnn)")

    def load_vocoder(self, sd):
        self._load(self.lib.bvc_load_vocoder, sd, "load_state_dict(generator)")

    # ---- device entry points -------------------------------------------------
    def _dev(self, t, name):
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name} must be a torch.Tensor")
        if t.device != self.device:
            raise RuntimeError(f"{name} is on {t.device}, the codec runs on {self.device}")
        return t.to(torch.float32).contiguous()

    def logmel(self, x, scale):
        x = self._dev(x, "x")
        B, L = x.shape
        mel = torch.empty(B, L // self.hop, self.X, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bvc_logmel(self.handle, _ptr(x), B, L, scale, _ptr(mel), _stream(self.device)),
                       "mel_spectrogram")
        return mel

    def encode(self, mel, bits, bits_scalar, h0, want_logits=False, want_all_h=True, want_packed=False, want_mel=False,
               uniforms=None, want_prior=False):
        """-> (codes, all_h, h_final, logits, packed[, mel_hat][, prior]); mel_hat (want_mel) is the decoder's mel the encoder
        forms inside its loop (bvrnn.py:198-206) = BVRNN.decode(codes, h0)[0]; uniforms [B,T,Z]: sampled bits
        round(u - 0.5 + p) (bvrnn.py:126); prior (want_prior): the prior head's probabilities [B,T,Z] (bvrnn.py:68-73)."""
        mel = self._dev(mel, "y")
        B, T, _ = mel.shape
        dev = self.device
        codes = torch.empty(B, T, self.Z, device=dev, dtype=torch.float32)
        all_h = torch.empty(B, T, self.H, device=dev, dtype=torch.float32) if want_all_h else None
        logits = torch.empty(B, T, self.Z, device=dev, dtype=torch.float32) if want_logits else None
        packed = torch.empty(B, T, device=dev, dtype=torch.int64) if want_packed else None
        h_fin = torch.empty(B, self.H, device=dev, dtype=torch.float32)
        bits_t = self._dev(bits, "varBitrate") if bits is not None else None
        h0_t = self._dev(h0, "h") if h0 is not None else None
        mel_hat = torch.empty(B, T, self.X, device=dev, dtype=torch.float32) if want_mel else None
        prior = torch.empty(B, T, self.Z, device=dev, dtype=torch.float32) if want_prior else None
        u_t = self._dev(uniforms, "uniforms") if uniforms is not None else None
        if u_t is not None and tuple(u_t.shape) != (B, T, self.Z):
            raise ValueError(f"uniforms must have shape {(B, T, self.Z)}")
        with torch.cuda.device(dev):
            _lib.check(self.lib.bvc_encode_ex(self.handle, _ptr(mel), _ptr(bits_t), float(bits_scalar), _ptr(h0_t), _ptr(u_t),
                                              B, T, _ptr(codes), _ptr(packed), _ptr(logits), _ptr(all_h), _ptr(h_fin),
                                              _ptr(mel_hat), _ptr(prior), _stream(dev)), "BVRNN.encode")
        out = (codes, all_h, h_fin, logits, packed)
        if want_mel:
            out = out + (mel_hat,)
        if want_prior:
            out = out + (prior,)
        return out

    def unpack_codes(self, packed, bits, bits_scalar):
        """packed int64 [B, T] (bit i = code i) -> float codes [B, T, Z] with 0.5 for masked bits."""
        B, T = packed.shape
        packed = packed.contiguous()
        codes = torch.empty(B, T, self.Z, device=packed.device, dtype=torch.float32)
        _lib.check(self.lib.bvc_unpack_codes(self.handle, _ptr(packed), _ptr(bits), float(bits_scalar), B, T, _ptr(codes),
                                             _stream(self.device)), "unpack_codes")
        return codes

    def decode_mel_packed(self, packed, bits, bits_scalar, h0):
        """BVRNN.decode straight from the wire words (bvc_decode_packed): packed int64 [B, T], budgets [B, T] or a scalar."""
        packed = packed.contiguous()
        B, T = packed.shape
        mel = torch.empty(B, T, self.X, device=self.device, dtype=torch.float32)
        h_fin = torch.empty(B, self.H, device=self.device, dtype=torch.float32)
        bits_t = self._dev(bits, "bits") if bits is not None else None
        h0_t = self._dev(h0, "h") if h0 is not None else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bvc_decode_packed(self.handle, _ptr(packed), _ptr(bits_t), float(bits_scalar), _ptr(h0_t), B, T,
                                                  _ptr(mel), _ptr(h_fin), _stream(self.device)), "BVRNN.decode (packed)")
        return mel, h_fin

    def pack_bitstream(self, packed, bits, bits_scalar):
        """words int64 [B, T] -> uint8 [B, stride] bit-streams (header + n bits per frame), see include/bvc.h."""
        packed = packed.contiguous()
        B, T = packed.shape
        bits_t = self._dev(bits, "bits") if bits is not None else None
        stride = int(self.lib.bvc_bitstream_bytes(self.handle, T, 1 if bits_t is not None else 0))
        out = torch.empty(B, stride, device=self.device, dtype=torch.uint8)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bvc_pack_bitstream(self.handle, _ptr(packed), _ptr(bits_t), float(bits_scalar), B, T, _ptr(out),
                                                   stride, _stream(self.device)), "pack_bitstream")
        return out

    def unpack_bitstream(self, stream, T):
        """uint8 [B, stride] -> (words int64 [B, T], budgets float32 [B, T])."""
        stream = stream.contiguous()
        B, stride = stream.shape
        packed = torch.empty(B, T, device=self.device, dtype=torch.int64)
        bits = torch.empty(B, T, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bvc_unpack_bitstream(self.handle, _ptr(stream), stride, B, T, _ptr(packed), _ptr(bits),
                                                     _stream(self.device)), "unpack_bitstream")
        return packed, bits

    def decode_mel(self, codes, h0):
        codes = self._dev(codes, "z")
        B, T, _ = codes.shape
        mel = torch.empty(B, T, self.X, device=self.device, dtype=torch.float32)
        h_fin = torch.empty(B, self.H, device=self.device, dtype=torch.float32)
        h0_t = self._dev(h0, "h") if h0 is not None else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bvc_decode_mel(self.handle, _ptr(codes), _ptr(h0_t), B, T, _ptr(mel), _ptr(h_fin),
                                               _stream(self.device)), "BVRNN.decode")
        return mel, h_fin

    def vocode(self, mel_cl, length, inv_scale_div):
        mel_cl = self._dev(mel_cl, "mel")
        B, T, _ = mel_cl.shape
        n = min(int(length), int(self.lib.bvc_vocoder_out_len(self.handle, T)))
        wav = torch.empty(B, max(n, 0), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bvc_vocode(self.handle, _ptr(mel_cl), B, T, int(length), float(inv_scale_div),
                                           _ptr(wav), _stream(self.device)), "BigVGAN.forward")
        return wav

    # ---- host entry points (H2D + compute + D2H inside the call) ---------------
    def encode_host(self, x, scale, bits_scalar):
        x = x.to(torch.float32).contiguous()
        B, L = x.shape
        codes = self.pinned.empty((B, L // self.hop, self.Z))
        _lib.check(self.lib.bvc_encode_host(self.handle, _ptr(x), B, L, scale, float(bits_scalar), _ptr(codes)),
                   "encode")
        return codes

    def decode_host(self, codes, length, inv_scale_div):
        codes = codes.to(torch.float32).contiguous()
        B, T, _ = codes.shape
        n = min(int(length), int(self.lib.bvc_vocoder_out_len(self.handle, T)))
        wav = self.pinned.empty((B, max(n, 0)))
        _lib.check(self.lib.bvc_decode_host(self.handle, _ptr(codes), B, T, int(length), float(inv_scale_div),
                                            _ptr(wav)), "decode")
        return wav

    def debug_read(self, name, shape):
        out = torch.empty(*shape, dtype=torch.float32)
        _lib.check(self.lib.bvc_debug_read(self.handle, name.encode(), _ptr(out), out.numel()), "debug_read")
        return out

    def kernel_launches(self):
        return int(self.lib.bvc_kernel_launches(self.handle))

    def last_recurrent_ms(self):
        """Device time of the persistent recurrent kernel of the last encode / decode_mel call (CUDA events); waits for it."""
        return float(self.lib.bvc_last_recurrent_ms(self.handle))

    def recurrent_ms(self, kind, age=0):
        """Device time of the recurrent kernel of the age-th latest call of a kind (0 encode, 1 decode); -1 if not in the ring."""
        return float(self.lib.bvc_recurrent_ms(self.handle, int(kind), int(age)))

    def check(self):
        """Waits for the recurrent-kernel launches enqueued so far and raises if one of them aborted (bvc_check)."""
        _lib.check(self.lib.bvc_check(self.handle), "bvc_check")

    def set_precision(self, mode):
        _lib.check(self.lib.bvc_set_precision(self.handle, int(mode)), "set_precision")


class _BVRNN(nn.Module):
    """Operator-level mirror of reference ``BVRNN`` inference entry points (bvrnn.py:163,211)."""

    def __init__(self, engine, h_dim, z_dim, x_dim, var_bit):
        super().__init__()
        self._engine = engine
        self.h_dim, self.z_dim, self.x_dim, self.varBit = h_dim, z_dim, x_dim, var_bit

    def encode(self, y, varBitrate, h):
        """y (B,T,x_dim), varBitrate (B,T) bits per frame, h (1,B,h_dim) -> (z (B,T,z_dim), all_h (B,T,h_dim))."""
        h0 = h[-1] if h is not None else None
        codes, all_h, _, _, _ = self._engine.encode(y, varBitrate, 0.0, h0)
        return codes, all_h

    def encode_sampled(self, y, varBitrate, h, uniforms, want_prior=True):
        """The sampled-bit variant of encode (reference BVRNN.forward with p_use_gen = 1, greedy = False, bvrnn.py:86-160, with
        the uniforms of its torch.rand_like supplied by the caller): z_t = round(u_t - 0.5 + p_t).
        -> dict(z, all_h, p = sigmoid(logits) [B,T,z_dim], prior [B,T,z_dim] or None)."""
        h0 = h[-1] if h is not None else None
        out = self._engine.encode(y, varBitrate, 0.0, h0, want_logits=True, uniforms=uniforms, want_prior=want_prior)
        return dict(z=out[0], all_h=out[1], p=torch.sigmoid(out[3]), prior=out[5] if want_prior else None)

    def kld(self, p, prior, varBitrate):
        """KL(enc || prior) of the reference's training forward (bvrnn.py:148-158): host-side glue over the kernel outputs."""
        e = p * (torch.log(torch.clip(p, 1e-3)) - torch.log(torch.clip(prior, 1e-3))) + \
            (1 - p) * (torch.log(torch.clip(1 - p, 1e-3)) - torch.log(torch.clip(1 - prior, 1e-3)))
        if self.varBit:
            e = e * (varBitrate[:, :, None] > torch.arange(self.z_dim, device=p.device)[None, None, :]).float()
        return e.sum(-1).mean(0).mean()

    def decode(self, z, h):
        """z (B,T,z_dim), h (1,B,h_dim) -> (mel (B,T,x_dim), h (1,B,h_dim))."""
        mel, h_fin = self._engine.decode_mel(z, h[-1] if h is not None else None)
        return mel, h_fin[None]


class _Vocoder(nn.Module):
    """Mirror of reference ``BigVGAN.forward(x, length)`` (third_party/BigVGAN/models.py:207-238)."""

    def __init__(self, engine, h):
        super().__init__()
        self._engine = engine
        self.h = h

    def forward(self, x, length):
        """x (B, num_mels, T) -> (B, 1, min(length, 256 T + 294))."""
        return self._engine.vocode(x.permute(0, 2, 1), length, 1.0)[:, None, :]


class BVRNNCodecModel(nn.Module):
    def __init__(self, config_path=default_config, bvrnn_chkpt_path=default_chkpt_bvrnn,
                 vocoder_chkpt_path=default_chkpt_vocoder, *, device=None):
        '''
        config_path: path to the toml config file
        bvrnn_chkpt_path: path to the checkpoint of the BVRNN model
        vocoder_chkpt_path: path to the checkpoint of the vocoder model
        device (keyword-only extension): CUDA device the codec runs on (default: current CUDA device)
        '''
        super().__init__()
        with open(config_path, "rb") as fh:
            conf = tomllib.load(fh)
        self.conf = conf
        if not torch.cuda.is_available():
            raise RuntimeError("BVRNNCodecModel (b200) needs a CUDA sm_100 device: there is no CPU path")
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("BVRNNCodecModel (b200) runs on CUDA devices only")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self._engine = _Engine(conf, device)

        bvrnn_chkpt = torch.load(bvrnn_chkpt_path, map_location=torch.device('cpu'), weights_only=True)
        vocoder_chkpt = torch.load(vocoder_chkpt_path, map_location=torch.device('cpu'), weights_only=True)
        self._engine.load_bvrnn(bvrnn_chkpt['vrnn'])
        self._engine.load_vocoder(vocoder_chkpt['generator'])

        self.bvrnn = _BVRNN(self._engine, conf['h_dim'], conf['z_dim'], conf['num_mels'], bool(conf['var_bit']))
        self.vocoder = _Vocoder(self._engine, AttrDict(conf['vocoder_config']))

    @property
    def device(self):
        return self._engine.device

    def bits_per_frame(self, bitrate):
        return float(np.round(bitrate * self.conf['hopsize'] / self.conf['fs']))   # reference :58

    def encode(self, x, bitrate, *, uniforms=None):
        '''
        x: input waveform, shape (batch, length)
        bitrate: target bitrate in bits per second, will be rounded to the nearest valid bitrate
        uniforms (keyword-only extension): (batch, frames, z_dim) uniforms in [0, 1) -> sampled bits round(u - 0.5 + p)
                 (the reference's stochastic mode, bvrnn.py:123-126) instead of the greedy round(p)
        '''
        bits = self.bits_per_frame(bitrate)
        if x.device.type == "cpu" and uniforms is None:
            return self._engine.encode_host(x, SCALING, bits)
        on_cpu = x.device.type == "cpu"
        mel = self._engine.logmel(x.to(self.device), SCALING)
        u = uniforms.to(self.device, torch.float32).contiguous() if uniforms is not None else None
        codes = self._engine.encode(mel, None, bits, None, want_all_h=False, uniforms=u)[0]
        return codes.cpu() if on_cpu else codes

    def decode(self, codes, length):
        '''
        codes: latent binary codes, shape (batch, frames, z_dim)
        length: length of the output waveform
        '''
        if codes.device.type == "cpu":
            return self._engine.decode_host(codes, length, SCALING)
        mel, _ = self._engine.decode_mel(codes, None)
        return self._engine.vocode(mel, length, SCALING)

    def forward(self, x, bitrate):
        """decode(encode(x, bitrate), x.shape[1]) (reference :73-76) with ONE recurrence: the encoder already forms the
        decoder's mel for these codes inside its loop (analysis by synthesis, bvrnn.py:198-206 vs 222-227), so
        the vocoder is fed from the encode kernel's side output instead of running BVRNN.decode again.  Calling encode()
        and decode() separately runs both recurrences, as it must."""
        length = x.shape[1]
        on_cpu = x.device.type == "cpu"
        mel = self._engine.logmel(x.to(self.device), SCALING)
        out = self._engine.encode(mel, None, self.bits_per_frame(bitrate), None, want_all_h=False, want_mel=True)
        wav = self._engine.vocode(out[5], length, SCALING)
        return wav.cpu() if on_cpu else wav

    # ---- packed wire format (not in the reference, which moves 256 B of floats per 35-bit frame) ----
    def encode_packed(self, x, bitrate):
        """x (B, L) CUDA tensor -> (packed int64 (B, L // hop), bits per frame); bit i of a word is code i."""
        bits = self.bits_per_frame(bitrate)
        mel = self._engine.logmel(x.to(self.device), SCALING)
        _, _, _, _, packed = self._engine.encode(mel, None, bits, None, want_all_h=False, want_packed=True)
        return packed, bits

    def decode_packed(self, packed, bits, length):
        """Inverse of encode_packed: packed int64 (B, T), bits per frame (number or (B, T) tensor) -> waveform (B, length).
        The words go straight into the decoder (bvc_decode_packed); the float code tensor is never formed."""
        packed = packed.to(self.device)
        if torch.is_tensor(bits):
            mel, _ = self._engine.decode_mel_packed(packed, bits.to(self.device, torch.float32).contiguous(), 0.0, None)
        else:
            mel, _ = self._engine.decode_mel_packed(packed, None, float(bits), None)
        return self._engine.vocode(mel, length, SCALING)

    def encode_bitstream(self, x, bitrate):
        """x (B, L) -> uint8 (B, n_bytes) self-describing bit-streams: 16-byte header + round(bitrate * hop / fs) bits per
        frame (include/bvc.h; 3 kbps: 4.4 bytes per frame instead of the 256 bytes of the float codes)."""
        packed, bits = self.encode_packed(x, bitrate)
        return self._engine.pack_bitstream(packed, None, bits)

    @staticmethod
    def bitstream_header(stream_row):
        """Parses the 16-byte header of one stream: dict(z_dim, mode, n_bits, T, payload_bits)."""
        hdr = bytes(stream_row[:16].cpu().tolist())
        if hdr[:4] != b"BVC1":
            raise ValueError("not a BVC1 bit-stream")
        return dict(z_dim=int.from_bytes(hdr[4:6], "little"), mode=hdr[6], n_bits=hdr[7],
                    T=int.from_bytes(hdr[8:12], "little"), payload_bits=int.from_bytes(hdr[12:16], "little"))

    def decode_bitstream(self, stream, length):
        """Inverse of encode_bitstream (all rows must share T): uint8 (B, n_bytes) -> waveform (B, length)."""
        stream = stream.to(self.device)
        T = self.bitstream_header(stream[0])["T"]
        packed, bits = self._engine.unpack_bitstream(stream, T)
        mel, _ = self._engine.decode_mel_packed(packed, bits, 0.0, None)
        return self._engine.vocode(mel, length, SCALING)

    # ---- taps used by the parity tests (not part of the reference API) ----
    def encode_with_taps(self, x, bitrate, h0=None, mel=None):
        """Returns dict(mel, codes, logits, all_h, h_final, packed) for a CUDA input."""
        if mel is None:
            mel = self._engine.logmel(x, SCALING)
        codes, all_h, h_fin, logits, packed = self._engine.encode(
            mel, None, self.bits_per_frame(bitrate), h0, want_logits=True, want_all_h=True, want_packed=True)
        return dict(mel=mel, codes=codes, logits=logits, all_h=all_h, h_final=h_fin, packed=packed)


def mel_spectrogram(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, padding_left,
                    center=False, return_stft=False, *, _engine_cache={}):
    """Mirror of reference third_party/BigVGAN/meldataset.py:60 for CUDA inputs: (B, L) -> (B, num_mels, T)."""
    if center or return_stft:
        raise NotImplementedError("center=True / return_stft=True are not on the codec path")
    if y.device.type != "cuda":
        raise RuntimeError("mel_spectrogram (b200) needs a CUDA tensor")
    key = (y.device, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, padding_left)
    eng = _engine_cache.get(key)
    if eng is None:
        if win_size != n_fft:
            raise NotImplementedError("win_size must equal n_fft")
        conf = dict(num_mels=num_mels, h_dim=1024, z_dim=64, var_bit=True, fs=sampling_rate, winsize=n_fft,
                    hopsize=hop_size, mel_pad_left=padding_left if padding_left != -1 else (n_fft - hop_size) // 2,
                    fmin=fmin, fmax=fmax,
                    vocoder_config=dict(upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4],
                                        upsample_initial_channel=128, resblock_kernel_sizes=[3, 7, 11],
                                        resblock_dilation_sizes=[[1, 3, 5]] * 3, resblock="1"))
        eng = _engine_cache[key] = _Engine(conf, y.device)
    return eng.logmel(y, 1.0).permute(0, 2, 1)

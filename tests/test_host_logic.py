"""CPU tests of the host side: C-ABI surface, mel basis, checkpoint synthesis, sharding (gloo, world 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "bvc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bvc_[a-z_0-9]+)\s*\(", src)))


def test_library_builds_loads_and_exports_header_symbols():
    import __graft_entry__
    __graft_entry__.build()
    from bernoulli_var_speech_codec_b200 import _lib
    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 18
    for n in names:
        assert n in _lib.SYMBOLS, f"{n} declared in bvc.h but not bound"
        getattr(lib, n)
    assert lib.bvc_abi_version() == 3


def test_create_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from bernoulli_var_speech_codec_b200 import _lib, BVRNNCodecModel
    lib = _lib.load()
    cfg = _lib.BvcConfig()
    h = ctypes.c_void_p()
    rc = lib.bvc_create(ctypes.byref(h), ctypes.byref(cfg))
    assert rc == -3 and b"no CPU path" in lib.bvc_last_error()
    with pytest.raises(RuntimeError, match="no CPU path"):
        BVRNNCodecModel(os.path.join(ROOT, "configs", "config_varBitRate.toml"), "x", "y")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "bernoulli_var_speech_codec_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "/root/reference" not in txt, f


def test_mel_basis_properties():
    from bernoulli_var_speech_codec_b200.melbasis import slaney_mel_basis, sparse_rows
    from oracle.codec_oracle import slaney_mel
    b = slaney_mel_basis(22050, 1024, 80, 0, 8000)
    assert b.shape == (80, 513) and b.dtype == np.float32 and (b >= 0).all()
    assert np.array_equal(b, slaney_mel(22050, 1024, 80, 0, 8000))    # two independent restatements agree
    assert np.nonzero(b.sum(0))[0].max() == 371
    start, count, taps = sparse_rows(b)
    assert count.min() == 3 and count.max() == 27
    dense = np.zeros_like(b)
    for m in range(80):
        dense[m, start[m]:start[m] + count[m]] = taps[m, :count[m]]
    assert np.array_equal(dense, b)
    # Slaney area normalisation: each triangle integrates to ~1 over Hz (bin width 22050/1024)
    area = b.sum(1) * 22050 / 1024
    assert np.abs(area[5:] - 1.0).max() < 0.25


def test_synthetic_checkpoint_schema(ckpts):
    sd = torch.load(ckpts[0], map_location="cpu", weights_only=True)["vrnn"]
    assert len(sd) == 39 and sd["rnn.weight_ih_l0"].shape == (3072, 2048) and sd["enc.0.weight"].shape == (1024, 2048)
    gd = torch.load(ckpts[1], map_location="cpu", weights_only=True)["generator"]
    assert len(gd) == 380 and gd["ups.0.1.weight_v"].shape == (128, 64, 16)
    assert gd["resblocks.11.convs2.2.weight_v"].shape == (8, 8, 11) and gd["conv_post.weight_g"].shape == (1, 1, 1)


def test_shard_bounds_cover_everything():
    from bernoulli_var_speech_codec_b200.sharding import shard_bounds
    for n in (0, 1, 7, 256, 1024, 1025):
        for ws in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from bernoulli_var_speech_codec_b200.sharding import ShardedCodec, shard_bounds
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)

class FakeCodec:          # stands in for the CUDA codec: rows are independent, like the real path
    def encode(self, x, bitrate):
        T = x.shape[1] // 256
        return (x[:, : T * 256].reshape(x.shape[0], T, 256)[:, :, :64] > 0).float()
    def decode(self, codes, length):
        return codes.sum(-1).repeat_interleave(256, dim=1)[:, :length] * 0.5

g = torch.Generator().manual_seed(0)
x = torch.randn(5, 256 * 6 + 17, generator=g)          # 5 utterances over 2 ranks: ragged shards 3 + 2
sc = ShardedCodec(FakeCodec())
codes = sc.encode(x, 3000)
wav = sc.decode(codes, x.shape[1])
ref_codes = FakeCodec().encode(x, 3000)
ref_wav = FakeCodec().decode(ref_codes, x.shape[1])
assert torch.equal(codes, ref_codes) and torch.equal(wav, ref_wav)
assert torch.equal(sc(x, 3000), ref_wav)
lo, hi = shard_bounds(5, 2, dist.get_rank())
assert sc.encode(x, 3000, gather=False).shape[0] == hi - lo
dist.barrier()
dist.destroy_process_group()
print("ok")
"""


def test_sharded_codec_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run on the B200 box)")


@pytest.fixture(scope="session")
def ckpts():
    """Synthetic checkpoints with the reference's state-dict schema (seed/sharpen of the golden fixtures)."""
    from bernoulli_var_speech_codec_b200.synth import write_synthetic_checkpoints
    d = os.environ.get("BVC_CKPT_DIR", "/tmp/bvc_ckpts")
    return write_synthetic_checkpoints(d, seed=1, sharpen=30.0)


@pytest.fixture(scope="session")
def cfg_var():
    return os.path.join(ROOT, "configs", "config_varBitRate.toml")


@pytest.fixture(scope="session")
def cfg_fix():
    return os.path.join(ROOT, "configs", "config_64bit.toml")


@pytest.fixture(scope="session")
def cfg_aa():
    """config_varBitRate with the anti-aliased activations on (vocoder_config.layers_antialias / antialias_post)."""
    return os.path.join(ROOT, "configs", "config_varBitRate_antialias.toml")


@pytest.fixture(scope="session")
def ckpts_aa(ckpts, cfg_aa):
    """Same BVRNN checkpoint; vocoder checkpoint with the reference's Activation1d schema."""
    import toml
    from bernoulli_var_speech_codec_b200.synth import write_synthetic_checkpoints
    d = os.environ.get("BVC_CKPT_DIR", "/tmp/bvc_ckpts")
    return write_synthetic_checkpoints(d, seed=1, sharpen=30.0, vcfg=toml.load(cfg_aa)["vocoder_config"])


@pytest.fixture(scope="session")
def oracle_aa(ckpts_aa, cfg_aa):
    from oracle.codec_oracle import OracleCodec
    return OracleCodec(cfg_aa, *ckpts_aa)


@pytest.fixture(scope="session")
def oracle_var(ckpts, cfg_var):
    from oracle.codec_oracle import OracleCodec
    return OracleCodec(cfg_var, *ckpts)


@pytest.fixture(scope="session")
def oracle_fix(ckpts, cfg_fix):
    from oracle.codec_oracle import OracleCodec
    return OracleCodec(cfg_fix, *ckpts)


# Every GPU parity test runs in both arithmetic modes of the library: 1 = split-bf16 tensor-core kernels (the
# default and the path bench.py measures), 0 = fp32 FFMA kernels.
PRECISIONS = [pytest.param(1, id="tensorcore"), pytest.param(0, id="fp32")]


@pytest.fixture(scope="session", params=PRECISIONS)
def model_var(request, ckpts, cfg_var):
    from bernoulli_var_speech_codec_b200 import BVRNNCodecModel
    m = BVRNNCodecModel(cfg_var, *ckpts).eval()
    m._engine.set_precision(request.param)
    m.precision = request.param
    return m


@pytest.fixture(scope="session", params=PRECISIONS)
def model_fix(request, ckpts, cfg_fix):
    from bernoulli_var_speech_codec_b200 import BVRNNCodecModel
    m = BVRNNCodecModel(cfg_fix, *ckpts).eval()
    m._engine.set_precision(request.param)
    m.precision = request.param
    return m


@pytest.fixture(scope="session")
def model_aa(ckpts_aa, cfg_aa):
    from bernoulli_var_speech_codec_b200 import BVRNNCodecModel
    return BVRNNCodecModel(cfg_aa, *ckpts_aa).eval()


def golden(name):
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", name))

"""Import the UNMODIFIED reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY; used by ``oracle/gen_golden.py`` to produce the
committed golden vectors and by the optional ``-m "not gpu"`` cross-check that
runs when /root/reference is present.  /root/reference does not exist on the
GPU box, so nothing that runs there imports this module.

The reference imports three packages that are not installed here at module
scope: librosa (third_party/BigVGAN/meldataset.py:13-15),
matplotlib (third_party/BigVGAN/utils.py:6-10).  They are replaced by minimal
stub modules; the only arithmetic among them is ``librosa.filters.mel``, which
is served by the oracle's Slaney restatement.
"""
from __future__ import annotations

import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# On the GPU box /root/reference does not exist; __graft_entry__.build() (run in the build container) copies the handful
# of reference modules the codec path imports into the git-ignored baseline/_ref/, which travels with the snapshot, so
# that `bench.py --impl reference` can time the UNMODIFIED reference there.  Never part of the product, never committed.
BASELINE_REF = os.path.join(_REPO, "baseline", "_ref")
REFERENCE_ROOT = os.environ.get("BVC_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isfile("/root/reference/bvrnn_codec_model.py") else BASELINE_REF)

REF_FILES = [
    "bvrnn.py", "bvrnn_codec_model.py", "LICENSE",
    "third_party/BigVGAN/__init__.py", "third_party/BigVGAN/models.py", "third_party/BigVGAN/activations.py",
    "third_party/BigVGAN/meldataset.py", "third_party/BigVGAN/env.py", "third_party/BigVGAN/utils.py",
    "third_party/BigVGAN/LICENSE",
    "third_party/BigVGAN/alias_free_torch/__init__.py", "third_party/BigVGAN/alias_free_torch/act.py",
    "third_party/BigVGAN/alias_free_torch/filter.py", "third_party/BigVGAN/alias_free_torch/resample.py",
]


def populate_baseline_ref(src_root: str = "/root/reference") -> bool:
    """Copies the reference modules of the codec path, byte for byte, into baseline/_ref/ (git-ignored).  Build container only."""
    import shutil
    if not os.path.isfile(os.path.join(src_root, "bvrnn_codec_model.py")):
        return False
    for rel in REF_FILES:
        src = os.path.join(src_root, rel)
        if not os.path.isfile(src):
            continue
        dst = os.path.join(BASELINE_REF, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    return True


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "bvrnn_codec_model.py"))


def _install_stubs():
    from oracle.codec_oracle import slaney_mel

    if "librosa" not in sys.modules:
        librosa = types.ModuleType("librosa")
        util = types.ModuleType("librosa.util")
        filters = types.ModuleType("librosa.filters")

        def normalize(x, *a, **k):
            import numpy as np
            return x / np.max(np.abs(x))

        def mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **kw):
            return slaney_mel(sr, n_fft, n_mels, fmin, fmax if fmax is not None else sr / 2)

        util.normalize = normalize
        filters.mel = mel
        librosa.util, librosa.filters = util, filters
        sys.modules.update({"librosa": librosa, "librosa.util": util, "librosa.filters": filters})
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        pylab = types.ModuleType("matplotlib.pylab")
        mpl.pylab = pylab
        sys.modules.update({"matplotlib": mpl, "matplotlib.pylab": pylab})
    if "toml" not in sys.modules:
        try:
            import toml  # noqa: F401
        except ImportError:
            import tomllib
            t = types.ModuleType("toml")
            t.load = lambda p: tomllib.load(open(p, "rb"))
            sys.modules["toml"] = t


def import_reference():
    """Returns the reference's ``bvrnn_codec_model`` module (unmodified source)."""
    if not available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib.util
    # loaded under a private name: the product ships a root-level alias module called
    # ``bvrnn_codec_model`` too, and the two must be importable side by side.
    name = "_reference_bvrnn_codec_model"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_ROOT, "bvrnn_codec_model.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod

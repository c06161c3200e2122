"""Drop-in alias: ``from bvrnn_codec_model import BVRNNCodecModel`` as in the reference (bvrnn_codec_model.py:19)."""
from bernoulli_var_speech_codec_b200.codec import (  # noqa: F401
    BVRNNCodecModel, SCALING, default_config, default_chkpt_bvrnn, default_chkpt_vocoder)

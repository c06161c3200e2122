"""B200-native (sm_100a) implementation of the batched BVRNN speech-codec encode->decode path.

Public surface mirrors the reference's ``bvrnn_codec_model.py``: ``BVRNNCodecModel`` with
``encode(x, bitrate)``, ``decode(codes, length)``, ``forward(x, bitrate)``.
"""
from .codec import BVRNNCodecModel, SCALING, mel_spectrogram, AttrDict  # noqa: F401

__all__ = ["BVRNNCodecModel", "SCALING", "mel_spectrogram", "AttrDict"]

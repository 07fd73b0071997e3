"""B200-native WaveRNN batched generation (drop-in for the reference's WaveRNN.generate path).

Public surface: `WaveRNN` (mirror of WaveRNN/models/fatchord_version.py::WaveRNN on the
generation path), `Synthesize` (WaveRNN/synthesizer_wavernn.py), `hparams`.  The compute lives
in csrc/ behind the C ABI of include/wavernn_b200.h.
"""
from . import hparams
from .synthesizer import Synthesize
from .wavernn import WaveRNN

__all__ = ["WaveRNN", "Synthesize", "hparams"]

"""waveforms_b200 — B200-native drop-in for the sampling hot path of
feihoo87/waveforms (same public names as /root/reference/waveforms/__init__.py).

Symbolic construction stays host Python; ``Waveform.sample`` / ``wav(t)`` /
``WaveVStack`` evaluation, the sample-time IIR and the ``distortion`` apply
functions run as hand-written sm_100a CUDA kernels behind a C-ABI
(include/wfm_b200.h).  ``sample_batch`` is the batched, multi-GPU entry point.
"""
from numpy import e, pi

from .multy_drag import drag_sin, drag_sinx
from .version import __version__
from .waveform import (D, Waveform, WaveVStack, chirp, const, cos, cosh,
                       coshPulse, cosPulse, cut, drag, exp, function, gaussian,
                       general_cosine, hanning, interp, mixing, mollifier, one,
                       poly, registerBaseFunc, registerDerivative,
                       samplingPoints, sign, sin, sinc, sinh, square, step, t,
                       zero)
from .waveform_parser import wave_eval
from .batch import sample_batch

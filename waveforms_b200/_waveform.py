"""Name-compatible alias of the reference's compiled module
(/root/reference/waveforms/_waveform.pyx) for callers that import from
``waveforms._waveform`` (e.g. ``wave_sum`` in the reference's own tests).
``calc_parts`` is intentionally absent: evaluation is the CUDA path."""
from ._algebra import *  # noqa: F401,F403
from ._algebra import (_D, _baseFunc, _baseFunc_latex, _const,
                       _derivativeBaseFunc, _half, _one, _zero)  # noqa: F401

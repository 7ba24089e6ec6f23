"""fft_lines_b200 -- B200-native (sm_100a) spectrum hot path of ImNotan/fft_lines.

The product is the C-ABI shared library ``libfftlines_cuda.so`` built from
``csrc/`` (see ``include/*.h``).  This package is the thin Python host side:

* :mod:`fft_lines_b200.fft_calculate`, :mod:`fft_lines_b200.beat_detector` --
  mirrors of the reference's ``fft_calculate.h`` / ``beatDetector.h`` entry
  points (same names, same argument meaning, same error behaviour);
* :mod:`fft_lines_b200.stft` -- the batched STFT generalisation;
* :mod:`fft_lines_b200.sharding` -- frame-range sharding across GPUs.

There is no CPU fallback and nothing here imports ``oracle/``.
"""
from ._capi import (FftlConfig, FftlBeat, FftLinesError, WINDOW_RECT, WINDOW_HANN, BINNING_LINEAR,
                    BINNING_LOG, BEAT_SAMPLES, TEMPO_RING, LIB_PATH)
from .stft import (BatchedSpectrum, reference_config, extended_config, build_bins, build_sg, BEAT_DTYPE,
                   device_count, synth_pcm_device, PinnedArray, measure_fp32_peak)
from .stream import StreamingSpectrum
from ._capi import MOVE_EXACT, MOVE_REFERENCE, TICK_AUTO, TICK_THREE_LAUNCHES
from . import fft_calculate, beat_detector, sharding, sinks

__all__ = ["FftlConfig", "FftlBeat", "FftLinesError", "BatchedSpectrum", "reference_config",
           "extended_config", "build_bins", "build_sg", "BEAT_DTYPE", "device_count", "synth_pcm_device",
           "PinnedArray", "measure_fp32_peak", "StreamingSpectrum", "MOVE_EXACT", "MOVE_REFERENCE", "TICK_AUTO", "TICK_THREE_LAUNCHES", "fft_calculate", "beat_detector", "sharding", "sinks", "WINDOW_RECT", "WINDOW_HANN",
           "BINNING_LINEAR", "BINNING_LOG", "BEAT_SAMPLES", "TEMPO_RING", "LIB_PATH"]

"""Python mirror of the reference's ``fft_calculate.h`` (reference
fft_lines/fft_calculate.h:5-11), bound to libfftlines_cuda's drop-in symbols.

Buffers are numpy ``complex64`` arrays (layout-identical to ``fftwf_complex``
= ``float[2]``).  As in the reference (fft_calculate.c:22-30) the plan captures
the two buffers at ``initializefft`` time and ``executefft`` ignores its
arguments; the arrays must therefore stay alive while the plan does, which
this module guarantees by keeping references.
"""
import numpy as np

from . import _capi

_bound = {}


def _ptr(a, n):
    if not isinstance(a, np.ndarray) or a.dtype != np.complex64 or not a.flags.c_contiguous or a.size < n:
        raise TypeError("need a C-contiguous complex64 array with at least %d elements" % n)
    return a.ctypes.data


def set_n(n):
    """Transform size used by the next initializefft (reference: compile-time N=2048, global.h:3)."""
    _capi.check(_capi.lib().fftl_set_n(n))


def get_n():
    return _capi.lib().fftl_get_n()


def initializefft(inp, out):
    n = get_n()
    _bound["fft"] = (inp, out)
    _capi.lib().initializefft(_ptr(inp, n), _ptr(out, n))


def executefft(inp=None, out=None):
    """Runs the bound transform; like the reference, the arguments are ignored."""
    _capi.lib().executefft(None, None)


def cleanupfft():
    _capi.lib().cleanupfft()
    _bound.pop("fft", None)


def initializefft256(inp, out):
    _bound["fft256"] = (inp, out)
    _capi.lib().initializefft256(_ptr(inp, _capi.TEMPO_RING), _ptr(out, _capi.TEMPO_RING))


def executefft256(inp=None, out=None):
    _capi.lib().executefft256(None, None)


def cleanupfft256():
    _capi.lib().cleanupfft256()
    _bound.pop("fft256", None)


def bound_buffers(which="fft"):
    return _bound.get(which)

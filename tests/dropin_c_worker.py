"""Subprocess side of tests/test_gpu_dropin_c.py: loads ONLY the harness built against libfftlines_cuda (so that the
library resolves the application's `bassBeatBuffer`, a weak import, against the harness as it would against the
reference's settingsFile.c:83) and replays one golden fixture through the restated WndProc.

    python tests/dropin_c_worker.py <libharness_cuda.so> <fixture.npz> <out.npz>
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402  (the checker: synthetic PCM generator and struct layouts only)

so, fixture, out = sys.argv[1:4]
g = np.load(fixture)
ch, S, hop, bars = int(g["channels"]), int(g["samples"]), int(g["hop"]), int(g["bars"])
first = int(g["first_sample"]) if "first_sample" in g else 0
pcm = O.synth_pcm(ch, S, first_sample=first)
R = C.CDLL(so, mode=C.RTLD_GLOBAL)
i32p, i16p = C.POINTER(C.c_int32), C.POINTER(C.c_int16)
R.ref_open.argtypes = [C.c_int]
R.ref_run.restype = C.c_double
R.ref_run.argtypes = [i16p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_float, i32p, C.c_void_p]
assert R.ref_open(2048) == 0 and R.ref_provider_error() == 0
R.ref_reset_beat()
F = (S - 2048) // hop + 1
heights = np.zeros((ch, F, bars), np.int32)
fo = np.zeros(F, O.REF_FRAME_DTYPE)
R.ref_run(pcm.ctypes.data_as(i16p), ch, S, S, hop, bars, float(g["zoom"]), int(g["circle"]), 1, 48000.0,
          heights.ctypes.data_as(i32p), fo.ctypes.data)
err = R.ref_provider_error()
R.ref_close()
np.savez(out, heights=heights, w=fo["w"], beat_value=fo["beat_value"], bass=fo["bass"], movement=fo["movement_per_sec"], err=err)

"""The drop-in proved from C: translation units that know only include/*.h are compiled with gcc and
linked against libfftlines_cuda.so with the command line INTEGRATION.md gives.

  (a) examples/headless_tick.c -- the reference's WM_CREATE / WM_TIMER / WM_DESTROY calls on a synthetic sine;
  (b) oracle/ref_harness.c -- the SAME restated WndProc that drives the unmodified reference files, built with the
      CUDA library as the provider instead of fft_calculate.c + beatDetector.c + the FFT shim, replaying the golden
      fixtures the unmodified reference produced (tests/golden/*.npz).
"""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
LIBDIR = os.path.join(ROOT, "fft_lines_b200")
BUILD = os.path.join(ROOT, "tests", "_build")


def _gcc(args):
    os.makedirs(BUILD, exist_ok=True)
    r = subprocess.run(["gcc"] + args, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return r.stdout


def test_headless_tick_example_builds_and_runs(lib):
    exe = os.path.join(BUILD, "headless_tick")
    # INTEGRATION.md section 2, verbatim apart from the output name; -Wall -Werror: the headers are clean C
    _gcc(["-Wall", "-Werror", "-Iinclude", "examples/headless_tick.c", "-L" + LIBDIR, "-lfftlines_cuda", "-Wl,-rpath," + LIBDIR, "-lm", "-o", exe])
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    rows = re.findall(r"tick (\d+): height\[5\]=(-?\d+) beatValue=(-?\d+) beatMovementPerSec=(\S+)", r.stdout)
    assert [int(t[0]) for t in rows] == [0, 1, 2, 3, 4]
    # on-bin sine, amplitude 10000, bin 5, N = 2048, zoom 1e-4: |X[5]| = A*N/2 -> height 1023 or 1024;
    # w = (0 + 0 + h5 + 0)/4 = 255 or 256 every tick -> threshold (k+1)*w/160 -> beatValue = (int)(w/thr*128)
    for k, (_, h5, bv, _) in enumerate(rows):
        h5, bv = int(h5), int(bv)
        assert h5 in (1023, 1024)
        w = h5 // 4
        thr = (k + 1) * w // 160
        want = int(np.float32(np.float32(w) / np.float32(thr)) * np.float32(128)) if thr else -2 ** 31
        assert bv == want, (k, bv, want)


@pytest.mark.parametrize("name", ["ref_c1_stereo.npz", "ref_mono_circle.npz"])
def test_reference_harness_linked_against_the_cuda_library(lib, name, tmp_path):
    so = os.path.join(BUILD, "libharness_cuda.so")
    _gcc(["-O2", "-fPIC", "-shared", "-DFFTL_HARNESS_CUDA_PROVIDER", "-Iinclude", "oracle/ref_harness.c", "-L" + LIBDIR,
          "-lfftlines_cuda", "-Wl,-rpath," + LIBDIR, "-lm", "-o", so])
    out = str(tmp_path / "out.npz")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_c_worker.py"), so, os.path.join(GOLDEN, name), out],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout
    got, g = np.load(out), np.load(os.path.join(GOLDEN, name))
    assert int(got["err"]) == 0
    # heights: two correct fp32 FFTs differ by at most one count (the golden ones are (int) casts of the reference's)
    d = got["heights"].astype(np.int64) - g["heights"]
    assert np.abs(d).max() <= 1 and (d != 0).mean() < 2e-3
    assert np.abs(got["w"].astype(np.int64) - g["w"]).max() <= 1
    assert np.abs(got["bass"].astype(np.int64) - g["bass"]).max() <= 1     # the mirrored bassBeatBuffer slot
    # AddBeatSample chain: bit exact on the w sequence the harness fed it (replayed by the oracle's restatement of
    # beatDetector.c:30-44), and equal to the reference's own outputs wherever w itself agrees up to that frame
    from oracle import oracle as O
    rep = O.beat_replay(O.reference_config(), got["w"], np.zeros_like(got["w"]))
    assert (rep["beat_value"] == got["beat_value"]).all()
    if (got["w"] == g["w"]).all():
        assert (got["beat_value"] == g["beat_value"]).all()
    # tempo: equal to the reference, or an explained integer-boundary flip (helpers._peak_possible)
    from helpers import _peak_possible, TEMPO_REL
    n_dt = np.float32(np.float32(2048) * (np.float32(int(g["hop"])) / np.float32(48000.0)))
    peaks = np.rint(got["movement"] / n_dt).astype(np.int64)
    assert (got["movement"] == n_dt * peaks.astype(np.float32)).all()
    differ = np.nonzero(got["movement"] != g["movement"])[0]
    if differ.size:
        mag, norm = O.tempo_mag_replay(got["bass"])
        for f in differ:
            assert _peak_possible(mag[f], 1.0 + TEMPO_REL * norm[f], int(peaks[f])), "frame %d: peak %d unexplained" % (f, peaks[f])

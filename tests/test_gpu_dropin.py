"""GPU tests of the drop-in layer: the reference's own call sequence
(WM_CREATE / WM_TIMER / WM_DESTROY, fft_lines.c:108-116, :186-236, :281,
:420-433) driven through initializefft / executefft / BassBeatDetector /
AddBeatSample / GetBeatValue exactly as WndProc does, compared with the golden
vectors produced by the unmodified reference files."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REAL, IMAG = 0, 1


class App:
    """The slice of the reference's WndProc that touches the hot path, in Python."""

    def __init__(self, F, n=2048):
        self.F, self.N = F, n
        fc = F.fft_calculate
        fc.set_n(n)
        # WM_CREATE (fft_lines.c:110-116)
        self.input = np.zeros(n, np.complex64)
        self.output = np.zeros(n, np.complex64)
        fc.initializefft(self.input, self.output)
        self.beatinput = np.zeros(256, np.complex64)
        self.beatoutput = np.zeros(256, np.complex64)
        fc.initializefft256(self.beatinput, self.beatoutput)
        F.beat_detector.reset()

    def timer(self, left, right, barCount, zoom, circle, td):
        """WM_TIMER (fft_lines.c:190-236, :281) -- same statements, numpy for the loops."""
        fc, bd = self.F.fft_calculate, self.F.beat_detector
        self.input[:] = left.astype(np.float32)            # :190-194
        fc.executefft(self.input, self.output)             # :196
        o = self.output[:barCount]
        z = np.float32(zoom)
        hl = (np.sqrt(o.real.astype(np.float64) ** 2 + o.imag.astype(np.float64) ** 2) * float(z)).astype(np.int32)  # :202
        hl += 10 if circle else 0
        hr = None
        if right is not None:                              # :210-231
            self.input[:] = right.astype(np.float32)
            fc.executefft(self.input, self.output)
            o = self.output[:barCount]
            hr = (np.sqrt(o.real.astype(np.float64) ** 2 + o.imag.astype(np.float64) ** 2) * float(z)).astype(np.int32)
            hr += 10 if circle else 0
        bd.set_timeDelta(td)                               # :265
        o4 = self.output[4]                                # what BassBeatDetector writes to the ring (beatDetector.c:53)
        self.bass = int(np.sqrt(np.float64(o4.real) ** 2 + np.float64(o4.imag) ** 2))
        bd.BassBeatDetector(self.input, self.output, self.beatinput, self.beatoutput)  # :233-236
        w = int((int(hl[3]) + int(hl[4]) + int(hl[5]) + int(hl[6])) / 4)               # :281 (C int division)
        bd.AddBeatSample(w)
        return hl, hr, w, bd.GetBeatValue(), bd.get_beatMovementPerSec()

    def destroy(self):
        self.F.fft_calculate.cleanupfft()                  # :420-421
        self.F.fft_calculate.cleanupfft256()


@pytest.mark.parametrize("name", ["ref_c1_stereo.npz", "ref_mono_circle.npz"])
def test_wm_timer_sequence_against_reference_golden(oracle, lib, name):
    g = np.load(os.path.join(GOLDEN, name))
    ch, S, hop, bars = int(g["channels"]), int(g["samples"]), int(g["hop"]), int(g["bars"])
    first = int(g["first_sample"]) if "first_sample" in g else 0
    pcm = oracle.synth_pcm(ch, S, first_sample=first)
    frames = min(g["heights"].shape[1], 120)
    app = App(lib)
    td = np.float32(hop) / np.float32(48000.0)
    ws, bvs, movs, basses = [], [], [], []
    worst = 0
    for f in range(frames):
        l = pcm[0, f * hop:f * hop + 2048]
        r = pcm[1, f * hop:f * hop + 2048] if ch == 2 else None
        hl, hr, w, bv, mov = app.timer(l, r, bars, float(g["zoom"]), int(g["circle"]), float(td))
        worst = max(worst, np.abs(hl - g["heights"][0, f]).max())
        if ch == 2:
            worst = max(worst, np.abs(hr - g["heights"][1, f]).max())
        ws.append(w); bvs.append(bv); movs.append(mov); basses.append(app.bass)
    app.destroy()
    assert worst <= 1                                       # two correct fp32 FFTs differ by at most one count
    ws, bvs, movs = np.array(ws), np.array(bvs), np.array(movs, np.float32)
    assert (np.abs(ws - g["w"][:frames]) <= 1).all()
    # AddBeatSample on the GPU, replayed by the oracle on the same w sequence: bit exact
    rep = oracle.beat_replay(oracle.reference_config(), ws, np.zeros_like(ws))
    assert (rep["beat_value"] == bvs).all()
    same_w = (ws == g["w"][:frames]).all()
    if same_w:
        assert (bvs == g["beat_value"][:frames]).all()
    # BassBeatDetector: movement = N * timeDelta * peak index (beatDetector.c:105).  The index must be the one the
    # unmodified reference produced, or differ from it only by an integer-boundary flip of the peak scan
    # (helpers._peak_possible: every differing frame has to be explained, none may be merely tolerated).
    from helpers import _peak_possible, TEMPO_REL
    n_dt = np.float32(np.float32(2048) * td)
    peaks = np.rint(movs / n_dt).astype(np.int64)
    assert (movs == n_dt * peaks.astype(np.float32)).all()
    differ = np.nonzero(movs != g["movement"][:frames])[0]
    if differ.size:
        mag, norm = oracle.tempo_mag_replay(np.array(basses, np.int32))
        for f in differ:
            assert _peak_possible(mag[f], 1.0 + TEMPO_REL * norm[f], int(peaks[f])), "frame %d: peak %d unexplained" % (f, peaks[f])


def test_executefft_matches_definition_all_sizes(oracle, lib):
    fc = lib.fft_calculate
    rng = np.random.default_rng(5)
    for n in (256, 512, 1024, 2048, 4096, 8192, 16384):
        fc.set_n(n)
        x = (rng.integers(-32768, 32768, n) + 1j * rng.integers(-32768, 32768, n)).astype(np.complex64)
        inp, out = x.copy(), np.zeros(n, np.complex64)
        fc.initializefft(inp, out)
        fc.executefft(None, None)      # the reference ignores the arguments too (fft_calculate.c:27-30)
        ref = oracle.fft(x.astype(np.complex128))
        assert np.abs(out - ref).max() <= 3e-6 * np.abs(ref).max(), n
        inp[:] = 0; inp[1] = 1
        fc.executefft(inp, out)
        k = np.arange(n)
        assert np.abs(out - np.exp(-2j * np.pi * k / n)).max() < 2e-6
        fc.cleanupfft()
    fc.set_n(2048)
    from fft_lines_b200 import _capi
    assert _capi.last_error()[0] == 0


def test_executefft256_and_state_errors(oracle, lib):
    from fft_lines_b200 import _capi
    fc = lib.fft_calculate
    L = _capi.lib()
    fc.cleanupfft256()
    L.fftl_clear_error()
    fc.executefft256(None, None)                       # before initialise: latched error, no crash
    assert _capi.last_error()[0] == _capi.FFTL_ERR_STATE
    L.fftl_clear_error()
    x = (np.arange(256) % 7).astype(np.complex64)
    out = np.zeros(256, np.complex64)
    fc.initializefft256(x, out)
    fc.executefft256(x, out)
    ref = oracle.fft(x.astype(np.complex128))
    assert np.abs(out - ref).max() <= 2e-6 * np.abs(ref).max()
    fc.cleanupfft256(); fc.cleanupfft256()
    assert _capi.last_error()[0] == 0
    assert L.fftl_compat_launch_count() > 0


def test_add_beat_sample_golden(lib):
    g = np.load(os.path.join(GOLDEN, "ref_add_beat_sample.npz"))
    bd = lib.beat_detector
    bd.reset()
    got = []
    for w in g["w"]:
        bd.AddBeatSample(int(w))
        got.append(bd.GetBeatValue())
    assert (np.array(got, np.int32) == g["beat_value"]).all()   # includes the thr==0 -> INT32_MIN frames

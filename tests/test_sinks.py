"""Callers / sinks either side of the path: Settings.txt line, RANGE255,
serial/LED byte, beat-marker sweep -- product (C-ABI) against the oracle's restatement of the
reference lines (settingsFile.c:172-247,:376; drawBar2D.cpp:33; fft_lines.c:250-255,:266-274)."""
import numpy as np
import pytest


def _fields(s):
    return [s.bar_count, s.color_sel, s.dofft, s.background, s.gradient, s.ignore_serial, s.circle, s.waveform, s.stereo,
            s.beat_detection]


@pytest.mark.parametrize("line", [
    "500 0.000100 0 1 1 1 1 0 0 0 1",        # what writeSettings emits for the defaults
    "250 0.000500 3 1 0 1 1 1 0 1 0",
    "99 0.5 6 0 0 0 0 0 0 0 0",              # bar count, zoom, colour out of range -> defaults
    "1001 0.000001 -1 2 3 4 5 6 7 8 9",      # bools: any non-zero reads as true
    "100 0.001 5 1 1",                       # short line: the rest parse as 0
    "", "garbage", "700",
    "1000 0.00001 0 0 0 0 0 0 0 0 0" + " " * 900,
    "x" * 1001,                              # > 1000 bytes: file ignored, everything default
])
def test_settings_parse_matches_reference_semantics(lib, oracle, line):
    s = lib.sinks.settings_parse(line)
    v, zoom = oracle.settings_parse(line)
    assert _fields(s) == [int(v[0])] + [int(x) for x in v[2:]]
    assert np.float32(s.zoom) == np.float32(zoom)


def test_settings_roundtrip_and_config(lib):
    s = lib.sinks.settings_default()
    assert lib.sinks.settings_format(s) == "500 0.000100 0 1 1 1 1 0 0 0 1"
    s2 = lib.sinks.settings_parse("640 0.000250 2 1 1 0 1 1 0 1 1")
    assert lib.sinks.settings_format(s2) == "640 0.000250 2 1 1 0 1 1 0 1 1"
    c = lib.sinks.settings_to_config(s2, 2048, 512, 44100.0)
    assert (c.n, c.hop, c.bars, c.circle, c.window, c.binning) == (2048, 512, 640, 1, 0, 0)
    assert np.float32(c.zoom) == np.float32(0.00025) and c.sample_rate == 44100.0


def test_color_index(lib, oracle):
    for count in (100, 500, 1000, 2048):
        for i in list(range(0, count, max(1, count // 37))) + [count - 1]:
            assert lib.sinks.color_index(i, count) == oracle.range255(i, count)
    assert lib.sinks.color_index(0, 500) == 0 and lib.sinks.color_index(499, 500) == 254


@pytest.mark.gpu
def test_led_bytes(lib, oracle):
    import torch
    rng = np.random.default_rng(2)
    h = rng.integers(-50, 3000, (5000, 100)).astype(np.int32)
    h[:20, 4] = [0, 9, 10, 19, 1270, 1279, 1280, 2559, 2560, 2569, -1, -10, 2147483647, -2147483648, 5, 7, 1289, 12800, 25600, 99]
    b, v = lib.sinks.led_bytes_device(torch.from_numpy(h).cuda(), led_bar=4)
    rb, rv = oracle.led_bytes(h[:, 4])
    assert (b.cpu().numpy() == rb).all() and (v.cpu().numpy() == rv).all()


@pytest.mark.gpu
@pytest.mark.parametrize("start", [(0, 0), (0, 1), (2048, -1), (700, 1), (700, -1), (2048, 1)])
def test_marker_sweep_matches_reference_loop(lib, oracle, start):
    import torch
    rng = np.random.default_rng(3)
    frames = 5000
    mov = (rng.integers(0, 160, frames) * np.float32(2048 * 512 / 48000.0)).astype(np.float32)
    mov[100:140] = 0
    mov[200] = 1e6
    td = float(np.float32(512) / np.float32(48000))
    beat = np.zeros(frames, lib.BEAT_DTYPE)
    beat["movement_per_sec"] = mov
    pos, direction = lib.sinks.marker_sweep_device(torch.from_numpy(beat.view(np.int32).reshape(frames, 8)).cuda(), td, *start)
    rp, rd = oracle.marker_sweep(mov, td, *start)
    assert (pos.cpu().numpy() == rp).all()
    assert (direction.cpu().numpy() == rd).all()

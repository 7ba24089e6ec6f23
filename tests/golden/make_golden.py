"""Generates tests/golden/*.npz by running the UNMODIFIED reference files
(oracle/_ref: /root/reference/fft_lines/fft_calculate.c + beatDetector.c behind
oracle/ref_harness.c) in the build container.  /root/reference does not exist
on the GPU box, so the outputs are committed as fixtures.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    O.build()
    if not O.ref_available():
        raise SystemExit("oracle/_ref is not built (needs /root/reference)")
    # c1: the reference's own configuration -- N=2048, stereo, 100 bars, zoom 1e-4
    samples, hop, bars = 48000 * 3, 512, 100
    pcm = O.synth_pcm(2, samples)
    _, heights, fo = O.ref_run(pcm, hop, bars, zoom=1e-4, circle=0, beat=True)
    np.savez_compressed(os.path.join(HERE, "ref_c1_stereo.npz"), channels=2, samples=samples, hop=hop, bars=bars,
                        zoom=np.float32(1e-4), circle=0, seed=0x5EED0000, pcm_sum=np.int64(pcm.astype(np.int64).sum()), pcm_head=pcm[:, :32].copy(),
                        heights=heights.astype(np.int32), w=fo["w"], thr=fo["thr"], beat_value=fo["beat_value"],
                        bass=fo["bass"], movement=fo["movement_per_sec"])
    # mono, circle mode, 500 bars (the reference default bar count), larger zoom
    samples, hop, bars = 48000, 300, 500
    pcm = O.synth_pcm(1, samples, first_sample=123457)
    _, heights, fo = O.ref_run(pcm, hop, bars, zoom=1e-3, circle=1, beat=True)
    np.savez_compressed(os.path.join(HERE, "ref_mono_circle.npz"), channels=1, samples=samples, hop=hop, bars=bars,
                        zoom=np.float32(1e-3), circle=1, seed=0x5EED0000, first_sample=123457,
                        pcm_sum=np.int64(pcm.astype(np.int64).sum()), pcm_head=pcm[:, :32].copy(),
                        heights=heights.astype(np.int32), w=fo["w"], thr=fo["thr"], beat_value=fo["beat_value"],
                        bass=fo["bass"], movement=fo["movement_per_sec"])
    # AddBeatSample step response straight through the reference function
    R = O.ref()
    R.ref_open(0)
    R.ref_reset_beat()
    seq = np.concatenate([np.zeros(5, np.int32), np.full(200, 255, np.int32), np.arange(0, 400, 3, dtype=np.int32),
                          np.zeros(170, np.int32)])
    bv, th = [], []
    for w in seq:
        R.ref_add_beat_sample(int(w))
        bv.append(R.ref_get_beat_value())
        th.append(R.ref_get_threshold())
    R.ref_close()
    np.savez_compressed(os.path.join(HERE, "ref_add_beat_sample.npz"), w=seq, beat_value=np.array(bv, np.int32),
                        thr=np.array(th, np.int32))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()

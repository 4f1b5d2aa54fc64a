"""Command line of the renderer: the reference's `main` (apply_hrtf.py:559-649) on the CUDA path.

    python -m binaural_audio_synthesis_b200 INPUT.wav [--bank FILE.mat] [--trajectory NAME]
                                            [--samples-to-keep K] [--chunksize C] [--subchunksize S]

Like the reference it reads a wav file, scales it by its maximum (:576-577), folds stereo to mono
(:630), renders it along a trajectory (default `passing`, :633) and writes
INPUT-c<C>-s<S>-l<K>.wav as float32 (:636-640), then prints the speed relative to real time (:642-646).
The reference hard-codes the bank file name and the trajectory; here they are options with the
reference's values as defaults.
"""
from __future__ import annotations

import argparse
import sys
import time

import numpy as np


def trajectories(fs: float, period: float = 4.0, length: float = 30.0, turns: float = 15.0):
    """The named (elev, azim) trajectories of apply_hrtf.py:579-593, t in samples, radians out."""
    k = 2 * np.pi / (period * fs)
    return {
        'circle_front': lambda t: (np.sin(k * t), np.cos(k * t)),
        'circle_horizontal': lambda t: (0, (k * t) % (2 * np.pi)),
        'circle_askew': lambda t: ((np.pi / 4) * np.cos(k * t), (k * t) % (2 * np.pi)),
        'halfcircle_vertical': lambda t: ((np.pi / 2) * (1 - 1.5 * np.abs(np.cos(k * t))), (np.pi / 2) * np.sign(np.cos(k * t))),
        'passing': lambda t: (0, np.arctan(12 * np.cos(2 * k * t))),
        'spiral': lambda t: ((-np.pi / 4) + (3 * np.pi / 4) * (t / (fs * length)), 2 * np.pi * t * turns / (fs * length)),
    }


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog='binaural_audio_synthesis_b200', description=__doc__.split('\n')[0])
    ap.add_argument('input', help='input wav file (mono, or stereo folded to mono)')
    ap.add_argument('--bank', default='irs_and_delaydiffs_compensated_6.mat', help='file written by upsample_irs.m (apply_hrtf.py:602)')
    ap.add_argument('--trajectory', default='passing', choices=sorted(trajectories(1.0)))
    ap.add_argument('--samples-to-keep', type=int, default=100)     # apply_hrtf.py:595
    ap.add_argument('--chunksize', type=int, default=512)           # :597
    ap.add_argument('--subchunksize', type=int, default=32)         # :598
    ap.add_argument('--output', default=None)
    try:
        args = ap.parse_args(argv)
    except SystemExit as e:                                          # the reference exits 1 without an input file (:570-574)
        return 1 if e.code else 0
    from scipy.io import wavfile
    from . import apply_hrtf
    fs, y = wavfile.read(args.input)
    y = y.astype(np.float32) / y.max()                               # :576-577
    start = time.time()
    bank = apply_hrtf.load_irs_and_delaydiffs(args.bank, samples_to_keep=args.samples_to_keep)
    if len(y.shape) == 2 and y.shape[1] == 2:
        y = 0.5 * y[:, 0] + 0.5 * y[:, 1]                            # :630
    traj = trajectories(fs)[args.trajectory]
    out_sig = apply_hrtf.make_signal_move_2d(y, args.chunksize, args.subchunksize, traj, bank).astype(np.float32)
    out_filename = args.output or '{}-c{}-s{}-l{}.wav'.format(args.input.replace('.wav', ''), args.chunksize,
                                                              args.subchunksize, args.samples_to_keep)      # :636-639
    wavfile.write(out_filename, fs, np.ascontiguousarray(out_sig))
    elapsed = time.time() - start
    print("wrote to '{}' - took {:.2f} secs - {:.2f}x as fast as real time".format(out_filename, elapsed, (y.size / fs) / elapsed))
    return 0


if __name__ == '__main__':
    sys.exit(main())

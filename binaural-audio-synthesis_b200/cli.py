"""Command line of the renderer: the reference's `main` (apply_hrtf.py:559-649) on the CUDA path.

    python -m binaural_audio_synthesis_b200 INPUT.wav [--bank FILE.mat] [--trajectory NAME]
                                            [--samples-to-keep K] [--chunksize C] [--subchunksize S] [--adaptive]
    python -m binaural_audio_synthesis_b200 --scene SCENE.json [--bank FILE.mat] [--output OUT.wav]

Like the reference it reads a wav file, scales it by its maximum (:576-577), folds stereo to mono
(:630), renders it along a trajectory (default `passing`, :633) and writes
INPUT-c<C>-s<S>-l<K>.wav as float32 (:636-640), then prints the speed relative to real time (:642-646).
The reference hard-codes the bank file name and the trajectory; here they are options with the
reference's values as defaults.

--adaptive picks chunksize / subchunksize from the trajectory's speed (apply_hrtf.py:383-385, the reference's "idea
for the future"; apply_hrtf.suggest_chunk_sizes).

--scene renders SEVERAL sources into one binaural mix: a JSON file

    {"output": "mix.wav", "chunksize": 512, "subchunksize": 32, "samples_to_keep": 256,
     "sources": [{"wav": "voice.wav", "trajectory": "circle_horizontal", "gain": 0.5, "period": 4.0},
                 {"wav": "steps.wav", "trajectory": "passing", "gain": 1.0, "delay": 1.5}]}

Every source is read like the single input above (scaled by its maximum, stereo folded), multiplied by its gain,
delayed by `delay` seconds, padded to the longest; all sources are rendered and mixed in one call
(render_sources(mix=True): each source normalised on its own like a make_signal_move_2d call, then summed) and the
mix is written as float32.  All wav files must share one sample rate.
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


def read_mono(path):
    """wav -> (fs, float32 mono scaled by its maximum), apply_hrtf.py:576-577 and :630."""
    from scipy.io import wavfile
    fs, y = wavfile.read(path)
    y = y.astype(np.float32) / y.max()
    if len(y.shape) == 2 and y.shape[1] == 2:
        y = 0.5 * y[:, 0] + 0.5 * y[:, 1]
    return fs, np.ascontiguousarray(y, dtype=np.float32)


def render_scene(scene: dict, bank_file: str, output=None) -> str:
    """Render the sources of a scene description (see the module docstring) into one binaural wav."""
    from scipy.io import wavfile
    from . import apply_hrtf
    keep = int(scene.get('samples_to_keep', 100))
    chunk, sub = int(scene.get('chunksize', 512)), int(scene.get('subchunksize', 32))
    bank = apply_hrtf.load_irs_and_delaydiffs(bank_file, samples_to_keep=keep)
    signals, trajs, rate = [], [], None
    for src in scene['sources']:
        fs, y = read_mono(src['wav'])
        if rate is None:
            rate = fs
        if fs != rate:
            raise ValueError('all sources of a scene must share one sample rate (%s has %d, expected %d)' % (src['wav'], fs, rate))
        y = float(src.get('gain', 1.0)) * y
        lead = int(round(float(src.get('delay', 0.0)) * fs))
        signals.append(np.concatenate([np.zeros(lead, dtype=np.float32), y]))
        table = trajectories(fs, period=float(src.get('period', 4.0)), length=float(src.get('length', 30.0)), turns=float(src.get('turns', 15.0)))
        fn = table[src.get('trajectory', 'passing')]
        trajs.append((lambda t, fn=fn, lead=lead: fn(t - lead)) if lead else fn)
    n = max(len(y) for y in signals)
    x = np.zeros((len(signals), n), dtype=np.float32)
    for i, y in enumerate(signals):
        x[i, :len(y)] = y
    mix = apply_hrtf.render_sources(x, chunk, sub, trajs, bank, mix=True)          # (2, N_out)
    out = output or scene.get('output', 'scene-mix.wav')
    wavfile.write(out, rate, np.ascontiguousarray(mix.T.astype(np.float32)))
    return out


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog='binaural_audio_synthesis_b200', description=__doc__.split('\n')[0])
    ap.add_argument('input', nargs='?', help='input wav file (mono, or stereo folded to mono)')
    ap.add_argument('--scene', default=None, help='JSON scene description: several sources mixed into one output')
    ap.add_argument('--adaptive', action='store_true', help='choose chunksize / subchunksize from the trajectory speed')
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
    if args.scene:
        import json
        start = time.time()
        with open(args.scene) as f:
            scene = json.load(f)
        out = render_scene(scene, args.bank, args.output)
        print("wrote to '{}' - {} sources - took {:.2f} secs".format(out, len(scene['sources']), time.time() - start))
        return 0
    if not args.input:
        return 1                                                     # the reference exits 1 without an input file (:570-574)
    from scipy.io import wavfile
    from . import apply_hrtf
    fs, y = wavfile.read(args.input)
    y = y.astype(np.float32) / y.max()                               # :576-577
    start = time.time()
    bank = apply_hrtf.load_irs_and_delaydiffs(args.bank, samples_to_keep=args.samples_to_keep)
    if len(y.shape) == 2 and y.shape[1] == 2:
        y = 0.5 * y[:, 0] + 0.5 * y[:, 1]                            # :630
    traj = trajectories(fs)[args.trajectory]
    if args.adaptive:
        args.chunksize, args.subchunksize = apply_hrtf.suggest_chunk_sizes(traj, y.size)
    out_sig = apply_hrtf.make_signal_move_2d(y, args.chunksize, args.subchunksize, traj, bank).astype(np.float32)
    out_filename = args.output or '{}-c{}-s{}-l{}.wav'.format(args.input.replace('.wav', ''), args.chunksize,
                                                              args.subchunksize, args.samples_to_keep)      # :636-639
    wavfile.write(out_filename, fs, np.ascontiguousarray(out_sig))
    elapsed = time.time() - start
    print("wrote to '{}' - took {:.2f} secs - {:.2f}x as fast as real time".format(out_filename, elapsed, (y.size / fs) / elapsed))
    return 0


if __name__ == '__main__':
    sys.exit(main())

"""Multi-rank host logic (source sharding + SUM reduce, time sharding with halo + MAX reduce) on
two CPU ranks with the gloo backend.  The local renderer is the float64 oracle, injected through
`local_render`; the product's default local renderer is the CUDA path (no CPU fallback)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _bank():
    sys.path.insert(0, ROOT)
    from tests.conftest import GoldenBank
    g = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_vectors.npz')))
    return GoldenBank(g['bank_upsampling'], g['bank_diffs_left'], g['bank_diffs_right'], g['bank_irs_left'], g['bank_irs_right'])


def _traj(seed):
    rng = np.random.default_rng(seed)
    f1, f2, p1, p2 = rng.uniform(0.5, 3.0), rng.uniform(0.5, 3.0), rng.uniform(0, 6), rng.uniform(0, 6)

    def fn(t):
        return (np.float64(0.6 * np.sin(f1 * 0.003 * t + p1)), np.float64((f2 * 0.004 * t + p2) % (2 * np.pi)))
    return fn


def oracle_render(signals, chunksize, subchunksize, trajectories, bank, mix=False, normalise=True,
                  return_device=False, time_range=None, **_):
    """Same contract as apply_hrtf.render_sources, computed by the oracle on the CPU."""
    from oracle import binaural_oracle as oracle
    outs = []
    for x, fn in zip(np.asarray(signals), trajectories):
        k, n_in, _ = oracle.render_geometry(x.size, chunksize, subchunksize, bank)
        filters = oracle.boundary_filters(bank, fn, n_in, chunksize)
        y = oracle.render_unnormalised(x, chunksize, subchunksize, filters, k)
        outs.append(oracle.finish(y).T if normalise else y.astype(np.float32))
    out = np.stack(outs)
    if time_range is not None:
        out = out[:, :, time_range[0]:time_range[1]]
    if mix:
        out = out.astype(np.float64).sum(axis=0).astype(np.float32)
    return torch.from_numpy(np.ascontiguousarray(out))


def _worker(rank, world, port, mode, result_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import binaural_audio_synthesis_b200 as bas
    bank = _bank()
    rng = np.random.default_rng(99)
    if mode == 'source':
        n_src, n = 5, 2200
        x = (0.05 * rng.standard_normal((n_src, n))).astype(np.float32)
        mine = bas.distributed.shard_sources(n_src, rank, world)
        out = bas.distributed.render_mix_by_source(x[mine], 512, 32, [_traj(s) for s in mine], bank,
                                                   local_render=oracle_render)
        np.save(os.path.join(result_dir, 'source_%d.npy' % rank), out.numpy())
    elif mode == 'source_dst':
        n_src, n = 3, 1500
        x = (0.05 * rng.standard_normal((n_src, n))).astype(np.float32)
        mine = bas.distributed.shard_sources(n_src, rank, world)
        out = bas.distributed.render_mix_by_source(x[mine], 512, 32, [_traj(s) for s in mine], bank, dst=0,
                                                   local_render=oracle_render)
        assert (out is None) == (rank != 0)
        if rank == 0:
            np.save(os.path.join(result_dir, 'source_dst.npy'), out.numpy())
    else:
        n = 5000
        scale = 6.0 if mode == 'time_loud' else 0.05
        x = (scale * rng.standard_normal(n)).astype(np.float32)
        out = bas.distributed.render_by_time(x, 512, 32, _traj(7), bank, local_render=oracle_render)
        np.save(os.path.join(result_dir, '%s_%d.npy' % (mode, rank)), out)
    dist.destroy_process_group()


def _run(mode, tmp_path, world=2):
    mp.spawn(_worker, args=(world, _free_port(), mode, str(tmp_path)), nprocs=world, join=True)


def test_partition_helpers():
    sys.path.insert(0, ROOT)
    import binaural_audio_synthesis_b200 as bas
    d = bas.distributed
    assert d.shard_sources(5, 0, 2) == [0, 2, 4] and d.shard_sources(5, 1, 2) == [1, 3]
    assert sorted(sum((d.shard_sources(64, r, 8) for r in range(8)), [])) == list(range(64))
    segs = d.time_segments(n_in=5120, chunksize=512, ir_length=32, world=4)
    assert segs[0][0] == 0 and segs[-1][1] == 5120 + 31
    assert all(a[1] == b[0] for a, b in zip(segs, segs[1:])) and all(s[0] % 512 == 0 for s in segs)
    assert d.time_segments(1024, 512, 32, 8)[-1][1] == 1024 + 31          # more ranks than chunks
    assert d.segment_inputs(1536, 3072, 5120, 512, 32) == (1024, 3072)    # K-1 halo widened to a chunk
    assert d.segment_inputs(0, 1536, 5120, 512, 32) == (0, 1536)
    assert d.segment_inputs(4096, 5151, 5120, 512, 32) == (3584, 5120)    # last rank renders the tail
    assert d.segment_inputs(100, 100, 5120, 512, 32) == (0, 0)


def test_source_sharded_mix_equals_single_rank(tmp_path):
    _run('source', tmp_path)
    a = np.load(tmp_path / 'source_0.npy')
    b = np.load(tmp_path / 'source_1.npy')
    assert np.array_equal(a, b)                                           # all_reduce: same mix everywhere
    rng = np.random.default_rng(99)
    x = (0.05 * rng.standard_normal((5, 2200))).astype(np.float32)
    want = oracle_render(x, 512, 32, [_traj(s) for s in range(5)], _bank(), mix=True).numpy()
    assert np.abs(a - want).max() <= 1e-6 * np.abs(want).max()


def test_source_sharded_reduce_to_rank0(tmp_path):
    _run('source_dst', tmp_path)
    rng = np.random.default_rng(99)
    x = (0.05 * rng.standard_normal((3, 1500))).astype(np.float32)
    want = oracle_render(x, 512, 32, [_traj(s) for s in range(3)], _bank(), mix=True).numpy()
    got = np.load(tmp_path / 'source_dst.npy')
    assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max()


@pytest.mark.parametrize('mode', ['time', 'time_loud'])
def test_time_sharded_equals_single_rank(tmp_path, mode):
    """Segments cut at chunk boundaries with a K-1 halo reproduce the one-process render, including
    the global peak normalisation (apply_hrtf.py:462-464) in the 'loud' case."""
    _run(mode, tmp_path)
    from oracle import binaural_oracle as oracle
    rng = np.random.default_rng(99)
    x = ((6.0 if mode == 'time_loud' else 0.05) * rng.standard_normal(5000)).astype(np.float32)
    want = oracle.make_signal_move_2d(x, 512, 32, _traj(7), _bank())
    a = np.load(tmp_path / ('%s_0.npy' % mode))
    b = np.load(tmp_path / ('%s_1.npy' % mode))
    assert a.shape == want.shape and np.array_equal(a, b)
    assert np.abs(a - want).max() <= 2e-7 * max(1.0, np.abs(want).max())
    if mode == 'time_loud':
        assert abs(np.abs(a).max() - 1.0) < 1e-6


def test_mix_segments_cover_the_output_on_the_tile_grid():
    """distributed.mix_segments: the time segments of a by-source mix (upload / render pipelining) are contiguous,
    cover [0, n_out) and start on the render tile grid."""
    import binaural_audio_synthesis_b200 as bas
    d = bas.distributed
    for n_out, n_seg in [(2646271, None), (2646271, 1), (9000, 6), (8192, 3), (1, 4), (441599, 8)]:
        segs = d.mix_segments(n_out, n_seg)
        assert segs[0][0] == 0 and segs[-1][1] == n_out
        for (a, b), (c, e) in zip(segs[:-1], segs[1:]):
            assert b == c and a < b and a % 8192 == 0
        assert len(segs) <= (n_seg or d.MIX_SEGMENTS)

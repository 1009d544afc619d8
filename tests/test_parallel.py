"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: round-robin target sharding without data-path
collectives, result gathering, max-over-ranks timing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dmpfold2_b200 import parallel as P


def test_round_robin_shards_partition_the_targets():
    for n in (0, 1, 7, 256):
        for w in (1, 2, 4, 8):
            shards = [P.targets_for_rank(n, r, w) for r in range(w)]
            assert sorted(t for s in shards for t in s) == list(range(n))
            assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    with pytest.raises(ValueError):
        P.targets_for_rank(4, 2, 2)
    assert P.world() == (0, 1)
    assert P.max_over_ranks(3.5) == 3.5
    assert P.fold_many([1, 2, 3], lambda t: t * 10) == [10, 20, 30]
    assert P.exchange_handles(b'x' * 64) == [b'x' * 64]
    assert P.fold_many_batched([1, 2, 3], lambda ts: [t * 10 for t in ts]) == [10, 20, 30]


def test_stream_pool_host_logic():
    """StreamPool: round-robin stream assignment, per-engine SM carve-out, results in input order (engine stand-in)."""
    class Fake:
        def __init__(self, i):
            self.i, self.sms, self.seen, self.closed = i, None, [], False

        def set_conv_sms(self, n):
            self.sms = n

        def fold(self, msa, tmpl, n, m):
            self.seen.append((msa, tmpl, n, m))
            return ('coords%d' % msa, 'eng%d' % self.i)

        def close(self):
            self.closed = True
    pool = P.StreamPool(streams=3, conv_sms=132, engine_factory=Fake)
    engines = list(pool.engines)
    assert [e.sms for e in engines] == [132, 132, 132]
    assert pool.assignment(7) == [0, 1, 2, 0, 1, 2, 0]
    res = pool.fold_all(list(range(7)), templates=['t%d' % i for i in range(7)], iterations=2, minsteps=5, use_cuda_streams=False)
    assert res == [('coords%d' % t, 'eng%d' % (t % 3)) for t in range(7)]
    assert engines[1].seen == [(1, 't1', 2, 5), (4, 't4', 2, 5)]
    pool.close()
    assert all(e.closed for e in engines) and pool.engines == []
    with pytest.raises(ValueError):
        P.StreamPool(streams=0, engine_factory=Fake)
    # defaults: shared vgru scans (and the dynamic conv schedule) only when several streams share the GPU; the graph
    # option reaches every engine that has one
    assert P.StreamPool(streams=1, engine_factory=Fake).scan_rows == 0 and pool.scan_rows == 384
    assert P.StreamPool(streams=1, scan_rows=256, engine_factory=Fake).scan_rows == 256
    assert P.StreamPool(streams=1, engine_factory=Fake).conv_dynamic is False and P.StreamPool(streams=2, engine_factory=Fake).conv_dynamic is True

    class FakeG(Fake):
        graph = None

        def set_graph(self, on):
            self.graph = on
    assert [e.graph for e in P.StreamPool(streams=2, graph=True, engine_factory=FakeG).engines] == [True, True]
    assert [e.graph for e in P.StreamPool(streams=2, engine_factory=FakeG).engines] == [None, None]


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world_size)
    try:
        calls = []

        def fake_fold(t):                       # stands in for Engine.fold_host: a pure function of the target
            calls.append(t)
            return {'target': t, 'coords': torch.full((3, 5, 3), float(t)), 'rank': rank}
        res = P.fold_many(list(range(5)), fake_fold)
        assert calls == P.targets_for_rank(5, rank, world_size)           # no rank folds another rank's targets
        assert [r['target'] for r in res] == [0, 1, 2, 3, 4]
        assert [r['rank'] for r in res] == [0, 1, 0, 1, 0]
        assert all(float(r['coords'][0, 0, 0]) == r['target'] for r in res)
        resb = P.fold_many_batched(list(range(5)), lambda ts: [fake_fold(t) for t in ts])
        assert [r['target'] for r in resb] == [0, 1, 2, 3, 4] and [r['rank'] for r in resb] == [0, 1, 0, 1, 0]
        calls.clear()
        calls.extend(P.targets_for_rank(5, rank, world_size))
        local = P.fold_many(list(range(5)), fake_fold, gather=False)
        assert [x is not None for x in local] == [t % world_size == rank for t in range(5)]
        handles = P.exchange_handles(bytes([rank]) * 64)                  # the one-off window-handle exchange of a StripGroup
        assert handles == [bytes([r]) * 64 for r in range(world_size)]
        slow = P.max_over_ranks(10.0 + rank)                             # device time = max over ranks
        assert slow == 10.0 + world_size - 1
        with open(os.path.join(out_dir, f'ok{rank}'), 'w') as fh:
            fh.write('ok')
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_fold_many(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.isfile(tmp_path / 'ok0') and os.path.isfile(tmp_path / 'ok1')


class _FakeEngine:
    """Records the strip calls a StripGroup makes (the real ones need a GPU): same method names and return shapes."""

    def __init__(self, rank):
        self.rank, self.calls, self.device = rank, [], torch.device('cpu')

    def strip_setup(self, rank, world, l, reserve_n=0):
        self.calls.append(('setup', rank, world, l, reserve_n))
        return bytes([rank + 1]) * 64, 1000 + rank

    def strip_attach(self, handles):
        self.calls.append(('attach', tuple(handles)))

    def strip_detach(self):
        self.calls.append(('detach',))

    def fold_strip_host(self, msa, tmpl, n, m):
        self.calls.append(('fold', msa.shape, n, m))
        return 'coords', 'confs'


def _strip_worker(rank, world_size, port, out_dir):
    import numpy as np
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world_size)
    try:
        real_sync = torch.cuda.synchronize
        torch.cuda.synchronize = lambda *a, **k: None          # no device in the CPU suite
        eng = _FakeEngine(rank)
        grp = P.StripGroup(eng)
        assert (grp.rank, grp.world) == (rank, world_size)
        both = tuple(bytes([r + 1]) * 64 for r in range(world_size))
        assert grp.fold_host(np.zeros((7, 100), np.uint8), None, 1, 2) == ('coords', 'confs')
        grp.fold_host(np.zeros((9, 100), np.uint8), None, 0, 0)            # same length: the windows are reused
        grp.fold_host(np.zeros((5, 120), np.uint8), None, 0, 0)            # new length: collective re-setup
        grp.close()
        grp.close()                                                        # idempotent
        torch.cuda.synchronize = real_sync
        assert eng.calls == [
            ('setup', rank, world_size, 100, 7), ('attach', both), ('fold', (7, 100), 1, 2), ('fold', (9, 100), 0, 0),
            ('detach',), ('setup', rank, world_size, 120, 5), ('attach', both), ('fold', (5, 120), 0, 0), ('detach',)], eng.calls
        with open(os.path.join(out_dir, f'strip_ok{rank}'), 'w') as fh:
            fh.write('ok')
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_strip_group_host_logic(tmp_path):
    """StripGroup: one-off exchange of the window handles in rank order, windows re-made only when L changes,
    detach bracketed by barriers -- checked with two gloo ranks and a recording stand-in for the engine."""
    port = _free_port()
    mp.spawn(_strip_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ['strip_ok0', 'strip_ok1']


def test_scan_batches_groups_consecutive_targets_with_equal_depth():
    """StreamPool's shared vgru scans: consecutive targets with the same N whose columns fit `scan_rows` together."""
    from dmpfold2_b200.parallel import scan_batches
    assert scan_batches([(512, 150)] * 5, 384) == [[0, 1], [2, 3]]                  # the odd one scans alone
    assert scan_batches([(512, 150)] * 5, 768) == [[0, 1, 2, 3, 4]]
    assert scan_batches([(1000, 300), (1000, 300)], 384) == []                        # one target already fills a wave
    assert scan_batches([(1000, 300), (1000, 300), (1000, 300)], 600) == [[0, 1]]
    assert scan_batches([(10, 82), (11, 82), (11, 40), (11, 500), (11, 60), (11, 70)], 384) == [[1, 2], [4, 5]]
    assert scan_batches([], 384) == [] and scan_batches([(5, 50)], 384) == []
    assert scan_batches([(5, 128), (5, 128), (5, 128), (5, 128)], 384) == [[0, 1, 2]]

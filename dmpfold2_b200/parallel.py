"""Multi-GPU plumbing.

Independent alignments are sharded one-per-GPU, one process per GPU, with NO data-path collective (SURVEY.md
section 8e): torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only to agree on the shard, to gather the
small per-target results and to take the max-over-ranks of a device time.

One LONG target can instead be halo-sharded over the ranks (BASELINE.json configs[4]): `StripGroup` exchanges the
CUDA IPC handles of the engines' windows once through torch.distributed; after that the ranks talk through peer
stores + flags inside the kernels' stream (csrc/strip.cu) and torch.distributed is not involved in a fold.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def targets_for_rank(num_targets: int, rank: int, world_size: int) -> List[int]:
    """Round-robin shard: target t belongs to rank t % world_size (equal work per rank up to one target)."""
    if not (0 <= rank < world_size):
        raise ValueError(f'rank {rank} outside world of {world_size}')
    return list(range(rank, num_targets, world_size))


def max_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    """Max of a per-rank scalar (e.g. a CUDA-event duration) over all ranks."""
    rank, ws = world()
    if ws == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or 'cpu')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def fold_many(targets: Sequence, fold_fn: Callable[[object], object], gather: bool = True) -> List[object]:
    """Fold a list of targets across the ranks of the current process group.

    Every rank calls this with the SAME `targets` list; rank r runs `fold_fn` on targets r, r+W, r+2W, ... (no
    communication while folding).  With gather=True every rank returns the full result list in target order
    (results are small: (L,5,3) coordinates and (L,) confidences); otherwise only the local results, with None
    for targets folded elsewhere.
    """
    rank, ws = world()
    mine = targets_for_rank(len(targets), rank, ws)
    local = {t: fold_fn(targets[t]) for t in mine}
    if ws == 1 or not gather:
        return [local.get(t) for t in range(len(targets))]
    parts: List[Optional[dict]] = [None] * ws
    dist.all_gather_object(parts, local)
    merged = {}
    for p in parts:
        merged.update(p)
    return [merged[t] for t in range(len(targets))]


def fold_many_batched(targets: Sequence, batch_fn: Callable[[List[object]], List[object]], gather: bool = True) -> List[object]:
    """Like fold_many, but the rank's shard is handed to `batch_fn` in ONE call (e.g. StreamPool.fold_all, which keeps
    several targets in flight on the GPU); batch_fn returns one result per local target, in order."""
    rank, ws = world()
    mine = targets_for_rank(len(targets), rank, ws)
    res = batch_fn([targets[t] for t in mine]) if mine else []
    if len(res) != len(mine):
        raise RuntimeError('batch_fn returned %d results for %d targets' % (len(res), len(mine)))
    local = dict(zip(mine, res))
    if ws == 1 or not gather:
        return [local.get(t) for t in range(len(targets))]
    parts: List[Optional[dict]] = [None] * ws
    dist.all_gather_object(parts, local)
    merged = {}
    for p in parts:
        merged.update(p)
    return [merged[t] for t in range(len(targets))]


class StreamPool:
    """Throughput mode on ONE GPU: K engines on K CUDA streams fold independent targets concurrently
    (BASELINE.json configs[2]: "one target per stream").

    A single fold is a serial chain in which the tensor-core conv owns the whole GPU ~2/3 of the time and the rest is
    latency-bound (1000 dependent vgru steps, the eigensolver, the GRUs).  With K targets in flight the latency-bound
    stages of one target run beside the convs of another.  Every engine has its own workspace; the weights are
    uploaded once per engine.  `conv_sms` < #SMs keeps a few SMs free of the persistent conv kernel so that the small
    kernels of the other streams never wait for a whole conv launch to drain.

    engine_factory(i) -> engine lets the CPU tests (and other back ends) substitute the engine; an engine needs
    fold(msa, template, iterations, minsteps) -> (coords, confs), set_conv_sms(n) and close().
    """

    def __init__(self, state_dict=None, device_index: int = 0, streams: int = 2, conv_mode: Optional[str] = None,
                 conv_sms: int = 0, engine_factory: Optional[Callable[[int], object]] = None):
        if streams < 1:
            raise ValueError('streams must be >= 1')
        if engine_factory is None:
            from .engine import Engine

            def engine_factory(i):
                return Engine(state_dict, device_index, conv_mode=conv_mode)
        self.device_index = device_index
        self.engines = [engine_factory(i) for i in range(streams)]
        if conv_sms:
            for e in self.engines:
                e.set_conv_sms(conv_sms)
        self._streams = None

    def _cuda_streams(self):
        if self._streams is None:
            dev = torch.device('cuda', self.device_index)
            self._streams = [torch.cuda.Stream(device=dev) for _ in self.engines]
        return self._streams

    def assignment(self, num_targets: int) -> List[int]:
        """Stream index of every target: round-robin, so consecutive targets are in flight together."""
        return [t % len(self.engines) for t in range(num_targets)]

    def fold_all(self, msas: Sequence, templates: Optional[Sequence] = None, iterations: int = 10, minsteps: int = 100,
                 use_cuda_streams: bool = True) -> List[Tuple[object, object]]:
        """Fold every alignment; returns [(coords, confs)] in input order.  Asynchronous w.r.t. the host until the
        final synchronisation of the streams with the caller's current stream."""
        k = len(self.engines)
        out: List[Optional[Tuple[object, object]]] = [None] * len(msas)
        if not use_cuda_streams:                     # (CPU tests / single-stream fallback of a custom engine)
            for t, msa in enumerate(msas):
                out[t] = self.engines[t % k].fold(msa, None if templates is None else templates[t], iterations, minsteps)
            return out
        if len(msas) and hasattr(self.engines[0], 'reserve'):        # no allocation (= device sync) inside the folds
            lmax, nmax = max(int(m.shape[1]) for m in msas), max(int(m.shape[0]) for m in msas)
            for e in self.engines:
                e.reserve(lmax, nmax)
        streams = self._cuda_streams()
        cur = torch.cuda.current_stream(torch.device('cuda', self.device_index))
        for s in streams:
            s.wait_stream(cur)
        for t, msa in enumerate(msas):
            with torch.cuda.stream(streams[t % k]):
                out[t] = self.engines[t % k].fold(msa, None if templates is None else templates[t], iterations, minsteps)
        for s in streams:
            cur.wait_stream(s)
        return out

    def close(self):
        for e in self.engines:
            e.close()
        self.engines = []


def exchange_handles(handle: bytes) -> List[bytes]:
    """All-gather one small bytes object per rank, in rank order (works on gloo and NCCL groups)."""
    rank, ws = world()
    if ws == 1:
        return [handle]
    parts: List[Optional[bytes]] = [None] * ws
    dist.all_gather_object(parts, handle)
    return list(parts)


class StripGroup:
    """The ranks of the current process group folding ONE target together, every L x L map split in row strips.

    Every rank constructs it with its own engine and calls fold() with the same alignment; every rank gets the full
    (L,5,3) coordinates and (L,) confidences back (the parts of the path that are not sharded are replicated).
    The windows are sized for one alignment length; fold() re-creates them (a collective step) when L changes.
    """

    def __init__(self, engine):
        self.engine = engine
        self.rank, self.world = world()
        self.l = None

    def _setup(self, l: int, n: int = 0):
        if self.l == l:
            return
        self.close()
        handle, _ = self.engine.strip_setup(self.rank, self.world, l, n)
        self.engine.strip_attach(exchange_handles(handle))
        if self.world > 1:
            dist.barrier()            # nobody starts pushing before every rank has mapped every window
        self.l = l

    def fold(self, msa, template_ca=None, iterations: int = 10, minsteps: int = 100):
        self._setup(int(msa.shape[1]), int(msa.shape[0]))
        return self.engine.fold_strip(msa, template_ca, iterations, minsteps)

    def fold_host(self, msa, template_ca=None, iterations: int = 10, minsteps: int = 100):
        self._setup(int(msa.shape[1]), int(msa.shape[0]))
        return self.engine.fold_strip_host(msa, template_ca, iterations, minsteps)

    def close(self):
        if self.l is not None:
            torch.cuda.synchronize(self.engine.device)
            if self.world > 1:
                dist.barrier()        # every rank has finished using its peers' windows
            self.engine.strip_detach()
            if self.world > 1:
                dist.barrier()
            self.l = None

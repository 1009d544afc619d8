"""Multi-GPU plumbing.

Independent alignments are sharded one-per-GPU, one process per GPU, with NO data-path collective (SURVEY.md
section 8e): torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only to agree on the shard, to gather the
small per-target results and to take the max-over-ranks of a device time.

One LONG target can instead be halo-sharded over the ranks (BASELINE.json configs[4]): `StripGroup` exchanges the
CUDA IPC handles of the engines' windows once through torch.distributed; after that the ranks talk through peer
stores + flags inside the kernels' stream (csrc/strip.cu) and torch.distributed is not involved in a fold.
"""
from __future__ import annotations

import threading
from concurrent.futures import ThreadPoolExecutor
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


def scan_batches(shapes: Sequence[Tuple[int, int]], scan_rows: int) -> List[List[int]]:
    """Which targets share one vgru scan.  shapes[t] = (N, L) of target t.  Consecutive targets (they are in flight at
    the same time) with the SAME number of sequences N are grouped while their alignment columns together stay within
    `scan_rows`; groups of one are dropped (such a fold scans its own alignment).  Returns lists of target indices."""
    groups: List[List[int]] = []
    cur: List[int] = []
    cols = 0
    for t, (n, l) in enumerate(shapes):
        if cur and (n != shapes[cur[0]][0] or cols + l > scan_rows):
            groups.append(cur)
            cur, cols = [], 0
        if l <= scan_rows:
            cur.append(t)
            cols += l
    if cur:
        groups.append(cur)
    return [g for g in groups if len(g) > 1]


class StreamPool:
    """Throughput mode on ONE GPU: K engines on K CUDA streams fold independent targets concurrently
    (BASELINE.json configs[2]: "one target per stream").

    A single fold is a serial chain in which the tensor-core conv owns the whole GPU ~2/3 of the time and the rest is
    latency-bound (the dependent vgru steps, the eigensolver, the GRUs).  With K targets in flight the small-footprint
    stages of one target (eigensolver, GRUs, minimiser) run beside the convs of another.  Every engine has its own
    workspace; the weights are uploaded once per engine.

    * conv_dynamic (default: on when streams > 1): the conv kernel claims its work units from a device counter, so a
      launch that finds SMs occupied by another target's kernels is finished by the CTAs that did start
      (dmp2_set_conv_dynamic).
    * scan_rows (default 384 = one wave of the vgru kernel when streams > 1; 0 = off, the default for one stream): the vgru scan (network.py:223-224) is N dependent
      steps that each occupy ~all SMs for ~14 us, mostly waiting, whatever the alignment length up to 384 columns.  It is
      independent per alignment column, so the pool scans the columns of consecutive targets with the same N in ONE call
      on a separate engine and hands every fold its slice (dmp2_set_vgru_input): bit-identical, and the scan costs the
      batch what it used to cost one target.
    * conv_sms < #SMs keeps a few SMs free of the persistent conv kernel.
    * graph=True replays every engine's recycling iterations from a CUDA graph (dmp2_set_graph): a third less host time per
      fold for the enqueueing threads, same device time (None = the engines' default, DMP2_GRAPH).

    engine_factory(i) -> engine lets the CPU tests (and other back ends) substitute the engine; an engine needs
    fold(msa, template, iterations, minsteps) -> (coords, confs), set_conv_sms(n) and close().
    """

    def __init__(self, state_dict=None, device_index: int = 0, streams: int = 2, conv_mode: Optional[str] = None,
                 conv_sms: int = 0, engine_factory: Optional[Callable[[int], object]] = None,
                 conv_dynamic: Optional[bool] = None, scan_rows: Optional[int] = None, graph: Optional[bool] = None):
        if streams < 1:
            raise ValueError('streams must be >= 1')
        if engine_factory is None:
            from .engine import Engine

            def engine_factory(i):
                return Engine(state_dict, device_index, conv_mode=conv_mode)
        self.device_index = device_index
        self._factory = engine_factory
        self.engines = [engine_factory(i) for i in range(streams)]
        if conv_sms:
            for e in self.engines:
                e.set_conv_sms(conv_sms)
        self.conv_dynamic = (streams > 1) if conv_dynamic is None else bool(conv_dynamic)
        if self.conv_dynamic:
            for e in self.engines:
                if hasattr(e, 'set_conv_dynamic'):
                    e.set_conv_dynamic(True)
        # shared scans pay when other streams' folds fill the GPU meanwhile; with ONE stream the scan stream only competes with
        # the fold it feeds (measured at the cfg3 shape: 53.5 ms/target with, 39.5 without -- profiles/round2_graph_replay.txt)
        self.scan_rows = (384 if streams > 1 else 0) if scan_rows is None else int(scan_rows)
        if graph is not None:
            for e in self.engines:
                if hasattr(e, 'set_graph'):
                    e.set_graph(bool(graph))
        self._scan_engine = None
        self._scan_stream = None
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
                 use_cuda_streams: bool = True, host_threads: bool = True) -> List[Tuple[object, object]]:
        """Fold every alignment; returns [(coords, confs)] in input order.  Asynchronous w.r.t. the device until the
        final synchronisation of the streams with the caller's current stream."""
        k = len(self.engines)
        out: List[Optional[Tuple[object, object]]] = [None] * len(msas)
        if not use_cuda_streams:                     # (CPU tests / single-stream fallback of a custom engine)
            for t, msa in enumerate(msas):
                out[t] = self.engines[t % k].fold(msa, None if templates is None else templates[t], iterations, minsteps)
            return out
        if not len(msas):
            return out
        dev = torch.device('cuda', self.device_index)
        shapes = [(int(m.shape[0]), int(m.shape[1])) for m in msas]
        if hasattr(self.engines[0], 'reserve'):      # no allocation (= device sync) inside the folds
            for e in self.engines:
                e.reserve(max(l for _, l in shapes), max(n for n, _ in shapes))
        streams = self._cuda_streams()
        cur = torch.cuda.current_stream(dev)
        for s in streams:
            s.wait_stream(cur)

        # ---- shared vgru scans
        batches = scan_batches(shapes, self.scan_rows) if (self.scan_rows > 0 and hasattr(self.engines[0], 'vgru')) else []
        batch_of = {t: b for b, g in enumerate(batches) for t in g}
        scans: List[Optional[tuple]] = [None] * len(batches)          # (state of the whole batch, event, {target: first row})
        posted = [threading.Event() for _ in batches]
        failure: List[BaseException] = []
        if batches:
            if self._scan_engine is None:
                self._scan_engine = self._factory(k)
                self._scan_stream = torch.cuda.Stream(device=dev, priority=-1)   # short chains of whole-GPU steps: let them through
            self._scan_engine.reserve(max(sum(shapes[t][1] for t in g) for g in batches), max(shapes[g[0]][0] for g in batches))
            self._scan_stream.wait_stream(cur)
            msas = [m if (torch.is_tensor(m) and m.device == dev) else torch.as_tensor(m, dtype=torch.uint8).to(dev) for m in msas]

        def enqueue_scans():
            torch.cuda.set_device(self.device_index)
            try:
                with torch.cuda.stream(self._scan_stream):
                    for b, g in enumerate(batches):
                        cols = torch.cat([msas[t] for t in g], dim=1).contiguous()       # [N][L1 + L2 + ...]
                        state = self._scan_engine.vgru(cols)                              # [L1 + L2 + ...][512]
                        ev = torch.cuda.Event()
                        ev.record(self._scan_stream)
                        first, row = {}, 0
                        for t in g:
                            first[t] = row
                            row += shapes[t][1]
                            state.record_stream(streams[t % k])
                        scans[b] = (state, ev, first)
                        posted[b].set()
            except BaseException as exc:             # never leave a fold thread waiting for a scan that will not come
                failure.append(exc)
                for p in posted:
                    p.set()
                raise

        def fold_one(i, t):
            vg = None
            if t in batch_of:
                b = batch_of[t]
                posted[b].wait()
                if failure:
                    raise RuntimeError('shared vgru scan failed') from failure[0]
                state, ev, first = scans[b]
                streams[i].wait_event(ev)
                vg = state[first[t]:first[t] + shapes[t][1]]
            tm = None if templates is None else templates[t]
            out[t] = self.engines[i].fold(msas[t], tm, iterations, minsteps, vgru=vg) if vg is not None else \
                self.engines[i].fold(msas[t], tm, iterations, minsteps)

        def enqueue(i):                              # everything stream i folds, in order
            torch.cuda.set_device(self.device_index)
            with torch.cuda.stream(streams[i]):
                for t in range(i, len(msas), k):
                    fold_one(i, t)

        if k == 1 or not host_threads:
            if batches:
                enqueue_scans()
            for t in range(len(msas)):               # interleaved: consecutive targets are in flight together
                with torch.cuda.stream(streams[t % k]):
                    fold_one(t % k, t)
        else:
            # One host thread per stream (a fold is several thousand kernel launches; the C calls release the GIL), plus
            # one for the scans.
            with ThreadPoolExecutor(max_workers=k + 1) as ex:
                futs = [ex.submit(enqueue_scans)] if batches else []
                futs += [ex.submit(enqueue, i) for i in range(k)]
                for f in futs:
                    f.result()
        for s in streams:
            cur.wait_stream(s)
        if batches:
            cur.wait_stream(self._scan_stream)
        return out

    def close(self):
        for e in self.engines:
            e.close()
        self.engines = []
        if self._scan_engine is not None:
            self._scan_engine.close()
            self._scan_engine = None


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

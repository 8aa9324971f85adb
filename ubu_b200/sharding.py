"""Multi-GPU use of the path: reads shard trivially, so there is NO collective on the data path.

The reference already treats its two input BAMs and its regions as independent work
(src/main/java/edu/unc/bioinf/ubu/assembly/ReAligner.java:180-184,199-200,328-349) and hands parallelism to
`bwa -t N` (Aligner.java:19,49). Here the batch is cut into fixed-size chunks in input order; chunk c belongs to
rank c mod world (round-robin keeps mixed-length inputs balanced); every rank aligns its chunks on its own GPU
through its own stream pipeline (ubu_sw_submit / ubu_sw_wait); an ordered merge releases chunk c only after c-1,
so records come out in input order. Two drivers share that logic:

  * MultiGpuAligner  -- one process, one host thread + one handle per GPU (library use);
  * align_rank / merge_ordered -- one process per GPU (torchrun); torch.distributed is used only to gather the
    per-rank results to rank 0, never inside the alignment.
"""
from __future__ import annotations

import threading
from dataclasses import dataclass
from typing import Callable, Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from .aligner import RESULT_DTYPE, Alignments
from .seqs import SeqSet


@dataclass(frozen=True)
class Chunk:
    cid: int
    lo: int
    hi: int
    rank: int


def plan_chunks(n_tasks: int, chunk: int, world: int) -> List[Chunk]:
    """Input-order chunks, chunk c -> rank c % world."""
    if chunk <= 0 or world <= 0:
        raise ValueError("chunk and world must be positive")
    return [Chunk(c, lo, min(lo + chunk, n_tasks), c % world) for c, lo in enumerate(range(0, n_tasks, chunk))]


def window_slice(windows: SeqSet, idx: Optional[np.ndarray], lo: int, hi: int) -> Tuple[SeqSet, Optional[np.ndarray]]:
    """Windows a chunk of reads needs, re-indexed from 0 when they form a contiguous range (the common layout:
    reads grouped by window); otherwise the whole window set is kept."""
    if idx is None:
        return windows.slice(lo, hi), None
    part = idx[lo:hi]
    if part.size == 0:
        return windows.slice(0, 0), part
    w0, w1 = int(part.min()), int(part.max()) + 1
    if w1 - w0 <= 2 * (hi - lo) + 16:
        return windows.slice(w0, w1), (part - np.uint32(w0)).astype(np.uint32)
    return windows, part


class OrderedMerger:
    """Reorder buffer: push(cid, payload) in any order, pop_ready() yields payloads in chunk order."""

    def __init__(self, n_chunks: int):
        self.n, self.next, self.held = n_chunks, 0, {}  # type: int, int, Dict[int, object]
        self.lock = threading.Lock()

    def push(self, cid: int, payload) -> None:
        with self.lock:
            if cid in self.held or cid < self.next or cid >= self.n:
                raise ValueError(f"chunk {cid} pushed twice or out of range")
            self.held[cid] = payload

    def pop_ready(self) -> Iterator[Tuple[int, object]]:
        while True:
            with self.lock:
                if self.next not in self.held:
                    return
                cid, payload = self.next, self.held.pop(self.next)
                self.next += 1
            yield cid, payload

    @property
    def done(self) -> bool:
        return self.next == self.n


def merge_ordered(n_tasks: int, parts: Sequence[Tuple[Chunk, np.ndarray, np.ndarray]]) -> Alignments:
    """Concatenates per-chunk (results, cigar ops) in input order, rebasing cigar_off into one pool."""
    res = np.zeros(n_tasks, dtype=RESULT_DTYPE)
    pools, base = [], 0
    seen = 0
    for ch, r, pool in sorted(parts, key=lambda p: p[0].cid):
        if ch.lo != seen:
            raise ValueError("chunks do not tile the input")
        seen = ch.hi
        r = r.copy()
        r["cigar_off"] = (r["cigar_off"].astype(np.int64) + base).astype(np.uint32)
        res[ch.lo:ch.hi] = r
        pools.append(pool)
        base += int(pool.shape[0])
    if seen != n_tasks:
        raise ValueError("chunks do not cover the input")
    return Alignments(res, np.concatenate(pools) if pools else np.zeros(0, np.uint32))


AlignFn = Callable[[SeqSet, SeqSet, Optional[np.ndarray]], Alignments]


def align_rank(reads: SeqSet, windows: SeqSet, idx: Optional[np.ndarray], rank: int, world: int, chunk: int,
               align_fn: AlignFn) -> List[Tuple[Chunk, np.ndarray, np.ndarray]]:
    """This rank's share: its chunks, aligned one after the other with align_fn (SwAligner.align on a GPU box)."""
    out = []
    for ch in plan_chunks(reads.count, chunk, world):
        if ch.rank != rank:
            continue
        w, ix = window_slice(windows, idx, ch.lo, ch.hi)
        al = align_fn(reads.slice(ch.lo, ch.hi), w, ix)
        out.append((ch, al.results, al.cigar_pool))
    return out


def gather_to_rank0(parts, n_tasks: int, dist=None) -> Optional[Alignments]:
    """One process per GPU: collect every rank's chunks on rank 0 and merge them in input order."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return merge_ordered(n_tasks, parts)
    gathered = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(parts, gathered, dst=0)
    if dist.get_rank() != 0:
        return None
    return merge_ordered(n_tasks, [p for rank_parts in gathered for p in rank_parts])


class _PinnedStage:
    """Pinned host buffers of one in-flight chunk (inputs and outputs), grown on demand and reused: uploads from and downloads
    into pinned memory are truly asynchronous, which is what lets a handle's slots overlap."""

    def __init__(self):
        self.bufs: Dict[str, np.ndarray] = {}

    def _buf(self, key: str, nbytes: int) -> np.ndarray:
        from .aligner import pinned_empty

        b = self.bufs.get(key)
        if b is None or b.nbytes < nbytes:
            b = self.bufs[key] = pinned_empty((max(nbytes, 1) * 5 // 4 + 64,), np.uint8)
        return b

    def put(self, key: str, arr: np.ndarray) -> np.ndarray:
        v = self._buf(key, arr.nbytes)[: arr.nbytes].view(arr.dtype).reshape(arr.shape)
        np.copyto(v, arr)
        return v

    def empty(self, key: str, n: int, dtype) -> np.ndarray:
        dtype = np.dtype(dtype)
        return self._buf(key, n * dtype.itemsize)[: n * dtype.itemsize].view(dtype)

    def seqset(self, key: str, s: SeqSet) -> SeqSet:
        return SeqSet(self.put(key + "p", s.packed), self.put(key + "o", s.off), self.put(key + "l", s.len),
                      None if s.nmask is None else self.put(key + "n", s.nmask), None if s.nmask is None else self.put(key + "m", s.nmask_off))


class MultiGpuAligner:
    """One process driving several GPUs: a host thread and a SwAligner handle per device, chunks dealt round-robin,
    every handle keeps `slots` chunks in flight through pinned staging buffers, results merged in input order."""

    def __init__(self, devices: Sequence[int], chunk: int = 250_000, aligner_factory=None, **params):
        from .aligner import SwAligner

        make = aligner_factory or (lambda dev: SwAligner(device=dev, slots=4, **params))
        self.devices, self.chunk = list(devices), chunk
        self.aligners = [make(d) for d in self.devices]
        self._stages: Dict[int, List[_PinnedStage]] = {}

    def close(self) -> None:
        for a in self.aligners:
            if hasattr(a, "close"):
                a.close()
        self._stages.clear()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def align(self, reads: SeqSet, windows: SeqSet, idx: Optional[np.ndarray] = None) -> Alignments:
        world = len(self.aligners)
        chunks = plan_chunks(reads.count, self.chunk, world)
        res_all = np.empty(reads.count, dtype=RESULT_DTYPE)       # chunks land in their input-order slice: the merge of the records
        pools: Dict[int, np.ndarray] = {}
        errors = []

        def work(rank: int):
            try:
                al = self.aligners[rank]
                mine = [ch for ch in chunks if ch.rank == rank]
                if not hasattr(al, "submit"):                      # plain aligner (tests): one chunk after the other
                    for ch in mine:
                        w, ix = window_slice(windows, idx, ch.lo, ch.hi)
                        r = al.align(reads.slice(ch.lo, ch.hi), w, ix)
                        res_all[ch.lo:ch.hi] = r.results; pools[ch.cid] = r.cigar_pool
                    return
                stages = self._stages.setdefault(rank, [_PinnedStage() for _ in range(al.slots)])
                inflight = []

                def finish(item):
                    t, ch, res, pool, args = item
                    try:
                        used = al.wait(t)
                        res_all[ch.lo:ch.hi] = res; pools[ch.cid] = pool[:used].copy()
                    except Exception as e:                          # CIGAR pool guess too small: redo this chunk synchronously
                        if getattr(e, "status", None) != -3:
                            raise
                        r = al.align(*args)
                        res_all[ch.lo:ch.hi] = r.results; pools[ch.cid] = r.cigar_pool

                for k, ch in enumerate(mine):
                    if len(inflight) == al.slots:
                        finish(inflight.pop(0))                    # frees stage k % slots
                    st = stages[k % al.slots]
                    w, ix = window_slice(windows, idx, ch.lo, ch.hi)
                    rs, ws = st.seqset("r", reads.slice(ch.lo, ch.hi)), st.seqset("w", w)
                    ixp = None if ix is None else st.put("i", np.ascontiguousarray(ix, dtype=np.uint32))
                    n = ch.hi - ch.lo
                    res, pool = st.empty("res", n, RESULT_DTYPE), st.empty("pool", 6 * n + 4096, np.uint32)
                    inflight.append((al.submit(rs, ws, ixp, res, pool), ch, res, pool, (rs, ws, ixp)))
                for item in inflight:
                    finish(item)
            except Exception as e:  # surfaced on the calling thread
                errors.append(e)

        threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        # CIGAR pools in chunk order, offsets rebased into the one pool
        base = 0
        ordered = []
        for ch in chunks:
            p = pools[ch.cid]
            if base:
                res_all["cigar_off"][ch.lo:ch.hi] += np.uint32(base)
            ordered.append(p)
            base += int(p.shape[0])
        return Alignments(res_all, np.concatenate(ordered) if ordered else np.zeros(0, np.uint32))

"""Deterministic synthetic workloads of the BASELINE.json shapes (BUILD-DEFINED).

RNG: numpy Generator(PCG64(seed)). Windows are i.i.d. uniform ACGT; a read is a copy of a random
stretch of its window with substitutions and indel events (geometric lengths) applied. Forward strand
only: strand handling is the caller's job in the reference (Sam2Fastq.java:154-157). No N in the
throughput sets; `with_n` sprinkles N for the parity sets.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from .seqs import SeqSet, pack_codes, pack_uniform

SEEDS = {"cfg1": 1001, "cfg2": 1002, "cfg3": 1003, "cfg4": 1004, "cfg5": 1005}


@dataclass
class Workload:
    name: str
    reads: SeqSet
    windows: SeqSet
    pair_ref_idx: Optional[np.ndarray]
    cells: int                     # sum over tasks of read_len * window_len (full-matrix cells)
    read_codes: Optional[np.ndarray] = None      # kept for oracle checks on small sets
    window_codes: Optional[np.ndarray] = None

    @property
    def n_tasks(self) -> int:
        return self.reads.count


def _mutate(rng: np.random.Generator, windows: np.ndarray, win_idx: np.ndarray, L: int, sub_rate: float,
            indel_rate: float, geom_p: float, lo: Optional[np.ndarray] = None, hi: Optional[np.ndarray] = None,
            max_events: int = 6) -> np.ndarray:
    """Reads [n, L] drawn from windows[win_idx] (all windows one length W), vectorised over reads."""
    n = win_idx.shape[0]
    W = windows.shape[1]
    # indel events
    n_ev = np.minimum(rng.binomial(L, indel_rate, size=n), max_events) if indel_rate > 0 else np.zeros(n, dtype=np.int64)
    adv = np.ones((n, L), dtype=np.int32)        # window advance per read base
    ins = np.zeros((n, L), dtype=bool)           # read base is an inserted (random) base
    total_del = np.zeros(n, dtype=np.int64)
    for e in range(int(n_ev.max()) if n else 0):
        who = np.nonzero(n_ev > e)[0]
        pos = rng.integers(1, max(L - 1, 2), size=who.shape[0])
        ln = rng.geometric(geom_p, size=who.shape[0])
        is_del = rng.random(who.shape[0]) < 0.5
        d = who[is_del]
        adv[d, pos[is_del]] += ln[is_del].astype(np.int32)
        total_del[d] += ln[is_del]
        for w_, p_, l_ in zip(who[~is_del], pos[~is_del], ln[~is_del]):    # insertions: few, short
            e_ = min(p_ + l_, L)
            ins[w_, p_:e_] = True
    adv[ins] = 0
    span = adv.sum(axis=1)                       # window bases consumed (upper bound)
    lo_ = np.zeros(n, dtype=np.int64) if lo is None else lo
    hi_ = np.full(n, W, dtype=np.int64) if hi is None else hi
    room = np.maximum(hi_ - lo_ - span, 0)
    start = lo_ + (rng.random(n) * (room + 1)).astype(np.int64)
    start = np.minimum(start, np.maximum(W - span, 0))
    wpos = start[:, None] + np.cumsum(adv, axis=1) - adv          # window index feeding each read base
    wpos = np.clip(wpos, 0, W - 1)
    reads = windows[win_idx[:, None], wpos]
    # inserted bases are random
    k = int(ins.sum())
    if k:
        reads[ins] = rng.integers(0, 4, size=k, dtype=np.uint8)
    # substitutions: always to a different base
    sub = rng.random((n, L)) < sub_rate
    k = int(sub.sum())
    if k:
        reads[sub] = (reads[sub] + rng.integers(1, 4, size=k, dtype=np.uint8)) & 3
    return reads.astype(np.uint8)


def uniform(name: str, n_reads: int, L: int, W: int, n_windows: Optional[int] = None, seed: int = 1, sub_rate: float = 0.01,
            indel_rate: float = 0.002, geom_p: float = 0.5, paired: bool = False, keep_codes: bool = False) -> Workload:
    """n_reads reads of length L vs windows of length W. paired: mates 2k, 2k+1 share window k (opposite ends)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if paired:
        n_windows = (n_reads + 1) // 2
        win_idx = (np.arange(n_reads) // 2).astype(np.int64)
    else:
        n_windows = n_reads if n_windows is None else n_windows
        win_idx = np.arange(n_reads, dtype=np.int64) if n_windows == n_reads else rng.integers(0, n_windows, size=n_reads)
    windows = rng.integers(0, 4, size=(n_windows, W), dtype=np.uint8)
    lo = hi = None
    if paired:
        mate = np.arange(n_reads) % 2
        lo = np.where(mate == 0, 0, W // 2).astype(np.int64)
        hi = np.where(mate == 0, W // 2 + L // 2, W).astype(np.int64)
    reads = _mutate(rng, windows, win_idx, L, sub_rate, indel_rate, geom_p, lo, hi)
    identity = (not paired) and n_windows == n_reads
    return Workload(name, pack_uniform(reads), pack_uniform(windows), None if identity else win_idx.astype(np.uint32),
                    int(n_reads) * L * W, reads if keep_codes else None, windows if keep_codes else None)


def cfg1(n: int = 10_000, keep_codes: bool = False) -> Workload:
    """10k x 100 bp vs 10k distinct 1 kb windows (the CPU-runnable oracle case)."""
    return uniform("cfg1", n, 100, 1000, seed=SEEDS["cfg1"], keep_codes=keep_codes)


def cfg2(n: int = 1_000_000, keep_codes: bool = False) -> Workload:
    """1M x 76 bp paired-end (500k pairs, mates share a window) vs 500 bp windows. BASELINE metric config."""
    return uniform("cfg2", n, 76, 500, seed=SEEDS["cfg2"], paired=True, keep_codes=keep_codes)


def cfg3(n: int = 5_000_000, pool: Optional[int] = None, keep_codes: bool = False) -> Workload:
    """150 bp vs 2 kb contig windows, full traceback. Windows drawn from a pool (default n/5, many reads per window)."""
    return uniform("cfg3", n, 150, 2000, n_windows=pool if pool is not None else max(n // 5, 1), seed=SEEDS["cfg3"],
                   keep_codes=keep_codes)


def cfg4(n: int = 1_000_000, seed: Optional[int] = None, tie_stress: int = 0) -> Workload:
    """Mixed 50..250 bp reads, indel-rich (3% substitutions, 2% indel events, geometric p=0.3); window = 4x read length
    rounded up to 64; one window per read. Reads of one length are generated together and packed straight into their
    input-order places (no per-read Python work). tie_stress > 0 turns that many reads (spread over the batch) into the
    tie-break stress shapes of SURVEY.md section 8d: homopolymers, short tandem repeats, reads that match their window at
    two offsets with equal score."""
    rng = np.random.Generator(np.random.PCG64(SEEDS["cfg4"] if seed is None else seed))
    lens = rng.integers(50, 251, size=n)
    wlens = ((4 * lens + 63) // 64) * 64
    rb, wb = (lens + 3) // 4, wlens // 4
    roff = np.zeros(n, dtype=np.int64); woff = np.zeros(n, dtype=np.int64)
    if n:
        np.cumsum(rb[:-1], out=roff[1:]); np.cumsum(wb[:-1], out=woff[1:])
    rpacked = np.zeros(int(rb.sum()), dtype=np.uint8); wpacked = np.zeros(int(wb.sum()), dtype=np.uint8)
    stress = np.zeros(n, dtype=np.int8)
    if tie_stress > 0 and n:
        stress[np.linspace(0, n - 1, num=min(tie_stress, n)).astype(np.int64)] = 1 + (np.arange(min(tie_stress, n)) % 3)
    cells = 0
    for L in np.unique(lens):
        who = np.nonzero(lens == L)[0]
        L = int(L); W = int(wlens[who[0]]); m = who.shape[0]
        w = rng.integers(0, 4, size=(m, W), dtype=np.uint8)
        kind = stress[who]
        for k in np.nonzero(kind)[0]:                          # few: the stress shapes are built row by row
            if kind[k] == 1:                                   # homopolymer window (every placement ties)
                w[k, :] = rng.integers(0, 4)
            elif kind[k] == 2:                                 # tandem repeat of a 2..6-mer
                u = rng.integers(0, 4, size=int(rng.integers(2, 7)), dtype=np.uint8)
                w[k, :] = np.resize(u, W)
            else:                                              # the read's source stretch occurs twice in the window
                w[k, W // 2: W // 2 + L + 8] = w[k, 4: 4 + L + 8]
        r = _mutate(rng, w, np.arange(m), L, 0.03, 0.02, 0.3, max_events=12)
        for k in np.nonzero(kind == 3)[0]:
            r[k] = w[k, 8: 8 + L]                              # exact copy: two equal-score placements
        rp, wp = pack_uniform(r), pack_uniform(w)
        rpacked[(roff[who][:, None] + np.arange(rp.packed.size // m)[None, :]).reshape(-1)] = rp.packed
        wpacked[(woff[who][:, None] + np.arange(W // 4)[None, :]).reshape(-1)] = wp.packed
        cells += m * L * W
    if max(rpacked.size, wpacked.size) >= 2 ** 32:
        raise ValueError("batch larger than 4 GiB packed; split it")
    reads = SeqSet(rpacked, roff.astype(np.uint32), lens.astype(np.uint16), None, None)
    wins = SeqSet(wpacked, woff.astype(np.uint32), wlens.astype(np.uint16), None, None)
    return Workload("cfg4", reads, wins, None, cells)


def concat(name: str, parts: Sequence[Workload]) -> Workload:
    """Several workloads as one batch (windows and their indices are offset part by part)."""
    def cat(sets):
        base = np.cumsum([0] + [s.packed.size for s in sets[:-1]])
        off = np.concatenate([(s.off.astype(np.int64) + b) for s, b in zip(sets, base)])
        if sets and int(base[-1]) + sets[-1].packed.size >= 2 ** 32:
            raise ValueError("batch larger than 4 GiB packed; split it")
        return SeqSet(np.concatenate([s.packed for s in sets]), off.astype(np.uint32), np.concatenate([s.len for s in sets]), None, None)
    assert all(p.reads.nmask is None and p.windows.nmask is None for p in parts)
    wbase = np.cumsum([0] + [p.windows.count for p in parts[:-1]])
    idx = None
    if any(p.pair_ref_idx is not None for p in parts):
        idx = np.concatenate([(p.pair_ref_idx if p.pair_ref_idx is not None else np.arange(p.n_tasks, dtype=np.uint32)).astype(np.int64) + b
                              for p, b in zip(parts, wbase)]).astype(np.uint32)
    return Workload(name, cat([p.reads for p in parts]), cat([p.windows for p in parts]), idx, sum(p.cells for p in parts))


def cfg3_sharded(n: int = 5_000_000, shard: int = 500_000, seed: Optional[int] = None, sort_by_window: bool = True, threads: int = 8) -> Workload:
    """cfg3 generated shard by shard from (seed, shard index), as SURVEY.md section 8d prescribes for the large configs: 150 bp
    reads vs 2 kb windows, one window per ~5 reads. sort_by_window puts the reads of a window next to each other, which is what
    coordinate-sorted realignment input looks like (ReAligner.java:237 sorts by coordinate before the per-region work)."""
    base = SEEDS["cfg3"] if seed is None else seed

    def one(k_lo):
        k, lo = k_lo
        m = min(shard, n - lo)
        w = uniform("cfg3", m, 150, 2000, n_windows=max(m // 5, 1), seed=base + 104729 * k)
        if sort_by_window and w.pair_ref_idx is not None:
            order = np.argsort(w.pair_ref_idx, kind="stable")
            nb = (150 + 3) // 4
            w = Workload(w.name, SeqSet(np.ascontiguousarray(w.reads.packed.reshape(m, nb)[order]).reshape(-1), w.reads.off, w.reads.len, None, None),
                         w.windows, np.ascontiguousarray(w.pair_ref_idx[order]), w.cells)
        return w

    jobs = list(enumerate(range(0, n, shard)))
    if len(jobs) > 1 and threads > 1:                          # numpy releases the GIL in the heavy parts
        from concurrent.futures import ThreadPoolExecutor

        with ThreadPoolExecutor(max_workers=min(threads, len(jobs))) as ex:
            parts = list(ex.map(one, jobs))
    else:
        parts = [one(j) for j in jobs]
    return parts[0] if len(parts) == 1 else concat("cfg3", parts)


def with_n(w: Workload, rate: float = 0.01, seed: int = 7) -> Workload:
    """Copy of a code-keeping uniform workload with N sprinkled into reads and windows (parity sets only)."""
    assert w.read_codes is not None and w.window_codes is not None
    rng = np.random.Generator(np.random.PCG64(seed))
    r = w.read_codes.copy(); wn = w.window_codes.copy()
    for a in (r, wn):                                          # ~rate of the bases, positions drawn directly (fast on 1e8 bases)
        flat = a.reshape(-1)
        flat[rng.integers(0, flat.size, size=rng.binomial(flat.size, rate))] = 4
    return Workload(w.name + "+N", pack_uniform(r), pack_uniform(wn), w.pair_ref_idx, w.cells, r, wn)

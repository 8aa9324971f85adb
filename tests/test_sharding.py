"""Host-side multi-GPU logic on CPU: chunk plan, window slicing, ordered merge, and the one-process-per-GPU gather
over torch.distributed (gloo, world_size 2). The per-chunk aligner is a stand-in built on the CPU oracle -- these
tests exercise the sharding/merge code, not the CUDA path."""
import os
import sys

import numpy as np
import pytest

import oracle_lib as O
import ubu_b200 as U
from ubu_b200 import sharding as S
from ubu_b200 import workloads as WL


def oracle_align_fn(reads, windows, idx):
    """Stand-in for SwAligner.align with identical result/pool layout."""
    q = [reads.codes(k) for k in range(reads.count)]
    r = [windows.codes(k) for k in range(windows.count)]
    res, cigs = O.align_batch(q, r, idx, nthreads=2)
    out = np.zeros(reads.count, dtype=U.RESULT_DTYPE)
    pool = []
    for k in range(reads.count):
        for f in ("score", "ref_start", "ref_end", "q_start", "q_end", "edit_distance"):
            out[k][f] = res[k][f]
        out[k]["flags"] = res[k]["flags"]
        out[k]["cigar_off"] = len(pool)
        out[k]["cigar_n"] = len(cigs[k])
        pool += [(n << 4) | o for n, o in cigs[k]]
    return U.Alignments(out, np.array(pool, dtype=np.uint32))


def same(a: U.Alignments, b: U.Alignments):
    assert len(a) == len(b)
    for f in ("score", "ref_start", "ref_end", "q_start", "q_end", "edit_distance", "cigar_n", "flags"):
        assert np.array_equal(a.results[f], b.results[f]), f
    for k in range(len(a)):
        assert a.cigar(k) == b.cigar(k)


def test_plan_and_merger_order():
    ch = S.plan_chunks(1050, 100, 4)
    assert [c.rank for c in ch] == [0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2]
    assert ch[0].lo == 0 and ch[-1].hi == 1050 and all(a.hi == b.lo for a, b in zip(ch, ch[1:]))
    m = S.OrderedMerger(4)
    m.push(2, "c"); m.push(1, "b")
    assert list(m.pop_ready()) == []
    m.push(0, "a")
    assert [p for _, p in m.pop_ready()] == ["a", "b", "c"] and not m.done
    m.push(3, "d")
    assert [p for _, p in m.pop_ready()] == ["d"] and m.done
    with pytest.raises(ValueError):
        m.push(3, "again")
    assert S.plan_chunks(0, 10, 2) == []


def test_window_slice_reindexes_contiguous_ranges():
    w = WL.uniform("t", 40, 30, 90, seed=2, paired=True)
    ws, ix = S.window_slice(w.windows, w.pair_ref_idx, 10, 20)
    assert ws.count == 5 and ix.min() == 0 and ix.max() == 4
    for k in range(10):
        assert ws.text(int(ix[k])) == w.windows.text(int(w.pair_ref_idx[10 + k]))


@pytest.mark.parametrize("world", [1, 2, 3])
def test_sharded_equals_single_pass(world):
    w = WL.uniform("t", 230, 40, 120, seed=4, paired=True)
    whole = oracle_align_fn(w.reads, w.windows, w.pair_ref_idx)
    parts = []
    for rank in reversed(range(world)):                 # ranks finish in any order
        parts += S.align_rank(w.reads, w.windows, w.pair_ref_idx, rank, world, 37, oracle_align_fn)
    same(S.merge_ordered(w.n_tasks, parts), whole)


def test_merge_rejects_gaps():
    w = WL.uniform("t", 100, 30, 80, seed=6)
    parts = S.align_rank(w.reads, w.windows, w.pair_ref_idx, 0, 2, 25, oracle_align_fn)
    with pytest.raises(ValueError):
        S.merge_ordered(w.n_tasks, parts)


def test_multigpu_driver_threads_and_ordered_merge():
    class Fake:
        def __init__(self, dev): self.dev = dev
        def align(self, r, w, ix): return oracle_align_fn(r, w, ix)
        def close(self): pass

    w = WL.uniform("t", 301, 36, 100, seed=8, paired=True)
    with S.MultiGpuAligner([0, 1, 2], chunk=40, aligner_factory=Fake) as m:
        got = m.align(w.reads, w.windows, w.pair_ref_idx)
    same(got, oracle_align_fn(w.reads, w.windows, w.pair_ref_idx))


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = WL.uniform("t", 180, 40, 110, seed=12, paired=True)
    parts = S.align_rank(w.reads, w.windows, w.pair_ref_idx, rank, world, 32, oracle_align_fn)
    merged = S.gather_to_rank0(parts, w.n_tasks, dist)
    if rank == 0:
        whole = oracle_align_fn(w.reads, w.windows, w.pair_ref_idx)
        ok = all(np.array_equal(merged.results[f], whole.results[f]) for f in ("score", "ref_start", "q_start", "edit_distance", "cigar_n"))
        ok = ok and all(merged.cigar(k) == whole.cigar(k) for k in range(len(whole)))
        q.put(bool(ok))
    else:
        assert merged is None
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_gather_and_ordered_merge():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True

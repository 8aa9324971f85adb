#!/usr/bin/env python
"""bench.py -- SW realignment throughput (GCUPS, reads/s) on B200, one process per GPU.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the SPEC oracle on the box's host cores

Headline workload = BASELINE.json configs[2], the largest single-GPU configuration (BASELINE.json's `metric` names no
config): 5M synthetic 150 bp reads vs 2 kb assembled-contig windows (1M windows, ~5 reads per window, reads grouped by
window as coordinate-sorted realignment input is), affine gap, full traceback = 1.5e12 full-matrix cells per step.
A "step" is one pass of the hot path (forward kernels + traceback kernels) over that batch.
The other single-GPU configs (cfg1, cfg2, cfg4 with its tie-stress subset, cfg2 with N) are run once each at N=1 and
reported under `configs`, every one with its time split, its dominant kernel, its fraction of the integer roofline and a
sample checked against the CPU oracle.
For N > 1 every rank aligns its own shard of the headline shape (weak scaling; reads shard trivially, no collective on the
data path -- torch.distributed carries only the barrier and the max-reduce of the timings). Rank 0 then runs the
product's multi-GPU driver (ubu_sw_multi_*: one process, a host thread + handle per GPU, ordered merge) on ONE
cfg5-shaped input over all N GPUs: `multi` = strong scaling of that path, to merged binary results and to SAM text.

JSON keys follow the driver contract; DESIGN.md section 6 says how each number is measured.
There is no reference Java Smith-Waterman: the CPU arm is this repo's SPEC oracle ("port"),
threads = all host cores, and is labelled as such.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

READ_LEN, WIN_LEN = 150, 2000
ROUND = "r02"
HEADLINE = ("cfg3: {n} synthetic 150 bp reads vs 2 kb assembled-contig windows ({nw} windows, ~5 reads per window, grouped by window), "
            "affine gap (+1/-3/5/2), score + full traceback to CIGAR (BASELINE.json configs[2], the largest single-GPU config)")
FIELDS = ("score", "ref_start", "ref_end", "q_start", "q_end", "edit_distance")


def env_int(name: str, default: int) -> int:
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_int_peak() -> dict:
    """P_int16x2 for the roofline: bench/peak_int16x2 run live (40000 iterations per mix, ~60 ms of kernels each, so the clock
    has settled), else the committed measurement."""
    exe = os.path.join(ROOT, "bench", "peak_int16x2")
    src, data = "live, 40000 iterations", None
    if os.path.exists(exe):
        try:
            data = json.loads(subprocess.run([exe, "40000"], capture_output=True, text=True, timeout=300).stdout)
        except Exception:
            data = None
    if data is None:
        src = "profiles/peak_int16x2_r01.json (committed measurement)"
        with open(os.path.join(ROOT, "profiles", "peak_int16x2_r01.json")) as f:
            data = json.load(f)
    return {"source": src, "roofline_gcups": data["roofline_gcups"], "instr_per_clk_per_sm": data["sw_mix_warp_instr_per_clk_per_sm"],
            "sm_mhz": data["observed_sm_mhz"], "alu_only_roofline_gcups": data.get("alu_only_roofline_gcups")}


def hbm_peak_gbs():
    """Measured HBM copy bandwidth of this pool's B200s: MEASURED_PEAKS.json (driver-written) when present, else the recipe's figure."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6551.4, "fallback: the figure MEASURED_PEAKS.json held when this round was built"


def ncu_traffic(kernel_key: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel on the headline workload, from this
    round's `ncu --set full` capture (profiles/traffic_<round>.json, written by tools/ncu_summary.py from the .ncu-rep)."""
    path = os.path.join(ROOT, "profiles", f"traffic_{ROUND}.json")
    if not os.path.exists(path):
        return None, None
    try:
        with open(path) as f:
            t = json.load(f)
        e = t.get(kernel_key)
        return (float(e["dram_bytes"]), e.get("source")) if e else (None, None)
    except Exception:
        return None, None


def make_headline(n_reads: int, rank: int):
    from ubu_b200 import workloads as WL

    return WL.cfg3_sharded(n_reads, seed=WL.SEEDS["cfg3"] + 7919 * rank, threads=min(os.cpu_count() or 1, 16))


def unpack_uniform(sq, lo: int, hi: int) -> np.ndarray:
    """codes [hi-lo, L] of the uniform-length, N-free sequences lo..hi-1 of a SeqSet."""
    L = int(sq.len[lo]); nb = (L + 3) // 4
    b = sq.packed[int(sq.off[lo]): int(sq.off[lo]) + (hi - lo) * nb].reshape(hi - lo, nb)
    return ((b[:, :, None] >> np.array([0, 2, 4, 6], dtype=np.uint8)) & 3).reshape(hi - lo, nb * 4)[:, :L].astype(np.uint8)


def cpu_arm(w, n_sample: int, threads: int):
    """Times the SPEC oracle (C, pthreads over tasks) on the first n_sample reads of the headline workload.
    Returns (gcups, reads/s, seconds, oracle records)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O

    n = min(n_sample, w.n_tasks)
    q = np.ascontiguousarray(unpack_uniform(w.reads, 0, n)).reshape(-1)
    idx = None if w.pair_ref_idx is None else np.ascontiguousarray(w.pair_ref_idx[:n], dtype=np.uint32)
    nwin = int(idx.max()) + 1 if idx is not None else n
    r = np.ascontiguousarray(unpack_uniform(w.windows, 0, nwin)).reshape(-1)
    qlen = np.full(n, READ_LEN, dtype=np.uint32); rlen = np.full(nwin, WIN_LEN, dtype=np.uint32)
    qoff = (np.arange(n, dtype=np.uint64) * READ_LEN); roff = (np.arange(nwin, dtype=np.uint64) * WIN_LEN)
    res = np.zeros(n, dtype=O.SWO_RESULT)
    stride = READ_LEN + 8
    cig = np.zeros(n * stride, dtype=np.uint32)
    p = O.SwoParams(1, -3, 5, 2)
    t0 = time.perf_counter()
    st = O.lib().swo_align_batch(q.ctypes.data, qoff.ctypes.data, qlen.ctypes.data, r.ctypes.data, roff.ctypes.data, rlen.ctypes.data,
                                 None if idx is None else idx.ctypes.data, n, C.byref(p), res.ctypes.data, cig.ctypes.data, stride, threads)
    dt = time.perf_counter() - t0
    assert st == 0
    return n * READ_LEN * WIN_LEN / dt / 1e9, n / dt, dt, res


def oracle_sample_check(w, res, n_check: int, threads: int, seed: int = 0):
    """A random sample of a run's records against the CPU oracle, every field + CIGAR. Returns (checked, mismatching)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O

    rng = np.random.default_rng(seed)
    pick = np.sort(rng.choice(w.n_tasks, size=min(n_check, w.n_tasks), replace=False))
    q = [w.reads.codes(int(k)) for k in pick]
    idx = w.pair_ref_idx
    r = [w.windows.codes(int(idx[k]) if idx is not None else int(k)) for k in pick]
    ores, ocig = O.align_batch(q, r, None, nthreads=threads)
    bad = 0
    for n, k in enumerate(pick):
        g = res.results[int(k)]
        ok = all(int(g[f]) == int(ores[n][f]) for f in FIELDS) and res.cigar(int(k)) == ocig[n]
        bad += 0 if ok else 1
    return len(pick), bad


def packing_leg(w, threads: int):
    """Host-side 2-bit packing of the step's text (reads + windows) with ubu_sw_pack_batch on all host cores."""
    from ubu_b200 import _native as N

    chars = np.frombuffer(b"ACTG", dtype=np.uint8)
    out, total_s, total_bytes = {}, 0.0, 0
    n_reads = min(w.n_tasks, 1_000_000)                     # a 1M-read slice of the step and the windows it uses
    n_wins = int(w.pair_ref_idx[:n_reads].max()) + 1 if w.pair_ref_idx is not None else n_reads
    for name, sq, n in (("reads", w.reads, n_reads), ("windows", w.windows, n_wins)):
        L = int(sq.len[0])
        text = chars[unpack_uniform(sq, 0, n)].reshape(-1)
        toff = (np.arange(n, dtype=np.uint64) * np.uint64(L))
        lens = np.full(n, L, dtype=np.uint16)
        pb, nb = (L + 3) // 4, (L + 7) // 8
        off = (np.arange(n, dtype=np.int64) * pb).astype(np.uint32); noff = (np.arange(n, dtype=np.int64) * nb).astype(np.uint32)
        packed = np.zeros(n * pb, dtype=np.uint8); nmask = np.zeros(n * nb, dtype=np.uint8)
        cnt = C.c_uint32(0)
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            rc = N.lib().ubu_sw_pack_batch(text.ctypes.data, toff.ctypes.data, lens.ctypes.data, n, packed.ctypes.data, off.ctypes.data,
                                           nmask.ctypes.data, noff.ctypes.data, threads, C.byref(cnt))
            dt = time.perf_counter() - t0
            assert rc == 0 and cnt.value == 0
            best = dt if best is None else min(best, dt)
        assert np.array_equal(packed, sq.packed[: packed.size]), "batch packer disagrees with the workload's packing"
        out[name + "_ms"] = best * 1e3
        total_s += best; total_bytes += text.size
    out.update({"ms_per_step": total_s * 1e3 * w.n_tasks / n_reads, "text_gb_per_s": total_bytes / total_s / 1e9, "reads_per_s": n_reads / total_s,
                "threads": threads, "sample": f"first {n_reads} reads and their {n_wins} windows; ms_per_step scaled to the whole step", "how": "ubu_sw_pack_batch (AVX2: 32 characters per step per thread) over the step's ASCII reads and windows, "
                "best of 3, output buffers preallocated; reported beside value / e2e, not inside them (SURVEY.md section 8d)"})
    return out


def java_baseline(n_reads: int, threads: int) -> dict:
    """baseline/SwOracle.java (a line-for-line Java restatement of SPEC.md) run iff a JDK is on PATH."""
    import shutil

    if not (shutil.which("java") and shutil.which("javac")):
        return {"status": "Java baseline not run: no JVM on this box"}
    try:
        cls = os.path.join(ROOT, "baseline", "_classes")
        os.makedirs(cls, exist_ok=True)
        subprocess.run(["javac", "-d", cls, os.path.join(ROOT, "baseline", "SwOracle.java")], check=True, capture_output=True, timeout=120)
        out = subprocess.run(["java", "-Xmx8g", "-cp", cls, "SwOracle", str(n_reads), str(READ_LEN), str(WIN_LEN), str(threads), "1003"],
                             check=True, capture_output=True, text=True, timeout=600).stdout
        d = json.loads(out.strip().split("\n")[-1])
        d["status"] = "SPEC oracle (Java), NOT a reference aligner (none exists)"
        return d
    except Exception as e:
        return {"status": f"Java baseline failed: {e!r}"}


def text_leg(al, w, U, N, threads: int, steps: int, n_reads: int, chunk: int, res_ref):
    """End to end FROM TEXT: ASCII reads and windows in host memory -> 2-bit packing (ubu_sw_pack_batch on helper threads, into
    pinned staging) -> ubu_sw_submit / ubu_sw_wait -> records in host memory; the packing of chunk c+1.. runs while chunk c is on
    the GPU. Measured on the first n_reads reads of the step (same shape) to bound the text held in memory."""
    from concurrent.futures import ThreadPoolExecutor

    from ubu_b200 import sharding as SH

    n = min(n_reads, w.n_tasks)
    chars = np.frombuffer(b"ACTG", dtype=np.uint8)
    idx_all = w.pair_ref_idx
    bounds = [(lo, min(lo + chunk, n)) for lo in range(0, n, chunk)]
    L, W = READ_LEN, WIN_LEN
    rb, wb = (L + 3) // 4, (W + 3) // 4
    chunks = []
    for lo, hi in bounds:                                     # the caller's text, chunk by chunk (reads and the windows they use)
        w0, w1 = (int(idx_all[lo:hi].min()), int(idx_all[lo:hi].max()) + 1) if idx_all is not None else (lo, hi)
        chunks.append({"lo": lo, "hi": hi, "w0": w0, "w1": w1,
                       "rtext": np.ascontiguousarray(chars[unpack_uniform(w.reads, lo, hi)]).reshape(-1),
                       "wtext": np.ascontiguousarray(chars[unpack_uniform(w.windows, w0, w1)]).reshape(-1),
                       "idx": None if idx_all is None else np.ascontiguousarray(idx_all[lo:hi] - np.uint32(w0), dtype=np.uint32)})
    max_m, max_w = max(c["hi"] - c["lo"] for c in chunks), max(c["w1"] - c["w0"] for c in chunks)
    ring = al.slots + 2

    def stage_set():
        return {"rp": U.pinned_empty((max_m * rb,), np.uint8), "wp": U.pinned_empty((max_w * wb,), np.uint8),
                "roff": (np.arange(max_m, dtype=np.int64) * rb).astype(np.uint32), "woff": (np.arange(max_w, dtype=np.int64) * wb).astype(np.uint32),
                "rlen": np.full(max_m, L, dtype=np.uint16), "wlen": np.full(max_w, W, dtype=np.uint16),
                "rtoff": np.arange(max_m, dtype=np.uint64) * np.uint64(L), "wtoff": np.arange(max_w, dtype=np.uint64) * np.uint64(W),
                "pool": U.pinned_empty((6 * max_m + 4096,), np.uint32)}
    stages = [stage_set() for _ in range(ring)]
    res_p = U.pinned_empty((n,), U.RESULT_DTYPE)
    pack_threads = max(1, threads // 2)
    lib = N.lib()

    def pack(ci, st):
        c = chunks[ci]
        m, mw = c["hi"] - c["lo"], c["w1"] - c["w0"]
        cnt = C.c_uint32(0)
        rc = lib.ubu_sw_pack_batch(c["rtext"].ctypes.data, st["rtoff"].ctypes.data, st["rlen"].ctypes.data, m, st["rp"].ctypes.data, st["roff"].ctypes.data,
                                   None, None, pack_threads, C.byref(cnt))
        rc |= lib.ubu_sw_pack_batch(c["wtext"].ctypes.data, st["wtoff"].ctypes.data, st["wlen"].ctypes.data, mw, st["wp"].ctypes.data, st["woff"].ctypes.data,
                                    None, None, pack_threads, C.byref(cnt))
        assert rc == 0
        return (U.SeqSet(st["rp"][: m * rb], st["roff"][:m], st["rlen"][:m], None, None),
                U.SeqSet(st["wp"][: mw * wb], st["woff"][:mw], st["wlen"][:mw], None, None))

    def one_pass():
        with ThreadPoolExecutor(max_workers=2) as ex:
            futs = {ci: ex.submit(pack, ci, stages[ci % ring]) for ci in range(min(ring, len(chunks)))}
            inflight = []
            for ci, c in enumerate(chunks):
                if len(inflight) == al.slots:
                    t, done_ci = inflight.pop(0)
                    al.wait(t)
                    nxt = done_ci + ring                      # that chunk's staging set is free again
                    if nxt < len(chunks):
                        futs[nxt] = ex.submit(pack, nxt, stages[nxt % ring])
                rs, ws = futs.pop(ci).result()
                inflight.append((al.submit(rs, ws, c["idx"], res_p[c["lo"]:c["hi"]], stages[ci % ring]["pool"]), ci))
            for t, _ in inflight:
                al.wait(t)

    one_pass()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_pass()
    dt = (time.perf_counter() - t0) / steps
    same = all(np.array_equal(res_p[f], res_ref[f][:n]) for f in FIELDS + ("cigar_n",))
    cells = float(n) * L * W
    return {"value": cells / dt / 1e9, "unit": "GCUPS", "reads_per_s": n / dt, "ms": dt * 1e3, "reads": n,
            "text_bytes_per_pass": int(sum(c["rtext"].size + c["wtext"].size for c in chunks)), "matches_device_resident_results": bool(same),
            "how": f"ASCII text in host memory -> ubu_sw_pack_batch ({pack_threads} threads per call, two chunks being packed while earlier ones are "
                   f"on the GPU) -> ubu_sw_submit/ubu_sw_wait over {len(chunks)} chunks of {chunk} reads -> records in pinned host memory; the first {n} "
                   "reads of the step's workload"}


def run_reference_arm(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    n_sample = args.ref_sample
    w = make_headline(max(n_sample, 1000), 0)
    for _ in range(max(args.warmup, 0)):
        cpu_arm(w, min(n_sample, 5000), threads)
    times = []
    for _ in range(args.steps):
        _, _, dt, _ = cpu_arm(w, n_sample, threads)
        times.append(dt)
    dt = float(np.mean(times))
    gcups = n_sample * READ_LEN * WIN_LEN / dt / 1e9
    sample = f"{n_sample} of the {args.reads} cfg3 reads per step (the first {n_sample}, same generator and seed as the GPU arm)"
    line = {
        "impl": "reference", "metric": "SW realignment GCUPS", "value": gcups, "unit": "GCUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": HEADLINE.format(n=args.reads, nw=args.reads // 5), "reads_per_step": n_sample,
                   "note": "no Java Smith-Waterman exists in mozack/ubu (its Aligner forks bwa); this arm is this repo's SPEC "
                           "oracle (oracle/sw_oracle.c), full-matrix 32-bit + traceback, pthreads over tasks"},
        "reads_per_s": n_sample / dt,
        "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# dominant kernel of each config (profiles/launches_*_r02.txt hold the ncu launch lists these names come from)
CONFIG_KERNELS = {
    "cfg1": "sw_fwd16_kernel<8,13> (general forward, one window per read; launch-latency sized batch)",
    "cfg2": "sw_fwd16s_kernel<4,19> (forward, shared-window pairs)",
    "cfg4": "sw_fwd16_kernel<G,K> general forward over nine length classes + tb_strip_kernel (wide-band traceback)",
    "cfg2+N": "sw_fwd16s_kernel<4,19,HAS_N> (forward with window N rows) + tb_diag_kernel<HAS_N>",
}


def run_configs(al, peak_gcups: float, steps: int, flush, torch, threads: int):
    """The other single-GPU configs of BASELINE.json, device-resident, each checked against the oracle on a sample."""
    from ubu_b200 import workloads as WL

    makers = [
        ("cfg1", "10,000 x 100 bp vs 10,000 distinct 1 kb windows (BASELINE.json configs[0])", lambda: WL.cfg1(10_000), 10_000),
        ("cfg2", "1M x 76 bp paired-end (mates share a 500 bp window) (configs[1])", lambda: WL.cfg2(1_000_000), 3000),
        ("cfg4", "1M mixed 50-250 bp reads, indel-rich, window = 4x read length, 20,000-read tie-stress subset (configs[3])",
         lambda: WL.cfg4(1_000_000, tie_stress=20_000), 4000),
        ("cfg2+N", "cfg2 with 0.5 % N in reads and windows",
         lambda: WL.with_n(WL.uniform("cfg2", 1_000_000, 76, 500, seed=5, paired=True, keep_codes=True), 0.005), 3000),
    ]
    out = {}
    for name, desc, make, n_check in makers:
        t0 = time.perf_counter(); w = make(); gen_s = time.perf_counter() - t0
        al.stage(0, w.reads, w.windows, w.pair_ref_idx); al.sync(0)
        for _ in range(3):
            al.run(0)
        al.sync(0)
        ms = []
        for _ in range(steps):
            flush.zero_(); torch.cuda.synchronize()
            al.run(0); al.sync(0); ms.append(al.last_run_ms(0))
        m = np.mean(np.array(ms), axis=0)
        res = al.fetch(0, w.n_tasks)
        nchk, bad = oracle_sample_check(w, res, n_check, threads)
        value = w.cells / (m[0] * 1e-3) / 1e9
        launches = al.last_run_launches(0)
        # a batch whose traceback ran in phases underneath its forward launches: time the same staged batch once more
        # with the phases switched off, for the forward / traceback split of the kernels on their own (reported, not the value)
        split = None
        if name == "cfg4":
            os.environ["UBU_TB_PHASES"] = "1"
            try:
                al.run(0); al.sync(0)
                ms1 = []
                for _ in range(steps):
                    flush.zero_(); torch.cuda.synchronize()
                    al.run(0); al.sync(0); ms1.append(al.last_run_ms(0))
                m1 = np.mean(np.array(ms1), axis=0)
                split = {"ms_per_step": float(m1[0]), "fwd_ms": float(m1[1]), "traceback_ms": float(m1[2]),
                         "launches_per_step": al.last_run_launches(0)}
            finally:
                del os.environ["UBU_TB_PHASES"]
        fwd_alone = split["fwd_ms"] if split else m[1]
        out[name] = {"workload": desc, "tasks": w.n_tasks, "cells": int(w.cells), "value": value, "unit": "GCUPS",
                     "reads_per_s": w.n_tasks / (m[0] * 1e-3), "ms_per_step": float(m[0]), "fwd_ms": float(m[1]), "traceback_ms": float(m[2]),
                     "steps": steps, "launches_per_step": launches, "dominant_kernel": CONFIG_KERNELS[name],
                     "roofline": {"bound": "int_alu", "peak": peak_gcups, "unit": "GCUPS", "frac": value / peak_gcups,
                                  "fwd_only_frac": w.cells / (fwd_alone * 1e-3) / 1e9 / peak_gcups},
                     "oracle_check": {"checked": nchk, "mismatching": bad}, "gen_s": round(gen_s, 1)}
        if split:
            out[name]["traceback_phases"] = ("fwd_ms holds the forward launches WITH the traceback of earlier groups underneath, traceback_ms the "
                                             "exposed rest; 'unphased' = the same batch with UBU_TB_PHASES=1")
            out[name]["unphased"] = split
        del w, res
    return out


def multi_leg(n_gpus: int, n_reads: int, steps: int, warmup: int, threads: int):
    """Strong scaling of the product's multi-GPU driver: ONE cfg5-shaped input (150 bp vs 2 kb, reads grouped by window),
    sharded over n_gpus handles inside one process by ubu_sw_multi_align (chunk c -> GPU c mod N, ordered merge into the
    caller's arrays), timed by host clock to merged binary results, and again with the SAM text written behind it."""
    from ubu_b200 import multi as M
    from ubu_b200 import workloads as WL

    w = WL.cfg3_sharded(n_reads, seed=WL.SEEDS["cfg5"], threads=min(threads, 16))
    out = {"input": f"{n_reads} reads = {100.0 * n_reads / 100_000_000:.0f} % of BASELINE.json configs[4] (100M x 150 bp vs 2 kb), "
                    f"{w.windows.count} windows, generated shard by shard from (seed, shard)", "n_gpus": n_gpus, "cells": int(w.cells)}
    with M.MultiAligner(devices=list(range(n_gpus))) as ma:
        bufs = ma.pin_workload(w)
        for _ in range(max(warmup, 1)):
            ma.align_pinned(bufs)
        ts = []
        for _ in range(steps):
            t0 = time.perf_counter(); used = ma.align_pinned(bufs); ts.append(time.perf_counter() - t0)
        dt = float(np.mean(ts))
        out["to_merged_binary"] = {"value": w.cells / dt / 1e9, "unit": "GCUPS", "reads_per_s": n_reads / dt, "ms": dt * 1e3,
                                   "cigar_ops": int(used)}
        devnull = os.open(os.devnull, os.O_WRONLY)
        try:
            ts = []
            for _ in range(max(1, steps // 2)):
                t0 = time.perf_counter(); nbytes = ma.align_pinned(bufs, sam_fd=devnull); ts.append(time.perf_counter() - t0)
            dt = float(np.mean(ts))
            out["to_sam"] = {"value": w.cells / dt / 1e9, "unit": "GCUPS", "reads_per_s": n_reads / dt, "ms": dt * 1e3,
                             "sam_bytes": int(ma.last_sam_bytes), "how": "SAM text of every chunk formatted on helper threads and written "
                             "in input order to /dev/null while later chunks are still on the GPUs"}
        finally:
            os.close(devnull)
        nchk, bad = oracle_sample_check(w, bufs.alignments(), 1500, threads, seed=3)
        out["oracle_check"] = {"checked": nchk, "mismatching": bad}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=5_000_000, help="reads per GPU per step (cfg3 = 5,000,000)")
    ap.add_argument("--chunk", type=int, default=62_500, help="reads per submit() in the end-to-end leg")
    ap.add_argument("--slots", type=int, default=6, help="pipeline depth (submits in flight) of the end-to-end leg")
    ap.add_argument("--cpu-sample", type=int, default=200_000, help="reads in the bounded cpu_baseline sample (N=1)")
    ap.add_argument("--ref-sample", type=int, default=50_000, help="reads per step of the --impl reference arm")
    ap.add_argument("--multi-reads", type=int, default=10_000_000, help="reads of the cfg5-shaped input of the multi-GPU driver leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end legs (profiling runs)")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config runs (profiling runs)")
    ap.add_argument("--no-multi", action="store_true", help="skip the multi-GPU driver leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours" and not (args.no_cpu and args.no_e2e):
        args.warmup = 3            # timing hygiene: at least 3 warm-up steps

    if args.impl == "reference":
        return run_reference_arm(args)

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    import torch

    dist = None
    if world > 1:
        import torch.distributed as dist_

        dist = dist_
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        host_group = dist.new_group(backend="gloo")      # host-side waits: an NCCL barrier spins on the waiting ranks' GPUs
    else:
        torch.cuda.set_device(local)

    import ubu_b200 as U
    from ubu_b200 import _native as N

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    n = args.reads
    threads = os.cpu_count() or 1
    want_cpu = (rank == 0 and world == 1 and not args.no_cpu)
    w = make_headline(n, rank)
    cells = float(w.cells)
    al = U.SwAligner(device=local, slots=args.slots)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")      # > 126 MB L2

    sampler = ClockSampler(local)          # clocks / throttle reasons across the whole measured part (both legs)
    sampler.start()
    # ---------------- device-resident leg: inputs already in HBM ------------------------------------
    al.stage(0, w.reads, w.windows, w.pair_ref_idx)
    al.sync(0)
    for _ in range(args.warmup):
        al.run(0)
    al.sync(0)
    barrier()
    ev_ms, fwd_ms, tb_ms = [], [], []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (outside the event pair)
        torch.cuda.synchronize()
        al.run(0)
        al.sync(0)
        m = al.last_run_ms(0)
        ev_ms.append(m[0]); fwd_ms.append(m[1]); tb_ms.append(m[2])
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = al.last_run_launches(0) * args.steps
    dev_s = max_over_ranks(float(np.sum(ev_ms)) / 1e3)           # CUDA events on the launching stream, max over ranks
    total_cells = sum_over_ranks(cells) * args.steps
    total_reads = sum_over_ranks(float(n)) * args.steps
    value = total_cells / dev_s / 1e9
    fwd_s = float(np.mean(fwd_ms)) / 1e3
    res0 = al.fetch(0, n)

    # ---------------- end-to-end leg: host buffers in, host results out, every step -----------------
    e2e = None
    if not args.no_e2e:
        from ubu_b200 import sharding as SH

        chunk = min(args.chunk, n)
        bounds = [(lo, min(lo + chunk, n)) for lo in range(0, n, chunk)]
        # pinned host buffers for everything the library copies; every chunk carries only its own reads and windows
        def pin(a):
            p = U.pinned_empty(a.shape, a.dtype); p[...] = a; return p
        def pin_set(sq):
            return U.SeqSet(pin(sq.packed), pin(sq.off), pin(sq.len), None, None)
        chunks = []
        for lo, hi in bounds:
            wsl, ix = SH.window_slice(w.windows, w.pair_ref_idx, lo, hi)
            chunks.append((pin_set(w.reads.slice(lo, hi)), pin_set(wsl), None if ix is None else pin(ix)))
        res_p = U.pinned_empty((n,), U.RESULT_DTYPE)
        pool_cap = 6 * chunk + 4096
        pools = [U.pinned_empty((pool_cap,), np.uint32) for _ in range(min(len(bounds), al.slots + 1))]
        h2d = d2h = 0

        def one_pass(count_bytes: bool):
            nonlocal h2d, d2h
            inflight = []
            for ci, (lo, hi) in enumerate(bounds):
                if len(inflight) == al.slots:
                    t, c = inflight.pop(0)
                    used = al.wait(t)
                    if count_bytes:
                        d2h += used * 4
                rs, ws, ix = chunks[ci]
                t = al.submit(rs, ws, ix, res_p[lo:hi], pools[ci % len(pools)])
                inflight.append((t, ci))
                if count_bytes:
                    h2d += rs.packed.nbytes + rs.off.nbytes + rs.len.nbytes + ws.packed.nbytes + ws.off.nbytes + ws.len.nbytes
                    h2d += 0 if ix is None else ix.nbytes
                    d2h += (hi - lo) * 32
            for t, c in inflight:
                used = al.wait(t)
                if count_bytes:
                    d2h += used * 4

        for _ in range(max(min(args.warmup, 3), 1)):
            one_pass(False)
        barrier()
        t0 = time.perf_counter()
        for s in range(args.steps):
            one_pass(s == 0)
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": total_cells / e2e_s / 1e9, "unit": "GCUPS", "reads_per_s": total_reads / e2e_s,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "how": f"ubu_sw_submit/ubu_sw_wait over {len(bounds)} chunks of {chunk} reads, {al.slots} slots in flight, pinned host "
                      "buffers, each chunk carries its own reads and windows; timed by host clock between barriers"}
        # the end-to-end results must equal the device-resident ones
        same = all(np.array_equal(res_p[f], res0.results[f]) for f in FIELDS + ("cigar_n",))
        e2e["matches_device_resident_results"] = bool(same)
        del chunks, pools
        # the same measurement through ONE ubu_sw_align_batch call per step (the library cuts the batch into chunks and runs
        # them through its slots itself): what a caller gets without managing tickets
        whole_r, whole_w, whole_i = pin_set(w.reads), pin_set(w.windows), None if w.pair_ref_idx is None else pin(w.pair_ref_idx)
        pool_q = U.pinned_empty((6 * n + 4096,), np.uint32)
        rs_, ws_, used_ = whole_r.as_struct(), whole_w.as_struct(), C.c_size_t(0)

        def one_call():
            rc = N.lib().ubu_sw_align_batch(al._h, C.byref(rs_), C.byref(ws_), None if whole_i is None else whole_i.ctypes.data,
                                            res_p.ctypes.data, pool_q.ctypes.data, pool_q.shape[0], C.byref(used_))
            if rc:
                raise N.UbuSwError(rc)

        for _ in range(max(min(args.warmup, 3), 1)):
            one_call()
        barrier()
        t0 = time.perf_counter()
        for s in range(args.steps):
            one_call()
        barrier()
        oc_s = max_over_ranks(time.perf_counter() - t0)
        same_q = all(np.array_equal(res_p[f], res0.results[f]) for f in FIELDS + ("cigar_n",))
        e2e["one_call"] = {"value": total_cells / oc_s / 1e9, "unit": "GCUPS", "reads_per_s": total_reads / oc_s,
                           "how": "one ubu_sw_align_batch call per step on the whole pinned batch; chunking and slot pipeline inside the library",
                           "matches_device_resident_results": bool(same_q)}
        del whole_r, whole_w, whole_i, pool_q
        if rank == 0 and world == 1:
            try:
                e2e["from_text"] = text_leg(al, w, U, N, threads, max(2, min(args.steps, 3)), 1_000_000, min(args.chunk, n), res0.results)
            except Exception as e:
                e2e["from_text"] = {"error": repr(e)}

    clocks = sampler.stop()
    # ---------------- roofline of the dominant kernel (forward pass) ----------------------------------
    line = None
    if rank == 0:
        peak = measured_int_peak()
        fwd_gcups = cells / fwd_s / 1e9
        algo_bytes = (w.reads.packed.nbytes + w.windows.packed.nbytes + n * (4 + 2 + 4 + 4) + w.windows.count * 6 + n * 16)
        kernel_key = "sw_fwd16s_kernel<8, 19"
        traffic, traffic_src = ncu_traffic(kernel_key) if n == 5_000_000 else (None, None)
        roofline = {
            "bound": "int_alu", "kernel": "sw_fwd16s_kernel<G=8,K=19> (forward score pass, shared-window pairs; 150 bp class) -- "
                                          "the ~90 % of the reads that pair up inside their window; the odd read of a window runs on sw_fwd16_kernel<8,19> behind it",
            "achieved": fwd_gcups, "peak": peak["roofline_gcups"], "unit": "GCUPS", "frac": fwd_gcups / peak["roofline_gcups"],
            "peak_source": f"bench/peak_int16x2.cu ({peak['source']}): {peak['instr_per_clk_per_sm']:.2f} warp-instr/clk/SM on the "
                           f"4 ALU : 3 VIADD cell mix at {peak['sm_mhz']:.0f} MHz x 148 SMs x 32 lanes x 2 cells / 7 ops",
            "alu_pipe_only_peak": peak.get("alu_only_roofline_gcups"),
            "ops_per_cell": 7, "launch_ms": fwd_s * 1e3, "share_of_step": float(np.mean(fwd_ms) / np.mean(ev_ms)),
            "traffic": traffic, "traffic_unit": "bytes per step's forward launches (dram__bytes_read.sum + dram__bytes_write.sum)",
            "traffic_source": traffic_src,
            "hbm": {"algorithmic_bytes_per_launch": int(algo_bytes), "achieved_gbs": algo_bytes / fwd_s / 1e9, "peak_gbs": hbm_peak_gbs()[0],
                    "peak_source": hbm_peak_gbs()[1], "frac": algo_bytes / fwd_s / 1e9 / hbm_peak_gbs()[0], "note": "not the bound: ~5e-4 B/cell"},
        }
        line = {
            "metric": "SW realignment GCUPS", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_s * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "s16x2", "data": "synthetic",
            "config": {"workload": HEADLINE.format(n=n, nw=w.windows.count),
                       "reads_per_gpu_per_step": n, "cells_per_gpu_per_step": cells, "parallelism": f"shard{world}",
                       "l2": "256 MB device buffer rewritten between timed steps (L2 flush), outside the CUDA-event pair; the step's own "
                             "inputs (740 MB) exceed the 126 MB L2 as well"},
            "reads_per_s": total_reads / dev_s,
            "wall_s_timed_region": t_wall, "fwd_ms": float(np.mean(fwd_ms)), "traceback_ms": float(np.mean(tb_ms)),
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline,
            "step_roofline_frac": value / world / peak["roofline_gcups"],
        }
        if e2e is not None:
            line["e2e"] = e2e
        unaligned = int((res0.results["flags"] & 1).sum())
        line["sanity"] = {"unaligned": unaligned, "mean_score": float(res0.results["score"].mean())}

    if want_cpu:
        ns = min(args.cpu_sample, n)
        g, rps, dt, ores = cpu_arm(w, ns, threads)
        agree = all(np.array_equal(ores[f], res0.results[f][:ns]) for f in FIELDS)
        line["cpu_baseline"] = {"value": g, "unit": "GCUPS", "cores": threads, "kind": "port", "reads_per_s": rps,
                                "sample": f"first {ns} of the {n} reads of this run's workload, {dt:.2f} s wall; SPEC oracle "
                                          "(oracle/sw_oracle.c, full-matrix 32-bit + traceback); NOT a reference Java aligner (none exists)",
                                "gpu_results_match_on_sample": bool(agree)}
        line["packing"] = packing_leg(w, threads)
        line["java_baseline"] = java_baseline(min(args.cpu_sample, 100_000), threads)
    del res0
    if rank == 0 and world == 1 and not args.no_configs:
        del w
        line["configs"] = run_configs(al, line["roofline"]["peak"], max(3, min(args.steps, 5)), flush, torch, threads)
    al.close()
    del flush
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier(group=host_group)              # every rank's GPU is idle from here on: rank 0 drives all of them in the multi leg
    if rank == 0 and not args.no_multi and hasattr(N.lib(), "ubu_sw_multi_create"):
        try:
            line["multi"] = multi_leg(world, args.multi_reads, max(2, min(args.steps, 4)), 1, threads)
        except Exception as e:                      # the headline must still be printed
            line["multi"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier(group=host_group)
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

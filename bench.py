#!/usr/bin/env python
"""bench.py -- read k-mers/s through get_node_counts (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path over one batch: reset counters -> fused hash+probe+count of the rank's read
batch (gki_count_reads) -> get_node_counts (gki_node_counts) [-> NCCL all-reduce of the node-count vector at N>1].
Workload at N=1: c2 = BASELINE configs[1] (chr20-scale: 60M k=31 index entries, modulo 452930477, 10M x 150 bp reads,
both strands -> 2.4 G read k-mers per step); the line also carries a "c3" record: the human-scale configuration on
the same GPU.  At N>1: c3 = BASELINE configs[2] (1 B entries replicated on every rank, 37.5 M reads per GPU = 300 M
over 8, weak scaling, one NCCL all-reduce of the 50 M node counts per step).  --config overrides either.
Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: entries, nodes, modulo, reads per GPU, read length, k, p_hit (permille of reads drawn from the indexed sequence)
    "c1": dict(entries=1_000_000, nodes=100_000, modulo=19_999_999, reads=100_000, read_len=150, k=31, p_hit=100),
    "c2": dict(entries=60_000_000, nodes=6_000_000, modulo=452_930_477, reads=10_000_000, read_len=150, k=31, p_hit=100),
    "c3": dict(entries=1_000_000_000, nodes=50_000_000, modulo=452_930_477, reads=37_500_000, read_len=150, k=31, p_hit=100),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS), help="default: c2 on one GPU, c3 when the reads are sharded over several")
    ap.add_argument("--no-c3", action="store_true", help="N=1: do not attach the c3 (human-scale) record to the c2 line")
    ap.add_argument("--entries", type=int)
    ap.add_argument("--reads", type=int)
    ap.add_argument("--p-hit", type=int, dest="p_hit")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU-baseline time box")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload(args, name):
    cfg = dict(CONFIGS[name])
    if args.entries:
        cfg["entries"] = args.entries
        cfg["nodes"] = max(args.entries // 10, 1)
    if args.reads:
        cfg["reads"] = args.reads
    if args.p_hit is not None:
        cfg["p_hit"] = args.p_hit
    return cfg


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        """Summarise the samples taken in [t0, t1] (the timed region); if the region was too short to catch one,
        fall back to the samples nearest to it (the sampler runs from before the warm-up)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        lines = self.lines
        if t0 is not None:
            inside = [l for l in lines if t0 - 0.15 <= l[0] <= t1 + 0.15]
            lines = inside or sorted(lines, key=lambda l: abs(l[0] - 0.5 * (t0 + t1)))[:3]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for _, line in lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_count_rate(idx, reads, k, seconds, chunk=50_000):
    """Oracle C port (oracle/gki_oracle.c, OpenMP over all host threads) on a time-boxed sample of the same reads.
    Returns (read k-mers/s, k-mers processed, seconds, node counts of the sample's entry counters)."""
    from oracle import c_oracle
    prepared = c_oracle._index_args(idx)
    ec = np.zeros(len(prepared[2]), dtype=np.uint32)
    nk = (reads.shape[1] - k + 1) * 2
    c_oracle.count_reads(idx, reads[:1000], k, True, np.zeros_like(ec), prepared)      # warm
    done, t0 = 0, time.perf_counter()
    while done < len(reads) and time.perf_counter() - t0 < seconds:
        c_oracle.count_reads(idx, reads[done:done + chunk], k, True, ec, prepared)
        done += min(chunk, len(reads) - done)
    dt = time.perf_counter() - t0
    return done * nk / dt, done * nk, dt, done, ec


def reference_native_rate(idx, reads, k, n_nodes, seconds=6.0, chunk=2000):
    """The reference's own compiled code where it exists: cython_kmer_index.pyx built as-is into oracle/_ref (pyx:47-109, two-pass probe
    -> hit list; gates as shipped) + np.bincount over the hit nodes, fed by the np.convolve hashing of read_kmers.py:67-70 read by read
    (oracle/numpy_oracle.read_kmer_hashes), on one thread like the reference.  Time-boxed prefix of the step's reads.  None when the
    extension was not built (no reference tree at build time)."""
    import contextlib
    import glob
    import importlib.util
    import io
    from oracle import numpy_oracle as no
    so = glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle", "_ref", "cython_kmer_index*.so"))
    if not so:
        return None
    spec = importlib.util.spec_from_file_location("cython_kmer_index", so[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    class View:   # pyx:33-41 reads these attributes; hashes_to_index must be long[:]
        pass
    view = View()
    for key in ("_n_kmers", "_nodes", "_ref_offsets", "_kmers", "_frequencies", "_allele_frequencies", "_modulo"):
        setattr(view, key, idx[key])
    view._hashes_to_index = idx["_hashes_to_index"].astype(np.int64)
    with contextlib.redirect_stdout(io.StringIO()):
        cy = mod.CythonKmerIndex(view)
    counts = np.zeros(n_nodes, dtype=np.float64)
    done, t_hash, t_probe, t0 = 0, 0.0, 0.0, time.perf_counter()
    while done < len(reads) and time.perf_counter() - t0 < seconds:
        t1 = time.perf_counter()
        hashes = []
        for row in reads[done:done + chunk]:
            hashes.append(no.read_kmer_hashes(row, k))
            hashes.append(no.read_kmer_hashes(no.reverse_complement_ascii(row), k))
        hashes = np.concatenate(hashes)
        t2 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            hits = cy.get(hashes)
        counts += np.bincount(hits[0].astype(np.int64), minlength=n_nodes)[:n_nodes]
        t_hash += t2 - t1
        t_probe += time.perf_counter() - t2
        done += min(chunk, len(reads) - done)
    dt = time.perf_counter() - t0
    nk = (reads.shape[1] - k + 1) * 2
    return {"value": done * nk / dt, "unit": "kmers/s", "cores": 1, "kind": "reference",
            "hash_kmers_per_s": done * nk / max(t_hash, 1e-9), "probe_kmers_per_s": done * nk / max(t_probe, 1e-9),
            "sample": "first %d reads of the step in %.1f s: np.convolve hashing per read (read_kmers.py:67-70) + the reference's compiled "
                      "CythonKmerIndex.get (oracle/_ref, built from cython_kmer_index.pyx as-is) + np.bincount, single thread" % (done, dt)}


def run_reference(args, cfg, rank):
    """--impl reference: the reference's CPU path for the metric (oracle port; the reference itself is pure Python
    whose counting step lives in an absent third-party package) on the host cores, bounded sample per step."""
    if rank != 0:
        return
    from graph_kmer_index_b200 import synthetic
    from oracle import c_oracle
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the arm uses every core this process may run on (the other ranks exit at once)
    c_oracle.set_num_threads(cpu_threads())
    threads = c_oracle.num_threads()
    n, k, L = cfg["entries"], cfg["k"], cfg["read_len"]
    nk = (L - k + 1) * 2
    t_setup = time.perf_counter()
    big = n > 200_000_000
    hashes = nodes = None
    if big:
        # The human-scale index (1 B entries): the synthetic FlatKmers come from the device generator when a GPU is visible
        # (bit-identical to graph_kmer_index_b200/synthetic.py, tests/test_gpu_fuzz.py; numpy needs minutes and ~60 GB for them).
        # That is input synthesis only: the index is built and probed by the CPU code below.
        try:
            import torch
            from graph_kmer_index_b200 import _lib
            if torch.cuda.is_available():
                torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
                dev = torch.device("cuda")
                glen = synthetic.genome_length(n, k)
                genome = torch.empty(glen, dtype=torch.uint8, device=dev)
                _lib.call("gki_synth_genome", _lib.ptr(genome), glen, None)
                d_h = torch.empty(n, dtype=torch.int64, device=dev)
                d_n = torch.empty(n, dtype=torch.int32, device=dev)
                _lib.call("gki_synth_flat_kmers", _lib.ptr(genome), n, cfg["nodes"], k, _lib.ptr(d_h), _lib.ptr(d_n), None, None, None)
                torch.cuda.synchronize()
                hashes, nodes, codes = d_h.cpu().numpy().view(np.uint64), d_n.cpu().numpy().view(np.uint32), genome.cpu().numpy()
                del d_h, d_n, genome
                torch.cuda.empty_cache()
        except Exception as e:                              # no usable device: fall through to numpy
            sys.stderr.write("reference arm: device generator unavailable (%s), using numpy\n" % e)
            hashes = None
    if hashes is None:
        codes = synthetic.genome_codes(synthetic.genome_length(n, k))
        hashes, nodes, _, _ = synthetic.flat_kmers(n, cfg["nodes"], k, codes=codes)
    idx = c_oracle.build_index_kmers_nodes(hashes, nodes, cfg["modulo"])      # == build_index's k-mer/node columns and tables, all threads
    del hashes, nodes
    prepared = c_oracle._index_args(idx)
    n_nodes = cfg["nodes"]
    # Bounded sample: a step counts `sample_reads` of the step's reads; the per-step costs over the whole index (counter reset, node counts
    # over all entries) are charged in proportion (sample_reads / reads), so the rate is that of a whole step.  Sized from a probe so that
    # warm-up + steps end within about a minute.
    probe = synthetic.reads(20_000, L, n, k, cfg["p_hit"], codes=codes)
    ec = np.zeros(n, dtype=np.uint32)
    c_oracle.count_reads(idx, probe[:2000], k, True, ec, prepared)
    t0 = time.perf_counter()
    c_oracle.count_reads(idx, probe, k, True, ec, prepared)
    probe_rate = len(probe) / (time.perf_counter() - t0)
    budget_s = float(os.environ.get("GKI_REFERENCE_BUDGET_S", "50"))
    per_step = budget_s / max(args.warmup + args.steps, 1)
    sample_reads = int(max(20_000, min(cfg["reads"], 2_000_000, probe_rate * per_step * 0.8)))
    reads = synthetic.reads(sample_reads, L, n, k, cfg["p_hit"], codes=codes)
    t0 = time.perf_counter()
    ec[:] = 0
    c_oracle.node_counts_from_entry_counts(idx, ec, n_nodes, parallel=True, size=n_nodes)
    fixed_s = time.perf_counter() - t0                      # reset + get_node_counts over the whole index, once
    share = sample_reads / cfg["reads"]
    setup_s = time.perf_counter() - t_setup
    times = []
    for step in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        c_oracle.count_reads(idx, reads, k, True, ec, prepared)
        dt = time.perf_counter() - t0 + fixed_s * share
        if step >= args.warmup:
            times.append(dt)
    counts = c_oracle.node_counts_from_entry_counts(idx, ec, n_nodes, parallel=True, size=n_nodes)
    total = sum(times)
    value = sample_reads * nk * len(times) / total
    line = {"impl": "reference", "metric": "read_kmers_per_s_through_get_node_counts", "value": value, "unit": "kmers/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": bench_config(cfg, args, max(args.gpus, 1), sample_reads=sample_reads),
            "cpu_baseline": {"value": value, "unit": "kmers/s", "cores": threads, "kind": "port",
                             "sample": "%d of %d reads per step counted by oracle/gki_oracle.c (OpenMP, %d host threads); the per-step pass over the whole index "
                                       "(counter reset + node counts, %.2f s) is charged in the same proportion" % (sample_reads, cfg["reads"], threads, fixed_s)},
            "e2e": {"value": value, "unit": "kmers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "setup_s": setup_s, "node_count_sum": float(counts.sum())}
    if not args.no_cpu_baseline and not big:
        # reported next to the port: the port (all host threads, C) is the faster, i.e. the conservative, baseline and stays `value`
        idx_full = dict(idx)
        idx_full["_ref_offsets"] = np.zeros(n, dtype=np.uint64)
        idx_full["_frequencies"] = np.zeros(n, dtype=np.uint16)
        idx_full["_allele_frequencies"] = np.ones(n, dtype=np.float32)
        line["reference_native"] = reference_native_rate(idx_full, reads, k, n_nodes)
    print(json.dumps(line), flush=True)


def bench_config(cfg, args, world, sample_reads=None, name=None):
    c = {"workload": "%s: synthetic variant-graph index, %d k=%d entries (%d distinct k-mers x 2 nodes), modulo %d, %d nodes; "
                     "%d x %d bp reads per GPU, both strands, %.0f%% of reads drawn from the indexed sequence"
                     % (name or config_name(args, world), cfg["entries"], cfg["k"], (cfg["entries"] + 1) // 2, cfg["modulo"], cfg["nodes"], cfg["reads"],
                        cfg["read_len"], cfg["p_hit"] / 10.0),
         "index_entries": cfg["entries"], "modulo": cfg["modulo"], "n_nodes": cfg["nodes"], "reads_per_gpu": cfg["reads"],
         "read_len": cfg["read_len"], "k": cfg["k"], "kmers_per_step_per_gpu": cfg["reads"] * (cfg["read_len"] - cfg["k"] + 1) * 2,
         "parallelism": "reads sharded x%d, index replicated, 1 all-reduce of node counts" % world if world > 1 else "single GPU",
         "l2": "inputs larger than L2 (reads %.2f GB, bucket cells %.2f GB per step)" % (cfg["reads"] * cfg["read_len"] / 1e9, cfg["modulo"] * 8 / 1e9)}
    if sample_reads is not None:
        c["reads_per_step_sampled"] = sample_reads
    return c


def config_name(args, world):
    """c2 (BASELINE configs[1], the chr20-scale case the single-GPU number is quoted on) on one GPU; c3 (configs[2]: the human-scale
    index replicated, 37.5 M reads per GPU = 300 M over 8) whenever the reads are sharded over several GPUs.  --config overrides."""
    if args.config:
        return args.config
    return "c2" if world == 1 else "c3"


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback (of fallback)"


def timed(fn, repeats=2):
    """best-of device time of fn() in ms (CUDA events on the current stream), after one untimed call"""
    import torch
    fn()
    torch.cuda.synchronize()
    best = None
    for _ in range(repeats):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        best = ms if best is None else min(best, ms)
    return best


def measure(args, cfg, name, rank, world, local_rank, full):
    """One configuration through the hot path.  `full`: the main line (CPU baseline, stage rooflines, host-array build);
    otherwise the reduced record attached to the main line (c3 next to c2 at N=1)."""
    import torch
    import torch.distributed as dist
    from graph_kmer_index_b200 import CounterKmerIndex, DeviceIndex, _lib, distributed, synthetic
    from graph_kmer_index_b200.collision_free_kmer_index import DeviceCounter

    dev = torch.device("cuda", local_rank)
    n, k, L, R, modulo, n_nodes = cfg["entries"], cfg["k"], cfg["read_len"], cfg["reads"], cfg["modulo"], cfg["nodes"]
    nk_per_read = (L - k + 1) * 2
    stream = torch.cuda.current_stream().cuda_stream
    peak, peak_src = hbm_peak()

    # ---------------- setup (untimed): synthetic FlatKmers + reads on the device, index built by K2 ----------------
    glen = synthetic.genome_length(n, k)
    genome = torch.empty(glen, dtype=torch.uint8, device=dev)
    _lib.call("gki_synth_genome", _lib.ptr(genome), glen, stream)
    hashes = torch.empty(n, dtype=torch.int64, device=dev)
    nodes = torch.empty(n, dtype=torch.int32, device=dev)
    ref = torch.empty(n, dtype=torch.int64, device=dev)
    af = torch.empty(n, dtype=torch.float32, device=dev)
    _lib.call("gki_synth_flat_kmers", _lib.ptr(genome), n, n_nodes, k, _lib.ptr(hashes), _lib.ptr(nodes), _lib.ptr(ref), _lib.ptr(af), stream)
    h2i = torch.empty(modulo, dtype=torch.int32, device=dev)
    nkm = torch.empty(modulo, dtype=torch.int32, device=dev)
    s_kmers, s_nodes = torch.empty_like(hashes), torch.empty_like(nodes)
    s_ref, s_af, s_freq = torch.empty_like(ref), torch.empty_like(af), torch.empty(n, dtype=torch.int16, device=dev)

    def build_all(cols=None):   # the FlatKmers -> index call of cfki:422-467 with every column (skip_frequencies as the reference's throughput runs)
        c_h, c_n, c_r, c_a = cols or (hashes, nodes, ref, af)
        _lib.call("gki_index_build", _lib.ptr(c_h), _lib.ptr(c_n), _lib.ptr(c_r), _lib.ptr(c_a), n, modulo, _lib.GKI_BUILD_SKIP_FREQUENCIES,
                  _lib.ptr(h2i), _lib.ptr(nkm), _lib.ptr(s_kmers), _lib.ptr(s_nodes), _lib.ptr(s_ref), _lib.ptr(s_af), _lib.ptr(s_freq), None, stream)

    def build_narrow():   # k-mers + nodes only (what counting needs)
        _lib.call("gki_index_build", _lib.ptr(hashes), _lib.ptr(nodes), None, None, n, modulo, _lib.GKI_BUILD_SKIP_FREQUENCIES,
                  _lib.ptr(h2i), _lib.ptr(nkm), _lib.ptr(s_kmers), _lib.ptr(s_nodes), None, None, None, None, stream)

    build_bytes = 50.0 * n + 8.0 * modulo          # 24 B/entry in, 24 + 2 B/entry out, both dense tables
    # the same FlatKmers in the order DenseKmerFinder emits them (kmer_finder.py:223-240: the rows of a k-mer, one per node of its path,
    # follow each other; the synthetic generator shuffles them apart, the worst case for the scatter) -- timed first, the index the steps
    # below use is the one built from the shuffled rows
    j = torch.arange(n, dtype=torch.int64, device=dev)
    pos = j * synthetic.PERM_MULT          # the generator's shuffle, undone: row j was drawn from position (j * MULT + ADD) % n
    pos += synthetic.PERM_ADD
    pos %= n
    order = torch.empty_like(j)
    order[pos] = j
    del j, pos
    ordered = tuple(c[order] for c in (hashes, nodes, ref, af))
    del order
    ordered_ms = timed(lambda: build_all(ordered))
    del ordered
    torch.cuda.empty_cache()
    build_ms = timed(build_all)
    del ref, af, s_ref, s_af, s_freq
    narrow_ms = timed(build_narrow)
    narrow_bytes = 24.0 * n + 8.0 * modulo
    index_build = {"entries_per_s": n / (build_ms / 1e3), "ms": build_ms, "entries": n, "columns": "kmers, nodes, ref_offsets, allele_frequencies (+ zeroed frequencies)",
                   "roofline": {"bound": "hbm", "bytes": build_bytes, "ms": build_ms, "achieved": build_bytes / build_ms / 1e6, "peak": peak, "unit": "GB/s",
                                "frac": build_bytes / build_ms / 1e6 / peak, "model": "compulsory traffic 50 N + 8 modulo (SURVEY 8d)"},
                   "kmers_nodes_only": {"ms": narrow_ms, "bytes": narrow_bytes, "frac": narrow_bytes / narrow_ms / 1e6 / peak},
                   "rows_in_finder_order": {"ms": ordered_ms, "bytes": build_bytes, "frac": build_bytes / ordered_ms / 1e6 / peak,
                                            "what": "same entries, the rows of a k-mer adjacent as DenseKmerFinder emits them (runs of 2 here)"},
                   "note": "gki_index_build (skip_frequencies), device-resident, best of two after one warm call; slab path: append-scatter into "
                           "fixed-capacity slabs + per-slab ordering in shared memory (csrc/build.cu)"}
    part_build = None
    if world > 1:                      # hash-range partitioned build of the same FlatKmers, sharded over the ranks
        lo, hi = distributed.shard_bounds(n, rank, world)
        pb_ms = None
        for _ in range(2):
            dist.barrier()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            part = distributed.build_index_partitioned(hashes[lo:hi], nodes[lo:hi], None, None, modulo, skip_frequencies=True, replicate=False)
            ev[1].record()
            torch.cuda.synchronize()
            t = torch.tensor([ev[0].elapsed_time(ev[1])], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            pb_ms = float(t.item())
        part_build = {"entries_per_s": n / (pb_ms / 1e3), "ms": pb_ms, "entries": n, "ranks": world,
                      "note": "partition by bucket range -> NCCL all-to-all -> gki_index_build_range per rank; max over ranks"}
        del part
    index = DeviceIndex(h2i, nkm, s_kmers, s_nodes, modulo)
    t0 = time.perf_counter()
    index.prepare_counting(k)          # Bloom filter + count table (otherwise built inside the first counting call)
    torch.cuda.synchronize()
    prepare_ms = 1e3 * (time.perf_counter() - t0)
    info = index.info()
    del hashes, nodes
    torch.cuda.empty_cache()
    _lib.call("gki_release_scratch")    # the builds' slab scratch (43 GB at 1 B entries) goes back to the driver
    reads = torch.empty((R, L), dtype=torch.uint8, device=dev)
    _lib.call("gki_synth_reads", _lib.ptr(genome), glen, rank * R, R, L, cfg["p_hit"], 0, _lib.ptr(reads), stream)
    counts = torch.zeros(max(n_nodes, index.max_node + 1), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()

    def step():
        index.reset_counts()
        index.count_reads(reads, k, True)
        index.node_counts(n_nodes, out=counts)
        if world > 1:
            distributed.allreduce_node_counts(counts)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = _lib.launch_count()
    wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    ev0.record()
    for i in range(args.steps):
        index.reset_counts()
        kern_ev[i][0].record()
        index.count_reads(reads, k, True)
        kern_ev[i][1].record()
        index.node_counts(n_nodes, out=counts)
        kern_ev[i][2].record()
        if world > 1:
            distributed.allreduce_node_counts(counts)
    ev1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop(wall0, time.time()) if rank == 0 else None
    elapsed_ms = ev0.elapsed_time(ev1)
    kernel_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in kern_ev]))
    node_counts_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in kern_ev]))
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    total_kmers_per_step = R * nk_per_read * world
    value = total_kmers_per_step * args.steps / (elapsed_ms / 1e3)

    # hit statistics of the last step (for the algorithmic-bytes model) -- untimed
    total_counts = float(counts.sum().item())
    entry_hits = None
    if rank == 0:
        ec = torch.empty(n, dtype=torch.int32, device=dev)
        _lib.call("gki_entry_counts", index.handle, _lib.ptr(ec), stream)
        torch.cuda.synchronize()
        # every distinct k-mer has its counter replicated on each of its entries; hits = sum over distinct k-mers
        entry_hits = float(ec.to(torch.float64).sum().item())
        del ec

    # ---------------- e2e: host reads through the same C-ABI call, node counts read back to the host ----------------
    e2e = None
    if not args.no_e2e:
        host_reads = torch.empty((R, L), dtype=torch.uint8, pin_memory=True)
        host_reads.copy_(reads)
        host_counts = torch.empty(counts.shape[0], dtype=torch.float64, pin_memory=True)
        torch.cuda.synchronize()

        def e2e_step():
            index.reset_counts()
            index.count_reads(host_reads, k, True)                 # H2D chunks overlap the count kernels inside the call
            index.node_counts(n_nodes, out=counts)
            if world > 1:
                distributed.allreduce_node_counts(counts)
            host_counts.copy_(counts, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        def run_e2e(step_fn, settle, steps):
            for _ in range(settle):                              # the host pipeline settles its choice of mode in its first seven calls
                step_fn()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                step_fn()
            b.record()
            barrier()
            t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return total_kmers_per_step * steps / (float(t.item()) / 1e3)

        e2e_steps = max(1, min(args.steps, 10))
        e2e_value = run_e2e(e2e_step, 8, e2e_steps)
        assert float(host_counts.sum().item()) == (total_counts if world == 1 else float(counts.sum().item()))
        lanes = int(os.environ.get("GKI_PACK_THREADS", max((os.cpu_count() or 2) - 2, 0) if world == 1 else
                                   max((os.cpu_count() or 2) // int(os.environ.get("LOCAL_WORLD_SIZE", world)) - 1, 0)))
        # The ceiling of the host side, measured under contention: this rank's packing lanes sweep the batch (gki_host_read_bandwidth) WHILE
        # the copy engine moves the same pinned batch to the device -- the two ways the reads leave host memory, competing for it as they
        # do inside the call; every rank probes at the same time, as every rank packs and copies at the same time.
        gbs = ctypes.c_double()
        barrier()
        ca, cb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ca.record()
        reads.copy_(host_reads, non_blocking=True)
        cb.record()
        _lib.call("gki_host_read_bandwidth", host_reads.data_ptr(), int(R) * int(L), max(lanes, 1), ctypes.byref(gbs))
        torch.cuda.synchronize()
        both = torch.tensor([gbs.value, R * L / (ca.elapsed_time(cb) / 1e3) / 1e9], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(both)                                 # sums over the ranks
        host_read_gbs, pcie_gbs = float(both[0].item()), float(both[1].item())
        ascii_gbs = e2e_value / nk_per_read * L / 1e9
        ceiling = host_read_gbs + pcie_gbs
        e2e = {"value": e2e_value, "unit": "kmers/s",
               "h2d_bytes_per_step": int(R * L), "d2h_bytes_per_step": int(counts.shape[0] * 8), "steps": e2e_steps,
               "host_memory": "pinned", "pack_lanes": lanes,
               "host_read_gbs": host_read_gbs, "pcie_h2d_gbs": pcie_gbs, "host_ascii_gbs_consumed": ascii_gbs, "host_frac": ascii_gbs / ceiling if ceiling else None,
               "note": "per GPU bytes of the caller's ASCII reads (1 byte per base); inside the call pack_lanes host threads re-encode "
                       "chunks to 2 bits per base before the bus while the copy engine moves the other chunks as ASCII (csrc/count.cu); "
                       "host_frac = ASCII bytes consumed per second / (host-DRAM read bandwidth of the packing lanes on this batch + copy-engine "
                       "rate of the same pinned batch), the two measured here at the same time (they compete for host memory as inside the call) "
                       "and summed over the ranks (all ranks probe at once)"}
        # the call a KAGE user makes (cfki:33-40): pageable numpy reads through CounterKmerIndex, node counts returned as a fresh numpy array
        pageable = np.empty((R, L), dtype=np.uint8)
        pageable[:] = host_reads.numpy()
        counter = CounterKmerIndex(None, None, DeviceCounter(index))
        got = {}

        def e2e_pageable_step():
            counter.reset()
            counter.count_reads(pageable, k)
            if world == 1:
                got["counts"] = counter.get_node_counts(n_nodes)
            else:                                        # what distributed.count_reads_sharded does after counting its shard
                index.node_counts(n_nodes, out=counts)
                distributed.allreduce_node_counts(counts)
                got["counts"] = counts.cpu().numpy()

        pg_steps = max(1, min(args.steps, 5))
        e2e["pageable_numpy"] = {"value": run_e2e(e2e_pageable_step, 8, pg_steps), "unit": "kmers/s", "steps": pg_steps,
                                 "call": "CounterKmerIndex.reset(); .count_reads(numpy uint8 reads, k); .get_node_counts(n_nodes) -> numpy float64"
                                         if world == 1 else "CounterKmerIndex.reset(); .count_reads(numpy uint8 reads, k); node counts all-reduced on the device -> numpy float64"}
        assert float(got["counts"].sum()) == float(host_counts.sum().item())
        # callers that already hold 2-bit packed reads (read_kmers.pack_reads layout): gki_count_packed_reads on a pinned host batch
        if world == 1 or full:
            words = (L + 31) // 32
            packed = torch.empty((R, words), dtype=torch.int64, pin_memory=True)
            n_clean, n_dirty = ctypes.c_int64(), ctypes.c_int64()
            _lib.call("gki_pack_reads", host_reads.data_ptr(), R, L, L, packed.data_ptr(), None, 0, ctypes.byref(n_clean), ctypes.byref(n_dirty), 0, 0)
            if n_dirty.value == 0:
                def e2e_packed_step():
                    index.reset_counts()
                    index.count_packed_reads(packed, L, k, True)
                    index.node_counts(n_nodes, out=counts)
                    if world > 1:
                        distributed.allreduce_node_counts(counts)
                    host_counts.copy_(counts, non_blocking=True)
                    torch.cuda.current_stream().synchronize()
                e2e["packed_2bit"] = {"value": run_e2e(e2e_packed_step, 2, pg_steps), "unit": "kmers/s", "steps": pg_steps,
                                      "h2d_bytes_per_step": int(R * words * 8), "call": "gki_count_packed_reads on a pinned host batch of 2-bit rows"}
            del packed
        del host_reads, pageable

    result = {"value": value, "ms_per_step": elapsed_ms / args.steps, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
              "index_build": index_build, "index_build_partitioned": part_build,
              "stages_ms": {"count_reads_kernel": kernel_ms, "get_node_counts": node_counts_ms, "prepare_counting_once": prepare_ms},
              "index": {"device_bytes": info["device_bytes"], "has_filter": info["has_filter"], "nonempty_buckets": info["nonempty_buckets"]}}
    if rank != 0:
        return result

    # ---------------- roofline of the dominant kernel (count_reads_kernel) ----------------
    kmers_per_launch = R * nk_per_read
    # counts summed over nodes counts every hit once per entry of the k-mer (2 entries per distinct k-mer)
    h = (entry_hits / 2.0) / kmers_per_launch
    occ = info["nonempty_buckets"] / modulo
    # Algorithmic HBM bytes of one launch.  The kernel probes the count table (csrc/count.cu), not
    # the reference's modulo buckets, so SURVEY 8(d)'s sector figure is re-derived for that layout.  Two granularities are
    # reported: the 128-byte line this B200 fills L2 from HBM with on a random access (measured: 127 B of DRAM reads per 8-byte
    # gather, profiles/r1/calibrate_gather_v2_ncu.txt) -- the `achieved` figure -- and the 32-byte sector the probe actually uses.
    # Per read position (= 2 k-mers, one canonical key): its share of the ASCII read; for a hit the table line / key sector and the
    # counter sector written back; the Bloom filter word comes from L2 when the filter is resident there, else it is one more line.
    positions = kmers_per_launch / 2.0
    h_pos = 2.0 * h                                         # hit positions / positions (a position hits on one strand)
    filter_resident = info["has_filter"] and (n // 2) <= (32 << 20)
    # a filter larger than L2 is addressed by minimizer (csrc/count.cu, filter_m): consecutive windows share a line as long as
    # their minimizer does, on average (w + 1) / 2 = 9 windows for the 17 m-mers of a k-mer
    minimizer_filter = (not filter_resident) and k in (27, 29, 31) and os.environ.get("GKI_FILTER_MZ", "1") != "0"
    filter_lines = 0.0 if (filter_resident or not info["has_filter"]) else (2.0 / 18.0 if minimizer_filter else 1.0)

    def model_bytes(gran):      # bytes per read position at a fill granularity of `gran` bytes per random access
        return 2.0 * L / nk_per_read + gran * filter_lines + (gran + 32.0) * h_pos
    bytes_per_kmer = model_bytes(128.0) / 2.0
    achieved = kmers_per_launch * bytes_per_kmer / (kernel_ms / 1e3) / 1e9
    achieved_sector = kmers_per_launch * model_bytes(32.0) / 2.0 / (kernel_ms / 1e3) / 1e9
    # the same for the reference's own bucket layout (SURVEY 8(d) as written: 1 sector empty bucket / 3 miss / 5 hit)
    survey_sectors = (1 - h) * (1 - occ) * 1 + (1 - h) * occ * 3 + h * 5
    survey_bytes_per_kmer = L / nk_per_read + 32.0 * survey_sectors
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("count_reads_kernel_dram_bytes_per_launch_%s" % name)
        except Exception:
            traffic = None
    # the other two resources the kernel leans on, against ceilings measured on this pool's B200 (profiles/r1/calibrate_gather_v2.jsonl):
    # per-thread global requests through L1TEX (1 sector per cycle per SM) and random HBM line fetches
    requests = positions * (1.0 + 2.0 * h_pos + 0.05)      # filter word + (keys load + RED) per hit + ~5 % false positives
    hbm_lines = positions * (h_pos + filter_lines + 0.05)
    result["roofline"] = {
        "bound": "hbm", "kernel": "count_reads_kernel<both=true,paired=true> (Bloom filter: %s)" % ("none" if not info["has_filter"] else ("L2-resident" if filter_resident else ("HBM, minimizer-addressed" if minimizer_filter else "HBM"))),
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
        "granularity": {"line_128B": {"bytes_per_kmer": bytes_per_kmer, "achieved": achieved, "frac": achieved / peak},
                        "sector_32B": {"bytes_per_kmer": model_bytes(32.0) / 2.0, "achieved": achieved_sector, "frac": achieved_sector / peak}},
        "peak_source": peak_src, "kernel_ms": kernel_ms, "kernel_share_of_step": kernel_ms * args.steps / elapsed_ms,
        "algorithmic_bytes_per_kmer": bytes_per_kmer, "hit_fraction_of_kmers": h, "positions_per_launch": positions,
        "kmers_per_s_kernel_only": kmers_per_launch / (kernel_ms / 1e3),
        "co_limits": {"l1tex_requests_per_s": requests / (kernel_ms / 1e3), "l1tex_ceiling_per_s": 291e9,
                      "l1tex_frac": requests / (kernel_ms / 1e3) / 291e9,
                      "hbm_random_lines_per_s": hbm_lines / (kernel_ms / 1e3), "hbm_random_ceiling_per_s": 37e9,
                      "hbm_random_frac": hbm_lines / (kernel_ms / 1e3) / 37e9,
                      "source": "profiles/r1/calibrate_gather_v2.jsonl: 291 G L2-resident gathers/s (L1TEX at 99 %), 37 G HBM gathers/s"},
        "reference_layout_model": {"bytes_per_kmer": survey_bytes_per_kmer, "sectors_per_kmer": survey_sectors, "bucket_occupancy": occ,
                                   "gbs": kmers_per_launch * survey_bytes_per_kmer / (kernel_ms / 1e3) / 1e9},
        "note": "HBM bytes of the count-table layout: ASCII + per hit one table access and the counter sector written back (+ one access per "
                "position when the filter does not fit L2), at the measured 128-byte fill granularity (achieved / frac) and at 32-byte sectors; "
                "traffic = ncu dram bytes of one launch (false-positive and chain probes are the excess)"}

    # ---------------- K1 (unfused hashing API, read_kmers.py:67-70) against the HBM roofline ----------------
    if full:
        hr = min(R, 2_000_000)
        sub = reads[:hr]
        fwd = torch.empty((hr, L - k + 1), dtype=torch.uint64, device=dev)
        rcv = torch.empty((hr, L - k + 1), dtype=torch.uint64, device=dev)
        both_ms = timed(lambda: _lib.call("gki_hash_reads", _lib.ptr(sub), hr, L, L, k, _lib.ptr(fwd), _lib.ptr(rcv), stream), 3)
        fwd_ms = timed(lambda: _lib.call("gki_hash_reads", _lib.ptr(sub), hr, L, L, k, _lib.ptr(fwd), None, stream), 3)
        both_b, fwd_b = hr * (L + 2.0 * (L - k + 1) * 8), hr * (L + (L - k + 1) * 8.0)
        result["stages"] = {"k1_frac": both_b / both_ms / 1e6 / peak, "k1_fwd_only_frac": fwd_b / fwd_ms / 1e6 / peak, "k2_frac": index_build["roofline"]["frac"],
                            "k1": {"reads": hr, "fwd_rc_ms": both_ms, "fwd_rc_bytes": both_b, "fwd_only_ms": fwd_ms, "fwd_only_bytes": fwd_b,
                                   "model": "L + strands x (L - k + 1) x 8 bytes per read (SURVEY 8d)"},
                            "peak": peak, "peak_source": peak_src}
        del fwd, rcv

    # ---------------- index build end to end: the reference-facing call on host numpy arrays (cfki:422-467) ----------------
    if full and world == 1 and not args.no_e2e:
        from graph_kmer_index_b200 import CollisionFreeKmerIndex, FlatKmers
        d_h, d_n = torch.empty(n, dtype=torch.int64, device=dev), torch.empty(n, dtype=torch.int32, device=dev)
        d_r, d_a = torch.empty(n, dtype=torch.int64, device=dev), torch.empty(n, dtype=torch.float32, device=dev)
        _lib.call("gki_synth_flat_kmers", _lib.ptr(genome), n, n_nodes, k, _lib.ptr(d_h), _lib.ptr(d_n), _lib.ptr(d_r), _lib.ptr(d_a), stream)
        flat = FlatKmers(d_h.cpu().numpy().view(np.uint64), d_n.cpu().numpy().view(np.uint32), d_r.cpu().numpy().view(np.uint64), d_a.cpu().numpy())
        del d_h, d_n, d_r, d_a
        secs = []
        for _ in range(2):
            t0 = time.perf_counter()
            built = CollisionFreeKmerIndex.from_flat_kmers(flat, modulo=modulo, skip_frequencies=True)
            secs.append(time.perf_counter() - t0)
        assert np.array_equal(built._hashes_to_index, h2i.cpu().numpy()) and np.array_equal(built._kmers, s_kmers.cpu().numpy().view(np.uint64))
        bytes_in, bytes_out = 24 * n, 26 * n + 8 * modulo
        result["index_build_e2e"] = {"entries_per_s": n / min(secs), "ms": 1e3 * min(secs), "first_call_ms": 1e3 * secs[0], "h2d_bytes": bytes_in, "d2h_bytes": bytes_out,
                                     "note": "CollisionFreeKmerIndex.from_flat_kmers(FlatKmers of host numpy arrays: k-mers, nodes, ref offsets, allele frequencies; "
                                             "skip_frequencies) -> the reference's eight host arrays, pageable memory both ways; best of two calls"}
        del flat, built

    # ---------------- CPU baseline + parity of this very configuration (oracle port on the host cores, bounded sample) ----------------
    if not args.no_cpu_baseline and world == 1:
        from oracle import c_oracle
        c_oracle.set_num_threads(cpu_threads())
        idx = {"_hashes_to_index": h2i.cpu().numpy(), "_n_kmers": nkm.cpu().numpy().view(np.uint32),
               "_kmers": s_kmers.cpu().numpy().view(np.uint64), "_nodes": s_nodes.cpu().numpy().view(np.uint32), "_modulo": modulo}
        sample = reads[:min(R, 4_000_000)].cpu().numpy()
        seconds = args.cpu_seconds if full else min(args.cpu_seconds, 4.0)
        rate, kmers_done, secs, reads_done, ec = cpu_count_rate(idx, sample, k, seconds)
        result["cpu_baseline"] = {"value": rate, "unit": "kmers/s", "cores": c_oracle.num_threads(), "kind": "port",
                                  "sample": "first %d reads (%d k-mers) of the step's batch in %.1f s; oracle/gki_oracle.c, OpenMP, host has %d threads"
                                            % (reads_done, kmers_done, secs, cpu_threads())}
        # parity at the benchmarked configuration: the GPU's node counts of exactly those reads against the oracle's
        want = c_oracle.node_counts_from_entry_counts(idx, ec, n_nodes, parallel=True, size=int(counts.shape[0]))
        index.reset_counts()
        index.count_reads(reads[:reads_done], k, True)
        index.node_counts(n_nodes, out=counts)
        torch.cuda.synchronize()
        got = counts.cpu().numpy()
        result["parity_checked"] = {"reads": int(reads_done), "kmers": int(kmers_done), "node_count_sum": float(want.sum()),
                                    "equal": bool(np.array_equal(got, want)),
                                    "what": "get_node_counts of the first `reads` reads of the step's batch: GPU (gki_count_reads + gki_node_counts) == "
                                            "oracle/gki_oracle.c on the index this run built, np.array_equal over all %d nodes" % int(counts.shape[0])}
        del idx, sample, ec, want, got
    index.close()
    return result


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    name = config_name(args, max(world, args.gpus))      # the reference arm runs on rank 0 alone: --gpus names the configuration there too
    cfg = workload(args, name)
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return

    import torch
    import torch.distributed as dist
    from graph_kmer_index_b200 import distributed

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        distributed.init_process_group("nccl")
    res = measure(args, cfg, name, rank, world, local_rank, True)
    extra = None
    if world == 1 and name == "c2" and not args.no_c3 and not args.entries and not args.reads:
        # the human-scale configuration (BASELINE configs[2]) on this one GPU, so that the N = 2, 4, 8 lines (which run it) have their N = 1 point
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        cfg3 = workload(args, "c3")
        sub_args = argparse.Namespace(**vars(args))
        sub_args.steps, sub_args.warmup = min(args.steps, 10), min(args.warmup, 3)
        extra = {"config": bench_config(cfg3, args, world, name="c3"), "steps": sub_args.steps, "warmup": sub_args.warmup}
        try:
            r3 = measure(sub_args, cfg3, "c3", rank, world, local_rank, False)
            extra.update({key: r3.get(key) for key in ("value", "ms_per_step", "e2e", "stages_ms", "index_build", "roofline", "cpu_baseline", "parity_checked", "index")})
        except Exception as e:            # the attached record must never cost the main line
            extra["error"] = "%s: %s" % (type(e).__name__, e)
    if rank == 0:
        line = {"metric": "read_kmers_per_s_through_get_node_counts", "value": res["value"], "unit": "kmers/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": bench_config(cfg, args, world, name=name), "clocks": res["clocks"], "e2e": res["e2e"], "gpu_launches": res["gpu_launches"],
                "roofline": res.get("roofline"), "cpu_baseline": res.get("cpu_baseline"), "parity_checked": res.get("parity_checked"),
                "stages": res.get("stages"), "stages_ms": res["stages_ms"],
                "index_build": res["index_build"], "index_build_e2e": res.get("index_build_e2e"), "index_build_partitioned": res["index_build_partitioned"],
                "index": res["index"]}
        if extra is not None:
            line["c3"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- read k-mers/s through get_node_counts (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path over one batch: reset counters -> fused hash+probe+count of the rank's read
batch (gki_count_reads) -> get_node_counts (gki_node_counts) [-> NCCL all-reduce of the node-count vector at N>1].
Workload at N=1: BASELINE configs[1] (chr20-scale: 60M k=31 index entries, modulo 452930477, 10M x 150 bp reads,
both strands -> 2.4 G read k-mers per step).  At N>1 every rank holds the replicated index and its own 10M-read
shard (weak scaling) -- configs[2]'s sharding at the per-GPU batch of configs[1].
Prints ONE JSON line (rank 0).
"""
import argparse
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
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--entries", type=int)
    ap.add_argument("--reads", type=int)
    ap.add_argument("--p-hit", type=int, dest="p_hit")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU-baseline time box")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload(args):
    cfg = dict(CONFIGS[args.config])
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
    return done * nk / dt, done * nk, dt, done


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
    n, k, L = cfg["entries"], cfg["k"], cfg["read_len"]
    codes = synthetic.genome_codes(synthetic.genome_length(n, k))
    hashes, nodes, ref, af = synthetic.flat_kmers(n, cfg["nodes"], k, codes=codes)
    idx = c_oracle.build_index(hashes, nodes, ref, af, cfg["modulo"], skip_frequencies=True)
    del hashes, ref, af
    sample_reads = min(cfg["reads"], 2_000_000)   # a step of ~3.5 s on 16 threads: the per-step costs over the whole index (counter reset,
                                                    # node counts over all entries) weigh as in a full step
    reads = synthetic.reads(sample_reads, L, n, k, cfg["p_hit"], codes=codes)
    prepared = c_oracle._index_args(idx)
    nk = (L - k + 1) * 2
    ec = np.zeros(n, dtype=np.uint32)
    times = []
    for step in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        ec[:] = 0
        c_oracle.count_reads(idx, reads, k, True, ec, prepared)
        c_oracle.node_counts_from_entry_counts(idx, ec, cfg["nodes"])
        if step >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = sample_reads * nk * len(times) / total
    threads = c_oracle.num_threads()
    line = {"impl": "reference", "metric": "read_kmers_per_s_through_get_node_counts", "value": value, "unit": "kmers/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": bench_config(cfg, args, 1, sample_reads=sample_reads),
            "cpu_baseline": {"value": value, "unit": "kmers/s", "cores": threads, "kind": "port",
                             "sample": "%d of %d reads per step, oracle/gki_oracle.c with OpenMP over %d host threads" % (sample_reads, cfg["reads"], threads)},
            "e2e": {"value": value, "unit": "kmers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if not args.no_cpu_baseline:
        # reported next to the port: the port (all host threads, C) is the faster, i.e. the conservative, baseline and stays `value`
        line["reference_native"] = reference_native_rate(idx, reads, k, cfg["nodes"])
    print(json.dumps(line), flush=True)


def bench_config(cfg, args, world, sample_reads=None):
    c = {"workload": "%s: synthetic variant-graph index, %d k=%d entries (%d distinct k-mers x 2 nodes), modulo %d, %d nodes; "
                     "%d x %d bp reads per GPU, both strands, %.0f%% of reads drawn from the indexed sequence"
                     % (args.config, cfg["entries"], cfg["k"], (cfg["entries"] + 1) // 2, cfg["modulo"], cfg["nodes"], cfg["reads"],
                        cfg["read_len"], cfg["p_hit"] / 10.0),
         "index_entries": cfg["entries"], "modulo": cfg["modulo"], "n_nodes": cfg["nodes"], "reads_per_gpu": cfg["reads"],
         "read_len": cfg["read_len"], "k": cfg["k"], "kmers_per_step_per_gpu": cfg["reads"] * (cfg["read_len"] - cfg["k"] + 1) * 2,
         "parallelism": "reads sharded x%d, index replicated, 1 all-reduce of node counts" % world if world > 1 else "single GPU",
         "l2": "inputs larger than L2 (reads %.2f GB, bucket cells %.2f GB per step)" % (cfg["reads"] * cfg["read_len"] / 1e9, cfg["modulo"] * 8 / 1e9)}
    if sample_reads is not None:
        c["reads_per_step_sampled"] = sample_reads
    return c


def main():
    args = parse_args()
    cfg = workload(args)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return

    import torch
    import torch.distributed as dist
    from graph_kmer_index_b200 import DeviceIndex, _lib, distributed, synthetic

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        distributed.init_process_group("nccl")
    dev = torch.device("cuda", local_rank)
    n, k, L, R, modulo, n_nodes = cfg["entries"], cfg["k"], cfg["read_len"], cfg["reads"], cfg["modulo"], cfg["nodes"]
    nk_per_read = (L - k + 1) * 2
    stream = torch.cuda.current_stream().cuda_stream

    # ---------------- setup (untimed): synthetic FlatKmers + reads on the device, index built by K2 ----------------
    glen = synthetic.genome_length(n, k)
    genome = torch.empty(glen, dtype=torch.uint8, device=dev)
    _lib.call("gki_synth_genome", _lib.ptr(genome), glen, stream)
    hashes = torch.empty(n, dtype=torch.int64, device=dev)
    nodes = torch.empty(n, dtype=torch.int32, device=dev)
    _lib.call("gki_synth_flat_kmers", _lib.ptr(genome), n, n_nodes, k, _lib.ptr(hashes), _lib.ptr(nodes), None, None, stream)
    h2i = torch.empty(modulo, dtype=torch.int32, device=dev)
    nkm = torch.empty(modulo, dtype=torch.int32, device=dev)
    s_kmers, s_nodes = torch.empty_like(hashes), torch.empty_like(nodes)
    build_ms = None
    for _ in range(2):                 # the second run is the measured one (first touches the stream-ordered pool)
        build_ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        build_ev[0].record()
        _lib.call("gki_index_build", _lib.ptr(hashes), _lib.ptr(nodes), None, None, n, modulo, _lib.GKI_BUILD_SKIP_FREQUENCIES,
                  _lib.ptr(h2i), _lib.ptr(nkm), _lib.ptr(s_kmers), _lib.ptr(s_nodes), None, None, None, None, stream)
        build_ev[1].record()
        torch.cuda.synchronize()
        build_ms = build_ev[0].elapsed_time(build_ev[1])
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
    index.prepare_counting(k)          # Bloom filter + count table (otherwise built inside the first counting call)
    info = index.info()
    del hashes, nodes
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
    kern_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev0.record()
    for i in range(args.steps):
        index.reset_counts()
        kern_ev[i][0].record()
        index.count_reads(reads, k, True)
        kern_ev[i][1].record()
        index.node_counts(n_nodes, out=counts)
        if world > 1:
            distributed.allreduce_node_counts(counts)
    ev1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop(wall0, time.time()) if rank == 0 else None
    elapsed_ms = ev0.elapsed_time(ev1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kern_ev]))
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

    # ---------------- e2e: host (pinned) reads through the same C-ABI call, node counts read back to the host -------
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

        e2e_steps = max(1, min(args.steps, 5))
        for _ in range(5):                                       # the host pipeline settles its copy-lane choice in its first five calls
            e2e_step()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(e2e_steps):
            e2e_step()
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": total_kmers_per_step * e2e_steps / (float(t.item()) / 1e3), "unit": "kmers/s",
               "h2d_bytes_per_step": int(R * L), "d2h_bytes_per_step": int(counts.shape[0] * 8), "steps": e2e_steps,
               "host_memory": "pinned", "pack_lanes": int(os.environ.get("GKI_PACK_THREADS", max((os.cpu_count() or 2) - 2, 0) if world == 1 else
                                                 max((os.cpu_count() or 2) // int(os.environ.get("LOCAL_WORLD_SIZE", world)) - 1, 0))),
               "note": "per GPU bytes of the caller's ASCII reads (1 byte per base); inside the call pack_lanes host threads re-encode "
                       "chunks to 2 bits per base before the bus while the copy engine moves the other chunks as ASCII (csrc/count.cu)"}
        assert float(host_counts.sum().item()) == (total_counts if world == 1 else float(counts.sum().item()))
        del host_reads

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (count_reads_kernel) ----------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    kmers_per_launch = R * nk_per_read
    # counts summed over nodes counts every hit once per entry of the k-mer (2 entries per distinct k-mer)
    h = (entry_hits / 2.0) / kmers_per_launch
    occ = info["nonempty_buckets"] / modulo
    # Algorithmic HBM bytes of one launch.  The kernel probes the count table (csrc/count.cu), not
    # the reference's modulo buckets, so SURVEY 8(d)'s sector figure is re-derived for that layout, at the granularity
    # this B200 fills L2 from HBM with: a whole 128-byte line per random access (measured: 127 B of DRAM reads per 8-byte
    # gather, profiles/r1/calibrate_gather_v2_ncu.txt).  Per read position (= 2 k-mers, one canonical key): its share of
    # the ASCII read; for a hit one table line (keys + counters) and the 32-byte counter sector written back; the Bloom
    # filter word comes from L2 when the filter is resident there, else it is one more line.
    LINE = 128.0
    positions = kmers_per_launch / 2.0
    h_pos = 2.0 * h                                         # hit positions / positions (a position hits on one strand)
    filter_resident = info["has_filter"] and (n // 2) <= (32 << 20)
    # a filter larger than L2 is addressed by minimizer (csrc/count.cu, filter_m): consecutive windows share a line as long as
    # their minimizer does, on average (w + 1) / 2 = 9 windows for the 17 m-mers of a k-mer
    minimizer_filter = (not filter_resident) and k in (27, 29, 31) and os.environ.get("GKI_FILTER_MZ", "1") != "0"
    filter_lines = 0.0 if (filter_resident or not info["has_filter"]) else (2.0 / 18.0 if minimizer_filter else 1.0)
    bytes_per_position = 2.0 * L / nk_per_read + LINE * filter_lines + (LINE + 32.0) * h_pos
    bytes_per_kmer = bytes_per_position / 2.0
    achieved = kmers_per_launch * bytes_per_kmer / (kernel_ms / 1e3) / 1e9
    # the same for the reference's own bucket layout (SURVEY 8(d) as written: 1 sector empty bucket / 3 miss / 5 hit)
    survey_sectors = (1 - h) * (1 - occ) * 1 + (1 - h) * occ * 3 + h * 5
    survey_bytes_per_kmer = L / nk_per_read + 32.0 * survey_sectors
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("count_reads_kernel_dram_bytes_per_launch_%s" % args.config)
        except Exception:
            traffic = None
    # the other two resources the kernel leans on, against ceilings measured on this pool's B200 (profiles/r1/calibrate_gather_v2.jsonl):
    # per-thread global requests through L1TEX (1 sector per cycle per SM) and random HBM line fetches
    requests = positions * (1.0 + 2.0 * h_pos + 0.05)      # filter word + (keys load + RED) per hit + ~5 % false positives
    hbm_lines = positions * (h_pos + filter_lines + 0.05)
    roofline = {"bound": "hbm", "kernel": "count_reads_kernel<both=true,paired=true> (Bloom filter: %s)" % ("none" if not info["has_filter"] else ("L2-resident" if filter_resident else ("HBM, minimizer-addressed" if minimizer_filter else "HBM"))),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
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
                "note": "HBM bytes of the count-table layout at the measured 128-byte fill granularity: ASCII + per hit one table line "
                        "and the counter sector written back (+ one line per position when the filter does not fit L2); "
                        "traffic = ncu dram bytes of one launch (false-positive and chain probes are the excess)"}

    # ---------------- index build end to end: the reference-facing call on host numpy arrays (cfki:422-467) ----------------
    build_e2e = None
    if world == 1 and not args.no_e2e:
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
        build_e2e = {"entries_per_s": n / min(secs), "ms": 1e3 * min(secs), "first_call_ms": 1e3 * secs[0], "h2d_bytes": bytes_in, "d2h_bytes": bytes_out,
                     "note": "CollisionFreeKmerIndex.from_flat_kmers(FlatKmers of host numpy arrays: k-mers, nodes, ref offsets, allele frequencies; "
                             "skip_frequencies) -> the reference's eight host arrays, pageable memory both ways; best of two calls"}
        del flat, built

    # ---------------- CPU baseline (oracle port on the host cores, bounded sample) ----------------
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import c_oracle
        idx = {"_hashes_to_index": h2i.cpu().numpy(), "_n_kmers": nkm.cpu().numpy().view(np.uint32),
               "_kmers": s_kmers.cpu().numpy().view(np.uint64), "_nodes": s_nodes.cpu().numpy().view(np.uint32), "_modulo": modulo}
        sample = reads[:min(R, 4_000_000)].cpu().numpy()
        rate, kmers_done, secs, reads_done = cpu_count_rate(idx, sample, k, args.cpu_seconds)
        cpu = {"value": rate, "unit": "kmers/s", "cores": c_oracle.num_threads(), "kind": "port",
               "sample": "first %d reads (%d k-mers) of the step's batch in %.1f s; oracle/gki_oracle.c, OpenMP, host has %d threads"
                         % (reads_done, kmers_done, secs, cpu_threads())}

    line = {"metric": "read_kmers_per_s_through_get_node_counts", "value": value, "unit": "kmers/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": bench_config(cfg, args, world), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu,
            "index_build": {"entries_per_s": n / (build_ms / 1e3), "ms": build_ms, "entries": n,
                            "compulsory_gbs": (50.0 * n + 8.0 * modulo) / (build_ms / 1e3) / 1e9,
                            "note": "gki_index_build (skip_frequencies, kmers+nodes), device-resident, second of two runs; binned path (one scatter + per-bin ordering, csrc/build.cu)"},
            "index_build_e2e": build_e2e, "index_build_partitioned": part_build,
            "index": {"device_bytes": info["device_bytes"], "has_filter": info["has_filter"], "nonempty_buckets": info["nonempty_buckets"]}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

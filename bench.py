#!/usr/bin/env python
"""Benchmark of the candidate-extension hot path (BASELINE.json configs[1]).

One "step" = one pass of the hot path over one batch of synthetic candidates: ungapped scoring (K1) of every candidate
followed by banded Smith-Waterman + traceback + re-scoring (K2+K4) of every candidate, 150 bp reads.
  value          banded-SW GCUPS = candidates * 16 * L cell updates / step time, inputs resident in HBM
  e2e            the same through the C ABI with pinned HOST buffers (H2D of the candidates, D2H of all results inside
                 the timed region)
  roofline       dominant kernel (swForwardKernel, timed together with the swTraceScoreKernel launches that overlap it =
                 the isaac_ext_gapped_batch_device call): algorithmic integer operations (23 per cell) per second
                 against the integer-pipe peak measured live on this GPU (MEASURED_PEAKS.json has no INT32 figure)
  cpu_baseline   the reference's own code (oracle/_ref, kind "reference") or the scalar restatement (kind "port")
                 timed on this box's host cores on a bounded sample of the same workload
`--impl reference` times only that CPU arm.  N > 1: one rank per GPU (torchrun), candidates sharded with no data-path
collective (weak scaling, a fixed batch per GPU); rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

# Only the JSON line may reach stdout: NCCL and other libraries print there (e.g. "NCCL version ..."), so file
# descriptor 1 points at stderr for the whole run and emit() writes the line to the saved descriptor.
_STDOUT_FD = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_STDOUT_FD, (line + "\n").encode())


OPS_PER_CELL = 23          # F 7 + G 8 + E 8 integer operations per (G,E,F) cell
BAND = 16


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--candidates", type=int, default=10_000_000, help="(read, window) pairs per GPU and step")
    p.add_argument("--read-length", type=int, default=150)
    p.add_argument("--genome-bases", type=int, default=5_000_000)
    p.add_argument("--per-read", type=int, default=8, help="candidates generated per read")
    p.add_argument("--cigar-stride", type=int, default=32)
    p.add_argument("--cpu-sample-per-core", type=int, default=40_000)
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--band", type=int, default=32, help="--workload wide: lanes of the band (16 = the reference's, 32 = the widened band)")
    p.add_argument("--alignments", type=int, default=2_000_000, help="--workload wide: alignments per GPU and step")
    p.add_argument("--workload", default="micro", choices=["micro", "pairs", "pack", "wide", "realign"],
                   help="micro: BASELINE configs[1] (default, the bench line); pairs: build + rescue pipeline, read pairs/s; "
                        "pack: the io::FragmentHeader bin records of a tile's templates (isaac_ext_pack_fragments), fragments/s")
    p.add_argument("--bins", type=int, default=16, help="--workload realign: bins per step")
    p.add_argument("--bin-pairs", type=int, default=40_000, help="--workload realign: read pairs sampled per bin")
    p.add_argument("--variant-spacing", type=int, default=300, help="--workload realign: a SNP or indel of the haplotype every ~this many bases")
    p.add_argument("--compact", action="store_true", help="--workload pack: records cut to their total length instead of FragmentBuffer slots")
    p.add_argument("--pairs", type=int, default=None,
                   help="read pairs per GPU of the pairs pipeline (default 2M: BASELINE configs[2] sharded, next to the micro run and for --workload pairs; 0 disables)")
    p.add_argument("--pairs-genome-bases", type=int, default=3_100_000_000, help="genome of the pairs pipeline (SURVEY 8(d) G3100)")
    p.add_argument("--pairs-contigs", type=int, default=24)
    p.add_argument("--pairs-n-fraction", type=float, default=0.001)
    p.add_argument("--contigs", type=int, default=1, help="contigs of the synthetic genome of the micro workload")
    p.add_argument("--n-fraction", type=float, default=0.0, help="fraction of the micro workload's genome replaced by runs of N")
    p.add_argument("--indel-rate", type=float, default=5e-4, help="indel events per base of the simulated pairs (config 4: 1e-2)")
    return p.parse_args()


def make_workload(args, rank, n_candidates):
    from isaac_aligner_b200 import synth
    from isaac_aligner_b200.types import ReadSet
    L = args.read_length
    genome = synth.make_genome(args.genome_bases, n_contigs=args.contigs, seed=synth.SEED_G5, n_fraction=args.n_fraction)
    n_pairs = max(1, -(-n_candidates // (2 * args.per_read)))
    sim = synth.simulate_pairs(genome, n_pairs, L=L, seed=synth.SEED_READS + 1 + 1000 * rank)
    reads = ReadSet(sim.bcl, (L, L))
    cand = synth.microbench_candidates(sim, genome, per_read=args.per_read, seed=synth.SEED_READS + 2 + 1000 * rank)
    return genome, reads, cand[:n_candidates]


def bind_to_gpu_numa(device):
    """Pins this rank to the host cores next to its GPU (NVML's ideal CPU affinity), so that the page-locked buffers it allocates
    from here on and the threads that fill them live on the GPU's NUMA node.  Returns what it did, for the bench line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = sorted(64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1)
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return {"bound": False, "why": "no ideal core in this process's cpuset"}
        os.sched_setaffinity(0, allowed)
        return {"bound": True, "cores": "%d-%d (%d)" % (allowed[0], allowed[-1], len(allowed))}
    except Exception as e:      # noqa: BLE001
        return {"bound": False, "why": "%s: %s" % (type(e).__name__, e)}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed regions run (B200_PROFILING.md)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.lines = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(device), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.lines:
            if t < t0 or t > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_arm(args, genome, reads, cand, config, steps, warmup):
    """Times the CPU implementation (all host threads) on `cand`; returns (GCUPS, seconds per step, kind, cores)."""
    import oracle_lib
    chk = oracle_lib.Oracle(oracle_lib.REF_SO) if os.path.exists(oracle_lib.REF_SO) else oracle_lib.port()
    cores = os.cpu_count() or 1
    g = oracle_lib.GenomeHolder(genome)
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        chk.gapped(g, reads, config, cand, cigar_stride=args.cigar_stride, threads=cores)   # alignUngapped + alignGapped
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    sec = float(np.mean(times))
    cells = len(cand) * BAND * args.read_length
    return cells / sec / 1e9, sec, chk.kind, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from isaac_aligner_b200.types import Config
    cores = os.cpu_count() or 1
    n = min(args.candidates, args.cpu_sample_per_core * cores)
    genome, reads, cand = make_workload(args, 0, n)
    config = Config.default(max_read_length=2 * args.read_length)
    gcups, sec, kind, cores = cpu_arm(args, genome, reads, cand, config, args.steps, args.warmup)
    sample = "%d of %d candidates per step (same generator, ungapped + gapped per candidate), %d host threads" % (
        n, args.candidates, cores)
    emit(json.dumps({
        "impl": "reference", "metric": "banded_sw_gcups", "value": gcups, "unit": "GCUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args):
    return {"workload": "BASELINE configs[1]: candidate-fragment microbench, %d ungapped + banded-SW extensions per GPU "
                        "per step, %d bp reads, %d bp random genome, 60%% true / 20%% shifted / 20%% random loci, bwa scores"
                        % (args.candidates, args.read_length, args.genome_bases),
            "candidates_per_gpu": args.candidates, "read_length": args.read_length, "band": BAND,
            "l2": "inputs+outputs per step (%.1f GB) exceed the 126 MB L2, no flush needed"
                  % (args.candidates * (16 + 2 * 64 + (3 + args.cigar_stride) * 4) / 1e9)}


def make_pairs_workload(args, rank, world, n_pairs):
    """BASELINE configs[2] sharded over the ranks: the human-scale genome G3100 (3.1 Gbp in 24 contigs, 0.1 % N runs; every rank
    holds the whole of it, like every GPU does), this rank's simulated FR pairs, seed matches from the error-free auto seeds +
    decoys, explicit template length statistics.  Runs fork workers: call it before the process touches CUDA."""
    from isaac_aligner_b200 import synth
    L = args.read_length
    workers = max(1, (os.cpu_count() or 1) // max(1, world))
    genome = synth.make_genome_parallel(args.pairs_genome_bases, n_contigs=args.pairs_contigs, seed=synth.SEED_G3100,
                                        n_fraction=args.pairs_n_fraction, workers=workers)
    sim = synth.simulate_pairs_parallel(genome, n_pairs, seed=synth.SEED_READS + 7 + 1000 * rank, workers=workers, L=L,
                                        indel_rate=args.indel_rate, seed_offsets=synth.auto_seed_offsets(L))
    matches, begin = synth.make_matches(sim, genome, seed=synth.SEED_READS + 8 + 1000 * rank, decoy_rate=0.2)
    return genome, sim.bcl, matches, begin, synth.seed_table(sim)


def pinned_like(a):
    """a copy of numpy array `a` in page-locked host memory (what a caller that cares about transfer speed hands to the ABI)"""
    import torch
    t = torch.empty(a.nbytes, dtype=torch.uint8).pin_memory()
    v = t.numpy().view(a.dtype).reshape(a.shape)
    v[...] = a
    return v, t


def pairs_pipeline_gpu(args, workload, local_rank, world, steps, warmup):
    """aligned read pairs/s: per step ONE tile = isaac_ext_set_reads (BCL bytes up, decode) + isaac_ext_build_templates (seed matches
    up, the whole TemplateBuilder on the device, templates down), host-pointer ABI with page-locked buffers = end to end.  Every rank
    runs its own tiles (no data-path collective); the time is the maximum over the ranks; the one exchange is the sum of the tile
    statistics.  Returns the dict of the bench line (rank 0) and what the CPU arm needs."""
    import torch
    from isaac_aligner_b200 import capi, distributed
    from isaac_aligner_b200.batch import MatchBatch, Tls
    from isaac_aligner_b200.types import Config, ReadSet
    genome, bcl, matches, begin, seeds = workload
    L = args.read_length
    n = bcl.shape[0]
    keep = []
    bcl_p, t = pinned_like(bcl); keep.append(t)
    bcl_q, t = pinned_like(bcl); keep.append(t)                  # tiles alternate between two host buffers (double buffering)
    matches_p, t = pinned_like(matches); keep.append(t)
    begin_p, t = pinned_like(begin); keep.append(t)
    matches_q, t = pinned_like(matches); keep.append(t)
    begin_q, t = pinned_like(begin); keep.append(t)
    tiles = [ReadSet(bcl_p, (L, L)), ReadSet(bcl_q, (L, L))]
    batches = [MatchBatch(matches_p, begin_p, seeds, with_gaps=True), MatchBatch(matches_q, begin_q, seeds, with_gaps=True)]
    reads, mb = tiles[0], batches[0]
    tls = Tls.make()
    config = Config.default(max_read_length=2 * L, device=local_rank, host_threads=max(1, (os.cpu_count() or 1) // world))
    ctx = capi.Context(config)
    ctx.set_reference(genome)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def run_tiles(count):
        """`count` tiles one after the other, every one with its own upload of the BCL bytes: while tile k is processed the bytes of
        tile k + 1 go up and are decoded (isaac_ext_prefetch_reads), the way the reference loads the next tile meanwhile, and the
        templates of tile k - 1 come down (isaac_ext_build_templates_deferred / isaac_ext_fetch_templates)"""
        trace = os.environ.get("ISAAC_BENCH_TRACE")
        ctx.prefetch_reads(tiles[0])
        ctx.prefetch_batch(batches[0], n)
        ctx.set_reads(tiles[0])
        res, waiting = None, None
        for k in range(count):
            t = [time.perf_counter()]
            if k + 1 < count:
                ctx.prefetch_reads(tiles[(k + 1) & 1])
                ctx.prefetch_batch(batches[(k + 1) & 1], n)
            t.append(time.perf_counter())
            handle = ctx.build_templates_deferred(batches[k & 1], tls)
            if waiting is not None:
                res = ctx.fetch_templates(waiting, copy=False)       # tile k - 1: its download ran next to the kernels of tile k
            waiting = handle
            t.append(time.perf_counter())
            if k + 1 < count:
                ctx.set_reads(tiles[(k + 1) & 1])
            t.append(time.perf_counter())
            if trace:
                print("[bench] tile %d: prefetch calls %.2f ms, build_templates %.2f ms, set_reads %.2f ms"
                      % (k, (t[1] - t[0]) * 1e3, (t[2] - t[1]) * 1e3, (t[3] - t[2]) * 1e3), file=sys.stderr)
        return ctx.fetch_templates(waiting, copy=False)              # the last tile's download is inside the timed region too

    res = run_tiles(max(2, warmup))
    barrier()
    l0 = ctx.launches
    t0 = time.perf_counter()
    res = run_tiles(steps)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
    launches = (ctx.launches - l0) // max(1, steps)
    for _ in range(2):
        res = ctx.build_templates(mb, tls, copy=False)           # warm-up of the blocking call (its own page-locked result buffers)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        res = ctx.build_templates(mb, tls, copy=False)           # the reads stay resident: the TemplateBuilder call alone
    templates_ms = (time.perf_counter() - t0) * 1e3 / steps
    d2h = n * 16 + 2 * n * 64 + int(res.cigarWords) * 4
    times = torch.tensor([e2e_ms, templates_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        torch.distributed.all_reduce(times, op=torch.distributed.ReduceOp.MAX)
    e2e_ms, templates_ms = (float(x) for x in times.cpu())
    # MatchSelectorStats of the tile (TileBarcodeStats per read and pass filter), summed over the ranks: the path's one exchange
    templates = ctx._templates(res)
    t0 = time.perf_counter()
    cycle_stats = ctx.tile_cycle_stats()                         # matchSelector::TileStats: 4 x 47105 u64 (score histograms, per-cycle arrays)
    cycle_ms = (time.perf_counter() - t0) * 1e3
    tile_stats = ctx.template_stats(mb, tls, templates)
    payload = torch.from_numpy(np.concatenate([tile_stats.reshape(-1), cycle_stats.reshape(-1)]).view(np.int64).copy()).cuda()
    t0 = time.perf_counter()
    if world > 1:
        payload = distributed.allreduce_stats(payload)
        torch.cuda.synchronize()
    allreduce_ms = (time.perf_counter() - t0) * 1e3
    payload = payload.cpu().numpy().view(np.uint64)
    read1 = payload[:32]
    cycles = payload[tile_stats.size:].reshape(4, -1)
    line = {"workload": "BASELINE configs[2] sharded: %d simulated 2x%d bp FR pairs per GPU per step on the %d bp / %d contig / %.1f %% N "
                        "synthetic genome resident on every GPU, indel events %.0e/base, seed matches from error-free auto seeds + 20 %% "
                        "decoys, explicit TLS 245/350/455; one step = isaac_ext_set_reads + isaac_ext_build_templates of one tile"
                        % (n, L, sum(int(c.size) for c in genome), len(genome), 100 * args.pairs_n_fraction, args.indel_rate),
            "pairs_per_gpu": n, "n_gpus": world,
            "value": world * n / (templates_ms * 1e-3), "unit": "pairs/s", "ms_per_step": templates_ms,
            "value_is": "isaac_ext_build_templates alone (reads resident; matches up and templates down inside), max over ranks",
            "e2e": {"value": world * n / (e2e_ms * 1e-3), "unit": "pairs/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(bcl.nbytes + matches.nbytes + begin.nbytes), "d2h_bytes_per_step": int(d2h),
                    "api": "per tile isaac_ext_set_reads + isaac_ext_build_templates_deferred / isaac_ext_fetch_templates: the next tile's bytes go "
                           "up (isaac_ext_prefetch_reads / _batch) and the previous tile's templates come down while a tile is processed; "
                           "page-locked host buffers; %d tiles back to back" % steps},
            "gpu_launches_per_step": int(launches),
            "matches_per_gpu": int(len(matches)), "rescue_requests": int(templates.rescue_requests),
            "templates_built": int(templates.templates["built"].sum()), "proper_pairs": int(templates.templates["properPair"].sum()),
            "match_selector_stats_all_ranks": dict(zip(distributed.TEMPLATE_STAT_NAMES, (int(x) for x in read1[:16]))),
            "tile_stats_all_ranks": {"payload_u64": int(payload.size), "tile_cycle_stats_ms": cycle_ms, "allreduce_ms": allreduce_ms,
                                     "read1_cycleMismatches": int(cycles[0, 34816:35840].sum()), "read1_cycleBlanks": int(cycles[0, 32768:33792].sum()),
                                     "read1_uniquelyAlignedFragments": int(cycles[0, 47104]),
                                     "is": "TileBarcodeStats summary + the full matchSelector::TileStats of every (read, pass filter), summed over "
                                           "the ranks with one all-reduce: the path's only collective"}}
    ctx.close()
    return line, (genome, ReadSet(bcl, (L, L)), MatchBatch(matches, begin, seeds, with_gaps=True), tls, config)


def pairs_pipeline_cpu(genome, reads, mb, tls, config, sample_clusters):
    """the reference's own TemplateBuilder (oracle/_ref), one per host thread, on the first `sample_clusters` clusters"""
    import oracle_lib
    from isaac_aligner_b200.batch import MatchBatch, TemplateOptions
    from isaac_aligner_b200.types import ReadSet
    if not os.path.exists(oracle_lib.REF_SO):
        return {"unavailable": "oracle/_ref/libisaac_ref.so did not travel to this box; TemplateBuilder has no second checker"}
    chk = oracle_lib.Oracle(oracle_lib.REF_SO)
    cores = os.cpu_count() or 1
    k = min(sample_clusters, reads.cluster_count)
    sub_reads = ReadSet(reads.bcl[:k], reads.read_lengths)
    sub_mb = MatchBatch(mb.matches[:int(mb.begin[k])], mb.begin[:k + 1], mb.seeds, with_gaps=True)
    g = oracle_lib.GenomeHolder(genome)
    oracle_lib.build_templates(chk, g, sub_reads, config, sub_mb, tls, TemplateOptions.make(), threads=cores)   # warm the genome cache
    t0 = time.perf_counter()
    oracle_lib.build_templates(chk, g, sub_reads, config, sub_mb, tls, TemplateOptions.make(), threads=cores)
    sec = time.perf_counter() - t0
    return {"value": k / sec, "unit": "pairs/s", "cores": cores, "kind": chk.kind,
            "sample": "first %d pairs of rank 0 through the reference's TemplateBuilder, one pass, %d host threads, %.2f s" % (k, cores, sec)}


def run_b200(args):
    import torch
    from isaac_aligner_b200 import capi
    from isaac_aligner_b200.types import CANDIDATE_DTYPE, FRAGMENT_DTYPE, Config

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # the pairs workload is drawn by fork workers: before this process has a CUDA context
    n_pairs = 2_000_000 if args.pairs is None else args.pairs
    pairs_workload, pairs_error = None, None
    if n_pairs:
        try:
            pairs_workload = make_pairs_workload(args, rank, world, n_pairs)
        except Exception as e:      # noqa: BLE001
            pairs_error = "%s: %s" % (type(e).__name__, e)
    realign_workload = None
    if world == 1 and n_pairs:                                       # the side measurements go together: --pairs 0 gives the bare micro run
        try:
            realign_workload = make_realign_workload(args, rank)
        except Exception as e:      # noqa: BLE001
            realign_workload = "%s: %s" % (type(e).__name__, e)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the candidate-extension path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    L, n, stride = args.read_length, args.candidates, args.cigar_stride
    genome, reads, cand = make_workload(args, rank, n)
    all_cores = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa(local_rank)
    n = len(cand)
    # the ranks of one box share its host cores: each context gets its share of the host threads
    config = Config.default(max_read_length=2 * L, device=local_rank, host_threads=max(1, (os.cpu_count() or 1) // world))
    ctx = capi.Context(config)
    ctx.set_reference(genome)
    ctx.set_reads(reads)

    dev = torch.device("cuda", local_rank)
    d_cand = torch.from_numpy(cand.view(np.uint8).reshape(n, 16)).to(dev)
    d_frag_u = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    d_cig_u = torch.empty((n, 3), dtype=torch.int32, device=dev)
    d_frag_g = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    d_cig_g = torch.empty((n, stride), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    d_stats = torch.zeros(64, dtype=torch.int64, device=dev)
    from isaac_aligner_b200 import distributed

    def step(events=None):
        if events:
            events[0].record()
        ctx.ungapped_device(n, d_cand.data_ptr(), d_frag_u.data_ptr(), d_cig_u.data_ptr(), 0, stream)
        if events:
            events[1].record()
        ctx.gapped_device(n, d_cand.data_ptr(), stride, d_frag_g.data_ptr(), d_cig_g.data_ptr(), 0, stream)
        if events:
            events[2].record()
        # per-tile statistics: K6 counters of this rank's results, summed over the ranks (the path's only collective)
        d_stats.zero_()
        ctx.tile_stats_device(n, d_frag_g.data_ptr(), d_stats.data_ptr(), stream)
        distributed.allreduce_stats(d_stats)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = ctx.launches
    t_wall0 = time.time()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for s in range(args.steps):
        step(evs[s])
    stop.record()
    barrier()
    gpu_launches = ctx.launches - launches0
    total_ms = start.elapsed_time(stop)
    ms_ungapped = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    ms_gapped = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    cells = float(n) * BAND * L

    # ---- end to end through the host-pointer ABI, pinned buffers, H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
        h_cand_t = pin((n, 16), torch.uint8)
        h_cand_t.numpy()[:] = cand.view(np.uint8).reshape(n, 16)
        h_cand = h_cand_t.numpy().reshape(-1).view(CANDIDATE_DTYPE)
        # isaac_ext_align_batch_packed: what FragmentBuilder::alignFragments keeps of every candidate, 32 bytes each + the words of
        # the accepted gapped CIGARs, chunked with overlapped copies
        from isaac_aligner_b200.batch import expected_alignments
        from isaac_aligner_b200.types import ALIGNMENT_DTYPE
        out_a, out_p = pin((n, 32), torch.uint8), pin((n * 4,), torch.int32)
        view_a, view_p = out_a.numpy().reshape(-1).view(ALIGNMENT_DTYPE), out_p.numpy().view(np.uint32)
        words = [0]

        def e2e_step():
            words[0] = ctx.align_packed(h_cand, view_a, view_p)

        for _ in range(max(1, args.warmup // 2)):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item()) / args.steps
        # the host results of the e2e path must be what the device-resident results say (the acceptance rule applied in numpy)
        res_u = d_frag_u.cpu().numpy().reshape(-1).view(FRAGMENT_DTYPE)
        res_g = d_frag_g.cpu().numpy().reshape(-1).view(FRAGMENT_DTYPE)
        want_a, want_p = expected_alignments(res_u, d_cig_u.cpu().numpy().view(np.uint32), res_g, d_cig_g.cpu().numpy().view(np.uint32), stride, L)
        assert want_a.tobytes() == view_a.tobytes(), "resident and end-to-end results differ"
        assert words[0] == len(want_p) and np.array_equal(view_p[:words[0]], want_p), "resident and end-to-end CIGAR pools differ"
        del res_u, res_g, want_a, want_p
        e2e = {"value": world * cells / (e2e_ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": n * 16, "d2h_bytes_per_step": n * 32 + 4 * words[0],
               "api": "isaac_ext_align_batch_packed (ungapped + gapped alignment of every candidate, the 32-byte record FragmentBuilder::alignFragments "
                      "keeps of each + the accepted gapped CIGARs), page-locked host buffers; checked against the device-resident results"}
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    # ---- aligned read pairs/s (the other half of BASELINE.json's metric): configs[2] sharded over the ranks, every rank runs it.
    # Whatever goes wrong in it is reported in its place and must not cost the bench line, but every rank must still take part
    # in the collectives of the others: a failure on one rank is made a failure of all before any collective is entered.
    pairs_line, pairs_cpu_inputs = None, None
    if n_pairs:
        ok = torch.tensor([0 if pairs_workload is None else 1], dtype=torch.int32, device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()):
            del d_cand, d_frag_u, d_cig_u, d_frag_g, d_cig_g
            torch.cuda.empty_cache()
            try:
                pairs_line, pairs_cpu_inputs = pairs_pipeline_gpu(args, pairs_workload, local_rank, world, args.steps, args.warmup)
            except Exception as e:      # noqa: BLE001
                pairs_line = {"error": "%s: %s" % (type(e).__name__, e)}
        else:
            pairs_line = {"error": pairs_error or "the pairs workload could not be generated on another rank"}

    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return
    tile_stats = dict(zip(distributed.STAT_NAMES, (int(x) for x in d_stats.cpu().numpy().view(np.uint64)[:8])))
    # ---- roofline of the dominant kernel, integer-pipe peak measured live
    peak_add = ctx.measure_int32_peak(0)
    peak_max = ctx.measure_int32_peak(1)
    peak_dpx = ctx.measure_int32_peak(2)
    sw_gcups_kernel = cells / (ms_gapped * 1e-3) / 1e9
    achieved = sw_gcups_kernel * 1e9 * OPS_PER_CELL
    hbm_peak = 6545.6
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        hbm_src = "MEASURED_PEAKS.json"
    except (OSError, KeyError, ValueError):
        hbm_src = "fallback"
    per_cand = -(-L // 4) + -(-L // 8) + 16 + 64 + 12       # window 2-bit + mask, descriptor, record, cigar
    per_read = -(-L // 4) + -(-L // 8) + L
    ungapped_gbs = (n * per_cand + reads.cluster_count * 2 * per_read) / (ms_ungapped * 1e-3) / 1e9
    traffic = None
    try:
        entry = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["gapped_pass"]     # one ncu --set full capture
        traffic = int(entry["dram_bytes_per_launch"] * float(n) / entry["candidates"]) if L == 150 else None
    except (OSError, KeyError, ValueError):
        pass
    traffic_ungapped = None
    try:
        entry = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["ungapped_pass"]
        traffic_ungapped = int(entry["dram_bytes_per_launch"] * float(n) / entry["candidates"]) if L == 150 else None
    except (OSError, KeyError, ValueError):
        pass
    # ---- CPU baseline on a bounded sample, same box, on ALL its host cores (the GPU arm's NUMA binding is lifted)
    os.sched_setaffinity(0, all_cores)
    cores = os.cpu_count() or 1
    ns = min(n, args.cpu_sample_per_core * cores)
    cpu_gcups, cpu_sec, kind, cores = cpu_arm(args, genome, reads, cand[:ns], config, 1, 0)
    realign_line = None
    if isinstance(realign_workload, str):
        realign_line = {"error": realign_workload}
    elif realign_workload is not None:
        try:
            realign_line = run_realign(args, workload=realign_workload, embedded=True)       # build::GapRealigner over 16 bins
        except Exception as e:      # noqa: BLE001
            realign_line = {"error": "%s: %s" % (type(e).__name__, e)}
    if pairs_line is not None and pairs_cpu_inputs is not None:
        try:
            pgenome, preads, pmb, ptls, pconfig = pairs_cpu_inputs
            pairs_line["cpu_baseline"] = pairs_pipeline_cpu(pgenome, preads, pmb, ptls, pconfig, 4000 * cores)
        except Exception as e:      # noqa: BLE001
            pairs_line["cpu_baseline"] = {"error": "%s: %s" % (type(e).__name__, e)}

    emit(json.dumps({
        "metric": "banded_sw_gcups", "value": world * cells / (ms_per_step * 1e-3) / 1e9, "unit": "GCUPS",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": workload_config(args),
        "e2e": e2e, "gpu_launches": int(gpu_launches), "clocks": clocks, "host": {"cores": os.cpu_count(), "numa_binding_rank0": numa},
        "tile_stats": tile_stats,
        "pairs_pipeline": pairs_line,
        "gap_realigner": realign_line,
        "roofline": {"bound": "int32", "kernel": "swForwardKernel (timed with the swTraceScoreKernel launches it overlaps: "
                                                  "the whole isaac_ext_gapped_batch_device call)",
                     "achieved": achieved / 1e12, "peak": peak_add / 1e12,
                     "unit": "TOP/s", "frac": achieved / peak_add,
                     # the same against the packed 16x2 rate (two cells per instruction is what the kernel's VIMNMX / VIADDMNMX do; the adds,
                     # flag bookkeeping and multiply-adds around them are 32-bit instructions, so 1.0 here is not reachable)
                     "frac_vs_16x2_peak": achieved / peak_dpx, "traffic": traffic,
                     "hbm_gbs": (traffic / (ms_gapped * 1e-3) / 1e9) if traffic else None,
                     "hbm_frac": (traffic / (ms_gapped * 1e-3) / 1e9 / hbm_peak) if traffic else None,
                     "plane_bytes_per_launch": int(n) * L * 12,
                     "ops_per_cell": OPS_PER_CELL, "gcups_kernel": sw_gcups_kernel, "ms_per_launch": ms_gapped,
                     "peak_source": "measured live by isaac_ext_measure_int32_peak: add.s32 %.1f, max.s32 %.1f, "
                                    "16x2 max %.1f TOP/s" % (peak_add / 1e12, peak_max / 1e12, peak_dpx / 1e12)},
        "roofline_ungapped": {"bound": "hbm", "kernel": "ungappedKernel", "achieved": ungapped_gbs, "peak": hbm_peak,
                              "unit": "GB/s", "frac": ungapped_gbs / hbm_peak, "traffic": traffic_ungapped, "ms_per_launch": ms_ungapped,
                              "limited_by": "the L1 data pipe and the issue slots, not HBM (ncu, profiles/r2_n_ncu_ungappedKernel.txt: LSU wavefronts 74 % of "
                                            "peak, issue 66 %, DRAM 1.4 GB per 10 M candidates = the algorithmic bytes): the reference's ordered FP64 sum is one "
                                            "8-byte shared-memory table lookup + one DADD per base (2 wavefronts each, 16 per 16 bases) and cannot be reassociated; "
                                            "230 instructions per 16 bases, 64 of them that chain",
                              "peak_source": hbm_src, "candidates_per_s": n / (ms_ungapped * 1e-3)},
        "cpu_baseline": {"value": cpu_gcups, "unit": "GCUPS", "cores": cores, "kind": kind,
                         "sample": "first %d of the %d candidates of rank 0, one pass, %d host threads, %.2f s"
                                   % (ns, n, cores, cpu_sec)},
    }))
    ctx.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def run_pairs(args):
    """--workload pairs: the pairs pipeline alone as the bench line (metric aligned_read_pairs_per_s)."""
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_pairs = 2_000_000 if args.pairs is None else args.pairs
    cores = os.cpu_count() or 1
    if args.impl == "reference":
        if rank != 0:
            return
        from isaac_aligner_b200.batch import MatchBatch, Tls
        from isaac_aligner_b200.types import Config, ReadSet
        k = min(n_pairs, 6000 * cores)
        genome, bcl, matches, begin, seeds = make_pairs_workload(args, 0, 1, k)
        L = args.read_length
        config = Config.default(max_read_length=2 * L)
        runs = [pairs_pipeline_cpu(genome, ReadSet(bcl, (L, L)), MatchBatch(matches, begin, seeds, with_gaps=True), Tls.make(), config, k)
                for _ in range(args.warmup + args.steps)][args.warmup:]
        if "unavailable" in runs[0]:
            emit(json.dumps({"impl": "reference", "unavailable": runs[0]["unavailable"]}))
            return
        v = float(np.mean([r["value"] for r in runs]))
        emit(json.dumps({"impl": "reference", "metric": "aligned_read_pairs_per_s", "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": k / v * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16+f64", "data": "synthetic",
                          "config": {"workload": "BASELINE configs[2] sharded, %d pairs per step of the same generator" % k},
                          "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": runs[0]["kind"], "sample": runs[0]["sample"]},
                          "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    workload = make_pairs_workload(args, rank, world, n_pairs)           # fork workers: before CUDA
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the candidate-extension path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    all_cores = os.sched_getaffinity(0)
    bind_to_gpu_numa(local_rank)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_wall0 = time.time()
    line, cpu_inputs = pairs_pipeline_gpu(args, workload, local_rank, world, args.steps, args.warmup)
    clocks = sampler.stop(t_wall0, time.time()) if sampler else None
    if world > 1:
        torch.distributed.destroy_process_group()
    if rank != 0:
        return
    pgenome, preads, pmb, ptls, pconfig = cpu_inputs
    os.sched_setaffinity(0, all_cores)
    cpu = pairs_pipeline_cpu(pgenome, preads, pmb, ptls, pconfig, 4000 * cores)
    emit(json.dumps({"metric": "aligned_read_pairs_per_s", "value": line["value"], "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": line["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "int16+f64", "data": "synthetic", "config": {"workload": line["workload"], "pairs_per_gpu": n_pairs,
                                                                              "l2": "per-step inputs + outputs exceed the 126 MB L2"},
                      "e2e": line["e2e"], "gpu_launches": int(line["gpu_launches_per_step"] * args.steps), "clocks": clocks, "pipeline": line,
                      "cpu_baseline": cpu}))


def small_pairs_workload(args, n_pairs):
    """a tile on the small genome for the kernel measurements that only need templates (--workload pack)"""
    from isaac_aligner_b200 import synth
    from isaac_aligner_b200.batch import MatchBatch, Tls
    from isaac_aligner_b200.types import ReadSet
    L = args.read_length
    genome = synth.make_genome(args.genome_bases, n_contigs=args.contigs, seed=synth.SEED_G5, n_fraction=args.n_fraction)
    sim = synth.simulate_pairs(genome, n_pairs, L=L, seed=synth.SEED_READS + 7, indel_rate=args.indel_rate, seed_offsets=synth.auto_seed_offsets(L))
    matches, begin = synth.make_matches(sim, genome, seed=synth.SEED_READS + 8, decoy_rate=0.2)
    return genome, ReadSet(sim.bcl, (L, L)), MatchBatch(matches, begin, synth.seed_table(sim), with_gaps=True), Tls.make()


def run_pack(args):
    """--workload pack: isaac_ext_pack_fragments over the templates of a simulated tile.  value = fragments/s of the
    kernel alone (CUDA events inside the library, records resident in HBM), e2e = the whole call with host templates in and host
    records out; roofline: HBM, algorithmic bytes = BCL + fragment / template records + CIGAR words read, record bytes written.
    One GPU (tiles shard over ranks like everywhere else; this workload is a kernel measurement)."""
    import torch
    from isaac_aligner_b200 import capi
    from isaac_aligner_b200.batch import PackOptions
    from isaac_aligner_b200.types import Config
    if args.impl == "reference":
        emit(json.dumps({"impl": "reference", "unavailable": "the pack workload has no reference arm: FragmentCollector needs the reference's bin storage (Boost.Filesystem)"}))
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the candidate-extension path has no CPU fallback")
    n_pairs = 500_000 if args.pairs is None else args.pairs
    genome, reads, mb, tls = small_pairs_workload(args, n_pairs)
    ctx = capi.Context(Config.default(max_read_length=2 * args.read_length))
    ctx.set_reference(genome)
    ctx.set_reads(reads)
    templates = ctx.build_templates(mb, tls)
    options = PackOptions(tile=1101, barcode_idx=0, keep_unaligned=True, compact=args.compact)
    sampler = ClockSampler(0)
    for _ in range(max(3, args.warmup)):
        res = ctx.pack_fragments(templates, options, copy=False)
    torch.cuda.synchronize()
    t_wall0 = time.time()
    kernel_ms, call_ms = [], []
    launches0 = ctx.launches
    for _ in range(args.steps):
        t0 = time.perf_counter()
        res = ctx.pack_fragments(templates, options, copy=False)
        call_ms.append((time.perf_counter() - t0) * 1e3)
        kernel_ms.append(float(res.kernelMs))
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1)
    fragments = int(res.storedFragments)
    L = args.read_length
    read_bytes = n_pairs * (2 * L + 2 * 64 + 16) + 4 * int(templates.fragments["cigarLength"].astype(np.int64).sum())
    write_bytes = int(res.recordBytes) + n_pairs * 2 * 9
    km, cm = float(np.mean(kernel_ms)), float(np.mean(call_ms))
    hbm_peak, hbm_src = 6545.6, "fallback"
    try:
        hbm_peak, hbm_src = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json"
    except (OSError, KeyError, ValueError):
        pass
    achieved = (read_bytes + write_bytes) / (km * 1e-3) / 1e9
    emit(json.dumps({"metric": "packed_fragments_per_s", "value": fragments / (km * 1e-3), "unit": "fragments/s", "n_gpus": 1,
                      "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": km, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                      "config": {"workload": "io::FragmentHeader bin records (%s) of the templates of %d simulated 2x%d bp pairs, --keep-unaligned"
                                             % ("compact" if args.compact else "FragmentBuffer slots", n_pairs, L),
                                 "l2": "records written per step (%d MB) exceed the 126 MB L2" % (int(res.recordBytes) >> 20)},
                      "e2e": {"value": fragments / (cm * 1e-3), "unit": "fragments/s", "ms_per_step": cm,
                              "h2d_bytes_per_step": int(n_pairs * (2 * 64 + 16) + 4 * templates.cigars.size),
                              "d2h_bytes_per_step": int(res.recordBytes) + n_pairs * 2 * 9},
                      "gpu_launches": int(ctx.launches - launches0), "clocks": clocks,
                      "roofline": {"bound": "hbm", "kernel": "packFragmentsKernel", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                                   "frac": achieved / hbm_peak, "traffic": None, "peak_source": hbm_src,
                                   "read_bytes_per_launch": int(read_bytes), "write_bytes_per_launch": int(write_bytes), "ms_per_launch": km}}))
    ctx.close()


def make_wide_workload(args, rank, n):
    """BASELINE configs[4]: 2x250 bp reads, widened band.  Every read of the simulated pairs against the window of the genome
    around its true locus: band / 2 bases in front, band / 2 - 1 behind (GappedAligner's flanks, GappedAligner.cpp:51-82, scaled to
    the band).  Returns ASCII queries [n, L] (strand order, 'n' for no-calls) and windows [n, L + band - 1]."""
    from isaac_aligner_b200 import synth
    L, W = args.read_length, args.band
    genome = synth.make_genome(args.genome_bases, n_contigs=1, seed=synth.SEED_G5)
    sim = synth.simulate_pairs(genome, -(-n // 2), L=L, seed=synth.SEED_READS + 5 + 1000 * rank, indel_rate=args.indel_rate)
    bcl = sim.bcl.reshape(-1, 2, L)
    fwd = bcl[:, 0, :]
    rev = bcl[:, 1, ::-1]                                            # read 2 in strand order: reversed ...
    acgt, comp = np.frombuffer(b"ACGT", dtype=np.uint8), np.frombuffer(b"TGCA", dtype=np.uint8)
    q = np.empty((bcl.shape[0], 2, L), dtype=np.uint8)
    q[:, 0, :] = np.where(fwd == 0, ord("n"), acgt[fwd & 3])
    q[:, 1, :] = np.where(rev == 0, ord("n"), comp[rev & 3])       # ... and complemented
    q = q.reshape(-1, L)[:n]
    pos = sim.position.reshape(-1)[:n] - W // 2
    idx = pos[:, None] + np.arange(L + W - 1)[None, :]
    db = genome[0][idx]
    return np.ascontiguousarray(q), np.ascontiguousarray(db)


def run_wide(args):
    """--workload wide: K3, the warp-wavefront banded Smith-Waterman with the widened band (isaac_ext_banded_sw_wide_batch).
    value = GCUPS with the strings resident in HBM (band * L cells per alignment), e2e = the host-pointer entry point."""
    import torch
    from isaac_aligner_b200 import capi
    from isaac_aligner_b200.types import Config
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.read_length == 150:
        args.read_length = 250                                       # the workload's own default
    L, W, n, stride = args.read_length, args.band, args.alignments, 32
    scores = (0, -3, 11, 4)
    cores = os.cpu_count() or 1
    workload = {"workload": "BASELINE configs[4]: %d banded Smith-Waterman alignments per GPU per step, %d bp reads of simulated pairs against "
                            "the %d-base windows around their loci, band of %d lanes (warp-wavefront kernel), bwa scores" % (n, L, L + W - 1, W),
                "alignments_per_gpu": n, "read_length": L, "band": W,
                "l2": "strings + results per step (%.1f GB) exceed the 126 MB L2" % (n * (2 * L + W + stride * 4 + 8) / 1e9)}
    if args.impl == "reference":
        if rank != 0:
            return
        import oracle_lib
        k = min(n, 2000 * cores)
        q, db = make_wide_workload(args, 0, k)
        qoff, doff = np.arange(k, dtype=np.uint64) * L, np.arange(k, dtype=np.uint64) * (L + W - 1)
        qlen = np.full(k, L, dtype=np.uint32)
        # the reference itself has 16 lanes only: its own code at band 16, the band-width-parametrised restatement otherwise
        chk = oracle_lib.Oracle(oracle_lib.REF_SO) if W == 16 and os.path.exists(oracle_lib.REF_SO) else oracle_lib.port()
        times = []
        for s_ in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            chk.banded_sw_flat(q.reshape(-1), qoff, qlen, db.reshape(-1), doff, scores, max_read_length=2 * L, cigar_stride=stride, threads=cores,
                               band=None if W == 16 else W)
            if s_ >= args.warmup:
                times.append(time.perf_counter() - t0)
        sec = float(np.mean(times))
        v = k * W * L / sec / 1e9
        emit(json.dumps({"impl": "reference", "metric": "banded_sw_gcups", "value": v, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "int16", "data": "synthetic", "config": workload,
                          "cpu_baseline": {"value": v, "unit": "GCUPS", "cores": cores, "kind": chk.kind,
                                           "sample": "%d alignments per step, %d host threads%s" % (k, cores, "" if W == 16 else
                                                     "; the reference hard-wires 16 lanes, a wider band runs its scalar restatement")},
                          "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the candidate-extension path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    q, db = make_wide_workload(args, rank, n)
    all_cores = os.sched_getaffinity(0)
    bind_to_gpu_numa(local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = capi.Context(Config.default(max_read_length=2 * L, device=local_rank))
    qoff, doff = np.arange(n, dtype=np.uint64) * L, np.arange(n, dtype=np.uint64) * (L + W - 1)
    qlen = np.full(n, L, dtype=np.uint32)
    d_q, d_db = torch.from_numpy(q).to(dev), torch.from_numpy(db).to(dev)
    d_qoff, d_doff = torch.from_numpy(qoff.view(np.int64)).to(dev), torch.from_numpy(doff.view(np.int64)).to(dev)
    d_qlen = torch.from_numpy(qlen.view(np.int32)).to(dev)
    d_cig = torch.empty((n, stride), dtype=torch.int32, device=dev)
    d_len, d_off = torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        ctx.banded_sw_wide_device(W, n, d_q.data_ptr(), d_qoff.data_ptr(), d_qlen.data_ptr(), d_db.data_ptr(), d_doff.data_ptr(), L, scores,
                                  stride, d_cig.data_ptr(), d_len.data_ptr(), d_off.data_ptr(), stream)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    t_wall0 = time.time()
    l0 = ctx.launches
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        step()
    stop.record()
    barrier()
    launches = ctx.launches - l0
    t = torch.tensor([start.elapsed_time(stop)], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t.item()) / args.steps
    cells = float(n) * W * L
    # ---- end to end: host strings in, CIGARs out
    hq, keep1 = pinned_like(q)
    hdb, keep2 = pinned_like(db)

    def e2e_step():
        return ctx._check(capi._lib.isaac_ext_banded_sw_wide_batch(
            ctx._h, W, n, hq.ctypes.data, qoff.ctypes.data, qlen.ctypes.data, hdb.ctypes.data, doff.ctypes.data, scores[0], scores[1], scores[2],
            scores[3], stride, h_cig.ctypes.data, h_len.ctypes.data, h_off.ctypes.data))

    h_cig, k3 = pinned_like(np.zeros((n, stride), dtype=np.uint32))
    h_len, k4 = pinned_like(np.zeros(n, dtype=np.uint32))
    h_off, k5 = pinned_like(np.zeros(n, dtype=np.uint32))
    import ctypes
    capi._lib.isaac_ext_banded_sw_wide_batch.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32] + [ctypes.c_void_p] * 5 + [ctypes.c_int] * 4 + \
        [ctypes.c_uint32] + [ctypes.c_void_p] * 3
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    dt = torch.tensor([(time.perf_counter() - t0) * 1e3 / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(dt, op=torch.distributed.ReduceOp.MAX)
    e2e_ms = float(dt.item())
    assert np.array_equal(h_len, d_len.cpu().numpy().view(np.uint32)) and np.array_equal(h_cig, d_cig.cpu().numpy().view(np.uint32)), "resident and end-to-end results differ"
    clocks = sampler.stop(t_wall0, time.time()) if sampler else None
    if world > 1:
        torch.distributed.destroy_process_group()
    if rank != 0:
        return
    peak_add = ctx.measure_int32_peak(0)
    # ---- CPU baseline + parity of the sample: the band-width-parametrised restatement (the reference has 16 lanes only)
    import oracle_lib
    os.sched_setaffinity(0, all_cores)
    k = min(n, 2000 * cores)
    chk = oracle_lib.port()
    t0 = time.perf_counter()
    cr, lr, orf = chk.banded_sw_flat(q[:k].reshape(-1), qoff[:k], qlen[:k], db[:k].reshape(-1), doff[:k], scores, max_read_length=2 * L,
                                     cigar_stride=stride, threads=cores, band=W)
    sec = time.perf_counter() - t0
    assert np.array_equal(lr, h_len[:k]) and np.array_equal(cr, h_cig[:k]) and np.array_equal(orf, h_off[:k]), "GPU and CPU model differ"
    achieved = cells / (ms * 1e-3) * OPS_PER_CELL
    emit(json.dumps({"metric": "banded_sw_gcups", "value": world * cells / (ms * 1e-3) / 1e9, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
                      "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "int16", "data": "synthetic", "config": workload,
                      "e2e": {"value": world * cells / (e2e_ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": e2e_ms,
                              "h2d_bytes_per_step": int(q.nbytes + db.nbytes + qoff.nbytes + doff.nbytes + qlen.nbytes),
                              "d2h_bytes_per_step": int(h_cig.nbytes + h_len.nbytes + h_off.nbytes), "api": "isaac_ext_banded_sw_wide_batch, page-locked host buffers"},
                      "gpu_launches": int(launches), "clocks": clocks,
                      "roofline": {"bound": "int32", "kernel": "bandedSwWideKernel<%d>" % W, "achieved": achieved / 1e12, "peak": peak_add / 1e12, "unit": "TOP/s",
                                   "frac": achieved / peak_add, "traffic": None, "ops_per_cell": OPS_PER_CELL, "ms_per_launch": ms,
                                   "peak_source": "measured live by isaac_ext_measure_int32_peak (add.s32)",
                                   "note": "a band lane per warp lane: the exchanges between band lanes are shuffles, which the thread-per-alignment kernel of the "
                                           "16-lane band does not pay; direction planes stay in shared memory (no HBM traffic for them)"},
                      "cpu_baseline": {"value": k * W * L / sec / 1e9, "unit": "GCUPS", "cores": cores, "kind": chk.kind,
                                       "sample": "first %d alignments of rank 0 through the band-width-parametrised restatement (the reference hard-wires 16 lanes), "
                                                 "%d host threads, %.2f s; results equal the GPU's" % (k, cores, sec)}}))
    ctx.close()


def _realign_bin_worker(job):
    from isaac_aligner_b200 import bins
    contig_bases, region, n_pairs, L, seed, spacing = job
    b = bins.simulate_bin([contig_bases], contig=0, region=region, n_pairs=n_pairs, read_length=L, seed=seed, template_mean=int(2.6 * L) + 60,
                          variant_spacing=spacing, max_indel=12)
    return b.data, b.record_offset, b.index, b.bin_start, b.bin_end


def make_realign_workload(args, rank):
    """the genome and the bins of one rank, drawn by fork workers (before the process has a CUDA context)"""
    import multiprocessing
    from isaac_aligner_b200 import synth
    L, B, n_pairs = args.read_length, args.bins, args.bin_pairs
    span = int(n_pairs * 2 * L / 20)                                 # ~20x
    genome = synth.make_genome((span + 4000) * B + 4000, n_contigs=1, seed=synth.SEED_G5)
    contig = genome[0]
    # bins are independent: every rank realigns its own --bins bins of the run (weak scaling, no exchange of any kind)
    jobs = [(contig, (2000 + k * (span + 4000), 2000 + k * (span + 4000) + span), n_pairs, L, 900 + k + 1000 * rank, args.variant_spacing) for k in range(B)]
    cores = sorted(os.sched_getaffinity(0))
    with multiprocessing.get_context("fork").Pool(min(len(cores), B)) as pool:
        made = pool.map(_realign_bin_worker, jobs)
    return genome, made


def run_realign(args, workload=None, embedded=False):
    """--workload realign: isaac_ext_realign_bin (build::GapRealigner) over --bins bins of one contig, each the records
    of --bin-pairs pairs sampled at ~20x from a haplotype with shared indels.  One step = every bin once (BinSorter::process per bin).
    value = index entries/s over the device phases alone (CUDA events inside the library: collectGaps + realignGaps), e2e = the same
    through the C call with host buffers (records up, updated records + positions + CIGARs down); cpu_baseline / --impl reference =
    the reference's own GapRealigner on the same bins, one bin per host thread.  Results are compared bin by bin.
    embedded: called from the default run on an initialised device (one GPU); returns the line instead of printing it."""
    import ctypes
    import torch
    from isaac_aligner_b200 import bins
    from isaac_aligner_b200.batch import Tls
    from isaac_aligner_b200.types import Config
    world = 1 if embedded else int(os.environ.get("WORLD_SIZE", "1"))
    rank = 0 if embedded else int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return                                                       # the reference arm runs on rank 0 alone
    L, B, n_pairs = args.read_length, args.bins, args.bin_pairs
    cores = sorted(os.sched_getaffinity(0))
    genome, made = workload if workload is not None else make_realign_workload(args, rank)
    the_bins = [bins.Bin(*m) for m in made]
    tls = [Tls.make(mn=int(2.0 * L), mx=int(3.4 * L) + 120, median=int(2.6 * L) + 60)]
    options = [bins.RealignOptions(b.bin_start, b.bin_end, tls, clip_semialigned=True) for b in the_bins]
    entries = sum(len(b.index) for b in the_bins)
    records = sum(len(b.record_offset) for b in the_bins)
    data_bytes = sum(int(b.data.size) for b in the_bins)
    workload = {"workload": "build::GapRealigner over %d bins of one contig: %d index entries / %d records / %d MB per step, 2x%d bp pairs at ~20x from a "
                            "haplotype with an indel or SNP every ~%d bases, --clip-semialigned, costs 3/4/0" % (B, entries, records, data_bytes >> 20, L, args.variant_spacing),
                "l2": "each bin's records are uploaded fresh and read once"}

    def reference_pass(threads):
        import oracle_lib
        ref = oracle_lib.require_reference() if False else oracle_lib.reference()
        if ref is None:
            return None
        holder = oracle_lib.GenomeHolder(genome)
        out = [None] * B
        import threading
        lock = threading.Lock()
        todo = list(range(B))

        def work():
            while True:
                with lock:
                    if not todo:
                        return
                    k = todo.pop()
                out[k] = oracle_lib.realign_bin(ref, holder, the_bins[k], options[k])
        oracle_lib.realign_bin(ref, holder, the_bins[0], options[0])          # the contig copy of the checker is made outside the timing
        t0 = time.perf_counter()
        pool = [threading.Thread(target=work) for _ in range(min(threads, B))]
        for t in pool:
            t.start()
        for t in pool:
            t.join()
        return out, time.perf_counter() - t0, min(threads, B)

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    if args.impl == "reference":
        done = reference_pass(len(cores))
        if done is None:
            emit(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libisaac_ref.so did not travel"}))
            return
        times = []
        for _ in range(max(1, args.steps)):
            _, sec, used = reference_pass(len(cores))
            times.append(sec)
        sec = float(np.mean(times))
        emit(json.dumps({"impl": "reference", "metric": "gap_realigner_fragments_per_s", "value": entries / sec, "unit": "fragments/s", "n_gpus": 1,
                          "steps": args.steps, "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "u8", "data": "synthetic", "config": workload,
                          "cpu_baseline": {"value": entries / sec, "unit": "fragments/s", "cores": used, "kind": "reference",
                                           "sample": "every bin of the step, one bin per host thread"},
                          "e2e": {"value": entries / sec, "unit": "fragments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the gap realigner has no CPU fallback")
    from isaac_aligner_b200 import capi
    if not embedded:
        torch.cuda.set_device(local_rank)
        if world > 1:
            torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        bind_to_gpu_numa(local_rank)
    cfg = Config.default(max_read_length=2 * L)
    cfg.device = local_rank
    ctx = capi.Context(cfg)
    ctx.set_reference(genome)
    lib = capi._lib
    # page-locked working copies: a step starts from the original records every time (the call updates them in place)
    work = [torch.empty(int(b.data.size), dtype=torch.uint8).pin_memory() for b in the_bins]
    offs = [torch.from_numpy(np.ascontiguousarray(b.record_offset, dtype=np.uint64).view(np.int64)).pin_memory() for b in the_bins]
    idx = [torch.from_numpy(np.ascontiguousarray(b.index).view(np.int64).reshape(-1)).pin_memory() for b in the_bins]
    results = [bins.RealignResultC() for _ in the_bins]
    # outputs of the batched call: caller's page-locked memory
    o_pos = [torch.empty(len(b.index), dtype=torch.int64).pin_memory() for b in the_bins]
    o_off = [torch.empty(len(b.index), dtype=torch.int32).pin_memory() for b in the_bins]
    o_len = [torch.empty(len(b.index), dtype=torch.int32).pin_memory() for b in the_bins]
    o_cig = [torch.empty(8 * len(b.index) + 4096, dtype=torch.int32).pin_memory() for b in the_bins]
    jobs = (bins.RealignJobC * B)()
    for k, b in enumerate(the_bins):
        j = jobs[k]
        j.options = ctypes.addressof(options[k].c)
        j.data, j.dataBytes = work[k].data_ptr(), int(b.data.size)
        j.recordOffset, j.recordCount = offs[k].data_ptr(), len(b.record_offset)
        j.index, j.indexCount = idx[k].data_ptr(), len(b.index)
        j.position, j.cigarOffset, j.cigarLength = o_pos[k].data_ptr(), o_off[k].data_ptr(), o_len[k].data_ptr()
        j.realignedCigars, j.realignedCigarCapacity = o_cig[k].data_ptr(), o_cig[k].numel()

    def step(collect=None):
        """every bin once through isaac_ext_realign_bin (one after the other: the device phases are timed by the library's events)"""
        for k, b in enumerate(the_bins):
            work[k].numpy()[:] = b.data
        device_ms = 0.0
        t0 = time.perf_counter()
        for k, b in enumerate(the_bins):
            ctx._check(lib.isaac_ext_realign_bin(ctx._h, ctypes.byref(options[k].c), ctypes.c_void_p(work[k].data_ptr()), ctypes.c_uint64(b.data.size),
                                                 ctypes.c_void_p(offs[k].data_ptr()), ctypes.c_uint64(len(b.record_offset)),
                                                 ctypes.c_void_p(idx[k].data_ptr()), ctypes.c_uint64(len(b.index)), ctypes.byref(results[k])))
            device_ms += float(results[k].collectMs) + float(results[k].realignMs)
            if collect is not None:
                n = len(b.index)
                collect.append((np.frombuffer((ctypes.c_char * (8 * n)).from_address(results[k].position), dtype=np.uint64).copy(),
                                work[k].numpy().copy(), int(results[k].realignedFragments)))
        return (time.perf_counter() - t0) * 1e3, device_ms

    def step_batched(collect=None):
        """every bin once through ONE isaac_ext_realign_bins call: the end-to-end number"""
        for k, b in enumerate(the_bins):
            work[k].numpy()[:] = b.data
        t0 = time.perf_counter()
        ctx._check(lib.isaac_ext_realign_bins(ctx._h, jobs, ctypes.c_uint32(B)))
        ms = (time.perf_counter() - t0) * 1e3
        if collect is not None:
            for k in range(B):
                collect.append((o_pos[k].numpy().view(np.uint64).copy(), work[k].numpy().copy(), int(jobs[k].realignedFragments)))
        return ms

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()

    def over_ranks(values, op):
        t = torch.tensor(values, dtype=torch.float64, device="cuda")
        if world > 1:
            torch.distributed.all_reduce(t, op=op)
        return [float(x) for x in t.tolist()]

    sampler = ClockSampler(local_rank) if rank == 0 and not embedded else None
    for _ in range(max(3, args.warmup)):
        step()
        step_batched()
    launches0 = ctx.launches
    t_wall0 = time.time()
    call_ms, dev_ms, single_ms = [], [], []
    for _ in range(args.steps):
        c, d = step()
        single_ms.append(c); dev_ms.append(d)
    launches = ctx.launches - launches0
    for _ in range(args.steps):
        barrier()                                                            # every rank starts its step together (they share the host)
        call_ms.append(step_batched())
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    got, got_batched = [], []
    step(got)
    step_batched(got_batched)
    for k in range(B):
        assert np.array_equal(got[k][0], got_batched[k][0]) and np.array_equal(got[k][1], got_batched[k][1]), "single and batched calls differ in bin %d" % k
    realigned = sum(g[2] for g in got)
    # the slowest rank counts; the fragments of all ranks are the job
    cm, dm = over_ranks([float(np.mean(call_ms)), float(np.mean(dev_ms))], torch.distributed.ReduceOp.MAX)
    entries_rank = entries
    entries, realigned_all = [int(x) for x in over_ranks([entries, realigned], torch.distributed.ReduceOp.SUM)]
    if world > 1:
        torch.distributed.destroy_process_group()
    if rank != 0:
        ctx.close()
        return
    line = {"metric": "gap_realigner_fragments_per_s", "value": entries / (dm * 1e-3), "unit": "fragments/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": dm, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": workload,
            "e2e": {"value": entries / (cm * 1e-3), "unit": "fragments/s", "ms_per_step": cm, "h2d_bytes_per_step": int(data_bytes + 8 * records + 16 * entries_rank),
                    "d2h_bytes_per_step": int(16 * entries_rank + 56 * 2 * realigned + 4 * sum(int(r.realignedCigarWords) for r in results) +
                                              16 * sum(int(r.gapCount) + int(r.deletionCount) for r in results)),
                    "api": "one isaac_ext_realign_bins call per step and GPU (three slots of the context take the bins in turn), page-locked host buffers; "
                           "of the records only the rewritten headers come back",
                    "one_bin_per_call_ms_per_step": float(np.mean(single_ms))},
            "gpu_launches": int(launches), "clocks": clocks, "realigned_fragments": int(realigned_all), "fragments_per_gpu": int(entries_rank)}
    hbm_peak, hbm_src = 6545.6, "fallback"
    try:
        hbm_peak, hbm_src = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json"
    except (OSError, KeyError, ValueError):
        pass
    achieved = (2 * data_bytes + 8 * records + 32 * entries) / (dm * 1e-3) / 1e9
    line["roofline"] = {"bound": "hbm", "kernel": "realignBinKernel (+ the collectGaps passes)", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                        "frac": achieved / hbm_peak, "traffic": None, "peak_source": hbm_src, "ms_per_launch": dm,
                        "note": "algorithmic bytes = the records read by collectGaps and by realign, offsets, index and results; the kernel is bound by the "
                                "divergent per-fragment search over gap combinations, not by HBM"}
    done = reference_pass(len(cores))
    if done is not None:
        want, sec, used = done
        for k in range(B):
            assert np.array_equal(got[k][0], want[k].position) and np.array_equal(got[k][1], want[k].data), "GPU and reference differ in bin %d" % k
        line["cpu_baseline"] = {"value": entries_rank / sec, "unit": "fragments/s", "cores": used, "kind": "reference",
                                "sample": "every bin of the step through the reference's own GapRealigner, one bin per host thread, %.2f s (the bins of rank 0); "
                                          "records and positions equal the GPU's" % sec}
    ctx.close()
    if embedded:
        return line
    emit(json.dumps(line))


if __name__ == "__main__":
    a = parse_args()
    if a.workload == "realign":
        run_realign(a)
    elif a.workload == "wide":
        run_wide(a)
    elif a.workload == "pack":
        run_pack(a)
    elif a.workload == "pairs":
        run_pairs(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)

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
    p.add_argument("--workload", default="micro", choices=["micro", "pairs", "pack"],
                   help="micro: BASELINE configs[1] (default, the bench line); pairs: build + rescue pipeline, read pairs/s; "
                        "pack: the io::FragmentHeader bin records of a tile's templates (isaac_ext_pack_fragments), fragments/s")
    p.add_argument("--compact", action="store_true", help="--workload pack: records cut to their total length instead of FragmentBuffer slots")
    p.add_argument("--pairs", type=int, default=None,
                   help="read pairs per GPU of the pairs pipeline (default 200k as a side measurement of the micro run, 1M for --workload pairs; 0 disables)")
    p.add_argument("--contigs", type=int, default=1, help="contigs of the synthetic genome (SURVEY 8(d) G3100: 24)")
    p.add_argument("--n-fraction", type=float, default=0.0, help="fraction of the genome replaced by runs of N (G3100: 0.001)")
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


def make_pairs_workload(args, rank, n_pairs):
    """BASELINE configs[0]/[2]-style input: simulated FR pairs, seed matches from the error-free auto seeds + decoys,
    explicit template length statistics."""
    from isaac_aligner_b200 import synth
    from isaac_aligner_b200.batch import MatchBatch, Tls
    from isaac_aligner_b200.types import ReadSet
    L = args.read_length
    genome = synth.make_genome(args.genome_bases, n_contigs=args.contigs, seed=synth.SEED_G5, n_fraction=args.n_fraction)
    sim = synth.simulate_pairs(genome, n_pairs, L=L, seed=synth.SEED_READS + 7 + 1000 * rank, indel_rate=args.indel_rate,
                               seed_offsets=synth.auto_seed_offsets(L))
    matches, begin = synth.make_matches(sim, genome, seed=synth.SEED_READS + 8 + 1000 * rank, decoy_rate=0.2)
    return genome, ReadSet(sim.bcl, (L, L)), MatchBatch(matches, begin, synth.seed_table(sim), with_gaps=True), Tls.make()


def pairs_pipeline_gpu(ctx, reads, mb, tls, steps, warmup, all_ranks=False):
    """FragmentBuilder::build for every cluster, then ShadowAligner::rescueShadow for every request of the stand-in
    template policy, both through the host-pointer ABI (H2D / D2H inside).  Returns a dict with pairs/s."""
    from isaac_aligner_b200 import synth
    from isaac_aligner_b200.batch import copy_result
    n = reads.cluster_count
    flat = ctx.build_fragments(mb)                                       # also the first warm-up pass
    req = synth.rescue_policy(flat.fragments, flat.begin, n)
    resc = ctx.rescue_shadows(tls, req)
    for _ in range(max(0, warmup - 1)):
        ctx.build_fragments(mb, copy=False); ctx.rescue_shadows(tls, req, copy=False)
    tb = tr = 0.0
    l0 = ctx.launches
    for _ in range(steps):
        t0 = time.perf_counter(); ctx.build_fragments(mb, copy=False)
        t1 = time.perf_counter(); ctx.rescue_shadows(tls, req, copy=False)
        t2 = time.perf_counter()
        tb += t1 - t0; tr += t2 - t1
    tb /= steps; tr /= steps
    # the whole TemplateBuilder: build + the rescues the templates really ask for + pair selection / MAPQ
    templates = ctx.build_templates(mb, tls)
    tt = 0.0
    for _ in range(steps):
        t0 = time.perf_counter(); ctx.build_templates(mb, tls, copy=False); tt += time.perf_counter() - t0
    tt /= steps
    gaps = flat.fragments["gapCount"] > 0
    # MatchSelectorStats of the tile (TileBarcodeStats per read and pass filter), summed over the ranks: the path's one exchange
    import torch
    from isaac_aligner_b200 import distributed
    t0 = time.perf_counter()
    tile_stats = ctx.template_stats(mb, tls, templates)
    stats_ms = (time.perf_counter() - t0) * 1e3
    summed = torch.from_numpy(tile_stats.view(np.int64).copy())
    # the all-reduce is a collective: only where every rank runs this function (the pairs workload; the side measurement of the
    # micro workload runs on rank 0 alone)
    if all_ranks and torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
        summed = distributed.allreduce_stats(summed.cuda()).cpu()
    read1 = summed.numpy().view(np.uint64)[0]
    return {"pairs": n, "template_pairs_per_s": n / tt, "templates_ms": tt * 1e3, "template_stats_ms": stats_ms,
            "match_selector_stats": dict(zip(distributed.TEMPLATE_STAT_NAMES, (int(x) for x in read1[:16]))),
            "template_rescue_requests": int(templates.rescue_requests),
            "templates_built": int(templates.templates["built"].sum()), "proper_pairs": int(templates.templates["properPair"].sum()),
            "pairs_per_s": n / (tb + tr), "build_ms": tb * 1e3, "rescue_ms": tr * 1e3,
            "matches": int(len(mb.matches)), "fragments": int(flat.fragments.size), "fragments_with_gaps": int(gaps.sum()),
            "rescue_requests": int(len(req)), "rescued": int(resc.flags.sum()), "shadow_candidates": int(resc.fragments.size),
            "gpu_launches_per_step": (ctx.launches - l0) // max(1, steps),
            "policy": "template_*: isaac_ext_build_templates = the reference's TemplateBuilder end to end; pairs_per_s / build_ms / "
                      "rescue_ms: the two batch calls alone with a stand-in policy (mates of all candidates are rescued unless both "
                      "reads have an edit-distance-0 candidate)"}, flat, req


def pairs_pipeline_cpu(genome, reads, mb, tls, config, sample_clusters):
    """the same two calls through the reference's own code (or the scalar restatement) on all host threads, on the first
    `sample_clusters` clusters"""
    import oracle_lib
    from isaac_aligner_b200 import synth
    from isaac_aligner_b200.batch import MatchBatch
    from isaac_aligner_b200.types import ReadSet
    chk = oracle_lib.Oracle(oracle_lib.REF_SO) if os.path.exists(oracle_lib.REF_SO) else oracle_lib.port()
    cores = os.cpu_count() or 1
    k = min(sample_clusters, reads.cluster_count)
    sub_reads = ReadSet(reads.bcl[:k], reads.read_lengths)
    sub_mb = MatchBatch(mb.matches[:int(mb.begin[k])], mb.begin[:k + 1], mb.seeds, with_gaps=True)
    g = oracle_lib.GenomeHolder(genome)
    t0 = time.perf_counter()
    flat = oracle_lib.build_fragments(chk, g, sub_reads, config, sub_mb, threads=cores)
    t1 = time.perf_counter()
    req = synth.rescue_policy(flat.fragments, flat.begin, k)
    t2 = time.perf_counter()
    oracle_lib.rescue_shadows(chk, g, sub_reads, config, tls, req, threads=cores, fragments_per_request=96)
    t3 = time.perf_counter()
    sec = (t1 - t0) + (t3 - t2)
    out = {"pairs": k, "pairs_per_s": k / sec, "build_ms": (t1 - t0) * 1e3, "rescue_ms": (t3 - t2) * 1e3, "cores": cores,
           "kind": chk.kind}
    if chk.kind == "reference":                          # the verbatim TemplateBuilder, one per host thread
        from isaac_aligner_b200.batch import TemplateOptions
        oracle_lib.build_templates(chk, g, sub_reads, config, sub_mb, tls, TemplateOptions.make(), threads=cores)   # warm the genome cache
        t4 = time.perf_counter()
        oracle_lib.build_templates(chk, g, sub_reads, config, sub_mb, tls, TemplateOptions.make(), threads=cores)
        out["templates_ms"] = (time.perf_counter() - t4) * 1e3
        out["template_pairs_per_s"] = k / (out["templates_ms"] * 1e-3)
    return out


def run_b200(args):
    import torch
    from isaac_aligner_b200 import capi
    from isaac_aligner_b200.types import CANDIDATE_DTYPE, FRAGMENT_DTYPE, Config

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the candidate-extension path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    L, n, stride = args.read_length, args.candidates, args.cigar_stride
    genome, reads, cand = make_workload(args, rank, n)
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
        # isaac_ext_*_batch_compact: 64-byte records + a dense CIGAR pool, chunked with overlapped copies
        out_u = (pin((n, 64), torch.uint8), pin((n * 3,), torch.int32))
        out_g = (pin((n, 64), torch.uint8), pin((n * 12,), torch.int32))
        views = [(o[0].numpy().reshape(-1).view(FRAGMENT_DTYPE), o[1].numpy().view(np.uint32)) for o in (out_u, out_g)]
        words = [0, 0]

        def e2e_step():
            words[0], words[1] = ctx.extend_compact_both(h_cand, views[0][0], views[0][1], views[1][0], views[1][1])

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
        # the device results of the resident path and the host results of the e2e path must agree (all but cigarOffset)
        res_dev = d_frag_g.cpu().numpy().reshape(-1).view(FRAGMENT_DTYPE)
        for name in FRAGMENT_DTYPE.names:
            if name != "cigarOffset":
                assert np.array_equal(res_dev[name], views[1][0][name]), "resident and end-to-end results differ: " + name
        e2e = {"value": world * cells / (e2e_ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": n * 16, "d2h_bytes_per_step": 2 * n * 64 + 4 * (words[0] + words[1]),
               "api": "isaac_ext_extend_batch_compact (ungapped + gapped records of every candidate), pinned host buffers"}
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    if rank != 0:
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
    # ---- CPU baseline on a bounded sample, same box
    cores = os.cpu_count() or 1
    ns = min(n, args.cpu_sample_per_core * cores)
    cpu_gcups, cpu_sec, kind, cores = cpu_arm(args, genome, reads, cand[:ns], config, 1, 0)
    # ---- side measurement: the two TemplateBuilder-facing calls on simulated pairs (BASELINE "aligned read pairs/sec")
    pairs_line = None
    n_pairs = 200_000 if args.pairs is None else args.pairs
    if n_pairs:
        # a side measurement must not cost the bench line: whatever goes wrong in it is reported in its place
        try:
            pgenome, preads, pmb, ptls = make_pairs_workload(args, rank, n_pairs)
            ctx.set_reads(preads)
            pairs_line, _, _ = pairs_pipeline_gpu(ctx, preads, pmb, ptls, max(1, args.steps // 2), 1)
            pairs_line["cpu_baseline"] = pairs_pipeline_cpu(pgenome, preads, pmb, ptls, config, 4000 * cores)
        except Exception as e:      # noqa: BLE001
            pairs_line = {"error": "%s: %s" % (type(e).__name__, e)}

    emit(json.dumps({
        "metric": "banded_sw_gcups", "value": world * cells / (ms_per_step * 1e-3) / 1e9, "unit": "GCUPS",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": workload_config(args),
        "e2e": e2e, "gpu_launches": int(gpu_launches), "clocks": clocks,
        "tile_stats": tile_stats,
        "pairs_pipeline": pairs_line,
        "roofline": {"bound": "int32", "kernel": "swForwardKernel (timed with the swTraceScoreKernel launches it overlaps: "
                                                  "the whole isaac_ext_gapped_batch_device call)",
                     "achieved": achieved / 1e12, "peak": peak_add / 1e12,
                     "unit": "TOP/s", "frac": achieved / peak_add, "traffic": traffic,
                     "hbm_gbs": (traffic / (ms_gapped * 1e-3) / 1e9) if traffic else None,
                     "hbm_frac": (traffic / (ms_gapped * 1e-3) / 1e9 / hbm_peak) if traffic else None,
                     "plane_bytes_per_launch": int(n) * L * 12 * 2,
                     "ops_per_cell": OPS_PER_CELL, "gcups_kernel": sw_gcups_kernel, "ms_per_launch": ms_gapped,
                     "peak_source": "measured live by isaac_ext_measure_int32_peak: add.s32 %.1f, max.s32 %.1f, "
                                    "16x2 max %.1f TOP/s" % (peak_add / 1e12, peak_max / 1e12, peak_dpx / 1e12)},
        "roofline_ungapped": {"bound": "hbm", "kernel": "ungappedKernel", "achieved": ungapped_gbs, "peak": hbm_peak,
                              "unit": "GB/s", "frac": ungapped_gbs / hbm_peak, "traffic": traffic_ungapped, "ms_per_launch": ms_ungapped,
                              "limited_by": "integer pipe (ncu: ALU 66 %, issue 61 %; profiles/r1_u_ncu_ungappedKernel.txt): the ordered "
                                            "FP64 sum and the longest-run counter cost about 290 instructions per 16 bases",
                              "peak_source": hbm_src, "candidates_per_s": n / (ms_ungapped * 1e-3)},
        "cpu_baseline": {"value": cpu_gcups, "unit": "GCUPS", "cores": cores, "kind": kind,
                         "sample": "first %d of the %d candidates of rank 0, one pass, %d host threads, %.2f s"
                                   % (ns, n, cores, cpu_sec)},
    }))
    ctx.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def pairs_config(args, n_pairs):
    return {"workload": "BASELINE configs[0]/[2] style: %d simulated 2x%d bp FR pairs per GPU per step on a %d bp random genome, "
                        "indel events %.0e/base, seed matches from error-free auto seeds + 20%% decoys, explicit TLS 245/350/455, "
                        "TemplateBuilder::buildFragments + buildTemplate of every cluster (isaac_ext_build_templates)"
                        % (n_pairs, args.read_length, args.genome_bases, args.indel_rate),
            "pairs_per_gpu": n_pairs, "read_length": args.read_length,
            "l2": "per-step inputs+outputs exceed the 126 MB L2 for >= 500k pairs"}


def run_pairs(args):
    """--workload pairs: aligned read pairs/s of isaac_ext_build_templates (seed matches in, templates out, host-pointer ABI =
    end to end), with the two batch calls underneath it timed separately in "pipeline"."""
    import torch
    from isaac_aligner_b200 import capi
    from isaac_aligner_b200.types import Config
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_pairs = 1_000_000 if args.pairs is None else args.pairs
    if args.impl == "reference":
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        genome, reads, mb, tls = make_pairs_workload(args, 0, min(n_pairs, 6000 * cores))
        config = Config.default(max_read_length=2 * args.read_length)
        runs = [pairs_pipeline_cpu(genome, reads, mb, tls, config, reads.cluster_count) for _ in range(args.warmup + args.steps)][args.warmup:]
        v = float(np.mean([r.get("template_pairs_per_s", r["pairs_per_s"]) for r in runs]))
        sample = "%d pairs per step of the same generator, %d host threads" % (reads.cluster_count, cores)
        emit(json.dumps({"impl": "reference", "metric": "aligned_read_pairs_per_s", "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": reads.cluster_count / v * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16+f64", "data": "synthetic",
                          "config": pairs_config(args, n_pairs),
                          "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": runs[0]["kind"], "sample": sample},
                          "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the candidate-extension path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    genome, reads, mb, tls = make_pairs_workload(args, rank, n_pairs)
    config = Config.default(max_read_length=2 * args.read_length, device=local_rank,
                            host_threads=max(1, (os.cpu_count() or 1) // world))
    ctx = capi.Context(config)
    ctx.set_reference(genome)
    ctx.set_reads(reads)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    line, flat, req = pairs_pipeline_gpu(ctx, reads, mb, tls, args.steps, args.warmup, all_ranks=True)
    t = torch.tensor([line["templates_ms"]], dtype=torch.float64, device="cuda")
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    cpu = pairs_pipeline_cpu(genome, reads, mb, tls, config, 4000 * cores)
    # one build_templates call: matches in; candidate records, rescue requests and shadow records cross PCIe inside it
    h2d = len(mb.matches) * 16 + line["template_rescue_requests"] * 32
    d2h = flat.fragments.size * 64 + flat.cigars.size * 4
    v = world * n_pairs / (ms * 1e-3)
    emit(json.dumps({"metric": "aligned_read_pairs_per_s", "value": v, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "int16+f64", "data": "synthetic", "config": pairs_config(args, n_pairs),
                      "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                      "gpu_launches": int(line["gpu_launches_per_step"] * args.steps), "pipeline": line,
                      "cpu_baseline": {"value": cpu.get("template_pairs_per_s", cpu["pairs_per_s"]), "unit": "pairs/s", "cores": cpu["cores"], "kind": cpu["kind"],
                                       "sample": "first %d pairs of rank 0, one pass, %d host threads" % (cpu["pairs"], cpu["cores"])}}))
    ctx.close()


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
    genome, reads, mb, tls = make_pairs_workload(args, 0, n_pairs)
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


if __name__ == "__main__":
    a = parse_args()
    if a.workload == "pack":
        run_pack(a)
    elif a.workload == "pairs":
        run_pairs(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)

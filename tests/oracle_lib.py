"""ctypes loader of the two CPU checkers (oracle/oracle_api.h).  Test infrastructure only."""
import ctypes
import os
import subprocess

import numpy as np

from isaac_aligner_b200.types import CANDIDATE_DTYPE, FRAGMENT_DTYPE, MASK_WORDS, Config, Reads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
PORT_SO = os.path.join(ORACLE_DIR, "libisaac_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libisaac_ref.so")


class Genome(ctypes.Structure):
    _fields_ = [("contigCount", ctypes.c_uint32), ("contigBases", ctypes.POINTER(ctypes.c_void_p)),
                ("contigLengths", ctypes.POINTER(ctypes.c_uint64))]


class GenomeHolder:
    def __init__(self, contigs):
        self.contigs = [np.ascontiguousarray(c, dtype=np.uint8) for c in contigs]
        n = len(self.contigs)
        self.ptrs = (ctypes.c_void_p * n)(*[c.ctypes.data for c in self.contigs])
        self.lens = (ctypes.c_uint64 * n)(*[c.size for c in self.contigs])
        self.c = Genome(n, self.ptrs, self.lens)


def build(target="port"):
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, target])


class Oracle:
    """One of the two checker libraries; both export the same symbols."""

    def __init__(self, path):
        self.lib = ctypes.CDLL(path)
        self.lib.oracle_kind.restype = ctypes.c_char_p
        self.kind = self.lib.oracle_kind().decode()

    def set_adapters(self, adapters):
        """process-wide adapter list of this checker: (sequence, reverse, clipLength) tuples; () clears"""
        from isaac_aligner_b200.types import adapter_array
        if not hasattr(self.lib, "oracle_set_adapters"):
            raise RuntimeError("this checker has no adapter support")
        arr = adapter_array(adapters)
        rc = self.lib.oracle_set_adapters(ctypes.c_uint32(len(adapters)), arr)
        if rc:
            raise RuntimeError("oracle_set_adapters failed: %d" % rc)

    def banded_sw(self, queries, dbs, scores, max_read_length=None, cigar_stride=64, threads=1, band=None):
        """queries/dbs: lists of bytes.  scores = (match, mismatch, gapOpen>0, gapExtend>0).
        returns (list of cigar word arrays, offsets)"""
        n = len(queries)
        qbuf = np.frombuffer(b"".join(queries), dtype=np.uint8)
        dbuf = np.frombuffer(b"".join(dbs), dtype=np.uint8)
        qlen = np.array([len(q) for q in queries], dtype=np.uint32)
        qoff = np.concatenate([[0], np.cumsum(qlen[:-1], dtype=np.uint64)]).astype(np.uint64)
        dlen = np.array([len(d) for d in dbs], dtype=np.uint64)
        assert np.all(dlen == qlen + (band or 16) - 1)
        doff = np.concatenate([[0], np.cumsum(dlen[:-1], dtype=np.uint64)]).astype(np.uint64)
        return self.banded_sw_flat(qbuf, qoff, qlen, dbuf, doff, scores, max_read_length, cigar_stride, threads, band)

    def banded_sw_flat(self, qbuf, qoff, qlen, dbuf, doff, scores, max_read_length=None, cigar_stride=64, threads=1, band=None):
        """band=None: BandedSmithWaterman::align (both libraries); band=16/32/64: the band-width-parametrised model of the port"""
        n = len(qlen)
        if max_read_length is None:
            max_read_length = int(qlen.max())
        cig = np.zeros((n, cigar_stride), dtype=np.uint32)
        ciglen = np.zeros(n, dtype=np.uint32)
        off = np.zeros(n, dtype=np.uint32)
        if band is not None:
            rc = self.lib.oracle_banded_sw_wide_batch(
                ctypes.c_uint32(band), ctypes.c_uint32(n), ctypes.c_void_p(qbuf.ctypes.data), ctypes.c_void_p(qoff.ctypes.data),
                ctypes.c_void_p(qlen.ctypes.data), ctypes.c_void_p(dbuf.ctypes.data), ctypes.c_void_p(doff.ctypes.data),
                ctypes.c_int(scores[0]), ctypes.c_int(scores[1]), ctypes.c_int(scores[2]), ctypes.c_int(scores[3]),
                ctypes.c_uint32(max_read_length), ctypes.c_uint32(cigar_stride), ctypes.c_void_p(cig.ctypes.data),
                ctypes.c_void_p(ciglen.ctypes.data), ctypes.c_void_p(off.ctypes.data), ctypes.c_uint32(threads))
            if rc:
                raise RuntimeError("oracle_banded_sw_wide_batch failed: %d" % rc)
            return cig, ciglen, off
        rc = self.lib.oracle_banded_sw_batch(
            ctypes.c_uint32(n), ctypes.c_void_p(qbuf.ctypes.data), ctypes.c_void_p(qoff.ctypes.data),
            ctypes.c_void_p(qlen.ctypes.data), ctypes.c_void_p(dbuf.ctypes.data), ctypes.c_void_p(doff.ctypes.data),
            ctypes.c_int(scores[0]), ctypes.c_int(scores[1]), ctypes.c_int(scores[2]), ctypes.c_int(scores[3]),
            ctypes.c_uint32(max_read_length), ctypes.c_uint32(cigar_stride), ctypes.c_void_p(cig.ctypes.data),
            ctypes.c_void_p(ciglen.ctypes.data), ctypes.c_void_p(off.ctypes.data), ctypes.c_uint32(threads))
        if rc:
            raise RuntimeError("oracle_banded_sw_batch failed: %d" % rc)
        return cig, ciglen, off

    def _extend(self, fn, genome, reads, config, candidates, cigar_stride, threads, gapped):
        n = len(candidates)
        cand = np.ascontiguousarray(candidates, dtype=CANDIDATE_DTYPE)
        frags = np.zeros(n, dtype=FRAGMENT_DTYPE)
        cig = np.zeros((n, cigar_stride), dtype=np.uint32)
        mask = np.zeros((n, MASK_WORDS), dtype=np.uint64)
        args = [ctypes.byref(genome.c), ctypes.byref(reads.c), ctypes.byref(config), ctypes.c_uint32(n),
                ctypes.c_void_p(cand.ctypes.data)]
        if gapped:
            args.append(ctypes.c_uint32(cigar_stride))
        args += [ctypes.c_void_p(frags.ctypes.data), ctypes.c_void_p(cig.ctypes.data),
                 ctypes.c_void_p(mask.ctypes.data), ctypes.c_uint32(threads)]
        rc = fn(*args)
        if rc:
            raise RuntimeError("oracle extend batch failed: %d" % rc)
        return frags, cig, mask

    def ungapped(self, genome, reads, config, candidates, threads=1):
        return self._extend(self.lib.oracle_ungapped_batch, genome, reads, config, candidates, 3, threads, False)

    def gapped(self, genome, reads, config, candidates, cigar_stride=32, threads=1):
        return self._extend(self.lib.oracle_gapped_batch, genome, reads, config, candidates, cigar_stride, threads, True)


def _flat_call(fn, head_args, n_groups, n_flags, fragment_capacity, cigar_capacity, threads):
    from isaac_aligner_b200.batch import FlatFragments
    frags = np.zeros(fragment_capacity, dtype=FRAGMENT_DTYPE)
    begin = np.zeros(n_groups + 1, dtype=np.uint64)
    cigars = np.zeros(cigar_capacity, dtype=np.uint32)
    flags = np.zeros(n_flags, dtype=np.uint8)
    nf, nc = ctypes.c_uint64(), ctypes.c_uint64()
    rc = fn(*head_args, ctypes.c_uint64(fragment_capacity), ctypes.c_void_p(frags.ctypes.data),
            ctypes.c_void_p(begin.ctypes.data), ctypes.c_uint64(cigar_capacity), ctypes.c_void_p(cigars.ctypes.data),
            ctypes.c_void_p(flags.ctypes.data), ctypes.byref(nf), ctypes.byref(nc), ctypes.c_uint32(threads))
    if rc:
        raise RuntimeError("oracle call failed: %d (fragments %d, cigar words %d)" % (rc, nf.value, nc.value))
    return FlatFragments(frags[:nf.value].copy(), begin, cigars[:nc.value].copy(), flags)


def build_fragments(oracle, genome, reads, config, match_batch, threads=1):
    """FragmentBuilder::build for every cluster -> FlatFragments (begin per cluster*readCount, flags = built)"""
    cap = len(match_batch.matches) + 16
    return _flat_call(oracle.lib.oracle_build_fragments,
                      [ctypes.byref(genome.c), ctypes.byref(reads.c), ctypes.byref(config), ctypes.byref(match_batch.c)],
                      reads.cluster_count * reads.read_count, reads.cluster_count, cap, cap * 12, threads)


def rescue_shadows(oracle, genome, reads, config, tls, requests, threads=1, fragments_per_request=64):
    """ShadowAligner::rescueShadow for every request -> FlatFragments (begin per request, flags = rescued)"""
    from isaac_aligner_b200.batch import RESCUE_REQUEST_DTYPE
    req = np.ascontiguousarray(requests, dtype=RESCUE_REQUEST_DTYPE)
    cap = len(req) * fragments_per_request + 1024
    return _flat_call(oracle.lib.oracle_rescue_shadows,
                      [ctypes.byref(genome.c), ctypes.byref(reads.c), ctypes.byref(config), ctypes.byref(tls),
                       ctypes.c_uint32(len(req)), ctypes.c_void_p(req.ctypes.data)],
                      len(req), len(req), cap, cap * 4, threads)


def build_templates(oracle, genome, reads, config, match_batch, tls, options, threads=1, cigar_capacity=None):
    """TemplateBuilder over every cluster (reference build only) -> batch.Templates"""
    from isaac_aligner_b200.batch import TEMPLATE_DTYPE, Templates
    n = reads.cluster_count
    templates = np.zeros(n, dtype=TEMPLATE_DTYPE)
    frags = np.zeros(n * reads.read_count, dtype=FRAGMENT_DTYPE)
    cigars = np.zeros(cigar_capacity or 64 * n + 1024, dtype=np.uint32)
    nc = ctypes.c_uint64()
    rc = oracle.lib.oracle_build_templates(
        ctypes.byref(genome.c), ctypes.byref(reads.c), ctypes.byref(config), ctypes.byref(match_batch.c), ctypes.byref(tls),
        ctypes.byref(options), ctypes.c_void_p(templates.ctypes.data), ctypes.c_void_p(frags.ctypes.data),
        ctypes.c_uint64(cigars.size), ctypes.c_void_p(cigars.ctypes.data), ctypes.byref(nc), ctypes.c_uint32(threads))
    if rc:
        raise RuntimeError("oracle_build_templates failed: %d" % rc)
    return Templates(templates, frags, cigars[:nc.value].copy())


def template_stats(oracle, genome, reads, config, match_batch, tls, options, pf=None, threads=1):
    """TileBarcodeStats of the tile's templates through the reference's own classes (reference build only) -> uint64 [4, 32]"""
    pf_arr = None if pf is None else np.ascontiguousarray(pf, dtype=np.uint8)
    out = np.zeros((4, 32), dtype=np.uint64)
    rc = oracle.lib.oracle_template_stats(
        ctypes.byref(genome.c), ctypes.byref(reads.c), ctypes.byref(config), ctypes.byref(match_batch.c), ctypes.byref(tls),
        ctypes.byref(options), ctypes.c_void_p(pf_arr.ctypes.data) if pf_arr is not None else None,
        ctypes.c_void_p(out.ctypes.data), ctypes.c_uint32(threads))
    if rc:
        raise RuntimeError("oracle_template_stats failed: %d" % rc)
    return out


TILE_CYCLE_STATS_WORDS = 47105
TILE_CYCLE_STATS_FIELDS = [("alignmentScoreFragments", 8192), ("alignmentScoreMismatches", 8192), ("alignmentScoreTemplates", 8192),
                           ("alignmentScoreTemplateMismatches", 8192), ("cycleBlanks", 1024), ("cycleUniquelyAlignedBlanks", 1024),
                           ("cycleMismatches", 1024), ("cycleUniquelyAlignedMismatches", 1024)] + \
                          [("cycleUniquelyAligned%sMismatchFragments" % k, 1024) for k in ("1", "2", "3", "4", "More")] + \
                          [("cycle%sMismatchFragments" % k, 1024) for k in ("1", "2", "3", "4", "More")] + [("uniquelyAlignedFragmentCount", 1)]


def tile_cycle_stats(oracle, genome, reads, config, match_batch, tls, options, pf=None, finalize=False, threads=1):
    """matchSelector::TileStats of the tile's templates through the reference's own class (reference build only) -> uint64 [4, 47105]"""
    pf_arr = None if pf is None else np.ascontiguousarray(pf, dtype=np.uint8)
    out = np.zeros((4, TILE_CYCLE_STATS_WORDS), dtype=np.uint64)
    rc = oracle.lib.oracle_tile_cycle_stats(
        ctypes.byref(genome.c), ctypes.byref(reads.c), ctypes.byref(config), ctypes.byref(match_batch.c), ctypes.byref(tls),
        ctypes.byref(options), ctypes.c_void_p(pf_arr.ctypes.data) if pf_arr is not None else None,
        ctypes.c_void_p(out.ctypes.data), ctypes.c_uint32(1 if finalize else 0), ctypes.c_uint32(threads))
    if rc:
        raise RuntimeError("oracle_tile_cycle_stats failed: %d" % rc)
    return out


def determine_template_length(oracle, genome, reads, config, match_batch, pf=None, mate_drift_range=-1):
    """MatchSelector::determineTemplateLength for the tile (reference build only) -> (batch.Tls, stable)"""
    from isaac_aligner_b200.batch import Tls
    tls = Tls()
    stable = ctypes.c_uint32()
    pf_arr = None if pf is None else np.ascontiguousarray(pf, dtype=np.uint8)
    rc = oracle.lib.oracle_determine_template_length(
        ctypes.byref(genome.c), ctypes.byref(reads.c), ctypes.byref(config), ctypes.byref(match_batch.c),
        ctypes.c_void_p(pf_arr.ctypes.data) if pf_arr is not None else None, ctypes.c_int32(mate_drift_range),
        ctypes.byref(tls), ctypes.byref(stable))
    if rc:
        raise RuntimeError("oracle_determine_template_length failed: %d" % rc)
    return tls, bool(stable.value)


def trim_low_quality_ends(oracle, reads, base_quality_cutoff):
    """alignment::trimLowQualityEnds on every cluster (reference build only) -> endCyclesMasked [clusters, readCount]"""
    out = np.zeros((reads.cluster_count, reads.read_count), dtype=np.uint16)
    rc = oracle.lib.oracle_trim_low_quality_ends(ctypes.byref(reads.c), ctypes.c_uint32(base_quality_cutoff),
                                                 ctypes.c_void_p(out.ctypes.data))
    if rc:
        raise RuntimeError("oracle_trim_low_quality_ends failed: %d" % rc)
    return out


def pack_fragments(oracle, reads, templates, options, barcode_bytes=None, records=True):
    """FragmentCollector::add for every stored template through the reference's own io::FragmentHeader (reference build only)
    -> (batch.PackedFragments, header mask [headerLength] with 0xFF on the bytes that carry a member)"""
    from isaac_aligner_b200.batch import PackedFragments
    n, rc = reads.cluster_count, reads.read_count
    layout = np.zeros(4, dtype=np.uint32)
    mask = np.zeros(256, dtype=np.uint8)
    t = np.ascontiguousarray(templates.templates)
    f = np.ascontiguousarray(templates.fragments)
    cig = np.ascontiguousarray(templates.cigars, dtype=np.uint32)
    bar = None if barcode_bytes is None else np.ascontiguousarray(barcode_bytes, dtype=np.uint8).reshape(n, -1)

    def call(rec, pos, init):
        rc_ = oracle.lib.oracle_pack_fragments(
            ctypes.byref(reads.c), ctypes.c_void_p(t.ctypes.data), ctypes.c_void_p(f.ctypes.data),
            ctypes.c_void_p(cig.ctypes.data) if cig.size else None, ctypes.c_uint64(cig.size), ctypes.byref(options.c),
            ctypes.c_void_p(bar.ctypes.data) if bar is not None else None, ctypes.c_uint32(bar.shape[1] if bar is not None else 0),
            ctypes.c_void_p(rec.ctypes.data) if rec is not None else None, ctypes.c_void_p(pos.ctypes.data) if pos is not None else None,
            ctypes.c_void_p(init.ctypes.data) if init is not None else None, ctypes.c_void_p(mask.ctypes.data),
            ctypes.c_void_p(layout.ctypes.data))
        if rc_:
            raise RuntimeError("oracle_pack_fragments failed: %d" % rc_)

    call(None, None, None)
    rec = np.zeros((n, int(layout[0])), dtype=np.uint8)
    pos = np.zeros((n, rc), dtype=np.uint64)
    init = np.zeros((n, rc), dtype=np.uint8)
    if records:
        call(rec, pos, init)
    return (PackedFragments(rec, pos, init, int(layout[0]), (int(layout[1]), int(layout[2])), int(layout[3]), int(init.sum())),
            mask[:int(layout[3])].copy())


def port():
    if not os.path.exists(PORT_SO):
        build("port")
    return Oracle(PORT_SO)


def require_reference():
    """The reference's own code for the `-m gpu` parity tests: a missing oracle/_ref/libisaac_ref.so is a FAILURE there, never a
    silent downgrade to "parity against our own restatement" (build it with `make -C oracle ref` where /root/reference is
    mounted; the built file travels to the GPU box with the snapshot)."""
    ref = reference()
    if ref is None:
        import pytest
        pytest.fail("oracle/_ref/libisaac_ref.so is missing: the GPU parity tests compare with the reference's own code")
    return ref


def gpu_checkers():
    """both CPU checkers, for the `-m gpu` parity tests"""
    return [port(), require_reference()]


def reference():
    """The reference's own code; None when it was never built (needs /root/reference at build time)."""
    if not os.path.exists(REF_SO):
        if os.path.isdir("/root/reference/src/c++"):
            build("ref")
        else:
            return None
    return Oracle(REF_SO)


def _realign_call(fn, head_args, bin_, options, cigar_capacity=None, gap_capacity=None):
    """shared argument plumbing of oracle_realign_bin and tests/cpp's realign_bin_host; returns bins.RealignResult + counts"""
    from isaac_aligner_b200 import bins
    n = len(bin_.index)
    data = bin_.data.copy()
    cigar_capacity = cigar_capacity or (n * 96 + 64)
    gap_capacity = gap_capacity or (len(bin_.record_offset) * 8 + 16)
    position = np.zeros(n, dtype=np.uint64)
    cigar_offset = np.zeros(n, dtype=np.uint32)
    cigar_length = np.zeros(n, dtype=np.uint32)
    cigars = np.zeros(cigar_capacity, dtype=np.uint32)
    gaps = np.zeros(gap_capacity, dtype=bins.GAP_DTYPE)
    deletions = np.zeros(gap_capacity, dtype=bins.GAP_DTYPE)
    counts = np.zeros(8, dtype=np.uint64)
    index = np.ascontiguousarray(bin_.index)
    offsets = np.ascontiguousarray(bin_.record_offset, dtype=np.uint64)
    rc = fn(*head_args, ctypes.byref(options.c), ctypes.c_void_p(data.ctypes.data), ctypes.c_uint64(data.size),
            ctypes.c_void_p(offsets.ctypes.data), ctypes.c_uint64(offsets.size), ctypes.c_void_p(index.ctypes.data), ctypes.c_uint64(n),
            ctypes.c_void_p(position.ctypes.data), ctypes.c_void_p(cigar_offset.ctypes.data), ctypes.c_void_p(cigar_length.ctypes.data),
            ctypes.c_void_p(cigars.ctypes.data), ctypes.c_uint64(cigar_capacity), ctypes.c_void_p(gaps.ctypes.data),
            ctypes.c_void_p(deletions.ctypes.data), ctypes.c_uint64(gap_capacity), ctypes.c_void_p(counts.ctypes.data))
    if rc:
        raise RuntimeError("realign call failed: %d" % rc)
    res = bins.RealignResult(data, position, cigar_offset, cigar_length, cigars[:int(counts[2])], gaps[:int(counts[0])],
                             deletions[:int(counts[1])])
    return res, counts


def realign_bin(oracle, genome, bin_, options):
    """BinSorter::collectGaps + realignGaps through the reference's own classes (liboracle_ref only)"""
    res, _ = _realign_call(oracle.lib.oracle_realign_bin, (ctypes.byref(genome.c),), bin_, options)
    return res

"""isaac_ext_pack_fragments (matchSelector::FragmentCollector::add, io::FragmentHeader bin records) against the
reference's own io::FragmentHeader constructors (oracle/_ref), byte for byte.

CPU (-m "not gpu"): the host/device functions of csrc/pack_fragments.cuh -- everything the warps of packFragmentsKernel run --
driven lane after lane by tests/cpp/test_pack_fragments.cpp.  GPU (-m gpu): the kernel through the C ABI, on random template
records and on the templates isaac_ext_build_templates leaves for a simulated tile.

Compared: every header byte that carries a member of io::FragmentHeader (its padding bytes are unspecified in the reference,
zero here), the BCL bytes, the CIGAR words, the zeros behind them, and the FragmentBuffer index (fStrandPos_, initialized())."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle_lib
from isaac_aligner_b200.batch import TEMPLATE_DTYPE, PackOptions, Templates
from isaac_aligner_b200.types import FRAGMENT_DTYPE, ReadSet

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_reference = pytest.mark.skipif(not os.path.exists(oracle_lib.REF_SO) and not os.path.isdir("/root/reference/src/c++"),
                                     reason="io::FragmentHeader comes from the reference build only")
NO_MATCH_CONTIG = 0x7FFFFF
CONTIG_LENGTHS = (250_000, 90_000, 1_000)
BIN_SIZE = 4096


def random_bcl(rng, n, total):
    bcl = (rng.integers(2, 42, size=(n, total)).astype(np.uint8) << 2) | rng.integers(0, 4, size=(n, total)).astype(np.uint8)
    bcl[rng.random((n, total)) < 0.02] = 0                               # BCL N: quality bits zero, any base bits
    where = rng.random((n, total)) < 0.01
    bcl[where] = rng.integers(0, 4, size=int(where.sum())).astype(np.uint8)
    return bcl


def random_cigar(rng, L, aligned_bases_zero=False):
    """-> (words, observedLength, gapCount, lowClip, highClip) of a plausible CIGAR for a read of L bases"""
    if aligned_bases_zero:
        return [(L << 4) | 4], 0, 0, 0, 0
    lead = int(rng.integers(0, 9)) if rng.random() < 0.3 else 0
    trail = int(rng.integers(0, 9)) if rng.random() < 0.3 else 0
    body = L - lead - trail
    words, observed, gaps = [], 0, 0
    if lead:
        words.append((lead << 4) | 4)
    n_gaps = int(rng.integers(0, 4)) if rng.random() < 0.4 and body >= 60 else 0
    left = body
    for g in range(n_gaps):
        m = int(rng.integers(5, max(6, (left - 15) // (n_gaps - g + 1))))
        words.append((m << 4) | 0)
        observed += m
        left -= m
        length = int(rng.integers(1, 12))
        if rng.random() < 0.5 and left > length + 5:
            words.append((length << 4) | 1)                              # insertion: read bases only
            left -= length
        else:
            words.append((length << 4) | 2)                              # deletion: reference bases only
            observed += length
        gaps += 1
    words.append((left << 4) | 0)
    observed += left
    if trail:
        words.append((trail << 4) | 4)
    return words, observed, gaps, lead, trail


def random_templates(rng, n, read_lengths):
    """Template records the way TemplateBuilder leaves them: pairs, singleton + shadow, unaligned no-match clusters"""
    rc = len(read_lengths)
    t = np.zeros(n, dtype=TEMPLATE_DTYPE)
    f = np.zeros(n * rc, dtype=FRAGMENT_DTYPE)
    pool = [0]                                                           # word 0 never belongs to a fragment
    scores = np.array([0, 1, 3, 4, 60, 250, 1300, 0xFFFF, 0x12345, 0xFFFFFFFF], dtype=np.uint64)
    for c in range(n):
        kind = rng.random()
        aligned = [True] * rc
        if kind < 0.12:
            aligned = [False] * rc                                       # nothing aligned: no-match records
        elif rc == 2 and kind < 0.3:
            aligned[int(rng.integers(0, 2))] = False                     # singleton + shadow
        contig = int(rng.integers(0, len(CONTIG_LENGTHS)))
        anchor = int(rng.integers(0, CONTIG_LENGTHS[contig] - 700))
        for r in range(rc):
            rec = f[c * rc + r]
            L = read_lengths[r]
            rec["readId"], rec["readIndex"] = c * rc + r, r
            rec["reverse"] = int(rng.integers(0, 2))
            rec["editDistance"], rec["mismatchCount"] = int(rng.integers(0, 40)), int(rng.integers(0, 20))
            if aligned[r]:
                words, observed, gaps, low, high = random_cigar(rng, L, aligned_bases_zero=rng.random() < 0.03)
                other = rng.random() < 0.1                               # mates on different contigs now and then
                rec["contigId"] = int(rng.integers(0, len(CONTIG_LENGTHS))) if other else contig
                limit = CONTIG_LENGTHS[int(rec["contigId"])] - 700
                rec["position"] = 0 if rng.random() < 0.03 else min(limit, anchor + int(rng.integers(0, 500))) if not other else int(rng.integers(0, limit))
                rec["observedLength"], rec["gapCount"] = observed, gaps
                rec["lowClipped"], rec["highClipped"] = (high, low) if rec["reverse"] else (low, high)
                rec["cigarOffset"], rec["cigarLength"] = len(pool), len(words)
                pool.extend(words)
            else:
                rec["cigarOffset"] = int(rng.integers(0, 5))             # stale offsets of unaligned records must not be read
                rec["observedLength"] = int(rng.integers(0, 3)) * 77    # nor their stale observed length
        for r in range(rc):
            rec = f[c * rc + r]
            if not aligned[r]:
                mate = f[c * rc + (1 - r)] if rc == 2 else None
                if mate is not None and aligned[1 - r]:                  # shadow: sits at its orphan's position
                    rec["contigId"], rec["position"] = mate["contigId"], mate["position"]
                else:
                    rec["contigId"], rec["position"] = NO_MATCH_CONTIG, 0
        t[c]["built"] = 1 if any(aligned) and rng.random() < 0.95 else 0
        t[c]["hadFragments"] = 1 if any(aligned) else 0
        t[c]["properPair"] = 1 if rc == 2 and all(aligned) and rng.random() < 0.7 else 0
        t[c]["alignmentScore"] = scores[int(rng.integers(0, len(scores)))]
        for r in range(rc):
            t[c]["fragmentAlignmentScore"][r] = scores[int(rng.integers(0, len(scores)))]
    return Templates(t, f, np.array(pool + [0, 0], dtype=np.uint32))


def bin_index():
    """a BinIndexMap: one output bin index per distribution bin of every contig, growing along the genome"""
    out, current = [], 1
    for length in CONTIG_LENGTHS:
        bins = length // BIN_SIZE + 1
        out.append((current + np.arange(bins) // 7).astype(np.uint32))
        current = int(out[-1][-1]) + 1
    return out


def make_options(rng, n, keep_unaligned, with_arrays=True, barcode_length=6, compact=False):
    if not with_arrays:
        return PackOptions(tile=7, barcode_idx=3, keep_unaligned=keep_unaligned, compact=compact), None
    xy = rng.integers(-5000, 250000, size=(n, 2)).astype(np.int32)
    xy[rng.random(n) < 0.1] = 0x7FFFFFFF
    barcode = random_bcl(rng, n, barcode_length)
    sequence = np.zeros(n, dtype=np.uint64)
    for i in range(barcode_length):                                      # oligo::packBclBases (Nucleotides.hh:280-293)
        sequence |= (barcode[:, i].astype(np.uint64) & np.uint64(3)) << np.uint64(2 * i)
    return PackOptions(tile=1101, barcode_idx=5, keep_unaligned=keep_unaligned, pf=(rng.random(n) < 0.8).astype(np.uint8), xy=xy,
                       barcode_sequence=sequence, distribution_bin_size=BIN_SIZE, bin_index=bin_index(), compact=compact), barcode


def assert_packed_equal(got, want, mask, read_lengths, what):
    assert (got.record_length, got.read_offset, got.header_length) == (want.record_length, want.read_offset, want.header_length), what
    assert np.array_equal(got.initialized, want.initialized), what + ": IndexRecord::initialized()"
    assert np.array_equal(got.f_strand_pos, want.f_strand_pos), what + ": IndexRecord::fStrandPos_"
    assert got.stored == int(want.initialized.sum()), what + ": storedFragments"
    H = want.header_length
    names = {0: "bamTlen_", 4: "observedLength_", 8: "fStrandPosition_", 16: "lowClipped_", 18: "highClipped_", 20: "alignmentScore_",
             22: "templateAlignmentScore_", 24: "mateFStrandPosition_", 32: "readLength_", 34: "cigarLength_", 36: "gapCount_",
             38: "editDistance_", 40: "flags_", 48: "tile_", 56: "barcode_", 64: "barcodeSequence_", 72: "clusterId_", 80: "clusterX_",
             84: "clusterY_", 88: "duplicateClusterRank_", 96: "mateAnchor_", 104: "mateStorageBin_"}
    for r, L in enumerate(read_lengths):
        begin = want.read_offset[r]
        end = want.read_offset[r + 1] if r + 1 < len(read_lengths) else want.record_length
        g, w = got.records[:, begin:end], want.records[:, begin:end]
        bad = np.argwhere((g[:, :H] & mask) != (w[:, :H] & mask))
        if bad.size:
            c, b = bad[0]
            member = names[max(k for k in names if k <= b)]
            raise AssertionError("%s: header of cluster %d read %d differs at byte %d (%s): %r vs %r" % (
                what, c, r, b, member, g[c, :H].tobytes().hex(), w[c, :H].tobytes().hex()))
        assert not (g[:, :H] & ~mask).any(), what + ": padding bytes of the header are zero"
        bad = np.argwhere(g[:, H:] != w[:, H:])
        assert not bad.size, "%s: data bytes of read %d differ at %d places, first cluster %d byte %d" % (
            what, r, len(bad), bad[0][0] if bad.size else -1, bad[0][1] if bad.size else -1)


# ---- CPU: the warp functions lane after lane ------------------------------------------------------------------------

@pytest.fixture(scope="module")
def lanes_lib():
    so = os.path.join(ROOT, "build", "libtest_pack_fragments.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", "-shared", "-fPIC",
                           os.path.join(ROOT, "tests", "cpp", "test_pack_fragments.cpp"), "-o", so])
    return ctypes.CDLL(so)


def pack_lanes(lib, reads, templates, options, lanes, misalign):
    from isaac_aligner_b200.batch import PackedFragments
    n, rc = reads.cluster_count, reads.read_count
    layout = np.zeros(4, dtype=np.uint32)
    rec = np.zeros(n * 8192, dtype=np.uint8)
    pos = np.zeros((n, rc), dtype=np.uint64)
    init = np.full((n, rc), 9, dtype=np.uint8)
    stored = ctypes.c_uint64()
    offsets = np.zeros(n * rc + 1, dtype=np.uint64)
    t, f, cig = templates.templates, templates.fragments, templates.cigars
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    assert lib.pack_fragments_lanes(ctypes.byref(reads.c), p(t), p(f), p(cig), ctypes.byref(options.c), ctypes.c_uint(lanes),
                                    ctypes.c_uint(misalign), p(rec), p(pos), p(init), p(layout), ctypes.byref(stored), p(offsets)) == 0
    record_length = int(layout[0])
    if options.c.compact:
        return PackedFragments(rec[:int(offsets[-1])].copy(), pos, init, record_length, (int(layout[1]), int(layout[2])), int(layout[3]),
                               int(stored.value), offsets)
    return PackedFragments(rec[:n * record_length].reshape(n, record_length).copy(), pos, init, record_length,
                           (int(layout[1]), int(layout[2])), int(layout[3]), int(stored.value))


@needs_reference
@pytest.mark.parametrize("read_lengths", [(150, 150), (100, 75), (36,), (251, 33), (151,)])
def test_record_layout_is_the_fragment_buffers(lanes_lib, read_lengths):
    """FragmentBuffer::getRecordLength / getReadOffsets, sizeof(io::FragmentHeader)"""
    rng = np.random.default_rng(1)
    reads = ReadSet(random_bcl(rng, 2, sum(read_lengths)), read_lengths)
    templates = random_templates(rng, 2, read_lengths)
    options, _ = make_options(rng, 2, True, with_arrays=False)
    want, _ = oracle_lib.pack_fragments(oracle_lib.reference(), reads, templates, options, records=False)
    got = pack_lanes(lanes_lib, reads, templates, options, 32, 0)
    assert (got.record_length, got.read_offset, got.header_length) == (want.record_length, want.read_offset, want.header_length)
    assert want.header_length == 112


@needs_reference
@pytest.mark.parametrize("read_lengths,lanes,misalign,keep,arrays", [
    ((150, 150), 32, 0, True, True), ((150, 150), 32, 0, False, True), ((150, 150), 1, 3, True, False),
    ((101, 76), 32, 1, True, True), ((101, 76), 5, 2, False, True), ((36, 250), 32, 3, True, True),
    ((151,), 32, 0, True, True), ((75,), 7, 1, False, False), ((33, 32), 32, 2, True, True),
])
def test_warp_functions_against_the_reference_on_the_cpu(lanes_lib, read_lengths, lanes, misalign, keep, arrays):
    rng = np.random.default_rng(hash((read_lengths, lanes, misalign, keep)) & 0xFFFFFFFF)
    n = 600
    reads = ReadSet(random_bcl(rng, n, sum(read_lengths)), read_lengths)
    templates = random_templates(rng, n, read_lengths)
    options, barcode = make_options(rng, n, keep, with_arrays=arrays)
    want, mask = oracle_lib.pack_fragments(oracle_lib.reference(), reads, templates, options, barcode_bytes=barcode)
    got = pack_lanes(lanes_lib, reads, templates, options, lanes, misalign)
    assert_packed_equal(got, want, mask, read_lengths, "lanes %d, buffer at %d mod 8" % (lanes, misalign))
    if not keep:
        assert 0 < want.initialized.sum() < want.initialized.size


def assert_compact_equal(got, want, mask, what):
    """got: a compact result; want: the reference's FragmentBuffer, cut to the records' total lengths"""
    data, offsets = want.compacted()
    assert np.array_equal(got.record_offset, offsets), what + ": record offsets"
    assert got.records.size == data.size, what
    assert np.array_equal(got.initialized, want.initialized) and np.array_equal(got.f_strand_pos, want.f_strand_pos), what
    keep = np.ones(data.size, dtype=bool)                                # all bytes but the padding of the headers
    H = want.header_length
    pad = np.nonzero(mask != 0xFF)[0]
    starts = offsets[:-1][np.diff(offsets.astype(np.int64)) > 0].astype(np.int64)
    for b in pad:
        if mask[b] == 0:
            keep[starts + b] = False
    g, w = got.records.copy(), data.copy()
    for b in pad:
        if mask[b] != 0:
            g[starts + b] &= mask[b]; w[starts + b] &= mask[b]
    bad = np.nonzero((g != w) & keep)[0]
    assert not bad.size, "%s: %d bytes differ, first at %d" % (what, bad.size, bad[0] if bad.size else -1)
    assert H == 112


@needs_reference
@pytest.mark.parametrize("read_lengths,lanes,misalign,keep", [((150, 150), 32, 0, True), ((101, 76), 32, 3, False), ((151,), 3, 1, False),
                                                              ((33, 32), 32, 2, True)])
def test_compact_records_against_the_reference_on_the_cpu(lanes_lib, read_lengths, lanes, misalign, keep):
    """options.compact: the bytes BufferingFragmentStorage::flush writes per fragment, back to back"""
    rng = np.random.default_rng(hash((read_lengths, lanes, misalign, keep, "compact")) & 0xFFFFFFFF)
    n = 500
    reads = ReadSet(random_bcl(rng, n, sum(read_lengths)), read_lengths)
    templates = random_templates(rng, n, read_lengths)
    options, barcode = make_options(rng, n, keep, compact=True)
    want, mask = oracle_lib.pack_fragments(oracle_lib.reference(), reads, templates, options, barcode_bytes=barcode)
    got = pack_lanes(lanes_lib, reads, templates, options, lanes, misalign)
    assert_compact_equal(got, want, mask, "compact, lanes %d" % lanes)


def golden_cases():
    import base64
    import json
    from isaac_aligner_b200.batch import PackedFragments
    data = json.load(open(os.path.join(ROOT, "tests", "golden", "pack_fragments.json")))
    for case in data["cases"]:
        dec = lambda key, dtype: None if case[key] is None else np.frombuffer(base64.b64decode(case[key]), dtype=dtype).copy()
        rl, n = tuple(case["readLengths"]), case["clusters"]
        reads = ReadSet(dec("bcl", np.uint8), rl)
        templates = Templates(dec("templates", TEMPLATE_DTYPE), dec("fragments", FRAGMENT_DTYPE), dec("cigars", np.uint32))
        options = PackOptions(tile=case["tile"], barcode_idx=case["barcodeIdx"], keep_unaligned=case["keepUnaligned"], pf=dec("pf", np.uint8),
                              xy=dec("xy", np.int32), barcode_sequence=dec("barcodeSequence", np.uint64),
                              distribution_bin_size=case["distributionBinSize"],
                              bin_index=[np.array(b, dtype=np.uint32) for b in case["binIndex"]] if case["binIndex"] else None)
        want = PackedFragments(dec("records", np.uint8).reshape(n, case["recordLength"]), dec("fStrandPos", np.uint64).reshape(n, len(rl)),
                               dec("initialized", np.uint8).reshape(n, len(rl)), case["recordLength"], tuple(case["readOffset"]),
                               case["headerLength"], 0)
        yield reads, templates, options, want, dec("headerMask", np.uint8), rl


def test_warp_functions_against_the_golden_vectors(lanes_lib):
    """tests/golden/pack_fragments.json (made by make_pack_goldens.py from the reference's io::FragmentHeader): no reference build needed"""
    count = 0
    for reads, templates, options, want, mask, rl in golden_cases():
        got = pack_lanes(lanes_lib, reads, templates, options, 32, count % 4)
        assert_packed_equal(got, want, mask, rl, "golden case %d" % count)
        count += 1
    assert count == 5


@pytest.mark.gpu
def test_pack_fragments_golden_vectors(capi):
    count = 0
    for reads, templates, options, want, mask, rl in golden_cases():
        ctx = gpu_context(capi, reads)
        assert_packed_equal(ctx.pack_fragments(templates, options), want, mask, rl, "GPU, golden case %d" % count)
        ctx.close()
        count += 1
    assert count == 5


def assert_round_trip(packed, reads, templates, options):
    """size-independent properties of a FragmentBuffer-layout result, no checker involved: every stored record decodes back to the
    template it was made from (positions, flags, scores, lengths), its bases are the read's BCL bytes (reverse-complemented back
    for reverse fragments), its CIGAR words the fragment's, and everything behind them is zero"""
    from isaac_aligner_b200.batch import NO_MATCH_POSITION, reference_position
    n, rc = reads.cluster_count, reads.read_count
    t, f = templates.templates, templates.fragments.reshape(n, rc)
    stored = (t["built"] != 0) | bool(options.c.keepUnaligned)
    assert np.array_equal(packed.initialized != 0, np.repeat(stored[:, None], rc, axis=1))
    assert packed.stored == int(stored.sum()) * rc
    offsets = np.concatenate([[0], np.cumsum(reads.read_lengths)])
    for r in range(rc):
        h, fr, L = packed.headers(r)[stored], f[stored, r], reads.read_lengths[r]
        aligned = fr["cigarLength"] != 0
        assert np.array_equal(h["readLength"], np.full(len(h), L)) and np.array_equal(h["cigarLength"], fr["cigarLength"])
        assert np.array_equal(h["observedLength"], np.where(aligned, fr["observedLength"], 0))
        assert np.array_equal(h["clusterId"], np.nonzero(stored)[0]) and (h["tile"] == options.c.tile).all() and (h["barcode"] == options.c.barcodeIdx).all()
        assert np.array_equal(h["alignmentScore"], (t["fragmentAlignmentScore"][stored, r] & 0xFFFF).astype(np.uint16))
        assert np.array_equal((h["flags"] >> 1) & 1, (~aligned).astype(np.uint16)) and np.array_equal((h["flags"] >> 3) & 1, fr["reverse"].astype(np.uint16))
        assert np.array_equal((h["flags"] >> 8) & 1, t["properPair"][stored].astype(np.uint16)) if rc == 2 else ((h["flags"] >> 8) & 1 == 0).all()
        assert np.array_equal((h["flags"] >> (5 + r)) & 1, np.ones(len(h), np.uint16)) and np.array_equal(h["flags"] & 1, np.full(len(h), rc - 1, np.uint16))
        own = aligned | (rc == 1)                                        # a shadow carries its mate's position (Fragment.hh:111-113)
        contig, position = reference_position(h["fStrandPosition"])
        placed = own & (fr["contigId"] != 0x7FFFFF)
        assert np.array_equal(contig[placed], fr["contigId"][placed].astype(np.int64)) and np.array_equal(position[placed], fr["position"][placed])
        assert (h["fStrandPosition"][own & (fr["contigId"] == 0x7FFFFF)] == NO_MATCH_POSITION).all()
        assert np.array_equal(packed.f_strand_pos[stored, r] == NO_MATCH_POSITION, fr["contigId"] == 0x7FFFFF)
        # bases: forward fragments carry the BCL bytes, reverse ones their reverse complement (N stays 0)
        begin = packed.read_offset[r] + packed.header_length
        bases = packed.records[stored, begin:begin + L]
        bcl = reads.bcl[stored, offsets[r]:offsets[r + 1]]
        is_n = (bcl & 0xFC) == 0
        back = np.where(is_n, 0, (bcl & 0xFC) | (3 - (bcl & 3)))[:, ::-1]
        want = np.where((fr["reverse"] != 0)[:, None], back, bcl)
        assert np.array_equal(bases, want)
        # CIGAR words and the zero fill
        end = packed.read_offset[r + 1] if r + 1 < rc else packed.record_length
        tail = packed.records[stored, begin + L:end]
        words = np.ascontiguousarray(tail[:, :(tail.shape[1] // 4) * 4]).view(np.uint32)
        k = np.arange(words.shape[1])[None, :]
        live = k < fr["cigarLength"][:, None]
        index = np.minimum(fr["cigarOffset"][:, None].astype(np.int64) + k, templates.cigars.size - 1)
        assert np.array_equal(np.where(live, words, 0), np.where(live, templates.cigars[index], 0)) and not words[~live].any()
        assert not tail[:, words.shape[1] * 4:].any()
    assert not packed.records[~stored].any()


def test_round_trip_properties_on_the_cpu(lanes_lib):
    rng = np.random.default_rng(77)
    for read_lengths, keep in (((150, 150), True), ((101, 76), False), ((151,), True)):
        n = 3000
        reads = ReadSet(random_bcl(rng, n, sum(read_lengths)), read_lengths)
        templates = random_templates(rng, n, read_lengths)
        options, _ = make_options(rng, n, keep, with_arrays=False)
        assert_round_trip(pack_lanes(lanes_lib, reads, templates, options, 32, 0), reads, templates, options)


def test_pack_kernel_compiles_for_the_device():
    """the kernel is part of libisaac_ext.so; its resources as ptxas reports them (no spills, no stack frame beyond the header)"""
    out = subprocess.run(["cuobjdump", "-res-usage", os.path.join(ROOT, "isaac_aligner_b200", "libisaac_ext.so")],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.splitlines()
    hit = [i for i, l in enumerate(lines) if "packFragmentsKernel" in l]
    assert hit, "packFragmentsKernel is missing from libisaac_ext.so"


# ---- GPU: the kernel through the C ABI --------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def capi():
    from isaac_aligner_b200 import capi
    return capi


def gpu_context(capi, reads, genome=None):
    from isaac_aligner_b200.types import Config
    ctx = capi.Context(Config.default(max_read_length=2 * max(sum(reads.read_lengths), 100)))
    if genome is not None:
        ctx.set_reference(genome)
    ctx.set_reads(reads)
    return ctx


@pytest.mark.gpu
@needs_reference
@pytest.mark.parametrize("read_lengths,keep,arrays", [((150, 150), True, True), ((150, 150), False, False), ((101, 76), True, True),
                                                      ((36, 250), False, True), ((151,), True, True), ((33, 32), True, False)])
def test_pack_fragments_random_templates(capi, read_lengths, keep, arrays):
    rng = np.random.default_rng(hash((read_lengths, keep, arrays)) & 0xFFFFFFFF)
    n = 5000
    reads = ReadSet(random_bcl(rng, n, sum(read_lengths)), read_lengths)
    templates = random_templates(rng, n, read_lengths)
    options, barcode = make_options(rng, n, keep, with_arrays=arrays)
    want, mask = oracle_lib.pack_fragments(oracle_lib.reference(), reads, templates, options, barcode_bytes=barcode)
    ctx = gpu_context(capi, reads)
    launches = ctx.launches
    got = ctx.pack_fragments(templates, options)
    assert ctx.launches == launches + 1
    assert_packed_equal(got, want, mask, read_lengths, "GPU, reads %r" % (read_lengths,))
    again = ctx.pack_fragments(templates, options)                       # buffers of the context reused
    assert np.array_equal(again.records, got.records)
    ctx.close()


@pytest.mark.gpu
@needs_reference
@pytest.mark.parametrize("read_lengths,keep", [((150, 150), True), ((101, 76), False), ((151,), False)])
def test_pack_fragments_compact(capi, read_lengths, keep):
    rng = np.random.default_rng(hash((read_lengths, keep, "gpu compact")) & 0xFFFFFFFF)
    n = 4000
    reads = ReadSet(random_bcl(rng, n, sum(read_lengths)), read_lengths)
    templates = random_templates(rng, n, read_lengths)
    options, barcode = make_options(rng, n, keep, compact=True)
    want, mask = oracle_lib.pack_fragments(oracle_lib.reference(), reads, templates, options, barcode_bytes=barcode)
    ctx = gpu_context(capi, reads)
    got = ctx.pack_fragments(templates, options)
    assert_compact_equal(got, want, mask, "GPU compact, reads %r" % (read_lengths,))
    ctx.close()


@pytest.mark.gpu
@needs_reference
def test_pack_fragments_of_a_simulated_tile(capi):
    """matches -> isaac_ext_build_templates (with the end clippers) -> isaac_ext_pack_fragments, against io::FragmentHeader fed with
    the same templates"""
    from common_build import build_workload
    from isaac_aligner_b200.batch import Tls, TemplateOptions
    from isaac_aligner_b200.types import Config
    genome, sim, reads, mb = build_workload(n_pairs=3000, L=100, seed=23)
    ctx = capi.Context(Config.default(max_read_length=200))
    ctx.set_reference(genome)
    ctx.set_reads(reads)
    templates = ctx.build_templates(mb, Tls.make(), TemplateOptions.make(clip_semialigned=True, clip_overlapping=True))
    rng = np.random.default_rng(5)
    lengths = [len(c) for c in genome]
    bins = [(1 + 3 * i + np.arange(l // 1000 + 1) // 50).astype(np.uint32) for i, l in enumerate(lengths)]
    for keep in (False, True):
        options = PackOptions(tile=2203, barcode_idx=1, keep_unaligned=keep, pf=(rng.random(reads.cluster_count) < 0.9).astype(np.uint8),
                              distribution_bin_size=1000, bin_index=bins)
        want, mask = oracle_lib.pack_fragments(oracle_lib.reference(), reads, templates, options)
        got = ctx.pack_fragments(templates, options)
        assert_packed_equal(got, want, mask, reads.read_lengths, "simulated tile, keepUnaligned %d" % keep)
        assert got.stored >= 2 * int(templates.templates["built"].sum())
    ctx.close()


@pytest.mark.gpu
def test_pack_fragments_round_trip_at_tile_size(capi):
    """200 000 clusters (a 329 MB record buffer, larger than the 126 MB L2): the size-independent properties, no checker"""
    rng = np.random.default_rng(78)
    n, read_lengths = 200_000, (150, 150)
    reads = ReadSet(random_bcl(rng, n, sum(read_lengths)), read_lengths)
    small = random_templates(rng, 2000, read_lengths)                    # 2000 random templates tiled over the clusters
    reps = n // 2000
    fragments = np.tile(small.fragments, reps)
    fragments["readId"] = np.arange(2 * n, dtype=np.uint32)
    templates = Templates(np.tile(small.templates, reps), fragments, small.cigars)
    options = PackOptions(tile=1101, barcode_idx=2, keep_unaligned=False)
    ctx = gpu_context(capi, reads)
    assert_round_trip(ctx.pack_fragments(templates, options), reads, templates, options)
    ctx.close()


@pytest.mark.gpu
def test_pack_fragments_argument_errors(capi):
    rng = np.random.default_rng(2)
    reads = ReadSet(random_bcl(rng, 10, 100), (50, 50))
    templates = random_templates(rng, 10, (50, 50))
    ctx = gpu_context(capi, reads)
    bad = Templates(templates.templates, templates.fragments.copy(), templates.cigars[:1].copy())
    bad.fragments["cigarLength"][0], bad.fragments["cigarOffset"][0] = 3, 0
    with pytest.raises(capi.ExtError):
        ctx.pack_fragments(bad, PackOptions())
    ctx.close()

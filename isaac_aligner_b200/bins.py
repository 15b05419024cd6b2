"""Bins of io::FragmentHeader records for build::GapRealigner: the record layout of the reference's bin files
(io/Fragment.hh:73-404), a synthetic bin whose reads come from a haplotype with shared insertions and deletions -- so that the
gaps some reads carry in their CIGARs can repair the mismatching ends of others, which is what the realigner is for -- and the
flat view of PackedFragmentBuffer::Index (build/PackedFragmentBuffer.hh:36-91).  Harness-side code: the product is
isaac_ext_realign_bin."""
import ctypes

import numpy as np

HEADER_DTYPE = np.dtype({
    "names": ["bamTlen", "observedLength", "fStrandPosition", "lowClipped", "highClipped", "alignmentScore", "templateAlignmentScore",
              "mateFStrandPosition", "readLength", "cigarLength", "gapCount", "editDistance", "flags", "tile", "barcode",
              "barcodeSequence", "clusterId", "clusterX", "clusterY", "duplicateClusterRank", "mateAnchor", "mateStorageBin"],
    "formats": ["<i4", "<u4", "<u8", "<u2", "<u2", "<u2", "<u2", "<u8", "<u2", "<u2", "<u2", "<u2", "<u2", "<u8", "<u8", "<u8", "<u8",
                "<i4", "<i4", "<u8", "<u8", "<u4"],
    "offsets": [0, 4, 8, 16, 18, 20, 22, 24, 32, 34, 36, 38, 40, 48, 56, 64, 72, 80, 84, 88, 96, 104],
    "itemsize": 112})
BIN_INDEX_DTYPE = np.dtype([("dataOffset", "<u8"), ("mateDataOffset", "<u8")])
GAP_DTYPE = np.dtype([("position", "<u8"), ("length", "<i4"), ("group", "<u4")])
FLAG_PAIRED, FLAG_UNMAPPED, FLAG_MATE_UNMAPPED, FLAG_REVERSE, FLAG_MATE_REVERSE = 1, 2, 4, 8, 16
FLAG_FIRST_READ, FLAG_SECOND_READ, FLAG_FAIL_FILTER, FLAG_PROPER_PAIR = 32, 64, 128, 256
OP_ALIGN, OP_INSERT, OP_DELETE, OP_SOFT_CLIP = 0, 1, 2, 4
OWN_CIGAR = 0xFFFFFFFF
DODGY_ALIGNMENT_SCORE = 0xFFFF


def reference_position(contig, position):
    """ReferencePosition(contigId, position).getValue() (ReferencePosition.hh:68-78)"""
    return (((int(contig) + 1) << 40) | int(position)) << 1


def position_of(value):
    value = int(value) >> 1
    return (value >> 40) - 1, value & ((1 << 40) - 1)


class RealignOptionsC(ctypes.Structure):
    """isaac_ext_realign_options_t"""
    _fields_ = [("binStart", ctypes.c_uint64), ("binEnd", ctypes.c_uint64), ("realignGapsVigorously", ctypes.c_uint32),
                ("realignDodgyFragments", ctypes.c_uint32), ("mismatchCost", ctypes.c_uint32), ("gapOpenCost", ctypes.c_uint32),
                ("gapExtendCost", ctypes.c_uint32), ("clipSemialigned", ctypes.c_uint32), ("barcodeCount", ctypes.c_uint32),
                ("pad", ctypes.c_uint32), ("barcodeTls", ctypes.c_void_p), ("barcodeGapGroup", ctypes.c_void_p)]


class RealignResultC(ctypes.Structure):
    """isaac_ext_realign_result_t"""
    _fields_ = [("position", ctypes.c_void_p), ("cigarOffset", ctypes.c_void_p), ("cigarLength", ctypes.c_void_p),
                ("realignedCigars", ctypes.c_void_p), ("realignedCigarWords", ctypes.c_uint64), ("realignedFragments", ctypes.c_uint64),
                ("gaps", ctypes.c_void_p), ("deletionsByEnd", ctypes.c_void_p), ("gapCount", ctypes.c_uint64),
                ("deletionCount", ctypes.c_uint64), ("collectMs", ctypes.c_float), ("realignMs", ctypes.c_float)]


class RealignJobC(ctypes.Structure):
    """isaac_ext_realign_job_t"""
    _fields_ = [("options", ctypes.c_void_p), ("data", ctypes.c_void_p), ("dataBytes", ctypes.c_uint64), ("recordOffset", ctypes.c_void_p),
                ("recordCount", ctypes.c_uint64), ("index", ctypes.c_void_p), ("indexCount", ctypes.c_uint64), ("position", ctypes.c_void_p),
                ("cigarOffset", ctypes.c_void_p), ("cigarLength", ctypes.c_void_p), ("realignedCigars", ctypes.c_void_p),
                ("realignedCigarCapacity", ctypes.c_uint64), ("realignedCigarWords", ctypes.c_uint64), ("realignedFragments", ctypes.c_uint64),
                ("status", ctypes.c_int32), ("pad", ctypes.c_uint32)]


class RealignOptions:
    """keeps the arrays the C struct points to alive"""

    def __init__(self, bin_start, bin_end, tls_list, vigorous=False, dodgy=False, clip_semialigned=False, gap_groups=None,
                 mismatch_cost=3, gap_open_cost=4, gap_extend_cost=0):
        from .batch import Tls
        self.tls = (Tls * len(tls_list))(*tls_list)
        self.groups = np.ascontiguousarray(gap_groups, dtype=np.uint32) if gap_groups is not None else None
        self.c = RealignOptionsC(int(bin_start), int(bin_end), int(vigorous), int(dodgy), mismatch_cost, gap_open_cost, gap_extend_cost,
                                 int(clip_semialigned), len(tls_list), 0, ctypes.addressof(self.tls),
                                 self.groups.ctypes.data if self.groups is not None else None)


class RealignResult:
    def __init__(self, data, position, cigar_offset, cigar_length, cigars, gaps, deletions, realigned=None, collect_ms=0.0, realign_ms=0.0):
        self.data, self.position, self.cigar_offset, self.cigar_length, self.cigars = data, position, cigar_offset, cigar_length, cigars
        self.gaps, self.deletions, self.realigned, self.collect_ms, self.realign_ms = gaps, deletions, realigned, collect_ms, realign_ms

    def cigar(self, i, bin_):
        """the CIGAR words of index entry i after the pass"""
        n = int(self.cigar_length[i])
        if self.cigar_offset[i] == OWN_CIGAR:
            return bin_.record_cigar(int(bin_.index["dataOffset"][i]), self.data)[:n]
        o = int(self.cigar_offset[i])
        return self.cigars[o:o + n]


class Bin:
    def __init__(self, data, record_offset, index, bin_start, bin_end):
        self.data, self.record_offset, self.index, self.bin_start, self.bin_end = data, record_offset, index, bin_start, bin_end

    def header(self, offset, data=None):
        data = self.data if data is None else data
        return np.frombuffer(data[offset:offset + 112].tobytes(), dtype=HEADER_DTYPE)[0]

    def record_cigar(self, offset, data=None):
        data = self.data if data is None else data
        h = self.header(offset, data)
        begin = offset + 112 + int(h["readLength"])
        return np.frombuffer(data[begin:begin + 4 * int(h["cigarLength"])].tobytes(), dtype=np.uint32)


def cigar_string(words):
    return "".join("%d%s" % (int(w) >> 4, "MID?S"[int(w) & 15] if (int(w) & 15) < 5 else "?") for w in words)


_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _a in enumerate(b"ACGT"):
    _CODE[_a] = _i


def _tlen(f_begin, f_end, m_begin, m_end, first_read):
    """io::FragmentHeader::getTlen (Fragment.hh:209-226) on plain positions of one contig"""
    distance = max(f_end, m_end) - min(f_begin, m_begin)
    if f_begin < m_begin:
        return distance
    return -distance if (f_begin > m_begin or not first_read) else distance


def simulate_bin(contigs, contig=0, region=(2000, 42000), n_pairs=2000, read_length=100, seed=1, variant_spacing=220, error_rate=0.004,
                 gapped_fraction=0.6, clip_fraction=0.08, singleton_fraction=0.03, single_ended_fraction=0.03, dodgy_fraction=0.03,
                 duplicate_fraction=0.03, n_fraction=0.002, barcodes=1, template_mean=330, template_sd=35, max_indel=14,
                 edge_fraction=0.0):
    """One bin of a run: pairs sampled from a haplotype of contigs[contig] (ASCII uint8 arrays) that differs from the reference by
    SNPs, insertions and deletions every ~variant_spacing bases.  An aligner's view of every read is emulated: a read that crosses an
    indel well inside keeps the true gapped CIGAR with probability gapped_fraction, otherwise it is laid down without gaps, anchored
    on its longer gap-free side (the other side then mismatches until a realigner brings the gap in).  Returns Bin."""
    rng = np.random.default_rng(seed)
    ref = np.asarray(contigs[contig], dtype=np.uint8)
    a, b = region
    L = read_length
    # ---- the haplotype over [a - 600, b + 1200): bases + for every base its reference coordinate and whether it is inserted
    lo, hi = max(0, a - 600), min(len(ref), b + 1200)
    hap, coord, inserted = [], [], []
    p = lo
    next_variant = lo + int(rng.integers(10, max(11, variant_spacing)))
    while p < hi:
        if p >= next_variant:
            kind = rng.integers(0, 4)
            n = int(rng.integers(1, max_indel + 1))
            if kind == 0 and ref[p] != ord("N"):                                         # SNP
                hap.append(b"ACGT"[(int(_CODE[ref[p]]) + int(rng.integers(1, 4))) % 4]); coord.append(p); inserted.append(False); p += 1
            elif kind in (1, 3):                                                          # insertion in front of p
                for _ in range(n):
                    hap.append(b"ACGT"[int(rng.integers(0, 4))]); coord.append(p); inserted.append(True)
            else:                                                                         # deletion of [p, p + n)
                p += n
            next_variant = p + int(rng.integers(min(30, variant_spacing), 2 * variant_spacing))
            continue
        hap.append(int(ref[p])); coord.append(p); inserted.append(False); p += 1
    hap = np.array(hap, dtype=np.uint8); coord = np.array(coord, dtype=np.int64); inserted = np.array(inserted, dtype=bool)

    def true_alignment(s, e):
        """CIGAR operations [(len, op)] and reference position of hap[s:e); inserted bases at either end become soft clips"""
        ops = []
        k = s
        while k < e and inserted[k]:
            k += 1
        lead = k - s
        t = e
        while t > k and inserted[t - 1]:
            t -= 1
        trail = e - t
        if k >= t:
            return None
        pos = int(coord[k])
        run = 0
        prev = None
        i = k
        while i < t:
            if inserted[i]:
                if run:
                    ops.append((run, OP_ALIGN)); run = 0
                j = i
                while j < t and inserted[j]:
                    j += 1
                ops.append((j - i, OP_INSERT)); i = j
                continue
            if prev is not None and coord[i] != prev + 1:
                if run:
                    ops.append((run, OP_ALIGN)); run = 0
                ops.append((int(coord[i] - prev - 1), OP_DELETE))
            run += 1; prev = int(coord[i]); i += 1
        if run:
            ops.append((run, OP_ALIGN))
        merged = []
        for n_, op in ops:                                                                # I directly followed by D etc. stay as they are
            merged.append((n_, op))
        return pos, lead, merged, trail

    def lay_down(bases, s, e):
        """(position, cigar ops, editDistance, gapCount, observedLength) the way the emulated aligner reports hap[s:e) = bases"""
        t = true_alignment(s, e)
        if t is None:
            return None
        pos, lead, ops, trail = t
        low_clip = int(rng.integers(1, 9)) if rng.random() < clip_fraction else 0
        high_clip = int(rng.integers(1, 9)) if rng.random() < clip_fraction else 0
        gaps = [op for _, op in ops if op != OP_ALIGN]
        interior = bool(gaps) and ops[0][0] >= 12 + low_clip and ops[-1][0] >= 12 + high_clip
        if gaps and interior and not lead and not trail and rng.random() < gapped_fraction:
            # true gapped CIGAR; alignment-independent clips eat into the first / last match run
            cig = list(ops)
            start = pos
            if low_clip:
                cig[0] = (cig[0][0] - low_clip, OP_ALIGN); cig.insert(0, (low_clip, OP_SOFT_CLIP)); start += low_clip
            if high_clip:
                cig[-1] = (cig[-1][0] - high_clip, OP_ALIGN); cig.append((high_clip, OP_SOFT_CLIP))
        else:
            # ungapped, anchored on the side with the longer gap-free run
            left_run = ops[0][0] if not lead else 0
            right_run = ops[-1][0] if not trail else 0
            if left_run >= right_run:
                start0 = pos - lead                                                       # reference coordinate of read base 0
            else:
                last = e - 1 - trail
                start0 = int(coord[last]) + trail - (len(bases) - 1)
            if start0 < 0 or start0 + len(bases) > len(ref):
                return None
            start = start0 + low_clip
            cig = ([(low_clip, OP_SOFT_CLIP)] if low_clip else []) + [(len(bases) - low_clip - high_clip, OP_ALIGN)] + \
                  ([(high_clip, OP_SOFT_CLIP)] if high_clip else [])
        # edit distance of that CIGAR against the reference (chars differ; inserted and deleted bases count)
        ed, r, q = 0, start, 0
        for n_, op in cig:
            if op == OP_SOFT_CLIP:
                q += n_
            elif op == OP_ALIGN:
                if r < 0 or r + n_ > len(ref):
                    return None
                ed += int(np.count_nonzero(ref[r:r + n_] != bases[q:q + n_])); r += n_; q += n_
            elif op == OP_INSERT:
                ed += n_; q += n_
            else:
                ed += n_; r += n_
        return start, cig, ed, sum(1 for _, op in cig if op in (OP_INSERT, OP_DELETE)), r - start, low_clip, high_clip

    records = []                     # dicts
    hap_lo = int(np.searchsorted(coord, a))
    hap_hi = int(np.searchsorted(coord, b))
    for pair in range(n_pairs):
        tl = max(L + 10, int(rng.normal(template_mean, template_sd)))
        if edge_fraction and rng.random() < edge_fraction:
            s = int(rng.integers(0, 40))
        else:
            s = int(rng.integers(max(0, hap_lo - tl // 2), hap_hi))
        e = s + tl
        if e > len(hap):
            continue
        single_ended = rng.random() < single_ended_fraction
        reads = []
        for read_index, (rs, re_, reverse) in enumerate(((s, s + L, False), (e - L, e, True))):
            bases = hap[rs:re_].copy()
            err = rng.random(L) < error_rate
            for k in np.flatnonzero(err):
                bases[k] = b"ACGT"[(int(_CODE[bases[k]]) + int(rng.integers(1, 4))) % 4] if _CODE[bases[k]] < 4 else bases[k]
            qual = rng.integers(20, 41, size=L).astype(np.uint8)
            isn = (rng.random(L) < n_fraction) | (_CODE[bases] > 3)
            bases_n = bases.copy(); bases_n[isn] = ord("N")
            laid = lay_down(bases_n, rs, re_)
            reads.append((read_index, reverse, bases_n, qual, isn, laid))
            if single_ended:
                break
        if any(r[5] is None for r in reads):
            continue
        singleton = (not single_ended) and rng.random() < singleton_fraction
        dodgy = rng.random() < dodgy_fraction
        barcode = int(rng.integers(0, barcodes))
        made = []
        for read_index, reverse, bases_n, qual, isn, laid in reads:
            start, cig, ed, gap_count, observed, low_clip, high_clip = laid
            bcl = np.where(isn, 0, (qual << 2) | np.where(isn, 0, _CODE[bases_n] & 3)).astype(np.uint8)
            made.append(dict(read_index=read_index, reverse=reverse, bcl=bcl, start=start, cigar=cig, ed=ed, gaps=gap_count, observed=observed,
                             low=low_clip if not reverse else high_clip, high=high_clip if not reverse else low_clip,
                             pair=pair, barcode=barcode, dodgy=dodgy, unmapped=False, single=single_ended))
        if singleton and len(made) == 2:
            shadow = made[int(rng.integers(0, 2))]
            shadow.update(unmapped=True, cigar=[], ed=0, gaps=0, observed=0, low=0, high=0)
        records.append(made)

    # ---- serialise: the fragments whose position lies in [a, b) belong to this bin (shadows sit at their singleton's position)
    blobs, offsets, entries = [], [], []     # entries: (dataOffset, mateDataOffset, kind) kind 0 se, 1 r-strand / shadow, 2 f-strand
    at = 0
    record_count = 0
    for made in records:
        positions = []
        for m in made:
            mate = made[1 - made.index(m)] if len(made) == 2 else None
            positions.append(mate["start"] if (m["unmapped"] and mate is not None) else m["start"])
        in_bin = [a <= p_ < b for p_ in positions]
        placed = []
        for k, m in enumerate(made):
            if not in_bin[k]:
                placed.append(None)
                continue
            mate = made[1 - k] if len(made) == 2 else None
            h = np.zeros(1, dtype=HEADER_DTYPE)[0]
            flags = 0
            if mate is not None:
                flags |= FLAG_PAIRED | (FLAG_FIRST_READ if m["read_index"] == 0 else FLAG_SECOND_READ)
                flags |= (FLAG_MATE_UNMAPPED if mate["unmapped"] else 0) | (FLAG_MATE_REVERSE if mate["reverse"] else 0)
                both = not m["unmapped"] and not mate["unmapped"]
                tlen = _tlen(m["start"], m["start"] + m["observed"], mate["start"], mate["start"] + mate["observed"], m["read_index"] == 0) if both else 0
                mate_pos = mate["start"] if not mate["unmapped"] else m["start"]
                proper = both and rng.random() < 0.9
                flags |= FLAG_PROPER_PAIR if proper else 0
            else:
                flags |= FLAG_MATE_UNMAPPED | FLAG_FIRST_READ | FLAG_SECOND_READ
                tlen = 0
                mate_pos = None
            flags |= (FLAG_UNMAPPED if m["unmapped"] else 0) | (FLAG_REVERSE if m["reverse"] else 0)
            h["bamTlen"] = tlen
            h["observedLength"] = m["observed"]
            h["fStrandPosition"] = reference_position(contig, positions[k])
            h["lowClipped"], h["highClipped"] = m["low"], m["high"]
            score = DODGY_ALIGNMENT_SCORE if m["dodgy"] else int(rng.integers(0, 1500))
            h["alignmentScore"] = score
            h["templateAlignmentScore"] = DODGY_ALIGNMENT_SCORE if m["dodgy"] else int(rng.integers(0, 3000))
            h["mateFStrandPosition"] = reference_position(contig, mate_pos) if mate_pos is not None else (0x7FFFFF << 41)
            h["readLength"] = len(m["bcl"]); h["cigarLength"] = len(m["cigar"]); h["gapCount"] = m["gaps"]; h["editDistance"] = m["ed"]
            h["flags"] = flags; h["tile"] = 1; h["barcode"] = m["barcode"]; h["clusterId"] = m["pair"]
            h["clusterX"] = h["clusterY"] = 0x7FFFFFFF
            words = np.array([(n_ << 4) | op for n_, op in m["cigar"]], dtype=np.uint32)
            blob = h.tobytes() + m["bcl"].tobytes() + words.tobytes()
            placed.append((at, len(blob)))
            blobs.append(blob); offsets.append(at); at += len(blob); record_count += 1
        duplicate = rng.random() < duplicate_fraction                                     # stays in the data, leaves the index
        for k, m in enumerate(made):
            if placed[k] is None or duplicate:
                continue
            own = placed[k][0]
            mate_at = placed[1 - k][0] if (len(made) == 2 and placed[1 - k] is not None) else own
            kind = 0 if len(made) == 1 else (1 if (m["reverse"] or m["unmapped"]) else 2)
            entries.append((own, mate_at, kind))
    data = np.frombuffer(b"".join(blobs), dtype=np.uint8).copy() if blobs else np.zeros(0, dtype=np.uint8)
    entries.sort(key=lambda t: t[2])                                                      # stable: se, then r / shadow, then f, in file order
    index = np.array([(o, m_) for o, m_, _ in entries], dtype=BIN_INDEX_DTYPE)
    return Bin(data, np.array(offsets, dtype=np.uint64), index, reference_position(contig, a), reference_position(contig, b))

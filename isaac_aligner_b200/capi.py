"""ctypes binding of libisaac_ext.so (include/isaac_ext.h).

The library is the product: this module raises at import time when it has not been built and every compute
call raises ExtError when no CUDA device is usable -- there is no CPU path to fall back to.
"""
import ctypes
import os

import numpy as np

from .types import CANDIDATE_DTYPE, FRAGMENT_DTYPE, MASK_WORDS, Config, ReadSet

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ISAAC_EXT_LIB", os.path.join(_HERE, "libisaac_ext.so"))   # override: kernel-variant experiments

if not os.path.exists(LIB_PATH):
    raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(isaac_aligner_b200 has no CPU fallback)" % LIB_PATH)

_lib = ctypes.CDLL(LIB_PATH)
_lib.isaac_ext_last_error.restype = ctypes.c_char_p
_lib.isaac_ext_last_error.argtypes = [ctypes.c_void_p]
_lib.isaac_ext_version.restype = ctypes.c_char_p
_lib.isaac_ext_launch_count.restype = ctypes.c_uint64
_lib.isaac_ext_launch_count.argtypes = [ctypes.c_void_p]
_lib.isaac_ext_destroy.argtypes = [ctypes.c_void_p]
_lib.isaac_ext_destroy.restype = None

# every symbol include/isaac_ext.h declares (tests/test_abi.py checks the header against this list)
EXPORTS = [
    "isaac_ext_create", "isaac_ext_destroy", "isaac_ext_last_error", "isaac_ext_version",
    "isaac_ext_set_reference", "isaac_ext_set_reads", "isaac_ext_banded_sw_batch", "isaac_ext_ungapped_batch",
    "isaac_ext_gapped_batch", "isaac_ext_ungapped_batch_device", "isaac_ext_gapped_batch_device",
    "isaac_ext_launch_count", "isaac_ext_measure_int32_peak", "isaac_ext_build_fragments", "isaac_ext_rescue_shadows",
    "isaac_ext_tile_stats_device", "isaac_ext_ungapped_batch_compact", "isaac_ext_gapped_batch_compact",
    "isaac_ext_build_templates", "isaac_ext_trim_low_quality_ends", "isaac_ext_set_adapters",
    "isaac_ext_determine_template_length", "isaac_ext_extend_batch_compact",
    "isaac_ext_template_stats", "isaac_ext_pack_fragments", "isaac_ext_align_batch_packed",
    "isaac_ext_banded_sw_wide_batch", "isaac_ext_banded_sw_wide_batch_device", "isaac_ext_select_tile", "isaac_ext_tile_packed", "isaac_ext_prefetch_reads", "isaac_ext_prefetch_batch", "isaac_ext_tile_cycle_stats", "isaac_ext_tile_cycle_stats_finalize",
    "isaac_ext_submit_build_fragments", "isaac_ext_submit_rescue_shadows", "isaac_ext_submit_build_templates", "isaac_ext_wait",
    "isaac_ext_realign_bin", "isaac_ext_realign_bins", "isaac_ext_build_templates_deferred", "isaac_ext_fetch_templates",
]


class ExtError(RuntimeError):
    def __init__(self, code, message):
        RuntimeError.__init__(self, "isaac_ext error %d: %s" % (code, message))
        self.code = code


def _p(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


class Context:
    """One isaac_ext_ctx (one per GPU)."""

    def __init__(self, config=None):
        self.config = config if config is not None else Config.default()
        self._h = ctypes.c_void_p()
        rc = _lib.isaac_ext_create(ctypes.byref(self.config), ctypes.byref(self._h))
        if rc:
            raise ExtError(rc, _lib.isaac_ext_last_error(None).decode())
        self._keep = []

    def close(self):
        if self._h:
            _lib.isaac_ext_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise ExtError(rc, _lib.isaac_ext_last_error(self._h).decode())

    @property
    def launches(self):
        return int(_lib.isaac_ext_launch_count(self._h))

    def set_reference(self, contigs):
        contigs = [np.ascontiguousarray(c, dtype=np.uint8) for c in contigs]
        n = len(contigs)
        ptrs = (ctypes.c_void_p * n)(*[c.ctypes.data for c in contigs])
        lens = (ctypes.c_uint64 * n)(*[c.size for c in contigs])
        self._check(_lib.isaac_ext_set_reference(self._h, ctypes.c_uint32(n), ptrs, lens))
        self.contig_lengths = [c.size for c in contigs]

    def set_adapters(self, adapters):
        """matchSelector::SequencingAdapterList for every later call: (sequence, reverse, clipLength) tuples; () clears"""
        from .types import adapter_array
        arr = adapter_array(adapters)
        self._check(_lib.isaac_ext_set_adapters(self._h, ctypes.c_uint32(len(adapters)), arr))

    def prefetch_reads(self, reads):
        """isaac_ext_prefetch_reads: the next tile's reads go up while the calls on the current tile run; set_reads(reads) takes them over"""
        assert isinstance(reads, ReadSet)
        self._check(_lib.isaac_ext_prefetch_reads(self._h, ctypes.byref(reads.c)))
        self._prefetched = reads          # keeps the host buffers alive

    def prefetch_batch(self, match_batch, cluster_count):
        """isaac_ext_prefetch_batch: the next tile's seed matches go up while the current tile is processed"""
        self._check(_lib.isaac_ext_prefetch_batch(self._h, ctypes.byref(match_batch.c), ctypes.c_uint32(cluster_count)))
        self._prefetched_batch = match_batch

    def set_reads(self, reads):
        assert isinstance(reads, ReadSet)
        self._check(_lib.isaac_ext_set_reads(self._h, ctypes.byref(reads.c)))
        self.reads = reads

    def banded_sw(self, queries, dbs, scores, cigar_stride=64):
        qbuf = np.frombuffer(b"".join(queries), dtype=np.uint8)
        dbuf = np.frombuffer(b"".join(dbs), dtype=np.uint8)
        qlen = np.array([len(q) for q in queries], dtype=np.uint32)
        qoff = np.concatenate([[0], np.cumsum(qlen[:-1], dtype=np.uint64)]).astype(np.uint64)
        dlen = np.array([len(d) for d in dbs], dtype=np.uint64)
        doff = np.concatenate([[0], np.cumsum(dlen[:-1], dtype=np.uint64)]).astype(np.uint64)
        return self.banded_sw_flat(qbuf, qoff, qlen, dbuf, doff, scores, cigar_stride)

    def banded_sw_flat(self, qbuf, qoff, qlen, dbuf, doff, scores, cigar_stride=64):
        n = len(qlen)
        cig = np.zeros((n, cigar_stride), dtype=np.uint32)
        ciglen = np.zeros(n, dtype=np.uint32)
        off = np.zeros(n, dtype=np.uint32)
        self._check(_lib.isaac_ext_banded_sw_batch(
            self._h, ctypes.c_uint32(n), _p(qbuf), _p(qoff), _p(qlen), _p(dbuf), _p(doff),
            ctypes.c_int(scores[0]), ctypes.c_int(scores[1]), ctypes.c_int(scores[2]), ctypes.c_int(scores[3]),
            ctypes.c_uint32(cigar_stride), _p(cig), _p(ciglen), _p(off)))
        return cig, ciglen, off

    def banded_sw_wide(self, band, queries, dbs, scores, cigar_stride=64):
        """isaac_ext_banded_sw_wide_batch: the warp-wavefront kernel on a band of `band` lanes; lists of bytes in"""
        qbuf = np.frombuffer(b"".join(queries), dtype=np.uint8)
        dbuf = np.frombuffer(b"".join(dbs), dtype=np.uint8)
        qlen = np.array([len(q) for q in queries], dtype=np.uint32)
        qoff = np.concatenate([[0], np.cumsum(qlen[:-1], dtype=np.uint64)]).astype(np.uint64)
        dlen = np.array([len(d) for d in dbs], dtype=np.uint64)
        doff = np.concatenate([[0], np.cumsum(dlen[:-1], dtype=np.uint64)]).astype(np.uint64)
        n = len(qlen)
        cig, ciglen, off = np.zeros((n, cigar_stride), dtype=np.uint32), np.zeros(n, dtype=np.uint32), np.zeros(n, dtype=np.uint32)
        self._check(_lib.isaac_ext_banded_sw_wide_batch(
            self._h, ctypes.c_uint32(band), ctypes.c_uint32(n), _p(qbuf), _p(qoff), _p(qlen), _p(dbuf), _p(doff), ctypes.c_int(scores[0]),
            ctypes.c_int(scores[1]), ctypes.c_int(scores[2]), ctypes.c_int(scores[3]), ctypes.c_uint32(cigar_stride), _p(cig), _p(ciglen), _p(off)))
        return cig, ciglen, off

    def banded_sw_wide_device(self, band, n, d_q, d_qoff, d_qlen, d_db, d_doff, max_len, scores, cigar_stride, d_cig, d_ciglen, d_off, stream):
        self._check(_lib.isaac_ext_banded_sw_wide_batch_device(
            self._h, ctypes.c_uint32(band), ctypes.c_uint32(n), ctypes.c_void_p(d_q), ctypes.c_void_p(d_qoff), ctypes.c_void_p(d_qlen),
            ctypes.c_void_p(d_db), ctypes.c_void_p(d_doff), ctypes.c_uint32(max_len), ctypes.c_int(scores[0]), ctypes.c_int(scores[1]),
            ctypes.c_int(scores[2]), ctypes.c_int(scores[3]), ctypes.c_uint32(cigar_stride), ctypes.c_void_p(d_cig), ctypes.c_void_p(d_ciglen),
            ctypes.c_void_p(d_off), ctypes.c_void_p(stream)))

    def ungapped(self, candidates, with_masks=True, out=None):
        cand = np.ascontiguousarray(candidates, dtype=CANDIDATE_DTYPE)
        n = len(cand)
        frags, cig, mask = out if out is not None else (
            np.zeros(n, dtype=FRAGMENT_DTYPE), np.zeros((n, 3), dtype=np.uint32),
            np.zeros((n, MASK_WORDS), dtype=np.uint64) if with_masks else None)
        self._check(_lib.isaac_ext_ungapped_batch(self._h, ctypes.c_uint32(n), _p(cand), _p(frags), _p(cig), _p(mask)))
        return frags, cig, mask

    def gapped(self, candidates, cigar_stride=32, with_masks=True, out=None):
        cand = np.ascontiguousarray(candidates, dtype=CANDIDATE_DTYPE)
        n = len(cand)
        frags, cig, mask = out if out is not None else (
            np.zeros(n, dtype=FRAGMENT_DTYPE), np.zeros((n, cigar_stride), dtype=np.uint32),
            np.zeros((n, MASK_WORDS), dtype=np.uint64) if with_masks else None)
        self._check(_lib.isaac_ext_gapped_batch(self._h, ctypes.c_uint32(n), _p(cand), ctypes.c_uint32(cigar_stride),
                                                _p(frags), _p(cig), _p(mask)))
        return frags, cig, mask

    def extend_compact(self, candidates, gapped, fragments_out, pool_out):
        """end-to-end variant: fixed 64-byte records + dense CIGAR pool, chunked and overlapped; returns words used"""
        cand = np.ascontiguousarray(candidates, dtype=CANDIDATE_DTYPE)
        words = ctypes.c_uint64()
        fn = _lib.isaac_ext_gapped_batch_compact if gapped else _lib.isaac_ext_ungapped_batch_compact
        self._check(fn(self._h, ctypes.c_uint32(len(cand)), _p(cand), _p(fragments_out), _p(pool_out),
                       ctypes.c_uint64(pool_out.size), ctypes.byref(words)))
        return int(words.value)

    def extend_compact_both(self, candidates, ungapped_out, ungapped_pool, gapped_out, gapped_pool):
        """ungapped + gapped pass over the same candidates in one chunked call; returns (ungapped words, gapped words)"""
        cand = np.ascontiguousarray(candidates, dtype=CANDIDATE_DTYPE)
        wu, wg = ctypes.c_uint64(), ctypes.c_uint64()
        self._check(_lib.isaac_ext_extend_batch_compact(
            self._h, ctypes.c_uint32(len(cand)), _p(cand), _p(ungapped_out), _p(ungapped_pool), ctypes.c_uint64(ungapped_pool.size),
            ctypes.byref(wu), _p(gapped_out), _p(gapped_pool), ctypes.c_uint64(gapped_pool.size), ctypes.byref(wg)))
        return int(wu.value), int(wg.value)

    def align_packed(self, candidates, alignments_out, pool_out):
        """isaac_ext_align_batch_packed: one 32-byte record per candidate (types.ALIGNMENT_DTYPE) + the words of the accepted gapped
        CIGARs; returns the number of words"""
        cand = np.ascontiguousarray(candidates, dtype=CANDIDATE_DTYPE)
        w = ctypes.c_uint64()
        self._check(_lib.isaac_ext_align_batch_packed(self._h, ctypes.c_uint32(len(cand)), _p(cand), _p(alignments_out), _p(pool_out),
                                                      ctypes.c_uint64(pool_out.size), ctypes.byref(w)))
        return int(w.value)

    def build_fragments(self, match_batch, copy=True):
        """FragmentBuilder::build for every cluster of the resident read set -> batch.FlatFragments
        (begin per cluster * readCount + readIndex, flags = return value of build())"""
        from .batch import BuildResult, copy_result
        res = BuildResult()
        self._check(_lib.isaac_ext_build_fragments(self._h, ctypes.byref(match_batch.c), ctypes.byref(res)))
        if not copy:
            return res
        return copy_result(res, self.reads.cluster_count * self.reads.read_count, "readFragmentBegin", "built",
                           self.reads.cluster_count)

    def determine_template_length(self, match_batch, pf=None, mate_drift_range=-1):
        """MatchSelector::determineTemplateLength for the resident tile -> (batch.Tls, stable)"""
        from .batch import Tls
        tls = Tls()
        stable = ctypes.c_uint32()
        pf_arr = None if pf is None else np.ascontiguousarray(pf, dtype=np.uint8)
        self._check(_lib.isaac_ext_determine_template_length(self._h, ctypes.byref(match_batch.c), _p(pf_arr),
                                                             ctypes.c_int32(mate_drift_range), ctypes.byref(tls), ctypes.byref(stable)))
        return tls, bool(stable.value)

    def rescue_shadows(self, tls, requests, copy=True):
        """ShadowAligner::rescueShadow for every request -> batch.FlatFragments (begin per request, flags = rescued)"""
        from .batch import RESCUE_REQUEST_DTYPE, RescueResult, copy_result
        req = np.ascontiguousarray(requests, dtype=RESCUE_REQUEST_DTYPE)
        res = RescueResult()
        self._check(_lib.isaac_ext_rescue_shadows(self._h, ctypes.byref(tls), ctypes.c_uint32(len(req)), _p(req),
                                                  ctypes.byref(res)))
        if not copy:
            return res
        return copy_result(res, len(req), "requestFragmentBegin", "rescued", len(req))

    def trim_low_quality_ends(self, base_quality_cutoff):
        """alignment::trimLowQualityEnds on the resident reads; returns the new endCyclesMasked [clusters, readCount]"""
        out = np.zeros((self.reads.cluster_count, self.reads.read_count), dtype=np.uint16)
        self._check(_lib.isaac_ext_trim_low_quality_ends(self._h, ctypes.c_uint32(base_quality_cutoff), _p(out)))
        return out

    def build_templates(self, match_batch, tls, options=None, copy=True):
        """TemplateBuilder::buildFragments + buildTemplate for every cluster of the resident read set -> batch.Templates"""
        from .batch import TemplateOptions, TemplateResult
        options = options if options is not None else TemplateOptions.make()
        res = TemplateResult()
        self._check(_lib.isaac_ext_build_templates(self._h, ctypes.byref(match_batch.c), ctypes.byref(tls), ctypes.byref(options),
                                                   ctypes.byref(res)))
        return self._templates(res) if copy else res

    def build_templates_deferred(self, match_batch, tls, options=None):
        """isaac_ext_build_templates_deferred: the tile's kernels are queued and sized, the download runs on a copy stream next to
        whatever is called next; returns the handle fetch_templates takes (at most two may be waiting)"""
        from .batch import TemplateOptions, TemplateResult
        options = options if options is not None else TemplateOptions.make()
        res = TemplateResult()
        self._check(_lib.isaac_ext_build_templates_deferred(self._h, ctypes.byref(match_batch.c), ctypes.byref(tls), ctypes.byref(options),
                                                            ctypes.byref(res)))
        return res, self.reads.cluster_count, self.reads.read_count

    def fetch_templates(self, handle, copy=True):
        """isaac_ext_fetch_templates: waits for the download of a deferred tile -> batch.Templates"""
        res, n, rc = handle
        self._check(_lib.isaac_ext_fetch_templates(self._h, ctypes.byref(res)))
        return self._templates(res, n, rc) if copy else res

    def _templates(self, res, n=None, read_count=None):
        from .batch import TEMPLATE_DTYPE, Templates
        n = self.reads.cluster_count if n is None else n
        read_count = self.reads.read_count if read_count is None else read_count

        def arr(ptr, dtype, count):
            if not count:
                return np.zeros(0, dtype=dtype)
            buf = (ctypes.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
            return np.frombuffer(buf, dtype=dtype).copy()

        return Templates(arr(res.templates, TEMPLATE_DTYPE, n), arr(res.fragments, FRAGMENT_DTYPE, n * read_count),
                         arr(res.cigars, np.uint32, int(res.cigarWords)), int(res.rescueRequests))

    def submit_build_templates(self, match_batch, tls, options=None):
        """isaac_ext_submit_build_templates: returns a ticket at once, the call runs on a worker thread of the context; the
        arrays of match_batch must stay alive and unchanged until wait_templates"""
        from .batch import TemplateOptions
        options = options if options is not None else TemplateOptions.make()
        ticket = ctypes.c_uint64()
        self._check(_lib.isaac_ext_submit_build_templates(self._h, ctypes.byref(match_batch.c), ctypes.byref(tls), ctypes.byref(options),
                                                          ctypes.byref(ticket)))
        self._in_flight = match_batch
        return int(ticket.value)

    def wait_templates(self, ticket):
        """isaac_ext_wait for a ticket of submit_build_templates -> batch.Templates"""
        from .batch import TemplateResult
        res = TemplateResult()
        self._check(_lib.isaac_ext_wait(self._h, ctypes.c_uint64(ticket), ctypes.byref(res)))
        self._in_flight = None
        return self._templates(res)

    def template_stats(self, match_batch, tls, templates, pf=None):
        """matchSelector::TileBarcodeStats of a tile's templates (batch.Templates) -> uint64 [4, TEMPLATE_STATS_COUNTERS],
        row = readIndex * 2 + passesFilter"""
        from .batch import TemplateResult
        from .distributed import TEMPLATE_STATS_COUNTERS
        t = np.ascontiguousarray(templates.templates)
        f = np.ascontiguousarray(templates.fragments)
        cig = np.ascontiguousarray(templates.cigars, dtype=np.uint32)
        res = TemplateResult(t.ctypes.data, f.ctypes.data, cig.ctypes.data if cig.size else None, cig.size, 0)
        pf_arr = None if pf is None else np.ascontiguousarray(pf, dtype=np.uint8)
        out = np.zeros((4, TEMPLATE_STATS_COUNTERS), dtype=np.uint64)
        self._check(_lib.isaac_ext_template_stats(self._h, ctypes.byref(match_batch.c), ctypes.byref(tls), ctypes.byref(res),
                                                  _p(pf_arr), _p(out)))
        return out

    def tile_cycle_stats(self, pf=None, finalize=False):
        """matchSelector::TileStats (score histograms + per-cycle arrays) of the templates the last build_templates / select_tile
        left on the device -> uint64 [4, TILE_CYCLE_STATS_WORDS], row = readIndex * 2 + passesFilter"""
        words = 47105
        out = np.zeros((4, words), dtype=np.uint64)
        pf_arr = None if pf is None else np.ascontiguousarray(pf, dtype=np.uint8)
        self._check(_lib.isaac_ext_tile_cycle_stats(self._h, _p(pf_arr), _p(out)))
        if finalize:
            for k in range(4):
                _lib.isaac_ext_tile_cycle_stats_finalize(ctypes.c_void_p(out[k].ctypes.data))
        return out

    def pack_fragments(self, templates, options=None, copy=True):
        """matchSelector::FragmentCollector::add for every stored template of the resident tile (batch.Templates ->
        batch.PackedFragments: the io::FragmentHeader bin records in FragmentBuffer layout)"""
        from .batch import PackedFragments, PackOptions, PackResultC, TemplateResult
        options = options if options is not None else PackOptions()
        t = np.ascontiguousarray(templates.templates)
        f = np.ascontiguousarray(templates.fragments)
        cig = np.ascontiguousarray(templates.cigars, dtype=np.uint32)
        tr = TemplateResult(t.ctypes.data, f.ctypes.data, cig.ctypes.data if cig.size else None, cig.size, 0)
        res = PackResultC()
        self._check(_lib.isaac_ext_pack_fragments(self._h, ctypes.byref(tr), ctypes.byref(options.c), ctypes.byref(res)))
        if not copy:
            return res                     # pointers into the context's buffers (bench.py: no host copy inside the timed region)
        return self._packed(res, bool(options.c.compact))

    def realign_bin(self, bin_, options):
        """build::GapRealigner over one bin (bins.Bin, bins.RealignOptions -> bins.RealignResult): BinSorter::collectGaps +
        BinSorter::realignGaps on the GPU against the resident reference; bin_.data is left as it was, the updated records are in
        the result"""
        from . import bins
        data = np.ascontiguousarray(bin_.data).copy()
        offsets = np.ascontiguousarray(bin_.record_offset, dtype=np.uint64) if bin_.record_offset is not None else None
        index = np.ascontiguousarray(bin_.index)
        res = bins.RealignResultC()
        self._check(_lib.isaac_ext_realign_bin(self._h, ctypes.byref(options.c), ctypes.c_void_p(data.ctypes.data), ctypes.c_uint64(data.size),
                                               ctypes.c_void_p(offsets.ctypes.data) if offsets is not None else None,
                                               ctypes.c_uint64(offsets.size if offsets is not None else 0),
                                               ctypes.c_void_p(index.ctypes.data), ctypes.c_uint64(index.size), ctypes.byref(res)))

        def arr(ptr, dtype, count):
            if not count:
                return np.zeros(0, dtype=dtype)
            buf = (ctypes.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
            return np.frombuffer(buf, dtype=dtype).copy()

        n = index.size
        return bins.RealignResult(data, arr(res.position, np.uint64, n), arr(res.cigarOffset, np.uint32, n), arr(res.cigarLength, np.uint32, n),
                                  arr(res.realignedCigars, np.uint32, int(res.realignedCigarWords)), arr(res.gaps, bins.GAP_DTYPE, int(res.gapCount)),
                                  arr(res.deletionsByEnd, bins.GAP_DTYPE, int(res.deletionCount)), int(res.realignedFragments),
                                  float(res.collectMs), float(res.realignMs))

    def realign_bins(self, bin_list, options_list):
        """isaac_ext_realign_bins: every bin of the list in one call (two slots of the context take them in turn); returns one
        bins.RealignResult per bin (without the gap lists, which only the single-bin call hands out)"""
        from . import bins
        n = len(bin_list)
        jobs = (bins.RealignJobC * n)()
        keep = []
        for k, (b, o) in enumerate(zip(bin_list, options_list)):
            data = np.ascontiguousarray(b.data).copy()
            offsets = np.ascontiguousarray(b.record_offset, dtype=np.uint64) if b.record_offset is not None else None
            index = np.ascontiguousarray(b.index)
            m = index.size
            position, cigar_offset, cigar_length = np.zeros(m, np.uint64), np.zeros(m, np.uint32), np.zeros(m, np.uint32)
            cigars = np.zeros(16 * m + 64, np.uint32)
            keep.append((data, offsets, index, position, cigar_offset, cigar_length, cigars))
            j = jobs[k]
            j.options = ctypes.addressof(o.c)
            j.data, j.dataBytes = data.ctypes.data if data.size else None, data.size
            j.recordOffset, j.recordCount = (offsets.ctypes.data if offsets is not None and offsets.size else None), (offsets.size if offsets is not None else 0)
            j.index, j.indexCount = index.ctypes.data if m else None, m
            j.position, j.cigarOffset, j.cigarLength = position.ctypes.data, cigar_offset.ctypes.data, cigar_length.ctypes.data
            j.realignedCigars, j.realignedCigarCapacity = cigars.ctypes.data, cigars.size
        self._check(_lib.isaac_ext_realign_bins(self._h, jobs, ctypes.c_uint32(n)))
        out = []
        for k, (data, offsets, index, position, cigar_offset, cigar_length, cigars) in enumerate(keep):
            out.append(bins.RealignResult(data, position, cigar_offset, cigar_length, cigars[:int(jobs[k].realignedCigarWords)], None, None,
                                          int(jobs[k].realignedFragments)))
        return out

    def _packed(self, res, compact=False):
        from .batch import PackedFragments
        n, rc = self.reads.cluster_count, self.reads.read_count

        def arr(ptr, dtype, count):
            buf = (ctypes.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
            return np.frombuffer(buf, dtype=dtype).copy()

        records = arr(res.records, np.uint8, int(res.recordBytes)) if res.recordBytes else np.zeros(0, np.uint8)
        if not compact:
            records = records.reshape(n, res.recordLength)
        return PackedFragments(records, arr(res.fStrandPos, np.uint64, n * rc).reshape(n, rc),
                               arr(res.initialized, np.uint8, n * rc).reshape(n, rc), int(res.recordLength),
                               (int(res.readOffset[0]), int(res.readOffset[1])), int(res.headerLength), int(res.storedFragments),
                               arr(res.recordOffset, np.uint64, n * rc + 1))

    def tile_stats_device(self, n, d_fragments, d_stats, stream):
        """adds the K6 counters of n device-resident fragment records to the 64 u64 at d_stats"""
        self._check(_lib.isaac_ext_tile_stats_device(self._h, ctypes.c_uint32(n), ctypes.c_void_p(d_fragments),
                                                     ctypes.c_void_p(d_stats), ctypes.c_void_p(stream)))

    def measure_int32_peak(self, kind=0):
        """operations per second of the integer pipes (0: add.s32, 1: max.s32, 2: 16x2 max counted twice)"""
        v = ctypes.c_double()
        self._check(_lib.isaac_ext_measure_int32_peak(self._h, ctypes.c_int(kind), ctypes.byref(v)))
        return v.value

    # device-resident variants: arguments are raw device pointers (e.g. torch tensor .data_ptr()) and a stream handle
    def ungapped_device(self, n, d_candidates, d_fragments, d_cigars, d_masks, stream):
        self._check(_lib.isaac_ext_ungapped_batch_device(
            self._h, ctypes.c_uint32(n), ctypes.c_void_p(d_candidates), ctypes.c_void_p(d_fragments),
            ctypes.c_void_p(d_cigars), ctypes.c_void_p(d_masks) if d_masks else None, ctypes.c_void_p(stream)))

    def gapped_device(self, n, d_candidates, cigar_stride, d_fragments, d_cigars, d_masks, stream):
        self._check(_lib.isaac_ext_gapped_batch_device(
            self._h, ctypes.c_uint32(n), ctypes.c_void_p(d_candidates), ctypes.c_uint32(cigar_stride),
            ctypes.c_void_p(d_fragments), ctypes.c_void_p(d_cigars), ctypes.c_void_p(d_masks) if d_masks else None,
            ctypes.c_void_p(stream)))


def version():
    return _lib.isaac_ext_version().decode()

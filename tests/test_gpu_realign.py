"""isaac_ext_realign_bin on the GPU against the reference's own build::RealignerGaps / build::GapRealigner /
build::SemialignedEndsClipper (oracle_realign_bin): every byte of the bin's data afterwards, Index::pos_ and the CIGAR of every index
entry, gapGroups_ and deletionEndGroups_ -- the latter also where the reference's unstable std::sort decides the order of ties."""
import numpy as np
import pytest

import oracle_lib
from isaac_aligner_b200 import bins
from isaac_aligner_b200.batch import Tls
from isaac_aligner_b200.types import Config
from test_realign_host import compare, make_contigs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    return oracle_lib.require_reference()


@pytest.fixture(scope="module")
def ctx_and_contigs():
    from isaac_aligner_b200 import capi
    contigs = make_contigs(77, lengths=(3000, 120000))
    ctx = capi.Context(Config.default(max_read_length=256))
    ctx.set_reference(contigs)
    yield ctx, contigs
    ctx.close()


def data_only_record(contig, position, cigar, read_length=100, barcode=0):
    """a record that only brings gaps along (it is in the bin's data, not in its index)"""
    h = np.zeros(1, dtype=bins.HEADER_DTYPE)[0]
    h["fStrandPosition"] = bins.reference_position(contig, position)
    h["mateFStrandPosition"] = bins.reference_position(contig, position)
    h["readLength"], h["cigarLength"] = read_length, len(cigar)
    h["gapCount"] = sum(1 for _, op in cigar if op in (bins.OP_INSERT, bins.OP_DELETE))
    h["flags"] = bins.FLAG_PAIRED | bins.FLAG_FIRST_READ
    h["barcode"] = barcode
    words = np.array([(n << 4) | op for n, op in cigar], dtype=np.uint32)
    return h.tobytes() + bytes(read_length) + words.tobytes()


@pytest.mark.parametrize("seed,vigorous,clip,dodgy,L,spacing", [(11, False, False, False, 100, 220), (12, False, True, False, 150, 120),
                                                                  (13, True, True, True, 100, 60), (14, True, False, False, 250, 220),
                                                                  (15, False, True, True, 36, 25)])
def test_realign_bin_equals_reference(ctx_and_contigs, ref, seed, vigorous, clip, dodgy, L, spacing):
    ctx, contigs = ctx_and_contigs
    genome = oracle_lib.GenomeHolder(contigs)
    bin_ = bins.simulate_bin(contigs, contig=1, region=(1500, 90000), n_pairs=12000, read_length=L, seed=seed, barcodes=3,
                             variant_spacing=spacing, template_mean=int(2.6 * L) + 60, edge_fraction=0.01)
    tls = [Tls.make(mn=int(2.0 * L), mx=int(3.4 * L) + 120, median=int(2.6 * L) + 60) for _ in range(3)]
    options = bins.RealignOptions(bin_.bin_start, bin_.bin_end, tls, vigorous=vigorous, dodgy=dodgy, clip_semialigned=clip,
                                  gap_groups=[0, 1, 0] if seed % 2 else None)
    want = oracle_lib.realign_bin(ref, genome, bin_, options)
    got = ctx.realign_bin(bin_, options)
    realigned = compare(bin_, got, want)
    assert realigned > 200 and got.realigned == realigned


def test_deletions_that_end_at_the_same_base(ctx_and_contigs, ref):
    """ties in deletionEndGroups_: more than 16 deletions with pairwise equal ends, so that the reference's introsort partitions"""
    ctx, contigs = ctx_and_contigs
    genome = oracle_lib.GenomeHolder(contigs)
    bin_ = bins.simulate_bin(contigs, contig=1, region=(1500, 40000), n_pairs=3000, read_length=100, seed=21)
    rng = np.random.default_rng(5)
    extra, offsets, at = [], [], int(bin_.data.size)
    for k in range(160):
        p = 2000 + 37 * (k // 4)
        first = 30 + 2 * (k % 4)                                             # 30M8D, 32M6D, 34M4D, 36M2D: all four end at p + 38
        blob = data_only_record(1, p, [(first, bins.OP_ALIGN), (8 - 2 * (k % 4), bins.OP_DELETE), (100 - first, bins.OP_ALIGN)])
        extra.append(blob); offsets.append(at); at += len(blob)
    order = rng.permutation(len(extra))
    data = np.concatenate([bin_.data] + [np.frombuffer(extra[i], dtype=np.uint8) for i in order])
    offs, at = [], int(bin_.data.size)
    for i in order:
        offs.append(at); at += len(extra[i])
    tied = bins.Bin(data, np.concatenate([bin_.record_offset, np.array(offs, dtype=np.uint64)]), bin_.index, bin_.bin_start, bin_.bin_end)
    options = bins.RealignOptions(tied.bin_start, tied.bin_end, [Tls.make()], clip_semialigned=True)
    want = oracle_lib.realign_bin(ref, genome, tied, options)
    ends = want.deletions["position"] + 2 * want.deletions["length"].astype(np.uint64)
    assert np.count_nonzero(ends[1:] == ends[:-1]) >= 100
    stable = np.lexsort((want.deletions["length"], want.deletions["position"], ends))
    assert not np.array_equal(want.deletions[stable], want.deletions)        # the reference's order is not the stable one: the host step matters
    got = ctx.realign_bin(tied, options)
    compare(tied, got, want)
    # the chain of records walked by the library itself gives the same
    walked = bins.Bin(data, None, bin_.index, bin_.bin_start, bin_.bin_end)
    again = ctx.realign_bin(walked, options)
    assert np.array_equal(again.data, got.data) and np.array_equal(again.deletions, got.deletions)


def test_several_bins_in_one_call(ctx_and_contigs, ref):
    """isaac_ext_realign_bins: two slots of the context take the bins in turn; every bin as if it had been realigned alone"""
    from isaac_aligner_b200 import capi
    ctx, contigs = ctx_and_contigs
    genome = oracle_lib.GenomeHolder(contigs)
    bin_list, options = [], []
    for k in range(7):
        b = bins.simulate_bin(contigs, contig=1, region=(1500 + 9000 * k, 1500 + 9000 * (k + 1)), n_pairs=400 + 700 * (k % 3), read_length=100, seed=40 + k)
        bin_list.append(b)
        options.append(bins.RealignOptions(b.bin_start, b.bin_end, [Tls.make()], vigorous=bool(k % 2), clip_semialigned=True))
    empty = bins.Bin(np.zeros(0, dtype=np.uint8), np.zeros(0, dtype=np.uint64), np.zeros(0, dtype=bins.BIN_INDEX_DTYPE), bin_list[0].bin_start, bin_list[0].bin_end)
    bin_list.insert(3, empty); options.insert(3, options[0])
    got = ctx.realign_bins(bin_list, options)
    for k, (b, o) in enumerate(zip(bin_list, options)):
        if not len(b.index):
            assert got[k].position.size == 0
            continue
        want = oracle_lib.realign_bin(ref, genome, b, o)
        got[k].gaps, got[k].deletions = want.gaps, want.deletions           # the batched call does not hand the gap lists out
        assert compare(b, got[k], want) == got[k].realigned
    # a pool that is too small for one job: that job alone fails, loudly
    jobs_small = ctx.realign_bins  # the harness sizes the pools generously; the C call with a tiny pool is exercised here
    import ctypes
    b, o = bin_list[0], options[0]
    data = b.data.copy(); index = np.ascontiguousarray(b.index); m = index.size
    pos, co, cl, cig = np.zeros(m, np.uint64), np.zeros(m, np.uint32), np.zeros(m, np.uint32), np.zeros(4, np.uint32)
    job = (bins.RealignJobC * 1)()
    job[0].options = ctypes.addressof(o.c); job[0].data = data.ctypes.data; job[0].dataBytes = data.size
    job[0].index = index.ctypes.data; job[0].indexCount = m
    job[0].position, job[0].cigarOffset, job[0].cigarLength = pos.ctypes.data, co.ctypes.data, cl.ctypes.data
    job[0].realignedCigars, job[0].realignedCigarCapacity = cig.ctypes.data, cig.size
    rc = capi._lib.isaac_ext_realign_bins(ctx._h, job, ctypes.c_uint32(1))
    assert rc == 5 and job[0].status == 5                                   # ISAAC_EXT_E_CAPACITY


def test_argument_errors(ctx_and_contigs):
    from isaac_aligner_b200 import capi
    ctx, contigs = ctx_and_contigs
    bin_ = bins.simulate_bin(contigs, contig=1, region=(1500, 9000), n_pairs=300, read_length=100, seed=3)
    good = bins.RealignOptions(bin_.bin_start, bin_.bin_end, [Tls.make()])
    broken = bins.Bin(bin_.data[:-5], bin_.record_offset, bin_.index, bin_.bin_start, bin_.bin_end)
    with pytest.raises(capi.ExtError) as e:
        ctx.realign_bin(broken, good)                                        # the last record runs out of the data
    assert e.value.code == 1
    elsewhere = bins.RealignOptions(bins.reference_position(7, 0), bins.reference_position(7, 100), [Tls.make()])
    with pytest.raises(capi.ExtError):
        ctx.realign_bin(bin_, elsewhere)                                     # no such contig
    empty = bins.Bin(np.zeros(0, dtype=np.uint8), np.zeros(0, dtype=np.uint64), np.zeros(0, dtype=bins.BIN_INDEX_DTYPE), bin_.bin_start, bin_.bin_end)
    res = ctx.realign_bin(empty, good)
    assert res.position.size == 0 and res.gaps.size == 0


def test_records_of_the_tile_pipeline_feed_the_realigner(ref):
    """the chain a run takes: match selection of a tile -> io::FragmentHeader records (isaac_ext_select_tile with pack, compact =
    the bin file) -> gap realigner over those very bytes.  The reference's GapRealigner runs on the same bytes: it asserts the
    coherence of what it reads (TLEN of mates, soft clips against lowClipped / highClipped, CIGAR lengths), so this is also the
    check that the packed records are bins the build stage accepts."""
    from common_build import build_workload
    from isaac_aligner_b200 import capi, tile
    from isaac_aligner_b200.batch import PackOptions, TemplateOptions
    n, L = 12000, 100
    genome, sim, reads, mb = build_workload(n_pairs=n, L=L, seed=808, genome_bases=150_000, indel_rate=6e-3, masked=False)   # ~16x, gaps everywhere
    ctx = capi.Context(Config.default(max_read_length=2 * L))
    ctx.set_reference(genome)
    pack = PackOptions(tile=1101, barcode_idx=0, keep_unaligned=False, compact=True)
    tls = Tls.make()
    got_tile = tile.select_tile(ctx, reads.bcl, (L, L), mb.matches, mb.seeds, tls=tls, options=TemplateOptions.make(clip_semialigned=True), pack=pack)
    packed = got_tile.packed
    offsets = packed.record_offset
    lengths = np.diff(offsets.astype(np.int64))
    stored = np.flatnonzero(lengths > 0)
    assert stored.size > 1.5 * n
    data = np.ascontiguousarray(packed.records)
    # one bin per contig: the records whose position lies on it; index in (cluster, read) order, mates linked when both are in the bin
    total = 0
    for contig, bases in enumerate(genome):
        start, end = bins.reference_position(contig, 0), bins.reference_position(contig, len(bases))
        headers = {int(i): np.frombuffer(data[int(offsets[i]):int(offsets[i]) + 112].tobytes(), dtype=bins.HEADER_DTYPE)[0] for i in stored}
        inside = [i for i in stored if start <= int(headers[int(i)]["fStrandPosition"]) < end]
        if not inside:
            continue
        # the bin's own data: its records back to back
        blob, new_offset = [], {}
        at = 0
        for i in inside:
            new_offset[int(i)] = at
            blob.append(data[int(offsets[i]):int(offsets[i + 1])]); at += int(lengths[i])
        bin_data = np.concatenate(blob)
        index = []
        for i in inside:
            mate = int(i) ^ 1
            index.append((new_offset[int(i)], new_offset.get(mate, new_offset[int(i)])))
        b = bins.Bin(bin_data, np.array([new_offset[int(i)] for i in inside], dtype=np.uint64), np.array(index, dtype=bins.BIN_INDEX_DTYPE), start, end)
        options = bins.RealignOptions(start, end, [tls], clip_semialigned=True)
        want = oracle_lib.realign_bin(ref, oracle_lib.GenomeHolder(genome), b, options)
        got = ctx.realign_bin(b, options)
        total += compare(b, got, want)
        assert got.gaps.size > 100
    assert total > 0                                             # some read was repaired with a gap another read brought along
    ctx.close()

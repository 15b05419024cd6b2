/*
 * isaac_ext.h -- C ABI of the B200-native candidate-extension path of the Isaac aligner.
 *
 * Drop-in boundary.  The reference has no plugin/FFI layer: its seam is the C++ class API that
 * alignment::TemplateBuilder calls once per cluster (reference citations are path:line under
 * src/c++/ of sequencing/isaac_aligner):
 *
 *   FragmentBuilder::build                 include/alignment/FragmentBuilder.hh:62-70
 *   FragmentBuilder::getFragments/Cigar    include/alignment/FragmentBuilder.hh:71-72
 *   ShadowAligner::rescueShadow            include/alignment/ShadowAligner.hh:81-88
 *   UngappedAligner::alignUngapped         include/alignment/fragmentBuilder/UngappedAligner.hh:55-60
 *   GappedAligner::alignGapped             include/alignment/fragmentBuilder/GappedAligner.hh:49-54
 *   SimpleIndelAligner::alignSimpleIndels  include/alignment/fragmentBuilder/SimpleIndelAligner.hh:50-55
 *   BandedSmithWaterman::align             include/alignment/BandedSmithWaterman.hh:81-86
 *
 * This library replaces those per-cluster calls by per-tile batch calls over plain pointers.  The
 * C++ classes carrying the reference's names (isaac_aligner_b200/host/, namespace isaac_b200) are
 * thin callers of this ABI.  There is no CPU fallback: every compute entry point runs CUDA kernels
 * on an sm_100a device and returns ISAAC_EXT_E_NO_DEVICE when none is usable.
 *
 * Conventions: every function returns an int status (0 = ok), no exceptions cross the boundary,
 * caller owns all buffers it passes in, result buffers owned by the context stay valid until the next
 * call on that context.  One context per GPU; a context is not re-entrant (like the reference's
 * builders, MatchSelector.cpp:143-163 keeps one per thread).
 */
#ifndef ISAAC_EXT_H
#define ISAAC_EXT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------------------------- */
#define ISAAC_EXT_OK             0
#define ISAAC_EXT_E_INVALID_ARG  1 /* reference: common::InvalidParameterException / assert()       */
#define ISAAC_EXT_E_NO_DEVICE    2 /* no usable CUDA device: there is deliberately no CPU fallback   */
#define ISAAC_EXT_E_CUDA         3 /* CUDA runtime error, see isaac_ext_last_error                   */
#define ISAAC_EXT_E_UNSUPPORTED  4 /* a combination this library refuses (see isaac_ext_create)       */
#define ISAAC_EXT_E_CAPACITY     5 /* a fixed capacity of the reference was exceeded (cigar stride)  */
#define ISAAC_EXT_E_NO_REFERENCE 6 /* isaac_ext_set_reference has not been called                    */

/* ---- constants fixed by the reference ------------------------------------------------------- */
#define ISAAC_EXT_BAND_WIDTH          16 /* BandedSmithWaterman::WIDEST_GAP_SIZE  BandedSmithWaterman.hh:88-89   */
#define ISAAC_EXT_SW_MISMATCH_CUTOFF   5 /* BandedSmithWaterman::mismatchesCutoff BandedSmithWaterman.hh:94      */
#define ISAAC_EXT_SW_DISTANCE_CUTOFF   7 /* BandedSmithWaterman::distanceCutoff   BandedSmithWaterman.hh:91-92   */
#define ISAAC_EXT_SHADOW_KMER          7 /* ShadowAligner::shadowKmerLength_      ShadowAligner.hh:95            */
#define ISAAC_EXT_SHADOW_POSITIONS 10000 /* candidate position cap                ShadowAligner.hh:92            */
#define ISAAC_EXT_MAX_CYCLES        1024 /* FragmentMetadata::maxCycles_          FragmentMetadata.hh:368        */
#define ISAAC_EXT_MASK_WORDS (ISAAC_EXT_MAX_CYCLES / 64)

/* CIGAR word = length << 4 | op, ops as in alignment::Cigar::OpCode (Cigar.hh:52-63,156-168). */
#define ISAAC_EXT_CIGAR_ALIGN     0u
#define ISAAC_EXT_CIGAR_INSERT    1u
#define ISAAC_EXT_CIGAR_DELETE    2u
#define ISAAC_EXT_CIGAR_SOFT_CLIP 4u

typedef struct isaac_ext_ctx isaac_ext_ctx;

/* Mirrors the constructor arguments of FragmentBuilder / ShadowAligner (FragmentBuilder.hh:49-60,
 * ShadowAligner.hh:52-59) plus the device plumbing.  Scores are passed exactly as the reference takes
 * them (bwa preset 0:-3:-11:-4:-20, eland 2:-1:-15:-3:-25; AlignOptions.cpp:55-56,687-741). */
typedef struct isaac_ext_config {
    int32_t  gapMatchScore;
    int32_t  gapMismatchScore;
    int32_t  gapOpenScore;
    int32_t  gapExtendScore;
    int32_t  minGapExtendScore;
    uint32_t repeatThreshold;       /* --repeat-threshold, default 10      */
    uint32_t maxSeedsPerRead;
    uint32_t gappedMismatchesMax;   /* --gapped-mismatches, default 5      */
    uint32_t semialignedGapLimit;   /* --semialigned-gap-limit, default 100; 0 disables simple indels */
    uint32_t avoidSmithWaterman;    /* --avoid-smith-waterman: GappedAligner::makesSenseToGapAlign decides (GappedAligner.cpp:88-165) */
    uint32_t maxReadLength;         /* flowcell::getMaxTotalReadLength; bounds the SW overflow check  */
    int32_t  device;                /* CUDA device ordinal                                            */
    uint32_t hostThreads;           /* threads for the per-cluster bookkeeping (0 = hardware)          */
} isaac_ext_config_t;

/* Flat equivalent of the FragmentMetadata fields the template layer reads (FragmentMetadata.hh:330-414).
 * 64 bytes.  mismatchCycles[] is replaced by a bit mask over strand-order base indices (see
 * isaac_ext_*_batch 'mismatchMask'): cycle = reverse ? lastCycle - i : firstCycle + i
 * (AlignerBase.cpp:171), emitted in increasing i exactly like addMismatchCycle does. */
typedef struct isaac_ext_fragment {
    int64_t  position;               /* FragmentMetadata::position                                   */
    double   logProbability;         /* ordered FP64 sum, bit-identical to the reference             */
    uint32_t contigId;
    uint32_t readId;                 /* cluster * readCount + readIndex                              */
    uint32_t cigarOffset;            /* word index into the cigar pool of the call that produced it  */
    uint32_t smithWatermanScore;
    uint32_t observedLength;
    uint16_t mismatchCount;
    uint16_t matchesInARow;
    uint16_t gapCount;
    uint16_t editDistance;
    uint16_t uniqueSeedCount;
    uint16_t repeatSeedsCount;
    uint16_t nonUniqueSeedOffsetFirst;  /* 0xFFFF = unset (reference: UINT_MAX)                      */
    uint16_t nonUniqueSeedOffsetSecond;
    int16_t  firstSeedIndex;
    uint16_t lowClipped;
    uint16_t highClipped;
    uint16_t cigarLength;            /* 0 = unaligned (FragmentMetadata::isAligned)                  */
    uint8_t  reverse;
    uint8_t  readIndex;
    uint16_t matchCount;             /* return value of updateFragmentCigar for this alignment       */
} isaac_ext_fragment_t;

/* One tile's clusters in the reference's own BclClusters layout (BclClusters.hh:33-124): one byte per
 * base = quality << 2 | base, bytes 0..3 = N (Nucleotides.hh:91-94), reads of a cluster back to back. */
typedef struct isaac_ext_reads {
    uint32_t clusterCount;
    uint32_t readCount;              /* 1 or 2                                                        */
    uint32_t readLength[2];          /* flowcell::ReadMetadata::getLength                             */
    uint32_t firstCycle[2];          /* flowcell::ReadMetadata::getFirstCycle (cycles are contiguous) */
    const uint8_t  *bcl;             /* clusterCount * (readLength[0] + readLength[1]) bytes          */
    const uint16_t *endCyclesMasked; /* clusterCount * readCount (Read::endCyclesMasked_) or NULL     */
} isaac_ext_reads_t;

/* A candidate placement of one read strand (what FragmentBuilder::addMatch / ShadowAligner produce). */
typedef struct isaac_ext_candidate {
    int64_t  position;               /* leftmost base on the forward strand, may be negative          */
    uint32_t readId;                 /* cluster * readCount + readIndex                               */
    uint32_t contigStrand;           /* contigId << 1 | reverse                                       */
} isaac_ext_candidate_t;             /* 16 bytes */

/* The reference's 16-byte Match record, bit for bit (Match.hh:38-73): seedId packs tile:12 barcode:12 cluster:31
 * seed:8 reverse:1 from the top (SeedId.hh:37-127), location packs (contigId+1):23 position:40 neighbors:1
 * (ReferencePosition.hh:51-188; contig field 0 = TooManyMatch, all ones = NoMatch).  A tile's match file can be
 * handed over unchanged. */
typedef struct isaac_ext_match {
    uint64_t seedId;
    uint64_t location;
} isaac_ext_match_t;

/* alignment::SeedMetadata (SeedMetadata.hh:46-98); the index of a seed is its position in the array. */
typedef struct isaac_ext_seed {
    uint16_t offset;
    uint16_t length;                 /* 16, 32 or 64 */
    uint32_t readIndex;
} isaac_ext_seed_t;

/* FragmentBuilder::build for every cluster of the resident read set (FragmentBuilder.cpp:82-145).  Matches are sorted
 * the way SelectMatchesTransition sorts them (cluster, location, seed; SelectMatchesTransition.cpp:242-254) and
 * delimited per cluster by clusterMatchBegin (clusterCount + 1 offsets; the cluster field of seedId is not consulted). */
typedef struct isaac_ext_build_batch {
    const isaac_ext_match_t *matches;
    const uint64_t *clusterMatchBegin;
    const isaac_ext_seed_t *seeds;
    uint32_t seedCount;
    uint32_t withGaps;               /* the 'withGaps' argument of build()                              */
} isaac_ext_build_batch_t;

/* getFragments() / getCigarBuffer() of all clusters, flattened.  Fragments of (cluster, readIndex) are
 * fragments[readFragmentBegin[cluster * readCount + readIndex] .. readFragmentBegin[... + 1]) in the reference's
 * final order; fragment.cigarOffset indexes 'cigars'.  Owned by the context, valid until its next call. */
typedef struct isaac_ext_build_result {
    const isaac_ext_fragment_t *fragments;
    const uint64_t *readFragmentBegin;   /* clusterCount * readCount + 1                                 */
    const uint32_t *cigars;
    const uint8_t *built;                /* per cluster: return value of build()                         */
    uint64_t fragmentCount;
    uint64_t cigarWords;
} isaac_ext_build_result_t;

/* alignment::TemplateLengthStatistics as its unit-test constructor takes it (TemplateLengthStatistics.hh:66-81);
 * models are the AlignmentModel enum values FFp=0 FRp=1 RFp=2 RRp=3 FFm=4 FRm=5 RFm=6 RRm=7 (:48-59). */
typedef struct isaac_ext_tls {
    uint32_t min, max, median, lowStdDev, highStdDev;
    uint32_t bestModel[2];
    int32_t  mateDriftRange;         /* -1: mateMin = min, mateMax = max (TemplateLengthStatistics.hh:205-214) */
} isaac_ext_tls_t;

/* MatchSelector::determineTemplateLength for the resident tile (MatchSelector.cpp:188-249): FragmentBuilder::build without gaps
 * for every cluster (one GPU pass), then TemplateLengthDistribution::addTemplate cluster by cluster in match-list order until
 * the statistics are stable, or finalize() at the end of the tile (TemplateLengthStatistics.cpp:95-160,266-357).  pf: one byte
 * per cluster (BclClusters::pf, only passing clusters are used) or NULL = all pass.  batch->withGaps is ignored.  Single-ended
 * read sets give the cleared statistics (all fields -1U, models 8 = InvalidAlignmentModel), like the reference.
 * With several GPUs the rank that owns the first tile calls this and broadcasts the eight words (MatchSelector.cpp:401-417). */
int isaac_ext_determine_template_length(isaac_ext_ctx *ctx, const struct isaac_ext_build_batch *batch, const uint8_t *pf,
                                        int32_t mateDriftRange, struct isaac_ext_tls *tlsOut, uint32_t *stableOut);

/* One ShadowAligner::rescueShadow call (ShadowAligner.hh:81-88): the orphan fields the call reads. */
typedef struct isaac_ext_rescue_request {
    int64_t  orphanPosition;
    int64_t  bestTemplateLength;     /* 0 = no best template                                            */
    uint32_t orphanReadId;           /* cluster * readCount + readIndex of the ORPHAN; the mate is rescued */
    uint32_t orphanContigStrand;     /* contigId << 1 | reverse                                         */
    uint32_t orphanObservedLength;
    uint32_t pad;
} isaac_ext_rescue_request_t;

/* shadowList + getCigarBuffer() of every request, flattened like isaac_ext_build_result_t; rescued[i] is the return
 * value of rescueShadow (the list can be non-empty when it is 0, ShadowAligner.cpp:151-154). */
typedef struct isaac_ext_rescue_result {
    const isaac_ext_fragment_t *fragments;
    const uint64_t *requestFragmentBegin;   /* requestCount + 1 */
    const uint32_t *cigars;
    const uint8_t *rescued;
    uint64_t fragmentCount;
    uint64_t cigarWords;
} isaac_ext_rescue_result_t;

/* What alignment::TemplateBuilder is constructed / called with beyond isaac_ext_config_t (TemplateBuilder.hh:68-82,
 * MatchSelector.cpp:330-334). */
typedef struct isaac_ext_template_options {
    uint32_t scatterRepeats;         /* --scatter-repeats                                                        */
    int32_t  dodgyAlignmentScore;    /* -1 = DODGY_ALIGNMENT_SCORE_UNALIGNED, 255 = _UNKNOWN, else numeric (:60-61) */
    uint32_t mapqThreshold;          /* --mapq-threshold                                                          */
    uint32_t clipFlags;              /* ISAAC_EXT_CLIP_* applied to every template that is kept (MatchSelector.cpp:336-346) */
} isaac_ext_template_options_t;
#define ISAAC_EXT_CLIP_SEMIALIGNED 1u    /* --clip-semialigned: matchSelector::SemialignedEndsClipper               */
#define ISAAC_EXT_CLIP_OVERLAPPING 2u    /* --clip-overlapping: matchSelector::OverlappingEndsClipper               */

/* alignment::BamTemplate of one cluster (BamTemplate.hh:40-137) next to its two fragment records. */
typedef struct isaac_ext_template {
    uint32_t alignmentScore;             /* BamTemplate::getAlignmentScore(), 0xFFFFFFFF = unknown (-1U)          */
    uint32_t fragmentAlignmentScore[2];  /* FragmentMetadata::alignmentScore of read 1 / read 2                   */
    uint8_t  properPair;
    uint8_t  built;                      /* buildFragments() && buildTemplate() (MatchSelector.cpp:323-334)       */
    uint8_t  hadFragments;               /* buildFragments()                                                       */
    uint8_t  pad;
} isaac_ext_template_t;

/* Owned by the context, valid until its next call.  fragments[cluster * readCount + readIndex] =
 * BamTemplate::getFragmentMetadata(readIndex); unaligned / no-match records keep the reference's field values
 * (contigId 0x7FFFFF = ReferencePosition::MAX_CONTIG_ID for "no match", FragmentMetadata.hh:258-260). */
typedef struct isaac_ext_template_result {
    const isaac_ext_template_t *templates;   /* clusterCount                                                  */
    const isaac_ext_fragment_t *fragments;   /* clusterCount * readCount                                      */
    const uint32_t *cigars;
    uint64_t cigarWords;
    uint64_t rescueRequests;                 /* ShadowAligner::rescueShadow calls the tile needed             */
} isaac_ext_template_result_t;

/* ---- life cycle ----------------------------------------------------------------------------- */
/* One context per GPU stands for what MatchSelector keeps per compute thread: a TemplateBuilder with its FragmentBuilder and
 * ShadowAligner (MatchSelector.cpp:143-163; constructor arguments FragmentBuilder.hh:49-60, ShadowAligner.hh:52-59).
 * isaac_ext_create fails with ISAAC_EXT_E_INVALID_ARG where the BandedSmithWaterman constructor throws
 * common::InvalidParameterException (BandedSmithWaterman.cpp:47-53). */
int  isaac_ext_create(const isaac_ext_config_t *config, isaac_ext_ctx **ctx);
void isaac_ext_destroy(isaac_ext_ctx *ctx);
const char *isaac_ext_last_error(const isaac_ext_ctx *ctx);   /* ctx may be NULL: last create error */
const char *isaac_ext_version(void);

/* Replaces reference::ContigLoader's product (ContigLoader.cpp:29-65): contigs arrive as 1 byte/base
 * upper-case ACGTN and are packed to 2 bit + N-mask and kept resident in HBM. */
int isaac_ext_set_reference(isaac_ext_ctx *ctx, uint32_t contigCount,
                            const char *const *contigBases, const uint64_t *contigLengths);

/* flowcell::SequencingAdapterMetadata (SequencingAdapterMetadata.hh:36-72). */
typedef struct isaac_ext_adapter {
    const char *sequence;            /* in the direction of the reference, upper-case ACGT, 5..126 bases            */
    uint32_t reverse;                /* isReverse(): the direction the adapter is sequenced in                      */
    uint32_t clipLength;             /* getClipLength(): 0 = unbounded (clip to the end of the read)               */
} isaac_ext_adapter_t;

/* The matchSelector::SequencingAdapterList every later call clips with: what FragmentBuilder::build / rescueShadow take as
 * 'sequencingAdapters' (FragmentBuilder.hh:62-70, ShadowAligner.hh:81-88) and hand to
 * matchSelector::FragmentSequencingAdapterClipper (FragmentSequencingAdapterClipper.cpp:102-277, SequencingAdapter.cpp:30-139).
 * The tile calls keep one clipper per read list / rescue like the reference (the first candidate of a strand locates the
 * adapter); the micro entry points treat every candidate as its own clipper.  count = 0 removes the adapters (the default:
 * empty list, DefaultAdaptersOption).  One deviation: decideWhichSideToClip reads the contig without a bounds check when
 * the adapter reaches one end of the read (:190-216); here bases outside the contig count as mismatches. */
int isaac_ext_set_adapters(isaac_ext_ctx *ctx, uint32_t count, const isaac_ext_adapter_t *adapters);

/* Decodes one tile's BCL bytes (Read::decodeBcl, Read.cpp:32-73) into the device-resident read set used
 * by the batch calls below.  Stays valid until the next isaac_ext_set_reads on this context. */
int isaac_ext_set_reads(isaac_ext_ctx *ctx, const isaac_ext_reads_t *reads);

/* Double buffering of tiles (SelectMatchesTransition.cpp:316-340 loads the next tile while the current one is processed): starts
 * the upload and decode of the NEXT tile's BCL bytes into the context's second read-set slot on its own stream and returns at
 * once; the calls on the current tile go on next to it (copy engine + a few SMs).  A later isaac_ext_set_reads with the same
 * description (same pointers, counts, lengths) takes the prefetched slot over instead of uploading again; any other
 * isaac_ext_set_reads simply uploads as usual.  reads->bcl (page-locked memory, for the copy to overlap) and
 * reads->endCyclesMasked must stay unchanged until that isaac_ext_set_reads.  One prefetch at a time. */
int isaac_ext_prefetch_reads(isaac_ext_ctx *ctx, const isaac_ext_reads_t *reads);
/* The same for the NEXT tile's seed matches (clusterCount = the clusters of that tile), after its isaac_ext_prefetch_reads: the
 * isaac_ext_set_reads that takes the prefetched reads over takes the matches with them, and the first isaac_ext_build_fragments /
 * isaac_ext_build_templates / isaac_ext_determine_template_length on that tile with the same batch (same pointers) finds them on
 * the device.  With both prefetches a tile starts computing at once: its inputs went up under the previous tile's kernels. */
int isaac_ext_prefetch_batch(isaac_ext_ctx *ctx, const struct isaac_ext_build_batch *batch, uint32_t clusterCount);

/* alignment::trimLowQualityEnds (Quality.cpp:71-120, --base-quality-cutoff, MatchSelector.cpp:300): recomputes
 * Read::endCyclesMasked_ of every resident read from its qualities, replacing what isaac_ext_set_reads was given.
 * baseQualityCutoff 0 = no masking.  endCyclesMaskedOut (clusterCount * readCount values) may be NULL. */
int isaac_ext_trim_low_quality_ends(isaac_ext_ctx *ctx, uint32_t baseQualityCutoff, uint16_t *endCyclesMaskedOut);

/* ---- the two calls TemplateBuilder makes --------------------------------------------------------- */

/* FragmentBuilder::build over the resident read set: seeds -> candidates -> ungapped -> simple indels -> gapped, with
 * the reference's consolidation and acceptance rules (FragmentBuilder.cpp:82-324).  The per-cluster bookkeeping
 * (candidate lists, std::sort + consolidate, acceptance) runs on config.hostThreads host threads between the kernel
 * passes; every base comparison, score and Smith-Waterman cell runs on the GPU. */
int isaac_ext_build_fragments(isaac_ext_ctx *ctx, const isaac_ext_build_batch_t *batch, isaac_ext_build_result_t *result);

/* ShadowAligner::rescueShadow for every request against the resident read set (ShadowAligner.cpp:155-291). */
int isaac_ext_rescue_shadows(isaac_ext_ctx *ctx, const isaac_ext_tls_t *tls, uint32_t requestCount,
                             const isaac_ext_rescue_request_t *requests, isaac_ext_rescue_result_t *result);

/* SURVEY 8(f) #1, the caller of the two: TemplateBuilder::buildFragments + buildTemplate for every cluster of the resident
 * read set (MatchSelector.cpp:323-334; TemplateBuilder.cpp:97-1089: pickBestPair / locateBestPair /
 * buildPairedEndTemplate / buildDisjoinedTemplate / rescueShadow / scoreDisjoinedTemplate / updateMappingScore, MAPQ
 * threshold filter).  One isaac_ext_build_fragments pass, the rescue requests the templates need in one
 * isaac_ext_rescue_shadows pass, then pair selection and mapping scores per cluster on the host threads.  The
 * rest-of-genome correction is computed from the resident reference and read lengths (RestOfGenomeCorrection.hh:45-57).
 * With options.clipFlags the kept templates then go through the end clippers (SemialignedEndsClipper.cpp:32-205,
 * OverlappingEndsClipper.cpp:45-180) in one kernel pass. */
int isaac_ext_build_templates(isaac_ext_ctx *ctx, const isaac_ext_build_batch_t *batch, const isaac_ext_tls_t *tls,
                              const isaac_ext_template_options_t *options, isaac_ext_template_result_t *result);

/* isaac_ext_build_templates whose result comes back on a copy stream of its own: the call returns once the tile's kernels are
 * queued (it still waits for the handful of totals that size its passes), result points at one of two host result sets of the
 * context and is VALID ONLY AFTER isaac_ext_fetch_templates(ctx, result) has returned ISAAC_EXT_OK -- which also is where a malformed
 * match batch or a capacity error of that tile is reported.  What the caller gains: the download of tile k (≈ 150 B per pair) runs
 * next to the kernels of tile k + 1 instead of in front of them,
 *     set_reads(k); build_templates_deferred(k, &r[k % 2]); set_reads(k + 1); build_templates_deferred(k + 1, &r[(k + 1) % 2]);
 *     fetch_templates(&r[k % 2]); ... use tile k ...
 * At most two tiles may be waiting to be fetched (a third deferred call returns ISAAC_EXT_E_UNSUPPORTED); a result set is reused by
 * the deferred call after the next one, so tile k must have been consumed before tile k + 2 is built.  The device-resident result
 * (isaac_ext_template_stats, isaac_ext_tile_cycle_stats, isaac_ext_pack_fragments) is that of the last tile built, as always. */
int isaac_ext_build_templates_deferred(isaac_ext_ctx *ctx, const isaac_ext_build_batch_t *batch, const isaac_ext_tls_t *tls,
                                       const isaac_ext_template_options_t *options, isaac_ext_template_result_t *result);
int isaac_ext_fetch_templates(isaac_ext_ctx *ctx, const isaac_ext_template_result_t *result);

/* The three tile calls above without blocking the caller: a submitted call runs on a worker thread of the context, on the
 * context's own streams and buffers, while the caller loads and sorts the next tile's matches, the way the reference runs its
 * load / compute / flush slots side by side (SelectMatchesTransition.cpp:316-340).  The small argument structs are copied at
 * submit; the arrays they point to (matches, offsets, seeds, requests) must stay valid and unchanged until the wait.  One call is
 * in flight per context (a context is no more re-entrant than the reference's per-thread builders, MatchSelector.cpp:143-163):
 * until the wait, a second submit AND every blocking entry point of the context (isaac_ext_set_reads included) return
 * ISAAC_EXT_E_UNSUPPORTED without touching the context -- all but isaac_ext_prefetch_reads / isaac_ext_prefetch_batch, which stage
 * the next tile into the standby slots on their own stream, and isaac_ext_last_error / isaac_ext_version.  isaac_ext_wait blocks,
 * returns the status of the call and fills *resultOut, which is the result struct of the call that was submitted
 * (isaac_ext_build_result_t / isaac_ext_rescue_result_t / isaac_ext_template_result_t), valid until the next call on the context. */
int isaac_ext_submit_build_fragments(isaac_ext_ctx *ctx, const isaac_ext_build_batch_t *batch, uint64_t *ticketOut);
int isaac_ext_submit_rescue_shadows(isaac_ext_ctx *ctx, const isaac_ext_tls_t *tls, uint32_t requestCount,
                                    const isaac_ext_rescue_request_t *requests, uint64_t *ticketOut);
int isaac_ext_submit_build_templates(isaac_ext_ctx *ctx, const isaac_ext_build_batch_t *batch, const isaac_ext_tls_t *tls,
                                     const isaac_ext_template_options_t *options, uint64_t *ticketOut);
int isaac_ext_wait(isaac_ext_ctx *ctx, uint64_t ticket, void *resultOut);

/* ---- micro entry points (unit parity + kernel benchmarks) ------------------------------------ */

/* BandedSmithWaterman::align over n independent (query, database) pairs given as ASCII, database i is
 * queryLength[i] + 15 bytes (BandedSmithWaterman.cpp:93).  Scores as the BandedSmithWaterman constructor
 * takes them (gapOpen/gapExtend positive, BandedSmithWaterman.hh:44-46).  Query alphabet ACGTn, database
 * ACGTN.  cigarOut: n * cigarStride words, cigarLengthOut/offsetOut: n.  offsetOut = return value of
 * align() (stripped leading deletion). */
int isaac_ext_banded_sw_batch(isaac_ext_ctx *ctx, uint32_t n,
                              const char *queries, const uint64_t *queryOffsets, const uint32_t *queryLengths,
                              const char *databases, const uint64_t *databaseOffsets,
                              int matchScore, int mismatchScore, int gapOpenScore, int gapExtendScore,
                              uint32_t cigarStride, uint32_t *cigarOut, uint32_t *cigarLengthOut,
                              uint32_t *offsetOut);

/* UngappedAligner::alignUngapped (UngappedAligner.cpp:39-92) on every candidate against the resident
 * reference and read set.  fragmentsOut: n records; cigarOut: n * 3 words (fragment.cigarOffset = 3 * i);
 * mismatchMaskOut: n * ISAAC_EXT_MASK_WORDS words or NULL. */
int isaac_ext_ungapped_batch(isaac_ext_ctx *ctx, uint32_t n, const isaac_ext_candidate_t *candidates,
                             isaac_ext_fragment_t *fragmentsOut, uint32_t *cigarOut, uint64_t *mismatchMaskOut);

/* GappedAligner::alignGapped (GappedAligner.cpp:167-249) on every candidate.  matchCount == 0 and
 * cigarLength == 0 where the reference returns 0 without aligning (GappedAligner.cpp:204-208).
 * cigarOut: n * cigarStride words (fragment.cigarOffset = cigarStride * i). */
int isaac_ext_gapped_batch(isaac_ext_ctx *ctx, uint32_t n, const isaac_ext_candidate_t *candidates,
                           uint32_t cigarStride, isaac_ext_fragment_t *fragmentsOut, uint32_t *cigarOut,
                           uint64_t *mismatchMaskOut);

/* End-to-end variants of the two calls above (UngappedAligner.cpp:39-92, GappedAligner.cpp:167-249) for large batches: same
 * semantics, but the CIGARs come back as a DENSE pool
 * (fragment.cigarOffset indexes cigarPoolOut; *cigarWordsOut = words used) and the batch is processed in chunks whose
 * host-to-device copy, kernels and device-to-host copies overlap.  Pass page-locked host buffers for full copy speed.
 * ISAAC_EXT_E_CAPACITY if cigarPoolCapacity (in words) is too small; *cigarWordsOut then holds the required size. */
int isaac_ext_ungapped_batch_compact(isaac_ext_ctx *ctx, uint32_t n, const isaac_ext_candidate_t *candidates,
                                     isaac_ext_fragment_t *fragmentsOut, uint32_t *cigarPoolOut, uint64_t cigarPoolCapacity,
                                     uint64_t *cigarWordsOut);
int isaac_ext_gapped_batch_compact(isaac_ext_ctx *ctx, uint32_t n, const isaac_ext_candidate_t *candidates,
                                   isaac_ext_fragment_t *fragmentsOut, uint32_t *cigarPoolOut, uint64_t cigarPoolCapacity,
                                   uint64_t *cigarWordsOut);

/* Both calls above over the same candidates in one pass over the batch: what FragmentBuilder::alignFragments does with a
 * candidate (alignUngapped, then alignGapped on a copy, FragmentBuilder.cpp:174,199-200), here for every candidate.  The
 * candidates are uploaded once and the ungapped records of a chunk travel back while its Smith-Waterman runs. */
int isaac_ext_extend_batch_compact(isaac_ext_ctx *ctx, uint32_t n, const isaac_ext_candidate_t *candidates,
                                   isaac_ext_fragment_t *ungappedOut, uint32_t *ungappedPoolOut, uint64_t ungappedPoolCapacity,
                                   uint64_t *ungappedWordsOut, isaac_ext_fragment_t *gappedOut, uint32_t *gappedPoolOut,
                                   uint64_t gappedPoolCapacity, uint64_t *gappedWordsOut);

/* What FragmentBuilder::alignFragments keeps of one candidate (FragmentBuilder.cpp:174-209), 32 bytes: the ungapped alignment, or
 * the gapped one where the reference would have run the gapped aligner (more than BandedSmithWaterman::mismatchesCutoff = 5
 * mismatches, :190-200) and its 5-clause rule accepts it (:202-209).  cigarLength = the alignment's words in the pool, one
 * alignment after the other in candidate order.  An aligned record with cigarLength 0 has the CIGAR its clip counts imply --
 * [leftClip S] [readLength - leftClip - rightClip M] [rightClip S], leftClip = reverse ? highClipped : lowClipped, rightClip the
 * other one (UngappedAligner.cpp:64-81, FragmentMetadata::incrementClipLeft / Right) -- which is nearly every kept ungapped
 * alignment (the exception: soft clips at a contig end, which the reference does not count in lowClipped / highClipped). */
typedef struct isaac_ext_alignment {
    int64_t  position;
    double   logProbability;
    uint16_t observedLength;
    uint16_t mismatchCount;
    uint16_t matchesInARow;
    uint16_t editDistance;
    uint16_t smithWatermanScore;
    uint16_t lowClipped;
    uint16_t highClipped;
    uint8_t  gapsAndFlags;           /* gapCount | ISAAC_EXT_ALIGNMENT_ALIGNED | ISAAC_EXT_ALIGNMENT_GAPPED */
    uint8_t  cigarLength;            /* words in the pool; 0 with _ALIGNED set: the implied CIGAR                     */
} isaac_ext_alignment_t;
#define ISAAC_EXT_ALIGNMENT_GAPS    0x3Fu
#define ISAAC_EXT_ALIGNMENT_ALIGNED 0x40u   /* FragmentMetadata::isAligned of the kept alignment */
#define ISAAC_EXT_ALIGNMENT_GAPPED  0x80u   /* the gapped alignment was accepted                 */

/* alignUngapped + alignGapped + the acceptance rule for every candidate, chunked and overlapped like the calls above, with
 * ONE 32-byte record per candidate on the way back (a quarter of the bytes of isaac_ext_extend_batch_compact: the copy back to the
 * host is what bounds these calls, and on a box with several GPUs the host's ingest is shared).  The gapped aligner runs on every
 * candidate (BASELINE configs[1] counts its cells that way); its result is only taken where the reference would have run it.
 * ISAAC_EXT_E_UNSUPPORTED if a Smith-Waterman score does not fit 16 bits (scores far beyond the presets). */
int isaac_ext_align_batch_packed(isaac_ext_ctx *ctx, uint32_t n, const isaac_ext_candidate_t *candidates,
                                 isaac_ext_alignment_t *alignmentsOut, uint32_t *cigarPoolOut, uint64_t cigarPoolCapacity,
                                 uint64_t *cigarWordsOut);

/* Device-resident variants of isaac_ext_ungapped_batch / isaac_ext_gapped_batch (UngappedAligner::alignUngapped,
 * UngappedAligner.cpp:39-92; GappedAligner::alignGapped, GappedAligner.cpp:167-249) used to time the kernels alone: the
 * candidate and result arrays are device pointers, the launch goes to 'cudaStream' (a cudaStream_t passed as void*) and returns
 * without synchronising.  Same semantics as the host variants above. */
int isaac_ext_ungapped_batch_device(isaac_ext_ctx *ctx, uint32_t n, const void *dCandidates,
                                    void *dFragmentsOut, void *dCigarOut, void *dMismatchMaskOut,
                                    void *cudaStream);
int isaac_ext_gapped_batch_device(isaac_ext_ctx *ctx, uint32_t n, const void *dCandidates,
                                  uint32_t cigarStride, void *dFragmentsOut, void *dCigarOut,
                                  void *dMismatchMaskOut, void *cudaStream);

/* Per-tile statistics (K6).  Adds the counters of n fragment records that live in DEVICE memory to the 64 u64
 * counters at dStats (device memory, caller zeroes them) on 'cudaStream': [0] records, [1] aligned, [2] with gaps,
 * [3] edit distance 0, [4] mismatches, [5] edit distance, [6] gaps, [7] observed bases, [8..40] histogram of the
 * mismatch count clipped at 32.  The ranks of a multi-GPU run sum these vectors with one NCCL all-reduce, the only
 * collective of the path (the reference sums per-thread TileStats the same way, MatchSelector.cpp:439-442). */
#define ISAAC_EXT_STATS_COUNTERS 64
int isaac_ext_tile_stats_device(isaac_ext_ctx *ctx, uint32_t n, const void *dFragments, void *dStats, void *cudaStream);

/* matchSelector::TileBarcodeStats of MatchSelectorStats (TileBarcodeStats.hh:40-160, MatchSelectorStats.hh:77-103) for the
 * templates of one tile (the result of isaac_ext_build_templates, host pointers): one kernel pass over the tile on the GPU.
 * statsOut: 4 blocks of ISAAC_EXT_TEMPLATE_STATS_COUNTERS u64, block = readIndex * 2 + passesFilter (the pass-filter block
 * counts pf clusters only, the other one all clusters; pair-level counters live under read index 0).  Counters of a block:
 * [0] yield [1] yieldQ30 [2] qualityScoreSum [3] clusterCount [4] unanchoredClusterCount [5] nmnmClusterCount
 * [6] rmClusterCount [7] qcClusterCount [8] alignedFragmentCount [9] uniquelyAlignedFragmentCount
 * [10] uniquelyAlignedPerfectFragmentCount [11] alignmentScoreSum [12] basesOutsideIndels [13] uniquelyAlignedBasesOutsideIndels
 * [14] mismatches [15] uniquelyAlignedMismatches [16..24] alignmentModelCounts (FFp..RRm, Invalid) [25..28] nominalModelCounts
 * (Oversized, Undersized, Nominal, NoMatch) [29] fragmentCount.  The template type of a cluster follows
 * MatchSelector.cpp:300-365: no match list or a NoMatch record first = NmNm (Qc if that record carries the N-seed id), match
 * list whose fragments did not build = Rm.  pf: one byte per cluster or NULL = all pass.  The ranks of a multi-GPU run sum
 * these vectors with one all-reduce (the reference sums its per-thread MatchSelectorStats, MatchSelector.cpp:439-442). */
#define ISAAC_EXT_TEMPLATE_STATS_COUNTERS 32
int isaac_ext_template_stats(isaac_ext_ctx *ctx, const struct isaac_ext_build_batch *batch, const struct isaac_ext_tls *tls,
                             const struct isaac_ext_template_result *templates, const uint8_t *pf, uint64_t *statsOut);

/* ---- io::FragmentHeader bin records ------------------------------------------------------------------ */

/* What matchSelector::FragmentCollector::add reads beyond the template itself (FragmentCollector.cpp:42-77): the Cluster
 * fields MatchSelector::processMatchList initialises (MatchSelector.cpp:301-302) and the BinIndexMap of the run
 * (BinIndexMap.hh:45-107, handed over as its per-contig vectors). */
typedef struct isaac_ext_pack_options {
    uint64_t tile;                    /* Cluster::getTile of the resident tile                                            */
    uint32_t barcodeIdx;              /* the barcode index FragmentCollector::add is called with                          */
    uint32_t keepUnaligned;           /* --keep-unaligned: store the templates that did not build too
                                         (MatchSelector.cpp:311-314,331,355-358); else only templates with 'built'       */
    uint32_t compact;                 /* 0: FragmentBuffer layout (fixed-size slots); 1: the stored records back to back, each
                                         FragmentHeader::getTotalLength() bytes = what BufferingFragmentStorage::flush writes
                                         to a bin file per fragment (BufferingFragmentStorage.cpp:108-116,184-188)           */
    uint32_t pad;
    const uint8_t  *pf;               /* Cluster::getPf per cluster, or NULL = every cluster passes                       */
    const int32_t  *xy;               /* Cluster::getXy: x, y per cluster, or NULL = unset (POSITION_NOT_SET)             */
    const uint64_t *barcodeSequence;  /* Cluster::getBarcodeSequence per cluster, or NULL = 0                             */
    uint32_t distributionBinSize;     /* MatchDistribution::getBinSize; 0 = no bin map, mateStorageBin_ stays 0           */
    uint32_t contigCount;             /* contigs the bin map covers                                                       */
    const uint64_t *contigBinBegin;   /* contigCount + 1 offsets into binIndex: BinIndexMap::at(contigId + 1)             */
    const uint32_t *binIndex;
} isaac_ext_pack_options_t;

/* matchSelector::FragmentBuffer after the tile (FragmentCollector.hh:44-310), owned by the context, valid until its next
 * isaac_ext_pack_fragments.  records = data_: clusterCount fixed-size records, the fragment of (cluster, readIndex) at
 * cluster * recordLength + readOffset[readIndex] as io::FragmentHeader (headerLength = sizeof = 112 bytes; the padding bytes 42..47,
 * 108..111 and the six spare bits of flags_ are zero here, unspecified in the reference) followed by readLength BCL bytes
 * (reverse-complemented for reverse fragments) and cigarLength CIGAR words; everything behind is zero, like the reference's
 * freshly resized buffer.  fStrandPos / initialized = index_: IndexRecord::fStrandPos_ (ReferencePosition::getValue) and
 * IndexRecord::initialized() per cluster * readCount + readIndex.  With options.compact the records lie back to back in
 * (cluster, readIndex) order at recordOffset[], each cut to its FragmentHeader::getTotalLength(). */
typedef struct isaac_ext_pack_result {
    const uint8_t  *records;
    const uint64_t *fStrandPos;
    const uint8_t  *initialized;
    uint32_t recordLength;            /* FragmentBuffer::getRecordLength (FragmentCollector.hh:283-289)                   */
    uint32_t readOffset[2];           /* FragmentBuffer::getReadOffsets (:295-308)                                        */
    uint32_t headerLength;
    uint64_t storedFragments;         /* records initialised by this call                                                 */
    const uint64_t *recordOffset;     /* clusterCount * readCount + 1: byte offset of every record in 'records' (compact: a
                                         record that is not stored has length 0)                                           */
    uint64_t recordBytes;             /* bytes at 'records'                                                               */
    float    kernelMs;                /* duration of the pack kernel alone (CUDA events on the context's stream), for bench.py */
    uint32_t pad;
} isaac_ext_pack_result_t;

/* FragmentCollector::add for every fragment of every template of the resident tile that MatchSelector::processMatchList stores
 * (BufferingFragmentStorage::add, BufferingFragmentStorage.hh:64-70): io::FragmentHeader's paired / single-ended constructors
 * (Fragment.hh:100-186: TLEN :209-237, mate anchor :489-497, duplicate rank :66-71), storeBclAndCigar
 * (FragmentCollector.cpp:79-103).  templates = the result of isaac_ext_build_templates (host pointers); the BCL bytes are the
 * resident read set's.  One kernel pass, one warp per cluster; the records travel back in one copy. */
int isaac_ext_pack_fragments(isaac_ext_ctx *ctx, const struct isaac_ext_template_result *templates,
                             const isaac_ext_pack_options_t *options, isaac_ext_pack_result_t *result);

/* matchSelector::TileStats of MatchSelectorStats (TileStats.hh:68-142, recorded like MatchSelectorStats.hh:77-103) for the templates
 * the last isaac_ext_build_templates / isaac_ext_select_tile of this context left on the device (no other tile call in between): the
 * alignment score histograms of fragments and templates and the per-cycle arrays (blanks, mismatches, fragments with 1 / 2 / 3 /
 * 4 / 5 mismatches so far; each again for uniquely aligned fragments), one kernel pass.  statsOut: 4 blocks of
 * ISAAC_EXT_TILE_CYCLE_STATS_WORDS u64, block = readIndex * 2 + passesFilter, each laid out like the struct:
 * alignmentScoreFragments_[8192], alignmentScoreMismatches_[8192], alignmentScoreTemplates_[8192],
 * alignmentScoreTemplateMismatches_[8192], cycleBlanks_[1024], cycleUniquelyAlignedBlanks_, cycleMismatches_,
 * cycleUniquelyAlignedMismatches_, cycleUniquelyAligned{1,2,3,4,More}MismatchFragments_, cycle{1,2,3,4,More}MismatchFragments_,
 * uniquelyAlignedFragmentCount_ -- the raw counters, which the ranks of a multi-GPU run sum with one all-reduce (the reference sums
 * its per-thread TileStats, MatchSelector.cpp:439-442, TileStats.hh:144-237); isaac_ext_tile_cycle_stats_finalize then applies
 * TileStats::finalize (:239-340) to one block.  A mismatch cycle is the reference's: the cycles of the alignment the template took
 * the fragment from, cut to the fragment's final mismatch count (the end clippers and the low-quality filter leave the list alone). */
#define ISAAC_EXT_TILE_CYCLE_STATS_WORDS 47105
int isaac_ext_tile_cycle_stats(isaac_ext_ctx *ctx, const uint8_t *pf, uint64_t *statsOut);
void isaac_ext_tile_cycle_stats_finalize(uint64_t *statsBlock);

/* ---- one tile, the way MatchSelector::parallelSelect drives it (MatchSelector.cpp:370-443) ------------------------------ */
/* Inputs as isaac-align leaves them for the match selector: the tile's BclClusters buffer and its match records (raw 16-byte
 * alignment::Match as io::MatchWriter writes them, sorted by cluster / location / seed, SelectMatchesTransition.cpp:242-254;
 * clusters are delimited by the cluster field of the seed ids like findNextCluster does, MatchSelector.cpp:262-277). */
typedef struct isaac_ext_tile {
    isaac_ext_reads_t reads;
    const isaac_ext_match_t *matches;
    uint64_t matchCount;
    const isaac_ext_seed_t *seeds;
    uint32_t seedCount;
    uint32_t withGaps;
    const uint8_t *pf;                    /* BclClusters::pf per cluster, or NULL = all pass                                  */
    uint32_t baseQualityCutoff;           /* --base-quality-cutoff, 0 = none (MatchSelector.cpp:300)                          */
    int32_t  mateDriftRange;
    const isaac_ext_tls_t *tls;           /* user-defined / earlier tile's template length statistics, or NULL = determine them
                                             from this tile (MatchSelector.cpp:401-417)                                       */
    isaac_ext_template_options_t options;
    const struct isaac_ext_pack_options *pack;   /* non-NULL: also leave the io::FragmentHeader records of the tile          */
    uint32_t cycleStats;                  /* non-zero: also the TileStats of the tile (isaac_ext_tile_cycle_stats)             */
    uint32_t pad;
} isaac_ext_tile_t;

typedef struct isaac_ext_tile_result {
    isaac_ext_template_result_t templates;       /* owned by the context, valid until its next call                           */
    isaac_ext_tls_t tls;                         /* the statistics the templates were built with                              */
    uint32_t tlsStable;                          /* 1 when given by the caller or stable within the tile                      */
    uint32_t packedValid;
    uint64_t stats[4 * 32];                      /* isaac_ext_template_stats of the tile                                      */
    const uint16_t *endCyclesMasked;             /* Read::endCyclesMasked_ after quality trimming, or NULL (no cutoff)        */
    const uint64_t *cycleStats;                  /* 4 * ISAAC_EXT_TILE_CYCLE_STATS_WORDS raw counters, or NULL                */
} isaac_ext_tile_result_t;

/* set_reads -> trim_low_quality_ends -> determine_template_length (unless given) -> build_templates -> template_stats
 * (-> tile_cycle_stats) (-> pack_fragments, result through isaac_ext_tile_packed): every step is a GPU pass of this library.  Tiles are independent once
 * the statistics are fixed: with several GPUs rank r runs tiles r, r + G, ... and the callers sum result.stats. */
int isaac_ext_select_tile(isaac_ext_ctx *ctx, const isaac_ext_tile_t *tile, isaac_ext_tile_result_t *result);
/* the io::FragmentHeader records the last isaac_ext_select_tile with tile.pack left (see isaac_ext_pack_fragments) */
int isaac_ext_tile_packed(isaac_ext_ctx *ctx, struct isaac_ext_pack_result *packedOut);

/* BandedSmithWaterman::align on a band of bandWidth lanes as a warp-level wavefront (band lanes across the lanes of a warp,
 * __shfl_sync between neighbouring band lanes, direction planes in shared memory): the variant for long reads and the widened band
 * of BASELINE configs[4] (2x250 bp, bandWidth 32).  Same arguments and results as isaac_ext_banded_sw_batch; databases hold
 * queryLength + bandWidth - 1 bases.  bandWidth 16 is the reference's band (bit-exact with BandedSmithWaterman.cpp:84-462);
 * bandWidth 32 has no reference behaviour (the reference hard-wires 16 lanes, BandedSmithWaterman.hh:88-89): it is bit-exact
 * with the band-width-parametrised model of oracle/ whose 16-lane instance equals the reference.  The _device variant takes
 * device pointers and a stream and returns without synchronising. */
int isaac_ext_banded_sw_wide_batch(isaac_ext_ctx *ctx, uint32_t bandWidth, uint32_t n, const char *queries, const uint64_t *queryOffsets,
                                   const uint32_t *queryLengths, const char *databases, const uint64_t *databaseOffsets,
                                   int matchScore, int mismatchScore, int gapOpenScore, int gapExtendScore,
                                   uint32_t cigarStride, uint32_t *cigarOut, uint32_t *cigarLengthOut, uint32_t *offsetOut);
int isaac_ext_banded_sw_wide_batch_device(isaac_ext_ctx *ctx, uint32_t bandWidth, uint32_t n, const void *dQueries,
                                          const void *dQueryOffsets, const void *dQueryLengths, const void *dDatabases,
                                          const void *dDatabaseOffsets, uint32_t maxQueryLength, int matchScore, int mismatchScore,
                                          int gapOpenScore, int gapExtendScore, uint32_t cigarStride, void *dCigarOut,
                                          void *dCigarLengthOut, void *dOffsetOut, void *cudaStream);

/* ---- SURVEY 8(f) #4, last part: build::GapRealigner over one bin -------------------------------------------------------- */
/* The reference realigns bin by bin (BinSorter::process, BinSorter.hh:161-175): collectGaps walks every record of the bin's data and
 * gathers the insertions and deletions of their CIGARs per gap group (BinSorter.cpp:389-405, RealignerGaps::addGaps
 * GapRealigner.hh:52-97, finalizeGaps GapRealigner.cpp:86-94), then realignGaps calls GapRealigner::realign for every entry of the
 * bin's index in index order (BinSorter.cpp:407-418, GapRealigner.cpp:1061-1267).  A bin here is what the reference holds at that
 * point: the io::FragmentHeader records back to back (the layout isaac_ext_pack_fragments leaves with options.compact, = the bin
 * file) and PackedFragmentBuffer::Index without its pointers. */
typedef struct isaac_ext_bin_index {
    uint64_t dataOffset;              /* PackedFragmentBuffer::Index::dataOffset_: byte offset of the record in the bin's data  */
    uint64_t mateDataOffset;          /* ::mateDataOffset_; = dataOffset for single-ended records and mates stored in another bin */
} isaac_ext_bin_index_t;

/* gapRealigner::Gap (Gap.hh:32-80) */
typedef struct isaac_ext_gap {
    uint64_t position;                /* ReferencePosition::getValue of Gap::pos_                                               */
    int32_t  length;                  /* > 0 deletion from the reference, < 0 insertion                                         */
    uint32_t group;                   /* gap group (BinSorter::getGapGroupIndex, BinSorter.cpp:355-370)                         */
} isaac_ext_gap_t;

typedef struct isaac_ext_realign_options {
    uint64_t binStart, binEnd;        /* ReferencePosition::getValue of BinMetadata::getBinStart / getBinEnd                    */
    uint32_t realignGapsVigorously;   /* --realign-vigorously                                                                  */
    uint32_t realignDodgyFragments;   /* --realign-dodgy                                                                       */
    uint32_t mismatchCost, gapOpenCost, gapExtendCost;   /* BinSorter constructs the realigner with 3, 4, 0 (BinSorter.hh:97)  */
    uint32_t clipSemialigned;         /* --clip-semialigned: build::SemialignedEndsClipper after a realignment                 */
    uint32_t barcodeCount;            /* barcodes the records may name (FragmentHeader::barcode_ < barcodeCount)                */
    uint32_t pad;
    const struct isaac_ext_tls *barcodeTls;     /* barcodeCount template length statistics (updatePairDetails, :267-269)        */
    const uint32_t *barcodeGapGroup;  /* gap group of every barcode (REALIGN_SAMPLE / REALIGN_PROJECT), NULL = one group (REALIGN_ALL) */
} isaac_ext_realign_options_t;

#define ISAAC_EXT_REALIGN_OWN_CIGAR 0xFFFFFFFFu
/* Owned by the context, valid until its next isaac_ext_realign_bin.  The records themselves are updated in the caller's buffer the
 * way the reference updates its PackedFragmentBuffer (fStrandPosition_, observedLength_, editDistance_, bamTlen_,
 * mateFStrandPosition_, flags_.properPair_ of the fragment and its mate; the CIGAR bytes of a record are never rewritten). */
typedef struct isaac_ext_realign_result {
    const uint64_t *position;         /* Index::pos_ of every index entry after the pass (ReferencePosition::getValue)          */
    const uint32_t *cigarOffset;      /* first word of the entry's CIGAR in realignedCigars, ISAAC_EXT_REALIGN_OWN_CIGAR = the
                                         record's own CIGAR is still the one (Index::cigarBegin_ / cigarEnd_)                    */
    const uint32_t *cigarLength;      /* operations of the entry's CIGAR                                                       */
    const uint32_t *realignedCigars;  /* GapRealigner::realignedCigars_: one CIGAR per realigned entry, in no particular order  */
    uint64_t realignedCigarWords;
    uint64_t realignedFragments;      /* entries that left with a new CIGAR                                                     */
    const isaac_ext_gap_t *gaps;      /* RealignerGaps::gapGroups_ of every group after finalizeGaps: by group, start, signed length */
    const isaac_ext_gap_t *deletionsByEnd;   /* RealignerGaps::deletionEndGroups_: the deletions by group and end position       */
    uint64_t gapCount, deletionCount;
    float    collectMs, realignMs;    /* duration of the two device phases alone (CUDA events), for bench.py                    */
} isaac_ext_realign_result_t;

/* collectGaps + realignGaps of one bin on the context's GPU against the resident reference.  data: the bin's records (read and
 * updated in place); recordOffset: byte offset of each of the recordCount records of the bin in 'data' (collectGaps reads them all,
 * whether the index still names them or not), NULL = the library walks the chain of FragmentHeader::getTotalLength itself.
 * Independent of the order of bins and, inside a bin, of the order of templates: the two mates of a pair are realigned in index
 * order by one thread, because the second one reads what the first one's updatePairDetails left (GapRealigner.cpp:222-270,
 * 1112-1113); nothing else is shared between index entries (the gaps are final before the first realign call).
 * Limits: records of more than 512 bases, or original CIGARs of more than 64 operations, are ISAAC_EXT_E_UNSUPPORTED when the
 * realigner has gaps to try on them; more than 255 gap groups are ISAAC_EXT_E_UNSUPPORTED. */
int isaac_ext_realign_bin(isaac_ext_ctx *ctx, const isaac_ext_realign_options_t *options, uint8_t *data, uint64_t dataBytes,
                          const uint64_t *recordOffset, uint64_t recordCount, const isaac_ext_bin_index_t *index,
                          uint64_t indexCount, isaac_ext_realign_result_t *result);

/* Several bins in one call (BinSorter::process runs on a pool of threads, one bin each): the jobs are taken in turn by two slots of
 * the context, each with its own stream, so that the upload of one bin runs next to the kernels and the download of the other.
 * Everything a job names is the caller's memory (page-locked memory makes the copies asynchronous): data is updated in place like
 * isaac_ext_realign_bin does; position / cigarOffset / cigarLength hold indexCount entries; realignedCigars has room for
 * realignedCigarCapacity words (8 + 2 per index entry is ample; a job whose CIGARs do not fit ends with ISAAC_EXT_E_CAPACITY).
 * status is the job's own result; the call returns the first status that is not ISAAC_EXT_OK. */
typedef struct isaac_ext_realign_job {
    const isaac_ext_realign_options_t *options;
    uint8_t *data;
    uint64_t dataBytes;
    const uint64_t *recordOffset;     /* or NULL */
    uint64_t recordCount;
    const isaac_ext_bin_index_t *index;
    uint64_t indexCount;
    uint64_t *position;               /* out */
    uint32_t *cigarOffset;            /* out */
    uint32_t *cigarLength;            /* out */
    uint32_t *realignedCigars;        /* out */
    uint64_t realignedCigarCapacity;
    uint64_t realignedCigarWords;     /* out */
    uint64_t realignedFragments;      /* out */
    int32_t  status;                  /* out */
    uint32_t pad;
} isaac_ext_realign_job_t;
int isaac_ext_realign_bins(isaac_ext_ctx *ctx, isaac_ext_realign_job_t *jobs, uint32_t jobCount);

/* Integer-pipe throughput probe used as the roofline denominator of the Smith-Waterman kernel (operations per
 * second over the whole chip).  kind 0: 32-bit add, 1: 32-bit max, 2: packed 16x2 max counted as two operations. */
int isaac_ext_measure_int32_peak(isaac_ext_ctx *ctx, int kind, double *opsPerSecond);

/* Number of kernel launches this context has issued so far (bench.py reports it as gpu_launches). */
uint64_t isaac_ext_launch_count(const isaac_ext_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* ISAAC_EXT_H */

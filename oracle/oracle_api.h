/*
 * oracle_api.h -- flat C interface shared by the two CPU checkers of the candidate-extension path:
 *
 *   oracle/_ref/libisaac_ref.so   the reference's own sources (BandedSmithWaterman.cpp, FragmentBuilder.cpp, ...)
 *                                 compiled unmodified from /root/reference by oracle/Makefile behind ref_capi.cpp
 *   oracle/libisaac_oracle.so     isaac_oracle.cpp, a scalar restatement of the same algorithms
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under isaac_aligner_b200/ may include, link or load this; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, as the checker and the
 * CPU baseline, never as the product.  Both libraries export exactly these symbols so the tests can diff them
 * against each other and against the CUDA path with one harness.
 */
#ifndef ISAAC_ORACLE_API_H
#define ISAAC_ORACLE_API_H
#include "../include/isaac_ext.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_genome {
    uint32_t contigCount;
    const char *const *contigBases;     /* upper-case ACGTN, 1 byte per base */
    const uint64_t *contigLengths;
} oracle_genome_t;

const char *oracle_kind(void);          /* "reference" or "port" */

/* The matchSelector::SequencingAdapterList every later call of this library clips with (see isaac_ext_set_adapters);
 * process-wide, count = 0 clears it. */
int oracle_set_adapters(uint32_t count, const isaac_ext_adapter_t *adapters);

/* BandedSmithWaterman::align, see isaac_ext_banded_sw_batch.  threads > 1 splits the batch over std::threads
 * (one BandedSmithWaterman object per thread, like the reference keeps one per TemplateBuilder). */
/* liboracle_port only: the same recurrence on a band of bandWidth lanes (16, 32, 64); database windows have
 * queryLength + bandWidth - 1 bases.  bandWidth 16 is oracle_banded_sw_batch (the same template instance). */
int oracle_banded_sw_wide_batch(uint32_t bandWidth, uint32_t n, const char *queries, const uint64_t *queryOffsets,
                                const uint32_t *queryLengths, const char *databases, const uint64_t *databaseOffsets,
                                int matchScore, int mismatchScore, int gapOpenScore, int gapExtendScore,
                                uint32_t maxReadLength, uint32_t cigarStride, uint32_t *cigarOut,
                                uint32_t *cigarLengthOut, uint32_t *offsetOut, uint32_t threads);
int oracle_banded_sw_batch(uint32_t n, const char *queries, const uint64_t *queryOffsets,
                           const uint32_t *queryLengths, const char *databases, const uint64_t *databaseOffsets,
                           int matchScore, int mismatchScore, int gapOpenScore, int gapExtendScore,
                           uint32_t maxReadLength, uint32_t cigarStride, uint32_t *cigarOut,
                           uint32_t *cigarLengthOut, uint32_t *offsetOut, uint32_t threads);

/* UngappedAligner::alignUngapped on every candidate, see isaac_ext_ungapped_batch. */
int oracle_ungapped_batch(const oracle_genome_t *genome, const isaac_ext_reads_t *reads,
                          const isaac_ext_config_t *config, uint32_t n, const isaac_ext_candidate_t *candidates,
                          isaac_ext_fragment_t *fragmentsOut, uint32_t *cigarOut, uint64_t *mismatchMaskOut,
                          uint32_t threads);

/* alignUngapped followed by GappedAligner::alignGapped on a copy (what FragmentBuilder::alignFragments does,
 * FragmentBuilder.cpp:199-200), see isaac_ext_gapped_batch. */
int oracle_gapped_batch(const oracle_genome_t *genome, const isaac_ext_reads_t *reads,
                        const isaac_ext_config_t *config, uint32_t n, const isaac_ext_candidate_t *candidates,
                        uint32_t cigarStride, isaac_ext_fragment_t *fragmentsOut, uint32_t *cigarOut,
                        uint64_t *mismatchMaskOut, uint32_t threads);

/* FragmentBuilder::build for every cluster, see isaac_ext_build_fragments.  Outputs go to caller-provided buffers
 * (ISAAC_EXT_E_CAPACITY if too small): fragmentsOut[fragmentCapacity], cigarsOut[cigarCapacity],
 * readFragmentBegin[clusterCount * readCount + 1], builtOut[clusterCount]. */
int oracle_build_fragments(const oracle_genome_t *genome, const isaac_ext_reads_t *reads, const isaac_ext_config_t *config,
                           const isaac_ext_build_batch_t *batch, uint64_t fragmentCapacity, isaac_ext_fragment_t *fragmentsOut,
                           uint64_t *readFragmentBegin, uint64_t cigarCapacity, uint32_t *cigarsOut, uint8_t *builtOut,
                           uint64_t *fragmentCount, uint64_t *cigarWords, uint32_t threads);

/* ShadowAligner::rescueShadow for every request, see isaac_ext_rescue_shadows.  requestFragmentBegin[requestCount + 1],
 * rescuedOut[requestCount]. */
int oracle_rescue_shadows(const oracle_genome_t *genome, const isaac_ext_reads_t *reads, const isaac_ext_config_t *config,
                          const isaac_ext_tls_t *tls, uint32_t requestCount, const isaac_ext_rescue_request_t *requests,
                          uint64_t fragmentCapacity, isaac_ext_fragment_t *fragmentsOut, uint64_t *requestFragmentBegin,
                          uint64_t cigarCapacity, uint32_t *cigarsOut, uint8_t *rescuedOut,
                          uint64_t *fragmentCount, uint64_t *cigarWords, uint32_t threads);

/* TemplateBuilder::buildFragments + buildTemplate for every cluster the way MatchSelector::processMatchList drives them
 * (MatchSelector.cpp:323-349), see isaac_ext_build_templates.  templatesOut[clusterCount],
 * fragmentsOut[clusterCount * readCount].  Only the reference build exports it. */
int oracle_build_templates(const oracle_genome_t *genome, const isaac_ext_reads_t *reads, const isaac_ext_config_t *config,
                           const isaac_ext_build_batch_t *batch, const isaac_ext_tls_t *tls,
                           const isaac_ext_template_options_t *options, isaac_ext_template_t *templatesOut,
                           isaac_ext_fragment_t *fragmentsOut, uint64_t cigarCapacity, uint32_t *cigarsOut,
                           uint64_t *cigarWords, uint32_t threads);

/* TileBarcodeStats of the tile's templates, see isaac_ext_template_stats (statsOut: 4 * ISAAC_EXT_TEMPLATE_STATS_COUNTERS).
 * Only the reference build exports it. */
int oracle_template_stats(const oracle_genome_t *genome, const isaac_ext_reads_t *reads, const isaac_ext_config_t *config,
                          const isaac_ext_build_batch_t *batch, const isaac_ext_tls_t *tls,
                          const isaac_ext_template_options_t *options, const uint8_t *pf, uint64_t *statsOut, uint32_t threads);

/* MatchSelector::determineTemplateLength for the tile (MatchSelector.cpp:188-249), see isaac_ext_determine_template_length. */
int oracle_determine_template_length(const oracle_genome_t *genome, const isaac_ext_reads_t *reads,
                                     const isaac_ext_config_t *config, const isaac_ext_build_batch_t *batch,
                                     const uint8_t *pf, int32_t mateDriftRange, isaac_ext_tls_t *tlsOut, uint32_t *stableOut);

/* alignment::trimLowQualityEnds on every cluster (Quality.cpp:71-120), see isaac_ext_trim_low_quality_ends.  Only the
 * reference build exports it. */
int oracle_trim_low_quality_ends(const isaac_ext_reads_t *reads, uint32_t baseQualityCutoff, uint16_t *endCyclesMaskedOut);

/* calculateShadowRescueRange (ShadowAligner.cpp:119-149) and TemplateLengthStatistics::mateOrientation for every request:
 * rangeOut[2 * i] / [2 * i + 1] = first / second of the pair the reference's function returns, orientationOut[i] the strand
 * rescueShadow gives the shadow (:171).  Only the reference build exports it. */
int oracle_shadow_rescue_range(const isaac_ext_reads_t *reads, const isaac_ext_tls_t *tls, uint32_t requestCount,
                               const isaac_ext_rescue_request_t *requests, int64_t *rangeOut, uint8_t *orientationOut);

/* matchSelector::FragmentCollector::add for every stored template of a tile, see isaac_ext_pack_fragments: the records are
 * written by the reference's own io::FragmentHeader constructors (Fragment.hh:100-186) from BamTemplate / FragmentMetadata /
 * Cluster objects rebuilt from the flat template result.  barcodeBytes: barcodeLength BCL bytes per cluster in front of the
 * reads (Cluster::getBarcodeSequence) or NULL.  headerMaskOut[headerLength]: 0xFF for the bytes of io::FragmentHeader that carry
 * a member, 0 for padding.  layoutOut: recordLength, readOffset[0], readOffset[1], sizeof(io::FragmentHeader).  recordsOut may be
 * NULL to query the layout only.  Only the reference build exports it. */
int oracle_pack_fragments(const isaac_ext_reads_t *reads, const isaac_ext_template_t *templates,
                          const isaac_ext_fragment_t *fragments, const uint32_t *cigars, uint64_t cigarWords,
                          const isaac_ext_pack_options_t *options, const uint8_t *barcodeBytes, uint32_t barcodeLength,
                          uint8_t *recordsOut, uint64_t *fStrandPosOut, uint8_t *initializedOut, uint8_t *headerMaskOut,
                          uint32_t *layoutOut);

/* liboracle_ref only: matchSelector::TileStats (score histograms + per-cycle arrays) of the tile's templates, see
 * isaac_ext_tile_cycle_stats; finalize != 0 applies TileStats::finalize */
int oracle_tile_cycle_stats(const oracle_genome_t *genome, const isaac_ext_reads_t *reads, const isaac_ext_config_t *config,
                            const isaac_ext_build_batch_t *batch, const isaac_ext_tls_t *tls,
                            const isaac_ext_template_options_t *options, const uint8_t *pf, uint64_t *statsOut, uint32_t finalize,
                            uint32_t threads);

/* liboracle_ref only: BinSorter::collectGaps + BinSorter::realignGaps of one bin through the reference's own RealignerGaps and
 * GapRealigner (see isaac_ext_realign_bin): data is updated in place; positionOut / cigarOffsetOut / cigarLengthOut per index entry
 * (cigarOffset 0xFFFFFFFF = the record's own CIGAR), the realigned CIGARs copied to cigarsOut back to back in index order;
 * gapsOut / deletionsOut = gapGroups_ / deletionEndGroups_ of every group in group order, countsOut = {gaps, deletions, cigar words}.
 * threadsafe (no static state but the contig cache). */
int oracle_realign_bin(const oracle_genome_t *genome, const isaac_ext_realign_options_t *options, uint8_t *data, uint64_t dataBytes,
                       const uint64_t *recordOffset, uint64_t recordCount, const isaac_ext_bin_index_t *index, uint64_t indexCount,
                       uint64_t *positionOut, uint32_t *cigarOffsetOut, uint32_t *cigarLengthOut, uint32_t *cigarsOut,
                       uint64_t cigarCapacity, isaac_ext_gap_t *gapsOut, isaac_ext_gap_t *deletionsOut, uint64_t gapCapacity,
                       uint64_t *countsOut);

#ifdef __cplusplus
}
#endif
#endif

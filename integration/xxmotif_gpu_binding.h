/*
 * xxmotif_gpu_binding.h -- the reference-side binding of libxxmotif_gpu.so, as a maintainer of
 * soedinglab/xxmotif would add it: force-included into every translation unit of the reference (-include), it DECLARES explicit
 * specialisations of the four scan / re-score member templates of StartPosUpdater (refinementPhase/StartPosUpdater.h:34-44,
 * bodies in StartPosUpdater-inl.h:6-266) for the DNA k-mer generator, so that no translation unit instantiates the CPU loops
 * for it; xxmotif_gpu_binding.cpp DEFINES them on top of the C ABI (include/xxmotif_gpu.h).  Nothing else of the reference
 * changes: reinitialize still runs (its host tables are simply not read by the replaced loops), getBestSubset,
 * updatePWM_OOL, the writers of _MotifFile.txt / _Pvals.txt / .pwm and the command line are the reference's own code.
 *
 * No reference source is copied or modified; this header only names reference declarations.
 */
#ifndef XXMOTIF_GPU_BINDING_H
#define XXMOTIF_GPU_BINDING_H
#ifdef __cplusplus
#include "refinementPhase/StartPosUpdater.h"

typedef Kmer_Generator_Reloaded<dna_hash_t> XxmDnaKmerGenerator;

template <>
void StartPosUpdater::fill_sites_with_startPosList<XxmDnaKmerGenerator>(XxmDnaKmerGenerator& kmerHash, Motif* motif, bool instanceThreshold);
template <>
void StartPosUpdater::fill_sites_with_possible_startPositions<XxmDnaKmerGenerator>(XxmDnaKmerGenerator& kmerHash, bool keepPosSetSize, Motif* motif);
template <>
void StartPosUpdater::fill_elongSites_with_possible_startPositions<XxmDnaKmerGenerator>(XxmDnaKmerGenerator& kmerHash, Motif* motif, double** pwm,
                                                                                        motif_columns_type& tmpColumns, bool palin);
template <>
void StartPosUpdater::fill_sites_with_possible_startPositions_oops<XxmDnaKmerGenerator>(XxmDnaKmerGenerator& kmerHash, Motif* motif);

/* the one context of the process (created, configured and given Global::posSet at the first call), shared by the bindings */
struct xxm_multi;
xxm_multi* xxm_binding_context();
void xxm_binding_die(const char* what);
#endif
#endif

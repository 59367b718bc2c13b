// See host_region.cpp.
#pragma once
#include <cstdint>

#include "host_stats.h"

namespace xxm {

struct Region {
	int set = 0, max = 0, start = 0, end = 0;
};

// MotifRegion(startRange, endRange).getRegion(startPosCounts, counts)
Region motif_region(LogSum& ls, const int* startPosCounts, int counts, int startRange, int endRange);

// setPositionalPval: false when the reference would exit ("In no man's land")
bool positional_pval(int max_leng, bool revcomp, int enrMax, int firstMotifColumn, int motifLength, const int32_t* sitePos, int nSites, double* out);
// the same in two steps, for site lists spread over several ranks: the sum over (this rank's) sites, then getCombinedPval_log
bool positional_logprod(int max_leng, bool revcomp, int enrMax, int firstMotifColumn, int motifLength, const int32_t* sitePos, int nSites, double* logProd,
                        int* trivial);
double positional_combine(double logProd, int64_t nSites);

}  // namespace xxm

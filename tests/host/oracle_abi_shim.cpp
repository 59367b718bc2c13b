// TEST INFRASTRUCTURE ONLY -- the nine data-sized C-ABI calls that csrc/refine.cpp makes (include/xxmotif_gpu.h), served by the
// oracle (oracle/xxm_oracle.h, the C restatement of the reference pinned on its golden traces) instead of the GPU library, so
// that the callers of the path -- optimizeMotif, maximizeMotifLength, the per-motif loop, and their several-contexts variants --
// can be held to the reference's own results (stages A, E, F of tests/golden/*.trace.gz) in a container without a GPU.
// Linked into a test-only shared object together with the product's refine.cpp; nothing under xxmotif_b200/ knows of it.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/xxmotif_gpu.h"
#include "../../oracle/xxm_oracle.h"

struct xxm_ctx {
	xo_ctx* xo = nullptr;
	const xo_seqset* set = nullptr;
	const xo_nullmodel* nm = nullptr;
	double bgLog[5] = {0, 0, 0, 0, 0};
	int revcomp = 0;
	xo_tables* tab = nullptr;
	std::vector<int32_t> cols;
	std::vector<xo_site> kept;
	bool haveKept = false;
	long calls = 0;
	std::string err;
	int fail(const char* m) { err = m; return 1; }
};

extern "C" {

xxm_ctx* shim_wrap(xo_ctx* xo, const xo_seqset* set, const xo_nullmodel* nm, const double* bgLog5, int revcomp) {
	xxm_ctx* c = new xxm_ctx();
	c->xo = xo; c->set = set; c->nm = nm; c->revcomp = revcomp;
	std::memcpy(c->bgLog, bgLog5, sizeof c->bgLog);
	return c;
}
void shim_free(xxm_ctx* c) {
	if (!c) return;
	if (c->tab) xo_tables_free(c->tab);
	delete c;
}
long shim_calls(const xxm_ctx* c) { return c->calls; }

const char* xxm_last_error(const xxm_ctx* c) { return c ? c->err.c_str() : ""; }

int xxm_reinitialize(xxm_ctx* c, const double* pwm, const int32_t* cols, int32_t ncols, int32_t posSetSize, int32_t totalPositions) {
	c->calls++;
	if (c->tab) xo_tables_free(c->tab);
	c->tab = xo_tables_new(pwm, cols, ncols, posSetSize, totalPositions, c->bgLog, c->nm);
	c->cols.assign(cols, cols + ncols);
	if (!c->tab || xo_tables_error(c->tab)) return c->fail("table build failed");
	return 0;
}

int xxm_scan_full(xxm_ctx* c, int32_t posSetSize, int32_t enrStart, int32_t enrEnd, int32_t* nbSites, int32_t* nbBestSites, int32_t* maxSites) {
	c->calls++;
	if (!c->tab) return c->fail("no model");
	xo_scan_full(c->xo, c->tab, c->cols.data(), static_cast<int32_t>(c->cols.size()), posSetSize, enrStart, enrEnd, 1);
	*nbSites = xo_ctx_nbSites(c->xo); *nbBestSites = xo_ctx_nbBestSites(c->xo); *maxSites = xo_ctx_maxSites(c->xo);
	return 0;
}

int xxm_rescore(xxm_ctx* c, int32_t firstMotifColumn, int32_t posSetSize, int32_t enrStart, int32_t enrEnd, int32_t palin, int32_t* nbSites, int32_t* maxSites) {
	c->calls++;
	if (!c->tab) return c->fail("no model");
	xo_rescore(c->xo, c->tab, firstMotifColumn, c->cols.data(), static_cast<int32_t>(c->cols.size()), posSetSize, enrStart, enrEnd, palin);
	*nbSites = xo_ctx_nbSites(c->xo); *maxSites = xo_ctx_maxSites(c->xo);
	return 0;
}

int xxm_update_pwm_with_best_subset(xxm_ctx* c, int32_t which, double* pwm, int32_t firstColumn, int32_t posSetSize, int32_t plus, double pseudo, int32_t* updated,
                                    int32_t* bestMotifNb, double* bestPval) {
	c->calls++;
	if (c->cols.empty() || c->cols[0] != firstColumn) return c->fail("first column differs from the current model's");
	int32_t nb = 0;
	double bp = 0;
	*updated = xo_update_pwm_with_best_subset(c->xo, which, pwm, c->cols.data(), static_cast<int32_t>(c->cols.size()), posSetSize, 0, plus, pseudo, &nb, &bp);
	if (*updated) { *bestMotifNb = nb; *bestPval = bp; }
	return 0;
}

int xxm_elong_keep(xxm_ctx* c) {
	c->calls++;
	const int64_t cap = xo_ctx_cap(c->xo);
	c->kept.assign(xo_ctx_elong(c->xo), xo_ctx_elong(c->xo) + cap);
	c->haveKept = true;
	return 0;
}

int xxm_startposlist_from_kept(xxm_ctx* c, int32_t bestMotifNb, int32_t firstTmpColumn, int32_t* outSeq, int32_t* outPos, int32_t cap, int32_t* nOut, double* worstScore) {
	c->calls++;
	if (!c->haveKept) return c->fail("nothing kept");
	xo_ctx_set_elong(c->xo, c->kept.data(), static_cast<int64_t>(c->kept.size()));
	double worst = 1;
	const int32_t n = xo_startposlist_from_elong(c->xo, bestMotifNb, firstTmpColumn, outSeq, outPos, cap, &worst);
	if (n > cap) return c->fail("output capacity too small");
	*nOut = n;
	*worstScore = worst;
	return 0;
}

int xxm_share_candidates(xxm_ctx* c, xxm_ctx* owner) {
	c->calls++;
	if (!owner || owner->set != c->set) return c->fail("not a view of that owner");
	xo_ctx_copy_candidates(c->xo, owner->xo);
	c->haveKept = false;
	return 0;
}

int xxm_update_enriched_region(xxm_ctx* c, int32_t firstMotifColumn, int32_t motifLength, int32_t* enr) {
	c->calls++;
	xo_update_enriched_region(c->xo, xo_ctx_best(c->xo), c->revcomp, xo_seqset_max_leng(c->set), firstMotifColumn, motifLength, enr);
	return 0;
}

int xxm_positional_pval(xxm_ctx* c, int32_t enrMax, int32_t firstMotifColumn, int32_t motifLength, const int32_t* sitePos, int32_t nSites, double* out) {
	c->calls++;
	const double p = xo_positional_pval(c->xo, enrMax, firstMotifColumn, motifLength, sitePos, nSites);
	if (p != p) return 2;  // "In no man's land": the reference exits
	*out = p;
	return 0;
}

}  // extern "C"

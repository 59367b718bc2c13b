// libxxmotif_gpu.so -- host side of the C ABI in include/xxmotif_gpu.h: owns the device buffers, stages the model of
// every reinitialize(), launches the kernels of xxm_kernels.cuh and runs the sequential host statistics.
// There is no CPU fallback: every entry point needs the CUDA device the context was created on.
#include "../../include/xxmotif_gpu.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <csignal>
#include <execinfo.h>
#include <unistd.h>
#include <cstring>
#include <atomic>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "host_pack.h"
#include "host_stats.h"
#include "pvalue_tables.h"
#include "xxm_kernels.cuh"
#include "xxm_select2.cuh"
#include "xxm_select3.cuh"
#include "xxm_select4.cuh"
#include "xxm_kmerhist.cuh"
#include "xxm_gappedhist.cuh"
#include "xxm_scan4.cuh"
#include "xxm_scan5.cuh"
#include "xxm_scan6.cuh"
#include "xxm_scan7.cuh"
#include "xxm_tables_device.cuh"
#include "pair_table.h"
#include "nccl_dyn.h"

using namespace xxm;

// NVTX ranges: one per C-ABI call (nsys / ncu --nvtx show which call of the reference's loop a kernel belongs to)
#include <nvtx3/nvToolsExt.h>
namespace {
// XXM_TRACE: wall time per C-ABI entry point, summed per name and printed when a context is destroyed
struct CallTimes {
	struct E { const char* name; uint64_t n; double us; };
	std::vector<E> e;
	bool on = std::getenv("XXM_TRACE") != nullptr;
	std::mutex mu;  // contexts may be driven from several host threads (xxm_multi, xxm_iterate_motifs)
	void add(const char* name, double us) {
		std::lock_guard<std::mutex> g(mu);
		for (E& x : e) if (x.name == name) { x.n++; x.us += us; return; }
		e.push_back(E{name, 1, us});
	}
	void print() {
		if (!on) return;
		std::lock_guard<std::mutex> g(mu);
		if (e.empty()) return;
		std::sort(e.begin(), e.end(), [](const E& a, const E& b) { return a.us > b.us; });
		for (const E& x : e) std::fprintf(stderr, "xxm calls: %-36s n=%6llu total %10.1f us  avg %9.1f us\n", x.name, static_cast<unsigned long long>(x.n), x.us, x.us / x.n);
		e.clear();
	}
};
CallTimes g_callTimes;
struct NvtxRange {
	const char* name;
	std::chrono::steady_clock::time_point t0;
	explicit NvtxRange(const char* nm) : name(nm) { nvtxRangePushA(nm); if (g_callTimes.on) t0 = std::chrono::steady_clock::now(); }
	~NvtxRange() {
		nvtxRangePop();
		if (g_callTimes.on) g_callTimes.add(name, std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count());
	}
};
}  // namespace
#define XXM_NVTX(name) NvtxRange nvtx_range__(name)
// XXM_TRACE: named time marks inside a call (summed like the calls themselves)
struct TraceMark {
	std::chrono::steady_clock::time_point t;
	TraceMark() { if (g_callTimes.on) t = std::chrono::steady_clock::now(); }
	void mark(const char* name) {
		if (!g_callTimes.on) return;
		const auto n = std::chrono::steady_clock::now();
		g_callTimes.add(name, std::chrono::duration<double, std::micro>(n - t).count());
		t = n;
	}
};

namespace {

thread_local std::string g_create_error;

// XXM_DEBUG_GUARDS: every device buffer gets a 256-byte canary behind its last element; xxm_check_guards verifies that no
// kernel wrote past the end of any buffer (compute-sanitizer is not available on every pool; tests/test_gpu_guards.py runs the
// ragged / low-complexity stress cases and the golden replays under it)
constexpr size_t kGuardBytes = 256;
struct GuardRegistry {
	bool on = std::getenv("XXM_DEBUG_GUARDS") != nullptr;
	std::vector<std::pair<void*, size_t>> bufs;  // (device pointer, user bytes)
	std::mutex mu;
	void add(void* p, size_t bytes) { std::lock_guard<std::mutex> g(mu); bufs.emplace_back(p, bytes); }
	void remove(void* p) {
		std::lock_guard<std::mutex> g(mu);
		for (size_t i = 0; i < bufs.size(); ++i)
			if (bufs[i].first == p) { bufs[i] = bufs.back(); bufs.pop_back(); return; }
	}
};
GuardRegistry g_guards;

template <typename T>
struct DevBuf {
	T* p = nullptr;
	size_t cap = 0;
	bool borrowed = false;  // p belongs to another context (xxm_share_seqset): never written through, never freed here
	DevBuf() = default;
	DevBuf(const DevBuf&) = delete;
	DevBuf& operator=(const DevBuf&) = delete;
	~DevBuf() { release(); }  // (xxm_destroy releases the buffers with the device current; this only catches what it does not name)
	void borrow(const DevBuf& o) { release(); p = o.p; cap = o.cap; borrowed = true; }
	cudaError_t ensure(size_t n) {
		if (n <= cap) return cudaSuccess;
		release();
		size_t want = n + n / 8 + 64;
		cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), want * sizeof(T) + (g_guards.on ? kGuardBytes : 0));
		if (e == cudaSuccess) {
			cap = want;
			if (g_guards.on) {
				cudaMemset(reinterpret_cast<uint8_t*>(p) + want * sizeof(T), 0xA5, kGuardBytes);
				g_guards.add(p, want * sizeof(T));
			}
		}
		return e;
	}
	void release() {
		if (p && !borrowed) { if (g_guards.on) g_guards.remove(p); cudaFree(p); }
		p = nullptr;
		cap = 0;
		borrowed = false;
	}
};

// a device-resident site list (members of StartPosUpdater: bestStartPos / sortedSites / sortedSites_elong)
enum ListKind { kListEmpty = 0, kListCompact, kListFlagged, kListSlots };
struct SiteList {
	ListKind kind = kListEmpty;
	const uint32_t* seq = nullptr;  // entry arrays (may alias another list's arrays)
	const uint32_t* pos = nullptr;
	const uint32_t* vidx = nullptr; // only meaningful while the model that produced it is current
	uint32_t n = 0;                 // entries
	int shift = 0;                  // startPos = pos + shift
	const float* score = nullptr;   // per entry (compact/flagged) or per slot (slots)
	const uint8_t* keep = nullptr;  // flagged
	const uint32_t* slotIdx = nullptr;
	uint32_t nslots = 0;
	uint32_t members = 0;           // sites with score < 1
	uint32_t maxPerSeq = 0;
	bool orderedBySeqPos = true;
	bool tableValued = false;       // float score is a function of vidx under the current model
	bool mopsValues = true;         // ... taken from the mops (true) or the mode (false) finalPval table
	uint64_t modelSerial = 0;
	uint64_t histSerial = 0;        // != 0: ctx->hist already holds this list's value-index histogram (fused into the select kernel)
	const double* lazyPS = nullptr; // score == nullptr: float(lazyPS[vidx]) on demand (snapshot of the pS table of the scan's model)
	int lazyDst = 0;                // where ensure_scores writes them: 0 = b_score (bestStartPos), 1 = s_kscore (sortedSites)
};

}  // namespace

struct xxm_ctx {
	int device = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr, evMid = nullptr;
	std::string err;
	xxm_options opt{};
	bool haveOpt = false;
	int numSMs = 148;
	xxm_stats stats{};

	// ---- sequence set
	int nent = 0;
	bool revcomp = false;
	std::vector<int32_t> len, total;
	std::vector<double> logAcc;  // running sum of log(non-N residues): avgLength[i] = exp(logAcc[i] / i), evaluated on demand
	double avg_length(int i) const { return i ? std::exp(logAcc[i] / i) : 0.0; }  // calculateSequenceFeatures, seqset.cpp:31-34
	int max_leng = 0, min_leng = 0;
	std::vector<uint32_t> wordOff;
	uint32_t nwords = 0;
	DevBuf<uint64_t> d_codes;
	DevBuf<uint32_t> d_nmask, d_wordOff, d_blk2seq;
	DevBuf<int32_t> d_len;
	SeqStream S{};
	// several contexts of one device on one resident set (xxm_share_seqset): a view borrows the five read-only buffers above and
	// the NullModel tables of its owner; the owner counts its views and refuses to reload or be re-modelled while any exist
	xxm_ctx* seqOwner = nullptr;
	std::atomic<int> views{0};

	// ---- background + model
	NullModelView nm;
	std::vector<float> rprobs, rconds;
	PvalueTables tab;
	BgProbCache bgCache;          // background probabilities of left k-mers under the current null model (cleared by xxm_set_background)
	bool haveModel = false;
	uint64_t modelSerial = 0;
	int modelPosSetSize = 0;
	int modelFirstCol = 1;  // cols[0] of the current model
	Model M{};
	DevBuf<uint32_t> d_rank0, d_rank1, d_hit0, d_ids[2], d_cls[2];
	DevBuf<uint8_t> d_clsOk[2];
	// batched staging of the small per-model tables (xxm_reinitialize): everything goes through one pinned buffer, one
	// H2D copy and one scatter kernel instead of a dozen pageable copies
	struct UpSeg { void* dst; uint32_t off, bytes; };
	struct FillJob { int k; int n; uint32_t* rank; uint32_t* bitmap; };
	bool batching = false, stageInFlight = false;
	std::vector<UpSeg> upSegs;
	std::vector<FillJob> fillJobs;
	uint8_t* h_stage = nullptr;
	size_t h_stageCap = 0, stageUsed = 0;
	DevBuf<uint8_t> d_stage;
	cudaEvent_t evStage = nullptr;
	// a full scan whose launches are queued but whose counts have not been read yet (xxm_optimize_step)
	struct PendingScan { bool active = false; size_t cap = 0; int posSetSize = 0; uint32_t scanWords = 0; };
	PendingScan pend;
	unsigned long long* h_pin = nullptr;  // pinned: [0] hits of the pending scan, [1] its select counters (2 x u32)
	DevBuf<double> d_cpval0, d_cpval1, d_leftProb;
	DevBuf<int32_t> d_cscore0, d_cscore1;
	DevBuf<int32_t> d_memoKey;
	DevBuf<double> d_memoVal;
	DevBuf<uint32_t> d_memoState, d_flags32;
	// value tables
	uint32_t nval = 0;
	DevBuf<double> d_pS, d_finalMops, d_finalMode, d_posProb, d_perCand;
	std::vector<double> h_pS, h_finalMops, h_finalMode;  // nparts == 1: per class; nparts == 2: per slot after a round trip
	bool memoModeValid = false;  // one-table models: h_finalMode / d_finalMode hold the mode table of the current model
	double modeExponent = 0;
	std::vector<double> posProb, sum_i;
	bool posProbValid = false;

	// ---- lists
	DevBuf<uint32_t> b_seq, b_pos, b_vidx, t_seq, t_pos, t_vidx, u_seq, u_pos, u_vidx;
	DevBuf<double> b_pSsnap;
	DevBuf<float> b_score, s_kscore, e_kscoreSet[2], slotScore, e_slotScoreSet[2], tmpScore, u_score;
	DevBuf<uint8_t> s_keep, e_keepSet[2], flag8;
	// sortedSites_elong is double buffered: xxm_elong_keep sets the current list aside (the best column variant so far) and the
	// following re-scores write the other set
	int eSet = 0;
	SiteList keptElong;
	int keptNbSites = 0;
	float keptLastScore = 0;
	bool keptLastScoreValid = false;
	DevBuf<uint16_t> s_kcls;      // value index per kept entry of sortedSites (0xFFFF: dropped), valid while kclsSerial == the list's histSerial
	uint64_t kclsSerial = 0;
	// class order of the current one-table model (stage_class_order): fin[] non-decreasing, T[k] and the number of site classes
	std::vector<uint16_t> h_clsThr;
	DevBuf<uint16_t> d_clsThr;
	bool clsOrderOK = false;
	uint32_t clsNSite = 0;
	int lastLoadRawChunks = 0, lastLoadChunks = 0;  // hybrid load: how many chunks went over the bus raw
	float lastKmerHistMs = 0;     // device time of the last xxm_kmer_counts (all passes)
	bool sortedLazy = false;      // the last deferred scan left sortedSites' float scores implicit (k_select_cls_fwd)
	DevBuf<uint32_t> stk, slotIdx, e_slotIdxSet[2], counters, hist, idxA, idxB, keysA, keysB, rsHist;
	DevBuf<uint64_t> status, total64;
	DevBuf<uint32_t> cmpCounts, runStart, memoIdx;
	// device table build (xxm_gpu_part8.inc)
	DevBuf<float> d_nmProbs, d_nmConds;
	bool nmOnDevice = false;
	DevBuf<int32_t> tbScore[2], tbClassScore, tbClassCount;
	DevBuf<uint32_t> tbId[2], tbCls;
	DevBuf<double> tbProb[2], tbClassPval;
	DevBuf<uint8_t> memoFlag;
	DevBuf<double> memoTmp;
	std::vector<double> h_memoVals;
	std::vector<xxm_site> h_sites;     // landing buffer of materialised site lists
	std::vector<int32_t> h_sortedSeq;  // sequence of every site in score order, kept when a per-sequence cap binds
	DevBuf<unsigned long long> runPack, sweepCand;
	DevBuf<double> sweepApprox;
	DevBuf<uint32_t> r_seq, d_pbits;
	DevBuf<uint8_t> d_t8, d_raw;
	DevBuf<uint32_t> d_pairTab, d_dead, d_scanCounts;
	DevBuf<uint32_t> d_gapped;      // bins of xxm_count_gapped_kmers
	DevBuf<uint32_t> d_siteBits;    // one bit per value index: final p-value < 1
	DevBuf<uint32_t> sweepLong;     // runs whose tail loop goes to the warp-per-run kernel
	std::vector<uint32_t> h_gapped;
	std::vector<int> deadKey;   // (load serial, L, column offsets) d_dead was built for
	uint64_t loadSerial = 0;
	DevBuf<unsigned long long> d_xchg;  // value-index histogram + scalars on their way through NCCL
	std::vector<unsigned long long> h_xchg;
	unsigned long long* h_pinXchg = nullptr;  // pinned landing buffer of the exchange vector / the count matrix
	size_t h_pinXchgCap = 0;
	std::vector<ScoreRun> h_runs;
	void* comm = nullptr;               // ncclComm_t of a sharded set (xxm_comm_init), else null
	int commRank = 0, commSize = 1;
	std::vector<uint32_t> h_pairIds, h_pairTab;
	bool pairOK = false;           // the pair table represents the left table exactly: k_scan5, else k_scan4
	bool forceGenericScan = false; // tests: always take the generic find kernel
	int scanGen = 7;               // 7: window-order masks + k_expand7 (xxm_scan7.cuh); 6: even/odd masks + k_expand6
	DevBuf<uint64_t> d_rawOff, d_layoutSums;
	DevBuf<int32_t> d_notX, d_n;
	std::vector<int32_t> notX;  // residues that are not N, per sequence (host_pack.h)
	int32_t* h_notX = nullptr;  // pinned landing buffer of the same counts when the device packs
	size_t h_notXCap = 0;
	cudaStream_t copyStream = nullptr;
	std::vector<cudaEvent_t> evPool;
	uint8_t* h_pack = nullptr;  // pinned staging of the packed stream: h_packCap code words, then as many N masks
	size_t h_packCap = 0;
	DevBuf<unsigned long long> d_counts;
	DevBuf<SiteRec> d_sites;
	DevBuf<unsigned int> d_hf, d_hb;
	uint32_t nBest = 0;
	bool scanFindsOnlySites = false;  // single table, no localization: every window the scan finds has finalPval < 1
	bool histValid = false;      // ctx->hist holds the value-index histogram of the list made by the last select_mops
	uint64_t histSerial = 0, histCounter = 0;
	SiteList L[3];
	int nbSites = 0;
	float lastChoiceScore[3] = {0, 0, 0};  // score of the last site of the optimal subset last chosen on a list
	bool lastChoiceValid[3] = {false, false, false};
	int64_t nbSitesGlobal = 0;  // sites of all ranks seen by the last subset choice on a sharded set
	int64_t maxSites = 0;  // (8 x 10^9 windows of a sharded set exceed the reference's int)

	LogSum logsum;
	// the sharded set (xxm_gpu_part7.inc): global counts, valid while synced for the current load
	struct ShardInfo { bool synced = false, external = false, haveLogAcc = false; uint64_t loadSerial = 0; int64_t nent = 0, total = 0, seqOffset = 0; double logAcc = 0; int maxLeng = 0; } G;
	const int32_t* notXPtr = nullptr;  // residues that are not N per sequence of the current load (host or device packing path)
	double phaseUs[7] = {0, 0, 0, 0, 0, 0, 0};  // xxm_step_phases
	int64_t phaseSteps = 0;

	int fail(const char* fmt, ...) {
		char buf[512];
		va_list ap;
		va_start(ap, fmt);
		vsnprintf(buf, sizeof buf, fmt, ap);
		va_end(ap);
		err = buf;
		return 1;
	}
};

#define CK(call)                                                                                              \
	do {                                                                                                      \
		cudaError_t e__ = (call);                                                                             \
		if (e__ != cudaSuccess) return ctx->fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
	} while (0)
#define CKK()                                                                                                 \
	do {                                                                                                      \
		ctx->stats.launches++;                                                                                \
		cudaError_t e__ = cudaGetLastError();                                                                 \
		if (e__ != cudaSuccess) return ctx->fail("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
	} while (0)
#define TRY(expr)            \
	do {                     \
		int r__ = (expr);    \
		if (r__) return r__; \
	} while (0)

static inline uint32_t cdiv(uint64_t a, uint32_t b) { return static_cast<uint32_t>((a + b - 1) / b); }
static int stage_scan(xxm_ctx* ctx);  // xxm_gpu_scan5.inc
static void drop_shared_seqset(xxm_ctx* ctx);

#define NCK(call)                                                                                              \
	do {                                                                                                       \
		ncclResult_t r__ = (call);                                                                             \
		if (r__ != ncclSuccess) return ctx->fail("%s failed: %s", #call, nccl_api().GetErrorString(r__));      \
	} while (0)

static int ensure_pinned_xchg(xxm_ctx* ctx, size_t bytes) {
	if (ctx->h_pinXchgCap >= bytes) return 0;
	if (ctx->h_pinXchg) cudaFreeHost(ctx->h_pinXchg);
	ctx->h_pinXchg = nullptr;
	ctx->h_pinXchgCap = 0;
	const size_t want = std::max<size_t>(bytes + bytes / 4, 1u << 16);
	CK(cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_pinXchg), want, cudaHostAllocDefault));
	ctx->h_pinXchgCap = want;
	return 0;
}

#include "xxm_gpu_part7.inc"

// =====================================================================================================================
// life cycle
// =====================================================================================================================
static void xxm_segv_handler(int sig) {
	void* frames[64];
	const int n = backtrace(frames, 64);
	const char msg[] = "libxxmotif_gpu: fatal signal, native backtrace:\n";
	if (write(2, msg, sizeof msg - 1) < 0) {}
	backtrace_symbols_fd(frames, n, 2);
	signal(sig, SIG_DFL);
	raise(sig);
}

extern "C" void xxm_destroy(xxm_ctx* ctx);
extern "C" int xxm_create(int device, xxm_ctx** out) {
	*out = nullptr;
	if (std::getenv("XXM_DEBUG_SEGV")) { signal(SIGSEGV, xxm_segv_handler); signal(SIGABRT, xxm_segv_handler); }  // debugging aid: native backtrace on a crash
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0) {
		g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libxxmotif_gpu has no CPU fallback)";
		return 1;
	}
	if (device < 0 || device >= ndev) { g_create_error = "device index out of range"; return 1; }
	e = cudaSetDevice(device);
	if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return 1; }
	xxm_ctx* ctx = new xxm_ctx();
	ctx->device = device;
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->numSMs = prop.multiProcessorCount;
	if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&ctx->ev0) != cudaSuccess ||
	    cudaEventCreate(&ctx->ev1) != cudaSuccess || cudaEventCreate(&ctx->evMid) != cudaSuccess ||
	    cudaEventCreateWithFlags(&ctx->evStage, cudaEventDisableTiming) != cudaSuccess) {
		g_create_error = "cannot create stream/events";
		delete ctx;
		return 1;
	}
	if (cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_pin), 64, cudaHostAllocDefault) != cudaSuccess) {
		g_create_error = "cannot allocate pinned memory";
		delete ctx;
		return 1;
	}
	// opt-in shared-memory sizes are per DEVICE, not per process: set them for the device of every new context
	if (cudaFuncSetAttribute(k_scan5<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairTabWords * 4) != cudaSuccess ||
	    cudaFuncSetAttribute(k_scan5<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairTabWords * 4) != cudaSuccess ||
	    cudaFuncSetAttribute(k_scan4<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 32 * 1024) != cudaSuccess ||
	    cudaFuncSetAttribute(k_scan4<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 16) != cudaSuccess ||
	    cudaFuncSetAttribute(k_scan4<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100) != cudaSuccess ||
	    cudaFuncSetAttribute(k_scan4<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100) != cudaSuccess ||
	    cudaFuncSetAttribute(k_select_cls_rc, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 6) != cudaSuccess ||
	    cudaFuncSetAttribute(k_memo_eval_list, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
		g_create_error = std::string("cannot opt in to large shared memory: ") + cudaGetErrorString(cudaGetLastError());
		xxm_destroy(ctx);
		return 1;
	}
	ctx->opt.maxMotifsPerSequence = 1000;
	ctx->opt.maxSeqCount = 1000;
	ctx->opt.overrepCorrection = 1.05;
	ctx->forceGenericScan = std::getenv("XXM_FORCE_GENERIC_SCAN") != nullptr;  // tests: exercise k_scan4 on every model
	// XXM_SCAN_GEN=6: the previous generation of the pair-table scan (even/odd masks + k_expand6), kept for A/B timing and tests
	if (const char* g = std::getenv("XXM_SCAN_GEN")) ctx->scanGen = (std::atoi(g) == 6) ? 6 : 7;
	*out = ctx;
	return 0;
}

// a view lets go of its owner's set (before it loads its own, adopts another one, or dies)
static void drop_shared_seqset(xxm_ctx* ctx) {
	if (!ctx->seqOwner) return;
	ctx->d_codes.release(); ctx->d_nmask.release(); ctx->d_wordOff.release(); ctx->d_blk2seq.release(); ctx->d_len.release();
	std::memset(&ctx->S, 0, sizeof ctx->S);
	ctx->nm = NullModelView();
	ctx->bgCache.clear();
	ctx->nmOnDevice = false;
	ctx->nent = 0;
	ctx->haveModel = false;
	ctx->seqOwner->views.fetch_sub(1);
	ctx->seqOwner = nullptr;
}

extern "C" int xxm_comm_destroy(xxm_ctx* ctx);
extern "C" void xxm_destroy(xxm_ctx* ctx) {
	if (!ctx) return;
	g_callTimes.print();
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	xxm_comm_destroy(ctx);
	drop_shared_seqset(ctx);
	if (ctx->views.load() > 0)
		std::fprintf(stderr, "libxxmotif_gpu: xxm_destroy of a context whose sequence set is still shared by %d view(s): destroy the views first\n", ctx->views.load());
	// DevBuf members are released explicitly (no destructor: the struct is copied nowhere, but keep it trivial)
	ctx->d_codes.release(); ctx->d_nmask.release(); ctx->d_wordOff.release(); ctx->d_blk2seq.release(); ctx->d_len.release();
	ctx->d_rank0.release(); ctx->d_rank1.release(); ctx->d_hit0.release(); for (int k = 0; k < 2; ++k) { ctx->d_ids[k].release(); ctx->d_cls[k].release(); ctx->d_clsOk[k].release(); }
	ctx->d_stage.release(); if (ctx->h_stage) cudaFreeHost(ctx->h_stage); if (ctx->h_pin) cudaFreeHost(ctx->h_pin); if (ctx->h_pack) cudaFreeHost(ctx->h_pack); if (ctx->h_notX) cudaFreeHost(ctx->h_notX); ctx->d_raw.release(); ctx->d_notX.release(); ctx->d_n.release(); ctx->d_dead.release(); ctx->d_gapped.release(); ctx->d_siteBits.release(); ctx->sweepLong.release(); ctx->d_scanCounts.release(); ctx->s_kcls.release(); ctx->d_clsThr.release(); ctx->d_layoutSums.release();
	for (cudaEvent_t e : ctx->evPool) cudaEventDestroy(e);
	if (ctx->copyStream) cudaStreamDestroy(ctx->copyStream);
	if (ctx->h_pinXchg) cudaFreeHost(ctx->h_pinXchg); if (ctx->evStage) cudaEventDestroy(ctx->evStage);
	ctx->d_cpval0.release(); ctx->d_cpval1.release(); ctx->d_leftProb.release(); ctx->d_cscore0.release(); ctx->d_cscore1.release();
	ctx->d_memoKey.release(); ctx->d_memoVal.release(); ctx->d_memoState.release(); ctx->d_flags32.release();
	ctx->d_pS.release(); ctx->d_finalMops.release(); ctx->d_finalMode.release(); ctx->d_posProb.release(); ctx->d_perCand.release();
	ctx->b_seq.release(); ctx->b_pos.release(); ctx->b_vidx.release(); ctx->t_seq.release(); ctx->t_pos.release(); ctx->t_vidx.release();
	ctx->u_seq.release(); ctx->u_pos.release(); ctx->u_vidx.release(); ctx->b_score.release(); ctx->b_pSsnap.release(); ctx->s_kscore.release(); for (int k = 0; k < 2; ++k) { ctx->e_kscoreSet[k].release(); ctx->e_slotScoreSet[k].release(); ctx->e_keepSet[k].release(); ctx->e_slotIdxSet[k].release(); }
	ctx->slotScore.release(); ctx->tmpScore.release(); ctx->s_keep.release(); ctx->flag8.release(); ctx->stk.release();
	ctx->slotIdx.release(); ctx->u_score.release(); ctx->counters.release(); ctx->hist.release(); ctx->idxA.release(); ctx->idxB.release(); ctx->keysA.release();
	ctx->keysB.release(); ctx->rsHist.release(); ctx->d_nmProbs.release(); ctx->d_nmConds.release(); for (int k = 0; k < 2; ++k) { ctx->tbScore[k].release(); ctx->tbId[k].release(); ctx->tbProb[k].release(); } ctx->tbCls.release(); ctx->tbClassScore.release(); ctx->tbClassCount.release(); ctx->tbClassPval.release(); ctx->cmpCounts.release(); ctx->memoIdx.release(); ctx->memoFlag.release(); ctx->memoTmp.release(); ctx->runStart.release(); ctx->runPack.release(); ctx->sweepCand.release(); ctx->sweepApprox.release(); ctx->status.release(); ctx->total64.release(); ctx->d_counts.release(); ctx->d_sites.release();
	ctx->d_hf.release(); ctx->d_hb.release();
	ctx->r_seq.release(); ctx->d_pbits.release(); ctx->d_t8.release(); ctx->d_pairTab.release();
	cudaEventDestroy(ctx->ev0);
	cudaEventDestroy(ctx->ev1);
	cudaEventDestroy(ctx->evMid);
	cudaStreamDestroy(ctx->stream);
	delete ctx;
}

extern "C" const char* xxm_last_error(const xxm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int xxm_set_options(xxm_ctx* ctx, const xxm_options* opt) {
	ctx->opt = *opt;
	ctx->haveOpt = true;
	ctx->haveModel = false;
	return 0;
}

// XXM_DEBUG_GUARDS: number of device buffers whose canary was overwritten (0 = clean); *checked = buffers looked at
extern "C" int xxm_check_guards(xxm_ctx* ctx, int32_t* checked, int32_t* damaged) {
	CK(cudaSetDevice(ctx->device));
	CK(cudaDeviceSynchronize());
	*checked = 0;
	*damaged = 0;
	if (!g_guards.on) return ctx->fail("xxm_check_guards: set XXM_DEBUG_GUARDS in the environment before the library is loaded");
	uint8_t tail[kGuardBytes];
	for (const auto& b : g_guards.bufs) {
		CK(cudaMemcpy(tail, static_cast<uint8_t*>(b.first) + b.second, kGuardBytes, cudaMemcpyDeviceToHost));
		++*checked;
		for (size_t i = 0; i < kGuardBytes; ++i)
			if (tail[i] != 0xA5) { ++*damaged; break; }
	}
	return 0;
}
// XXM_TRACE: print and clear the per-call wall times collected so far
extern "C" int xxm_trace_dump(void) { g_callTimes.print(); return 0; }
extern "C" void* xxm_stream(const xxm_ctx* ctx) { return reinterpret_cast<void*>(ctx->stream); }
extern "C" int xxm_get_stats(const xxm_ctx* ctx, xxm_stats* out) { *out = ctx->stats; return 0; }

// =====================================================================================================================
// a1: sequence set
// =====================================================================================================================
#include "xxm_load.cuh"  // k_seq_stats, k_seq_layout_sums, k_seq_layout, k_blk2seq

// The packing step on the host cores (host_pack.h): worker threads pack chunks of whole sequences into a pinned staging
// buffer, every finished chunk goes on the bus at once (0.375 bytes per stored base), and this thread folds the residue
// counts of a chunk into total[] / avgLength[] while the next chunks are being packed and copied.
static int load_pack_on_host(xxm_ctx* ctx, int32_t nent, const int32_t* n, const uint8_t* codes, int32_t revcomp, uint64_t rawTotal, int nthreads, bool trace,
                             std::chrono::steady_clock::time_point tStart) {
	uint64_t off = 0, words = 0;
	// pinned staging of the packed words (kept across loads): code words, then N masks
	const size_t wordsAll = static_cast<size_t>(ctx->nwords) + 2;
	if (ctx->h_packCap < wordsAll) {
		if (ctx->h_pack) cudaFreeHost(ctx->h_pack);
		ctx->h_pack = nullptr;
		ctx->h_packCap = 0;
		CK(cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_pack), wordsAll * 12, cudaHostAllocDefault));
		ctx->h_packCap = wordsAll;
	}
	uint64_t* const hCodes = reinterpret_cast<uint64_t*>(ctx->h_pack);
	uint32_t* const hMask = reinterpret_cast<uint32_t*>(ctx->h_pack + ctx->h_packCap * 8);
	ctx->notX.resize(nent);
	// ---- chunks of whole sequences: cut by this thread (which owns wordOff / len), packed by the workers, copied in order
	// (the workers take chunks as they come, so the load ends up to one chunk's packing time after the ideal, while every chunk
	// costs two copy calls: 1M x 1 kb with 14 threads loads in 13.2-13.4 ms with 3-4 MB chunks, 13.5-13.7 ms with 8 MB, 14.1 ms with
	// 2 MB, 17.9 ms with 1 MB -- profiles/r02e_pack_chunk_size.txt; XXM_PACK_CHUNK_KB overrides it for measurements)
	uint64_t kChunkBytes = 4ull << 20;
	if (const char* e = std::getenv("XXM_PACK_CHUNK_KB")) kChunkBytes = std::max<uint64_t>(64ull << 10, static_cast<uint64_t>(std::atoll(e)) << 10);
	struct Chunk { int s0, s1; uint64_t b0; uint32_t w0, w1; uint32_t bad; };
	const size_t maxChunks = static_cast<size_t>(rawTotal / kChunkBytes) + 2;
	std::vector<Chunk> chunks(maxChunks);
	std::unique_ptr<std::atomic<uint32_t>[]> done(new std::atomic<uint32_t>[maxChunks]);
	for (size_t k = 0; k < maxChunks; ++k) done[k].store(0, std::memory_order_relaxed);
	std::atomic<uint32_t> nReady{0}, nextChunk{0};
	std::atomic<bool> cutDone{false};
	const bool rc = revcomp != 0;
	auto packChunk = [&](uint32_t k) {
		Chunk& c = chunks[k];
		c.bad = pack_range(n, static_cast<uint32_t>(c.s0), static_cast<uint32_t>(c.s1), codes + c.b0, rawTotal - c.b0, rc, hCodes + c.w0, hMask + c.w0, ctx->notX.data() + c.s0);
		done[k].store(1, std::memory_order_release);
	};
	auto worker = [&]() {
		for (;;) {
			const uint32_t k = nextChunk.fetch_add(1, std::memory_order_relaxed);
			while (k >= nReady.load(std::memory_order_acquire)) {
				if (cutDone.load(std::memory_order_acquire) && k >= nReady.load(std::memory_order_acquire)) return;
				std::this_thread::yield();
			}
			packChunk(k);
		}
	};
	if (nthreads == 1) nthreads = 0;  // one packer: this thread
	struct Joiner {
		std::vector<std::thread> t;
		std::atomic<bool>* cutDone;
		~Joiner() { cutDone->store(true, std::memory_order_release); for (auto& th : t) if (th.joinable()) th.join(); }
	} pool{{}, &cutDone};
	for (int t = 0; t < nthreads; ++t) pool.t.emplace_back(worker);
	// fold of calculateSequenceFeatures (seqset.cpp:21-36) on the raw set, then -- with revcomp -- the recount at the end of
	// createRevcomp (seqset.cpp:165-177) which compares codes with the CHARACTER '0' and therefore counts every position
	// of the doubled sequence
	double acc = 0, lastLog = 0;
	int lastNotX = 0;
	int64_t tot = 0;
	uint32_t issued = 0, badSeq = 0;
	bool overflow = false;
	auto issue = [&](uint32_t k) -> int {
		const Chunk& c = chunks[k];
		if (c.w1 > c.w0) {
			CK(cudaMemcpyAsync(ctx->d_codes.p + c.w0, hCodes + c.w0, static_cast<size_t>(c.w1 - c.w0) * 8, cudaMemcpyHostToDevice, ctx->stream));
			CK(cudaMemcpyAsync(ctx->d_nmask.p + c.w0, hMask + c.w0, static_cast<size_t>(c.w1 - c.w0) * 4, cudaMemcpyHostToDevice, ctx->stream));
		}
		if (c.bad && !badSeq) badSeq = c.bad;
		for (int i = c.s0; i < c.s1; ++i) {
			int notX = ctx->notX[i];
			if (revcomp) notX = ctx->len[i];
			if (notX != 0) {
				if (notX != lastNotX) { lastNotX = notX; lastLog = std::log(notX); }  // same argument, same glibc result
				acc += lastLog;
			}
			tot += notX;
			if (tot > 2147483647LL) overflow = true;
			ctx->total[i + 1] = static_cast<int32_t>(tot);
			ctx->logAcc[i + 1] = acc;
		}
		return 0;
	};
	auto tFirst = tStart;
	uint32_t ncut = 0;
	{
		Chunk c{0, 0, 0, 0, 0, 0};
		off = 0;
		words = 0;
		for (int i = 0; i < nent; ++i) {
			const int n0 = n[i];
			const int nstore = revcomp ? 2 * n0 + 1 : n0;
			ctx->len[i] = nstore;
			ctx->wordOff[i] = static_cast<uint32_t>(words);
			words += (static_cast<uint64_t>(nstore) + kPadBases + 31) / 32;
			off += static_cast<uint64_t>(n0);
			if (off - c.b0 < kChunkBytes && i + 1 != nent) continue;
			c.s1 = i + 1; c.w1 = static_cast<uint32_t>(words);
			chunks[ncut] = c;
			nReady.store(++ncut, std::memory_order_release);
			if (nthreads == 0) packChunk(nextChunk.fetch_add(1, std::memory_order_relaxed));
			while (issued < ncut && done[issued].load(std::memory_order_acquire)) {
				if (issued == 0 && trace) tFirst = std::chrono::steady_clock::now();
				TRY(issue(issued++));
			}
			c = Chunk{i + 1, i + 1, off, static_cast<uint32_t>(words), static_cast<uint32_t>(words), 0};
		}
	}
	cutDone.store(true, std::memory_order_release);
	ctx->wordOff[nent] = static_cast<uint32_t>(words);
	const auto tQueued = std::chrono::steady_clock::now();
	for (; issued < ncut; ++issued) {
		while (!done[issued].load(std::memory_order_acquire)) std::this_thread::yield();
		TRY(issue(issued));
	}
	// end-of-stream guard words
	hCodes[ctx->nwords] = hCodes[ctx->nwords + 1] = 0;
	hMask[ctx->nwords] = hMask[ctx->nwords + 1] = 0xFFFFFFFFu;
	CK(cudaMemcpyAsync(ctx->d_codes.p + ctx->nwords, hCodes + ctx->nwords, 16, cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaMemcpyAsync(ctx->d_nmask.p + ctx->nwords, hMask + ctx->nwords, 8, cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));  // the staging buffer is reused by the next load
	ctx->notXPtr = ctx->notX.data();
	if (badSeq) return ctx->fail("sequence %u holds a code outside 0..4", badSeq);
	if (overflow) return ctx->fail("more than 2^31-1 residues: the reference's int totals overflow (Alignment.h:40-41)");
	if (trace) {
		const auto tEnd = std::chrono::steady_clock::now();
		std::fprintf(stderr, "xxm load_seqset (host pack): %u chunks, %d packing threads, first copy issued after %.2f ms, all cut after %.2f ms, done after %.2f ms\n", ncut,
		             nthreads ? nthreads : 1, std::chrono::duration<double, std::milli>(tFirst - tStart).count(),
		             std::chrono::duration<double, std::milli>(tQueued - tStart).count(), std::chrono::duration<double, std::milli>(tEnd - tStart).count());
	}
	return 0;
}

// The packing step on the device, for hosts with few cores per GPU: the raw codes travel in chunks of whole sequences on a
// copy stream (1 byte per base over PCIe); behind every chunk the per-sequence statistics (non-N residues, range check) and
// the packing kernel run on the compute stream, and the host folds the statistics of a chunk into total[] / avgLength[]
// while the next chunks are still on the bus.
static int load_pack_on_device(xxm_ctx* ctx, int32_t nent, const int32_t* n, const uint8_t* codes, int32_t revcomp, uint64_t rawTotal, bool trace,
                               std::chrono::steady_clock::time_point tStart) {
	uint64_t off = 0, words = 0;
	if (!ctx->copyStream) CK(cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
	CK(ctx->d_raw.ensure(std::max<uint64_t>(1, rawTotal)));
	CK(ctx->d_notX.ensure(nent));
	CK(ctx->counters.ensure(8));
	CK(cudaMemsetAsync(ctx->counters.p, 0, 32, ctx->stream));
	// pinned landing buffer of the per-sequence counts
	if (ctx->h_notXCap < static_cast<size_t>(nent)) {
		if (ctx->h_notX) cudaFreeHost(ctx->h_notX);
		ctx->h_notX = nullptr;
		ctx->h_notXCap = 0;
		CK(cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_notX), static_cast<size_t>(nent) * 4, cudaHostAllocDefault));
		ctx->h_notXCap = static_cast<size_t>(nent);
	}
	// ---- the same layout on the host (it owns wordOff / len / total), cutting chunks of whole sequences as it goes; every
	// chunk is queued as soon as it is cut: copy(k) on the copy stream, stats(k) + counts-to-host(k) + encode(k) on the compute stream
	constexpr uint64_t kChunkBytes = 48ull << 20;
	struct Chunk { int s0, s1; uint64_t b0, b1; uint32_t w0, w1; };
	std::vector<Chunk> chunks;
	auto tFirst = tStart;
	{
		Chunk c{0, 0, 0, 0, 0, 0};
		off = 0;
		words = 0;
		for (int i = 0; i < nent; ++i) {
			const int n0 = n[i];
			const int nstore = revcomp ? 2 * n0 + 1 : n0;
			ctx->len[i] = nstore;
			ctx->wordOff[i] = static_cast<uint32_t>(words);
			words += (static_cast<uint64_t>(nstore) + kPadBases + 31) / 32;
			off += static_cast<uint64_t>(n0);
			if (off - c.b0 < kChunkBytes && i + 1 != nent) continue;
			c.s1 = i + 1; c.b1 = off; c.w1 = static_cast<uint32_t>(words);
			const size_t k = chunks.size();
			chunks.push_back(c);
			while (ctx->evPool.size() < 2 * (k + 1)) {
				cudaEvent_t e;
				CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
				ctx->evPool.push_back(e);
			}
			if (trace && k == 0) { CK(cudaEventRecord(ctx->ev0, ctx->copyStream)); tFirst = std::chrono::steady_clock::now(); }
			CK(cudaMemcpyAsync(ctx->d_raw.p + c.b0, codes + c.b0, c.b1 - c.b0, cudaMemcpyHostToDevice, ctx->copyStream));
			CK(cudaEventRecord(ctx->evPool[2 * k], ctx->copyStream));
			if (trace && i + 1 == nent) CK(cudaEventRecord(ctx->ev1, ctx->copyStream));
			CK(cudaStreamWaitEvent(ctx->stream, ctx->evPool[2 * k], 0));
			const uint32_t ns = static_cast<uint32_t>(c.s1 - c.s0);
			k_seq_stats<<<std::min<uint32_t>(cdiv(ns, 8), static_cast<uint32_t>(ctx->numSMs) * 8), 256, 0, ctx->stream>>>(
				ctx->d_raw.p, ctx->d_rawOff.p, static_cast<uint32_t>(c.s0), static_cast<uint32_t>(c.s1), ctx->d_notX.p, ctx->counters.p);
			CKK();
			CK(cudaMemcpyAsync(ctx->h_notX + c.s0, ctx->d_notX.p + c.s0, static_cast<size_t>(ns) * 4, cudaMemcpyDeviceToHost, ctx->stream));
			CK(cudaEventRecord(ctx->evPool[2 * k + 1], ctx->stream));
			const uint32_t w1 = (i + 1 == nent) ? ctx->nwords + 2 : c.w1;
			if (w1 > c.w0) {
				k_encode<<<cdiv(w1 - c.w0, 256), 256, 0, ctx->stream>>>(ctx->d_raw.p, ctx->d_rawOff.p, ctx->S, c.w0, w1, ctx->d_codes.p, ctx->d_nmask.p, revcomp ? 1 : 0);
				CKK();
			}
			c = Chunk{i + 1, i + 1, off, off, static_cast<uint32_t>(words), static_cast<uint32_t>(words)};
		}
	}
	ctx->wordOff[nent] = static_cast<uint32_t>(words);
	const auto tQueued = std::chrono::steady_clock::now();
	// ---- calculateSequenceFeatures (seqset.cpp:21-36) on the raw set, then -- with revcomp -- the recount at the end of
	// createRevcomp (seqset.cpp:165-177) which compares codes with the CHARACTER '0' and therefore counts every
	// position of the doubled sequence.
	double acc = 0, lastLog = 0;
	int lastNotX = 0;
	int64_t tot = 0;
	for (size_t k = 0; k < chunks.size(); ++k) {
		CK(cudaEventSynchronize(ctx->evPool[2 * k + 1]));
		for (int i = chunks[k].s0; i < chunks[k].s1; ++i) {
			int notX = ctx->h_notX[i];
			if (revcomp) notX = ctx->len[i];
			if (notX != 0) {
				if (notX != lastNotX) { lastNotX = notX; lastLog = std::log(notX); }  // same argument, same glibc result
				acc += lastLog;
			}
			tot += notX;
			if (tot > 2147483647LL) { cudaStreamSynchronize(ctx->stream); return ctx->fail("more than 2^31-1 residues: the reference's int totals overflow (Alignment.h:40-41)"); }
			ctx->total[i + 1] = static_cast<int32_t>(tot);
			ctx->logAcc[i + 1] = acc;
		}
	}
	uint32_t bad = 0;
	CK(cudaMemcpyAsync(ctx->h_pin + 4, ctx->counters.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	std::memcpy(&bad, ctx->h_pin + 4, 4);
	ctx->notXPtr = ctx->h_notX;
	if (bad) return ctx->fail("sequence %u holds a code outside 0..4", bad);
	if (ctx->d_raw.cap > (8ull << 30)) ctx->d_raw.release();  // the raw codes are only needed for packing
	if (trace) {
		const auto tEnd = std::chrono::steady_clock::now();
		float copyMs = 0;
		cudaEventElapsedTime(&copyMs, ctx->ev0, ctx->ev1);
		std::fprintf(stderr, "xxm load_seqset (device pack): %zu chunks, first copy issued after %.2f ms, all queued after %.2f ms, copies took %.2f ms, done after %.2f ms\n",
		             chunks.size(), std::chrono::duration<double, std::milli>(tFirst - tStart).count(),
		             std::chrono::duration<double, std::milli>(tQueued - tStart).count(), copyMs, std::chrono::duration<double, std::milli>(tEnd - tStart).count());
	}
	return 0;
}

// Both at once, for hosts with a few cores per GPU (8 ranks on a 32-core box): the bus carries raw chunks from the FRONT of the
// set (packed behind it by k_seq_stats / k_encode, as in load_pack_on_device) while the worker threads pack chunks from the BACK
// (as in load_pack_on_host) and their packed words follow at 0.375 B per base; the two meet wherever the hardware of the day puts
// the balance.  At most kRawDepth raw chunks are in flight, so the bus is never committed far ahead of the packers.
static int load_pack_hybrid(xxm_ctx* ctx, int32_t nent, const int32_t* n, const uint8_t* codes, int32_t revcomp, uint64_t rawTotal, int nthreads, bool trace,
                            std::chrono::steady_clock::time_point tStart) {
	if (!ctx->copyStream) CK(cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
	CK(ctx->d_raw.ensure(std::max<uint64_t>(1, rawTotal)));
	CK(ctx->d_notX.ensure(nent));
	CK(ctx->counters.ensure(8));
	CK(cudaMemsetAsync(ctx->counters.p, 0, 32, ctx->stream));
	if (ctx->h_notXCap < static_cast<size_t>(nent)) {
		if (ctx->h_notX) cudaFreeHost(ctx->h_notX);
		ctx->h_notX = nullptr;
		ctx->h_notXCap = 0;
		CK(cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_notX), static_cast<size_t>(nent) * 4, cudaHostAllocDefault));
		ctx->h_notXCap = static_cast<size_t>(nent);
	}
	const size_t wordsAll = static_cast<size_t>(ctx->nwords) + 2;
	if (ctx->h_packCap < wordsAll) {
		if (ctx->h_pack) cudaFreeHost(ctx->h_pack);
		ctx->h_pack = nullptr;
		ctx->h_packCap = 0;
		CK(cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_pack), wordsAll * 12, cudaHostAllocDefault));
		ctx->h_packCap = wordsAll;
	}
	uint64_t* const hCodes = reinterpret_cast<uint64_t*>(ctx->h_pack);
	uint32_t* const hMask = reinterpret_cast<uint32_t*>(ctx->h_pack + ctx->h_packCap * 8);
	ctx->notX.resize(nent);
	// ---- chunks of whole sequences, cut by this thread as it lays the stream out (it owns wordOff / len)
	constexpr uint64_t kChunkBytes = 8ull << 20;
	constexpr int kRawDepth = 3;
	struct Chunk { int s0, s1; uint64_t b0, b1; uint32_t w0, w1; uint32_t bad; };
	std::vector<Chunk> chunks;
	chunks.reserve(static_cast<size_t>(rawTotal / kChunkBytes) + 2);  // the workers read it while nothing is appended any more
	// who takes which chunk: the bus from the front, the packers (started once the cut is complete) from the back
	std::mutex claimMu;
	int front = 0, back = 0;  // [front, back) is cut and unclaimed
	auto claimFront = [&]() { std::lock_guard<std::mutex> g(claimMu); return front < back ? front++ : -1; };
	auto claimBack = [&]() { std::lock_guard<std::mutex> g(claimMu); return back > front ? --back : -1; };
	int rawIssued = 0, rawDone = 0;   // raw chunks are 0 .. rawIssued-1
	auto issueRaw = [&](int k) -> int {
		while (ctx->evPool.size() < 2 * static_cast<size_t>(k + 1)) {
			cudaEvent_t e;
			CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
			ctx->evPool.push_back(e);
		}
		const Chunk& c = chunks[k];
		CK(cudaMemcpyAsync(ctx->d_raw.p + c.b0, codes + c.b0, c.b1 - c.b0, cudaMemcpyHostToDevice, ctx->copyStream));
		CK(cudaEventRecord(ctx->evPool[2 * k], ctx->copyStream));
		CK(cudaStreamWaitEvent(ctx->stream, ctx->evPool[2 * k], 0));
		const uint32_t ns = static_cast<uint32_t>(c.s1 - c.s0);
		k_seq_stats<<<std::min<uint32_t>(cdiv(ns, 8), static_cast<uint32_t>(ctx->numSMs) * 8), 256, 0, ctx->stream>>>(
			ctx->d_raw.p, ctx->d_rawOff.p, static_cast<uint32_t>(c.s0), static_cast<uint32_t>(c.s1), ctx->d_notX.p, ctx->counters.p);
		CKK();
		CK(cudaMemcpyAsync(ctx->h_notX + c.s0, ctx->d_notX.p + c.s0, static_cast<size_t>(ns) * 4, cudaMemcpyDeviceToHost, ctx->stream));
		if (c.w1 > c.w0) {
			k_encode<<<cdiv(c.w1 - c.w0, 256), 256, 0, ctx->stream>>>(ctx->d_raw.p, ctx->d_rawOff.p, ctx->S, c.w0, c.w1, ctx->d_codes.p, ctx->d_nmask.p, revcomp ? 1 : 0);
			CKK();
		}
		CK(cudaEventRecord(ctx->evPool[2 * k + 1], ctx->stream));
		return 0;
	};
	// calculateSequenceFeatures / createRevcomp recount, in sequence order (see load_pack_on_host): the raw chunks are folded as
	// their counts arrive, the packed ones after the loop
	double acc = 0, lastLog = 0;
	int lastNotX = 0;
	int64_t tot = 0;
	bool overflow = false;
	auto fold = [&](int k, bool raw) {
		const Chunk& c = chunks[k];
		if (raw) std::memcpy(ctx->notX.data() + c.s0, ctx->h_notX + c.s0, static_cast<size_t>(c.s1 - c.s0) * 4);
		for (int i = c.s0; i < c.s1; ++i) {
			int notX = ctx->notX[i];
			if (revcomp) notX = ctx->len[i];
			if (notX != 0) {
				if (notX != lastNotX) { lastNotX = notX; lastLog = std::log(notX); }  // same argument, same glibc result
				acc += lastLog;
			}
			tot += notX;
			if (tot > 2147483647LL) overflow = true;
			ctx->total[i + 1] = static_cast<int32_t>(tot);
			ctx->logAcc[i + 1] = acc;
		}
	};
	// one turn of the bus side: fold what has arrived, put the next raw chunk on the bus if fewer than kRawDepth are in flight
	bool claimsOpen = true, cutComplete = false;
	auto pumpRaw = [&](bool* progress) -> int {
		while (rawDone < rawIssued && cudaEventQuery(ctx->evPool[2 * rawDone + 1]) == cudaSuccess) { fold(rawDone++, true); *progress = true; }
		if (claimsOpen && rawIssued - rawDone < kRawDepth) {
			const int k = claimFront();
			if (k >= 0) { TRY(issueRaw(k)); rawIssued = k + 1; *progress = true; }
			else if (cutComplete) claimsOpen = false;  // front met back
		}
		return 0;
	};
	{
		uint64_t off = 0, words = 0;
		Chunk c{0, 0, 0, 0, 0, 0, 0};
		for (int i = 0; i < nent; ++i) {
			const int n0 = n[i];
			const int nstore = revcomp ? 2 * n0 + 1 : n0;
			ctx->len[i] = nstore;
			ctx->wordOff[i] = static_cast<uint32_t>(words);
			words += (static_cast<uint64_t>(nstore) + kPadBases + 31) / 32;
			off += static_cast<uint64_t>(n0);
			if (off - c.b0 < kChunkBytes && i + 1 != nent) continue;
			c.s1 = i + 1; c.b1 = off; c.w1 = static_cast<uint32_t>(words);
			chunks.push_back(c);
			// the bus starts while the rest is still being cut -- but only chunks of the first half, the packers get their share
			if (chunks.size() * kChunkBytes < rawTotal / 2) {
				back = static_cast<int>(chunks.size());
				bool dummy = false;
				TRY(pumpRaw(&dummy));
			}
			c = Chunk{i + 1, i + 1, off, off, static_cast<uint32_t>(words), static_cast<uint32_t>(words), 0};
		}
		ctx->wordOff[nent] = static_cast<uint32_t>(words);
	}
	const int nchunks = static_cast<int>(chunks.size());
	back = nchunks;  // no worker is running yet
	cutComplete = true;
	std::unique_ptr<std::atomic<uint32_t>[]> done(new std::atomic<uint32_t>[static_cast<size_t>(nchunks)]);
	for (int k = 0; k < nchunks; ++k) done[k].store(0, std::memory_order_relaxed);
	const bool rc = revcomp != 0;
	auto worker = [&]() {
		for (int k = claimBack(); k >= 0; k = claimBack()) {
			Chunk& c = chunks[k];
			c.bad = pack_range(n, static_cast<uint32_t>(c.s0), static_cast<uint32_t>(c.s1), codes + c.b0, rawTotal - c.b0, rc, hCodes + c.w0, hMask + c.w0, ctx->notX.data() + c.s0);
			done[k].store(1, std::memory_order_release);
		}
	};
	struct Joiner {
		std::vector<std::thread> t;
		~Joiner() { for (auto& th : t) if (th.joinable()) th.join(); }
	} pool;
	for (int t = 0; t < nthreads; ++t) pool.t.emplace_back(worker);
	// ---- this thread: raw chunks onto the bus (bounded depth), packed chunks onto the bus as they finish
	int packedIssued = nchunks;       // packed chunks packedIssued .. nchunks-1 have been copied
	auto issuePacked = [&](int k) -> int {
		const Chunk& c = chunks[k];
		if (c.w1 > c.w0) {
			CK(cudaMemcpyAsync(ctx->d_codes.p + c.w0, hCodes + c.w0, static_cast<size_t>(c.w1 - c.w0) * 8, cudaMemcpyHostToDevice, ctx->copyStream));
			CK(cudaMemcpyAsync(ctx->d_nmask.p + c.w0, hMask + c.w0, static_cast<size_t>(c.w1 - c.w0) * 4, cudaMemcpyHostToDevice, ctx->copyStream));
		}
		return 0;
	};
	for (;;) {
		bool progress = false;
		TRY(pumpRaw(&progress));
		// the packers work downwards from the end: copy whatever is finished below what has been copied already
		while (packedIssued > rawIssued && done[packedIssued - 1].load(std::memory_order_acquire)) { TRY(issuePacked(--packedIssued)); progress = true; }
		if (!claimsOpen && packedIssued == rawIssued) break;
		if (!progress) std::this_thread::yield();
	}
	const int firstPacked = rawIssued;
	hCodes[ctx->nwords] = hCodes[ctx->nwords + 1] = 0;
	hMask[ctx->nwords] = hMask[ctx->nwords + 1] = 0xFFFFFFFFu;
	CK(cudaStreamSynchronize(ctx->stream));  // k_encode may have touched the guard words' neighbours; the guards go last
	CK(cudaMemcpyAsync(ctx->d_codes.p + ctx->nwords, hCodes + ctx->nwords, 16, cudaMemcpyHostToDevice, ctx->copyStream));
	CK(cudaMemcpyAsync(ctx->d_nmask.p + ctx->nwords, hMask + ctx->nwords, 8, cudaMemcpyHostToDevice, ctx->copyStream));
	uint32_t badSeq = 0;
	for (; rawDone < rawIssued; ++rawDone) fold(rawDone, true);  // the compute stream is idle: every count has landed
	for (int k = firstPacked; k < nchunks; ++k) {
		if (chunks[k].bad && !badSeq) badSeq = chunks[k].bad;
		fold(k, false);
	}
	if (overflow) { cudaStreamSynchronize(ctx->copyStream); return ctx->fail("more than 2^31-1 residues: the reference's int totals overflow (Alignment.h:40-41)"); }
	uint32_t bad = 0;
	CK(cudaMemcpyAsync(ctx->h_pin + 4, ctx->counters.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	CK(cudaStreamSynchronize(ctx->copyStream));  // the staging buffer is reused by the next load; the compute stream must see the packed words
	std::memcpy(&bad, ctx->h_pin + 4, 4);
	ctx->notXPtr = ctx->notX.data();
	if (bad && (!badSeq || bad < badSeq)) badSeq = bad;
	if (badSeq) return ctx->fail("sequence %u holds a code outside 0..4", badSeq);
	if (ctx->d_raw.cap > (8ull << 30)) ctx->d_raw.release();
	if (trace) {
		const auto tEnd = std::chrono::steady_clock::now();
		std::fprintf(stderr, "xxm load_seqset (hybrid): %d chunks, %d over the bus raw + %d packed by %d host threads, done after %.2f ms\n", nchunks, firstPacked,
		             nchunks - firstPacked, nthreads, std::chrono::duration<double, std::milli>(tEnd - tStart).count());
	}
	ctx->lastLoadRawChunks = firstPacked;
	ctx->lastLoadChunks = nchunks;
	return 0;
}

// a1.  Layout from the lengths, then the packing step where it is cheaper: on the host cores when there are at least
// kMinPackThreads of them for this GPU (XXM_PACK_THREADS, default: cores - 2 up to 14 on hosts with 12 or more cores, else 3/4 of the cores), else on the device.
extern "C" int xxm_load_seqset(xxm_ctx* ctx, int32_t nent, const int32_t* n, const uint8_t* codes, int32_t revcomp) {
	XXM_NVTX("xxm_load_seqset");
	CK(cudaSetDevice(ctx->device));
	if (nent <= 0) return ctx->fail("empty sequence set");
	if (ctx->views.load() > 0) return ctx->fail("xxm_load_seqset: %d view(s) share this context's set (xxm_share_seqset); destroy or re-point them first", ctx->views.load());
	CK(cudaStreamSynchronize(ctx->stream));
	drop_shared_seqset(ctx);  // a view that loads its own set stops being one (it must never write through borrowed pointers)
	static const bool trace = std::getenv("XXM_TRACE") != nullptr;
	const auto tStart = std::chrono::steady_clock::now();
	// whatever happens below, the previous set is gone: a failed load must not leave a valid-looking model over a partly
	// overwritten stream (nent is restored only on success)
	ctx->haveModel = false;
	ctx->loadSerial++;
	ctx->deadKey.clear();
	for (auto& l : ctx->L) l = SiteList();
	ctx->keptElong = SiteList();
	ctx->G.external = false;
	ctx->nBest = 0;
	ctx->nent = 0;
	ctx->pend.active = false;
	ctx->revcomp = revcomp != 0;
	ctx->len.resize(nent);
	ctx->total.resize(static_cast<size_t>(nent) + 1);
	ctx->logAcc.resize(static_cast<size_t>(nent) + 1);
	ctx->total[0] = 0;
	ctx->logAcc[0] = 0.0;
	ctx->wordOff.resize(static_cast<size_t>(nent) + 1);
	// ---- layout on the device from the lengths alone
	CK(ctx->d_n.ensure(nent));
	CK(ctx->d_rawOff.ensure(static_cast<size_t>(nent) + 1));
	CK(ctx->d_wordOff.ensure(static_cast<size_t>(nent) + 1));
	CK(ctx->d_len.ensure(nent));
	CK(cudaMemcpyAsync(ctx->d_n.p, n, static_cast<size_t>(nent) * 4, cudaMemcpyHostToDevice, ctx->stream));
	{
		const uint32_t nb = cdiv(static_cast<uint32_t>(nent), 1024);
		CK(ctx->d_layoutSums.ensure(2ull * nb));
		k_seq_layout_sums<<<nb, 1024, 0, ctx->stream>>>(ctx->d_n.p, static_cast<uint32_t>(nent), revcomp ? 1 : 0, ctx->d_layoutSums.p);
		CKK();
		k_seq_layout<<<nb, 1024, 0, ctx->stream>>>(ctx->d_n.p, static_cast<uint32_t>(nent), revcomp ? 1 : 0, ctx->d_layoutSums.p, ctx->d_rawOff.p,
		                                            ctx->d_wordOff.p, ctx->d_len.p);
		CKK();
	}
	// ---- sizes first (no stores: a fraction of a millisecond per million sequences), so that the buffers exist before the first chunk leaves
	uint64_t off = 0, words = 0;
	int mx = 0, mn = 2147483647;
	{
		int mxRaw = 0;
		uint64_t w0 = 0;
		for (int i = 0; i < nent; ++i) {  // branch free: vectorises
			const int n0 = n[i];
			mxRaw = std::max(mxRaw, n0);
			mn = std::min(mn, n0);
			off += static_cast<uint32_t>(n0);
			w0 += (static_cast<uint32_t>(n0) + static_cast<uint32_t>(kPadBases + 31)) >> 5;
		}
		if (mn < 1) {
			for (int i = 0; i < nent; ++i) if (n[i] < 1) return ctx->fail("sequence %d is empty", i + 1);
		}
		if (mxRaw > 1000000000) return ctx->fail("sequence longer than 10^9 residues");
		if (revcomp) {
			words = 0;
			for (int i = 0; i < nent; ++i) words += (2ull * static_cast<uint32_t>(n[i]) + 1 + kPadBases + 31) / 32;
			mx = 2 * mxRaw + 1;
			mn = 2 * mn + 1;
		} else {
			words = w0;
			mx = mxRaw;
		}
	}
	if (words > 0xFFFFFF00ull) return ctx->fail("sequence stream exceeds 2^32 words");
	const uint64_t rawTotal = off;
	ctx->nwords = static_cast<uint32_t>(words);
	ctx->max_leng = mx;
	ctx->min_leng = mn;
	const uint32_t nblk = ctx->nwords / kBlkWords + 1;
	CK(ctx->d_codes.ensure(static_cast<size_t>(ctx->nwords) + 2));
	CK(ctx->d_nmask.ensure(static_cast<size_t>(ctx->nwords) + 2));
	CK(ctx->d_blk2seq.ensure(nblk));
	ctx->S.codes = ctx->d_codes.p;
	ctx->S.nmask = ctx->d_nmask.p;
	ctx->S.wordOff = ctx->d_wordOff.p;
	ctx->S.len = ctx->d_len.p;
	ctx->S.blk2seq = ctx->d_blk2seq.p;
	ctx->S.nwords = ctx->nwords;
	ctx->S.nent = nent;
	ctx->S.uniformWps = (mx == mn) ? static_cast<uint32_t>((static_cast<uint64_t>(mx) + kPadBases + 31) / 32) : 0u;
	ctx->S.uniformLen = (mx == mn) ? mx : 0;
	k_blk2seq<<<cdiv(static_cast<uint32_t>(nent), 256), 256, 0, ctx->stream>>>(ctx->d_wordOff.p, static_cast<uint32_t>(nent), nblk, ctx->d_blk2seq.p);
	CKK();
	const int nthreads = pack_threads_default();
	constexpr int kMinPackThreads = 6;  // 3 threads pack ~17 GB/s of bases, the bus takes raw bytes at 20-48 GB/s (profiles/README.md)
	// XXM_LOAD_PATH=host|device|hybrid overrides the choice (tests, measurements)
	const char* forced = std::getenv("XXM_LOAD_PATH");
	const bool small = rawTotal < (4ull << 20);
	// (the hybrid of both is opt-in: on 1- and 2-GPU boxes, where the bus takes raw bytes at 42-51 GB/s, 3-5 packing threads did not
	// shorten the load -- profiles/r02_load_paths.json)
	int path = (nthreads >= kMinPackThreads || small) ? 0 : 1;
	if (forced && !small) path = !std::strcmp(forced, "host") ? 0 : !std::strcmp(forced, "device") ? 1 : !std::strcmp(forced, "hybrid") ? 2 : path;
	ctx->lastLoadRawChunks = ctx->lastLoadChunks = 0;
	if (path == 0) TRY(load_pack_on_host(ctx, nent, n, codes, revcomp, rawTotal, small ? 0 : nthreads, trace, tStart));
	else if (path == 1) TRY(load_pack_on_device(ctx, nent, n, codes, revcomp, rawTotal, trace, tStart));
	else TRY(load_pack_hybrid(ctx, nent, n, codes, revcomp, rawTotal, std::max(1, nthreads), trace, tStart));
	if (ctx->sum_i.size() != static_cast<size_t>(ctx->max_leng) + 1) init_sum_i(ctx->sum_i, ctx->max_leng);
	ctx->nent = nent;
	return 0;
}

// how the last xxm_load_seqset moved the set: chunks that crossed the bus as raw bytes (packed on the device) out of all chunks;
// 0 of 0 when the whole set was packed on the host cores or on the device
extern "C" int xxm_last_load_info(const xxm_ctx* ctx, int32_t* rawChunks, int32_t* chunks) {
	*rawChunks = ctx->lastLoadRawChunks;
	*chunks = ctx->lastLoadChunks;
	return 0;
}

extern "C" int xxm_seqset_info(const xxm_ctx* ctx, int32_t* total, double* avgLength, int32_t* max_leng, int32_t* min_leng) {
	if (total) std::memcpy(total, ctx->total.data(), ctx->total.size() * sizeof(int32_t));
	if (avgLength)
		for (size_t i = 0; i < ctx->logAcc.size(); ++i) avgLength[i] = ctx->avg_length(static_cast<int>(i));
	if (max_leng) *max_leng = ctx->max_leng;
	if (min_leng) *min_leng = ctx->min_leng;
	return 0;
}

// =====================================================================================================================
// a5: background + model
// =====================================================================================================================
extern "C" int xxm_set_background(xxm_ctx* ctx, int32_t order, const float* rprobs, const float* rconds) {
	XXM_NVTX("xxm_set_background");
	if (order < 0 || order > 9) return ctx->fail("background order %d not supported", order);
	if (ctx->views.load() > 0) return ctx->fail("xxm_set_background: %d view(s) read this context's NullModel tables (xxm_share_seqset)", ctx->views.load());
	const size_t n = static_cast<size_t>(1) << (3 * (order + 1));
	ctx->rprobs.assign(rprobs, rprobs + n);
	ctx->rconds.assign(rconds, rconds + n);
	ctx->nm.order = order;
	ctx->nm.rprobs = ctx->rprobs.data();
	ctx->nm.rconds = ctx->rconds.data();
	ctx->bgCache.clear();  // background probabilities of k-mers under the previous model
	ctx->nmOnDevice = false;
	ctx->haveModel = false;
	return 0;
}

#include "xxm_stage.cuh"  // kMaxUpSegs, UpSegs, k_scatter_stage

static int begin_uploads(xxm_ctx* ctx) {
	if (ctx->stageInFlight) { CK(cudaEventSynchronize(ctx->evStage)); ctx->stageInFlight = false; }
	ctx->batching = true;
	ctx->upSegs.clear();
	ctx->fillJobs.clear();
	ctx->stageUsed = 0;
	return 0;
}

// copy what has been staged so far and run the kernels that were waiting for it
static int flush_uploads(xxm_ctx* ctx) {
	if (!ctx->upSegs.empty()) {
		CK(ctx->d_stage.ensure(ctx->stageUsed));
		CK(cudaMemcpyAsync(ctx->d_stage.p, ctx->h_stage, ctx->stageUsed, cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaEventRecord(ctx->evStage, ctx->stream));
		ctx->stageInFlight = true;
		for (size_t b = 0; b < ctx->upSegs.size(); b += kMaxUpSegs) {
			UpSegs U{};
			const int ns = static_cast<int>(std::min<size_t>(kMaxUpSegs, ctx->upSegs.size() - b));
			uint32_t mx = 4;
			for (int i = 0; i < ns; ++i) {
				U.dst[i] = ctx->upSegs[b + i].dst; U.off[i] = ctx->upSegs[b + i].off; U.bytes[i] = ctx->upSegs[b + i].bytes;
				mx = std::max(mx, U.bytes[i]);
			}
			k_scatter_stage<<<dim3(std::min<uint32_t>(64, cdiv(mx / 4, 256)), ns), 256, 0, ctx->stream>>>(ctx->d_stage.p, U);
			CKK();
		}
		ctx->upSegs.clear();
	}
	for (const xxm_ctx::FillJob& j : ctx->fillJobs) {
		k_fill_rank<<<cdiv(j.n, 256), 256, 0, ctx->stream>>>(ctx->d_ids[j.k].p, ctx->d_cls[j.k].p, ctx->d_clsOk[j.k].p, j.n, j.rank, j.bitmap);
		CKK();
	}
	ctx->fillJobs.clear();
	return 0;
}

static int end_uploads(xxm_ctx* ctx) {
	const int r = flush_uploads(ctx);
	ctx->batching = false;
	// the pinned buffer is reused by the next batch: begin_uploads waits for this copy first
	ctx->stageUsed = 0;
	return r;
}

template <typename T>
static int upload(xxm_ctx* ctx, DevBuf<T>& d, const std::vector<T>& h) {
	CK(d.ensure(std::max<size_t>(1, h.size())));
	if (h.empty()) return 0;
	const size_t bytes = h.size() * sizeof(T);
	if (!ctx->batching || bytes > (1u << 20)) {
		CK(cudaMemcpyAsync(d.p, h.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
		return 0;
	}
	const size_t off = (ctx->stageUsed + 15) & ~static_cast<size_t>(15);
	if (off + bytes > ctx->h_stageCap) {
		// grow the pinned buffer; segments already staged in this batch are copied first
		TRY(flush_uploads(ctx));
		CK(cudaEventSynchronize(ctx->evStage));
		ctx->stageInFlight = false;
		if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
		ctx->h_stage = nullptr;
		ctx->h_stageCap = std::max<size_t>(4u << 20, 2 * (off + bytes));
		CK(cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_stage), ctx->h_stageCap, cudaHostAllocDefault));
		ctx->stageUsed = 0;
		return upload(ctx, d, h);
	}
	std::memcpy(ctx->h_stage + off, h.data(), bytes);
	ctx->stageUsed = off + bytes;
	ctx->upSegs.push_back(xxm_ctx::UpSeg{d.p, static_cast<uint32_t>(off), static_cast<uint32_t>(bytes)});
	return 0;
}

// k = 0 / 1: scratch buffers of the left / right table part
static int stage_part(xxm_ctx* ctx, const TablePart& P, bool single, DevBuf<uint32_t>& rank, DevBuf<uint32_t>* bitmap, int k) {
	const size_t size = static_cast<size_t>(1) << (2 * P.len);
	CK(rank.ensure(size));
	CK(cudaMemsetAsync(rank.p, 0, size * 4, ctx->stream));
	if (bitmap) {
		CK(bitmap->ensure(std::max<size_t>(1, size / 32)));
		CK(cudaMemsetAsync(bitmap->p, 0, std::max<size_t>(1, size / 32) * 4, ctx->stream));
	}
	std::vector<uint8_t> ok(P.classPval.size());
	for (size_t c = 0; c < ok.size(); ++c) ok[c] = (P.classPval[c] != 0.0) && !(single && P.classPval[c] == 1.0);
	TRY(upload(ctx, ctx->d_ids[k], P.ids));
	TRY(upload(ctx, ctx->d_cls[k], P.cls));
	TRY(upload(ctx, ctx->d_clsOk[k], ok));
	const int n = static_cast<int>(P.ids.size());
	if (n) ctx->fillJobs.push_back(xxm_ctx::FillJob{k, n, rank.p, bitmap ? bitmap->p : nullptr});
	if (!ctx->batching) TRY(flush_uploads(ctx));
	return 0;
}

static double zoops_final(double pS, double exponent) { return 1 - std::pow(1 - pS, exponent); }  // pValCalculation.h:96

static int device_build_part0(xxm_ctx* ctx, PvalueTables& T, float* ms);  // xxm_gpu_part8.inc
static int reinitialize_impl(xxm_ctx* ctx, const double* pwm, const int32_t* cols, int32_t ncols, int32_t posSetSize, int32_t totalPositions);
extern "C" int xxm_reinitialize(xxm_ctx* ctx, const double* pwm, const int32_t* cols, int32_t ncols, int32_t posSetSize, int32_t totalPositions) {
	XXM_NVTX("xxm_reinitialize");
	const int r = reinitialize_impl(ctx, pwm, cols, ncols, posSetSize, totalPositions);
	if (r) { ctx->batching = false; ctx->upSegs.clear(); ctx->fillJobs.clear(); ctx->haveModel = false; }
	return r;
}
// Order of the score classes under the mops finalPval table (what a full scan compares, StartPosUpdater-inl.h:99):
// updateMultOcc's test "(double)float(fin[kept]) <= fin[incoming]" (StartPosUtils.h:18) becomes "incoming >= T[kept]" on class
// indices when fin[] is non-decreasing in the class index -- checked here, k_select_cls_fwd is used only if it holds.
static int stage_class_order(xxm_ctx* ctx) {
	const std::vector<double>& fin = ctx->h_finalMops;
	const size_t nc = fin.size();
	ctx->clsOrderOK = nc > 0 && nc <= 4096;
	for (size_t c = 1; c < nc && ctx->clsOrderOK; ++c)
		if (!(fin[c - 1] <= fin[c])) ctx->clsOrderOK = false;
	if (!ctx->clsOrderOK) return 0;
	ctx->h_clsThr.resize(nc);
	for (size_t k = 0; k < nc; ++k) {
		const double f = static_cast<double>(static_cast<float>(fin[k]));
		ctx->h_clsThr[k] = static_cast<uint16_t>(std::lower_bound(fin.begin(), fin.end(), f) - fin.begin());
	}
	ctx->clsNSite = static_cast<uint32_t>(std::lower_bound(fin.begin(), fin.end(), 1.0) - fin.begin());
	TRY(upload(ctx, ctx->d_clsThr, ctx->h_clsThr));
	return 0;
}

static int reinitialize_impl(xxm_ctx* ctx, const double* pwm, const int32_t* cols, int32_t ncols, int32_t posSetSize, int32_t totalPositions) {
	CK(cudaSetDevice(ctx->device));
	if (!ctx->nent) return ctx->fail("no sequence set loaded");
	if (!ctx->nm.rprobs) return ctx->fail("no background model set");
	if (ncols < 1 || ncols > 26) return ctx->fail("%d match columns not supported (reference maximum is 26)", ncols);
	for (int i = 0; i < ncols; ++i)
		if (cols[i] < 1 || cols[i] > XXM_PWM_LENGTH || (i && cols[i] <= cols[i - 1])) return ctx->fail("invalid match columns");
	if (posSetSize < 1 || posSetSize > ctx->nent) return ctx->fail("posSetSize out of range");
	if (ctx->comm) {
		// a sharded set: the thresholds of the tables follow the WHOLE set (branch_and_bound-inl.h:186-187, :417)
		TRY(shard_sync(ctx));
		if (posSetSize != ctx->nent) return ctx->fail("with a communicator attached every call refers to the whole sharded set: posSetSize must be the rank's sequence count");
		build_pvalue_tables(pwm, cols, ncols, ctx->G.nent, ctx->G.total, ctx->opt.bgLog, ctx->nm, ctx->tab, &ctx->bgCache);
	} else if (ctx->G.external) {
		// the caller runs the exchange itself (split-phase entry points) and has told us the counts of the whole set
		if (ctx->G.loadSerial != ctx->loadSerial) return ctx->fail("xxm_set_global_counts must be repeated after xxm_load_seqset");
		if (posSetSize != ctx->nent) return ctx->fail("with global counts set every call refers to the whole sharded set: posSetSize must be the rank's sequence count");
		build_pvalue_tables(pwm, cols, ncols, ctx->G.nent, ctx->G.total, ctx->opt.bgLog, ctx->nm, ctx->tab, &ctx->bgCache);
	} else if (ncols <= 8 && std::getenv("XXM_DEVICE_TABLES")) {
		// the enumeration, the stable sort and the cumulative class p-values on the device
		model_header(pwm, cols, ncols, posSetSize, totalPositions, ctx->opt.bgLog, ctx->tab);
		if (ctx->tab.error == 0) TRY(device_build_part0(ctx, ctx->tab, nullptr));
	} else
	build_pvalue_tables(pwm, cols, ncols, posSetSize, totalPositions, ctx->opt.bgLog, ctx->nm, ctx->tab, &ctx->bgCache);
	const PvalueTables& T = ctx->tab;
	if (T.error == 1) return ctx->fail("gapped motif with >8 match columns and a left part spanning >10 positions: reference behaviour undefined");
	if (T.error == 2) return ctx->fail("PWM column not normalised (the reference asserts, branch_and_bound-inl.h:223)");
	if (T.part[0].len > 13 || (T.nparts == 2 && T.part[1].len > 13)) return ctx->fail("table part too long");
	ctx->modelSerial++;
	ctx->modelPosSetSize = posSetSize;
	ctx->modelFirstCol = cols[0];
	Model& M = ctx->M;
	std::memset(&M, 0, sizeof M);
	M.L = T.L; M.maxk = T.maxk; M.nparts = T.nparts; M.span = T.span;
	M.len0 = T.part[0].len; M.len1 = T.nparts == 2 ? T.part[1].len : 0;
	M.contiguous = (T.span == T.L);
	for (int k = 0; k < T.L; ++k) {
		M.coloff[k] = T.coloff[k];
		for (int a = 0; a < 4; ++a) M.logodds[k][a] = T.logodds[k + 1][a + 1];
	}
	M.pruneThr = T.pruneThr;
	M.corr = std::pow(ctx->opt.overrepCorrection, T.L);  // pOverrepCorrection[i] = pow(overrepCorrection, i), pValCalculation.cpp:23-24
	TRY(begin_uploads(ctx));
	TRY(stage_part(ctx, T.part[0], T.nparts == 1, ctx->d_rank0, &ctx->d_hit0, 0));
	M.rank0 = ctx->d_rank0.p;
	M.hit0 = ctx->d_hit0.p;
	TRY(upload(ctx, ctx->d_cpval0, T.part[0].classPval));
	TRY(upload(ctx, ctx->d_cscore0, T.part[0].classScore));
	M.cpval0 = ctx->d_cpval0.p; M.cscore0 = ctx->d_cscore0.p; M.nclass0 = static_cast<int>(T.part[0].classPval.size());
	const double exponent = set_avg_length(ctx, posSetSize) - T.span + 1;
	ctx->memoModeValid = false;
	ctx->posProbValid = false;
	if (T.nparts == 1) {
		const size_t nc = T.part[0].classPval.size();
		ctx->nval = static_cast<uint32_t>(nc);
		ctx->h_pS.resize(nc); ctx->h_finalMops.resize(nc);
		for (size_t c = 0; c < nc; ++c) {
			const double v = T.part[0].classPval[c];
			if (v == 0.0 || v == 1.0) { ctx->h_pS[c] = 1; ctx->h_finalMops[c] = 1; continue; }
			const double pS = v * M.corr;
			ctx->h_pS[c] = pS;
			ctx->h_finalMops[c] = pS;
		}
		// the ZOOPS / OOPS table (one glibc pow per class) is only built when a call with mops = false needs it
		// (ensure_mode_table): a full-scan step never does
		ctx->modeExponent = exponent;
		TRY(upload(ctx, ctx->d_pS, ctx->h_pS));
		TRY(upload(ctx, ctx->d_finalMops, ctx->h_finalMops));
		CK(ctx->d_finalMode.ensure(std::max<size_t>(1, nc)));
		TRY(stage_class_order(ctx));
	} else {
		ctx->clsOrderOK = false;
		TRY(stage_part(ctx, T.part[1], false, ctx->d_rank1, nullptr, 1));
		M.rank1 = ctx->d_rank1.p;
		TRY(upload(ctx, ctx->d_cpval1, T.part[1].classPval));
		TRY(upload(ctx, ctx->d_cscore1, T.part[1].classScore));
		TRY(upload(ctx, ctx->d_leftProb, T.leftClassProb));
		M.cpval1 = ctx->d_cpval1.p; M.cscore1 = ctx->d_cscore1.p; M.leftProb = ctx->d_leftProb.p;
		M.nclass1 = static_cast<int>(T.part[1].classPval.size());
		// memo capacity: distinct scores <= enumerated left ids x enumerated right ids
		// (... and <= the windows that can survive the product test: a generous 1/32 of the stored bases)
		uint64_t want = 4ull * std::min<uint64_t>(std::max<uint64_t>(1, T.part[0].ids.size()) * std::max<uint64_t>(1, T.part[1].ids.size()),
		                                          std::max<uint64_t>(1u << 16, static_cast<uint64_t>(ctx->nwords)));
		uint32_t cap = 1u << 12;
		while (cap < want && cap < (1u << 24)) cap <<= 1;
		CK(ctx->d_memoKey.ensure(cap));
		CK(ctx->d_memoVal.ensure(cap));
		CK(ctx->d_memoState.ensure(cap));
		CK(ctx->d_flags32.ensure(4));
		k_fill_u32<<<cdiv(cap, 256), 256, 0, ctx->stream>>>(reinterpret_cast<uint32_t*>(ctx->d_memoKey.p), static_cast<uint32_t>(kMemoEmpty), cap);
		CKK();
		CK(cudaMemsetAsync(ctx->d_memoState.p, 0, static_cast<size_t>(cap) * 4, ctx->stream));
		CK(cudaMemsetAsync(ctx->d_memoVal.p, 0, static_cast<size_t>(cap) * 8, ctx->stream));
		CK(cudaMemsetAsync(ctx->d_flags32.p, 0, 16, ctx->stream));
		M.memoKey = ctx->d_memoKey.p; M.memoVal = ctx->d_memoVal.p; M.memoState = ctx->d_memoState.p; M.memoCap = cap;
		M.memoOverflow = ctx->d_flags32.p;
		ctx->nval = cap;
		CK(ctx->d_pS.ensure(cap));
		CK(ctx->d_finalMops.ensure(cap));
		CK(ctx->d_finalMode.ensure(cap));
	}
	TRY(stage_scan(ctx));
	TRY(end_uploads(ctx));
	ctx->haveModel = true;
	return 0;
}

extern "C" int xxm_tables_info(const xxm_ctx* ctx, int32_t* L, int32_t* maxk, int32_t* nparts, int64_t size[2], int32_t nclasses[2], int64_t nids[2]) {
	const PvalueTables& T = ctx->tab;
	*L = T.L; *maxk = T.maxk; *nparts = T.nparts;
	for (int p = 0; p < 2; ++p) {
		size[p] = p < T.nparts ? (static_cast<int64_t>(1) << (2 * T.part[p].len)) : 0;
		nclasses[p] = p < T.nparts ? static_cast<int32_t>(T.part[p].classScore.size()) : 0;
		nids[p] = p < T.nparts ? static_cast<int64_t>(T.part[p].ids.size()) : 0;
	}
	return 0;
}

extern "C" int xxm_tables_get(const xxm_ctx* ctx, int32_t part, double* pvalues, double* probs, int32_t* classScore, int32_t* classCount,
                              int64_t* classIds) {
	const PvalueTables& T = ctx->tab;
	if (part < 0 || part >= T.nparts) return 1;
	const TablePart& P = T.part[part];
	const size_t size = static_cast<size_t>(1) << (2 * P.len);
	if (pvalues) std::fill(pvalues, pvalues + size, 0.0);
	if (probs) std::fill(probs, probs + size, 0.0);
	for (size_t i = 0; i < P.ids.size(); ++i) {
		if (pvalues) pvalues[P.ids[i]] = P.classPval[P.cls[i]];
		if (probs && !P.prob.empty()) probs[P.ids[i]] = P.prob[i];
		if (classIds) classIds[i] = P.ids[i];
	}
	for (size_t c = 0; c < P.classScore.size(); ++c) {
		if (classScore) classScore[c] = P.classScore[c];
		if (classCount) classCount[c] = P.classCount[c];
	}
	return 0;
}

// =====================================================================================================================
// value tables
// =====================================================================================================================
static int stage_pos_prob(xxm_ctx* ctx, int motifLength, int firstMotifColumn, int enrStart, int enrEnd, int offset) {
	init_positional_probs(ctx->posProb, set_max_leng(ctx), ctx->revcomp, motifLength, firstMotifColumn, enrStart, enrEnd, offset);  // Global::posSet->max_leng: of the whole (sharded) set
	TRY(upload(ctx, ctx->d_posProb, ctx->posProb));
	ctx->posProbValid = true;
	return 0;
}

// evaluate fresh memo slots; with needMode also fill the ZOOPS/OOPS finalPval of the occupied slots that lack it, on the host
// (1 - pow(1 - pS, avgLength - length + 1) needs the C library's pow to be bit-identical).  The table is
// sparse (a few 10^4 .. 10^6 distinct scores in up to 2^24 slots), so both steps run over compacted slot lists.
static int compact_list(xxm_ctx* ctx, const uint32_t* seq, const uint32_t* pos, const uint32_t* vidx, const uint8_t* flag, uint32_t n, uint32_t* oseq,
                        uint32_t* opos, uint32_t* ovidx, uint32_t* oidx, uint32_t* nOut, const uint32_t* keys = nullptr);
// one-table models: finalPval of every class for mops = false, 1 - pow(1 - pS, avgLength - length + 1) (pValCalculation.h:96)
static int ensure_mode_table(xxm_ctx* ctx) {
	if (ctx->M.nparts != 1 || ctx->memoModeValid) return 0;
	const size_t nc = ctx->h_pS.size();
	ctx->h_finalMode.resize(nc);
	for (size_t c = 0; c < nc; ++c) ctx->h_finalMode[c] = (ctx->h_pS[c] == 1.0) ? 1.0 : zoops_final(ctx->h_pS[c], ctx->modeExponent);
	if (nc) CK(cudaMemcpyAsync(ctx->d_finalMode.p, ctx->h_finalMode.data(), nc * 8, cudaMemcpyHostToDevice, ctx->stream));
	ctx->memoModeValid = true;
	return 0;
}

static int finalize_memo(xxm_ctx* ctx, bool needMode) {
	if (ctx->M.nparts != 2) return needMode ? ensure_mode_table(ctx) : 0;
	const uint32_t cap = ctx->M.memoCap;
	CK(ctx->memoFlag.ensure(cap));
	CK(ctx->memoIdx.ensure(cap));
	uint32_t* h_ovf = reinterpret_cast<uint32_t*>(ctx->h_pin + 5);
	CK(cudaMemcpyAsync(h_ovf, ctx->M.memoOverflow, 4, cudaMemcpyDeviceToHost, ctx->stream));
	k_memo_flag<<<cdiv(cap, 256), 256, 0, ctx->stream>>>(ctx->M.memoKey, ctx->M.memoState, cap, 0u, ctx->memoFlag.p);
	CKK();
	uint32_t nf = 0;
	TRY(compact_list(ctx, nullptr, nullptr, nullptr, ctx->memoFlag.p, cap, nullptr, nullptr, nullptr, ctx->memoIdx.p, &nf, nullptr));  // synchronises
	if (*h_ovf) return ctx->fail("memo table of the >8-column path overflowed (capacity %u)", cap);
	if (nf) {
		const size_t tabBytes = static_cast<size_t>(ctx->M.nclass0 + ctx->M.nclass1) * 12;
		const int smemTables = tabBytes <= 200 * 1024;
		k_memo_eval_list<<<cdiv(nf, 128), 128, smemTables ? tabBytes : 0, ctx->stream>>>(ctx->M, ctx->memoIdx.p, nf, smemTables, ctx->d_pS.p, ctx->d_finalMops.p);
		CKK();
	}
	if (!needMode) return 0;
	k_memo_flag<<<cdiv(cap, 256), 256, 0, ctx->stream>>>(ctx->M.memoKey, ctx->M.memoState, cap, 1u, ctx->memoFlag.p);
	CKK();
	uint32_t nm = 0;
	TRY(compact_list(ctx, nullptr, nullptr, nullptr, ctx->memoFlag.p, cap, nullptr, nullptr, nullptr, ctx->memoIdx.p, &nm, nullptr));
	if (!nm) return 0;
	CK(ctx->memoTmp.ensure(nm));
	k_gather_f64<<<cdiv(nm, 256), 256, 0, ctx->stream>>>(ctx->d_pS.p, ctx->memoIdx.p, nm, ctx->memoTmp.p);
	CKK();
	ctx->h_memoVals.resize(nm);
	CK(cudaMemcpyAsync(ctx->h_memoVals.data(), ctx->memoTmp.p, static_cast<size_t>(nm) * 8, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	const double exponent = set_avg_length(ctx, ctx->modelPosSetSize) - ctx->tab.span + 1;
	for (uint32_t k = 0; k < nm; ++k) {
		const double pS = ctx->h_memoVals[k];
		ctx->h_memoVals[k] = (pS == 1.0) ? 1.0 : zoops_final(pS, exponent);
	}
	CK(cudaMemcpyAsync(ctx->memoTmp.p, ctx->h_memoVals.data(), static_cast<size_t>(nm) * 8, cudaMemcpyHostToDevice, ctx->stream));
	k_scatter_mode<<<cdiv(nm, 256), 256, 0, ctx->stream>>>(ctx->memoTmp.p, ctx->memoIdx.p, nm, ctx->d_finalMode.p, ctx->M.memoState);
	CKK();
	CK(cudaStreamSynchronize(ctx->stream));  // h_memoVals is reused by the next call
	return 0;
}

static Values make_values(xxm_ctx* ctx, bool mops) {
	Values V{};
	V.pS = ctx->d_pS.p;
	V.posProb = ctx->d_posProb.p;
	V.perCand = ctx->d_perCand.p;
	if (!ctx->opt.localization) {
		V.kind = kFinalTable;
		V.finalTab = mops ? ctx->d_finalMops.p : ctx->d_finalMode.p;
	} else if (mops) {
		V.kind = kFinalLocMops;
	} else {
		V.kind = kFinalPerCand;
	}
	return V;
}

// ZOOPS/OOPS with localization: finalPval = calculateFinalPval(len, delta_0, pS * positionalProb[pos]) needs pow/exp
// (pValCalculation.h:82, :116-120) -> per-candidate host evaluation (exact, rare: --localization without --mops)
static int compute_per_candidate(xxm_ctx* ctx, const uint32_t* d_seq, const uint32_t* d_pos, const uint32_t* d_vidx, uint32_t n, int shift, int enrStart,
                                 int enrEnd) {
	CK(ctx->d_perCand.ensure(std::max<uint32_t>(1, n)));
	if (!n) return 0;
	std::vector<uint32_t> seq(n), pos(n), vidx(n);
	std::vector<double> pS(ctx->nval);
	CK(cudaMemcpyAsync(seq.data(), d_seq, n * 4ull, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaMemcpyAsync(pos.data(), d_pos, n * 4ull, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaMemcpyAsync(vidx.data(), d_vidx, n * 4ull, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaMemcpyAsync(pS.data(), ctx->d_pS.p, ctx->nval * 8ull, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	std::vector<double> fin(n, 1.0);
	const int region = enrEnd - enrStart + 1;
	const int length = ctx->tab.span;
	for (uint32_t i = 0; i < n; ++i) {
		if (vidx[i] == kNone) continue;
		const double p = pS[vidx[i]];
		if (p == 1.0) continue;
		const int len = ctx->len[seq[i] - 1] - length + 1;
		const int delta_0 = std::min(len, region);
		fin[i] = final_pval_localized(ctx->sum_i, len, delta_0, p * ctx->posProb[static_cast<int>(pos[i]) + shift]);
	}
	CK(cudaMemcpyAsync(ctx->d_perCand.p, fin.data(), n * 8ull, cudaMemcpyHostToDevice, ctx->stream));
	return 0;
}

static void set_max_sites_mops(xxm_ctx* ctx, int posSetSize) {
	const int length = ctx->tab.span;
	ctx->maxSites = set_total(ctx, posSetSize) - (ctx->revcomp ? 2 : 1) * set_nent(ctx, posSetSize) * (length - 1);  // StartPosUpdater-inl.h:135-138
}
static int32_t clamp_i32(int64_t v) { return static_cast<int32_t>(std::min<int64_t>(v, 2147483647LL)); }

// =====================================================================================================================
// scan machinery
// =====================================================================================================================
#include "xxm_gpu_scan4.inc"
#include "xxm_gpu_scan5.inc"

// the scan finds the windows whose left k-mer is in the table; unless ctx->scanFindsOnlySites the caller still has to
// drop the entries whose finalPval turns out to be >= 1 (or whose right part / product test fails)
static int run_scan(xxm_ctx* ctx, int nseq, const Values& V, int checkFinal, DevBuf<uint32_t>& oseq, DevBuf<uint32_t>& opos, DevBuf<uint32_t>& ovidx,
                    DevBuf<float>* oscore, uint32_t* nOut, bool timeIt, size_t* deferCap = nullptr) {
	(void) checkFinal;
	return run_scan5(ctx, nseq, V, oseq, opos, ovidx, oscore, nOut, timeIt, deferCap);
}

// selected = flag byte set, or (keys != nullptr) head of a run of equal keys
static int compact_list(xxm_ctx* ctx, const uint32_t* seq, const uint32_t* pos, const uint32_t* vidx, const uint8_t* flag, uint32_t n, uint32_t* oseq,
                        uint32_t* opos, uint32_t* ovidx, uint32_t* oidx, uint32_t* nOut, const uint32_t* keys) {
	*nOut = 0;
	if (!n) return 0;
	const uint32_t ntiles = cdiv(n, kCmpTile);
	CK(ctx->status.ensure(static_cast<size_t>(ntiles) + 2));
	CK(ctx->cmpCounts.ensure(ntiles));
	CK(ctx->total64.ensure(2));
	CompactIO io{};
	io.seq = seq; io.pos = pos; io.vidx = vidx; io.flag = flag; io.keys = keys; io.n = n;
	io.oseq = oseq; io.opos = opos; io.ovidx = ovidx; io.oidx = oidx;
	io.tileCount = ctx->cmpCounts.p; io.tileOff = ctx->status.p;
	k_cmp_count<<<ntiles, 256, 0, ctx->stream>>>(io);
	CKK();
	k_tile_scan<<<cdiv(ntiles, 1024), 1024, 0, ctx->stream>>>(ctx->cmpCounts.p, ntiles, ctx->status.p, ctx->total64.p);
	CKK();
	k_cmp_scatter<<<ntiles, 256, 0, ctx->stream>>>(io);
	CKK();
	CK(cudaMemcpyAsync(ctx->h_pin + 3, ctx->total64.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	*nOut = static_cast<uint32_t>(ctx->h_pin[3]);
	return 0;
}

// updateMultOcc over an entry list -> keep flags + float scores
// forward-only overlap removal (k_select_mops_fwd) with the class histogram fused in; nDev != nullptr: n is the capacity
// and the entry count is read on the device
static int launch_select_fwd(xxm_ctx* ctx, const Values& V, const uint32_t* seq, const uint32_t* pos, const uint32_t* vidx, uint32_t n,
                             const unsigned long long* nDev, int shift, DevBuf<uint8_t>& keep, DevBuf<float>& kscore, DevBuf<uint16_t>* kcls = nullptr) {
	SelectFwdIO f{};
	ctx->kclsSerial = 0;
	if (kcls && V.kind == kFinalTable && ctx->nval < 0xFFFFu) {
		CK(kcls->ensure(std::max<uint32_t>(8, n) + 8));
		f.kcls = kcls->p;
	}
	f.seq = seq; f.pos = pos; f.vidx = vidx; f.n = n; f.nDev = nDev; f.length = ctx->tab.span; f.shift = shift; f.kscore = kscore.p; f.keep = keep.p;
	f.counters = ctx->counters.p;
	const bool wantHist = (V.kind == kFinalTable);
	if (wantHist) {
		CK(ctx->hist.ensure(std::max<uint32_t>(1, ctx->nval)));
		CK(cudaMemsetAsync(ctx->hist.p, 0, ctx->nval * 4ull, ctx->stream));
		f.hist = ctx->hist.p; f.nval = ctx->nval; f.smemHist = ctx->nval <= 8192;
	}
	const size_t sm = (wantHist && f.smemHist) ? ctx->nval * 4ull : 0;
	uint32_t grid = static_cast<uint32_t>(ctx->numSMs) * 16;
	if (!nDev) grid = std::min<uint32_t>(cdiv(n, 256 * kSelChunk), grid);  // a thread owns kSelChunk entries per pass
	k_select_mops_fwd<<<grid, 256, sm, ctx->stream>>>(V, f);
	CKK();
	ctx->histValid = wantHist;
	ctx->histSerial = ++ctx->histCounter;
	if (f.kcls) ctx->kclsSerial = ctx->histSerial;
	return 0;
}

// the same on class indices (k_select_cls_fwd, xxm_select3.cuh): mops table values of a one-table model with a verified
// class order; no float scores are written
static int launch_select_cls(xxm_ctx* ctx, const uint32_t* seq, const uint32_t* pos, const uint32_t* vidx, uint32_t n, const unsigned long long* nDev,
                             DevBuf<uint8_t>& keep, DevBuf<uint16_t>& kcls) {
	SelectClsIO f{};
	CK(kcls.ensure(std::max<uint32_t>(8, n) + 8));
	CK(ctx->hist.ensure(std::max<uint32_t>(1, ctx->nval)));
	CK(cudaMemsetAsync(ctx->hist.p, 0, ctx->nval * 4ull, ctx->stream));
	f.seq = seq; f.pos = pos; f.vidx = vidx; f.n = n; f.nDev = nDev; f.length = ctx->tab.span;
	f.clsThr = ctx->d_clsThr.p; f.nSite = ctx->clsNSite; f.nval = ctx->nval;
	f.keep = keep.p; f.kcls = kcls.p; f.counters = ctx->counters.p; f.hist = ctx->hist.p;
	uint32_t grid = static_cast<uint32_t>(ctx->numSMs) * 16;
	if (!nDev) grid = std::min<uint32_t>(cdiv(n, 256 * kSelChunk), grid);
	k_select_cls_fwd<<<grid, 32 * kSel3Warps, ctx->nval * 6ull, ctx->stream>>>(f);
	CKK();
	ctx->histValid = true;
	ctx->histSerial = ++ctx->histCounter;
	ctx->kclsSerial = ctx->histSerial;
	return 0;
}

// --revcomp: forward rule + reverse-strand look-back on class indices (k_select_cls_rc, xxm_select4.cuh); also reports the
// largest number of kept sites of one sequence in counters[1]
static int launch_select_cls_rc(xxm_ctx* ctx, const uint32_t* seq, const uint32_t* pos, const uint32_t* vidx, uint32_t n, const unsigned long long* nDev,
                                DevBuf<uint8_t>& keep, DevBuf<uint16_t>& kcls) {
	SelectRcIO f{};
	CK(kcls.ensure(std::max<uint32_t>(8, n) + 8));
	CK(ctx->stk.ensure(std::max<uint32_t>(8, n) + 8));
	CK(ctx->hist.ensure(std::max<uint32_t>(1, ctx->nval)));
	CK(cudaMemsetAsync(ctx->hist.p, 0, ctx->nval * 4ull, ctx->stream));
	f.seq = seq; f.pos = pos; f.vidx = vidx; f.n = n; f.nDev = nDev; f.length = ctx->tab.span;
	f.clsThr = ctx->d_clsThr.p; f.nSite = ctx->clsNSite; f.nval = ctx->nval;
	f.stk = ctx->stk.p; f.keep = keep.p; f.kcls = kcls.p; f.counters = ctx->counters.p; f.hist = ctx->hist.p;
	uint32_t grid = static_cast<uint32_t>(ctx->numSMs) * 6;
	if (!nDev) grid = std::max<uint32_t>(1, std::min<uint32_t>(cdiv(n, kSel4Tile), grid));
	k_select_cls_rc<<<grid, kSel4Threads, ctx->nval * 6ull, ctx->stream>>>(ctx->S, f);
	CKK();
	ctx->histValid = true;
	ctx->histSerial = ++ctx->histCounter;
	ctx->kclsSerial = ctx->histSerial;
	return 0;
}

static bool select_cap_safe(const xxm_ctx* ctx) { return set_max_leng(ctx) / (ctx->tab.span + 1) + 1 <= ctx->opt.maxMotifsPerSequence; }

static int select_mops(xxm_ctx* ctx, const Values& V, const uint32_t* seq, const uint32_t* pos, const uint32_t* vidx, uint32_t n, int shift, bool ordered,
                       DevBuf<uint8_t>& keep, DevBuf<float>& kscore, uint32_t* nKept, uint32_t* maxPerSeq) {
	CK(keep.ensure(std::max<uint32_t>(1, n)));
	CK(kscore.ensure(std::max<uint32_t>(1, n)));
	CK(ctx->stk.ensure(std::max<uint32_t>(1, n)));
	CK(ctx->counters.ensure(8));
	CK(cudaMemsetAsync(ctx->counters.p, 0, 16, ctx->stream));
	*nKept = 0; *maxPerSeq = 0;
	if (!n) return 0;
	SelectIO io{};
	io.seq = seq; io.pos = pos; io.vidx = vidx; io.n = n; io.shift = shift; io.length = ctx->tab.span; io.revcomp = ctx->revcomp;
	io.stk = ctx->stk.p; io.kscore = kscore.p; io.keep = keep.p; io.counters = ctx->counters.p;
	ctx->histValid = false;
	// forward-only sets whose per-sequence site count cannot reach the cap: cluster walk with the class histogram fused in
	const bool capSafe = select_cap_safe(ctx);
	if (ordered && !ctx->revcomp && capSafe) {
		TRY(launch_select_fwd(ctx, V, seq, pos, vidx, n, nullptr, shift, keep, kscore));
	} else {
		CK(cudaMemsetAsync(keep.p, 0, n, ctx->stream));
		if (ordered) k_select_mops<<<cdiv(n, 128), 128, 0, ctx->stream>>>(ctx->S, V, io);
		else k_select_mops_serial<<<1, 32, 0, ctx->stream>>>(ctx->S, V, io);
		CKK();
	}
	uint32_t c[2] = {0, 0};
	CK(cudaMemcpyAsync(c, ctx->counters.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	*nKept = c[0]; *maxPerSeq = c[1];
	if (ordered && !ctx->revcomp && capSafe) *maxPerSeq = std::min<uint32_t>(c[0], static_cast<uint32_t>(ctx->max_leng / (ctx->tab.span + 1) + 1));
	return 0;
}

static int select_slots(xxm_ctx* ctx, const Values& V, const uint32_t* seq, const uint32_t* pos, const uint32_t* vidx, uint32_t n, int shift, bool ordered,
                        uint32_t nslots, DevBuf<uint32_t>& sIdx, DevBuf<float>& sScore, uint32_t* nFilled) {
	CK(sIdx.ensure(std::max<uint32_t>(1, nslots)));
	CK(sScore.ensure(std::max<uint32_t>(1, nslots)));
	CK(ctx->counters.ensure(8));
	CK(cudaMemsetAsync(ctx->counters.p, 0, 16, ctx->stream));
	k_fill_u32<<<cdiv(nslots, 256), 256, 0, ctx->stream>>>(sIdx.p, kNone, nslots);
	CKK();
	k_fill_f32<<<cdiv(nslots, 256), 256, 0, ctx->stream>>>(sScore.p, 1.0f, nslots);
	CKK();
	*nFilled = 0;
	if (n) {
		SlotIO io{};
		io.seq = seq; io.pos = pos; io.vidx = vidx; io.n = n; io.shift = shift;
		io.slotIdx = sIdx.p; io.slotScore = sScore.p; io.counters = ctx->counters.p;
		if (ordered) k_select_zoops<<<cdiv(n, static_cast<uint32_t>(kZoopsTile)), 256, 0, ctx->stream>>>(V, io);
		else k_select_zoops_serial<<<1, 32, 0, ctx->stream>>>(V, io);
		CKK();
		CK(cudaMemcpyAsync(nFilled, ctx->counters.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
	}
	CK(cudaStreamSynchronize(ctx->stream));
	return 0;
}

// =====================================================================================================================
// a2: full scan
// =====================================================================================================================
static void scan_full_set_best(xxm_ctx* ctx, uint32_t n, bool lazyScores) {
	SiteList& B = ctx->L[XXM_LIST_BEST_STARTPOS];
	B = SiteList();
	B.kind = kListCompact; B.seq = ctx->b_seq.p; B.pos = ctx->b_pos.p; B.vidx = ctx->b_vidx.p; B.n = n; B.score = lazyScores ? nullptr : ctx->b_score.p;
	B.lazyPS = lazyScores ? ctx->b_pSsnap.p : nullptr; B.members = n;
	B.modelSerial = ctx->modelSerial;
}

static void scan_full_set_sorted(xxm_ctx* ctx, const Values& V, uint32_t n, uint32_t kept, uint32_t mps, int posSetSize, bool lazyScores = false) {
	SiteList& Sd = ctx->L[XXM_LIST_SORTED_SITES];
	Sd = SiteList();
	Sd.kind = kListFlagged; Sd.seq = ctx->b_seq.p; Sd.pos = ctx->b_pos.p; Sd.vidx = ctx->b_vidx.p; Sd.n = n; Sd.score = lazyScores ? nullptr : ctx->s_kscore.p;
	// (one table, no localization: the mops finalPval of a class is its pS, so the snapshot of the pS table serves both lists)
	Sd.lazyPS = lazyScores ? ctx->b_pSsnap.p : nullptr; Sd.lazyDst = 1;
	Sd.keep = ctx->s_keep.p; Sd.members = kept; Sd.maxPerSeq = mps; Sd.tableValued = (V.kind == kFinalTable); Sd.modelSerial = ctx->modelSerial;
	Sd.histSerial = ctx->histValid ? ctx->histSerial : 0;
	ctx->nbSites = static_cast<int>(kept);
	set_max_sites_mops(ctx, posSetSize);
}

static bool scan_full_can_defer(const xxm_ctx* ctx);
static int scan_full_launch(xxm_ctx* ctx, int32_t posSetSize);
static int scan_full_finish(xxm_ctx* ctx, int32_t* nbSites, int32_t* nbBestSites, int32_t* maxSites, bool* overflow);
extern "C" int xxm_scan_full(xxm_ctx* ctx, int32_t posSetSize, int32_t enrStart, int32_t enrEnd, int32_t* nbSites, int32_t* nbBestSites, int32_t* maxSites) {
	XXM_NVTX("xxm_scan_full");
	CK(cudaSetDevice(ctx->device));
	if (!ctx->haveModel) return ctx->fail("xxm_scan_full: no model (call xxm_reinitialize)");
	if (posSetSize < 1 || posSetSize > ctx->nent) return ctx->fail("posSetSize out of range");
	const PvalueTables& T = ctx->tab;
	if (scan_full_can_defer(ctx) && !std::getenv("XXM_NO_DEFER")) {
		// the same launches xxm_optimize_step queues (find, expand, class-index overlap removal + histogram), read back at once
		TRY(scan_full_launch(ctx, posSetSize));
		CK(cudaStreamSynchronize(ctx->stream));
		bool overflow = false;
		TRY(scan_full_finish(ctx, nbSites, nbBestSites, maxSites, &overflow));
		if (!overflow) return 0;
		// the site arrays were too small: the ordinary path below grows them
	}
	if (ctx->opt.localization) TRY(stage_pos_prob(ctx, T.span, ctx->modelFirstCol, enrStart, enrEnd, 0));
	const Values V = make_values(ctx, true);  // the full scan always uses mops = true (StartPosUpdater-inl.h:99)
	uint32_t n = 0;
	const bool lazyScores = ctx->scanFindsOnlySites;  // bestStartPos[].score = float(pS[value index]): materialised only if somebody reads it
	if (ctx->scanFindsOnlySites) {
		TRY(run_scan(ctx, posSetSize, V, 1, ctx->b_seq, ctx->b_pos, ctx->b_vidx, nullptr, &n, true));
		CK(ctx->b_pSsnap.ensure(std::max<uint32_t>(1, ctx->nval)));
		CK(cudaMemcpyAsync(ctx->b_pSsnap.p, ctx->d_pS.p, ctx->nval * 8ull, cudaMemcpyDeviceToDevice, ctx->stream));
	} else {
		// candidates first; the p-value of a >8-column window is known only after the memo slots are evaluated
		uint32_t nc = 0;
		TRY(run_scan(ctx, posSetSize, V, 0, ctx->t_seq, ctx->t_pos, ctx->t_vidx, nullptr, &nc, true));
		TRY(finalize_memo(ctx, false));
		CK(ctx->flag8.ensure(std::max<uint32_t>(1, nc)));
		CK(ctx->counters.ensure(8));
		CK(cudaMemsetAsync(ctx->counters.p, 0, 16, ctx->stream));
		if (nc && V.kind == kFinalTable) {
			CK(ctx->d_siteBits.ensure(ctx->nval / 32 + 1));
			k_site_bits<<<cdiv(ctx->nval, 256), 256, 0, ctx->stream>>>(V.finalTab, ctx->nval, ctx->d_siteBits.p);
			CKK();
			k_flag_hits_bits<<<cdiv(nc, 256), 256, 0, ctx->stream>>>(ctx->d_siteBits.p, ctx->t_vidx.p, nc, ctx->flag8.p);
			CKK();
		} else if (nc) {
			k_flag_hits<<<cdiv(nc, 256), 256, 0, ctx->stream>>>(V, ctx->t_vidx.p, ctx->t_pos.p, nc, ctx->flag8.p, ctx->counters.p);
			CKK();
		}
		CK(ctx->b_seq.ensure(std::max<uint32_t>(1, nc)));
		CK(ctx->b_pos.ensure(std::max<uint32_t>(1, nc)));
		CK(ctx->b_vidx.ensure(std::max<uint32_t>(1, nc)));
		TRY(compact_list(ctx, ctx->t_seq.p, ctx->t_pos.p, ctx->t_vidx.p, ctx->flag8.p, nc, ctx->b_seq.p, ctx->b_pos.p, ctx->b_vidx.p, nullptr, &n));
	}
	ctx->nBest = n;
	// bestStartPos[].score = float(pValScore) (StartPosUpdater-inl.h:105)
	if (!lazyScores) CK(ctx->b_score.ensure(std::max<uint32_t>(1, n)));
	if (n && !ctx->scanFindsOnlySites) {
		k_entry_scores<<<cdiv(n, 256), 256, 0, ctx->stream>>>(V, ctx->b_vidx.p, ctx->b_pos.p, n, 0, 0, ctx->b_score.p);
		CKK();
	}
	scan_full_set_best(ctx, n, lazyScores);
	// updateMultOcc -> sortedSites
	uint32_t kept = 0, mps = 0;
	TRY(select_mops(ctx, V, ctx->b_seq.p, ctx->b_pos.p, ctx->b_vidx.p, n, 0, true, ctx->s_keep, ctx->s_kscore, &kept, &mps));
	scan_full_set_sorted(ctx, V, n, kept, mps, posSetSize);
	*nbSites = ctx->nbSites;
	*nbBestSites = static_cast<int32_t>(n);
	*maxSites = clamp_i32(ctx->maxSites);
	return 0;
}

// The same scan with nothing read back: find, expand, overlap removal (+ histogram) are queued, the hit count and the
// select counters travel to ctx->h_pin with the caller's next synchronisation, and scan_full_finish does the bookkeeping.
// Only for the case xxm_optimize_step's fast path handles (one table, no localization, forward strand only).
static bool scan_full_can_defer(const xxm_ctx* ctx) {
	// (--revcomp sets need the class-index kernel: there is no float-score variant of the reverse-strand look-back that runs
	// off a device-side entry count)
	return ctx->scanFindsOnlySites && !ctx->opt.localization && select_cap_safe(ctx) && ctx->M.nparts == 1 && (!ctx->revcomp || ctx->clsOrderOK);
}

static int scan_full_launch(xxm_ctx* ctx, int32_t posSetSize) {
	const Values V = make_values(ctx, true);
	if (V.kind != kFinalTable) return ctx->fail("internal: deferred scan needs table values");
	uint32_t n = 0;
	size_t cap = 0;
	TRY(run_scan(ctx, posSetSize, V, 1, ctx->b_seq, ctx->b_pos, ctx->b_vidx, nullptr, &n, true, &cap));
	CK(ctx->b_pSsnap.ensure(std::max<uint32_t>(1, ctx->nval)));
	CK(cudaMemcpyAsync(ctx->b_pSsnap.p, ctx->d_pS.p, ctx->nval * 8ull, cudaMemcpyDeviceToDevice, ctx->stream));
	const uint32_t capN = static_cast<uint32_t>(std::min<size_t>(cap, 0xFFFFFFF0u));
	CK(ctx->s_keep.ensure(capN));
	CK(ctx->s_kscore.ensure(capN));
	CK(ctx->counters.ensure(8));
	CK(cudaMemsetAsync(ctx->counters.p, 0, 16, ctx->stream));
	ctx->sortedLazy = ctx->clsOrderOK && ctx->M.nparts == 1 && V.finalTab == ctx->d_finalMops.p && capN <= 0x7FFFFF00u && (ctx->revcomp || !std::getenv("XXM_NO_CLS_SELECT"));
	if (ctx->revcomp && !ctx->sortedLazy) return ctx->fail("internal: deferred --revcomp scan without a class order");
	if (ctx->sortedLazy && ctx->revcomp)
		TRY(launch_select_cls_rc(ctx, ctx->b_seq.p, ctx->b_pos.p, ctx->b_vidx.p, capN, reinterpret_cast<const unsigned long long*>(ctx->total64.p), ctx->s_keep, ctx->s_kcls));
	else if (ctx->sortedLazy)
		TRY(launch_select_cls(ctx, ctx->b_seq.p, ctx->b_pos.p, ctx->b_vidx.p, capN, reinterpret_cast<const unsigned long long*>(ctx->total64.p), ctx->s_keep, ctx->s_kcls));
	else
		TRY(launch_select_fwd(ctx, V, ctx->b_seq.p, ctx->b_pos.p, ctx->b_vidx.p, capN, reinterpret_cast<const unsigned long long*>(ctx->total64.p), 0, ctx->s_keep,
		                      ctx->s_kscore, &ctx->s_kcls));
	CK(cudaMemcpyAsync(ctx->h_pin + 1, ctx->counters.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
	ctx->pend.active = true; ctx->pend.cap = capN; ctx->pend.posSetSize = posSetSize; ctx->pend.scanWords = ctx->wordOff[posSetSize];
	return 0;
}

// after the stream has been synchronised; *overflow: the site list did not fit its arrays, the scan has to be redone
static int scan_full_finish(xxm_ctx* ctx, int32_t* nbSites, int32_t* nbBestSites, int32_t* maxSites, bool* overflow) {
	ctx->pend.active = false;
	const unsigned long long tot = ctx->h_pin[0];
	*overflow = tot > ctx->pend.cap;
	if (*overflow) { ctx->histValid = false; return 0; }
	const uint32_t n = static_cast<uint32_t>(tot);
	uint32_t c[2];
	std::memcpy(c, ctx->h_pin + 1, 8);
	scan5_stats(ctx, ctx->pend.scanWords, tot);
	ctx->nBest = n;
	scan_full_set_best(ctx, n, true);
	const Values V = make_values(ctx, true);
	uint32_t mps = std::min<uint32_t>(c[0], static_cast<uint32_t>(ctx->max_leng / (ctx->tab.span + 1) + 1));
	if (ctx->revcomp) mps = c[1];  // k_select_cls_rc reports it
	scan_full_set_sorted(ctx, V, n, c[0], mps, ctx->pend.posSetSize, ctx->sortedLazy);
	*nbSites = ctx->nbSites;
	*nbBestSites = static_cast<int32_t>(n);
	*maxSites = clamp_i32(ctx->maxSites);
	return 0;
}

#include "xxm_gpu_part2.inc"
#include "xxm_gpu_part3.inc"
#include "xxm_gpu_part4.inc"
#include "xxm_gpu_part5.inc"
#include "xxm_gpu_part6.inc"
#include "xxm_gpu_part8.inc"
#include "xxm_gpu_part9.inc"
#include "xxm_gpu_part10.inc"

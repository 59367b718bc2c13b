// K3, third generation: updateMultOcc (StartPosUtils.h:12-28) on CLASS INDICES, for the fast path of a refinement step
// (one p-value table, forward strand only, no localization: the float score of a site is a function of its value index).
//
// The reference compares the float-stored score of the last kept site with the double p-value of the incoming window
// (`(double)lastScore <= finj` drops the incoming one, StartPosUtils.h:18).  With fin[] the finalPval table of the scan and
// fin[] non-decreasing in the class index (the host checks it for every table, xxm_gpu.cu: stage_class_order; classes are
// built in descending score order with cumulative p-values) that test is  c_incoming >= T[c_kept]  with
//     T[k] = min { c : fin[c] >= (double)(float)fin[k] },
// and "finj < 1.0" is  c < nSite.  So the whole greedy runs on 16-bit integers out of shared memory: no FP64, no dependent
// gathers from the p-value table in L2, and no per-entry float score is written (it is float(pS[vidx]), materialised when
// somebody asks for it -- ensure_scores).
//
// Work split as in the second generation (a thread owns 8 consecutive entries; it replays the part of its first entry's
// cluster that lies before the chunk and looks ahead while a later entry could still replace its last kept site), but the
// 256 entries of a warp are staged in shared memory first, so the replay / look-ahead loops leave the SM only at the
// edges of the warp's window.
#pragma once
#include "xxm_select2.cuh"

namespace xxm {

struct SelectClsIO {
	const uint32_t* seq;
	const uint32_t* pos;
	const uint32_t* vidx;
	uint32_t n;                      // entries, or -- with nDev -- the capacity of the arrays
	const unsigned long long* nDev;  // non-null: entry count on the device (written by the expand kernel)
	int length;
	const uint16_t* clsThr;          // [nval] T[k]
	uint32_t nSite;                  // classes [0, nSite) are sites
	uint32_t nval;
	uint8_t* keep;                   // [n]
	uint16_t* kcls;                  // [n] value index of a kept entry, 0xFFFF for a dropped one
	uint32_t* counters;              // [0] += kept sites
	uint32_t* hist;                  // [nval], zeroed
};

constexpr int kSel3Warps = 8;
constexpr int kSel3Window = 32 * kSelChunk;  // entries per warp and pass

// (5 CTAs per SM -- 48 registers, a few spills -- measured slower: 0.711 vs 0.689 ms per step)
__global__ void __launch_bounds__(32 * kSel3Warps, 4) k_select_cls_fwd(SelectClsIO io) {
	extern __shared__ __align__(16) uint8_t smem3[];
	uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem3);
	uint16_t* s_thr = reinterpret_cast<uint16_t*>(s_hist + io.nval);
	__shared__ __align__(16) uint32_t s_posAll[kSel3Warps][kSel3Window];
	__shared__ __align__(16) uint32_t s_seqAll[kSel3Warps][kSel3Window];
	__shared__ __align__(16) uint16_t s_clsAll[kSel3Warps][kSel3Window];
	__shared__ uint32_t s_cnt;
	for (uint32_t k = threadIdx.x; k < io.nval; k += blockDim.x) { s_hist[k] = 0; s_thr[k] = io.clsThr[k]; }
	if (threadIdx.x == 0) s_cnt = 0;
	__syncthreads();
	const uint32_t L = static_cast<uint32_t>(io.length);
	const uint32_t nSite = io.nSite;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	uint32_t* const s_pos = s_posAll[wid];
	uint32_t* const s_seq = s_seqAll[wid];
	uint16_t* const s_cls = s_clsAll[wid];
	// all entry indices fit 31 bits (the launcher falls back to k_select_mops_fwd for longer lists): 32-bit index arithmetic
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
	const int n = static_cast<int>(io.nDev ? min(*io.nDev, static_cast<unsigned long long>(io.n)) : io.n);
	uint32_t kept = 0;
	for (uint32_t uw = warp * kSel3Window; uw < static_cast<uint32_t>(n); uw += nwarps * kSel3Window) {
		const int wj0 = static_cast<int>(uw);                                    // first entry of the warp's window
		const uint32_t wlen = static_cast<uint32_t>(min(kSel3Window, n - wj0));  // entries of the warp's window
		const int r0 = kSelChunk * lane;                                         // first entry of the chunk, relative to the window
		const int j0 = wj0 + r0;
		const int m = max(0, min(kSelChunk, n - j0));
		const uint32_t* const gseq = io.seq + wj0;
		const uint32_t* const gpos = io.pos + wj0;
		const uint32_t* const gvid = io.vidx + wj0;
		uint32_t sq[kSelChunk], ps[kSelChunk], vi[kSelChunk];
		if (m == kSelChunk) {
			const uint4 a0 = *reinterpret_cast<const uint4*>(gseq + r0), a1 = *reinterpret_cast<const uint4*>(gseq + r0 + 4);
			const uint4 b0 = *reinterpret_cast<const uint4*>(gpos + r0), b1 = *reinterpret_cast<const uint4*>(gpos + r0 + 4);
			const uint4 v0 = *reinterpret_cast<const uint4*>(gvid + r0), v1 = *reinterpret_cast<const uint4*>(gvid + r0 + 4);
			sq[0] = a0.x; sq[1] = a0.y; sq[2] = a0.z; sq[3] = a0.w; sq[4] = a1.x; sq[5] = a1.y; sq[6] = a1.z; sq[7] = a1.w;
			ps[0] = b0.x; ps[1] = b0.y; ps[2] = b0.z; ps[3] = b0.w; ps[4] = b1.x; ps[5] = b1.y; ps[6] = b1.z; ps[7] = b1.w;
			vi[0] = v0.x; vi[1] = v0.y; vi[2] = v0.z; vi[3] = v0.w; vi[4] = v1.x; vi[5] = v1.y; vi[6] = v1.z; vi[7] = v1.w;
		} else {
#pragma unroll
			for (int k = 0; k < kSelChunk; ++k) {
				const bool in = k < m;
				sq[k] = in ? gseq[r0 + k] : 0u;
				ps[k] = in ? gpos[r0 + k] : 0u;
				vi[k] = in ? gvid[r0 + k] : kNone;
			}
		}
#pragma unroll
		for (int k = 0; k < kSelChunk; ++k) vi[k] = min(vi[k], 0xFFFFu);  // (kNone and anything >= nSite: not a site)
		__syncwarp();  // the previous pass has finished reading the window
		*reinterpret_cast<uint4*>(s_seq + r0) = make_uint4(sq[0], sq[1], sq[2], sq[3]);
		*reinterpret_cast<uint4*>(s_seq + r0 + 4) = make_uint4(sq[4], sq[5], sq[6], sq[7]);
		*reinterpret_cast<uint4*>(s_pos + r0) = make_uint4(ps[0], ps[1], ps[2], ps[3]);
		*reinterpret_cast<uint4*>(s_pos + r0 + 4) = make_uint4(ps[4], ps[5], ps[6], ps[7]);
		*reinterpret_cast<uint4*>(s_cls + r0) =
			make_uint4(vi[0] | (vi[1] << 16), vi[2] | (vi[3] << 16), vi[4] | (vi[5] << 16), vi[6] | (vi[7] << 16));
		__syncwarp();
		// entry r of the window (r < 0 or r >= wlen: outside of it, from global memory)
		auto pos_at = [&](int r) -> uint32_t { return (static_cast<uint32_t>(r) < wlen) ? s_pos[r] : gpos[r]; };
		auto seq_at = [&](int r) -> uint32_t { return (static_cast<uint32_t>(r) < wlen) ? s_seq[r] : gseq[r]; };
		auto cls_at = [&](int r) -> uint32_t { return (static_cast<uint32_t>(r) < wlen) ? static_cast<uint32_t>(s_cls[r]) : min(gvid[r], 0xFFFFu); };
		// state of the greedy: the kept site a later entry can still replace.  lastEnd = its startPos + length (0: none)
		uint32_t lastEnd = 0, lastT = 0;
		int lastK = -1;  // its index inside the chunk, -1: it lies before the chunk
		uint32_t keepBits = 0, lastSeq = 0;
		if (m > 0) {
			// ---- the part of the cluster before the chunk: back to its head, then forward again
			if (j0 > 0) {
				uint32_t pp = pos_at(r0 - 1);
				if (seq_at(r0 - 1) == sq[0] && pp + L >= ps[0]) {
					int h = r0 - 1;
					while (h > -wj0) {
						const uint32_t p2 = pos_at(h - 1);
						if (seq_at(h - 1) != sq[0] || p2 + L < pp) break;
						pp = p2;
						--h;
					}
					for (int r = h; r < r0; ++r) {
						const uint32_t pj = pos_at(r), cj = cls_at(r);
						if (cj < nSite && !(lastEnd >= pj && cj >= lastT)) { lastEnd = pj + L; lastT = s_thr[cj]; }
					}
				}
			}
			// ---- own entries
#pragma unroll
			for (int k = 0; k < kSelChunk; ++k) {
				if (k < m) {
					if (k > 0 && sq[k] != sq[k - 1]) lastEnd = 0;  // a new sequence starts inside the chunk
					const bool ov = lastEnd >= ps[k];
					if (vi[k] < nSite && !(ov && vi[k] >= lastT)) {
						if (ov && lastK >= 0) keepBits &= ~(1u << lastK);
						keepBits |= 1u << k;
						lastK = k;
						lastSeq = sq[k];
						lastEnd = ps[k] + L;
						lastT = s_thr[vi[k]];
					}
				}
			}
			// ---- can a later entry still replace my last kept site?
			if (lastK >= 0) {
				const uint32_t mySeq = lastSeq, better = min(lastT, nSite);
				for (int r = r0 + m; r < n - wj0; ++r) {
					const uint32_t pj = pos_at(r);
					if (seq_at(r) != mySeq || pj > lastEnd) break;
					if (cls_at(r) >= better) continue;  // not a site, or not better than mine
					keepBits &= ~(1u << lastK);
					break;
				}
			}
			// ---- flags and the compact class array
			if (m == kSelChunk) {
				uint2 kb;
				kb.x = (keepBits & 1u) | ((keepBits & 2u) << 7) | ((keepBits & 4u) << 14) | ((keepBits & 8u) << 21);
				kb.y = ((keepBits >> 4) & 1u) | ((keepBits & 32u) << 3) | ((keepBits & 64u) << 10) | ((keepBits & 128u) << 17);
				*reinterpret_cast<uint2*>(io.keep + j0) = kb;
				uint32_t cw[4];
#pragma unroll
				for (int k = 0; k < 4; ++k) {
					const uint32_t lo = ((keepBits >> (2 * k)) & 1u) ? vi[2 * k] : 0xFFFFu;
					const uint32_t hi = ((keepBits >> (2 * k + 1)) & 1u) ? vi[2 * k + 1] : 0xFFFFu;
					cw[k] = lo | (hi << 16);
				}
				*reinterpret_cast<uint4*>(io.kcls + j0) = make_uint4(cw[0], cw[1], cw[2], cw[3]);
			} else {
#pragma unroll
				for (int k = 0; k < kSelChunk; ++k)
					if (k < m) {
						io.keep[j0 + k] = (keepBits >> k) & 1u;
						io.kcls[j0 + k] = ((keepBits >> k) & 1u) ? static_cast<uint16_t>(vi[k]) : static_cast<uint16_t>(0xFFFFu);
					}
			}
			kept += __popc(keepBits);
#pragma unroll
			for (int k = 0; k < kSelChunk; ++k)
				if ((keepBits >> k) & 1u) atomicAdd(&s_hist[vi[k]], 1u);
		}
	}
	kept = __reduce_add_sync(0xFFFFFFFFu, kept);
	if (lane == 0 && kept) atomicAdd(&s_cnt, kept);
	__syncthreads();
	if (threadIdx.x == 0 && s_cnt) atomicAdd(&io.counters[0], s_cnt);
	for (uint32_t k = threadIdx.x; k < io.nval; k += blockDim.x)
		if (s_hist[k]) atomicAdd(&io.hist[k], s_hist[k]);
}

}  // namespace xxm

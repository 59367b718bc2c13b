// K3, fourth generation: updateMultOcc (StartPosUtils.h:12-65) on CLASS INDICES for --revcomp sets -- the forward rule AND the
// look-back over the reverse strand -- so that cfg2 / cfg3 style scans take the same no-read-back path as forward-only sets.
//
// With --revcomp a sequence is stored as fwd + N + revcomp (seqset.cpp:128-178; stored length n = 2 n0 + 1) and the reference,
// after the forward overlap test against the last kept site (StartPosUtils.h:14-27), walks back over the kept sites of the
// same sequence to the one that covers the mirrored window revPos = n - length + 2 - pos (:29-58): kept sites with
// startPos > revPos + length - 1 are skipped, the first one at or below that is tested against revPos - length + 1, and the
// better of the two survives (the loser is removed from the middle of the kept array).
//
// Restated without the walk.  Entries arrive in position order, so the forward strand (pos <= n0 - length + 1) is complete
// before the first reverse-strand entry (pos >= n0 + 2):
//   * for a forward-strand entry the mirrored window lies on the reverse strand, above every kept site: the walk skips
//     nothing, lands on the last kept site and finds it below revPos - length + 1 (that would need pos > n0 - length + 2).  Only
//     the forward rule applies.
//   * for a reverse-strand entry every kept reverse-strand site lies above revPos + length - 1 <= n0 and is skipped; the
//     walk ends on the kept FORWARD site with the largest startPos <= revPos + length - 1.  As pos grows that bound falls, so
//     one cursor into the (now frozen) forward kept list, moving down, replaces the walk; a forward site that loses is
//     tomb-stoned and the cursor steps past it, exactly what removing it from the array does to later walks.
//   * the forward rule of a reverse-strand entry can only meet the last kept reverse-strand site (the last forward site ends
//     at n0 at the latest; an earlier reverse site is more than `length` below the one it did not overlap).
// Only keep flags, kept classes and their histogram leave the kernel, so the order of the kept array is immaterial.
//
// A block stages a tile of 2048 entries (position, 16-bit class, sequence) in shared memory with coalesced vector loads,
// compacts the heads of the sequence runs that START in the tile, and one thread per head processes its run out of shared
// memory: O(1) work per entry, two short data-dependent loops.  The last run of a tile may leave it: that run alone is
// replayed from global memory with the reference's literal loop (walk_global), so results do not depend on the tiling.
// As in k_select_cls_fwd the comparison "(double)float(fin[kept]) <= fin[incoming]" is "c_incoming >= T[c_kept]" (the host
// verifies that fin[] is non-decreasing in the class index), classes >= nSite are not sites, and no float score is written.
#pragma once
#include "xxm_select3.cuh"

namespace xxm {

constexpr int kSel4Tile = 2048;
constexpr int kSel4Threads = 256;

struct SelectRcIO {
	const uint32_t* seq;
	const uint32_t* pos;
	const uint32_t* vidx;
	uint32_t n;                      // entries, or -- with nDev -- the capacity of the arrays
	const unsigned long long* nDev;  // non-null: entry count on the device (written by the expand kernel)
	int length;                      // motif span
	const uint16_t* clsThr;          // [nval] T[k]
	uint32_t nSite;                  // classes [0, nSite) are sites
	uint32_t nval;
	uint32_t* stk;                   // [n] scratch: kept array of runs that leave their tile
	uint8_t* keep;                   // [n]
	uint16_t* kcls;                  // [n] value index of a kept entry, 0xFFFF for a dropped one
	uint32_t* counters;              // [0] += kept sites, [1] = max kept sites of one sequence
	uint32_t* hist;                  // [nval], zeroed
};

// the reference's loop, literally, over entries [a, b) of one sequence in global memory (runs that leave their tile)
// (arguments by value: a reference to the kernel's parameter struct would force a copy of it into local memory)
struct RcGlobalArgs {
	const uint32_t* pos;
	const uint32_t* vidx;
	uint32_t* stk;
	uint8_t* keep;
	uint16_t* kcls;
	uint32_t nSite;
	int length;
};
__device__ __noinline__ uint32_t select_rc_walk_global(RcGlobalArgs io, const uint16_t* s_thr, uint32_t* s_hist, uint32_t a, uint32_t b, int seqLen) {
	const int L = io.length;
	uint32_t* stk = io.stk + a;
	int kc = 0;
	auto cls_of = [&](uint32_t i) { return min(io.vidx[i], 0xFFFFu); };
	for (uint32_t i = a; i < b; ++i) {
		io.keep[i] = 0;
		io.kcls[i] = 0xFFFFu;
		const uint32_t c = cls_of(i);
		if (c >= io.nSite) continue;
		const int p = static_cast<int>(io.pos[i]);
		bool reject = false;
		if (kc > 0) {
			const uint32_t t = stk[kc - 1];
			if (static_cast<int>(io.pos[t]) + L >= p) {
				if (c >= s_thr[cls_of(t)]) reject = true;
				else --kc;
			}
		}
		if (!reject && p > L / 2) {
			int back = 1;
			const int revPos = seqLen - L + 2 - p;
			while (kc - back >= 0 && static_cast<int>(io.pos[stk[kc - back]]) > revPos + L - 1) ++back;
			if (kc - back >= 0 && static_cast<int>(io.pos[stk[kc - back]]) >= revPos - L + 1) {
				if (c >= s_thr[cls_of(stk[kc - back])]) reject = true;
				else {
					while (back > 1) { stk[kc - back] = stk[kc - back + 1]; --back; }
					--kc;
				}
			}
		}
		if (reject) continue;
		stk[kc++] = i;
	}
	for (int k = 0; k < kc; ++k) {
		const uint32_t i = stk[k], c = cls_of(i);
		io.keep[i] = 1;
		io.kcls[i] = static_cast<uint16_t>(c);
		atomicAdd(&s_hist[c], 1u);
	}
	return static_cast<uint32_t>(kc);
}

__global__ void __launch_bounds__(kSel4Threads, 6) k_select_cls_rc(SeqStream S, SelectRcIO io) {
	extern __shared__ __align__(16) uint8_t smem4[];
	uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem4);
	uint16_t* s_thr = reinterpret_cast<uint16_t*>(s_hist + io.nval);
	__shared__ __align__(16) uint32_t s_pos[kSel4Tile];
	__shared__ __align__(16) uint32_t s_seq[kSel4Tile];
	__shared__ __align__(16) uint16_t s_cls[kSel4Tile];
	__shared__ __align__(16) uint16_t s_kc[kSel4Tile];    // out: class of a kept entry, 0xFFFF otherwise
	__shared__ uint16_t s_fwd[kSel4Tile];                 // kept forward-strand sites of a run, in place: slot k of the run at [head + k]
	__shared__ uint16_t s_head[kSel4Tile];
	__shared__ uint32_t s_wsum[kSel4Threads / 32];
	__shared__ uint32_t s_first, s_last;
	for (uint32_t k = threadIdx.x; k < io.nval; k += blockDim.x) { s_hist[k] = 0; s_thr[k] = io.clsThr[k]; }
	const int L = io.length;
	const uint32_t nSite = io.nSite;
	const uint32_t n = io.nDev ? static_cast<uint32_t>(min(*io.nDev, static_cast<unsigned long long>(io.n))) : io.n;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	uint32_t keptTot = 0, keptMax = 0;
	for (uint32_t base = blockIdx.x * kSel4Tile; base < n; base += gridDim.x * kSel4Tile) {
		__syncthreads();  // the previous tile is done with shared memory (and the tables are in place)
		const uint32_t tlen = min(static_cast<uint32_t>(kSel4Tile), n - base);
		// ---- stage the tile: 8 consecutive entries per thread
		{
			const uint32_t r0 = 8 * threadIdx.x;
			uint32_t sq[8], ps[8], vi[8];
			if (r0 + 8 <= tlen) {
				const uint4 a0 = *reinterpret_cast<const uint4*>(io.seq + base + r0), a1 = *reinterpret_cast<const uint4*>(io.seq + base + r0 + 4);
				const uint4 b0 = *reinterpret_cast<const uint4*>(io.pos + base + r0), b1 = *reinterpret_cast<const uint4*>(io.pos + base + r0 + 4);
				const uint4 v0 = *reinterpret_cast<const uint4*>(io.vidx + base + r0), v1 = *reinterpret_cast<const uint4*>(io.vidx + base + r0 + 4);
				sq[0] = a0.x; sq[1] = a0.y; sq[2] = a0.z; sq[3] = a0.w; sq[4] = a1.x; sq[5] = a1.y; sq[6] = a1.z; sq[7] = a1.w;
				ps[0] = b0.x; ps[1] = b0.y; ps[2] = b0.z; ps[3] = b0.w; ps[4] = b1.x; ps[5] = b1.y; ps[6] = b1.z; ps[7] = b1.w;
				vi[0] = v0.x; vi[1] = v0.y; vi[2] = v0.z; vi[3] = v0.w; vi[4] = v1.x; vi[5] = v1.y; vi[6] = v1.z; vi[7] = v1.w;
			} else {
#pragma unroll
				for (int k = 0; k < 8; ++k) {
					const bool in = r0 + k < tlen;
					sq[k] = in ? io.seq[base + r0 + k] : 0u;
					ps[k] = in ? io.pos[base + r0 + k] : 0u;
					vi[k] = in ? io.vidx[base + r0 + k] : kNone;
				}
			}
#pragma unroll
			for (int k = 0; k < 8; ++k) vi[k] = min(vi[k], 0xFFFFu);
			*reinterpret_cast<uint4*>(s_seq + r0) = make_uint4(sq[0], sq[1], sq[2], sq[3]);
			*reinterpret_cast<uint4*>(s_seq + r0 + 4) = make_uint4(sq[4], sq[5], sq[6], sq[7]);
			*reinterpret_cast<uint4*>(s_pos + r0) = make_uint4(ps[0], ps[1], ps[2], ps[3]);
			*reinterpret_cast<uint4*>(s_pos + r0 + 4) = make_uint4(ps[4], ps[5], ps[6], ps[7]);
			*reinterpret_cast<uint4*>(s_cls + r0) = make_uint4(vi[0] | (vi[1] << 16), vi[2] | (vi[3] << 16), vi[4] | (vi[5] << 16), vi[6] | (vi[7] << 16));
			*reinterpret_cast<uint4*>(s_kc + r0) = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
		}
		if (threadIdx.x == 0) { s_first = kSel4Tile; s_last = kSel4Tile; }
		__syncthreads();
		// ---- heads of the runs that start in this tile, compacted
		uint32_t hm = 0;
		{
			const uint32_t r0 = 8 * threadIdx.x;
#pragma unroll
			for (int k = 0; k < 8; ++k) {
				const uint32_t r = r0 + k;
				if (r >= tlen) break;
				const uint32_t prev = r ? s_seq[r - 1] : (base ? io.seq[base - 1] : 0u);
				if (s_seq[r] != prev) hm |= 1u << k;
			}
		}
		const uint32_t cnt = __popc(hm);
		uint32_t incl = cnt;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
			if (lane >= d) incl += t;
		}
		if (lane == 31) s_wsum[wid] = incl;
		__syncthreads();
		uint32_t off = incl - cnt, nheads = 0;
#pragma unroll
		for (int k = 0; k < kSel4Threads / 32; ++k) { if (k < wid) off += s_wsum[k]; nheads += s_wsum[k]; }
		if (hm) {
			if (off == 0) s_first = 8 * threadIdx.x + (__ffs(hm) - 1);  // the first head of the tile: entries before it belong to the previous tile's last run
#pragma unroll
			for (int k = 0; k < 8; ++k)
				if ((hm >> k) & 1u) s_head[off++] = static_cast<uint16_t>(8 * threadIdx.x + k);
		}
		__syncthreads();
		// ---- one thread per run
		for (uint32_t h = threadIdx.x; h < nheads; h += kSel4Threads) {
			const uint32_t r0 = s_head[h];
			const uint32_t sq = s_seq[r0];
			const int seqLen = S.uniformLen ? S.uniformLen : S.len[sq - 1];
			const bool lastRun = h + 1 == nheads;
			uint32_t rend = lastRun ? tlen : static_cast<uint32_t>(s_head[h + 1]);
			if (lastRun && base + tlen < n && io.seq[base + tlen] == sq) {
				// the run leaves the tile: replay all of it from global memory (the tile store below skips its entries)
				uint32_t e = base + tlen;
				while (e < n && io.seq[e] == sq) ++e;
				s_last = r0;
				const RcGlobalArgs ga{io.pos, io.vidx, io.stk, io.keep, io.kcls, io.nSite, io.length};
				const uint32_t kc = select_rc_walk_global(ga, s_thr, s_hist, base + r0, e, seqLen);
				keptTot += kc;
				keptMax = max(keptMax, kc);
				continue;
			}
			const int n0 = (seqLen - 1) / 2;   // the forward strand ends here; position n0 + 1 is the separator
			int nF = 0;                        // kept forward-strand sites: s_fwd[r0 .. r0 + nF)
			uint32_t r = r0;
			// forward strand: the forward rule only
			for (; r < rend; ++r) {
				const int p = static_cast<int>(s_pos[r]);
				if (p > n0 + 1) break;
				const uint32_t c = s_cls[r];
				if (c >= nSite) continue;
				if (nF > 0) {
					const uint32_t t = s_fwd[r0 + nF - 1];
					if (static_cast<int>(s_pos[t]) + L >= p) {
						if (c >= s_thr[s_cls[t]]) continue;
						s_kc[t] = 0xFFFFu;
						--nF;
					}
				}
				s_fwd[r0 + nF++] = static_cast<uint16_t>(r);
				s_kc[r] = static_cast<uint16_t>(c);
			}
			// reverse strand: forward rule against the last kept reverse site, then the cursor into the forward list
			int fp = nF - 1;       // forward sites above fp lie above every later bound (or are gone)
			int alive = nF;
			int lastRev = -1;
			for (; r < rend; ++r) {
				const uint32_t c = s_cls[r];
				if (c >= nSite) continue;
				const int p = static_cast<int>(s_pos[r]);
				if (lastRev >= 0 && static_cast<int>(s_pos[lastRev]) + L >= p) {
					if (c >= s_thr[s_cls[lastRev]]) continue;
					s_kc[lastRev] = 0xFFFFu;  // popped (stays popped even if the look-back rejects the new site)
					lastRev = -1;
					--alive;
				}
				if (p > L / 2) {
					const int revPos = seqLen - L + 2 - p;
					while (fp >= 0 && static_cast<int>(s_pos[s_fwd[r0 + fp]]) > revPos + L - 1) --fp;
					if (fp >= 0) {
						const uint32_t t = s_fwd[r0 + fp];
						if (static_cast<int>(s_pos[t]) >= revPos - L + 1) {
							if (c >= s_thr[s_cls[t]]) continue;
							s_kc[t] = 0xFFFFu;  // the forward site loses: gone for every later walk
							--fp;
							--alive;
						}
					}
				}
				s_kc[r] = static_cast<uint16_t>(c);
				lastRev = static_cast<int>(r);
				++alive;
			}
			keptMax = max(keptMax, static_cast<uint32_t>(alive));
		}
		__syncthreads();
		// ---- flags, classes and histogram of the entries this tile owns: from its first head to its end (or to the run that
		// was replayed from global memory)
		const uint32_t first = s_first, last = min(tlen, s_last);
		for (uint32_t r = threadIdx.x; r < last; r += kSel4Threads) {
			if (r < first) continue;
			const uint16_t c = s_kc[r];
			io.keep[base + r] = (c != 0xFFFFu);
			io.kcls[base + r] = c;
			if (c != 0xFFFFu) { atomicAdd(&s_hist[c], 1u); ++keptTot; }
		}
	}
	keptTot = __reduce_add_sync(0xFFFFFFFFu, keptTot);
	keptMax = __reduce_max_sync(0xFFFFFFFFu, keptMax);
	if (lane == 0 && (keptTot || keptMax)) { atomicAdd(&io.counters[0], keptTot); atomicMax(&io.counters[1], keptMax); }
	__syncthreads();
	for (uint32_t k = threadIdx.x; k < io.nval; k += blockDim.x)
		if (s_hist[k]) atomicAdd(&io.hist[k], s_hist[k]);
}

}  // namespace xxm

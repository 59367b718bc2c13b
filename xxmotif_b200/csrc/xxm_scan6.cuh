// K2 (sixth generation): the full scan without any inter-CTA dependency.
//
// The fifth generation's expand kernel was latency bound, not instruction bound (ncu: 45 % of the warp cycles at CTA
// barriers -- waiting for warp 0's decoupled look-back -- and 24 % on the loads in front of them).  What forced the
// look-back is that the number of sites of a tile was only known inside the expand kernel, after the windows with an N in
// a match column had been removed.  Now
//
//   k_dead_mask  (once per COLUMN SET, cached across the PWMs of a refinement run): for every word of the stream the
//                windows that can never be a site -- an N (or sequence padding) under one of the match columns,
//                branch_and_bound-inl.h:388 -- in the find kernel's even/odd bit order;
//   k_scan5      (find) clears those windows itself, so its masks are final, and counts them: one count per 64-word
//                chunk (a warp's unit) and, by one atomic per chunk, one per 1024-word tile;
//   k_tile_scan  one block: exclusive prefix over the tile counts (32 k tiles for 10^9 windows) + grand total;
//   k_expand6    every WARP expands its own 128 words: offset = tile offset + prefix over the tile's 16 chunk counts.  No
//                barrier, no status words, no spinning; a warp with no hit leaves after two loads.
//
// Records still come out globally ordered by (sequence, position).
#pragma once
#include "xxm_scan5.cuh"

namespace xxm {

constexpr int kChunkWords = 64;                           // the find kernel's unit: one warp, two words per lane
constexpr int kChunksPerTile = kTileWords / kChunkWords;  // 16

// window order -> even/odd order (inverse of eo_to_natural)
__device__ __forceinline__ uint32_t natural_to_eo(uint32_t x) {
	uint32_t t;
	t = (x ^ (x >> 1)) & 0x22222222u; x = x ^ t ^ (t << 1);
	t = (x ^ (x >> 2)) & 0x0C0C0C0Cu; x = x ^ t ^ (t << 2);
	t = (x ^ (x >> 4)) & 0x00F000F0u; x = x ^ t ^ (t << 4);
	return __byte_perm(x, 0u, 0x2031);
}

struct DeadCols {
	int L;           // match columns
	int contiguous;  // no gap columns: the union over the columns is a run of L positions
	int coloff[32];
};

// dead[w] (even/odd order, or window order with kNatural): window j of word w has an N under a match column.  Words [0, nwords + 8).
template <bool kNatural>
__global__ void k_dead_mask(const uint32_t* __restrict__ nmask, uint32_t nwords, DeadCols C, uint32_t* __restrict__ dead) {
	const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= nwords + 8) return;
	if (w >= nwords) { dead[w] = 0xFFFFFFFFu; return; }
	const uint64_t mm = (static_cast<uint64_t>(nmask[w + 1]) << 32) | nmask[w];  // nmask has nwords + 2 entries
	uint64_t x = mm;
	if (C.contiguous) {  // OR of (mm >> 0) .. (mm >> (L - 1)) by doubling
		int have = 1;
		while (have < C.L) {
			const int sh = min(have, C.L - have);
			x |= x >> sh;
			have += sh;
		}
	} else {
		for (int c = 1; c < C.L; ++c) x |= mm >> C.coloff[c];
	}
	dead[w] = kNatural ? static_cast<uint32_t>(x) : natural_to_eo(static_cast<uint32_t>(x));
}

// exclusive prefix over the tile counts.  Block b owns tiles [1024 b, 1024 b + 1024): it adds up everything before them
// itself (coalesced, all loads independent; the counts sit in L2 -- 128 KB for 10^9 windows) and scans its own 1024.
__global__ void __launch_bounds__(1024) k_tile_scan(const uint32_t* __restrict__ tileCount, uint32_t ntiles, uint64_t* __restrict__ tileOff,
                                                    uint64_t* __restrict__ total) {
	__shared__ uint64_t s_part[32];
	__shared__ uint32_t s_wsum[32];
	const uint32_t t = threadIdx.x, lane = t & 31u, wid = t >> 5;
	const uint32_t first = blockIdx.x * 1024u;
	uint64_t s = 0;
	for (uint32_t i = t; i < first; i += 1024u) s += tileCount[i];
#pragma unroll
	for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, d);
	const uint32_t g = first + t;
	const uint32_t c = (g < ntiles) ? tileCount[g] : 0u;
	uint32_t incl = c;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
		if (lane >= static_cast<uint32_t>(d)) incl += v;
	}
	if (lane == 0) s_part[wid] = s;
	if (lane == 31) s_wsum[wid] = incl;
	__syncthreads();
	uint64_t prefix = 0;
	uint32_t wbefore = 0;
#pragma unroll
	for (int k = 0; k < 32; ++k) {
		prefix += s_part[k];
		if (static_cast<uint32_t>(k) < wid) wbefore += s_wsum[k];
	}
	const uint64_t excl = prefix + wbefore + (incl - c);
	if (g < ntiles) tileOff[g] = excl;
	if (g + 1 == ntiles) { tileOff[ntiles] = excl + c; *total = excl + c; }
}

constexpr int kExpand6Threads = 256;
constexpr int kExpand6Warps = kExpand6Threads / 32;
constexpr int kWarpWords = kTileWords / kExpand6Warps;  // 128: four words per lane
constexpr int kWarpQueue = 512;                         // queue entries per warp and round

struct Expand6IO {
	const uint32_t* hitmask;     // final masks of the find kernel (even/odd order, dead windows cleared)
	const uint32_t* chunkCount;  // [ntiles * 16]
	const uint64_t* tileOff;     // [ntiles + 1]
	uint32_t scanWords;
	uint32_t *oseq, *opos, *ovidx;
	uint32_t cap;
	uint32_t wpsMagic;           // floor(2^32 / uniformWps), 0 for ragged sets
};

template <bool kFast>
__device__ __forceinline__ HitRec expand6_hit(const SeqStream& S, const Model& M, uint32_t wpsMagic, const uint64_t* s_codes, uint32_t wbase, uint32_t e) {
	const uint32_t wi = e >> 5;
	const uint32_t w = wbase + wi;
	const int j = static_cast<int>(e & 31u);
	HitRec r;
	if (kFast) {
		const uint2 c0 = *reinterpret_cast<const uint2*>(s_codes + wi);
		const uint32_t c1 = static_cast<uint32_t>(s_codes[wi + 1]);
		const uint32_t k = (j < 16) ? __funnelshift_r(c0.x, c0.y, 2 * j) : __funnelshift_r(c0.y, c1, 2 * j - 32);
		r.v = M.rank0[k & ((1u << (2 * M.len0)) - 1u)] - 1u;  // the window is in the table: the find kernel is exact
	} else {
		r.v = eval_window(M, s_codes[wi], s_codes[wi + 1], 0u, 0u, j);  // N windows are already gone
	}
	uint32_t sq, wo;
	if (wpsMagic) {
		sq = __umulhi(w, wpsMagic);
		wo = sq * S.uniformWps;
		if (w - wo >= S.uniformWps) { ++sq; wo += S.uniformWps; }
	} else {
		sq = find_seq(S, w);
		wo = S.wordOff[sq];
	}
	r.seq = sq + 1;
	r.pos = (w - wo) * 32 + static_cast<uint32_t>(j) + 1;
	return r;
}

template <bool kFast>
__global__ void __launch_bounds__(kExpand6Threads, 6) k_expand6(SeqStream S, Model M, Expand6IO io) {
	__shared__ __align__(16) uint64_t s_codesAll[kExpand6Warps][kWarpWords + 2];
	__shared__ uint16_t s_qAll[kExpand6Warps][kWarpQueue];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t tile = blockIdx.x;
	uint64_t* s_codes = s_codesAll[wid];
	uint16_t* s_q = s_qAll[wid];
	const uint32_t wbase = tile * kTileWords + wid * kWarpWords;
	const uint32_t w0 = wbase + 4 * lane;
	// ---- every load of the warp is issued before anything waits: the chunk counts and the tile offset (where the records
	// go), the lane's four masks and four code words (staged for the expansion).  A warp without a hit -- rare -- has
	// then read its words for nothing, but nobody pays a second DRAM round trip.
	uint32_t cc = 0;
	if (lane < kChunksPerTile) cc = __ldg(io.chunkCount + tile * kChunksPerTile + lane);
	const uint64_t tileBase = __ldg(io.tileOff + tile);
	uint32_t h[4];
	const bool full = w0 + 4 <= io.scanWords && w0 + 4 <= S.nwords + 2;
	uint4 ca = make_uint4(0u, 0u, 0u, 0u), cb = ca;
	if (full) {
		const uint4 hv = __ldg(reinterpret_cast<const uint4*>(io.hitmask + w0));
		ca = __ldg(reinterpret_cast<const uint4*>(S.codes + w0));
		cb = __ldg(reinterpret_cast<const uint4*>(S.codes + w0 + 2));
		h[0] = hv.x; h[1] = hv.y; h[2] = hv.z; h[3] = hv.w;
	} else {
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			const uint32_t w = w0 + k;
			h[k] = (w < io.scanWords) ? io.hitmask[w] : 0u;
			const uint64_t c = (w < S.nwords + 2) ? S.codes[w] : 0ull;
			if (k == 0) { ca.x = static_cast<uint32_t>(c); ca.y = static_cast<uint32_t>(c >> 32); }
			if (k == 1) { ca.z = static_cast<uint32_t>(c); ca.w = static_cast<uint32_t>(c >> 32); }
			if (k == 2) { cb.x = static_cast<uint32_t>(c); cb.y = static_cast<uint32_t>(c >> 32); }
			if (k == 3) { cb.z = static_cast<uint32_t>(c); cb.w = static_cast<uint32_t>(c >> 32); }
		}
	}
	uint64_t ctail = 0;
	if (lane == 0 && wbase + kWarpWords < S.nwords + 2) ctail = __ldg(S.codes + wbase + kWarpWords);
	// ---- where this warp's records go: tile offset + the counts of the chunks before it inside the tile
	const uint32_t mine = __shfl_sync(0xFFFFFFFFu, cc, 2 * wid) + __shfl_sync(0xFFFFFFFFu, cc, 2 * wid + 1);
	if (mine == 0) return;
	uint32_t before = (lane < 2 * wid) ? cc : 0u;
#pragma unroll
	for (int d = 16; d; d >>= 1) before += __shfl_xor_sync(0xFFFFFFFFu, before, d);
	const uint64_t base = tileBase + before;
	*reinterpret_cast<uint4*>(s_codes + 4 * lane) = ca;
	*reinterpret_cast<uint4*>(s_codes + 4 * lane + 2) = cb;
	if (lane == 0) s_codes[kWarpWords] = ctail;
#pragma unroll
	for (int k = 0; k < 4; ++k) h[k] = eo_to_natural(h[k]);
	const uint32_t cnt = __popc(h[0]) + __popc(h[1]) + __popc(h[2]) + __popc(h[3]);
	uint32_t incl = cnt;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
		if (lane >= d) incl += t;
	}
	const uint32_t myOff = incl - cnt;
	// records [base, base + mine) of the three output arrays; room = how many of them fit the arrays
	const uint32_t room = (base < io.cap) ? static_cast<uint32_t>(min(static_cast<uint64_t>(io.cap) - base, static_cast<uint64_t>(mine))) : 0u;
	uint32_t* pseq = io.oseq + base;
	uint32_t* ppos = io.opos + base;
	uint32_t* pvid = io.ovidx + base;
#ifndef XXM_SIMT_EMU
	asm volatile("" : "+l"(pseq), "+l"(ppos), "+l"(pvid));  // keep the three bases in registers: one IMAD.WIDE per store instead of 64-bit adds per record
#endif
	auto emit = [&](const HitRec& rec, uint32_t q) {
		if (q < room) {
#ifdef XXM_SIMT_EMU
			pseq[q] = rec.seq; ppos[q] = rec.pos; pvid[q] = rec.v;
#else
			asm volatile("st.global.u32 [%0], %1;" ::"l"(pseq + q), "r"(rec.seq) : "memory");
			asm volatile("st.global.u32 [%0], %1;" ::"l"(ppos + q), "r"(rec.pos) : "memory");
			asm volatile("st.global.u32 [%0], %1;" ::"l"(pvid + q), "r"(rec.v) : "memory");
#endif
		}
	};
	if (mine <= static_cast<uint32_t>(kWarpQueue)) {
		// the four words of the lane in lock step, highest window first, every word filling its queue range backwards: the
		// trip count of the warp is the largest hit count of any of its 128 words
		uint32_t r0 = h[0], r1 = h[1], r2 = h[2], r3 = h[3];
		smem_addr_t q0 = static_cast<smem_addr_t>(__cvta_generic_to_shared(s_q + myOff)) + 2u * __popc(r0);  // one past the last slot of word 0
		smem_addr_t q1 = q0 + 2u * __popc(r1);
		smem_addr_t q2 = q1 + 2u * __popc(r2);
		smem_addr_t q3 = q2 + 2u * __popc(r3);
		const uint32_t tag = static_cast<uint32_t>(4 * lane) << 5;
		while (r0 | r1 | r2 | r3) {
			pop_msb(r0, q0, tag);
			pop_msb(r1, q1, tag + 32);
			pop_msb(r2, q2, tag + 64);
			pop_msb(r3, q3, tag + 96);
		}
		__syncwarp();
		for (uint32_t q = lane; q < mine; q += 32) emit(expand6_hit<kFast>(S, M, io.wpsMagic, s_codes, wbase, s_q[q]), q);
		return;
	}
	// dense warp: queue the hits in rounds
	for (uint32_t qbase = 0; qbase < mine; qbase += kWarpQueue) {
		__syncwarp();
		uint32_t o = myOff;
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			uint32_t rem = h[k];
			while (rem) {
				const int j = __ffs(rem) - 1;
				rem &= rem - 1;
				if (o >= qbase && o < qbase + kWarpQueue) s_q[o - qbase] = static_cast<uint16_t>(((4 * lane + k) << 5) | j);
				++o;
			}
		}
		__syncwarp();
		const uint32_t nq = min(static_cast<uint32_t>(kWarpQueue), mine - qbase);
		for (uint32_t q = lane; q < nq; q += 32) emit(expand6_hit<kFast>(S, M, io.wpsMagic, s_codes, wbase, s_q[q]), qbase + q);
	}
}

}  // namespace xxm

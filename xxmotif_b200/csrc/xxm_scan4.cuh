// K2 (fourth generation): branch-free "find" kernel with a dense hit bitmap as output, and a balanced "expand" kernel.
//
// find   (k_scan4):  one thread per 32-base word.  16 LDS.U8 per word into the 64 KB shared-memory table T9, indexed by
//                    9 consecutive bases (18 bits; the byte is addressed by bases 0..7, base 8 selects one of its four
//                    2-bit entries).  An entry holds the EXACT answers for the two windows that start at base 0 and at
//                    base 1 of the group (left k-mer in the table) whenever all left match columns fit into the 9
//                    bases -- true for every ungapped motif; otherwise the bits are projections and the word takes one
//                    exact bitmap test per flagged window.  No compaction, no divergent loops in the common case: the
//                    thread ORs its 16 x 2 bits into a 32-bit mask (bit j = window starting at base j of the word is a
//                    site candidate), clears windows that touch an N, and stores the mask (4 B per 32 windows).  Each
//                    warp also sums the hits of a 1024-word tile.
// expand (k_expand4): one CTA per tile: prefix over the masks, hit offsets into a shared-memory queue in (word, bit)
//                    order, then all threads resolve queue entries into records (value index, sequence, position, float
//                    score) at the tile's global offset -- coalesced, balanced, and globally ordered by construction.
#pragma once
#include "xxm_kernels.cuh"

namespace xxm {

constexpr int kScan4Threads = 512;
constexpr int kScan4Warps = kScan4Threads / 32;
constexpr int kTileWords = 1024;

struct T9Build {
	int nvis[2];            // visible left match columns for window offsets k = 0, 1 inside the 9-base group
	int visoff[2][9];       // base offset of the visible column inside the group
	uint32_t poff[2];       // word offset of the projection bitmap of offset k inside `pbits`
	const uint32_t* pbits;
};

// one thread per table byte: bases 0..7 = byte index, base 8 = entry inside the byte
__global__ void k_build_t9(T9Build B, uint8_t* __restrict__ t9) {
	const uint32_t lo = blockIdx.x * blockDim.x + threadIdx.x;
	if (lo >= 65536u) return;
	uint32_t byte = 0;
	for (uint32_t b8 = 0; b8 < 4; ++b8) {
		const uint32_t idx = lo | (b8 << 16);
#pragma unroll
		for (int k = 0; k < 2; ++k) {
			const int m = B.nvis[k];
			uint32_t pid = 0;
			for (int c = 0; c < m; ++c) pid |= ((idx >> (2 * B.visoff[k][c])) & 3u) << (2 * c);
			const uint32_t bit = (m == 0) ? 1u : (B.pbits[B.poff[k] + (pid >> 5)] >> (pid & 31)) & 1u;
			byte |= bit << (2 * b8 + k);
		}
	}
	t9[lo] = static_cast<uint8_t>(byte);
}

template <bool kSmemBitmap>
__global__ void __launch_bounds__(kScan4Threads, 3) k_scan4(SeqStream S, Model M, uint32_t* __restrict__ hitmask, uint32_t* __restrict__ tileCount,
                                                            uint32_t scanWords, uint32_t wordsPerWarp) {
	extern __shared__ __align__(16) uint8_t smem4[];
	uint8_t* s_t9 = smem4;  // 64 KB
	uint32_t* s_bm = reinterpret_cast<uint32_t*>(smem4 + 65536);
	{
		const uint4* src = reinterpret_cast<const uint4*>(M.t8);
		uint4* dst = reinterpret_cast<uint4*>(s_t9);
		for (uint32_t i = threadIdx.x; i < 65536u / 16u; i += blockDim.x) dst[i] = src[i];
	}
	const uint32_t* bm = M.hit0;
	if (kSmemBitmap) {
		const uint32_t nw = max(1u, (1u << (2 * M.len0)) / 32u);
		for (uint32_t i = threadIdx.x; i < nw; i += blockDim.x) s_bm[i] = M.hit0[i];
		bm = s_bm;
	}
	__syncthreads();
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t gw = blockIdx.x * kScan4Warps + wid;
	const uint32_t wBeg = gw * wordsPerWarp;  // multiple of kTileWords
	const uint32_t wEnd = min(scanWords, wBeg + wordsPerWarp);
	const uint32_t inexact = ((M.exactBits & 1u) ? 0u : 0x55555555u) | ((M.exactBits & 2u) ? 0u : 0xAAAAAAAAu);
	uint32_t tileHits = 0;
	for (uint32_t base = wBeg; base < wEnd; base += 32) {
		const uint32_t w = base + lane;
		uint32_t hits = 0;
		if (w < wEnd) {
			const uint32_t m0 = S.nmask[w];
			if (m0 != 0xFFFFFFFFu) {  // not a padding word
				const uint64_t c0 = S.codes[w], c1 = S.codes[w + 1];
				const uint32_t r0 = static_cast<uint32_t>(c0), r1 = static_cast<uint32_t>(c0 >> 32), r2 = static_cast<uint32_t>(c1);
#pragma unroll
				for (int m = 0; m < 16; ++m) {
					const uint32_t t = (m < 8) ? __funnelshift_r(r0, r1, 4 * m) : __funnelshift_r(r1, r2, 4 * (m - 8));
					const uint32_t byte = s_t9[t & 0xFFFFu];
					hits |= ((byte >> ((t >> 15) & 6u)) & 3u) << (2 * m);
				}
				const uint32_t m1 = S.nmask[w + 1];
				if (m0 | m1) {  // an N (or padding) within reach: a window dies if any match column holds an N (branch_and_bound-inl.h:388)
					const uint64_t mm = (static_cast<uint64_t>(m1) << 32) | m0;
					uint32_t inv = 0;
					for (int k = 0; k < M.L; ++k) inv |= static_cast<uint32_t>(mm >> M.coloff[k]);
					hits &= ~inv;
				}
				uint32_t rem = hits & inexact;  // gapped column sets only: the table bits were projections
				while (rem) {
					const int j = __ffs(rem) - 1;
					rem &= rem - 1;
					uint32_t id = 0;
					for (int k = 0; k < M.len0; ++k) id |= base_at(c0, c1, j + M.coloff[k]) << (2 * k);
					if (!((bm[id >> 5] >> (id & 31)) & 1u)) hits &= ~(1u << j);
				}
			}
			hitmask[w] = hits;
		}
		tileHits += __popc(hits);
		if (((base + 32 - wBeg) & (kTileWords - 1)) == 0 || base + 32 >= wEnd) {  // end of a tile (or of the range)
			uint32_t tsum = tileHits;
#pragma unroll
			for (int d = 16; d; d >>= 1) tsum += __shfl_xor_sync(0xFFFFFFFFu, tsum, d);
			if (lane == 0) tileCount[base / kTileWords] = tsum;
			tileHits = 0;
		}
	}
}

constexpr int kExpandThreads = 256;
constexpr int kExpandQueue = 8192;

__global__ void __launch_bounds__(kExpandThreads) k_expand4(SeqStream S, Model M, Values V, const uint32_t* __restrict__ hitmask, const uint32_t* __restrict__ tileOff,
                                                            uint32_t scanWords, int wantScore, uint32_t* __restrict__ oseq, uint32_t* __restrict__ opos,
                                                            uint32_t* __restrict__ ovidx, float* __restrict__ oscore, uint32_t cap) {
	__shared__ uint16_t s_q[kExpandQueue];
	__shared__ uint32_t s_warp[kExpandThreads / 32];
	const uint32_t tile = blockIdx.x;
	const uint32_t wbase = tile * kTileWords;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	uint32_t h[4];
	{
		const uint4 hv = (wbase + 4 * tid + 3 < scanWords) ? reinterpret_cast<const uint4*>(hitmask + wbase)[tid] : make_uint4(0, 0, 0, 0);
		h[0] = hv.x; h[1] = hv.y; h[2] = hv.z; h[3] = hv.w;
		if (!(wbase + 4 * tid + 3 < scanWords))
			for (int k = 0; k < 4; ++k) { const uint32_t w = wbase + 4 * tid + k; h[k] = (w < scanWords) ? hitmask[w] : 0u; }
	}
	const uint32_t cnt = __popc(h[0]) + __popc(h[1]) + __popc(h[2]) + __popc(h[3]);
	uint32_t incl = cnt;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
		if (lane >= d) incl += t;
	}
	if (lane == 31) s_warp[wid] = incl;
	__syncthreads();
	uint32_t woff = 0, total = 0;
#pragma unroll
	for (int k = 0; k < kExpandThreads / 32; ++k) { if (k < wid) woff += s_warp[k]; total += s_warp[k]; }
	if (total == 0) return;
	const uint32_t myOff = woff + incl - cnt;
	const uint32_t dst = tileOff[tile];
	for (uint32_t qbase = 0; qbase < total; qbase += kExpandQueue) {
		__syncthreads();
		{  // queue the hits whose rank falls into [qbase, qbase + kExpandQueue)
			uint32_t o = myOff;
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				uint32_t rem = h[k];
				while (rem) {
					const int j = __ffs(rem) - 1;
					rem &= rem - 1;
					if (o >= qbase && o < qbase + kExpandQueue) s_q[o - qbase] = static_cast<uint16_t>(((4 * tid + k) << 5) | j);
					++o;
				}
			}
		}
		__syncthreads();
		const uint32_t nq = min(static_cast<uint32_t>(kExpandQueue), total - qbase);
		for (uint32_t q = tid; q < nq; q += kExpandThreads) {
			const uint32_t e = s_q[q];
			const uint32_t wi = e >> 5;
			const uint32_t w = wbase + wi;
			const int j = static_cast<int>(e & 31u);
			const uint32_t v = eval_window(M, S.codes[w], S.codes[w + 1], S.nmask[w], S.nmask[w + 1], j);
			uint32_t sq, wo;
			if (S.uniformWps) { sq = w / S.uniformWps; wo = sq * S.uniformWps; }
			else { sq = find_seq(S, w); wo = S.wordOff[sq]; }
			const uint32_t o = dst + qbase + q;
			if (o < cap) {
				oseq[o] = sq + 1;
				opos[o] = (w - wo) * 32 + static_cast<uint32_t>(j) + 1;
				ovidx[o] = v;
				if (wantScore) oscore[o] = (v == kNone) ? 1.0f : static_cast<float>(V.pS[v]);
			}
		}
	}
}

// exclusive prefix over the tile counts, three small kernels (block sums -> scan of block sums -> add back)
__global__ void __launch_bounds__(1024) k_tile_blocksum(const uint32_t* __restrict__ cnt, uint32_t n, uint32_t* __restrict__ bsum) {
	__shared__ uint32_t s_w[32];
	const uint32_t i = blockIdx.x * 1024 + threadIdx.x;
	uint32_t v = (i < n) ? cnt[i] : 0u;
#pragma unroll
	for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
	if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
	__syncthreads();
	if (threadIdx.x < 32) {
		uint32_t t = s_w[threadIdx.x];
#pragma unroll
		for (int d = 16; d; d >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, d);
		if (threadIdx.x == 0) bsum[blockIdx.x] = t;
	}
}
__global__ void __launch_bounds__(1024) k_tile_offsets(const uint32_t* __restrict__ cnt, uint32_t n, const uint32_t* __restrict__ boff, uint32_t* __restrict__ off) {
	__shared__ uint32_t s_w[32];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t i = blockIdx.x * 1024 + threadIdx.x;
	const uint32_t v = (i < n) ? cnt[i] : 0u;
	uint32_t incl = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
	if (lane == 31) s_w[wid] = incl;
	__syncthreads();
	uint32_t woff = 0;
	for (int k = 0; k < wid; ++k) woff += s_w[k];
	if (i < n) off[i] = boff[blockIdx.x] + woff + incl - v;
}

}  // namespace xxm

// K2 (third generation): the full-scan kernel, split into "find" and "resolve".
//
// find (k_scan3):   per 32-base word, 17 LDS.U8 lookups into the 64 KB shared-memory table T8 (index = 8 consecutive
//                   bases).  Every entry carries three bits about windows that overlap the group:
//                     bit0  window starting at base 0 of the group   (all of its left match columns visible -> exact)
//                     bit1  window starting at base 1                (projection on the columns visible in the group)
//                     bit2  window starting at base -1               (projection on the columns visible in the group)
//                   Even windows (2m) take bit0 of group m; odd windows (2m+1) take bit1 of group m AND bit2 of group
//                   m+1, two overlapping conservative projections whose conjunction is almost exact.  Windows whose
//                   bits are not exact get one exact left-k-mer bitmap test.  Survivors are written as 4-byte global
//                   window indices into the warp's private output region (contiguous range of the stream per warp, so
//                   regions are globally ordered) -- nothing else happens in the hot loop.
// resolve (k_resolve_regions): one thread per found window: value index (rank gather, >8-column path), sequence and
//                   position, float(pValScore); writes the final (sequence, position)-ordered list.  Coalesced and
//                   perfectly balanced, unlike an in-loop evaluation of the ~2 % hits.
#pragma once
#include "xxm_kernels.cuh"

namespace xxm {

constexpr int kScan3Threads = 512;
constexpr int kScan3Warps = kScan3Threads / 32;

struct T8Build3 {
	int nvis[3];            // visible left match columns for window offsets k = 0, +1, -1
	int visoff[3][8];       // base offset of the visible column inside the 8-base group
	uint32_t poff[3];       // word offset of the projection bitmap of offset k inside `pbits`
	const uint32_t* pbits;
};

__global__ void k_build_t8v3(T8Build3 B, uint8_t* __restrict__ t8) {
	const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= 65536u) return;
	uint32_t v = 0;
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		const int m = B.nvis[k];
		uint32_t pid = 0;
		for (int c = 0; c < m; ++c) pid |= ((idx >> (2 * B.visoff[k][c])) & 3u) << (2 * c);
		const uint32_t bit = (m == 0) ? 1u : (B.pbits[B.poff[k] + (pid >> 5)] >> (pid & 31)) & 1u;
		v |= bit << k;
	}
	t8[idx] = static_cast<uint8_t>(v);
}

struct ScanOut3 {
	uint32_t* gpos;      // staging: global window index (word * 32 + offset), warp-private regions of warpCap entries
	uint32_t warpCap;
	uint32_t* warpCount; // [nwarps]
	uint32_t* overflow;
};

// 16 window bits j -> nibble-spaced positions 4(j>>1) + (j&1)
__device__ __forceinline__ uint32_t spread_nibble(uint32_t x) {
	x = (x | (x << 8)) & 0x00FF00FFu;
	x = (x | (x << 4)) & 0x0F0F0F0Fu;
	x = (x | (x << 2)) & 0x33333333u;
	x = (x | (x << 1)) & 0x55555555u;  // bit j at 2j
	return (x & 0x11111111u) | ((x & 0x44444444u) >> 1);
}

template <bool kSmemBitmap>
__global__ void __launch_bounds__(kScan3Threads, 2) k_scan3(SeqStream S, Model M, ScanOut3 O, uint32_t scanWords, uint32_t wordsPerWarp) {
	extern __shared__ __align__(16) uint8_t smem3[];
	uint8_t* s_t8 = smem3;  // 64 KB
	uint32_t* s_bm = reinterpret_cast<uint32_t*>(smem3 + 65536);
	{
		const uint4* src = reinterpret_cast<const uint4*>(M.t8);
		uint4* dst = reinterpret_cast<uint4*>(s_t8);
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
	const uint32_t gw = blockIdx.x * kScan3Warps + wid;
	const uint32_t wBeg = gw * wordsPerWarp;
	const uint32_t wEnd = min(scanWords, wBeg + wordsPerWarp);
	uint32_t* out = O.gpos + static_cast<uint64_t>(gw) * O.warpCap;
	uint32_t ocount = 0;
	const uint32_t inexact = ((M.exactBits & 1u) ? 0u : 0x11111111u) | ((M.exactBits & 2u) ? 0u : 0x22222222u);
	for (uint32_t base = wBeg; base < wEnd; base += 32) {
		const uint32_t w = base + lane;
		uint32_t clo = 0, chi = 0;
		if (w < wEnd) {
			const uint64_t c0 = S.codes[w], c1 = S.codes[w + 1];
			const uint32_t m0 = S.nmask[w];
			const uint32_t r0 = static_cast<uint32_t>(c0), r1 = static_cast<uint32_t>(c0 >> 32), r2 = static_cast<uint32_t>(c1);
			// nibble-spaced accumulators: group m contributes its 3 bits at bit 4m (m = 0..7 -> lo, 8..15 -> hi, 16 -> nx)
			uint32_t lo = 0, hi = 0;
#pragma unroll
			for (int m = 0; m < 8; ++m) lo += static_cast<uint32_t>(s_t8[__funnelshift_r(r0, r1, 4 * m) & 0xFFFFu]) << (4 * m);
#pragma unroll
			for (int m = 0; m < 8; ++m) hi += static_cast<uint32_t>(s_t8[__funnelshift_r(r1, r2, 4 * m) & 0xFFFFu]) << (4 * m);
			const uint32_t nx = s_t8[r2 & 0xFFFFu];
			// even window 2m <- bit0 of group m ; odd window 2m+1 <- bit1 of group m & bit2 of group m+1
			const uint32_t elo = lo & 0x11111111u, ehi = hi & 0x11111111u;
			const uint32_t s2lo = (lo >> 2) & 0x11111111u, s2hi = (hi >> 2) & 0x11111111u;
			const uint32_t olo = ((lo >> 1) & 0x11111111u) & ((s2lo >> 4) | (s2hi << 28));
			const uint32_t ohi = ((hi >> 1) & 0x11111111u) & ((s2hi >> 4) | (((nx >> 2) & 1u) << 28));
			// candidates stay nibble-spaced: bit 4m+p of clo <-> window 2m+p, of chi <-> window 16+2m+p (ascending = scan order)
			clo = elo | (olo << 1);
			chi = ehi | (ohi << 1);
			if (m0 == 0xFFFFFFFFu) {
				clo = chi = 0;  // padding word
			} else {
				const uint32_t m1 = S.nmask[w + 1];
				if (m0 | m1) {  // N (or padding) within reach: a window dies if any match column holds an N (branch_and_bound-inl.h:388)
					const uint64_t mm = (static_cast<uint64_t>(m1) << 32) | m0;
					uint32_t inv = 0;
					for (int k = 0; k < M.L; ++k) inv |= static_cast<uint32_t>(mm >> M.coloff[k]);
					clo &= ~spread_nibble(inv & 0xFFFFu);
					chi &= ~spread_nibble(inv >> 16);
				}
			}
			// exact left-k-mer test for windows whose stage-1 bits are only projections
			if (inexact) {
#pragma unroll
				for (int half = 0; half < 2; ++half) {
					uint32_t rem = (half ? chi : clo) & inexact;
					while (rem) {
						const int b = __ffs(rem) - 1;
						rem &= rem - 1;
						const int j = 16 * half + 2 * (b >> 2) + (b & 1);
						uint32_t id;
						if (M.contiguous) id = kmer_at(c0, c1, j, M.len0);
						else { id = 0; for (int k = 0; k < M.len0; ++k) id |= base_at(c0, c1, j + M.coloff[k]) << (2 * k); }
						if (!((bm[id >> 5] >> (id & 31)) & 1u)) { if (half) chi &= ~(1u << b); else clo &= ~(1u << b); }
					}
				}
			}
		}
		const uint32_t cnt = __popc(clo) + __popc(chi);
		uint32_t incl = cnt;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
			if (lane >= d) incl += t;
		}
		const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
		uint32_t o = ocount + incl - cnt;
		while (clo) {
			const int b = __ffs(clo) - 1;
			clo &= clo - 1;
			if (o < O.warpCap) out[o] = w * 32u + static_cast<uint32_t>(2 * (b >> 2) + (b & 1));
			++o;
		}
		while (chi) {
			const int b = __ffs(chi) - 1;
			chi &= chi - 1;
			if (o < O.warpCap) out[o] = w * 32u + 16u + static_cast<uint32_t>(2 * (b >> 2) + (b & 1));
			++o;
		}
		ocount += total;
	}
	if (lane == 0) {
		O.warpCount[gw] = min(ocount, O.warpCap);
		if (ocount > O.warpCap) atomicMax(O.overflow, ocount);
	}
}

// resolve the found windows of every region into the final list (one CTA per region)
__global__ void __launch_bounds__(128) k_resolve_regions(SeqStream S, Model M, Values V, ScanOut3 O, const uint32_t* __restrict__ off, int wantScore,
                                                         uint32_t* __restrict__ oseq, uint32_t* __restrict__ opos, uint32_t* __restrict__ ovidx,
                                                         float* __restrict__ oscore, uint32_t cap) {
	const uint32_t r = blockIdx.x;
	const uint32_t n = O.warpCount[r];
	const uint32_t* src = O.gpos + static_cast<uint64_t>(r) * O.warpCap;
	const uint32_t dst = off[r];
	for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
		if (dst + i >= cap) break;
		const uint32_t g = src[i];
		const uint32_t w = g >> 5;
		const int j = static_cast<int>(g & 31u);
		const uint32_t v = eval_window(M, S.codes[w], S.codes[w + 1], S.nmask[w], S.nmask[w + 1], j);
		uint32_t sq, wo;
		if (S.uniformWps) { sq = w / S.uniformWps; wo = sq * S.uniformWps; }
		else { sq = find_seq(S, w); wo = S.wordOff[sq]; }
		oseq[dst + i] = sq + 1;
		opos[dst + i] = (w - wo) * 32 + static_cast<uint32_t>(j) + 1;
		ovidx[dst + i] = v;
		if (wantScore) oscore[dst + i] = (v == kNone) ? 1.0f : static_cast<float>(V.pS[v]);
	}
}

}  // namespace xxm

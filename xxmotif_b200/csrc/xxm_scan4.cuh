// K2, generic find kernel: used when the pair table of xxm_scan5.cuh cannot represent the left table exactly (match
// columns of the left part spanning more than 8 bases, i.e. long gapped motifs; tables that are not threshold-shaped).
//
// find   (k_scan4):  one thread per 32-base word.  16 LDS.U8 per word into the 64 KB shared-memory table T9, indexed by
//                    9 consecutive bases (18 bits; the byte is addressed by bases 0..7, base 8 selects one of its four
//                    2-bit entries).  An entry holds the EXACT answers for the two windows that start at base 0 and at
//                    base 1 of the group (left k-mer in the table) whenever all left match columns fit into the 9
//                    bases; otherwise the bits are projections and the word takes one exact bitmap test per flagged
//                    window.  The thread ORs its 16 x 2 bits into a 32-bit mask in window order (bit j = window starting at
//                    base j of the word is a site candidate) and stores it (4 B per 32 windows).  N handling, counting and
//                    record generation are k_expand5's.
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
__global__ void __launch_bounds__(kScan4Threads, 3) k_scan4(SeqStream S, Model M, uint32_t* __restrict__ hitmask, uint32_t scanWords, uint32_t wordsPerWarp) {
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
	const uint32_t wBeg = gw * wordsPerWarp;
	const uint32_t wEnd = min(scanWords, wBeg + wordsPerWarp);
	const uint32_t inexact = ((M.exactBits & 1u) ? 0u : 0x55555555u) | ((M.exactBits & 2u) ? 0u : 0xAAAAAAAAu);
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
	}
}

}  // namespace xxm

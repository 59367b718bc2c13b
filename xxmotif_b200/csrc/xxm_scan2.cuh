// K2 (second generation): the full-scan kernel.
//
// Budget: 0.375 B of sequence per window means ~60 windows per SM-clock at the HBM roofline, i.e. a few
// lane-instructions per window.  A table gather per window cannot get there (one LDS per window, 3-4-way bank
// conflicts on a random 16-bit index), so the scan is staged:
//   stage 1  one LDS.U8 per TWO windows into a 64 KB shared-memory byte table T8 indexed by 8 consecutive bases:
//            bit k (k = 0,1) says whether the window starting at base k of the group can still be a site judging from
//            the match columns visible inside the group (exact membership of the left k-mer when all of its columns
//            are visible, viability of its prefix otherwise);
//   stage 2  windows flagged by a prefix bit only: exact membership test of the left k-mer (bitmap in shared memory);
//   stage 3  the survivors (~2 % of the windows) are compacted across the warp and evaluated with all lanes busy:
//            rank gather, >8-column path, finalPval test, record write (coalesced).
// Every warp owns a contiguous range of the sequence stream and a private output region, so records come out in
// (sequence, position) order without any inter-CTA dependency; a small second kernel concatenates the regions.
#pragma once
#include "xxm_kernels.cuh"

namespace xxm {

constexpr int kScan2Threads = 512;
constexpr int kScan2Warps = kScan2Threads / 32;

struct T8Build {
	int nvis[2];            // visible left match columns for window offset k = 0,1
	int visoff[2][8];       // their base offsets inside the 8-base group
	uint32_t poff[9];       // word offset of the prefix-viability bitmap P_m inside `pbits`, m = 0..8
	const uint32_t* pbits;  // P_m[id & (4^m - 1)] = 1 iff some enumerated left k-mer starts with these m columns
};

__global__ void k_build_t8(T8Build B, uint8_t* __restrict__ t8) {
	const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= 65536u) return;
	uint32_t v = 0;
#pragma unroll
	for (int k = 0; k < 2; ++k) {
		const int m = B.nvis[k];
		uint32_t pid = 0;
		for (int c = 0; c < m; ++c) pid |= ((idx >> (2 * B.visoff[k][c])) & 3u) << (2 * c);
		const uint32_t bit = (m == 0) ? 1u : (B.pbits[B.poff[m] + (pid >> 5)] >> (pid & 31)) & 1u;
		v |= bit << k;
	}
	t8[idx] = static_cast<uint8_t>(v);
}

struct ScanOut2 {
	uint32_t* seq;       // staging records, warp-private regions of warpCap entries
	uint32_t* pos;
	uint32_t* vidx;
	uint32_t warpCap;
	uint32_t* warpCount; // [nwarps]
	uint32_t* overflow;  // set when a region was too small
};

__device__ __forceinline__ uint32_t left_id(const Model& M, uint64_t c0, uint64_t c1, int j) {
	if (M.contiguous) return kmer_at(c0, c1, j, M.len0);
	uint32_t id = 0;
	for (int k = 0; k < M.len0; ++k) id |= base_at(c0, c1, j + M.coloff[k]) << (2 * k);
	return id;
}

template <bool kSmemBitmap>
__global__ void __launch_bounds__(kScan2Threads, 2) k_scan2(SeqStream S, Model M, Values V, ScanOut2 O, uint32_t scanWords, uint32_t wordsPerWarp,
                                                         int checkFinal) {
	extern __shared__ __align__(16) uint8_t smem[];
	uint8_t* s_t8 = smem;                                                    // 64 KB
	uint16_t* s_q = reinterpret_cast<uint16_t*>(smem + 65536);               // kScan2Warps x 1024 queue entries
	uint32_t* s_bm = reinterpret_cast<uint32_t*>(smem + 65536 + kScan2Warps * 2048);
	{
		const uint4* src = reinterpret_cast<const uint4*>(M.t8);
		uint4* dst = reinterpret_cast<uint4*>(s_t8);
		for (uint32_t i = threadIdx.x; i < 65536u / 16u; i += blockDim.x) dst[i] = src[i];
	}
	const uint32_t* bm = M.hit0;
	if (kSmemBitmap) {
		const uint32_t nw = (1u << (2 * M.len0)) / 32u;
		for (uint32_t i = threadIdx.x; i < nw; i += blockDim.x) s_bm[i] = M.hit0[i];
		bm = s_bm;
	}
	__syncthreads();
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t gw = blockIdx.x * kScan2Warps + wid;  // global warp id = output region id
	uint16_t* q = s_q + wid * 1024;
	const uint32_t wBeg = gw * wordsPerWarp;
	const uint32_t wEnd = min(scanWords, wBeg + wordsPerWarp);
	const uint64_t obase = static_cast<uint64_t>(gw) * O.warpCap;
	uint32_t ocount = 0;
	const uint32_t exact = M.exactBits;
	for (uint32_t base = wBeg; base < wEnd; base += 32) {
		const uint32_t w = base + lane;
		uint32_t flagged = 0;
		uint64_t c0 = 0, c1 = 0;
		uint32_t m0 = 0xFFFFFFFFu, m1 = 0xFFFFFFFFu;
		if (w < wEnd) {
			c0 = S.codes[w]; c1 = S.codes[w + 1];
			m0 = S.nmask[w]; m1 = S.nmask[w + 1];
			const uint32_t r0 = static_cast<uint32_t>(c0), r1 = static_cast<uint32_t>(c0 >> 32), r2 = static_cast<uint32_t>(c1);
			// stage 1: 16 byte lookups, two windows each
#pragma unroll
			for (int m = 0; m < 16; ++m) {
				const int sh = 4 * m;
				const uint32_t x = (sh < 32 ? __funnelshift_r(r0, r1, sh) : __funnelshift_r(r1, r2, sh - 32)) & 0xFFFFu;
				flagged += static_cast<uint32_t>(s_t8[x]) << (2 * m);
			}
			flagged &= ~m0;
			// stage 2: windows flagged by a prefix bit -> exact membership of the left k-mer
			uint32_t rem = flagged & ~exact;
			while (rem) {
				const int j = __ffs(rem) - 1;
				rem &= rem - 1;
				const uint32_t id = left_id(M, c0, c1, j);
				if (!((bm[id >> 5] >> (id & 31)) & 1u)) flagged &= ~(1u << j);
			}
		}
		// stage 3: compact the survivors of the warp in (word, offset) order and evaluate them with all lanes
		const uint32_t cnt = __popc(flagged);
		uint32_t incl = cnt;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
			if (lane >= d) incl += t;
		}
		const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
		if (total == 0) continue;
		{
			uint32_t o = incl - cnt, rem = flagged;
			while (rem) {
				const int j = __ffs(rem) - 1;
				rem &= rem - 1;
				q[o++] = static_cast<uint16_t>((lane << 5) | j);
			}
		}
		__syncwarp();
		for (uint32_t qb = 0; qb < total; qb += 32) {
			const uint32_t qi = qb + lane;
			uint32_t v = kNone, seq = 0, pos = 0;
			if (qi < total) {
				const uint32_t e = q[qi];
				const uint32_t ww = base + (e >> 5);
				const int j = static_cast<int>(e & 31u);
				v = eval_window(M, S.codes[ww], S.codes[ww + 1], S.nmask[ww], S.nmask[ww + 1], j);
				if (v != kNone) {
					uint32_t sq, wo;
					if (S.uniformWps) { sq = ww / S.uniformWps; wo = sq * S.uniformWps; }
					else { sq = find_seq(S, ww); wo = S.wordOff[sq]; }
					seq = sq + 1;
					pos = (ww - wo) * 32 + j + 1;
					if (checkFinal && !(final_of(V, v, 0, static_cast<int>(pos)) < 1.0)) v = kNone;  // finalPval >= 1 -> continue
				}
			}
			const uint32_t ok = __ballot_sync(0xFFFFFFFFu, v != kNone);
			const uint32_t nOk = __popc(ok);
			if (v != kNone) {
				const uint32_t o = ocount + __popc(ok & ((1u << lane) - 1));
				if (o < O.warpCap) {
					O.seq[obase + o] = seq;
					O.pos[obase + o] = pos;
					O.vidx[obase + o] = v;
				}
			}
			ocount += nOk;
		}
		__syncwarp();
	}
	if (lane == 0) {
		O.warpCount[gw] = min(ocount, O.warpCap);
		if (ocount > O.warpCap) atomicExch(O.overflow, ocount);
	}
}

// exclusive prefix of the per-warp record counts (single CTA) + grand total
__global__ void __launch_bounds__(1024) k_region_offsets(const uint32_t* __restrict__ cnt, uint32_t n, uint32_t* __restrict__ off, uint64_t* __restrict__ total) {
	__shared__ uint32_t s_w[32];
	__shared__ uint32_t s_carry;
	if (threadIdx.x == 0) s_carry = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	for (uint32_t base = 0; base < n; base += 1024) {
		const uint32_t i = base + threadIdx.x;
		const uint32_t v = (i < n) ? cnt[i] : 0u;
		uint32_t incl = v;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
		if (lane == 31) s_w[wid] = incl;
		__syncthreads();
		uint32_t woff = 0, tot = 0;
		for (int k = 0; k < 32; ++k) { if (k < wid) woff += s_w[k]; tot += s_w[k]; }
		const uint32_t carry = s_carry;
		if (i < n) off[i] = carry + woff + incl - v;
		__syncthreads();
		if (threadIdx.x == 0) s_carry = carry + tot;
		__syncthreads();
	}
	if (threadIdx.x == 0) *total = s_carry;
}

// concatenate the warp regions into the final list and attach float(pValScore) (bestStartPos[].score, StartPosUpdater-inl.h:105)
__global__ void k_gather_regions(ScanOut2 O, const uint32_t* __restrict__ off, Values V, int wantScore, uint32_t* __restrict__ oseq, uint32_t* __restrict__ opos,
                                 uint32_t* __restrict__ ovidx, float* __restrict__ oscore, uint32_t cap) {
	const uint32_t r = blockIdx.x;
	const uint32_t n = O.warpCount[r];
	const uint64_t src = static_cast<uint64_t>(r) * O.warpCap;
	const uint32_t dst = off[r];
	for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
		if (dst + i >= cap) break;
		const uint32_t v = O.vidx[src + i];
		oseq[dst + i] = O.seq[src + i];
		opos[dst + i] = O.pos[src + i];
		ovidx[dst + i] = v;
		if (wantScore) oscore[dst + i] = static_cast<float>(V.pS[v]);
	}
}

}  // namespace xxm

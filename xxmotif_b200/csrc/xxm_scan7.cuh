// K2 (seventh generation): the expand kernel of the full scan, rebuilt around its instruction count.
//
// ncu on k_expand6 (profiles/r02_ncu_brief.txt): 77 % issue utilisation, DRAM at 45 % of peak -- the kernel is bound by the
// number of instructions it executes, ~560 per warp of 128 words on the bench workload, of which
//   ~70  bring the four mask words of a lane from the find kernel's even/odd order into window order,
//   ~190 are the lock-step extraction loop: four words per lane popped together, so a trip costs 32 instructions and the trip
//        count is the largest hit count of ANY of the warp's 128 words (~6 on clustered hits -- TATATA... -- although a word
//        holds 0.7-1 hits on average and half of the words none),
//   ~55  per 32 resolved hits, the rest offsets and loads.
// Here
//   * the find kernel writes its masks in window order (k_scan5<true>: 3 instructions per word there, xxm_scan5.cuh), so the
//     conversion is gone;
//   * the warp's offsets come from two REDUX instead of two shuffles plus a shuffle reduction;
//   * the k-mer of a hit is fetched as two 32-bit words at word index (entry >> 4) and one funnel shift by 2 * entry -- no
//     predicated second load, no select between the halves of the 64-bit code word;
//   * uniform and ragged sets are separate instantiations (no uniform branch and no reconvergence point in the resolve loop),
//     the record stores are predicated instead of branched around, and the loop is not unrolled (the unrolled form carried
//     three 64-bit store pointers along).
// Tried here and dropped after counting SASS instructions against the hit statistics of the bench workload (2-3 % of the
// windows, clustered: 0.6-1 hits per word, half of the words empty, ~6 in the fullest of 128): compacting the non-empty words of
// a warp and popping one word per lane (2 passes x 5.5 trips x 9 instructions + 44 to build the list + 30 per pass to set up)
// costs what the lock-step loop costs (6 trips x 32).  Measured and dropped: for ragged sets, per-word (sequence, position) origins in
// shared memory -- a lane resolves its first word once and walks on -- instead of one find_seq per hit: 0.333 ms per scan against
// 0.321 ms (profiles/r02c_ab_ragged.txt); the look-ups of neighbouring hits hit the same cache lines anyway.  Also measured and
// dropped: one lock-step trip, then the words with more than one hit compacted and finished one per lane (~60 instructions per
// warp fewer on paper, 150.2 us against 148.1 us on the clock: 8 KB more shared memory per CTA, two more warp barriers); 5 or 8
// resident CTAs per SM instead of 6 (148.4 / 152.3 us).
// Same inputs (final masks, chunk counts, tile offsets), same records, same order as k_expand6: globally ordered by
// (sequence, position), the reference's scan order (StartPosUpdater-inl.h:70-140).
#pragma once
#include "xxm_scan6.cuh"

namespace xxm {

constexpr int kExpand7Threads = 256;
constexpr int kExpand7Warps = kExpand7Threads / 32;
constexpr int kE7WarpWords = kTileWords / kExpand7Warps;  // 128: four words per lane
constexpr int kE7Queue = 512;                             // queue entries per warp and round

template <bool kFast, bool kUniform>
__global__ void __launch_bounds__(kExpand7Threads, 6) k_expand7(SeqStream S, Model M, Expand6IO io) {
	__shared__ __align__(16) uint64_t s_codesAll[kExpand7Warps][kE7WarpWords + 2];
	__shared__ uint16_t s_qAll[kExpand7Warps][kE7Queue];
	const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
	const uint32_t tile = blockIdx.x;
	uint64_t* s_codes = s_codesAll[wid];
	uint16_t* s_q = s_qAll[wid];
	const uint32_t wbase = tile * kTileWords + wid * kE7WarpWords;
	const uint32_t w0 = wbase + 4 * lane;
	// ---- every load of the warp is issued before anything waits (as in k_expand6)
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
	if (lane == 0 && wbase + kE7WarpWords < S.nwords + 2) ctail = __ldg(S.codes + wbase + kE7WarpWords);
	// ---- where this warp's records go: tile offset + the counts of the chunks before it inside the tile (chunks 2 wid, 2 wid + 1 are its own)
	const uint32_t mine = __reduce_add_sync(0xFFFFFFFFu, ((lane >> 1) == wid) ? cc : 0u);
	if (mine == 0) return;
	const uint32_t before = __reduce_add_sync(0xFFFFFFFFu, (lane < 2 * wid) ? cc : 0u);
	const uint64_t base = tileBase + before;
	*reinterpret_cast<uint4*>(s_codes + 4 * lane) = ca;
	*reinterpret_cast<uint4*>(s_codes + 4 * lane + 2) = cb;
	if (lane == 0) s_codes[kE7WarpWords] = ctail;
	const uint32_t p0 = __popc(h[0]), p1 = __popc(h[1]), p2 = __popc(h[2]), p3 = __popc(h[3]);
	const uint32_t cnt = p0 + p1 + p2 + p3;
	uint32_t incl = cnt;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
		if (lane >= static_cast<uint32_t>(d)) incl += t;
	}
	const uint32_t myOff = incl - cnt;
	// records [base, base + mine) of the three output arrays; room = how many of them fit the arrays
	const uint32_t room = (base < io.cap) ? static_cast<uint32_t>(min(static_cast<uint64_t>(io.cap) - base, static_cast<uint64_t>(mine))) : 0u;
	// the lane's own record slots: q = lane, lane + 32, ... (lane-dependent bases stay in vector registers, so that an address is
	// one IMAD.WIDE from the round offset)
	uint32_t* pseq = io.oseq + base + lane;
	uint32_t* ppos = io.opos + base + lane;
	uint32_t* pvid = io.ovidx + base + lane;
	const uint32_t roomL = (room > lane) ? room - lane : 0u;  // round offsets qr < roomL are inside the arrays
	const uint32_t* s_codes32 = reinterpret_cast<const uint32_t*>(s_codes);
	const uint32_t kmask = (1u << (2 * M.len0)) - 1u;
	const uint32_t* __restrict__ rank0 = M.rank0;
	const uint32_t wps = S.uniformWps, magic = io.wpsMagic;
	// one queue entry (word << 5 | window) -> record qr + lane of the warp
	auto resolve = [&](uint32_t ent, uint32_t qr) {
		const uint32_t wi = ent >> 5, j = ent & 31u;
		const uint32_t w = wbase + wi;
		uint32_t v;
		if (kFast) {
			// the 2 len0 <= 16 bits of the k-mer start at bit 2 ent of the staged code stream: two 32-bit words, one funnel shift
			const uint32_t lo = s_codes32[ent >> 4], hi = s_codes32[(ent >> 4) + 1];
			v = rank0[__funnelshift_r(lo, hi, 2u * ent) & kmask] - 1u;  // the window is in the table: the find kernel is exact
		} else {
			v = eval_window(M, s_codes[wi], s_codes[wi + 1], 0u, 0u, static_cast<int>(j));  // N windows are already gone
		}
		uint32_t seq1, pos;  // sequence and position, both 1-based
		if (kUniform) {
			uint32_t sq = __umulhi(w, magic);  // magic = floor(2^32 / wps): at most one too small
			uint32_t r = w - sq * wps;         // word inside the sequence
			if (r >= wps) { r -= wps; ++sq; }
			seq1 = sq + 1;
			pos = r * 32 + j + 1;
		} else {
			const uint32_t sq = find_seq(S, w);
			seq1 = sq + 1;
			pos = (w - S.wordOff[sq]) * 32 + j + 1;
		}
#ifdef XXM_SIMT_EMU
		if (qr < roomL) { pseq[qr] = seq1; ppos[qr] = pos; pvid[qr] = v; }
#else
		// three predicated stores, every address one multiply-add from its base (written in C the loop carries three 64-bit
		// pointers along and branches around the stores)
		asm volatile(
			"{\n"
			".reg .pred p;\n"
			".reg .u64 a, b, c;\n"
			"setp.lt.u32 p, %3, %4;\n"
			"mad.wide.u32 a, %3, 4, %0;\n"
			"mad.wide.u32 b, %3, 4, %1;\n"
			"mad.wide.u32 c, %3, 4, %2;\n"
			"@p st.global.u32 [a], %5;\n"
			"@p st.global.u32 [b], %6;\n"
			"@p st.global.u32 [c], %7;\n"
			"}" ::"l"(pseq), "l"(ppos), "l"(pvid), "r"(qr), "r"(roomL), "r"(seq1), "r"(pos), "r"(v)
			: "memory");
#endif
	};
	if (mine <= static_cast<uint32_t>(kE7Queue)) {
		// the four words of the lane in lock step, highest window first, every word filling its queue range backwards: the
		// trip count of the warp is the largest hit count of any of its 128 words
		uint32_t r0 = h[0], r1 = h[1], r2 = h[2], r3 = h[3];
		smem_addr_t q0 = static_cast<smem_addr_t>(__cvta_generic_to_shared(s_q + myOff)) + 2u * p0;  // one past the last slot of word 0
		smem_addr_t q1 = q0 + 2u * p1;
		smem_addr_t q2 = q1 + 2u * p2;
		smem_addr_t q3 = q2 + 2u * p3;
		const uint32_t tag = (4 * lane) << 5;
		while (r0 | r1 | r2 | r3) {
			pop_msb(r0, q0, tag);
			pop_msb(r1, q1, tag + 32);
			pop_msb(r2, q2, tag + 64);
			pop_msb(r3, q3, tag + 96);
		}
		__syncwarp();
#pragma unroll 1
		for (uint32_t qr = 0; qr + lane < mine; qr += 32) resolve(s_q[qr + lane], qr);
		return;
	}
	// dense warp (more than kE7Queue hits in 4096 windows): queue the hits in rounds
	for (uint32_t qbase = 0; qbase < mine; qbase += kE7Queue) {
		__syncwarp();
		uint32_t o = myOff;
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			uint32_t rem = h[k];
			while (rem) {
				const int j = __ffs(rem) - 1;
				rem &= rem - 1;
				if (o >= qbase && o < qbase + kE7Queue) s_q[o - qbase] = static_cast<uint16_t>(((4 * lane + k) << 5) | j);
				++o;
			}
		}
		__syncwarp();
		const uint32_t nq = min(static_cast<uint32_t>(kE7Queue), mine - qbase);
#pragma unroll 1
		for (uint32_t qr = 0; qr + lane < nq; qr += 32) resolve(s_q[qr + lane], qbase + qr);
	}
}

}  // namespace xxm

// K2 (fifth generation): the full scan as a pure streaming "find" kernel plus a single-pass "expand" kernel.
//
// find   (k_scan5):  one thread per 64 consecutive window starts (two 64-bit code words, one 128-bit load).  The test "is
//                    the left k-mer of this window in the p-value table" is the exact identity  a[x] + b[y] >= 128  over the
//                    two 4-mers x (bases 0..3) and y (bases 4..7) of the window (pair_table.h).  One LDS.32 fetches
//                    (a, b) for the two 4-mers starting at an even base and at the following odd base: the table is indexed
//                    by 5 bases (1024 entries) and replicated once per lane (entry q of lane l at word q * 32 + l, 128 KB),
//                    so every lookup is bank-conflict free.  Per pair of windows: 1 shift + 1 mask-or (address), 1 LDS,
//                    1 multiply-add (E[j] + (E[j+2] << 8): the two byte sums of the pair; bit 7 of a byte sum is the hit
//                    flag) and 1.5 instructions to gather the flags.  The kernel does not look at N masks and does not
//                    count; it writes one 32-bit mask per 32 windows in "even/odd" order: bit 16 + t = window 2t,
//                    bit t = window 2t + 1.
// expand (k_expand5): one CTA per tile of 1024 words, tile = block index (see the kernel on dispatch order).  Stages the tile's codes and N masks in
//                    shared memory (coalesced), brings the masks into window order, clears windows with an N in a match
//                    column (branch_and_bound-inl.h:388) and words outside the scanned range, counts, obtains the tile's
//                    global offset by a decoupled look-back over the tile status words (warp-parallel), and resolves the
//                    hits into records (value index, sequence, position, float score) in (sequence, position) order --
//                    the reference's scan order, so nothing downstream needs a sort by position.
#pragma once
#include "xxm_kernels.cuh"

namespace xxm {

constexpr int kScan5Threads = 1024;
constexpr int kPairTabWords = 1024 * 32;  // 128 KB of shared memory

template <int J>
__device__ __forceinline__ uint32_t pair_lookup(const uint32_t (&r)[5], const uint8_t* tab) {
	constexpr int bit = 4 * J, k = bit / 32, o = bit % 32;  // the 10-bit field of pair J starts at bit o of r[k]; wanted at bits [7, 17)
	uint32_t t;
	if constexpr (o == 0) t = r[k] << 7;
	else if constexpr (o == 4) t = r[k] << 3;
	else if constexpr (o + 10 <= 32) t = r[k] >> (o - 7);
	else t = __funnelshift_r(r[k], r[k + 1], o - 7);
	return *reinterpret_cast<const uint32_t*>(tab + (t & 0x1FF80u));  // tab already points at this lane's column
}

// prmt.b32 in its default mode: nibble n of the selector picks byte (n & 7) of {b, a}; bit 3 of the nibble replicates the sign of
// that byte instead.  (__byte_perm documents only the three low bits of a nibble, hence the PTX; XXM_SIMT_EMU: tests/simt.)
template <uint32_t kSel>
__device__ __forceinline__ uint32_t prmt_sx(uint32_t a, uint32_t b) {
#ifdef XXM_SIMT_EMU
	return emu::prmt(a, b, kSel);
#else
	uint32_t u;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(u) : "r"(a), "r"(b), "n"(kSel));
	return u;
#endif
}

template <int S>
__device__ __forceinline__ uint32_t gather_flags(uint32_t acc, uint32_t tLo, uint32_t tHi) {
	// bytes 3 / 1 of a T word hold the sums of its even / odd window; bit 7 of a sum is the hit flag.  PRMT with bit 3 of a
	// selector nibble set replicates the sign of the selected byte: U = [odd(lo), odd(hi), even(lo), even(hi)], every byte
	// 0x00 or 0xFF, and one and-or picks bit S of every byte
	const uint32_t u = prmt_sx<0xFBD9>(tLo, tHi);
	return acc | (u & (0x01010101u << S));
}

template <int G>
__device__ __forceinline__ uint32_t scan5_word(const uint32_t (&e)[34]) {
	uint32_t acc = 0;
#define XXM_T(j) (e[(j)] + (e[(j) + 2] << 8))
	acc = gather_flags<0>(acc, XXM_T(G + 0), XXM_T(G + 8));
	acc = gather_flags<1>(acc, XXM_T(G + 1), XXM_T(G + 9));
	acc = gather_flags<2>(acc, XXM_T(G + 2), XXM_T(G + 10));
	acc = gather_flags<3>(acc, XXM_T(G + 3), XXM_T(G + 11));
	acc = gather_flags<4>(acc, XXM_T(G + 4), XXM_T(G + 12));
	acc = gather_flags<5>(acc, XXM_T(G + 5), XXM_T(G + 13));
	acc = gather_flags<6>(acc, XXM_T(G + 6), XXM_T(G + 14));
	acc = gather_flags<7>(acc, XXM_T(G + 7), XXM_T(G + 15));
#undef XXM_T
	return acc;
}

// The same 32 flags in WINDOW order (bit n = window n of the word), for the seventh-generation expand kernel, which then needs
// no even/odd -> window conversion (13 instructions per word there against 3 here).  Windows 8B + s, B = 0..3, share bit s of
// their bytes; pair p (windows 2p, 2p + 1) sits in T word p.  For s' = 0..3 one PRMT gathers the four flags of pairs s', s' + 4
// into accA -- the even windows at bit 2s' of bytes 0 / 1 (their final place), the odd ones at bit 2s' + 1 of bytes 2 / 3 (two
// bytes too high) -- and one those of pairs s' + 8, s' + 12 into accB -- even windows in bytes 2 / 3 (final), odd ones in bytes
// 0 / 1 (two bytes too low); two byte permutes at the end put the odd windows in place.
template <int S>
__device__ __forceinline__ void gather_flags_nat(uint32_t& accA, uint32_t& accB, uint32_t t0, uint32_t t4, uint32_t t8, uint32_t t12) {
	accA |= prmt_sx<0xD9FB>(t0, t4) & ((0x00000101u << (2 * S)) | (0x01010000u << (2 * S + 1)));
	accB |= prmt_sx<0xFBD9>(t8, t12) & ((0x01010000u << (2 * S)) | (0x00000101u << (2 * S + 1)));
}

template <int G>
__device__ __forceinline__ uint32_t scan5_word_nat(const uint32_t (&e)[34]) {
	uint32_t accA = 0, accB = 0;
#define XXM_T(j) (e[(j)] + (e[(j) + 2] << 8))
	gather_flags_nat<0>(accA, accB, XXM_T(G + 0), XXM_T(G + 4), XXM_T(G + 8), XXM_T(G + 12));
	gather_flags_nat<1>(accA, accB, XXM_T(G + 1), XXM_T(G + 5), XXM_T(G + 9), XXM_T(G + 13));
	gather_flags_nat<2>(accA, accB, XXM_T(G + 2), XXM_T(G + 6), XXM_T(G + 10), XXM_T(G + 14));
	gather_flags_nat<3>(accA, accB, XXM_T(G + 3), XXM_T(G + 7), XXM_T(G + 11), XXM_T(G + 15));
#undef XXM_T
	return __byte_perm(accA, accB, 0x7610) | __byte_perm(accA, accB, 0x5432);
}

// dead: windows that cannot be a site whatever the PWM (k_dead_mask, xxm_scan6.cuh), cleared here so that the masks are
// final; chunkCount[c] = hits of the warp's 64 words, tileCount[c / 16] += the same (one atomic per chunk).
// kNatural: masks (and the dead-window masks read) in window order instead of even/odd order.
template <bool kNatural>
__global__ void __launch_bounds__(kScan5Threads, 1) k_scan5(const uint64_t* __restrict__ codes, const uint32_t* __restrict__ pairTab,
                                                           const uint32_t* __restrict__ dead, uint32_t* __restrict__ hitmask,
                                                           uint32_t* __restrict__ chunkCount, uint32_t* __restrict__ tileCount, uint32_t scanWords) {
	extern __shared__ __align__(16) uint8_t smem5[];
	{
		uint32_t* s = reinterpret_cast<uint32_t*>(smem5);
		for (uint32_t i = threadIdx.x; i < kPairTabWords; i += kScan5Threads) s[i] = pairTab[i >> 5];
	}
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const uint8_t* tab = smem5 + 4 * lane;
	const uint32_t warp = (blockIdx.x * kScan5Threads + threadIdx.x) >> 5;
	const uint32_t nwarps = (gridDim.x * kScan5Threads) >> 5;
	const uint32_t nchunks = (scanWords + 63) / 64;
	// the loads of a chunk are issued one chunk ahead (8 warps per scheduler do not hide a DRAM round trip otherwise); two
	// register sets alternate, so nothing is moved between them
	struct Chunk { uint4 v; uint2 d; uint32_t tail; };
	auto fetch = [&](Chunk& k, uint32_t c) {
		const uint32_t w = c * 64 + 2 * lane;
		k.v = make_uint4(0u, 0u, 0u, 0u);
		k.d = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
		k.tail = 0;
		if (c < nchunks && w < scanWords) {
			k.v = __ldg(reinterpret_cast<const uint4*>(codes + w));
			k.d = __ldg(reinterpret_cast<const uint2*>(dead + w));
			if (lane == 31 && w + 2 <= scanWords) k.tail = static_cast<uint32_t>(__ldg(codes + w + 2));
		}
	};
	auto process = [&](const Chunk& k, uint32_t c) {
		const uint32_t w = c * 64 + 2 * lane;
		uint32_t r[5] = {k.v.x, k.v.y, k.v.z, k.v.w, 0};
		r[4] = __shfl_down_sync(0xFFFFFFFFu, r[0], 1);
		if (lane == 31) r[4] = k.tail;
		uint32_t e[34];
		e[0] = pair_lookup<0>(r, tab);   e[1] = pair_lookup<1>(r, tab);   e[2] = pair_lookup<2>(r, tab);   e[3] = pair_lookup<3>(r, tab);
		e[4] = pair_lookup<4>(r, tab);   e[5] = pair_lookup<5>(r, tab);   e[6] = pair_lookup<6>(r, tab);   e[7] = pair_lookup<7>(r, tab);
		e[8] = pair_lookup<8>(r, tab);   e[9] = pair_lookup<9>(r, tab);   e[10] = pair_lookup<10>(r, tab); e[11] = pair_lookup<11>(r, tab);
		e[12] = pair_lookup<12>(r, tab); e[13] = pair_lookup<13>(r, tab); e[14] = pair_lookup<14>(r, tab); e[15] = pair_lookup<15>(r, tab);
		e[16] = pair_lookup<16>(r, tab); e[17] = pair_lookup<17>(r, tab); e[18] = pair_lookup<18>(r, tab); e[19] = pair_lookup<19>(r, tab);
		e[20] = pair_lookup<20>(r, tab); e[21] = pair_lookup<21>(r, tab); e[22] = pair_lookup<22>(r, tab); e[23] = pair_lookup<23>(r, tab);
		e[24] = pair_lookup<24>(r, tab); e[25] = pair_lookup<25>(r, tab); e[26] = pair_lookup<26>(r, tab); e[27] = pair_lookup<27>(r, tab);
		e[28] = pair_lookup<28>(r, tab); e[29] = pair_lookup<29>(r, tab); e[30] = pair_lookup<30>(r, tab); e[31] = pair_lookup<31>(r, tab);
		e[32] = pair_lookup<32>(r, tab); e[33] = pair_lookup<33>(r, tab);
		uint32_t h0 = kNatural ? scan5_word_nat<0>(e) : scan5_word<0>(e), h1 = kNatural ? scan5_word_nat<16>(e) : scan5_word<16>(e);
		uint32_t cnt = 0;
		if (w < scanWords) {
			h0 &= ~k.d.x;
			h1 = (w + 1 < scanWords) ? (h1 & ~k.d.y) : 0u;
			*reinterpret_cast<uint2*>(hitmask + w) = make_uint2(h0, h1);
			cnt = __popc(h0) + __popc(h1);
		}
		cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
		if (lane == 0) {
			chunkCount[c] = cnt;
			if (cnt) atomicAdd(tileCount + (c >> 4), cnt);
		}
	};
	Chunk A, B;
	fetch(A, warp);
	for (uint32_t c = warp; c < nchunks; c += 2 * nwarps) {
		fetch(B, c + nwarps);
		process(A, c);
		if (c + nwarps >= nchunks) break;
		fetch(A, c + 2 * nwarps);
		process(B, c + nwarps);
	}
}

// even/odd mask (bit 16 + t = window 2t, bit t = window 2t + 1) -> window order: a perfect shuffle of the two halves.  One
// PRMT swaps the halves and the two middle bytes (the first butterfly stage), three xor-swap stages finish it.
__device__ __forceinline__ uint32_t eo_to_natural(uint32_t m) {
	uint32_t x = __byte_perm(m, 0u, 0x1302), t;
	t = (x ^ (x >> 4)) & 0x00F000F0u; x = x ^ t ^ (t << 4);
	t = (x ^ (x >> 2)) & 0x0C0C0C0Cu; x = x ^ t ^ (t << 2);
	t = (x ^ (x >> 1)) & 0x22222222u; x = x ^ t ^ (t << 1);
	return x;
}

// One step of the hit extraction for one word, branch free: if r != 0, clear its highest set bit j, move the shared-memory
// cursor back by one 16-bit slot and store tagk + j there; if r == 0 nothing happens (bfind gives 0xFFFFFFFF, the shift
// clamps to 0, the store and the cursor update are predicated off).
#ifdef XXM_SIMT_EMU
using smem_addr_t = uintptr_t;  // the emulator's "shared memory" lives in the host address space
__device__ __forceinline__ void pop_msb(uint32_t& r, smem_addr_t& cursor, uint32_t tagk) {
	if (r == 0) return;
	const int j = 31 - __builtin_clz(r);
	r ^= 1u << j;
	cursor -= 2;
	*reinterpret_cast<uint16_t*>(cursor) = static_cast<uint16_t>(tagk + static_cast<uint32_t>(j));
}
#else
using smem_addr_t = uint32_t;
__device__ __forceinline__ void pop_msb(uint32_t& r, smem_addr_t& cursor, uint32_t tagk) {
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		".reg .u32 j, b, v;\n"
		".reg .u16 h;\n"
		"setp.ne.u32 p, %0, 0;\n"
		"bfind.u32 j, %0;\n"
		"shl.b32 b, 1, j;\n"
		"xor.b32 %0, %0, b;\n"
		"add.u32 v, %2, j;\n"
		"cvt.u16.u32 h, v;\n"
		"@p sub.u32 %1, %1, 2;\n"
		"@p st.shared.u16 [%1], h;\n"
		"}"
		: "+r"(r), "+r"(cursor)
		: "r"(tagk)
		: "memory");
}
#endif

constexpr int kExpand5Threads = 256;
constexpr int kExpand5Queue = 4096;

struct Expand5IO {
	const uint32_t* hitmask;
	uint32_t scanWords;
	int eoFormat;       // 1: masks in even/odd order (k_scan5), 0: window order (k_scan4)
	int wantScore;
	uint32_t *oseq, *opos, *ovidx;
	float* oscore;
	uint32_t cap;
	uint64_t* status;   // [ntiles] zeroed: per-tile count
	uint64_t* gstatus;  // [ntiles / 32 + 1] zeroed: per-group sum / inclusive prefix
	uint32_t* ticket;   // zeroed
	uint64_t* total;    // out
	uint32_t ntiles;
	uint32_t wpsMagic;  // floor(2^32 / uniformWps), 0 for ragged sets
};

struct HitRec { uint32_t seq, pos, v; };

// kFast: one table, no gap columns -- the value index is rank0[k-mer] - 1 and the k-mer is one funnel shift away
template <bool kFast>
__device__ __forceinline__ HitRec expand_hit(const SeqStream& S, const Model& M, const Expand5IO& io, const uint64_t* s_codes, uint32_t wbase, uint32_t e) {
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
	if (io.wpsMagic) {
		sq = __umulhi(w, io.wpsMagic);
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

// kTicket (XXM_EXPAND5_TICKET, off by default): the tile is the block's ticket from an atomic counter instead of its index, so a tile
// only ever waits for tiles whose blocks have already started -- no assumption about the order in which blocks become resident.
// Verified under the SIMT emulator with the blocks run in REVERSE order (tests/simt/scan4_emu.cpp); it has not run on a GPU yet,
// which is why the instantiation that has -- kTicket = false, machine code unchanged -- stays the default.
template <bool kFast, bool kTicket = false>
__global__ void __launch_bounds__(kExpand5Threads) k_expand5(SeqStream S, Model M, Values V, Expand5IO io) {
	__shared__ __align__(16) uint64_t s_codes[kTileWords + 4];
	__shared__ uint16_t s_q[kExpand5Queue];
	__shared__ uint32_t s_warp[kExpand5Threads / 32];
	__shared__ uint64_t s_base;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	// tile = block index: like CUB's decoupled look-back WITHOUT its ticket this relies on blocks being dispatched in index order (a
	// tile only ever waits for tiles with smaller indices, which then are resident or done)
	uint32_t tileOf = blockIdx.x;
	if constexpr (kTicket) {
		__shared__ uint32_t s_ticket;
		if (tid == 0) s_ticket = atomicAdd(io.ticket, 1u);
		__syncthreads();
		tileOf = s_ticket;
	}
	const uint32_t tile = tileOf;
	const uint32_t wbase = tile * kTileWords;
	const uint32_t w0 = wbase + 4 * tid;
	// ---- all loads of the thread up front: its four hit masks, five N masks, four code words (staged for the expansion)
	uint32_t h[4], nm[5];
	if (w0 + 4 < S.nwords + 2 && w0 + 3 < io.scanWords) {
		const uint4 hv = *reinterpret_cast<const uint4*>(io.hitmask + w0);
		const uint4 nv = *reinterpret_cast<const uint4*>(S.nmask + w0);
		nm[4] = S.nmask[w0 + 4];
		const uint4 ca = *reinterpret_cast<const uint4*>(S.codes + w0);
		const uint4 cb = *reinterpret_cast<const uint4*>(S.codes + w0 + 2);
		h[0] = hv.x; h[1] = hv.y; h[2] = hv.z; h[3] = hv.w;
		nm[0] = nv.x; nm[1] = nv.y; nm[2] = nv.z; nm[3] = nv.w;
		*reinterpret_cast<uint4*>(s_codes + 4 * tid) = ca;
		*reinterpret_cast<uint4*>(s_codes + 4 * tid + 2) = cb;
	} else {
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			const uint32_t w = w0 + k;
			const bool in = w < S.nwords + 2;
			h[k] = (w < io.scanWords) ? io.hitmask[w] : 0u;
			nm[k] = in ? S.nmask[w] : 0xFFFFFFFFu;
			s_codes[4 * tid + k] = in ? S.codes[w] : 0ull;
		}
		nm[4] = (w0 + 4 < S.nwords + 2) ? S.nmask[w0 + 4] : 0xFFFFFFFFu;
	}
	if (tid == 0) s_codes[kTileWords] = (wbase + kTileWords < S.nwords + 2) ? S.codes[wbase + kTileWords] : 0ull;
	// ---- window order, N windows out (a window dies if any match column holds an N, branch_and_bound-inl.h:388)
	if (io.eoFormat) {
#pragma unroll
		for (int k = 0; k < 4; ++k) h[k] = eo_to_natural(h[k]);
	}
	if (__any_sync(0xFFFFFFFFu, (nm[0] | nm[1] | nm[2] | nm[3] | nm[4]) != 0u)) {  // a sequence end or an N within reach of this warp
		if (M.contiguous) {  // OR of (mask >> 0) .. (mask >> (L - 1)) by doubling, all four words at once
			uint32_t x[5] = {nm[0], nm[1], nm[2], nm[3], nm[4]};
			int have = 1;
			while (have < M.L) {
				const int sh = min(have, M.L - have);
#pragma unroll
				for (int k = 0; k < 4; ++k) x[k] |= __funnelshift_r(x[k], x[k + 1], sh);
				x[4] |= x[4] >> sh;
				have += sh;
			}
#pragma unroll
			for (int k = 0; k < 4; ++k) h[k] &= ~x[k];
		} else {
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				const uint64_t mm = (static_cast<uint64_t>(nm[k + 1]) << 32) | nm[k];
				uint32_t inv = 0;
				for (int c = 0; c < M.L; ++c) inv |= static_cast<uint32_t>(mm >> M.coloff[c]);
				h[k] &= ~inv;
			}
		}
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
	for (int k = 0; k < kExpand5Threads / 32; ++k) { if (k < wid) woff += s_warp[k]; total += s_warp[k]; }
	if (tid == 0) reinterpret_cast<volatile uint64_t*>(io.status)[tile] = kFlagAgg | total;  // as early as possible
	const uint32_t myOff = woff + incl - cnt;
	const bool fits = total <= static_cast<uint32_t>(kExpand5Queue);
	if (fits) {  // the common case: the whole tile fits the queue
		// the four words of the thread in lock step, highest window first, every word filling its queue range backwards: the
		// trip count of a warp is the largest hit count of any of its 128 words
		uint32_t r0 = h[0], r1 = h[1], r2 = h[2], r3 = h[3];
		smem_addr_t q0 = static_cast<smem_addr_t>(__cvta_generic_to_shared(s_q + myOff)) + 2u * __popc(r0);  // one past the last slot of word 0
		smem_addr_t q1 = q0 + 2u * __popc(r1);
		smem_addr_t q2 = q1 + 2u * __popc(r2);
		smem_addr_t q3 = q2 + 2u * __popc(r3);
		const uint32_t tag = static_cast<uint32_t>(4 * tid) << 5;
		while (r0 | r1 | r2 | r3) {
			pop_msb(r0, q0, tag);
			pop_msb(r1, q1, tag + 32);
			pop_msb(r2, q2, tag + 64);
			pop_msb(r3, q3, tag + 96);
		}
	}
	__syncthreads();
	HitRec first{0, 0, 0};
	if (wid != 0 && fits && static_cast<uint32_t>(tid) < total) first = expand_hit<kFast>(S, M, io, s_codes, wbase, s_q[tid]);
	if (wid == 0) {
		// Two-level decoupled look-back.  A plain look-back advances its frontier by at most 32 tiles per L2 round trip,
		// which is slower than the tiles can be expanded.  Tiles are therefore grouped by 32: a tile sums the counts of
		// the tiles before it in its group (they hold earlier tickets, so they are running and will publish), the last
		// tile of a group publishes the group sum, and the running prefix is carried group to group (32 groups =
		// 1024 tiles per round trip).  Warps 1..7 expand their first hit meanwhile.
		volatile uint64_t* st = io.status;
		volatile uint64_t* gst = io.gstatus;
		const uint32_t g = tile >> 5, r = tile & 31u;
		const bool lastOfGroup = (r == 31u) || (tile == io.ntiles - 1);
		uint64_t v = 0;
		if (static_cast<uint32_t>(lane) < r) {
			do { v = st[(g << 5) + lane]; } while ((v >> 62) == 0);
			v &= kValMask;
		}
#pragma unroll
		for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
		const uint64_t local = v;  // hits of the tiles before this one in its group
		if (lastOfGroup && lane == 0 && g != 0) gst[g] = kFlagAgg | (local + total);
		uint64_t gprefix = 0;
		int64_t t = static_cast<int64_t>(g) - 1 - lane;
		while (g != 0) {
			uint64_t x = 2ull << 62;  // before the first group: inclusive prefix 0
			if (t >= 0) x = gst[t];
			while (__any_sync(0xFFFFFFFFu, (x >> 62) == 0)) {
				if ((x >> 62) == 0) x = gst[t];
			}
			const uint32_t inclMask = __ballot_sync(0xFFFFFFFFu, (x >> 62) == 2);
			const int firstIncl = inclMask ? (__ffs(inclMask) - 1) : 32;
			uint64_t c = (lane <= firstIncl) ? (x & kValMask) : 0ull;
#pragma unroll
			for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
			gprefix += c;
			if (inclMask) break;
			t -= 32;
		}
		if (lastOfGroup && lane == 0) gst[g] = kFlagIncl | (gprefix + local + total);
		if (lane == 0) {
			s_base = gprefix + local;
			if (tile == io.ntiles - 1) *io.total = gprefix + local + total;
		}
		if (fits && static_cast<uint32_t>(tid) < total) first = expand_hit<kFast>(S, M, io, s_codes, wbase, s_q[tid]);
	}
	__syncthreads();
	if (total == 0) return;
	const uint64_t base = s_base;
	auto emit = [&](const HitRec& rec, uint64_t o) {
		if (o < io.cap) {
			io.oseq[o] = rec.seq;
			io.opos[o] = rec.pos;
			io.ovidx[o] = rec.v;
			if (io.wantScore) io.oscore[o] = (rec.v == kNone) ? 1.0f : static_cast<float>(V.pS[rec.v]);
		}
	};
	if (fits) {
		if (static_cast<uint32_t>(tid) < total) emit(first, base + tid);
		for (uint32_t q = tid + kExpand5Threads; q < total; q += kExpand5Threads) emit(expand_hit<kFast>(S, M, io, s_codes, wbase, s_q[q]), base + q);
		return;
	}
	// dense tile: queue the hits in rounds
	for (uint32_t qbase = 0; qbase < total; qbase += kExpand5Queue) {
		__syncthreads();
		uint32_t o = myOff;
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			uint32_t rem = h[k];
			while (rem) {
				const int j = __ffs(rem) - 1;
				rem &= rem - 1;
				if (o >= qbase && o < qbase + kExpand5Queue) s_q[o - qbase] = static_cast<uint16_t>(((4 * tid + k) << 5) | j);
				++o;
			}
		}
		__syncthreads();
		const uint32_t nq = min(static_cast<uint32_t>(kExpand5Queue), total - qbase);
		for (uint32_t q = tid; q < nq; q += kExpand5Threads) emit(expand_hit<kFast>(S, M, io, s_codes, wbase, s_q[q]), base + qbase + q);
	}
}

}  // namespace xxm

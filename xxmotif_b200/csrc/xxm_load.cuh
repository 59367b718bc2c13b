// a1 on the device: the load-time kernels of xxm_load_seqset (xxm_gpu.cu) -- the per-sequence residue statistics of the device
// packing path, the layout of the stream from the raw lengths (two-launch block scan), and the coarse word -> sequence index.
// (In a header of their own so that tests/simt can run them from source; the library includes this file where the kernels used to
// stand.)
#pragma once
#include "xxm_device.cuh"

namespace xxm {
// per sequence: number of residues that are not N (code 0); any code > 4 raises the flag.  One warp per sequence of [s0, s1).
__global__ void k_seq_stats(const uint8_t* __restrict__ raw, const uint64_t* __restrict__ rawOff, uint32_t s0, uint32_t s1, int32_t* __restrict__ notX,
                            uint32_t* __restrict__ bad) {
	const int lane = threadIdx.x & 31;
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t s = s0 + warp; s < s1; s += nwarps) {
		const uint64_t a = rawOff[s], b = rawOff[s + 1];
		uint32_t cnt = 0, err = 0;
		// 16 bytes per lane once the pointer is aligned
		uint64_t i = a;
		const uint64_t aUp = min(b, (a + 15) & ~static_cast<uint64_t>(15));
		for (uint64_t k = i + lane; k < aUp; k += 32) { const uint8_t c = raw[k]; cnt += (c != 0); err |= (c > 4); }
		i = aUp;
		const uint64_t nv = (b - i) / 16;
		for (uint64_t v = lane; v < nv; v += 32) {
			const uint4 q = *reinterpret_cast<const uint4*>(raw + i + 16 * v);
			const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
			for (int t = 0; t < 4; ++t) {
				const uint32_t x = w4[t];
				// bytes != 0 (for codes 0..7): bit 0 | bit 1 | bit 2; codes > 4: bit 2 and (bit 0 or bit 1), or any higher bit (then the count is moot)
				const uint32_t nz = (x | (x >> 1) | (x >> 2)) & 0x01010101u;
				cnt += __popc(nz);
				err |= (x & 0xF8F8F8F8u) | ((x >> 2) & (x | (x >> 1)) & 0x01010101u);
			}
		}
		for (uint64_t k = i + 16 * nv + lane; k < b; k += 32) { const uint8_t c = raw[k]; cnt += (c != 0); err |= (c > 4); }
		for (int d = 16; d; d >>= 1) { cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, d); err |= __shfl_xor_sync(0xFFFFFFFFu, err, d); }
		if (lane == 0) { notX[s] = static_cast<int32_t>(cnt); if (err) atomicExch(bad, s + 1); }
	}
}

// layout of the stream from the raw lengths alone: rawOff (bytes), wordOff (words), len (stored length).  Two launches over
// blocks of 1024 sequences: per-block sums, then every block adds up the sums of the blocks before it and scans its own.
__device__ __forceinline__ void seq_extent(const int32_t* __restrict__ n, uint32_t i, uint32_t nent, int revcomp, uint64_t* raw, uint64_t* words, uint64_t* stored) {
	const uint64_t n0 = (i < nent) ? static_cast<uint64_t>(n[i]) : 0ull;
	const uint64_t ns = revcomp ? 2 * n0 + 1 : n0;
	*raw = n0;
	*stored = ns;
	*words = (i < nent) ? (ns + kPadBases + 31) / 32 : 0ull;
}
__global__ void __launch_bounds__(1024) k_seq_layout_sums(const int32_t* __restrict__ n, uint32_t nent, int revcomp, uint64_t* __restrict__ sums) {
	__shared__ uint64_t s_r[32], s_w[32];
	const uint32_t i = blockIdx.x * 1024u + threadIdx.x, lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
	uint64_t r, w, st;
	seq_extent(n, i, nent, revcomp, &r, &w, &st);
#pragma unroll
	for (int d = 16; d; d >>= 1) { r += __shfl_xor_sync(0xFFFFFFFFu, r, d); w += __shfl_xor_sync(0xFFFFFFFFu, w, d); }
	if (lane == 0) { s_r[wid] = r; s_w[wid] = w; }
	__syncthreads();
	if (threadIdx.x == 0) {
		uint64_t ar = 0, aw = 0;
		for (int k = 0; k < 32; ++k) { ar += s_r[k]; aw += s_w[k]; }
		sums[2 * blockIdx.x] = ar;
		sums[2 * blockIdx.x + 1] = aw;
	}
}
__global__ void __launch_bounds__(1024) k_seq_layout(const int32_t* __restrict__ n, uint32_t nent, int revcomp, const uint64_t* __restrict__ sums,
                                                     uint64_t* __restrict__ rawOff, uint32_t* __restrict__ wordOff, int32_t* __restrict__ len) {
	__shared__ uint64_t s_r[32], s_w[32], s_pr[32], s_pw[32];
	const uint32_t i = blockIdx.x * 1024u + threadIdx.x, lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
	uint64_t pr = 0, pw = 0;  // everything before this block
	for (uint32_t b = threadIdx.x; b < blockIdx.x; b += 1024u) { pr += sums[2 * b]; pw += sums[2 * b + 1]; }
#pragma unroll
	for (int d = 16; d; d >>= 1) { pr += __shfl_xor_sync(0xFFFFFFFFu, pr, d); pw += __shfl_xor_sync(0xFFFFFFFFu, pw, d); }
	uint64_t r, w, st;
	seq_extent(n, i, nent, revcomp, &r, &w, &st);
	uint64_t ir = r, iw = w;  // inclusive scans inside the warp
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint64_t a = __shfl_up_sync(0xFFFFFFFFu, ir, d), b2 = __shfl_up_sync(0xFFFFFFFFu, iw, d);
		if (lane >= static_cast<uint32_t>(d)) { ir += a; iw += b2; }
	}
	if (lane == 0) { s_pr[wid] = pr; s_pw[wid] = pw; }
	if (lane == 31) { s_r[wid] = ir; s_w[wid] = iw; }
	__syncthreads();
	uint64_t br = 0, bw = 0;
#pragma unroll
	for (int k = 0; k < 32; ++k) {
		br += s_pr[k]; bw += s_pw[k];
		if (static_cast<uint32_t>(k) < wid) { br += s_r[k]; bw += s_w[k]; }
	}
	const uint64_t er = br + ir - r, ew = bw + iw - w;  // exclusive
	if (i < nent) {
		rawOff[i] = er;
		wordOff[i] = static_cast<uint32_t>(ew);
		len[i] = static_cast<int32_t>(st);
	}
	if (i + 1 == nent) { rawOff[nent] = er + r; wordOff[nent] = static_cast<uint32_t>(ew + w); }
}

// coarse word -> sequence index: blk2seq[b] = the sequence that owns word b * kBlkWords
__global__ void k_blk2seq(const uint32_t* __restrict__ wordOff, uint32_t nent, uint32_t nblk, uint32_t* __restrict__ blk2seq) {
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= nent) return;
	const uint32_t w0 = wordOff[s], w1 = wordOff[s + 1];
	for (uint32_t b = (w0 + kBlkWords - 1) / kBlkWords; b < nblk && (b * kBlkWords < w1 || s + 1 == nent); ++b) blk2seq[b] = s;
}
}  // namespace xxm

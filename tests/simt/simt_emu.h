// TEST INFRASTRUCTURE ONLY -- a minimal SIMT emulator that runs the product's CUDA kernel SOURCES (xxmotif_b200/csrc/*.cuh) on the
// host, so that `-m "not gpu"` tests can hold the kernels themselves -- not a restatement of them -- to a brute-force expectation
// in a container that has no GPU.  Nothing under xxmotif_b200/ includes, links or loads this; libxxmotif_gpu.so has no CPU path.
//
// Model: one OS thread; every CUDA thread of the running block is a ucontext fiber; blocks of a grid run one after another (so
// `static` stands in for `__shared__`; kernels that spin on other blocks cannot be emulated).  A fiber runs until it reaches a
// warp collective or a block barrier, deposits its operand and yields; the last participant to arrive releases the others.  A
// full pass over the fibers without progress is reported as a deadlock.  Divergence, partial masks and early exits of whole
// warps behave as on the device; atomics are plain read-modify-writes (one OS thread).
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define XXM_SIMT_EMU 1

// ---- race detection (build.py sanitize="thread"): every CUDA thread is a ThreadSanitizer fiber.  Switching fibers creates NO
// happens-before edge; the edges are the ones the programming model gives -- __syncthreads among the threads of a block, __syncwarp
// (and, unless XXM_EMU_STRICT_SHFL is defined, the other *_sync collectives) among the lanes named in the mask, launch order on the
// stream, atomics -- so two accesses of one block to the same shared or global address with a write among them and no such edge
// between them are reported with both source lines: what `compute-sanitizer --tool racecheck` reports, without a GPU.  Blocks of a
// grid run one after another on the same `static` arrays that stand in for __shared__, so consecutive blocks are ordered
// (races BETWEEN blocks are out of scope, as they are for racecheck).  The emulator's own bookkeeping is not instrumented.
#if defined(__SANITIZE_THREAD__)
#define XXM_EMU_TSAN 1
extern "C" {
void* __tsan_get_current_fiber(void);
void* __tsan_create_fiber(unsigned flags);
void __tsan_destroy_fiber(void* fiber);
void __tsan_switch_to_fiber(void* fiber, unsigned flags);
void __tsan_acquire(void* addr);
void __tsan_release(void* addr);
}
#define EMU_NOTSAN __attribute__((no_sanitize("thread")))
#define EMU_TSAN_SWITCH(f) __tsan_switch_to_fiber((f), 1u /* no synchronisation */)
#define EMU_TSAN_ACQUIRE(p) __tsan_acquire(static_cast<void*>(p))
#define EMU_TSAN_RELEASE(p) __tsan_release(static_cast<void*>(p))
#else
#define EMU_NOTSAN
#define EMU_TSAN_SWITCH(f) ((void) 0)
#define EMU_TSAN_ACQUIRE(p) ((void) 0)
#define EMU_TSAN_RELEASE(p) ((void) 0)
#endif
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __noinline__
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
	unsigned x, y, z;
	dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

inline uint3 threadIdx{0, 0, 0}, blockIdx{0, 0, 0};
inline dim3 blockDim, gridDim;

namespace emu {

constexpr size_t kStackBytes = 256 * 1024;

struct Fiber {
	ucontext_t ctx;
	void* stack = nullptr;
	void* tsan = nullptr;  // ThreadSanitizer's view of this CUDA thread
	bool done = false;
};

struct WarpState {
	uint64_t val[32];
	uint64_t snap[32];
	uint32_t arrived = 0;
	uint32_t readers = 0;
	bool ready = false;
	char hb = 0;  // carries the happens-before edge of this warp's collectives
};

struct BlockState {
	std::vector<Fiber> fib;
	std::vector<WarpState> warp;
	ucontext_t sched;
	void* schedTsan = nullptr;
	// addresses that carry the happens-before edges of __syncthreads / of block and launch order.  A block's threads acquire
	// chain[b & 1] when they start and release into chain[(b + 1) & 1] when they end: the next block (and the host) sees them, the
	// threads of the SAME block that start later do not
	char barObj = 0, chain[2] = {0, 0};
	unsigned block = 0;
	int cur = 0;
	int alive = 0;
	int barCount = 0;
	uint64_t barGen = 0;
	uint64_t progress = 0;
	std::function<void()> body;
	// statistics of the running launch
	uint64_t collectives = 0, barriers = 0;
};

inline BlockState* B = nullptr;
inline unsigned blockY = 0;  // blockIdx.y of the next launches (the grid itself is one-dimensional: a harness launches once per y)
inline bool reverseBlocks = false;  // run the blocks of the next launches from the last to the first (a kernel must not care)

EMU_NOTSAN inline void yield() {
	EMU_TSAN_SWITCH(B->schedTsan);
	swapcontext(&B->fib[B->cur].ctx, &B->sched);
}

EMU_NOTSAN inline void release_barrier_if_complete() {
	if (B->barCount > 0 && B->barCount == B->alive) {
		B->barCount = 0;
		++B->barGen;
		++B->progress;
		++B->barriers;
	}
}

EMU_NOTSAN inline void fiber_entry() {
	EMU_TSAN_ACQUIRE(&B->chain[B->block & 1u]);  // after the host code before the launch and after the previous blocks
	BlockState* const b = B;
	Fiber* const self = &b->fib[b->cur];
	b->body();
	self->done = true;
	--b->alive;
	++b->progress;
	release_barrier_if_complete();  // exited threads do not take part in later barriers
	EMU_TSAN_RELEASE(&b->chain[(b->block + 1u) & 1u]);  // (nothing below touches instrumented memory)
	EMU_TSAN_SWITCH(b->schedTsan);
	swapcontext(&self->ctx, &b->sched);
}

EMU_NOTSAN inline void syncthreads() {
	const uint64_t g = B->barGen;
	EMU_TSAN_RELEASE(&B->barObj);
	++B->barCount;
	release_barrier_if_complete();
	while (B->barGen == g) yield();
	EMU_TSAN_ACQUIRE(&B->barObj);
}

// every lane of `mask` deposits v; returns once all have, with the 32 deposited values in out[]
EMU_NOTSAN inline void exchange(uint32_t mask, uint64_t v, uint64_t* out, bool orders = true) {
	const int t = B->cur, lane = t & 31;
	WarpState& W = B->warp[t >> 5];
	if (orders) EMU_TSAN_RELEASE(&W.hb);
	if (!((mask >> lane) & 1u)) { std::fprintf(stderr, "simt_emu: lane %d calls a collective whose mask %08x does not name it\n", lane, mask); std::abort(); }
	while (W.ready) yield();  // the previous collective of this warp is still being read
	W.val[lane] = v;
	W.arrived |= 1u << lane;
	if (W.arrived == mask) {
		for (int l = 0; l < 32; ++l) W.snap[l] = W.val[l];
		W.ready = true;
		W.readers = mask;
		W.arrived = 0;
		++B->progress;
		++B->collectives;
	} else if (W.arrived & ~mask) {
		std::fprintf(stderr, "simt_emu: lanes of one warp disagree on a collective's mask (%08x vs arrived %08x)\n", mask, W.arrived);
		std::abort();
	}
	while (!W.ready || !((W.readers >> lane) & 1u)) yield();
	for (int l = 0; l < 32; ++l) out[l] = W.snap[l];
	W.readers &= ~(1u << lane);
	if (W.readers == 0) { W.ready = false; ++B->progress; }
	if (orders) EMU_TSAN_ACQUIRE(&W.hb);
}
// __syncwarp orders memory among the named lanes; the CUDA programming guide does not promise that of the other *_sync collectives
#ifdef XXM_EMU_STRICT_SHFL
constexpr bool kShflOrders = false;
#else
constexpr bool kShflOrders = true;
#endif

struct LaunchStats { uint64_t collectives, barriers; };
inline LaunchStats lastLaunch{0, 0};

// runs body() once per CUDA thread of a grid x block launch (1-D)
template <class F>
EMU_NOTSAN void launch(unsigned grid, unsigned block, F&& body) {
	BlockState st;
	B = &st;
	gridDim = dim3(grid);
	blockDim = dim3(block);
	st.fib.resize(block);
	st.warp.resize((block + 31) / 32);
	for (auto& f : st.fib) f.stack = std::malloc(kStackBytes);
#ifdef XXM_EMU_TSAN
	st.schedTsan = __tsan_get_current_fiber();
	for (auto& f : st.fib) f.tsan = __tsan_create_fiber(0);
#endif
	st.body = body;
	EMU_TSAN_RELEASE(&st.chain[0]);  // what the host wrote before the launch is visible to every thread of the grid
	EMU_TSAN_RELEASE(&st.chain[1]);
	for (unsigned bi = 0; bi < grid; ++bi) {
		const unsigned b = reverseBlocks ? grid - 1 - bi : bi;
		blockIdx = uint3{b, blockY, 0};
		st.block = bi;  // (execution order: what the happens-before chain between consecutive blocks follows)
		st.alive = static_cast<int>(block);
		st.barCount = 0;
		for (auto& w : st.warp) { w.arrived = 0; w.readers = 0; w.ready = false; }
		for (unsigned t = 0; t < block; ++t) {
			Fiber& f = st.fib[t];
			f.done = false;
			getcontext(&f.ctx);
			f.ctx.uc_stack.ss_sp = f.stack;
			f.ctx.uc_stack.ss_size = kStackBytes;
			f.ctx.uc_link = &st.sched;
			makecontext(&f.ctx, reinterpret_cast<void (*)()>(fiber_entry), 0);
		}
		while (st.alive > 0) {
			const uint64_t before = st.progress;
			for (unsigned t = 0; t < block; ++t) {
				if (st.fib[t].done) continue;
				st.cur = static_cast<int>(t);
				threadIdx = uint3{t, 0, 0};
				EMU_TSAN_SWITCH(st.fib[t].tsan);
				swapcontext(&st.sched, &st.fib[t].ctx);
			}
			if (st.progress == before && st.alive > 0) {
				std::fprintf(stderr, "simt_emu: deadlock in block %u (%d threads alive, %d at the barrier)\n", b, st.alive, st.barCount);
				std::abort();
			}
		}
	}
	EMU_TSAN_ACQUIRE(&st.chain[0]);  // ... and what the grid wrote is visible to the host code after the launch
	EMU_TSAN_ACQUIRE(&st.chain[1]);
	lastLaunch = LaunchStats{st.collectives, st.barriers};
#ifdef XXM_EMU_TSAN
	for (auto& f : st.fib) __tsan_destroy_fiber(f.tsan);
#endif
	for (auto& f : st.fib) std::free(f.stack);
	B = nullptr;
}

template <class T>
inline uint64_t to_bits(T v) {
	static_assert(sizeof(T) <= 8, "collective operand wider than 64 bits");
	uint64_t u = 0;
	std::memcpy(&u, &v, sizeof(T));
	return u;
}
template <class T>
inline T from_bits(uint64_t u) {
	T v;
	std::memcpy(&v, &u, sizeof(T));
	return v;
}
inline int lane_id() { return B->cur & 31; }

// prmt.b32 in its default mode: selector nibble n picks byte (n & 7) of {b, a}; bit 3 of the nibble replicates that byte's sign
inline uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
	const uint64_t src = (static_cast<uint64_t>(b) << 32) | a;
	uint32_t r = 0;
	for (int i = 0; i < 4; ++i) {
		const uint32_t n = (sel >> (4 * i)) & 0xFu;
		uint32_t byte = static_cast<uint32_t>(src >> (8 * (n & 7u))) & 0xFFu;
		if (n & 8u) byte = (byte & 0x80u) ? 0xFFu : 0x00u;
		r |= byte << (8 * i);
	}
	return r;
}

}  // namespace emu

// ---- barriers and warp collectives
inline void __syncthreads() { emu::syncthreads(); }
inline void __syncwarp(unsigned mask = 0xFFFFFFFFu) { uint64_t o[32]; emu::exchange(mask, 0, o, true); }
template <class T>
inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
	uint64_t o[32];
	emu::exchange(mask, emu::to_bits(v), o, emu::kShflOrders);
	const int lane = emu::lane_id();
	const int base = lane & ~(width - 1);
	return emu::from_bits<T>(o[base + (src & (width - 1))]);
}
template <class T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) {
	uint64_t o[32];
	emu::exchange(mask, emu::to_bits(v), o, emu::kShflOrders);
	const int lane = emu::lane_id();
	const int base = lane & ~(width - 1);
	const int src = lane - static_cast<int>(d);
	return emu::from_bits<T>(o[src < base ? lane : src]);
}
template <class T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) {
	uint64_t o[32];
	emu::exchange(mask, emu::to_bits(v), o, emu::kShflOrders);
	const int lane = emu::lane_id();
	const int base = lane & ~(width - 1);
	const int src = lane + static_cast<int>(d);
	return emu::from_bits<T>(o[src >= base + width ? lane : src]);
}
template <class T>
inline T __shfl_xor_sync(unsigned mask, T v, int d, int width = 32) {
	(void) width;
	uint64_t o[32];
	emu::exchange(mask, emu::to_bits(v), o, emu::kShflOrders);
	return emu::from_bits<T>(o[emu::lane_id() ^ d]);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
	uint64_t o[32];
	emu::exchange(mask, pred ? 1u : 0u, o, emu::kShflOrders);
	unsigned r = 0;
	for (int l = 0; l < 32; ++l)
		if (((mask >> l) & 1u) && o[l]) r |= 1u << l;
	return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
	uint64_t o[32];
	emu::exchange(mask, v, o, emu::kShflOrders);
	unsigned r = 0;
	for (int l = 0; l < 32; ++l)
		if ((mask >> l) & 1u) r += static_cast<unsigned>(o[l]);
	return r;
}
inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
	uint64_t o[32];
	emu::exchange(mask, v, o, emu::kShflOrders);
	unsigned r = 0;
	for (int l = 0; l < 32; ++l)
		if (((mask >> l) & 1u) && static_cast<unsigned>(o[l]) > r) r = static_cast<unsigned>(o[l]);
	return r;
}
inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
	uint64_t o[32];
	emu::exchange(mask, v, o, emu::kShflOrders);
	unsigned r = 0;
	for (int l = 0; l < 32; ++l)
		if ((mask >> l) & 1u) r |= static_cast<unsigned>(o[l]);
	return r;
}
template <class T>
inline unsigned __match_any_sync(unsigned mask, T v) {
	uint64_t o[32];
	emu::exchange(mask, emu::to_bits(v), o, emu::kShflOrders);
	unsigned r = 0;
	const uint64_t mine = emu::to_bits(v);
	for (int l = 0; l < 32; ++l)
		if (((mask >> l) & 1u) && o[l] == mine) r |= 1u << l;
	return r;
}

// ---- scalar intrinsics
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
inline int __clz(int x) { return x ? __builtin_clz(static_cast<unsigned>(x)) : 32; }
inline unsigned __brev(unsigned x) {
	unsigned r = 0;
	for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i);
	return r;
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { return static_cast<unsigned>((((static_cast<uint64_t>(hi) << 32) | lo) >> (s & 31u))); }
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { return static_cast<unsigned>((((static_cast<uint64_t>(hi) << 32) | lo) << (s & 31u)) >> 32); }
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) { return emu::prmt(a, b, sel & 0x7777u); }
inline unsigned __umulhi(unsigned a, unsigned b) { return static_cast<unsigned>((static_cast<uint64_t>(a) * b) >> 32); }
template <class T>
inline T __ldg(const T* p) { return *p; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dsqrt_rn(double a) { return std::sqrt(a); }
inline int __double2int_rz(double a) { return static_cast<int>(a); }
inline long long __double_as_longlong(double a) { return emu::from_bits<long long>(emu::to_bits(a)); }
inline double __longlong_as_double(long long a) { return emu::from_bits<double>(emu::to_bits(a)); }
inline unsigned __float_as_uint(float f) { return emu::to_bits(f) & 0xFFFFFFFFu; }
inline float __uint_as_float(unsigned u) { return emu::from_bits<float>(u); }
inline size_t __cvta_generic_to_shared(const void* p) { return reinterpret_cast<size_t>(p); }

// ---- atomics (one OS thread: plain read-modify-write; under ThreadSanitizer real relaxed atomics, so that they do not count as races)
#ifdef XXM_EMU_TSAN
template <class T, class U>
inline T atomicAdd(T* p, U v) { return __atomic_fetch_add(p, static_cast<T>(v), __ATOMIC_RELAXED); }
template <class T, class U>
inline T atomicOr(T* p, U v) { return __atomic_fetch_or(p, static_cast<T>(v), __ATOMIC_RELAXED); }
template <class T, class U>
inline T atomicMax(T* p, U v) {
	T o = __atomic_load_n(p, __ATOMIC_RELAXED);
	while (static_cast<T>(v) > o && !__atomic_compare_exchange_n(p, &o, static_cast<T>(v), false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
	return o;
}
template <class T, class U>
inline T atomicMin(T* p, U v) {
	T o = __atomic_load_n(p, __ATOMIC_RELAXED);
	while (static_cast<T>(v) < o && !__atomic_compare_exchange_n(p, &o, static_cast<T>(v), false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
	return o;
}
template <class T, class U>
inline T atomicExch(T* p, U v) { return __atomic_exchange_n(p, static_cast<T>(v), __ATOMIC_RELAXED); }
template <class T, class U, class V>
inline T atomicCAS(T* p, U cmp, V v) {
	T o = static_cast<T>(cmp);
	__atomic_compare_exchange_n(p, &o, static_cast<T>(v), false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
	return o;
}
#else
template <class T, class U>
inline T atomicAdd(T* p, U v) { const T o = *p; *p = static_cast<T>(o + static_cast<T>(v)); return o; }
template <class T, class U>
inline T atomicOr(T* p, U v) { const T o = *p; *p = static_cast<T>(o | static_cast<T>(v)); return o; }
template <class T, class U>
inline T atomicMax(T* p, U v) { const T o = *p; if (static_cast<T>(v) > o) *p = static_cast<T>(v); return o; }
template <class T, class U>
inline T atomicMin(T* p, U v) { const T o = *p; if (static_cast<T>(v) < o) *p = static_cast<T>(v); return o; }
template <class T, class U>
inline T atomicExch(T* p, U v) { const T o = *p; *p = static_cast<T>(v); return o; }
template <class T, class U, class V>
inline T atomicCAS(T* p, U cmp, V v) { const T o = *p; if (o == static_cast<T>(cmp)) *p = static_cast<T>(v); return o; }
#endif

// ---- device libm stand-ins where the host's has side effects the device's does not: glibc's lgamma() writes the global `signgam`
// (every thread calling it would "race" on that); CUDA's has no such global.  The kernels live in namespace xxm, so an unqualified
// lgamma() there finds this one first.
namespace xxm {
inline double lgamma(double x) {
	int sign;
	return ::lgamma_r(x, &sign);
}
}  // namespace xxm

// ---- CUDA's global min / max overloads
template <class T>
inline T min(T a, T b) { return b < a ? b : a; }
template <class T>
inline T max(T a, T b) { return a < b ? b : a; }
inline unsigned min(unsigned a, int b) { return min(a, static_cast<unsigned>(b)); }
inline unsigned min(int a, unsigned b) { return min(static_cast<unsigned>(a), b); }
inline unsigned max(unsigned a, int b) { return max(a, static_cast<unsigned>(b)); }
inline unsigned max(int a, unsigned b) { return max(static_cast<unsigned>(a), b); }

// TEST INFRASTRUCTURE ONLY -- controls for the emulator's race detection (simt_emu.h built with -fsanitize=thread): kernels with a
// known shared-memory hazard must be reported, their repaired twins must not.  usage: race_control <case>
#include <cstdio>
#include <cstring>
#include <vector>

#include "simt_emu.h"

static int s_one;         // stands in for a __shared__ scalar
static int s_arr[256];    // ... and a __shared__ array

// lane 0 publishes a value, the other lanes of the warp read it
static void k_warp_publish(bool withSync, int* out) {
	if ((threadIdx.x & 31) == 0) s_one = 40 + static_cast<int>(threadIdx.x >> 5);
	if (withSync) __syncwarp();
	out[threadIdx.x] = s_one;
}
// every thread writes its slot, then reads its neighbour's from another warp
static void k_block_neighbour(bool withSync, int* out) {
	s_arr[threadIdx.x] = static_cast<int>(threadIdx.x) * 3;
	if (withSync) __syncthreads();
	out[threadIdx.x] = s_arr[(threadIdx.x + 32) % blockDim.x];
}
// a shuffle is not a memory fence for the strict model, but it is for the default one
static void k_shfl_then_read(int* out) {
	if ((threadIdx.x & 31) == 0) s_one = 7;
	const int v = __shfl_sync(0xFFFFFFFFu, static_cast<int>(threadIdx.x), 0);
	out[threadIdx.x] = s_one + v;
}
// all lanes add to one counter: atomics are not races
static void k_atomic_counter(unsigned* counter) { atomicAdd(counter, 1u); }

int main(int argc, char** argv) {
	const char* which = argc > 1 ? argv[1] : "";
	std::vector<int> out(256, -1);
	unsigned counter = 0;
	if (!std::strcmp(which, "warp_racy")) emu::launch(1, 32, [&] { k_warp_publish(false, out.data()); });
	else if (!std::strcmp(which, "warp_synced")) emu::launch(1, 32, [&] { k_warp_publish(true, out.data()); });
	else if (!std::strcmp(which, "block_racy")) emu::launch(1, 128, [&] { k_block_neighbour(false, out.data()); });
	else if (!std::strcmp(which, "block_synced")) emu::launch(2, 128, [&] { k_block_neighbour(true, out.data()); });
	else if (!std::strcmp(which, "shfl")) emu::launch(1, 32, [&] { k_shfl_then_read(out.data()); });
	else if (!std::strcmp(which, "atomics")) emu::launch(3, 64, [&] { k_atomic_counter(&counter); });
	else { std::fprintf(stderr, "unknown case\n"); return 2; }
	std::printf("%s done: out[1] = %d, counter = %u\n", which, out[1], counter);
	return 0;
}

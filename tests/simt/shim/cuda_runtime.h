// TEST INFRASTRUCTURE ONLY: stands in for <cuda_runtime.h> when the kernel sources are compiled for the SIMT emulator.
#pragma once
#include "../simt_emu.h"

/* TEST INFRASTRUCTURE — stands in for <cuda_runtime.h> when the product's CUDA sources are compiled
 * with g++ for the CPU emulation of the device (tests/emu/cuda_emu.h). */
#pragma once
#include "../cuda_emu.h"

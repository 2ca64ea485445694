// TEST-ONLY stand-in for the CUDA header of the same name (tests/emu/cuda_shim.h)
#include "../cuda_shim.h"

// Stand-in for <cuda_runtime.h> when the library's sources are compiled for the CPU emulator (tests/emulate; TEST
// INFRASTRUCTURE ONLY). Found first on the include path of tests/emulate/emu_build.py, never by nvcc.
#pragma once
#include "../cuda_emu.h"

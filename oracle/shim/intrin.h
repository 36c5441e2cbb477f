/* oracle shim (Linux): Utils/MemoryHelpers.cpp includes the MSVC header <intrin.h> */
#pragma once
#include <immintrin.h>

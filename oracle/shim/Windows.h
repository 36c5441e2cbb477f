/* oracle shim (Linux): the reference includes <Windows.h> unconditionally in Utils/Profiler.h and
 * Utils/Memory.cpp; only the QueryPerformance* timer API is needed by headers used on Linux. */
#pragma once
#include <time.h>
#include <stdint.h>
typedef union _LARGE_INTEGER { long long QuadPart; } LARGE_INTEGER;
static inline int QueryPerformanceCounter(LARGE_INTEGER* t)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    t->QuadPart = (long long)ts.tv_sec * 1000000000LL + ts.tv_nsec;
    return 1;
}
static inline int QueryPerformanceFrequency(LARGE_INTEGER* t) { t->QuadPart = 1000000000LL; return 1; }

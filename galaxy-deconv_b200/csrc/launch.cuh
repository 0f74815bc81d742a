// Launch bookkeeping shared by every .cu of libgdeconv: a process-wide launch counter (gd_launch_count(), the
// `gpu_launches` claim of bench.py) and the post-launch error check.
#pragma once
#include <atomic>
#include "gd_common.cuh"

namespace gd {
extern std::atomic<unsigned long long> g_launches;
}

#define GD_LAUNCHED()                                                     \
    do {                                                                  \
        gd::g_launches.fetch_add(1, std::memory_order_relaxed);           \
        GD_CUDA_CHECK(cudaGetLastError());                                \
    } while (0)

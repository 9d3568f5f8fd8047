// multigpu.h -- internal: NUMA placement helpers of multigpu.cu
#pragma once
#include <cstddef>

#include "splatapult_b200/splatcomm.h"

namespace splat
{
using Comm = ::SplatComm;
int numa_node_of_device(int device);                            // -1: unknown / single node
bool bind_thread_near(int device, int* nodeOut, int* cpusOut);  // sched_setaffinity to the GPU's node
void* host_alloc_near(size_t bytes, int device, int* nodeOut);  // pinned, pages on the GPU's node when that can be arranged
void host_free_any(void* p);                                    // frees blocks of host_alloc_near
}  // namespace splat

// render_launch.h — host-visible launch interface of the two kernel builds.
#pragma once
#include <cuda_runtime.h>
#include "device_layout.h"

struct RtLaunchCfg {
    int block_threads;  // 128 (default) or 64
    int min_ctas;       // __launch_bounds__ second argument: caps registers/thread
    bool work_counters; // RT_AOV_WORK build (counts inner visits and triangle tests)
    bool speculative;   // fast build: speculative traversal (postponed leaves)
    int wide;           // fast build: 0 = the reference's 2-wide tree, 1 = the 4-wide tree made from it, 2 = compressed 8-wide (both imply speculative)
    int grid;           // number of persistent CTAs
    bool park = false;  // fast build, 4-wide tree, no work counters: the instance that keeps the path state in shared memory (8 CTAs/SM)
};

cudaError_t rt_launch_fast(const RtDeviceScene& sc, const RtFrameArgs& fa, const RtLaunchCfg& cfg, cudaStream_t st);
cudaError_t rt_launch_strict(const RtDeviceScene& sc, const RtFrameArgs& fa, const RtLaunchCfg& cfg, cudaStream_t st);
// resident CTAs per SM and registers/thread of the instantiation cfg selects
cudaError_t rt_occupancy_fast(const RtLaunchCfg& cfg, int* ctas_per_sm, int* regs);
cudaError_t rt_occupancy_strict(const RtLaunchCfg& cfg, int* ctas_per_sm, int* regs);
// the cooperative drain kernel (fast build): finishes the paths queued in fa.drain_queue; one resident wave
cudaError_t rt_launch_drain(const RtDeviceScene& sc, const RtFrameArgs& fa, bool work_counters, int sm_count, cudaStream_t st);

// render_fast.cu — RT_MODE_FAST build of render_kernel.cuh (FMA contraction on, MUFU reciprocals).
#define RT_STRICT 0
#define RT_KERNEL_NS rt_fast
#include "render_kernel.cuh"
#include "render_launch.inl"

cudaError_t rt_launch_fast(const RtDeviceScene& sc, const RtFrameArgs& fa, const RtLaunchCfg& cfg, cudaStream_t st)
{
    return rt_fast::launch(sc, fa, cfg, st);
}
cudaError_t rt_occupancy_fast(const RtLaunchCfg& cfg, int* ctas_per_sm, int* regs)
{
    return rt_fast::occupancy(cfg, ctas_per_sm, regs);
}

cudaError_t rt_launch_drain(const RtDeviceScene& sc, const RtFrameArgs& fa, bool work_counters, int sm_count, cudaStream_t st)
{
    static int occ[2] = {0, 0};
    const int w = work_counters ? 1 : 0;
    if (!occ[w]) {
        cudaError_t e = work_counters ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[w], rt_fast::drain_kernel<true>, 128, 0)
                                      : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[w], rt_fast::drain_kernel<false>, 128, 0);
        if (e != cudaSuccess) return e;
        if (occ[w] < 1) occ[w] = 1;
    }
    const int grid = sm_count * occ[w];
    if (work_counters) rt_fast::drain_kernel<true><<<grid, 128, 0, st>>>(sc, fa);
    else rt_fast::drain_kernel<false><<<grid, 128, 0, st>>>(sc, fa);
    return cudaGetLastError();
}

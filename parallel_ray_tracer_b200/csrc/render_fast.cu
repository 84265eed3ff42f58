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

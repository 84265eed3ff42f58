// render_strict.cu — RT_MODE_STRICT build of render_kernel.cuh.  Compiled with -fmad=false
// -prec-div=true -prec-sqrt=true -ftz=false (csrc/Makefile) so that every FP32 operation is the
// IEEE operation the strict oracle (oracle/rt_oracle.c) performs, in the same order.
#define RT_STRICT 1
#define RT_KERNEL_NS rt_strict
#include "render_kernel.cuh"
#include "render_launch.inl"

cudaError_t rt_launch_strict(const RtDeviceScene& sc, const RtFrameArgs& fa, const RtLaunchCfg& cfg, cudaStream_t st)
{
    return rt_strict::launch(sc, fa, cfg, st);
}
cudaError_t rt_occupancy_strict(const RtLaunchCfg& cfg, int* ctas_per_sm, int* regs)
{
    return rt_strict::occupancy(cfg, ctas_per_sm, regs);
}

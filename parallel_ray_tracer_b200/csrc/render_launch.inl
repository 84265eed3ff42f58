// render_launch.inl — instantiation table; included inside render_{fast,strict}.cu.
#include "render_launch.h"

namespace RT_KERNEL_NS {

typedef void (*kernel_fn)(const RtDeviceScene, const RtFrameArgs);

template <bool WORK, bool SPEC, int WIDE>
static kernel_fn pick(int block, int minb)
{
    if (block == 64) return render_kernel<64, 12, WORK, SPEC, WIDE>;
#if RT_OPT_PARK
    if (minb >= 10) return render_kernel<128, 10, WORK, SPEC, WIDE>;
    if (minb >= 9) return render_kernel<128, 9, WORK, SPEC, WIDE>;
#endif
    if (minb >= 8) return render_kernel<128, 8, WORK, SPEC, WIDE>;
    if (minb >= 7) return render_kernel<128, 7, WORK, SPEC, WIDE>;
    if (minb >= 6) return render_kernel<128, 6, WORK, SPEC, WIDE>;
    if (minb >= 5) return render_kernel<128, 5, WORK, SPEC, WIDE>;
    return render_kernel<128, 4, WORK, SPEC, WIDE>;
}

static kernel_fn pick(const RtLaunchCfg& c)
{
#if RT_STRICT
    // the strict build never speculates and never uses the 4-wide tree: its visit order is the reference's
    return c.work_counters ? pick<true, false, 0>(c.block_threads, c.min_ctas) : pick<false, false, 0>(c.block_threads, c.min_ctas);
#else
    if (c.wide == 2) return c.work_counters ? pick<true, true, 2>(c.block_threads, c.min_ctas) : pick<false, true, 2>(c.block_threads, c.min_ctas);
    if (c.wide == 1 && c.park && !c.work_counters && c.block_threads != 64) {
        if (c.min_ctas >= 8) return render_kernel<128, 8, false, true, 1, true>;
        if (c.min_ctas >= 7) return render_kernel<128, 7, false, true, 1, true>;
        return render_kernel<128, 6, false, true, 1, true>;
    }
    if (c.wide == 1) return c.work_counters ? pick<true, true, 1>(c.block_threads, c.min_ctas) : pick<false, true, 1>(c.block_threads, c.min_ctas);
    if (c.speculative) return c.work_counters ? pick<true, true, 0>(c.block_threads, c.min_ctas) : pick<false, true, 0>(c.block_threads, c.min_ctas);
    return c.work_counters ? pick<true, false, 0>(c.block_threads, c.min_ctas) : pick<false, false, 0>(c.block_threads, c.min_ctas);
#endif
}

static cudaError_t launch(const RtDeviceScene& sc, const RtFrameArgs& fa, const RtLaunchCfg& cfg, cudaStream_t st)
{
    kernel_fn k = pick(cfg);
    const int block = cfg.block_threads == 64 ? 64 : 128;
    k<<<cfg.grid, block, 0, st>>>(sc, fa);
    return cudaGetLastError();
}

static cudaError_t occupancy(const RtLaunchCfg& cfg, int* ctas_per_sm, int* regs)
{
    kernel_fn k = pick(cfg);
    const int block = cfg.block_threads == 64 ? 64 : 128;
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, k);
    if (e != cudaSuccess) return e;
    if (regs) *regs = a.numRegs;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, k, block, 0);
}

} // namespace RT_KERNEL_NS

/*
 * rt_sampling.h — counter-based sub-pixel sample sequence shared by the CUDA
 * kernel, the CPU oracle restatement and the reference harness.
 *
 * The reference renderer has no anti-aliasing: it shoots exactly one ray through
 * the top-left CORNER of every pixel (reference cpu/src/main.c:228-234).  The
 * spp > 1 mode is new surface.  The convention fixed here:
 *
 *   - sample 0 always has offset (0,0), so spp == 1 is the reference's image;
 *   - samples s >= 1 get offsets (jx, jy) uniform in [0,1)^2 from a stateless
 *     hash of (x, y, s, seed) — no RNG state, any thread can evaluate any sample;
 *   - the pixel colour is the FP32 sum of the un-clamped sample colours in
 *     sample order, divided by (float)spp, then clamped to [0,1].
 *
 * Plain C, usable from C, C++ and CUDA (define RT_HD before including for
 * __host__ __device__ qualifiers).
 */
#ifndef RT_SAMPLING_H
#define RT_SAMPLING_H

#include <stdint.h>

#ifndef RT_HD
#define RT_HD
#endif

static inline RT_HD uint32_t rt_hash4(uint32_t x, uint32_t y, uint32_t s, uint32_t seed)
{
    uint32_t v = (x * 0x9E3779B1u) ^ (y * 0x85EBCA77u) ^ (s * 0xC2B2AE3Du) ^ (seed * 0x27D4EB2Fu);
    v ^= v >> 16; v *= 0x7FEB352Du;
    v ^= v >> 15; v *= 0x846CA68Bu;
    v ^= v >> 16;
    return v;
}

/* offsets in [0,1) with 24 random bits each; exact in float */
static inline RT_HD void rt_sample_offset(uint32_t x, uint32_t y, uint32_t s, uint32_t seed,
                                          float* jx, float* jy)
{
    if (s == 0u) { *jx = 0.0f; *jy = 0.0f; return; }
    uint32_t hx = rt_hash4(x, y, s, seed);
    uint32_t hy = rt_hash4(x, y, s, seed ^ 0x68BC21EBu);
    *jx = (float)(hx >> 8) * (1.0f / 16777216.0f);
    *jy = (float)(hy >> 8) * (1.0f / 16777216.0f);
}

#endif /* RT_SAMPLING_H */

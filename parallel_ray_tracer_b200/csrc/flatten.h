// flatten.h — host staging buffers of the HBM layout (device_layout.h), produced by flatten.cpp.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "rt_b200.h"

// host mirrors of the device_layout.h constants (that header needs CUDA vector types)
#define RT_REF_NONE_HOST ((int32_t)0x80000000)
#define RT_LEAF_CNT_ESC_HOST 15
#define RT_STACK_ENTRIES_HOST 40

namespace rt {

// std::allocator whose value-less construct() default-initialises: resize() of the multi-GB staging arrays does not
// zero-fill them on one thread first (every element is written by the parallel passes of flatten.cpp).
template <class T>
struct NoInitAlloc : std::allocator<T> {
    template <class U> struct rebind { using other = NoInitAlloc<U>; };
    NoInitAlloc() = default;
    template <class U> NoInitAlloc(const NoInitAlloc<U>&) {}
    template <class U> void construct(U* p) { ::new (static_cast<void*>(p)) U; }
    template <class U, class A0, class... A> void construct(U* p, A0&& a0, A&&... a) { ::new (static_cast<void*>(p)) U(std::forward<A0>(a0), std::forward<A>(a)...); }
};
typedef std::vector<float, NoInitAlloc<float>> FloatBuf;

struct FlatScene {
    FloatBuf             nodes;    // 16 floats (4 x float4) per inner node
    FloatBuf             nodes4;   // 32 floats (8 x float4) per 4-wide node (fast build)
    std::vector<uint32_t, NoInitAlloc<uint32_t>> nodes8; // 24 words per compressed 8-wide node (fast build; wide8.h)
    int depth8 = 0;                // levels of the 8-wide tree
    FloatBuf             tris;     // 16 floats (4 x float4) per leaf-order slot
    FloatBuf             shade;    // 4 floats per original triangle
    std::vector<float>   mats;     // 12 floats per material
    std::vector<float>   lights;   // 8 floats per light
    std::vector<int32_t> leaf_cnt; // empty unless some leaf holds >= 15 triangles
    uint32_t n_lights = 0;
    float ambient[3] = {0, 0, 0};
    int max_depth = 0;
    int stack_need4 = 0;           // stack slots a ray can need on the 4-wide tree

    size_t bytes() const
    {
        return 4 * (nodes.size() + nodes4.size() + nodes8.size() + tris.size() + shade.size() + mats.size() + lights.size() + leaf_cnt.size());
    }
};

// One axis of a 4-wide node's child box as centre and half extent (render_kernel.cuh: box_key4).  Plain IEEE
// round-to-nearest operations only (this file is compiled without contraction), so flatten_gpu.cu's device version gives
// the same bits: centre = mn / 2 + mx / 2, half = the larger one-sided distance, one ulp up unless it is exactly 0 —
// a round-to-nearest difference is at most half an ulp short, so [centre - half, centre + half] contains [mn, mx].
inline void box_center_half(float mn, float mx, float& ctr, float& half)
{
    ctr = mn * 0.5f + mx * 0.5f;
    const float up = mx - ctr, dn = ctr - mn;
    float h = up > dn ? up : dn;
    if (h > 0.0f && h < 3.0e38f) {
        uint32_t b;
        std::memcpy(&b, &h, 4);
        b += 1u;
        std::memcpy(&h, &b, 4);
    }
    half = h;
}

int flatten_scene(const rt_scene_desc& d, FlatScene& out, std::string& err);
void flatten_small(const rt_scene_desc& d, FlatScene& out); // mats, lights, n_lights, ambient only

} // namespace rt

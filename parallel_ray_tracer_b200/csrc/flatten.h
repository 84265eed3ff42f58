// flatten.h — host staging buffers of the HBM layout (device_layout.h), produced by flatten.cpp.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "rt_b200.h"

// host mirrors of the device_layout.h constants (that header needs CUDA vector types)
#define RT_REF_NONE_HOST ((int32_t)0x80000000)
#define RT_LEAF_CNT_ESC_HOST 15
#define RT_STACK_ENTRIES_HOST 40

namespace rt {

struct FlatScene {
    std::vector<float>   nodes;    // 16 floats (4 x float4) per inner node
    std::vector<float>   nodes4;   // 32 floats (8 x float4) per 4-wide node (fast build)
    std::vector<float>   tris;     // 16 floats (4 x float4) per leaf-order slot
    std::vector<float>   shade;    // 4 floats per original triangle
    std::vector<float>   mats;     // 12 floats per material
    std::vector<float>   lights;   // 8 floats per light
    std::vector<int32_t> leaf_cnt; // empty unless some leaf holds >= 15 triangles
    uint32_t n_lights = 0;
    float ambient[3] = {0, 0, 0};
    int max_depth = 0;
    int stack_need4 = 0;           // stack slots a ray can need on the 4-wide tree

    size_t bytes() const
    {
        return 4 * (nodes.size() + nodes4.size() + tris.size() + shade.size() + mats.size() + lights.size() + leaf_cnt.size());
    }
};

int flatten_scene(const rt_scene_desc& d, FlatScene& out, std::string& err);

} // namespace rt

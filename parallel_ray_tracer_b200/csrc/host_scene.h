// host_scene.h — host-side scene container behind the opaque rt_scene of include/rt_b200.h.
#pragma once
#include <cstdint>
#include <exception>
#include <new>
#include <string>
#include <vector>

#include "rt_b200.h"

struct rt_scene {
    std::vector<float>       tri;      // 9 per triangle: v0 v1 v2
    std::vector<uint32_t>    tri_mat;  // material index per triangle
    std::vector<float>       mats;     // 9 per material: ks kd kr
    std::vector<float>       lights;   // 6 per light: pos kl
    float                    ambient[3] = {0.5f, 0.5f, 0.5f}; // amb_light, cpu/src/main.c:37
    std::vector<rt_bvh_node> bvh;      // reference node layout
    std::vector<int32_t>     tri_idx;

    uint32_t n_tris() const { return (uint32_t)(tri.size() / 9); }
    uint32_t n_mats() const { return (uint32_t)(mats.size() / 9); }
    uint32_t n_lights() const { return (uint32_t)(lights.size() / 6); }
};

namespace rt {

// thread-local error text for calls that have no context yet
void set_error(const std::string& msg);
const char* get_error();

// No C++ exception may cross the C ABI: extern "C" entry points that allocate run their body through this.
template <class F>
inline int guarded(const char* who, F f)
{
    try { return f(); }
    catch (const std::bad_alloc&) { set_error(std::string(who) + ": out of memory"); return RT_ERR_NOMEM; }
    catch (const std::exception& e) { set_error(std::string(who) + ": " + e.what()); return RT_ERR_INVALID; }
}

// BVH build variants (bvh_build.cpp)
enum BvhArith {
    BVH_IEEE = 0,   // IEEE reading of the reference source (= the reference GPU program's host build)
    BVH_REFBIN = 1  // the four contracted expressions of the reference CPU binary (see bvh_build.cpp)
};
int build_bvh(rt_scene& s, int heuristic, BvhArith arith, int threads);

} // namespace rt

// gpu_tree.h — device-resident results of the GPU BVH build (bvh_build_gpu.cu) and of the device-side flatten
// (flatten_gpu.cu), shared with rt_api.cu: the scene can go from triangles to a render-ready context without its tree
// ever visiting the host.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <string>

#include "host_scene.h"

namespace rt {

// Reference-layout tree on one device.  Node numbering is the reference's (children allocated pairwise in the
// depth-first order of the splits, cpu/src/bvh.c:98-99, 265-266), so the k-th split node — pre-order rank k among the
// inner nodes — has its children at 1 + 2k.
struct GpuTree {
    int device = 0;
    size_t n_tris = 0, n_nodes = 0;
    float* tri = nullptr;             // 9 floats per triangle (as uploaded)
    int* tri_idx = nullptr;           // leaf-order permutation
    rt_bvh_node* nodes = nullptr;     // n_nodes
    unsigned char* depth = nullptr;   // depth of every node (root 0)
    bool fell_back = false;           // degenerate input: the host builder produced s.bvh / s.tri_idx instead
    void release()
    {
        if (tri || tri_idx || nodes || depth) cudaSetDevice(device);
        cudaFree(tri); cudaFree(tri_idx); cudaFree(nodes); cudaFree(depth);
        tri = nullptr; tri_idx = nullptr; nodes = nullptr; depth = nullptr;
    }
    ~GpuTree() { release(); }
    GpuTree() = default;
    GpuTree(const GpuTree&) = delete;
    GpuTree& operator=(const GpuTree&) = delete;
};

// Build the heuristic-6 tree of `s` on `device`.  keep != nullptr: the result stays on the device in *keep (and the host
// scene is only filled when the build fell back to the host builder); keep == nullptr: the arrays are copied into s.
int gpu_build_bvh(rt_scene& s, int refbin, int device, rt_bvh_gpu_stats* stats, GpuTree* keep);

// The HBM layout of device_layout.h produced on the device from a GpuTree (flatten_gpu.cu); same arrays, bit for bit,
// as flatten_scene (flatten.cpp) makes on the host.  Pointers are owned by the caller (cudaFree).
struct DeviceFlat {
    float4 *nodes = nullptr, *nodes4 = nullptr, *tris = nullptr, *shade = nullptr;
    uint4* nodes8 = nullptr;
    int* leaf_cnt = nullptr;
    size_t n_inner = 0, n_nodes4 = 0, n_nodes8 = 0, n_tris = 0;
    int max_depth = 0, stack_need4 = 0, depth8 = 0;
    size_t bytes() const { return 64 * n_inner + 128 * n_nodes4 + 96 * n_nodes8 + 64 * n_tris + 16 * n_tris + (leaf_cnt ? 4 * n_tris : 0); }
};
int flatten_gpu(const GpuTree& t, const uint32_t* host_tri_mat, uint32_t n_mats, DeviceFlat& out, std::string& err);

} // namespace rt

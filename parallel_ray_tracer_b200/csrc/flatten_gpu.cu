// flatten_gpu.cu — reference-layout tree (GpuTree, on the device) -> the HBM layout of device_layout.h, on the device.
// The device-side twin of flatten.cpp: same records bit for bit (this file is compiled -fmad=false -prec-div=true
// -prec-sqrt=true like the host file is compiled -ffp-contract=off), so a scene can go from triangles to a render-ready
// context without its tree visiting the host (gpu/src/gpu.cu:129-201 marshals on the host and copies synchronously).
//
// What makes it parallel: in the reference's node numbering the k-th node that was split — pre-order rank k among the
// inner nodes, which is the record index flatten.cpp assigns by a depth-first walk — has its children at 1 + 2k, so
// record index = (child index - 1) / 2 with no traversal.  The 4-wide collapse keeps the inner nodes at even depth
// (flatten.cpp: kids_of), in the same pre-order: their indices are an exclusive scan of the even-depth flags.  The
// per-node stack need of the 4-wide tree is a bottom-up maximum, done level by level (<= 17 levels).
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "device_layout.h"
#include "gpu_tree.h"
#include "staged_copy.h"
#include "wide8.h"

#ifndef RT_W8_DEVICE_BUILD_DEFAULT
#define RT_W8_DEVICE_BUILD_DEFAULT true
#endif

namespace {

__device__ __forceinline__ bool is_inner(const rt_bvh_node& b) { return b.tr_len == 0 && b.idx != 0; }
__device__ __forceinline__ int leaf_ref(const rt_bvh_node& b) // flatten.cpp: leaf_ref
{
    if (b.tr_len <= 0) return RT_REF_NONE; // empty leaf: never pushed
    const int cnt = b.tr_len >= RT_LEAF_CNT_ESC ? RT_LEAF_CNT_ESC : b.tr_len;
    return ~((b.idx << 4) | cnt);
}

struct FlatFlags { int bad_material, need_leaf_cnt, max_depth, bad_leaf; };

// triangles in leaf order: (v0, e1, e2, n = e1 x e2, original index), raytracer.c:36-38
__global__ void tris_kernel(int n, const float* __restrict__ tri, const int* __restrict__ tri_idx, float4* __restrict__ out)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int orig = tri_idx[j];
    const float* c = tri + 9 * (size_t)orig;
    const float e1[3] = {c[3] - c[0], c[4] - c[1], c[5] - c[2]};
    const float e2[3] = {c[6] - c[0], c[7] - c[1], c[8] - c[2]};
    const float nn[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]}; // vec_cross
    float4* q = out + 4 * (size_t)j;
    q[0] = make_float4(c[0], c[1], c[2], e1[0]);
    q[1] = make_float4(e1[1], e1[2], e2[0], e2[1]);
    q[2] = make_float4(e2[2], nn[0], nn[1], nn[2]);
    q[3] = make_float4(__int_as_float(orig), 0.f, 0.f, 0.f);
}

// unit normal norm[0] + material index per original triangle, triangle.c:14-17
__global__ void shade_kernel(int n, const float* __restrict__ tri, const unsigned* __restrict__ tri_mat, unsigned n_mats, float4* __restrict__ out, FlatFlags* flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* c = tri + 9 * (size_t)i;
    const float e1[3] = {c[3] - c[0], c[4] - c[1], c[5] - c[2]};
    const float e2[3] = {c[6] - c[0], c[7] - c[1], c[8] - c[2]};
    const float nn[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    const float mag = sqrtf(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]); // vec_mag
    unsigned m = tri_mat ? tri_mat[i] : 0u;
    if (m >= n_mats) { flags->bad_material = 1; m = 0; }
    out[i] = make_float4(nn[0] / mag, nn[1] / mag, nn[2] / mag, __uint_as_float(m)); // vec_normalize
}

__device__ __forceinline__ void child_box(const rt_bvh_node& c, int ref, float* mn, float* mx)
{
    // an empty child must never be entered: degenerate box at +infinity (flatten.cpp: put_child)
    for (int a = 0; a < 3; a++) { mn[a] = ref == RT_REF_NONE ? INFINITY : c.min[a]; mx[a] = ref == RT_REF_NONE ? INFINITY : c.max[a]; }
}

__device__ __forceinline__ int child_ref(const rt_bvh_node& c, int n_tris, int* leaf_cnt, FlatFlags* flags)
{
    if (is_inner(c)) return (c.idx - 1) >> 1;
    if (c.tr_len > 0 && (c.idx < 0 || (long long)c.idx + c.tr_len > n_tris)) { flags->bad_leaf = 1; return RT_REF_NONE; }
    if (c.tr_len >= RT_LEAF_CNT_ESC) {
        if (leaf_cnt) leaf_cnt[c.idx] = c.tr_len;
        else flags->need_leaf_cnt = 1;
    }
    return leaf_ref(c);
}

// 2-wide records: one per inner node, index = pre-order rank = (child index - 1) / 2; even-depth flags for the 4-wide scan
__global__ void nodes_kernel(int n_nodes, int n_tris, const rt_bvh_node* __restrict__ nodes, const unsigned char* __restrict__ depth, float4* __restrict__ out,
                             int* __restrict__ is_even, int* leaf_cnt, FlatFlags* flags)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_nodes) return;
    const rt_bvh_node b = nodes[v];
    if (!is_inner(b)) return;
    const int k = (b.idx - 1) >> 1;
    const rt_bvh_node c0 = nodes[b.idx], c1 = nodes[b.idx + 1];
    const int r0 = child_ref(c0, n_tris, leaf_cnt, flags), r1 = child_ref(c1, n_tris, leaf_cnt, flags);
    float mn0[3], mx0[3], mn1[3], mx1[3];
    child_box(c0, r0, mn0, mx0);
    child_box(c1, r1, mn1, mx1);
    float4* q = out + 4 * (size_t)k;
    q[0] = make_float4(mn0[0], mn0[1], mn0[2], mx0[0]);
    q[1] = make_float4(mx0[1], mx0[2], mn1[0], mn1[1]);
    q[2] = make_float4(mn1[2], mx1[0], mx1[1], mx1[2]);
    q[3] = make_float4(__int_as_float(r0), __int_as_float(r1), 0.f, 0.f);
    is_even[k] = (depth[v] & 1) == 0;
    atomicMax(&flags->max_depth, (int)depth[v]);
}

// ---- compressed 8-wide tree (wide8.h), level by level exactly as wide8.cpp builds it on the host ----
// pass A: inner children per node of the level
__global__ void w8_count_kernel(int n, int leaf_max, int width, const unsigned* __restrict__ level, const rt_bvh_node* __restrict__ bvh, int* __restrict__ n_inner)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    rt::W8Child ch[8];
    const int c = rt::w8_expand(bvh, level[i], leaf_max, ch, width);
    int k = 0;
    for (int j = 0; j < c; j++) k += ch[j].inner;
    n_inner[i] = k;
}
// pass B: the next level's node list (inner children in slot order, the order their indices are assigned in) and the
// level's plan: per (node, slot) the reference node whose box the slot holds (-1 = empty) and the slot's final reference
// (width 8: slots by octant order, wide8.h; width 4 — the FP32 4-wide tree, build_wide4 — children stay in expansion order)
__global__ void w8_plan_kernel(int n, int leaf_max, int width, const unsigned* __restrict__ level, const rt_bvh_node* __restrict__ bvh, const int* __restrict__ offset,
                               unsigned next_base, unsigned* __restrict__ next, int* __restrict__ plan_node, int* __restrict__ plan_ref)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    rt::W8Child ch[8];
    int slot_of[8];
    const int c = rt::w8_expand(bvh, level[i], leaf_max, ch, width);
    if (width == 8) rt::w8_assign_slots(ch, c, slot_of);
    else for (int j = 0; j < 8; j++) slot_of[j] = j;
    for (int s = 0; s < width; s++) { plan_node[width * (size_t)i + s] = -1; plan_ref[width * (size_t)i + s] = RT_REF_NONE; }
    int m = 0;
    for (int s = 0; s < width; s++)
        for (int j = 0; j < c; j++)
            if (slot_of[j] == s) {
                plan_node[width * (size_t)i + s] = ch[j].bnode;
                if (ch[j].inner) {
                    plan_ref[width * (size_t)i + s] = (int)(next_base + (unsigned)offset[i] + (unsigned)m);
                    next[offset[i] + m] = (unsigned)ch[j].bnode;
                    m++;
                } else plan_ref[width * (size_t)i + s] = rt::w8_leaf_ref(ch[j].first, ch[j].cnt);
            }
}
// the records of one level of the 4-wide FP32 tree (device_layout.h: nodes4), one thread per (node, slot): centre and half extent of
// the slot's box bit for bit as flatten.h: box_center_half computes them (this file: -fmad=false), the reference, the pad
__global__ void w4_encode_kernel(int n, const rt_bvh_node* __restrict__ bvh, const int* __restrict__ plan_node, const int* __restrict__ plan_ref,
                                 unsigned base, float* __restrict__ nodes4)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = tid >> 2, slot = tid & 3;
    if (i >= n) return;
    const int mine = plan_node[4 * (size_t)i + slot], ref = plan_ref[4 * (size_t)i + slot];
    float* q = nodes4 + 32 * (size_t)(base + (unsigned)i);
    const bool live = mine >= 0 && ref != RT_REF_NONE;
    for (int a = 0; a < 3; a++) {
        float ctr = INFINITY, h = 0.f; // empty slot: centre +inf, half extent 0
        if (live) {
            const float mn = bvh[mine].min[a], mx = bvh[mine].max[a];
            ctr = mn * 0.5f + mx * 0.5f;
            const float up = mx - ctr, dn = ctr - mn;
            h = up > dn ? up : dn;
            if (h > 0.0f && h < 3.0e38f) h = __uint_as_float(__float_as_uint(h) + 1u);
        }
        q[4 * a + slot] = ctr; q[12 + 4 * a + slot] = h;
    }
    q[24 + slot] = __int_as_float(live ? ref : RT_REF_NONE);
    q[28 + slot] = 0.f;
}
// stack need of a ray on the 4-wide tree, one level per launch, deepest first: a node with c live children enters one and leaves at
// most c - 1 pushed (wide8.cpp: build_wide4)
__global__ void w4_need_kernel(int n, unsigned base, const float* __restrict__ nodes4, int* __restrict__ need4)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* q = nodes4 + 32 * (size_t)(base + (unsigned)i);
    int live = 0, deepest = 0;
    for (int s = 0; s < 4; s++) {
        const int ref = __float_as_int(q[24 + s]);
        if (ref == RT_REF_NONE) continue;
        live++;
        if (ref >= 0) deepest = max(deepest, need4[ref]);
    }
    need4[base + (unsigned)i] = max(live - 1, 0) + deepest;
}
// pass C: the records of the level, one thread per (node, slot): the node's grid from the union of its children's boxes
// (recomputed by each of the eight threads), the slot's own six bytes and reference, the header by slot 0
__global__ void w8_encode_kernel(int n, const rt_bvh_node* __restrict__ bvh, const int* __restrict__ plan_node, const int* __restrict__ plan_ref,
                                 unsigned base, unsigned* __restrict__ words)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = tid >> 3, slot = tid & 7;
    if (i >= n) return;
    // every slot contributes its child's box; the union, the child count and the grid are formed once per node (lane 0 of
    // the eight) and handed to the other seven by shuffles (blockDim is a multiple of 32, a node never straddles a warp,
    // and whole 8-lane groups leave together above)
    const int mine = plan_node[8 * (size_t)i + slot];
    const unsigned gm = 0xffu << (threadIdx.x & 24u);
    float lo[3], hi[3];
    for (int a = 0; a < 3; a++) {
        lo[a] = mine >= 0 ? bvh[mine].min[a] : INFINITY;
        hi[a] = mine >= 0 ? bvh[mine].max[a] : -INFINITY;
    }
    for (int k = 1; k < 8; k <<= 1)
        for (int a = 0; a < 3; a++) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(gm, lo[a], k));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(gm, hi[a], k));
        }
    const int cnt = __popc(__ballot_sync(gm, mine >= 0) & gm);
    int e[3] = {0, 0, 0};
    float p[3] = {0, 0, 0};
    if (slot == 0)
        for (int a = 0; a < 3; a++) rt::w8_node_axis(lo[a], hi[a], &e[a], &p[a]);
    const int src = (int)(threadIdx.x & 24u);
    for (int a = 0; a < 3; a++) { e[a] = __shfl_sync(gm, e[a], src); p[a] = __shfl_sync(gm, p[a], src); }
    unsigned* w = words + (size_t)(base + (unsigned)i) * rt::kWide8Words;
    unsigned char* q = reinterpret_cast<unsigned char*>(w + 4);
    for (int a = 0; a < 3; a++) {
        unsigned char ql = 255, qh = 0; // empty slot: inverted, never hit
        if (mine >= 0) rt::w8_quantize(bvh[mine].min[a], bvh[mine].max[a], p[a], e[a], &ql, &qh);
        q[8 * a + slot] = ql;
        q[24 + 8 * a + slot] = qh;
    }
    w[16 + slot] = (unsigned)plan_ref[8 * (size_t)i + slot];
    if (slot == 0) {
        w[0] = __float_as_uint(p[0]); w[1] = __float_as_uint(p[1]); w[2] = __float_as_uint(p[2]);
        w[3] = rt::w8_header_word(e, cnt);
    }
}

struct Buf {
    void* p = nullptr;
    bool pooled = false; // stream-ordered allocation (cudaMallocAsync on the default stream): no device-wide sync per free
    ~Buf() { if (pooled) cudaFreeAsync(p, 0); else cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
    cudaError_t alloc_pooled(size_t bytes) { pooled = true; return cudaMallocAsync(&p, bytes ? bytes : 16, 0); }
    template <class T> T* as() const { return static_cast<T*>(p); }
    template <class T> T* take() { T* r = static_cast<T*>(p); p = nullptr; return r; }
};

} // namespace

int rt::flatten_gpu(const GpuTree& t, const uint32_t* host_tri_mat, uint32_t n_mats, DeviceFlat& out, std::string& err)
{
#define CKF(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { err = std::string("flatten_gpu: ") + #call + " failed: " + cudaGetErrorString(e__); return RT_ERR_CUDA; } } while (0)
    const bool timing = std::getenv("RT_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto stage = [&](const char* what) {
        if (!timing) return;
        cudaDeviceSynchronize();
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[flatten_gpu] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    const int n = (int)t.n_tris, nn = (int)t.n_nodes;
    if (n < 1 || nn < 3) { err = "flatten_gpu: the tree has no inner node"; return RT_ERR_INVALID; }
    const size_t n_inner = (size_t)(nn - 1) / 2; // every split allocated two nodes
    CKF(cudaSetDevice(t.device));
    Buf tris, shade, nodes, nodes4, leaf_cnt, is_even, need4, flags, mat;
    CKF(tris.alloc((size_t)n * 64));
    CKF(shade.alloc((size_t)n * 16));
    CKF(nodes.alloc(n_inner * 64));
    CKF(is_even.alloc(n_inner * 4));
    CKF(flags.alloc(sizeof(FlatFlags)));
    CKF(cudaMemset(flags.p, 0, sizeof(FlatFlags)));
    if (host_tri_mat) {
        CKF(mat.alloc((size_t)n * 4));
        CKF(rt::staged_h2d(mat.p, host_tri_mat, (size_t)n * 4, 0));
    }
    stage("alloc + tri_mat upload");
    const int B = 256;
    tris_kernel<<<(n + B - 1) / B, B>>>(n, t.tri, t.tri_idx, tris.as<float4>());
    shade_kernel<<<(n + B - 1) / B, B>>>(n, t.tri, host_tri_mat ? mat.as<unsigned>() : nullptr, n_mats ? n_mats : 1u, shade.as<float4>(), flags.as<FlatFlags>());
    nodes_kernel<<<(nn + B - 1) / B, B>>>(nn, n, t.nodes, t.depth, nodes.as<float4>(), is_even.as<int>(), nullptr, flags.as<FlatFlags>());
    CKF(cudaGetLastError());
    FlatFlags fl;
    CKF(cudaMemcpy(&fl, flags.p, sizeof fl, cudaMemcpyDeviceToHost));
    if (fl.bad_material) { err = "material index out of range"; return RT_ERR_INVALID; }
    if (fl.bad_leaf) { err = "BVH leaf range out of bounds"; return RT_ERR_INVALID; }
    if (fl.max_depth + 5 > RT_STACK_ENTRIES) { err = "BVH deeper than the traversal stack (" + std::to_string(fl.max_depth) + ")"; return RT_ERR_INVALID; }
    if (fl.need_leaf_cnt) { // leaves of >= 15 triangles (depth-capped): second pass fills the side table
        CKF(leaf_cnt.alloc((size_t)n * 4));
        CKF(cudaMemset(leaf_cnt.p, 0, (size_t)n * 4));
        nodes_kernel<<<(nn + B - 1) / B, B>>>(nn, n, t.nodes, t.depth, nodes.as<float4>(), is_even.as<int>(), leaf_cnt.as<int>(), flags.as<FlatFlags>());
        CKF(cudaGetLastError());
    }
    stage("tris / shade / nodes");
    // ---- the two wide trees of the fast build, level by level exactly as wide8.cpp builds them on the host: first the node lists
    // of all levels (count inner children, exclusive scan, plan every (node, slot)), then the records ----
    {   // keep freed blocks in the pool between levels instead of returning them to the driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, t.device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    const bool dbg = std::getenv("RT_SYNC_DEBUG") != nullptr; // name the failing kernel (no compute-sanitizer on the pool)
    struct Levels {
        std::vector<Buf> lists, plan_nodes, plan_refs; // per level: node list, (node, slot) plan
        std::vector<int> sizes;
        size_t n_nodes = 0;
        int depth = 0;
    };
    Buf tmp;
    size_t tmp_cap = 0;
    auto plan_levels = [&](int width, int leaf_max, Levels& L) -> int {
#define CKL(what) do { if (dbg) { cudaError_t e__ = cudaDeviceSynchronize(); if (e__ != cudaSuccess) { err = std::string("flatten_gpu: ") + what + " (width " + std::to_string(width) + ", level " + std::to_string(lv) + "): " + cudaGetErrorString(e__); return RT_ERR_CUDA; } } } while (0)
        L.lists.reserve(72); L.plan_nodes.reserve(72); L.plan_refs.reserve(72); // (Buf is not movable: no reallocation; <= 66 levels, checked below)
        L.lists.emplace_back();
        CKF(L.lists[0].alloc_pooled(4));
        CKF(cudaMemset(L.lists[0].p, 0, 4)); // level 0 = {reference node 0}
        L.sizes.push_back(1);
        size_t done = 0; // nodes of the levels before this one
        for (int lv = 0; L.sizes[lv] > 0; lv++) {
            if (lv > 64) { err = "flatten_gpu: wide tree deeper than 64 levels"; return RT_ERR_INVALID; }
            const int m = L.sizes[lv];
            Buf c2, off;
            CKF(c2.alloc_pooled(((size_t)m + 1) * 4));
            CKF(cudaMemset(c2.p, 0, ((size_t)m + 1) * 4));
            w8_count_kernel<<<(m + 127) / 128, 128>>>(m, leaf_max, width, L.lists[lv].as<unsigned>(), t.nodes, c2.as<int>());
            CKL("w8_count_kernel");
            CKF(off.alloc_pooled(((size_t)m + 1) * 4));
            size_t need = 0;
            CKF(cub::DeviceScan::ExclusiveSum(nullptr, need, c2.as<int>(), off.as<int>(), m + 1));
            if (need > tmp_cap) { if (tmp.p) cudaFreeAsync(tmp.p, 0); tmp.p = nullptr; CKF(tmp.alloc_pooled(need)); tmp_cap = need; }
            CKF(cub::DeviceScan::ExclusiveSum(tmp.p, need, c2.as<int>(), off.as<int>(), m + 1));
            int total = 0;
            CKF(cudaMemcpy(&total, off.as<int>() + m, 4, cudaMemcpyDeviceToHost));
            L.lists.emplace_back(); L.plan_nodes.emplace_back(); L.plan_refs.emplace_back();
            CKF(L.lists[lv + 1].alloc_pooled((size_t)std::max(total, 1) * 4));
            CKF(L.plan_nodes[lv].alloc_pooled((size_t)m * 4 * width));
            CKF(L.plan_refs[lv].alloc_pooled((size_t)m * 4 * width));
            w8_plan_kernel<<<(m + 127) / 128, 128>>>(m, leaf_max, width, L.lists[lv].as<unsigned>(), t.nodes, off.as<int>(), (unsigned)(done + (size_t)m),
                                                     L.lists[lv + 1].as<unsigned>(), L.plan_nodes[lv].as<int>(), L.plan_refs[lv].as<int>());
            CKF(cudaGetLastError());
            CKL("w8_plan_kernel");
            L.sizes.push_back(total);
            done += (size_t)m;
            L.n_nodes += (size_t)m;
            L.depth++;
        }
#undef CKL
        return RT_OK;
    };
    // 4-wide FP32 tree (wide8.h: build_wide4)
    Levels L4;
    int rc_l = plan_levels(4, rt::wide4_leaf_max(), L4);
    if (rc_l) return rc_l;
    const size_t n4 = L4.n_nodes;
    CKF(nodes4.alloc(n4 * 128));
    CKF(need4.alloc(n4 * 4));
    {
        size_t base = 0;
        std::vector<size_t> bases;
        for (int lv = 0; lv < L4.depth; lv++) {
            const int m = L4.sizes[lv];
            bases.push_back(base);
            w4_encode_kernel<<<(4 * m + 127) / 128, 128>>>(m, t.nodes, L4.plan_nodes[lv].as<int>(), L4.plan_refs[lv].as<int>(), (unsigned)base, nodes4.as<float>());
            base += (size_t)m;
        }
        for (int lv = L4.depth - 1; lv >= 0; lv--) {
            const int m = L4.sizes[lv];
            w4_need_kernel<<<(m + B - 1) / B, B>>>(m, (unsigned)bases[lv], nodes4.as<float>(), need4.as<int>());
        }
        CKF(cudaGetLastError());
    }
    int need_root = 0;
    CKF(cudaMemcpy(&need_root, need4.p, 4, cudaMemcpyDeviceToHost));
    stage("4-wide tree");
    // compressed 8-wide tree
    Buf nodes8;
    size_t n8 = 0;
    int depth8 = 0;
    // RT_W8_DEVICE_BUILD=0 builds the 8-wide tree on the host from a copy of the device tree (wide8.cpp) and uploads it
    const char* w8_env = std::getenv("RT_W8_DEVICE_BUILD");
    const bool w8_on_device = w8_env ? std::atoi(w8_env) != 0 : RT_W8_DEVICE_BUILD_DEFAULT;
    if (!w8_on_device) {
        std::vector<rt_bvh_node> host_nodes((size_t)nn);
        CKF(cudaDeviceSynchronize());
        CKF(rt::staged_d2h(host_nodes.data(), t.nodes, (size_t)nn * sizeof(rt_bvh_node), 0));
        rt::Wide8Tree w8;
        if (rt::build_wide8(host_nodes.data(), (uint32_t)nn, rt::wide8_leaf_max(), w8)) { err = "flatten_gpu: 8-wide collapse failed"; return RT_ERR_INVALID; }
        n8 = w8.n_nodes(); depth8 = w8.depth;
        CKF(nodes8.alloc(n8 * 96));
        CKF(rt::staged_h2d(nodes8.p, w8.words.data(), n8 * 96, 0));
    } else {
        Levels L8;
        rc_l = plan_levels(8, rt::wide8_leaf_max(), L8);
        if (rc_l) return rc_l;
        n8 = L8.n_nodes; depth8 = L8.depth;
        stage("8-wide: levels (count/scan/plan)");
        CKF(nodes8.alloc(n8 * 96));
        size_t base = 0;
        for (int lv = 0; lv < depth8; lv++) {
            const int m = L8.sizes[lv];
            w8_encode_kernel<<<(8 * m + 127) / 128, 128>>>(m, t.nodes, L8.plan_nodes[lv].as<int>(), L8.plan_refs[lv].as<int>(), (unsigned)base, nodes8.as<unsigned>());
            if (dbg) { cudaError_t e__ = cudaDeviceSynchronize(); if (e__ != cudaSuccess) { err = std::string("flatten_gpu: w8_encode_kernel (level ") + std::to_string(lv) + "): " + cudaGetErrorString(e__); return RT_ERR_CUDA; } }
            base += (size_t)m;
        }
        CKF(cudaGetLastError());
    }
    CKF(cudaDeviceSynchronize());

    out.nodes = nodes.take<float4>(); out.nodes4 = nodes4.take<float4>(); out.tris = tris.take<float4>(); out.shade = shade.take<float4>();
    out.leaf_cnt = fl.need_leaf_cnt ? leaf_cnt.take<int>() : nullptr;
    out.nodes8 = nodes8.take<uint4>(); out.n_nodes8 = n8; out.depth8 = depth8;
    out.n_inner = n_inner; out.n_nodes4 = (size_t)n4; out.n_tris = (size_t)n;
    out.max_depth = fl.max_depth;
    out.stack_need4 = need_root + 3; // + sentinel, postponed leaf, slack (flatten.cpp)
#undef CKF
    return RT_OK;
}

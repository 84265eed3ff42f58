// bvh_build_gpu.cu — the reference's BVH (bvh_build / bvh_split, cpu/src/bvh.c:78-267, 360-388, heuristic 6) built on the
// GPU, node for node and tri_idx entry for entry what csrc/bvh_build.cpp (and so the reference) produces.
// the host build gates the 50 M-triangle configuration and any dynamic scene.
//
// The reference algorithm is sequential in three places; each is restated in a parallel form that gives the same bits:
//
//   * candidate evaluation (bvh.c:138-177: 3 axes x 32 planes, one pass over the node's triangles per candidate):
//     the planes are non-decreasing in i, so one binning pass per axis (count + vertex bounds per threshold bin) and
//     prefix / suffix unions give every candidate's (cl, box_L, cr, box_R) exactly — min / max and integer counts are
//     order independent — and the cost is evaluated with the reference's float expression, candidates in its order
//     (same argument as bvh_build.cpp);
//   * the in-place forward partition (bvh.c:244-259): `for i: if left(A[i]) swap(A[i], A[first + n_left++])`.
//     Lefts end up in encounter order (a stable compaction).  The pending rights occupy [n_left, i) as a queue whose
//     FRONT is moved to position i by every left that is met; position P is consumed as the front exactly once, by the
//     left of rank P, which sits at posL[P].  So the right that starts at x ends at the first element of the chain
//     x -> posL[x] -> posL[posL[x]] ... that is >= n_left.  The chains are disjoint and their total length is <= n, so
//     following them in parallel is O(n) work; the depth is the longest chain (logarithmic for interleaved input);
//   * node numbering (children allocated pairwise when the parent is split, left subtree first, bvh.c:98-99, 265-266):
//     nodes are created in whatever order the GPU builds them and numbered afterwards from subtree sizes.
//
// Structure: the top of the tree (nodes of >= kSmall triangles) is built level by level with many CTAs per node
// (binning with shared-memory bins, one global merge per CTA; scan + scatter partition); every node below that size is
// the root of a subtree that ONE WARP finishes in the reference's own depth-first order (bins in shared memory, ballot
// compaction, the same chain rule), thousands of subtrees at once.  Arithmetic: this file is compiled -fmad=false -prec-div=true (csrc/Makefile), the two flavours of
// bvh_build.cpp (IEEE source reading / the reference CPU binary's four contracted expressions) are both available.
//
// Inputs that trip the reference's degenerate-input guards (bvh_len >= 2N, bvh.c:80-83) or would need pathological
// chain lengths are handed to the host builder (same output by construction; reported in rt_bvh_gpu_stats).
#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "gpu_tree.h"
#include "staged_copy.h"
#include "host_scene.h"

namespace {

constexpr int kMaxDepth = 32;     // BVH_MAX_ITER, cpu/include/options.h:64
constexpr int kLeafThreshold = 2; // BVH_ELEMENT_THRESHOLD, cpu/include/options.h:58
constexpr int kBins = 32;         // SAH_BIN_SIZE, cpu/include/options.h:61
constexpr int kNB = kBins + 1;    // threshold bins per axis
constexpr int kSmallDefault = 256; // nodes below this size are finished by one warp (RT_BVH_GPU_SMALL overrides, for experiments)
constexpr int kChunk = 2048;      // triangles per CTA work item of the level-synchronous top
constexpr int kCta = 256;
constexpr int kPer = kChunk / kCta;
constexpr int kChainCap = 1 << 16; // longest right-element chain followed on the device

struct BNode { // build node of the level-synchronous top (BFS order of creation)
    float mn[3], mx[3];
    int first, len, depth, child; // child: BFS index of the left child, -1 = not split
    int axis;
    float pos;
    int nl, pad;
};

struct BinSet { // per active node: 3 axes x 33 threshold bins
    int cnt[3 * kNB];
    float mn[3 * kNB * 3];
    float mx[3 * kNB * 3];
};

struct Flags { int chain_overflow, region_overflow; };

// ---- float atomics on non-negative-zero-free values (callers add +0.0f first) ----
__device__ __forceinline__ void atomic_min_f(float* a, float v)
{
    if (v >= 0.0f) atomicMin(reinterpret_cast<int*>(a), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned*>(a), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* a, float v)
{
    if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(a), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned*>(a), __float_as_uint(v));
}
__device__ __forceinline__ float fmn(float a, float b) { return b < a ? b : a; } // as bvh_build.cpp
__device__ __forceinline__ float fmx(float a, float b) { return b > a ? b : a; }

__device__ __forceinline__ float split_plane(float mn, float size, int i, int refbin)
{
    // bvh.c:156-157; REFBIN: fmaf((float)i, size * 0.03125f, min) as gcc contracts it in the reference binary
    return refbin ? __fmaf_rn((float)i, size * 0.03125f, mn) : mn + size * ((float)i / kBins);
}
__device__ __forceinline__ float diag2(const float* mn, const float* mx, int refbin)
{
    const float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
    if (refbin) return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, dy * dy));
    return dx * dx + dy * dy + dz * dz; // vec_dot(size, size), bvh.c:43-46
}
// k = min{ i : c < split[i] } (32 if none); split[] is non-decreasing
__device__ __forceinline__ int threshold_bin(const float* split, float c)
{
    int lo = 0, n = kBins;
    while (n > 0) {
        const int half = n >> 1;
        if (!(c < split[lo + half])) { lo += half + 1; n -= half + 1; } else n = half;
    }
    return lo;
}

// per-triangle record: q0 = (centroid.xyz, min.x) q1 = (min.y, min.z, max.x, max.y) q2 = (max.z, -, -, -)
__global__ void prepare_kernel(const float* __restrict__ tri, float4* __restrict__ info, int* __restrict__ tri_idx, int n, int refbin)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* t = tri + 9 * (size_t)i;
    float c[3], mn[3], mx[3];
    for (int a = 0; a < 3; a++) {
        const float p0 = t[a], p1 = t[3 + a], p2 = t[6 + a];
        c[a] = refbin ? ((p0 + p1) + p2) * 0.33333334f : (p0 + p1 + p2) / 3.0f; // triangle.c:21-23
        mn[a] = fmn(fmn(p0, p1), p2) + 0.0f; // +0.0f: no negative zeros in the bounds (see atomic_min_f)
        mx[a] = fmx(fmx(p0, p1), p2) + 0.0f;
    }
    info[3 * (size_t)i + 0] = make_float4(c[0], c[1], c[2], mn[0]);
    info[3 * (size_t)i + 1] = make_float4(mn[1], mn[2], mx[0], mx[1]);
    info[3 * (size_t)i + 2] = make_float4(mx[2], 0.f, 0.f, 0.f);
    tri_idx[i] = i; // bvh.c:366-368
}

__global__ void root_box_kernel(const float4* __restrict__ info, int n, BNode* root)
{
    __shared__ float s[6];
    if (threadIdx.x < 3) { s[threadIdx.x] = 1e10f; s[3 + threadIdx.x] = -1e10f; } // bvh.c:373-377
    __syncthreads();
    float mn[3] = {1e10f, 1e10f, 1e10f}, mx[3] = {-1e10f, -1e10f, -1e10f};
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 a = info[3 * i], b = info[3 * i + 1], c = info[3 * i + 2];
        mn[0] = fmn(mn[0], a.w); mn[1] = fmn(mn[1], b.x); mn[2] = fmn(mn[2], b.y);
        mx[0] = fmx(mx[0], b.z); mx[1] = fmx(mx[1], b.w); mx[2] = fmx(mx[2], c.x);
    }
    for (int a = 0; a < 3; a++) { atomic_min_f(&s[a], mn[a]); atomic_max_f(&s[3 + a], mx[a]); }
    __syncthreads();
    if (threadIdx.x < 3) { atomic_min_f(&root->mn[threadIdx.x], s[threadIdx.x]); atomic_max_f(&root->mx[threadIdx.x], s[3 + threadIdx.x]); }
}

__global__ void init_bins_kernel(BinSet* bins, int n_active)
{
    const size_t total = (size_t)n_active * (3 * kNB);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        BinSet& B = bins[i / (3 * kNB)];
        const int b = (int)(i % (3 * kNB));
        B.cnt[b] = 0;
        for (int a = 0; a < 3; a++) { B.mn[3 * b + a] = INFINITY; B.mx[3 * b + a] = -INFINITY; }
    }
}

// ---- level-synchronous top: binning, one CTA per (node, chunk) ----
__global__ void __launch_bounds__(kCta) bin_kernel(const int2* __restrict__ chunks, const int* __restrict__ active_ids, const BNode* __restrict__ nodes,
                                                   const int* __restrict__ tri_idx, const float4* __restrict__ info, BinSet* bins, int refbin)
{
    __shared__ float split[3][kBins];
    __shared__ int s_cnt[3 * kNB];
    __shared__ float s_mn[3 * kNB * 3], s_mx[3 * kNB * 3];
    const int slot = chunks[blockIdx.x].x, chunk = chunks[blockIdx.x].y;
    const BNode& P = nodes[active_ids[slot]];
    for (int b = threadIdx.x; b < 3 * kNB; b += kCta) {
        s_cnt[b] = 0;
        for (int a = 0; a < 3; a++) { s_mn[3 * b + a] = INFINITY; s_mx[3 * b + a] = -INFINITY; }
    }
    if (threadIdx.x < 3 * kBins) {
        const int axis = threadIdx.x / kBins, i = threadIdx.x % kBins;
        split[axis][i] = split_plane(P.mn[axis], P.mx[axis] - P.mn[axis], i, refbin);
    }
    __syncthreads();
    const int base = P.first + chunk * kChunk, cnt = min(kChunk, P.len - chunk * kChunk);
    for (int i = threadIdx.x; i < cnt; i += kCta) {
        const int ti = tri_idx[base + i];
        const float4 a = info[3 * (size_t)ti], b = info[3 * (size_t)ti + 1], c = info[3 * (size_t)ti + 2];
        const float cen[3] = {a.x, a.y, a.z}, mn[3] = {a.w, b.x, b.y}, mx[3] = {b.z, b.w, c.x};
        for (int axis = 0; axis < 3; axis++) {
            const int k = axis * kNB + threshold_bin(split[axis], cen[axis]);
            atomicAdd(&s_cnt[k], 1);
            for (int d = 0; d < 3; d++) { atomic_min_f(&s_mn[3 * k + d], mn[d]); atomic_max_f(&s_mx[3 * k + d], mx[d]); }
        }
    }
    __syncthreads();
    BinSet& B = bins[slot];
    for (int b = threadIdx.x; b < 3 * kNB; b += kCta) {
        if (!s_cnt[b]) continue;
        atomicAdd(&B.cnt[b], s_cnt[b]);
        for (int d = 0; d < 3; d++) { atomic_min_f(&B.mn[3 * b + d], s_mn[3 * b + d]); atomic_max_f(&B.mx[3 * b + d], s_mx[3 * b + d]); }
    }
}

// heuristic 6 (bvh.c:138-177) from the bins of one node, one axis: the first candidate plane i (ascending) whose cost is
// strictly below `best` wins (bvh.c:169-174).  bin(b): count and vertex bounds of threshold bin b of `axis`.
template <class BinCnt, class BinBox>
__device__ __forceinline__ void choose_axis(int axis, const float* pmn, const float* pmx, int refbin, BinCnt bin_cnt, BinBox bin_box,
                                            float& best, int& splitAxis, float& splitPos)
{
    const float size = pmx[axis] - pmn[axis];
    // suffix unions: R_i = bins i+1 .. 32
    float suf_mn[kNB][3], suf_mx[kNB][3];
    int sufc[kNB];
    float acc_mn[3] = {INFINITY, INFINITY, INFINITY}, acc_mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    int accc = 0;
    for (int b = kBins; b >= 1; b--) {
        float bm[3], bx[3];
        bin_box(axis, b, bm, bx);
        for (int a = 0; a < 3; a++) { acc_mn[a] = fmn(acc_mn[a], bm[a]); acc_mx[a] = fmx(acc_mx[a], bx[a]); }
        accc += bin_cnt(axis, b);
        for (int a = 0; a < 3; a++) { suf_mn[b - 1][a] = acc_mn[a]; suf_mx[b - 1][a] = acc_mx[a]; }
        sufc[b - 1] = accc;
    }
    float pre_mn[3] = {INFINITY, INFINITY, INFINITY}, pre_mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    int prec = 0;
    for (int i = 0; i < kBins; i++) {
        float bm[3], bx[3];
        bin_box(axis, i, bm, bx);
        for (int a = 0; a < 3; a++) { pre_mn[a] = fmn(pre_mn[a], bm[a]); pre_mx[a] = fmx(pre_mx[a], bx[a]); }
        prec += bin_cnt(axis, i);
        float al_mn[3], al_mx[3], ar_mn[3], ar_mx[3]; // candidate boxes start at min = FLT_MAX, max = FLT_MIN (bvh.c:149-150)
        for (int a = 0; a < 3; a++) {
            al_mn[a] = fmn(FLT_MAX, pre_mn[a]);
            al_mx[a] = fmx(FLT_MIN, pre_mx[a]);
            ar_mn[a] = fmn(FLT_MAX, suf_mn[i][a]);
            ar_mx[a] = fmx(FLT_MIN, suf_mx[i][a]);
        }
        const int cl = prec, cr = sufc[i];
        float score;
        if (refbin) score = __fmaf_rn((float)cl, diag2(al_mn, al_mx, 1), (float)cr * diag2(ar_mn, ar_mx, 1));
        else score = (float)cl * diag2(al_mn, al_mx, 0) + (float)cr * diag2(ar_mn, ar_mx, 0); // bvh.c:169
        if (score < best) { best = score; splitAxis = axis; splitPos = split_plane(pmn[axis], size, i, refbin); }
    }
}

template <class BinCnt, class BinBox>
__device__ __forceinline__ void choose_h6(const float* pmn, const float* pmx, int refbin, BinCnt bin_cnt, BinBox bin_box, int& splitAxis, float& splitPos)
{
    splitAxis = 0;
    splitPos = 0;
    float best = FLT_MAX;
    for (int axis = 0; axis < 3; axis++) choose_axis(axis, pmn, pmx, refbin, bin_cnt, bin_box, best, splitAxis, splitPos);
}

__global__ void choose_kernel(int n_active, const int* __restrict__ active_ids, BNode* nodes, const BinSet* __restrict__ bins, int child_base, int refbin)
{
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_active) return;
    BNode& P = nodes[active_ids[slot]];
    const BinSet& B = bins[slot];
    int axis;
    float pos;
    choose_h6(P.mn, P.mx, refbin, [&](int ax, int b) { return B.cnt[ax * kNB + b]; },
              [&](int ax, int b, float* m, float* x) { for (int a = 0; a < 3; a++) { m[a] = B.mn[3 * (ax * kNB + b) + a]; x[a] = B.mx[3 * (ax * kNB + b) + a]; } },
              axis, pos);
    P.axis = axis; P.pos = pos; P.child = child_base + 2 * slot; P.nl = 0;
    for (int k = 0; k < 2; k++) {
        BNode& C = nodes[P.child + k];
        for (int a = 0; a < 3; a++) { C.mn[a] = 1e10f; C.mx[a] = -1e10f; } // bvh.c:104-108
        C.first = P.first; C.len = 0; C.depth = P.depth + 1; C.child = -1; C.axis = 0; C.pos = 0; C.nl = 0; C.pad = 0;
    }
}

__device__ __forceinline__ float centroid_of(const float4* info, int ti, int axis)
{
    const float4 a = info[3 * (size_t)ti];
    return axis == 0 ? a.x : (axis == 1 ? a.y : a.z);
}

// lefts per chunk
__global__ void __launch_bounds__(kCta) count_kernel(const int2* __restrict__ chunks, const int* __restrict__ active_ids, const BNode* __restrict__ nodes,
                                                     const int* __restrict__ tri_idx, const float4* __restrict__ info, int* __restrict__ chunk_nl)
{
    const int slot = chunks[blockIdx.x].x, chunk = chunks[blockIdx.x].y;
    const BNode& P = nodes[active_ids[slot]];
    const int base = P.first + chunk * kChunk, cnt = min(kChunk, P.len - chunk * kChunk);
    int mine = 0;
    for (int i = threadIdx.x; i < cnt; i += kCta) mine += centroid_of(info, tri_idx[base + i], P.axis) < P.pos;
    __shared__ int s;
    if (threadIdx.x == 0) s = 0;
    __syncthreads();
    const int w = __reduce_add_sync(0xffffffffu, mine);
    if ((threadIdx.x & 31) == 0 && w) atomicAdd(&s, w);
    __syncthreads();
    if (threadIdx.x == 0) chunk_nl[blockIdx.x] = s;
}

// exclusive scan of chunk_nl inside every node's chunk range; one CTA per active node
__global__ void __launch_bounds__(kCta) scan_kernel(const int* __restrict__ chunk_base, const int* __restrict__ active_ids, BNode* nodes,
                                                    const int* __restrict__ chunk_nl, int* __restrict__ chunk_off)
{
    __shared__ int s[kCta];
    __shared__ int carry;
    const int slot = blockIdx.x;
    const int lo = chunk_base[slot], hi = chunk_base[slot + 1];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int t = lo; t < hi; t += kCta) {
        const int i = t + threadIdx.x;
        const int v = i < hi ? chunk_nl[i] : 0;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int d = 1; d < kCta; d <<= 1) {
            const int add = threadIdx.x >= d ? s[threadIdx.x - d] : 0;
            __syncthreads();
            s[threadIdx.x] += add;
            __syncthreads();
        }
        if (i < hi) chunk_off[i] = carry + s[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == kCta - 1) carry += s[kCta - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        BNode& P = nodes[active_ids[slot]];
        P.nl = carry;
        BNode &L = nodes[P.child], &R = nodes[P.child + 1];
        L.first = P.first; L.len = carry;            // bvh.c:248-255
        R.first = P.first + carry; R.len = P.len - carry;
    }
}

// lefts to their final place (stable), posL, child boxes
__global__ void __launch_bounds__(kCta) scatter_kernel(const int2* __restrict__ chunks, const int* __restrict__ active_ids, BNode* nodes,
                                                       const int* __restrict__ tri_idx, const float4* __restrict__ info, const int* __restrict__ chunk_off,
                                                       int* __restrict__ dst, int* __restrict__ posL)
{
    __shared__ int s[kCta];
    __shared__ float bx[2][6];
    const int slot = chunks[blockIdx.x].x, chunk = chunks[blockIdx.x].y;
    const BNode& P = nodes[active_ids[slot]];
    const int rel0 = chunk * kChunk, base = P.first + rel0, cnt = min(kChunk, P.len - rel0);
    if (threadIdx.x < 12) bx[threadIdx.x / 6][threadIdx.x % 6] = (threadIdx.x % 6) < 3 ? 1e10f : -1e10f;
    int ti[kPer];
    unsigned left_mask = 0;
    float lb[6] = {1e10f, 1e10f, 1e10f, -1e10f, -1e10f, -1e10f}, rb[6] = {1e10f, 1e10f, 1e10f, -1e10f, -1e10f, -1e10f};
    int n_right_mine = 0;
    for (int j = 0; j < kPer; j++) { // consecutive elements per thread: the scan keeps the encounter order
        const int i = threadIdx.x * kPer + j;
        if (i < cnt) {
            ti[j] = tri_idx[base + i];
            const float4 a = info[3 * (size_t)ti[j]], b = info[3 * (size_t)ti[j] + 1], c = info[3 * (size_t)ti[j] + 2];
            const float cen = P.axis == 0 ? a.x : (P.axis == 1 ? a.y : a.z);
            const bool inA = cen < P.pos; // bvh.c:245
            float* g = inA ? lb : rb;
            g[0] = fmn(g[0], a.w); g[1] = fmn(g[1], b.x); g[2] = fmn(g[2], b.y);
            g[3] = fmx(g[3], b.z); g[4] = fmx(g[4], b.w); g[5] = fmx(g[5], c.x);
            if (inA) left_mask |= 1u << j; else n_right_mine++;
        }
    }
    const int mine = __popc(left_mask);
    s[threadIdx.x] = mine;
    __syncthreads();
    for (int d = 1; d < kCta; d <<= 1) {
        const int add = threadIdx.x >= d ? s[threadIdx.x - d] : 0;
        __syncthreads();
        s[threadIdx.x] += add;
        __syncthreads();
    }
    int rank = chunk_off[blockIdx.x] + s[threadIdx.x] - mine;
    for (int j = 0; j < kPer; j++) {
        if (left_mask & (1u << j)) {
            dst[P.first + rank] = ti[j];
            posL[P.first + rank] = rel0 + threadIdx.x * kPer + j;
            rank++;
        }
    }
    if (mine) for (int a = 0; a < 3; a++) { atomic_min_f(&bx[0][a], lb[a]); atomic_max_f(&bx[0][3 + a], lb[3 + a]); }
    if (n_right_mine) for (int a = 0; a < 3; a++) { atomic_min_f(&bx[1][a], rb[a]); atomic_max_f(&bx[1][3 + a], rb[3 + a]); }
    __syncthreads();
    if (threadIdx.x < 12) {
        const int k = threadIdx.x / 6, a = threadIdx.x % 6;
        BNode& C = nodes[P.child + k];
        if (a < 3) atomic_min_f(&C.mn[a], bx[k][a]); else atomic_max_f(&C.mx[a - 3], bx[k][a]);
    }
}

// rights: follow the chain x -> posL[x] until it leaves the left region
__global__ void __launch_bounds__(kCta) rights_kernel(const int2* __restrict__ chunks, const int* __restrict__ active_ids, const BNode* __restrict__ nodes,
                                                      const int* __restrict__ tri_idx, const float4* __restrict__ info, const int* __restrict__ posL,
                                                      int* __restrict__ dst, Flags* flags)
{
    const int slot = chunks[blockIdx.x].x, chunk = chunks[blockIdx.x].y;
    const BNode& P = nodes[active_ids[slot]];
    const int rel0 = chunk * kChunk, cnt = min(kChunk, P.len - rel0);
    for (int i = threadIdx.x; i < cnt; i += kCta) {
        const int t = tri_idx[P.first + rel0 + i];
        if (centroid_of(info, t, P.axis) < P.pos) continue;
        int p = rel0 + i, steps = 0;
        while (p < P.nl) {
            p = posL[P.first + p];
            if (++steps > kChainCap) { flags->chain_overflow = 1; break; }
        }
        if (p >= P.nl) dst[P.first + p] = t;
    }
}

__global__ void __launch_bounds__(kCta) copyback_kernel(const int2* __restrict__ chunks, const int* __restrict__ active_ids, const BNode* __restrict__ nodes,
                                                        int* __restrict__ tri_idx, const int* __restrict__ dst)
{
    const int slot = chunks[blockIdx.x].x, chunk = chunks[blockIdx.x].y;
    const BNode& P = nodes[active_ids[slot]];
    const int base = P.first + chunk * kChunk, cnt = min(kChunk, P.len - chunk * kChunk);
    for (int i = threadIdx.x; i < cnt; i += kCta) tri_idx[base + i] = dst[base + i];
}

// ---- one WARP per small subtree: the reference algorithm node by node in its own (DFS) order, lanes cooperating ----
// Per node: lanes stride over the triangles for the binning pass (bins in the warp's slice of shared memory), lanes 0-2
// evaluate one axis each, the partition is the same compaction + chain-following as in the top levels (ballot instead
// of a block scan), child boxes by a shuffle reduction.  Nodes are written in the order the reference allocates them.
struct SubRoot { int bfs; int region; int cap; int final_root; int base; int pad[3]; };
constexpr int kSubWarps = 8; // warps per CTA

__global__ void __launch_bounds__(kSubWarps * 32) subtree_kernel(int n_sub, const SubRoot* __restrict__ roots, const BNode* __restrict__ nodes, int* tri_idx,
                                                                 int* tmp, int* posL, const float4* __restrict__ info, rt_bvh_node* region,
                                                                 unsigned char* region_depth, int* __restrict__ used, int refbin, Flags* flags, int* next_sub)
{
    __shared__ int s_cnt[kSubWarps][3 * kNB];
    __shared__ float s_mn[kSubWarps][3 * kNB * 3], s_mx[kSubWarps][3 * kNB * 3];
    __shared__ float s_split[kSubWarps][3][kBins];
    __shared__ int s_stack[kSubWarps][2 * (kMaxDepth + 2)];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    int* cnt = s_cnt[w];
    float* bmn = s_mn[w];
    float* bmx = s_mx[w];
    for (;;) {
        int k = 0;
        if (lane == 0) k = atomicAdd(next_sub, 1);
        k = __shfl_sync(FULL, k, 0);
        if (k >= n_sub) break;
        const BNode& R = nodes[roots[k].bfs];
        rt_bvh_node* T = region + roots[k].region;
        unsigned char* Tdepth = region_depth + roots[k].region;
        const int cap = roots[k].cap;
        if (lane == 0) {
            Tdepth[0] = (unsigned char)R.depth;
            rt_bvh_node r;
            for (int a = 0; a < 3; a++) { r.min[a] = R.mn[a]; r.max[a] = R.mx[a]; }
            r.tr_len = R.len; r.idx = R.first;
            T[0] = r;
            s_stack[w][0] = 0; s_stack[w][1] = R.depth;
        }
        __syncwarp();
        int len = 1, sp = 1; // warp-uniform
        bool overflow = false;
        while (sp) {
            sp--;
            const int node_idx = s_stack[w][2 * sp], depth = s_stack[w][2 * sp + 1];
            const rt_bvh_node parent = T[node_idx];
            __syncwarp();
            if (depth == kMaxDepth || parent.tr_len <= kLeafThreshold) { // bvh.c:84
                if (!parent.tr_len && lane == 0) T[node_idx].idx = 0;     // bvh.c:85-86
                continue;
            }
            if (len + 2 > cap) { overflow = true; break; }
            const int child_idx = len; // bvh.c:98-99
            len += 2;
            const int first = parent.idx, n_el = parent.tr_len;
            // ---- binning pass, all three axes ----
            for (int t = lane; t < 3 * kBins; t += 32) {
                const int axis = t / kBins, i = t % kBins;
                s_split[w][axis][i] = split_plane(parent.min[axis], parent.max[axis] - parent.min[axis], i, refbin);
            }
            for (int b = lane; b < 3 * kNB; b += 32) {
                cnt[b] = 0;
                for (int a = 0; a < 3; a++) { bmn[3 * b + a] = INFINITY; bmx[3 * b + a] = -INFINITY; }
            }
            __syncwarp();
            for (int i = lane; i < n_el; i += 32) {
                const int ti = tri_idx[first + i];
                const float4 a = info[3 * (size_t)ti], b = info[3 * (size_t)ti + 1], c = info[3 * (size_t)ti + 2];
                const float cen[3] = {a.x, a.y, a.z}, mn[3] = {a.w, b.x, b.y}, mx[3] = {b.z, b.w, c.x};
                for (int axis = 0; axis < 3; axis++) {
                    const int kb = axis * kNB + threshold_bin(s_split[w][axis], cen[axis]);
                    atomicAdd(&cnt[kb], 1);
                    for (int d = 0; d < 3; d++) { atomic_min_f(&bmn[3 * kb + d], mn[d]); atomic_max_f(&bmx[3 * kb + d], mx[d]); }
                }
            }
            __syncwarp();
            // ---- candidates (bvh.c:138-177): lane i evaluates plane i of one axis at a time.  Prefix / suffix unions of the
            // bins by warp scans (min / max / integer sums: exact in any order); the first strictly smaller cost in
            // axis-major, plane-ascending order wins = lexicographic minimum of (cost, plane) among the costs < FLT_MAX ----
            int splitAxis = 0;
            float splitPos = 0.f, best = FLT_MAX;
            for (int axis = 0; axis < 3; axis++) {
                const int bi = axis * kNB + lane;
                float pm[3], px[3], sm[3], sx[3];
                int pc = cnt[bi], sc_ = pc;
                for (int a = 0; a < 3; a++) { pm[a] = sm[a] = bmn[3 * bi + a]; px[a] = sx[a] = bmx[3 * bi + a]; }
                for (int o = 1; o < 32; o <<= 1) { // inclusive prefix over bins 0..lane, inclusive suffix over bins lane..31
                    const int pcu = __shfl_up_sync(FULL, pc, o), scd = __shfl_down_sync(FULL, sc_, o);
                    float pmu[3], pxu[3], smd[3], sxd[3];
                    for (int a = 0; a < 3; a++) {
                        pmu[a] = __shfl_up_sync(FULL, pm[a], o); pxu[a] = __shfl_up_sync(FULL, px[a], o);
                        smd[a] = __shfl_down_sync(FULL, sm[a], o); sxd[a] = __shfl_down_sync(FULL, sx[a], o);
                    }
                    if (lane >= o) { pc += pcu; for (int a = 0; a < 3; a++) { pm[a] = fmn(pm[a], pmu[a]); px[a] = fmx(px[a], pxu[a]); } }
                    if (lane + o < 32) { sc_ += scd; for (int a = 0; a < 3; a++) { sm[a] = fmn(sm[a], smd[a]); sx[a] = fmx(sx[a], sxd[a]); } }
                }
                // R_i = bins i+1 .. 32: the suffix of lane i+1 (none for lane 31) joined with bin 32
                const int b32 = axis * kNB + kBins;
                int cr = __shfl_down_sync(FULL, sc_, 1);
                float rm[3], rx[3];
                for (int a = 0; a < 3; a++) { rm[a] = __shfl_down_sync(FULL, sm[a], 1); rx[a] = __shfl_down_sync(FULL, sx[a], 1); }
                if (lane == 31) { cr = 0; for (int a = 0; a < 3; a++) { rm[a] = INFINITY; rx[a] = -INFINITY; } }
                cr += cnt[b32];
                for (int a = 0; a < 3; a++) { rm[a] = fmn(rm[a], bmn[3 * b32 + a]); rx[a] = fmx(rx[a], bmx[3 * b32 + a]); }
                float al_mn[3], al_mx[3], ar_mn[3], ar_mx[3]; // candidate boxes start at min = FLT_MAX, max = FLT_MIN (bvh.c:149-150)
                for (int a = 0; a < 3; a++) {
                    al_mn[a] = fmn(FLT_MAX, pm[a]); al_mx[a] = fmx(FLT_MIN, px[a]);
                    ar_mn[a] = fmn(FLT_MAX, rm[a]); ar_mx[a] = fmx(FLT_MIN, rx[a]);
                }
                float score;
                if (refbin) score = __fmaf_rn((float)pc, diag2(al_mn, al_mx, 1), (float)cr * diag2(ar_mn, ar_mx, 1));
                else score = (float)pc * diag2(al_mn, al_mx, 0) + (float)cr * diag2(ar_mn, ar_mx, 0); // bvh.c:169
                float key = score < FLT_MAX ? score : INFINITY; // NaN (0 * inf) and costs that can never win drop out
                int idx = lane;
                for (int o = 16; o; o >>= 1) {
                    const float ko = __shfl_xor_sync(FULL, key, o);
                    const int io = __shfl_xor_sync(FULL, idx, o);
                    if (ko < key || (ko == key && io < idx)) { key = ko; idx = io; }
                }
                if (key < best) { best = key; splitAxis = axis; splitPos = split_plane(parent.min[axis], parent.max[axis] - parent.min[axis], idx, refbin); }
            }
            // ---- partition (bvh.c:244-259): lefts compacted in encounter order, rights by their chains ----
            float lb[6] = {1e10f, 1e10f, 1e10f, -1e10f, -1e10f, -1e10f}, rb[6] = {1e10f, 1e10f, 1e10f, -1e10f, -1e10f, -1e10f}; // bvh.c:104-108
            int nl = 0;
            for (int base = 0; base < n_el; base += 32) {
                const int i = base + lane;
                bool inA = false;
                int ti = 0;
                if (i < n_el) {
                    ti = tri_idx[first + i];
                    const float4 a = info[3 * (size_t)ti], b = info[3 * (size_t)ti + 1], c = info[3 * (size_t)ti + 2];
                    const float cen = splitAxis == 0 ? a.x : (splitAxis == 1 ? a.y : a.z);
                    inA = cen < splitPos;
                    float* g = inA ? lb : rb;
                    g[0] = fmn(g[0], a.w); g[1] = fmn(g[1], b.x); g[2] = fmn(g[2], b.y);
                    g[3] = fmx(g[3], b.z); g[4] = fmx(g[4], b.w); g[5] = fmx(g[5], c.x);
                }
                const unsigned m = __ballot_sync(FULL, inA);
                if (inA) {
                    const int r = nl + __popc(m & ((1u << lane) - 1u));
                    tmp[first + r] = ti;
                    posL[first + r] = i;
                }
                nl += __popc(m);
            }
            __syncwarp();
            for (int i = lane; i < n_el; i += 32) {
                const int ti = tri_idx[first + i];
                const float4 a = info[3 * (size_t)ti];
                const float cen = splitAxis == 0 ? a.x : (splitAxis == 1 ? a.y : a.z);
                if (cen < splitPos) continue;
                int p = i, steps = 0; // a chain visits every position at most once: <= n_el steps
                while (p < nl && steps++ <= n_el) p = posL[first + p];
                if (p >= nl) tmp[first + p] = ti; else flags->chain_overflow = 1;
            }
            __syncwarp();
            for (int i = lane; i < n_el; i += 32) tri_idx[first + i] = tmp[first + i];
            for (int d = 0; d < 3; d++) {
                for (int o = 16; o; o >>= 1) {
                    lb[d] = fmn(lb[d], __shfl_xor_sync(FULL, lb[d], o)); lb[3 + d] = fmx(lb[3 + d], __shfl_xor_sync(FULL, lb[3 + d], o));
                    rb[d] = fmn(rb[d], __shfl_xor_sync(FULL, rb[d], o)); rb[3 + d] = fmx(rb[3 + d], __shfl_xor_sync(FULL, rb[3 + d], o));
                }
            }
            if (lane == 0) {
                rt_bvh_node left, right;
                left.tr_len = nl; left.idx = first; right.tr_len = n_el - nl; right.idx = first + nl;
                for (int a = 0; a < 3; a++) { left.min[a] = lb[a]; left.max[a] = lb[3 + a]; right.min[a] = rb[a]; right.max[a] = rb[3 + a]; }
                T[child_idx] = left;
                T[child_idx + 1] = right;
                Tdepth[child_idx] = Tdepth[child_idx + 1] = (unsigned char)(depth + 1);
                T[node_idx].idx = child_idx; // bvh.c:262-263
                T[node_idx].tr_len = 0;
                s_stack[w][2 * sp] = child_idx + 1; s_stack[w][2 * sp + 1] = depth + 1; // right is popped after the whole left subtree
                s_stack[w][2 * sp + 2] = child_idx; s_stack[w][2 * sp + 3] = depth + 1;
            }
            sp += 2;
            __syncwarp();
        }
        if (lane == 0) {
            used[k] = len;
            if (overflow) flags->region_overflow = 1;
        }
        __syncwarp();
    }
}

// ---- numbering: subtrees and top nodes into the reference's node order ----
__global__ void assemble_subtrees_kernel(const SubRoot* __restrict__ roots, const rt_bvh_node* __restrict__ region,
                                         const unsigned char* __restrict__ region_depth, const int* __restrict__ used,
                                         rt_bvh_node* __restrict__ out, unsigned char* __restrict__ out_depth)
{
    const SubRoot r = roots[blockIdx.x];
    const rt_bvh_node* T = region + r.region;
    const int m = used[blockIdx.x];
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        rt_bvh_node nd = T[j];
        if (nd.tr_len == 0 && nd.idx != 0) nd.idx = r.base + (nd.idx - 1);
        const int f = j == 0 ? r.final_root : r.base + j - 1;
        out[f] = nd;
        out_depth[f] = region_depth[r.region + j];
    }
}

struct TopRec { int f; int depth; rt_bvh_node nd; };
__global__ void scatter_top_kernel(const TopRec* __restrict__ recs, int n, rt_bvh_node* __restrict__ out, unsigned char* __restrict__ out_depth)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { out[recs[i].f] = recs[i].nd; out_depth[recs[i].f] = (unsigned char)recs[i].depth; }
}

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { if (p) { cudaFree(p); p = nullptr; } return cudaMalloc(&p, std::max<size_t>(bytes, 16)); }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

} // namespace

extern "C" int rt_scene_build_bvh_gpu(rt_scene* s, int heuristic, int device, rt_bvh_gpu_stats* stats)
{
    return rt::guarded("rt_scene_build_bvh_gpu", [&]() -> int {
    if (stats) std::memset(stats, 0, sizeof *stats);
    if (!s) { rt::set_error("rt_scene_build_bvh_gpu: null scene"); return RT_ERR_INVALID; }
    if ((heuristic & ~RT_BVH_REFBIN) != 6) { rt::set_error("rt_scene_build_bvh_gpu: only heuristic 6 is built on the GPU (use rt_scene_build_bvh for 0 / 1)"); return RT_ERR_INVALID; }
    return rt::gpu_build_bvh(*s, (heuristic & RT_BVH_REFBIN) ? 1 : 0, device, stats, nullptr);
    });
}

int rt::gpu_build_bvh(rt_scene& scene, int refbin, int device, rt_bvh_gpu_stats* stats, GpuTree* keep)
{
    rt_scene* s = &scene;
    rt_bvh_gpu_stats st;
    std::memset(&st, 0, sizeof st);
    if (stats) *stats = st;
    const size_t n = s->n_tris();
    if (n == 0) { rt::set_error("no triangles, cannot build bvh"); return RT_ERR_INVALID; } // bvh.c:361-364
    if (n >= (1u << 27)) { rt::set_error("more than 2^27 triangles"); return RT_ERR_INVALID; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); rt::set_error("rt_scene_build_bvh_gpu: no CUDA device"); return RT_ERR_NO_DEVICE; }
    if (device < 0 || device >= ndev) { rt::set_error("rt_scene_build_bvh_gpu: device index out of range"); return RT_ERR_INVALID; }

    std::string err;
#define CKB(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rt::set_error(std::string("rt_scene_build_bvh_gpu: ") + #call + " failed: " + cudaGetErrorString(e__)); return RT_ERR_CUDA; } } while (0)
    const double t_begin = now_ms();
    CKB(cudaSetDevice(device));
    CKB(cudaFree(nullptr)); // context creation is not part of the build
    const bool dbg = std::getenv("RT_BVH_GPU_DEBUG") != nullptr;
    auto mark = [&](const char* what) { if (dbg) { cudaDeviceSynchronize(); std::fprintf(stderr, "[bvh_gpu] %-28s %9.2f ms\n", what, now_ms() - t_begin); } };
    mark("context");
    cudaStream_t stream = nullptr; // legacy default stream: every step below is ordered, the host waits where it reads back

    DevBuf d_tri, d_info, d_idx, d_tmp, d_posl, d_nodes, d_bins, d_flags, d_chunks, d_active, d_chunk_base, d_chunk_nl, d_chunk_off;
    int kSmall = kSmallDefault;
    if (const char* e = std::getenv("RT_BVH_GPU_SMALL")) { const int v = std::atoi(e); if (v >= 3 && v <= (1 << 20)) kSmall = v; }
    const size_t max_active = n / (size_t)kSmall + 2;
    const size_t max_top_nodes = 4 * max_active + 64; // first allocation; the node array grows on demand (lopsided splits)
    CKB(d_tri.alloc(n * 36));
    CKB(d_info.alloc(n * 48));
    CKB(d_idx.alloc(n * 4));
    CKB(d_tmp.alloc(n * 4));
    CKB(d_posl.alloc(n * 4));
    CKB(d_bins.alloc(max_active * sizeof(BinSet)));
    CKB(d_flags.alloc(sizeof(Flags)));
    mark("allocations");
    CKB(cudaMemsetAsync(d_flags.p, 0, sizeof(Flags), stream));
    CKB(rt::staged_h2d(d_tri.p, s->tri.data(), n * 36, stream)); // pinned ring, PCIe rate (staged_copy.h)
    mark("triangles host->device");
    prepare_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_tri.as<float>(), d_info.as<float4>(), d_idx.as<int>(), (int)n, refbin);
    CKB(cudaGetLastError());
    mark("prepare kernel");

    // top nodes live in a host mirror + device array that grows level by level
    std::vector<BNode> top;
    top.reserve(1024);
    size_t top_cap = std::min<size_t>(std::max<size_t>(max_top_nodes, 1024), 2 * n + 64);
    // depth-32 chains of lopsided splits can create more top nodes than n / kSmall * 4; grow on demand
    CKB(d_nodes.alloc(top_cap * sizeof(BNode)));
    {
        BNode root;
        std::memset(&root, 0, sizeof root);
        for (int a = 0; a < 3; a++) { root.mn[a] = 1e10f; root.mx[a] = -1e10f; }
        root.first = 0; root.len = (int)n; root.depth = 0; root.child = -1;
        CKB(cudaMemcpyAsync(d_nodes.p, &root, sizeof root, cudaMemcpyHostToDevice, stream));
        root_box_kernel<<<148 * 4, 256, 0, stream>>>(d_info.as<float4>(), (int)n, d_nodes.as<BNode>());
        CKB(cudaGetLastError());
        BNode back;
        CKB(cudaMemcpy(&back, d_nodes.p, sizeof back, cudaMemcpyDeviceToHost));
        top.push_back(back);
    }
    st.upload_ms = (float)(now_ms() - t_begin);

    const double t_top = now_ms();
    std::vector<int> active, small_roots;
    auto classify = [&](int id) {
        const BNode& b = top[(size_t)id];
        if (b.depth == kMaxDepth || b.len <= kLeafThreshold) return; // leaf (bvh.c:84)
        if (b.len >= kSmall) active.push_back(id); else small_roots.push_back(id);
    };
    classify(0);
    std::vector<int2> chunks;
    std::vector<int> chunk_base;
    size_t chunk_cap = 0, active_cap = 0;
    while (!active.empty()) {
        const int n_active = (int)active.size();
        if ((size_t)n_active > max_active) { rt::set_error("rt_scene_build_bvh_gpu: internal: too many active nodes"); return RT_ERR_STATE; }
        chunks.clear();
        chunk_base.assign((size_t)n_active + 1, 0);
        for (int a = 0; a < n_active; a++) {
            const int nc = (top[(size_t)active[a]].len + kChunk - 1) / kChunk;
            chunk_base[(size_t)a] = (int)chunks.size();
            for (int c = 0; c < nc; c++) chunks.push_back(make_int2(a, c));
        }
        chunk_base[(size_t)n_active] = (int)chunks.size();
        const int n_chunks = (int)chunks.size();
        if ((size_t)n_chunks > chunk_cap) {
            chunk_cap = (size_t)n_chunks * 2;
            CKB(d_chunks.alloc(chunk_cap * sizeof(int2)));
            CKB(d_chunk_nl.alloc(chunk_cap * 4));
            CKB(d_chunk_off.alloc(chunk_cap * 4));
        }
        if ((size_t)n_active + 1 > active_cap) {
            active_cap = ((size_t)n_active + 1) * 2;
            CKB(d_active.alloc(active_cap * 4));
            CKB(d_chunk_base.alloc(active_cap * 4));
        }
        const int child_base = (int)top.size();
        if (top.size() + 2 * (size_t)n_active > top_cap) { // grow the device node array, keeping its contents
            const size_t new_cap = (top.size() + 2 * (size_t)n_active) * 2;
            DevBuf bigger;
            CKB(bigger.alloc(new_cap * sizeof(BNode)));
            CKB(cudaMemcpyAsync(bigger.p, d_nodes.p, top.size() * sizeof(BNode), cudaMemcpyDeviceToDevice, stream));
            CKB(cudaStreamSynchronize(stream));
            std::swap(bigger.p, d_nodes.p);
            top_cap = new_cap;
        }
        CKB(cudaMemcpyAsync(d_chunks.p, chunks.data(), (size_t)n_chunks * sizeof(int2), cudaMemcpyHostToDevice, stream));
        CKB(cudaMemcpyAsync(d_active.p, active.data(), (size_t)n_active * 4, cudaMemcpyHostToDevice, stream));
        CKB(cudaMemcpyAsync(d_chunk_base.p, chunk_base.data(), ((size_t)n_active + 1) * 4, cudaMemcpyHostToDevice, stream));
        BNode* nodes = d_nodes.as<BNode>();
        init_bins_kernel<<<std::min(n_active * 3 + 1, 148 * 8), 128, 0, stream>>>(d_bins.as<BinSet>(), n_active);
        bin_kernel<<<n_chunks, kCta, 0, stream>>>(d_chunks.as<int2>(), d_active.as<int>(), nodes, d_idx.as<int>(), d_info.as<float4>(), d_bins.as<BinSet>(), refbin);
        choose_kernel<<<(n_active + 63) / 64, 64, 0, stream>>>(n_active, d_active.as<int>(), nodes, d_bins.as<BinSet>(), child_base, refbin);
        count_kernel<<<n_chunks, kCta, 0, stream>>>(d_chunks.as<int2>(), d_active.as<int>(), nodes, d_idx.as<int>(), d_info.as<float4>(), d_chunk_nl.as<int>());
        scan_kernel<<<n_active, kCta, 0, stream>>>(d_chunk_base.as<int>(), d_active.as<int>(), nodes, d_chunk_nl.as<int>(), d_chunk_off.as<int>());
        scatter_kernel<<<n_chunks, kCta, 0, stream>>>(d_chunks.as<int2>(), d_active.as<int>(), nodes, d_idx.as<int>(), d_info.as<float4>(), d_chunk_off.as<int>(),
                                                      d_tmp.as<int>(), d_posl.as<int>());
        rights_kernel<<<n_chunks, kCta, 0, stream>>>(d_chunks.as<int2>(), d_active.as<int>(), nodes, d_idx.as<int>(), d_info.as<float4>(), d_posl.as<int>(),
                                                     d_tmp.as<int>(), d_flags.as<Flags>());
        copyback_kernel<<<n_chunks, kCta, 0, stream>>>(d_chunks.as<int2>(), d_active.as<int>(), nodes, d_idx.as<int>(), d_tmp.as<int>());
        CKB(cudaGetLastError());
        // read back the parents (axis, pos, child) and the new children
        top.resize(top.size() + 2 * (size_t)n_active);
        CKB(cudaMemcpyAsync(top.data() + child_base, nodes + child_base, 2 * (size_t)n_active * sizeof(BNode), cudaMemcpyDeviceToHost, stream));
        CKB(cudaStreamSynchronize(stream));
        for (int a = 0; a < n_active; a++) top[(size_t)active[a]].child = child_base + 2 * a;
        std::vector<int> prev;
        prev.swap(active);
        for (int a = 0; a < (int)prev.size(); a++) { classify(child_base + 2 * a); classify(child_base + 2 * a + 1); }
        st.levels++;
    }
    mark("top levels");
    st.top_ms = (float)(now_ms() - t_top);
    st.top_nodes = (int)top.size();

    // ---- small subtrees: one thread each ----
    const double t_sub = now_ms();
    const int n_sub = (int)small_roots.size();
    // largest first: the warps draw subtrees from a counter, so the long ones should not start last
    std::stable_sort(small_roots.begin(), small_roots.end(), [&](int a, int b) { return top[(size_t)a].len > top[(size_t)b].len; });
    std::vector<SubRoot> roots((size_t)n_sub);
    size_t region_total = 0;
    for (int k = 0; k < n_sub; k++) {
        SubRoot r;
        std::memset(&r, 0, sizeof r);
        r.bfs = small_roots[(size_t)k];
        r.region = (int)region_total;
        r.cap = 2 * top[(size_t)r.bfs].len + 64;
        region_total += (size_t)r.cap;
        roots[(size_t)k] = r;
    }
    if (region_total >= (size_t)INT32_MAX) { rt::set_error("rt_scene_build_bvh_gpu: subtree regions exceed 2^31 nodes"); return RT_ERR_NOMEM; }
    DevBuf d_roots, d_region, d_region_depth, d_used, d_out, d_out_depth, d_recs;
    std::vector<int> used((size_t)n_sub, 1);
    if (n_sub) {
        CKB(d_roots.alloc((size_t)n_sub * sizeof(SubRoot)));
        CKB(d_region.alloc(region_total * sizeof(rt_bvh_node)));
        CKB(d_region_depth.alloc(region_total));
        CKB(d_used.alloc((size_t)n_sub * 4));
        CKB(cudaMemcpyAsync(d_roots.p, roots.data(), (size_t)n_sub * sizeof(SubRoot), cudaMemcpyHostToDevice, stream));
        DevBuf d_next;
        CKB(d_next.alloc(4));
        CKB(cudaMemsetAsync(d_next.p, 0, 4, stream));
        const int sub_ctas = std::min((n_sub + kSubWarps - 1) / kSubWarps, 148 * 8);
        subtree_kernel<<<sub_ctas, kSubWarps * 32, 0, stream>>>(n_sub, d_roots.as<SubRoot>(), d_nodes.as<BNode>(), d_idx.as<int>(), d_tmp.as<int>(), d_posl.as<int>(),
                                                              d_info.as<float4>(), d_region.as<rt_bvh_node>(), d_region_depth.as<unsigned char>(), d_used.as<int>(), refbin,
                                                              d_flags.as<Flags>(), d_next.as<int>());
        CKB(cudaGetLastError());
        CKB(cudaStreamSynchronize(stream)); // d_next is released at the end of this scope
        CKB(cudaGetLastError());
        CKB(cudaMemcpyAsync(used.data(), d_used.p, (size_t)n_sub * 4, cudaMemcpyDeviceToHost, stream));
    }
    Flags fl;
    CKB(cudaMemcpyAsync(&fl, d_flags.p, sizeof fl, cudaMemcpyDeviceToHost, stream));
    CKB(cudaStreamSynchronize(stream));
    mark("subtrees");
    st.subtree_ms = (float)(now_ms() - t_sub);
    st.subtrees = n_sub;

    // ---- reference numbering (as Builder::run_parallel): children pairwise when the parent is split, left subtree first ----
    const double t_asm = now_ms();
    size_t total = 0;
    for (const BNode& b : top) (void)b, total++;
    for (int k = 0; k < n_sub; k++) total += (size_t)used[(size_t)k] - 1;
    const bool degenerate = fl.chain_overflow || fl.region_overflow || total >= 2 * n; // bvh.c:80-83 would have stopped the build
    if (degenerate) {
        st.fell_back = 1;
        st.total_ms = (float)(now_ms() - t_begin);
        if (stats) *stats = st;
        if (keep) keep->fell_back = true;
        return rt::build_bvh(*s, 6, refbin ? rt::BVH_REFBIN : rt::BVH_IEEE, 1); // serial: it carries the reference's 2N guard
    }
    std::vector<int> task_of(top.size(), -1);
    for (int k = 0; k < n_sub; k++) task_of[(size_t)roots[(size_t)k].bfs] = k;
    std::vector<TopRec> recs;
    recs.reserve(top.size());
    {
        int counter = 1;
        struct Item { int t, f; };
        std::vector<Item> stk{{0, 0}};
        while (!stk.empty()) {
            const Item it = stk.back();
            stk.pop_back();
            const BNode& b = top[(size_t)it.t];
            const int k = task_of[(size_t)it.t];
            if (k >= 0) {
                roots[(size_t)k].final_root = it.f;
                roots[(size_t)k].base = counter;
                counter += used[(size_t)k] - 1;
                continue;
            }
            TopRec r;
            r.f = it.f;
            r.depth = b.depth;
            for (int a = 0; a < 3; a++) { r.nd.min[a] = b.mn[a]; r.nd.max[a] = b.mx[a]; }
            if (b.child >= 0) {
                r.nd.tr_len = 0;
                r.nd.idx = counter;
                stk.push_back({b.child + 1, counter + 1}); // right pushed first: the whole left subtree is numbered before it
                stk.push_back({b.child, counter});
                counter += 2;
            } else {
                r.nd.tr_len = b.len;
                r.nd.idx = b.len ? b.first : 0; // bvh.c:85-86
            }
            recs.push_back(r);
        }
        total = (size_t)counter;
    }
    CKB(d_out.alloc(total * sizeof(rt_bvh_node)));
    CKB(d_out_depth.alloc(total));
    CKB(d_recs.alloc(recs.size() * sizeof(TopRec)));
    CKB(cudaMemcpyAsync(d_recs.p, recs.data(), recs.size() * sizeof(TopRec), cudaMemcpyHostToDevice, stream));
    if (!recs.empty())
        scatter_top_kernel<<<(unsigned)((recs.size() + 255) / 256), 256, 0, stream>>>(d_recs.as<TopRec>(), (int)recs.size(), d_out.as<rt_bvh_node>(),
                                                                                      d_out_depth.as<unsigned char>());
    if (n_sub) {
        CKB(cudaMemcpyAsync(d_roots.p, roots.data(), (size_t)n_sub * sizeof(SubRoot), cudaMemcpyHostToDevice, stream));
        assemble_subtrees_kernel<<<n_sub, 128, 0, stream>>>(d_roots.as<SubRoot>(), d_region.as<rt_bvh_node>(), d_region_depth.as<unsigned char>(), d_used.as<int>(),
                                                            d_out.as<rt_bvh_node>(), d_out_depth.as<unsigned char>());
    }
    CKB(cudaGetLastError());
    CKB(cudaStreamSynchronize(stream));
    mark("assemble");
    st.assemble_ms = (float)(now_ms() - t_asm);

    const double t_down = now_ms();
    if (keep) {
        // the tree stays on the device: hand the four arrays over, nothing visits the host
        keep->release();
        keep->device = device; keep->n_tris = n; keep->n_nodes = total; keep->fell_back = false;
        keep->tri = d_tri.as<float>(); d_tri.p = nullptr;
        keep->tri_idx = d_idx.as<int>(); d_idx.p = nullptr;
        keep->nodes = d_out.as<rt_bvh_node>(); d_out.p = nullptr;
        keep->depth = d_out_depth.as<unsigned char>(); d_out_depth.p = nullptr;
    } else {
        s->bvh.resize(total);
        s->tri_idx.resize(n);
        CKB(cudaStreamSynchronize(stream));
        CKB(rt::staged_d2h(s->bvh.data(), d_out.p, total * sizeof(rt_bvh_node), stream));
        CKB(rt::staged_d2h(s->tri_idx.data(), d_idx.p, n * 4, stream));
    }
    mark("download");
    st.download_ms = (float)(now_ms() - t_down);
    st.total_ms = (float)(now_ms() - t_begin);
    st.nodes = (unsigned)total;
    if (stats) *stats = st;
#undef CKB
    return RT_OK;
}

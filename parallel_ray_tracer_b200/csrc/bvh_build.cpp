// bvh_build.cpp — host BVH build that reproduces the reference tree node for node
// (bvh_build / bvh_split, cpu/src/bvh.c:78-267, 360-388) at O(n) per node instead of the
// reference's O(96 n).
//
// What is reproduced literally:
//   * node numbering: children are allocated as a pair (bvh_len, bvh_len+1) when the parent is
//     split, left subtree before right (bvh.c:98-99, 265-266);
//   * termination: depth == 32 or <= 2 triangles (bvh.c:84); an empty node gets idx = 0;
//   * child boxes grown from (+1e10, -1e10) by triangle vertices (bvh.c:104-108, 61-71);
//   * heuristic 6: 3 axes x 32 planes split_i = min + size * (i/32) (bvh.c:156-157), membership
//     centroid < split_i, candidate boxes initialised to min = FLT_MAX, max = FLT_MIN (positive!,
//     bvh.c:149-150), cost = cl * diag2(L) + cr * diag2(R) in float with diag2 the squared diagonal
//     (bvh.c:43-46, 169), first strictly smaller cost wins in axis-major / i-ascending order;
//     an empty side gives inf and 0 * inf = NaN, which never compares smaller;
//   * the in-place forward partition with swap-to-front (bvh.c:244-259) that fixes the order of
//     tri_idx inside leaves, hence the tie-break order of the traversal.
//
// Why O(n) gives the same bits: split_i is non-decreasing in i (rounding is monotone), so each
// triangle has one threshold bin k = min{ i : centroid < split_i }.  One pass accumulates per-bin
// counts and vertex bounds; prefix/suffix unions give every candidate's (cl, box_L, cr, box_R)
// exactly, because min/max are exact and order-independent.  The cost is then evaluated with the
// reference's float expression.
//
// Two arithmetic flavours (rt::BvhArith):
//   BVH_IEEE    the IEEE reading of the source — what the reference GPU program's host code does
//               (gpu/src/bvh.cu, compiled by the host compiler without fast-math);
//   BVH_REFBIN  the reference CPU program is compiled -O3 -ffast-math -march=native
//               (cpu/makefile:14); gcc 13 on FMA hardware turns four expressions into
//                   centroid = ((a + b) + c) * 0.33333334f
//                   split_i  = fmaf((float)i, size * 0.03125f, min)
//                   diag2    = fmaf(dz, dz, fmaf(dx, dx, dy * dy))
//                   cost     = fmaf(cl, diag2_L, cr * diag2_R)
//               which flips ~2 % of near-tied decisions.  With these the builder reproduces that
//               binary's bvh[] / tri_idx[] exactly (tests/test_bvh_build.py).
//
// This file must be compiled with -ffp-contract=off (csrc/Makefile): the compiler must not
// contract anything on its own.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include "host_scene.h"

namespace {

constexpr int kMaxDepth = 32;   // BVH_MAX_ITER, cpu/include/options.h:64
constexpr int kLeafThreshold = 2; // BVH_ELEMENT_THRESHOLD, cpu/include/options.h:58
constexpr int kBins = 32;       // SAH_BIN_SIZE, cpu/include/options.h:61

struct Box {
    float mn[3], mx[3];
};

struct Builder {
    rt_scene& s;
    int heuristic;
    rt::BvhArith arith;
    std::vector<float> centroid; // 3 per triangle
    std::vector<Box> tbox;       // vertex bounds per triangle
    int32_t bvh_len = 1;
    size_t n;

    Builder(rt_scene& sc, int h, rt::BvhArith a) : s(sc), heuristic(h), arith(a), n(sc.n_tris()) {}

    float diag2(const Box& b) const
    {
        const float dx = b.mx[0] - b.mn[0], dy = b.mx[1] - b.mn[1], dz = b.mx[2] - b.mn[2];
        if (arith == rt::BVH_REFBIN) return std::fmaf(dz, dz, std::fmaf(dx, dx, dy * dy));
        return dx * dx + dy * dy + dz * dz; // vec_dot(size, size), bvh.c:43-46
    }

    void prepare()
    {
        centroid.resize(3 * n);
        tbox.resize(n);
        const float* t = s.tri.data();
        for (size_t i = 0; i < n; i++, t += 9) {
            for (int a = 0; a < 3; a++) {
                const float p0 = t[a], p1 = t[3 + a], p2 = t[6 + a];
                centroid[3 * i + a] = (arith == rt::BVH_REFBIN) ? ((p0 + p1) + p2) * 0.33333334f
                                                                : (p0 + p1 + p2) / 3.0f; // triangle.c:21-23
                tbox[i].mn[a] = std::fmin(std::fmin(p0, p1), p2);
                tbox[i].mx[a] = std::fmax(std::fmax(p0, p1), p2);
            }
        }
    }

    // heuristic 6 (bvh.c:138-177) in one binning pass per axis
    void choose_h6(const rt_bvh_node& parent, int& splitAxis, float& splitPos) const
    {
        splitAxis = 0;
        splitPos = 0;
        float best = FLT_MAX;
        const int first = parent.idx, len = parent.tr_len;
        for (int axis = 0; axis < 3; axis++) {
            float split[kBins];
            const float size = parent.max[axis] - parent.min[axis];
            for (int i = 0; i < kBins; i++)
                split[i] = (arith == rt::BVH_REFBIN) ? std::fmaf((float)i, size * 0.03125f, parent.min[axis])
                                                     : parent.min[axis] + size * ((float)i / kBins);
            int cnt[kBins + 1] = {0};
            Box bin[kBins + 1];
            for (int b = 0; b <= kBins; b++)
                for (int a = 0; a < 3; a++) { bin[b].mn[a] = INFINITY; bin[b].mx[a] = -INFINITY; }
            for (int j = first; j < first + len; j++) {
                const int ti = s.tri_idx[j];
                const float c = centroid[3 * (size_t)ti + axis];
                // k = min{ i : c < split[i] } (32 if none); split[] is non-decreasing
                const int k = (int)(std::upper_bound(split, split + kBins, c) - split);
                cnt[k]++;
                const Box& tb = tbox[ti];
                Box& bb = bin[k];
                for (int a = 0; a < 3; a++) {
                    bb.mn[a] = std::fmin(bb.mn[a], tb.mn[a]);
                    bb.mx[a] = std::fmax(bb.mx[a], tb.mx[a]);
                }
            }
            // suffix unions: R_i = bins i+1 .. 32
            Box suf[kBins + 1];
            int sufc[kBins + 1];
            Box acc;
            for (int a = 0; a < 3; a++) { acc.mn[a] = INFINITY; acc.mx[a] = -INFINITY; }
            int accc = 0;
            for (int b = kBins; b >= 1; b--) {
                for (int a = 0; a < 3; a++) { acc.mn[a] = std::fmin(acc.mn[a], bin[b].mn[a]); acc.mx[a] = std::fmax(acc.mx[a], bin[b].mx[a]); }
                accc += cnt[b];
                suf[b - 1] = acc;
                sufc[b - 1] = accc;
            }
            Box pre;
            for (int a = 0; a < 3; a++) { pre.mn[a] = INFINITY; pre.mx[a] = -INFINITY; }
            int prec = 0;
            for (int i = 0; i < kBins; i++) {
                for (int a = 0; a < 3; a++) { pre.mn[a] = std::fmin(pre.mn[a], bin[i].mn[a]); pre.mx[a] = std::fmax(pre.mx[a], bin[i].mx[a]); }
                prec += cnt[i];
                Box al, ar; // candidate boxes start at min = FLT_MAX, max = FLT_MIN (bvh.c:149-150)
                for (int a = 0; a < 3; a++) {
                    al.mn[a] = std::fmin(FLT_MAX, pre.mn[a]);
                    al.mx[a] = std::fmax(FLT_MIN, pre.mx[a]);
                    ar.mn[a] = std::fmin(FLT_MAX, suf[i].mn[a]);
                    ar.mx[a] = std::fmax(FLT_MIN, suf[i].mx[a]);
                }
                const int cl = prec, cr = sufc[i];
                float score;
                if (arith == rt::BVH_REFBIN) score = std::fmaf((float)cl, diag2(al), (float)cr * diag2(ar));
                else score = cl * diag2(al) + cr * diag2(ar); // bvh.c:169
                if (score < best) { best = score; splitAxis = axis; splitPos = split[i]; }
            }
        }
    }

    void split(int node_idx, int depth)
    {
        rt_bvh_node* bvh = s.bvh.data();
        rt_bvh_node& parent = bvh[node_idx];
        if ((size_t)bvh_len >= 2 * n) {                                           // bvh.c:80-83
            // the reference returns without clearing an EMPTY node's union field, which then reads as a
            // dangling child index; clear it (only reachable with heuristics 0/1 on degenerate input)
            if (!parent.tr_len) parent.idx = 0;
            return;
        }
        if (depth == kMaxDepth || parent.tr_len <= kLeafThreshold) {              // bvh.c:84
            if (!parent.tr_len) parent.idx = 0;                                   // bvh.c:85-86
            return;
        }
        const int child_idx = bvh_len;                                            // bvh.c:98-99
        bvh_len += 2;
        rt_bvh_node& left = bvh[child_idx];
        rt_bvh_node& right = bvh[child_idx + 1];
        left.idx = parent.idx;
        right.idx = parent.idx;
        Box lb, rb;
        for (int a = 0; a < 3; a++) { lb.mn[a] = rb.mn[a] = 1e10f; lb.mx[a] = rb.mx[a] = -1e10f; } // bvh.c:104-108

        int splitAxis = 0;
        float splitPos = 0;
        if (heuristic == 6) {
            choose_h6(parent, splitAxis, splitPos);
        } else {
            // heuristics 0 / 1: spatial median (bvh.c:112-113, 214-223)
            float center[3], size[3];
            for (int a = 0; a < 3; a++) { center[a] = (parent.min[a] + parent.max[a]) * 0.5f; size[a] = parent.max[a] - parent.min[a]; }
            if (heuristic == 1) {
                if (size[1] > size[0]) splitAxis = 1;
                if (size[2] > size[0] && size[2] > size[1]) splitAxis = 2;
            }
            splitPos = center[splitAxis];
        }

        int32_t* tri_idx = s.tri_idx.data();
        for (int i = parent.idx; i < parent.idx + parent.tr_len; i++) {            // bvh.c:244-259
            const int t_idx = tri_idx[i];
            const bool inA = centroid[3 * (size_t)t_idx + splitAxis] < splitPos;
            Box& cb = inA ? lb : rb;
            const Box& tb = tbox[t_idx];
            for (int a = 0; a < 3; a++) { cb.mn[a] = std::fmin(cb.mn[a], tb.mn[a]); cb.mx[a] = std::fmax(cb.mx[a], tb.mx[a]); }
            if (inA) {
                left.tr_len += 1;
                const int swap = left.idx + left.tr_len - 1;
                tri_idx[i] = tri_idx[swap];
                tri_idx[swap] = t_idx;
                right.idx += 1;
            } else {
                right.tr_len += 1;
            }
        }
        std::memcpy(left.min, lb.mn, 12); std::memcpy(left.max, lb.mx, 12);
        std::memcpy(right.min, rb.mn, 12); std::memcpy(right.max, rb.mx, 12);
        parent.idx = child_idx;                                                   // bvh.c:262-263
        parent.tr_len = 0;
        split(child_idx, depth + 1);                                              // bvh.c:265-266
        split(child_idx + 1, depth + 1);
    }

    void run()
    {
        prepare();
        s.tri_idx.resize(n);
        for (size_t i = 0; i < n; i++) s.tri_idx[i] = (int32_t)i;                  // bvh.c:366-368
        // bvh.c:370-371 allocates 2N nodes, but the guard at bvh.c:80 lets a split begin at bvh_len == 2N-1
        // and write node 2N (reachable with empty children under heuristics 0/1): keep two spare nodes
        s.bvh.assign(2 * n + 2, rt_bvh_node{{0, 0, 0}, {0, 0, 0}, 0, 0});
        rt_bvh_node& root = s.bvh[0];
        root.tr_len = (int32_t)n;
        Box rb;
        for (int a = 0; a < 3; a++) { rb.mn[a] = 1e10f; rb.mx[a] = -1e10f; }       // bvh.c:373-377
        for (size_t i = 0; i < n; i++)
            for (int a = 0; a < 3; a++) { rb.mn[a] = std::fmin(rb.mn[a], tbox[i].mn[a]); rb.mx[a] = std::fmax(rb.mx[a], tbox[i].mx[a]); }
        std::memcpy(root.min, rb.mn, 12); std::memcpy(root.max, rb.mx, 12);
        split(0, 0);
        s.bvh.resize((size_t)bvh_len);
    }
};

} // namespace

namespace rt {

int build_bvh(rt_scene& s, int heuristic, BvhArith arith, int /*threads*/)
{
    if (s.n_tris() == 0) { set_error("no triangles, cannot build bvh"); return RT_ERR_INVALID; } // bvh.c:361-364
    if (heuristic != 6 && heuristic != 0 && heuristic != 1) {
        set_error("rt_scene_build_bvh: heuristic must be 6, 0 or 1 (2/3 depend on libc rand() and an out-of-bounds "
                  "axis, 4/5 on qsort tie order; see SURVEY.md C.1)");
        return RT_ERR_INVALID;
    }
    if (s.n_tris() >= (1u << 27)) { set_error("more than 2^27 triangles"); return RT_ERR_INVALID; }
    Builder b(s, heuristic, arith);
    b.run();
    return RT_OK;
}

} // namespace rt

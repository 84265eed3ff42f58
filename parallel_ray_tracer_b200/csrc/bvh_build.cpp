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
// Large scenes (>= 200 000 triangles, heuristic 6) are built in parallel: the top of the tree is split serially
// (its per-node binning is chunked over threads), subtrees below a size cut-off are built concurrently into
// private arrays, and a final pass assigns the reference's node numbers (children allocated pairwise in DFS
// order: a subtree's descendants are contiguous and keep their relative order) — same arrays as the serial build.
//
// This file must be compiled with -ffp-contract=off (csrc/Makefile): the compiler must not
// contract anything on its own.
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "host_scene.h"

namespace {

constexpr int kMaxDepth = 32;     // BVH_MAX_ITER, cpu/include/options.h:64
constexpr int kLeafThreshold = 2; // BVH_ELEMENT_THRESHOLD, cpu/include/options.h:58
constexpr int kBins = 32;         // SAH_BIN_SIZE, cpu/include/options.h:61

// fminf/fmaxf on finite inputs without the libm call (-fno-fast-math keeps gcc from inlining them)
inline float fmn(float a, float b) { return b < a ? b : a; }
inline float fmx(float a, float b) { return b > a ? b : a; }

struct Box {
    float mn[3], mx[3];
    void clear() { for (int a = 0; a < 3; a++) { mn[a] = INFINITY; mx[a] = -INFINITY; } }
    void grow(const Box& o) { for (int a = 0; a < 3; a++) { mn[a] = fmn(mn[a], o.mn[a]); mx[a] = fmx(mx[a], o.mx[a]); } }
};

struct Bins {
    int cnt[kBins + 1];
    Box box[kBins + 1];
    void clear() { for (int b = 0; b <= kBins; b++) { cnt[b] = 0; box[b].clear(); } }
};

// node storage with the reference's allocation counter (bvh_len)
struct Tree {
    std::vector<rt_bvh_node> nodes;
    int32_t len = 1;
};

struct Task { int32_t top_node; int depth; };

struct Builder {
    rt_scene& s;
    int heuristic;
    rt::BvhArith arith;
    int threads;
    std::vector<float> centroid; // 3 per triangle
    std::vector<Box> tbox;       // vertex bounds per triangle
    size_t n;

    Builder(rt_scene& sc, int h, rt::BvhArith a, int th) : s(sc), heuristic(h), arith(a), threads(th), n(sc.n_tris()) {}

    float diag2(const Box& b) const
    {
        const float dx = b.mx[0] - b.mn[0], dy = b.mx[1] - b.mn[1], dz = b.mx[2] - b.mn[2];
        if (arith == rt::BVH_REFBIN) return std::fmaf(dz, dz, std::fmaf(dx, dx, dy * dy));
        return dx * dx + dy * dy + dz * dz; // vec_dot(size, size), bvh.c:43-46
    }

    template <class F>
    void parallel_for(size_t count, size_t min_chunk, F f) const
    {
        int t = (int)std::min<size_t>((size_t)threads, count / min_chunk);
        if (t <= 1) { f(0, (size_t)0, count); return; }
        std::vector<std::thread> th;
        for (int i = 0; i < t; i++) th.emplace_back([=] { f(i, count * i / t, count * (i + 1) / t); });
        for (auto& x : th) x.join();
    }

    void prepare()
    {
        centroid.resize(3 * n);
        tbox.resize(n);
        parallel_for(n, 1 << 16, [&](int, size_t lo, size_t hi) {
            const float* t = s.tri.data() + 9 * lo;
            for (size_t i = lo; i < hi; i++, t += 9) {
                for (int a = 0; a < 3; a++) {
                    const float p0 = t[a], p1 = t[3 + a], p2 = t[6 + a];
                    centroid[3 * i + a] = (arith == rt::BVH_REFBIN) ? ((p0 + p1) + p2) * 0.33333334f
                                                                    : (p0 + p1 + p2) / 3.0f; // triangle.c:21-23
                    tbox[i].mn[a] = fmn(fmn(p0, p1), p2);
                    tbox[i].mx[a] = fmx(fmx(p0, p1), p2);
                }
            }
        });
    }

    void bin_range(const float* split, int axis, int lo, int hi, Bins& out) const
    {
        out.clear();
        const int32_t* tri_idx = s.tri_idx.data();
        for (int j = lo; j < hi; j++) {
            const int ti = tri_idx[j];
            const float c = centroid[3 * (size_t)ti + axis];
            // k = min{ i : c < split[i] } (32 if none); split[] is non-decreasing
            const int k = (int)(std::upper_bound(split, split + kBins, c) - split);
            out.cnt[k]++;
            out.box[k].grow(tbox[ti]);
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
            Bins bins;
            if (threads > 1 && len >= (1 << 18)) {
                // min/max and integer counts are exact and order independent: chunked binning gives the same bins
                std::vector<Bins> part((size_t)threads);
                parallel_for((size_t)len, 1 << 16, [&](int t, size_t lo, size_t hi) { bin_range(split, axis, first + (int)lo, first + (int)hi, part[t]); });
                bins.clear();
                const int t_used = (int)std::min<size_t>((size_t)threads, (size_t)len / (1 << 16));
                for (int t = 0; t < std::max(t_used, 1); t++)
                    for (int b = 0; b <= kBins; b++) { bins.cnt[b] += part[t].cnt[b]; bins.box[b].grow(part[t].box[b]); }
            } else {
                bin_range(split, axis, first, first + len, bins);
            }
            // suffix unions: R_i = bins i+1 .. 32
            Box suf[kBins + 1];
            int sufc[kBins + 1];
            Box acc;
            acc.clear();
            int accc = 0;
            for (int b = kBins; b >= 1; b--) {
                acc.grow(bins.box[b]);
                accc += bins.cnt[b];
                suf[b - 1] = acc;
                sufc[b - 1] = accc;
            }
            Box pre;
            pre.clear();
            int prec = 0;
            for (int i = 0; i < kBins; i++) {
                pre.grow(bins.box[i]);
                prec += bins.cnt[i];
                Box al, ar; // candidate boxes start at min = FLT_MAX, max = FLT_MIN (bvh.c:149-150)
                for (int a = 0; a < 3; a++) {
                    al.mn[a] = fmn(FLT_MAX, pre.mn[a]);
                    al.mx[a] = fmx(FLT_MIN, pre.mx[a]);
                    ar.mn[a] = fmn(FLT_MAX, suf[i].mn[a]);
                    ar.mx[a] = fmx(FLT_MIN, suf[i].mx[a]);
                }
                const int cl = prec, cr = sufc[i];
                float score;
                if (arith == rt::BVH_REFBIN) score = std::fmaf((float)cl, diag2(al), (float)cr * diag2(ar));
                else score = cl * diag2(al) + cr * diag2(ar); // bvh.c:169
                if (score < best) { best = score; splitAxis = axis; splitPos = split[i]; }
            }
        }
    }

    // bvh_split (bvh.c:78-267) on tree T.  guard_2n: apply the reference's "bvh_len >= 2N" stop (serial build of the
    // whole tree only).  tasks != nullptr: nodes smaller than stop_size are not split but recorded for a worker.
    void split(Tree& T, int node_idx, int depth, bool guard_2n, int stop_size, std::vector<Task>* tasks)
    {
        if (guard_2n && (size_t)T.len >= 2 * n) {                                 // bvh.c:80-83
            // the reference returns without clearing an EMPTY node's union field, which then reads as a
            // dangling child index; clear it (only reachable with heuristics 0/1 on degenerate input)
            if (!T.nodes[node_idx].tr_len) T.nodes[node_idx].idx = 0;
            return;
        }
        if (depth == kMaxDepth || T.nodes[node_idx].tr_len <= kLeafThreshold) {    // bvh.c:84
            if (!T.nodes[node_idx].tr_len) T.nodes[node_idx].idx = 0;             // bvh.c:85-86
            return;
        }
        if (tasks && T.nodes[node_idx].tr_len < stop_size) { tasks->push_back({node_idx, depth}); return; }
        const int child_idx = T.len;                                              // bvh.c:98-99
        T.len += 2;
        if ((size_t)T.len > T.nodes.size()) T.nodes.resize(std::max<size_t>(T.nodes.size() * 2, (size_t)T.len), rt_bvh_node{{0, 0, 0}, {0, 0, 0}, 0, 0});
        rt_bvh_node parent = T.nodes[node_idx];
        rt_bvh_node left{{0, 0, 0}, {0, 0, 0}, 0, parent.idx}, right{{0, 0, 0}, {0, 0, 0}, 0, parent.idx};
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
            (inA ? lb : rb).grow(tbox[t_idx]);
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
        T.nodes[child_idx] = left;
        T.nodes[child_idx + 1] = right;
        T.nodes[node_idx].idx = child_idx;                                        // bvh.c:262-263
        T.nodes[node_idx].tr_len = 0;
        split(T, child_idx, depth + 1, guard_2n, stop_size, tasks);               // bvh.c:265-266
        split(T, child_idx + 1, depth + 1, guard_2n, stop_size, tasks);
    }

    rt_bvh_node make_root()
    {
        s.tri_idx.resize(n);
        for (size_t i = 0; i < n; i++) s.tri_idx[i] = (int32_t)i;                  // bvh.c:366-368
        rt_bvh_node root{{0, 0, 0}, {0, 0, 0}, (int32_t)n, 0};
        Box rb;
        for (int a = 0; a < 3; a++) { rb.mn[a] = 1e10f; rb.mx[a] = -1e10f; }       // bvh.c:373-377
        for (size_t i = 0; i < n; i++) rb.grow(tbox[i]);
        std::memcpy(root.min, rb.mn, 12); std::memcpy(root.max, rb.mx, 12);
        return root;
    }

    void run_serial()
    {
        Tree T;
        // bvh.c:370-371 allocates 2N nodes, but the guard at bvh.c:80 lets a split begin at bvh_len == 2N-1
        // and write node 2N (reachable with empty children under heuristics 0/1): keep two spare nodes
        T.nodes.assign(2 * n + 2, rt_bvh_node{{0, 0, 0}, {0, 0, 0}, 0, 0});
        T.nodes[0] = make_root();
        split(T, 0, 0, true, 0, nullptr);
        T.nodes.resize((size_t)T.len);
        s.bvh.swap(T.nodes);
    }

    void run_parallel()
    {
        Tree top;
        top.nodes.assign(1024, rt_bvh_node{{0, 0, 0}, {0, 0, 0}, 0, 0});
        top.nodes[0] = make_root();
        std::vector<Task> tasks;
        const int stop = (int)std::max<size_t>(n / ((size_t)threads * 16), 4096);
        split(top, 0, 0, false, stop, &tasks);

        // subtrees, concurrently (disjoint tri_idx ranges)
        std::vector<Tree> sub(tasks.size());
        std::atomic<size_t> next{0};
        auto worker = [&] {
            for (size_t k; (k = next.fetch_add(1)) < tasks.size();) {
                Tree& T = sub[k];
                const rt_bvh_node r = top.nodes[tasks[k].top_node];
                T.nodes.assign((size_t)r.tr_len + 2, rt_bvh_node{{0, 0, 0}, {0, 0, 0}, 0, 0});
                T.nodes[0] = r;
                split(T, 0, tasks[k].depth, false, 0, nullptr);
            }
        };
        {
            Builder* self = this;
            const int saved = threads;
            std::vector<std::thread> th;
            const int t = std::max(1, std::min<int>(saved, (int)tasks.size()));
            self->threads = 1; // no nested parallel binning inside workers
            for (int i = 0; i < t; i++) th.emplace_back(worker);
            for (auto& x : th) x.join();
            self->threads = saved;
        }

        // reference numbering: children are allocated pairwise when the parent is split, left subtree first
        std::vector<int32_t> task_of(top.nodes.size(), -1);
        for (size_t k = 0; k < tasks.size(); k++) task_of[tasks[k].top_node] = (int32_t)k;
        size_t total = (size_t)top.len - tasks.size();
        for (const Tree& T : sub) total += (size_t)T.len;
        std::vector<rt_bvh_node> out(total);
        int32_t counter = 1;
        struct Item { int32_t t, f; };
        std::vector<Item> st{{0, 0}};
        while (!st.empty()) {
            const Item it = st.back();
            st.pop_back();
            const int32_t k = task_of[it.t];
            if (k >= 0) {
                const Tree& T = sub[(size_t)k];
                const int32_t m = T.len - 1, base = counter;
                counter += m;
                auto remap = [&](rt_bvh_node nd) { if (nd.tr_len == 0 && nd.idx != 0) nd.idx = base + (nd.idx - 1); return nd; };
                out[(size_t)it.f] = remap(T.nodes[0]);
                for (int32_t j = 1; j <= m; j++) out[(size_t)(base + j - 1)] = remap(T.nodes[(size_t)j]);
            } else {
                rt_bvh_node nd = top.nodes[(size_t)it.t];
                if (nd.tr_len == 0 && nd.idx != 0) {
                    const int32_t l = nd.idx, c = counter;
                    counter += 2;
                    nd.idx = c;
                    // right pushed first so that the whole left subtree is numbered before the right one
                    st.push_back({l + 1, c + 1});
                    st.push_back({l, c});
                }
                out[(size_t)it.f] = nd;
            }
        }
        out.resize((size_t)counter);
        s.bvh.swap(out);
    }

    void run()
    {
        prepare();
        if (heuristic == 6 && threads > 1 && n >= 200000) {
            run_parallel();
            // The parallel build has no `bvh_len >= 2N` stop (cpu/src/bvh.c:80-83): a tree that reaches 2N nodes is one the
            // reference would have cut short, so rebuild it with the serial builder, which carries the guard.
            if (s.bvh.size() >= 2 * (size_t)n) run_serial();
        } else run_serial();
    }
};

} // namespace

namespace rt {

int build_bvh(rt_scene& s, int heuristic, BvhArith arith, int threads)
{
    if (s.n_tris() == 0) { set_error("no triangles, cannot build bvh"); return RT_ERR_INVALID; } // bvh.c:361-364
    if (heuristic != 6 && heuristic != 0 && heuristic != 1) {
        set_error("rt_scene_build_bvh: heuristic must be 6, 0 or 1 (2/3 depend on libc rand() and an out-of-bounds "
                  "axis, 4/5 on qsort tie order; see SURVEY.md C.1)");
        return RT_ERR_INVALID;
    }
    if (s.n_tris() >= (1u << 27)) { set_error("more than 2^27 triangles"); return RT_ERR_INVALID; }
    if (threads <= 0) {
        const char* e = std::getenv("RT_BVH_THREADS"); // tests pin this to compare the serial and parallel builds
        threads = e ? std::atoi(e) : (int)std::thread::hardware_concurrency();
        if (threads <= 0) threads = 1;
    }
    Builder b(s, heuristic, arith, threads);
    b.run();
    return RT_OK;
}

} // namespace rt

// flatten.cpp — reference host layout (rt_scene_desc) -> the HBM layout of device_layout.h.
// Runs once per rt_create (the reference's load_to_gpu marshalling, gpu/src/gpu.cu:129-201).
// All derived quantities (edges, face normal, unit normal, |kr| > 0) are computed here in IEEE
// FP32 in the reference's operation order, so the strict kernel sees the bits the oracle computes.
// Must be compiled with -ffp-contract=off.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <atomic>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "flatten.h"
#include "host_scene.h"
#include "wide8.h"

namespace rt {

static inline void sub3(const float* a, const float* b, float* o) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
static inline void cross3(const float* a, const float* b, float* o) // vec_cross, cpu/src/vec.c:39-45
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

// The per-triangle and per-node passes are independent per element: chunk them over the host threads
// (RT_FLATTEN_THREADS overrides the count; small scenes stay on the calling thread).
template <class F>
static void parallel_for(size_t count, F f)
{
    static const int hw = [] {
        const char* e = std::getenv("RT_FLATTEN_THREADS");
        int t = e ? std::atoi(e) : (int)std::thread::hardware_concurrency();
        return t > 0 ? t : 1;
    }();
    const int t = (int)std::min<size_t>((size_t)hw, count / (1u << 15));
    if (t <= 1) { f((size_t)0, count); return; }
    std::vector<std::thread> th;
    for (int i = 0; i < t; i++) th.emplace_back([=] { f(count * i / t, count * (i + 1) / t); });
    for (auto& x : th) x.join();
}

// materials and lights (a few hundred bytes): shared by the host flatten and the device-side one (flatten_gpu.cu)
void flatten_small(const rt_scene_desc& d, FlatScene& out)
{
    const uint32_t n_mats = d.n_mats ? d.n_mats : 1;
    out.mats.assign(12 * (size_t)n_mats, 0.0f);
    for (uint32_t m = 0; m < d.n_mats; m++) {
        const float* k = d.materials + 9 * (size_t)m;
        float* q = &out.mats[12 * (size_t)m];
        q[0] = k[0]; q[1] = k[1]; q[2] = k[2];
        q[4] = k[3]; q[5] = k[4]; q[6] = k[5];
        q[8] = k[6]; q[9] = k[7]; q[10] = k[8];
        const float kr_mag = std::sqrt(k[6] * k[6] + k[7] * k[7] + k[8] * k[8]);
        q[11] = (kr_mag > 0.0) ? 1.0f : 0.0f; // vec_mag(&kr) > 0.0, raytracer.c:168
    }
    out.lights.assign(8 * (size_t)(d.n_lights ? d.n_lights : 1), 0.0f);
    for (uint32_t l = 0; l < d.n_lights; l++) {
        const float* s = d.lights + 6 * (size_t)l;
        float* q = &out.lights[8 * (size_t)l];
        q[0] = s[0]; q[1] = s[1]; q[2] = s[2];
        q[4] = s[3]; q[5] = s[4]; q[6] = s[5];
    }
    out.n_lights = d.n_lights;
    std::memcpy(out.ambient, d.ambient, 12);

}

int flatten_scene(const rt_scene_desc& d, FlatScene& out, std::string& err)
{
    if (!d.n_tris || !d.tri_coords) { err = "scene has no triangles"; return RT_ERR_INVALID; }
    if (!d.bvh || !d.tri_idx || !d.bvh_len) { err = "scene has no BVH (call rt_scene_build_bvh or supply bvh/tri_idx)"; return RT_ERR_INVALID; }
    if (d.n_tris >= (1u << 27)) { err = "more than 2^27 triangles"; return RT_ERR_INVALID; }
    const uint32_t n = d.n_tris, nb = d.bvh_len;
    const uint32_t n_mats = d.n_mats ? d.n_mats : 1;

    std::atomic<int> bad{0};
    parallel_for(n, [&](size_t lo, size_t hi) {
        for (size_t j = lo; j < hi; j++)
            if (d.tri_idx[j] < 0 || (uint32_t)d.tri_idx[j] >= n) bad.store(1);
    });
    if (bad.load()) { err = "tri_idx entry out of range"; return RT_ERR_INVALID; }

    // ---- triangles in leaf order ----
    out.tris.resize(16 * (size_t)n);
    out.shade.resize(4 * (size_t)n);
    parallel_for(n, [&](size_t lo, size_t hi) {
    for (size_t j = lo; j < hi; j++) {
        const uint32_t orig = (uint32_t)d.tri_idx[j];
        const float* c = d.tri_coords + 9 * (size_t)orig;
        float e1[3], e2[3], nn[3];
        sub3(c + 3, c, e1);  // raytracer.c:36
        sub3(c + 6, c, e2);  // raytracer.c:37
        cross3(e1, e2, nn);  // raytracer.c:38
        float* q = &out.tris[16 * (size_t)j];
        q[0] = c[0]; q[1] = c[1]; q[2] = c[2]; q[3] = e1[0];
        q[4] = e1[1]; q[5] = e1[2]; q[6] = e2[0]; q[7] = e2[1];
        q[8] = e2[2]; q[9] = nn[0]; q[10] = nn[1]; q[11] = nn[2];
        std::memcpy(&q[12], &orig, 4);
        q[13] = q[14] = q[15] = 0.0f;
    }
    });
    const uint32_t* tri_mat = d.tri_mat;
    parallel_for(n, [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) {
        const float* c = d.tri_coords + 9 * (size_t)i;
        float e1[3], e2[3], nn[3];
        sub3(c + 3, c, e1);
        sub3(c + 6, c, e2);
        cross3(e1, e2, nn);                                                  // triangle.c:14-17
        const float mag = std::sqrt(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]); // vec_mag
        float* s = &out.shade[4 * (size_t)i];
        s[0] = nn[0] / mag; s[1] = nn[1] / mag; s[2] = nn[2] / mag;          // vec_normalize
        uint32_t m = tri_mat ? tri_mat[i] : 0;
        if (m >= n_mats) { bad.store(1); m = 0; }
        std::memcpy(&s[3], &m, 4);
    }
    });
    if (bad.load()) { err = "material index out of range"; return RT_ERR_INVALID; }

    flatten_small(d, out);

    // ---- nodes: one record per inner node, DFS pre-order ----
    auto is_inner = [&](const rt_bvh_node& b) { return b.tr_len == 0 && b.idx != 0; };
    {   // leaves of >= 15 triangles (depth-capped) keep their count in a side table: allocate it before the parallel passes
        std::atomic<int> big{0};
        parallel_for(nb, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; i++) if (d.bvh[i].tr_len >= RT_LEAF_CNT_ESC_HOST) { big.store(1); break; }
        });
        if (big.load()) out.leaf_cnt.assign(n, 0);
    }
    auto leaf_ref = [&](const rt_bvh_node& b, int32_t& ref) -> bool {
        if (b.tr_len <= 0) { ref = RT_REF_NONE_HOST; return true; } // empty leaf: never pushed
        if (b.idx < 0 || (uint64_t)b.idx + (uint64_t)b.tr_len > n) return false;
        const int cnt = b.tr_len >= RT_LEAF_CNT_ESC_HOST ? RT_LEAF_CNT_ESC_HOST : b.tr_len;
        if (b.tr_len >= RT_LEAF_CNT_ESC_HOST) out.leaf_cnt[b.idx] = b.tr_len; // (a leaf is the child of one node only)
        ref = ~((b.idx << 4) | cnt);
        return true;
    };

    // pass 1: number the inner nodes (explicit stack; guards against cycles and bad indices)
    std::vector<int32_t> inner_of(nb, -1);
    std::vector<uint32_t> order; // reference node index of inner k
    struct Item { uint32_t node; int depth; };
    std::vector<Item> st;
    int max_depth = 0;
    if (is_inner(d.bvh[0])) st.push_back({0, 0});
    while (!st.empty()) {
        Item it = st.back();
        st.pop_back();
        const rt_bvh_node& b = d.bvh[it.node];
        if (inner_of[it.node] >= 0) { err = "BVH is not a tree (node reached twice)"; return RT_ERR_INVALID; }
        inner_of[it.node] = (int32_t)order.size();
        order.push_back(it.node);
        if (it.depth > max_depth) max_depth = it.depth;
        if (b.idx < 1 || (uint64_t)b.idx + 1 >= nb) {
            err = "BVH child index out of range";
            return RT_ERR_INVALID;
        }
        // right first so that the left subtree is numbered first (pre-order)
        if (is_inner(d.bvh[b.idx + 1])) st.push_back({(uint32_t)b.idx + 1, it.depth + 1});
        if (is_inner(d.bvh[b.idx])) st.push_back({(uint32_t)b.idx, it.depth + 1});
    }
    // live stack entries never exceed (inner depth + 2)
    if (max_depth + 5 > RT_STACK_ENTRIES_HOST) { err = "BVH deeper than the traversal stack (" + std::to_string(max_depth) + ")"; return RT_ERR_INVALID; }
    out.max_depth = max_depth;

    const size_t n_inner = order.empty() ? 1 : order.size();
    out.nodes.resize(16 * n_inner);
    if (order.empty()) std::fill(out.nodes.begin(), out.nodes.end(), 0.0f);
    auto put_child = [&](float* q, int which, const rt_bvh_node& c, int32_t ref) {
        float mn[3], mx[3];
        // An empty child must never be entered.  The slab test cannot see an inverted box (it takes min/max of
        // the two plane distances), so use a degenerate box at +infinity: every plane distance is +-inf, hence
        // either tmax < 0 (miss) or tmin = +inf, which is never < t.
        if (ref == RT_REF_NONE_HOST) { for (int a = 0; a < 3; a++) { mn[a] = INFINITY; mx[a] = INFINITY; } }
        else { std::memcpy(mn, c.min, 12); std::memcpy(mx, c.max, 12); }
        if (which == 0) { q[0] = mn[0]; q[1] = mn[1]; q[2] = mn[2]; q[3] = mx[0]; q[4] = mx[1]; q[5] = mx[2]; }
        else { q[6] = mn[0]; q[7] = mn[1]; q[8] = mn[2]; q[9] = mx[0]; q[10] = mx[1]; q[11] = mx[2]; }
        std::memcpy(&q[12 + which], &ref, 4);
    };
    if (order.empty()) {
        // the root itself is a leaf (<= 2 triangles): synthetic inner node, left = root, right = none
        int32_t ref;
        if (!leaf_ref(d.bvh[0], ref)) { err = "BVH leaf range out of bounds"; return RT_ERR_INVALID; }
        // the reference pops the root untested (cpu/src/bvh.c:321-324): give it a box no ray can miss
        rt_bvh_node all = d.bvh[0];
        for (int a = 0; a < 3; a++) { all.min[a] = -1e30f; all.max[a] = 1e30f; }
        put_child(out.nodes.data(), 0, all, ref);
        put_child(out.nodes.data(), 1, d.bvh[0], RT_REF_NONE_HOST);
    }
    parallel_for(order.size(), [&](size_t lo, size_t hi) {
        for (size_t k = lo; k < hi; k++) {
            const rt_bvh_node& b = d.bvh[order[k]];
            float* q = &out.nodes[16 * k];
            q[14] = q[15] = 0.0f;
            for (int w = 0; w < 2; w++) {
                const rt_bvh_node& c = d.bvh[b.idx + w];
                int32_t ref = RT_REF_NONE_HOST;
                if (is_inner(c)) ref = inner_of[b.idx + w];
                else if (!leaf_ref(c, ref)) { bad.store(1); ref = RT_REF_NONE_HOST; }
                put_child(q, w, c, ref);
            }
        }
    });
    if (bad.load()) { err = "BVH leaf range out of bounds"; return RT_ERR_INVALID; }

    // ---- 4-wide FP32 tree for the fast build (wide8.h: build_wide4) ----
    // A node holds up to four children of a reference subtree — its frontier after expanding the inner child of largest surface
    // area until four children exist — as SoA rows: cx[4] cy[4] cz[4] hx[4] hy[4] hz[4] (centre, half extent) refs[4] pad[4] =
    // 128 bytes; reference subtrees of <= wide4_leaf_max() triangles are one leaf.  Leaf references and triangle slots are
    // shared with the 2-wide layout.  (Until the middle of round 2 the tree was the collapse of every other reference level:
    // a third of its slots were empty, profiles/r02_notes.md §9.)
    if (!is_inner(d.bvh[0])) {
        // the root itself is a leaf: one node whose only child holds it, in a box no ray can miss
        out.nodes4.resize(32);
        float* q = out.nodes4.data();
        for (int i = 0; i < 12; i++) q[i] = INFINITY; // empty slot: centre at +inf, half extent 0, ref NONE
        for (int i = 12; i < 24; i++) q[i] = 0.0f;
        const int32_t none = RT_REF_NONE_HOST;
        for (int i = 0; i < 4; i++) std::memcpy(&q[24 + i], &none, 4);
        for (int i = 28; i < 32; i++) q[i] = 0.0f;
        int32_t ref = RT_REF_NONE_HOST;
        leaf_ref(d.bvh[0], ref);
        if (ref != RT_REF_NONE_HOST) {
            for (int a = 0; a < 3; a++) { q[4 * a] = 0.0f; q[12 + 4 * a] = 1e30f; }
            std::memcpy(&q[24], &ref, 4);
        }
        out.stack_need4 = 3;
    } else {
        std::vector<float> n4;
        int need = 0;
        const int rc4 = build_wide4(d.bvh, (uint32_t)nb, wide4_leaf_max(), n4, &need);
        if (rc4) { err = "BVH child index out of range (4-wide tree)"; return rc4; }
        out.nodes4.assign(n4.begin(), n4.end());
        out.stack_need4 = need + 3; // + sentinel, postponed leaf, slack
    }
    // ---- compressed 8-wide collapse (wide8.h): the fast build's default tree ----
    {
        Wide8Tree w8;
        const int rc8 = build_wide8(d.bvh, nb, wide8_leaf_max(), w8);
        if (rc8) { err = "BVH child index out of range (8-wide collapse)"; return rc8; }
        out.nodes8.assign(w8.words.begin(), w8.words.end());
        out.depth8 = w8.depth;
    }
    return RT_OK;
}

} // namespace rt

// Host-only view of the staging arrays (no device needed): lets the CPU test tier check the layout rules of
// device_layout.h and that the threaded passes do not depend on the thread count.
extern "C" int rt_debug_flatten_host(const rt_scene_desc* desc, int which, void* out, size_t cap_bytes, size_t* bytes_out, int* meta4)
{
    if (!desc || !bytes_out) { rt::set_error("rt_debug_flatten_host: null argument"); return RT_ERR_INVALID; }
    rt::FlatScene flat;
    std::string err;
    const int rc = rt::flatten_scene(*desc, flat, err);
    if (rc) { rt::set_error("rt_debug_flatten_host: " + err); return rc; }
    const void* src = nullptr;
    size_t n = 0;
    switch (which) {
    case 0: src = flat.nodes.data(); n = flat.nodes.size() * 4; break;
    case 1: src = flat.nodes4.data(); n = flat.nodes4.size() * 4; break;
    case 2: src = flat.tris.data(); n = flat.tris.size() * 4; break;
    case 3: src = flat.shade.data(); n = flat.shade.size() * 4; break;
    case 4: src = flat.leaf_cnt.data(); n = flat.leaf_cnt.size() * 4; break;
    case 5: src = flat.mats.data(); n = flat.mats.size() * 4; break;
    case 6: src = flat.lights.data(); n = flat.lights.size() * 4; break;
    case 7: src = flat.nodes8.data(); n = flat.nodes8.size() * 4; break;
    default: rt::set_error("rt_debug_flatten_host: bad array selector"); return RT_ERR_INVALID;
    }
    *bytes_out = n;
    if (meta4) { meta4[0] = flat.max_depth; meta4[1] = flat.stack_need4; meta4[2] = (int)flat.n_lights; meta4[3] = flat.depth8; }
    if (out) {
        if (cap_bytes < n) { rt::set_error("rt_debug_flatten_host: buffer too small"); return RT_ERR_INVALID; }
        if (n) std::memcpy(out, src, n);
    }
    return RT_OK;
}


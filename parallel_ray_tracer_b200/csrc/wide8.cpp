// wide8.cpp — host builder of the compressed 8-wide tree (layout and rules in wide8.h).  Level-synchronous: every level
// is three parallel passes over the previous level's inner children (expand + count, exclusive scan, encode), which is
// also how flatten_gpu.cu builds the same bytes on the device.
#include "wide8.h"
#include "flatten.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace rt {

namespace {
template <class F>
void par_for(size_t count, F f)
{
    static const int hw = [] {
        const char* e = std::getenv("RT_FLATTEN_THREADS");
        int t = e ? std::atoi(e) : (int)std::thread::hardware_concurrency();
        return t > 0 ? t : 1;
    }();
    const int t = (int)std::min<size_t>((size_t)hw, count / 2048);
    if (t <= 1) { f((size_t)0, count); return; }
    std::vector<std::thread> th;
    for (int i = 0; i < t; i++) th.emplace_back([=] { f(count * i / t, count * (i + 1) / t); });
    for (auto& x : th) x.join();
}
} // namespace

int wide8_leaf_max()
{
    const char* e = std::getenv("RT_W8_LEAF_MAX");
    int v = e ? std::atoi(e) : kWide8LeafMaxDefault;
    return v < 2 ? 2 : (v > RT_W8_LEAF_CAP ? RT_W8_LEAF_CAP : v);
}

int wide4_leaf_max()
{
    const char* e = std::getenv("RT_W4_LEAF_MAX");
    const int v = e ? std::atoi(e) : 3;
    return v < 1 ? 1 : (v > RT_W8_LEAF_CAP ? RT_W8_LEAF_CAP : v);
}

int build_wide8(const rt_bvh_node* bvh, uint32_t bvh_len, int leaf_max, Wide8Tree& out)
{
    out.words.clear();
    out.depth = 0;
    if (!bvh || !bvh_len) return RT_ERR_INVALID;
    auto is_inner = [&](uint32_t b) { return bvh[b].tr_len == 0 && bvh[b].idx != 0; };
    int32_t rf = 0, rc = 0;
    if (!is_inner(0)) {
        // the root itself is a leaf (<= 2 triangles): one node with one child
        W8Child c;
        for (int a = 0; a < 3; a++) { c.mn[a] = bvh[0].min[a]; c.mx[a] = bvh[0].max[a]; }
        c.bnode = 0; c.inner = 0;
        rf = bvh[0].idx; rc = bvh[0].tr_len;
        const int slot = 0;
        const int32_t ref = w8_leaf_ref(rf, rc);
        out.words.resize(kWide8Words);
        w8_encode(&c, 1, &slot, &ref, out.words.data());
        if (rc <= 0) out.words[3] &= 0x00ffffffu; // no triangles at all: zero children
        out.depth = 1;
        return RT_OK;
    }
    std::vector<uint32_t> level{0}, next;
    std::vector<uint32_t> n_inner, offset;
    size_t base = 0;
    while (!level.empty()) {
        const size_t n = level.size();
        n_inner.assign(n, 0);
        offset.assign(n + 1, 0);
        std::atomic<int> bad{0};
        par_for(n, [&](size_t lo, size_t hi) {
            W8Child ch[8];
            for (size_t i = lo; i < hi; i++) {
                const uint32_t r = level[i];
                if ((uint64_t)bvh[r].idx + 1 >= bvh_len || bvh[r].idx < 1) { bad.store(1); continue; }
                const int c = w8_expand(bvh, r, leaf_max, ch);
                uint32_t k = 0;
                for (int j = 0; j < c; j++) k += (uint32_t)ch[j].inner;
                n_inner[i] = k;
            }
        });
        if (bad.load()) return RT_ERR_INVALID;
        for (size_t i = 0; i < n; i++) offset[i + 1] = offset[i] + n_inner[i];
        const size_t next_base = base + n;
        next.assign(offset[n], 0);
        out.words.resize((next_base) * kWide8Words);
        par_for(n, [&](size_t lo, size_t hi) {
            W8Child ch[8];
            int slot_of[8];
            int32_t ref_of[8];
            for (size_t i = lo; i < hi; i++) {
                const int c = w8_expand(bvh, level[i], leaf_max, ch);
                w8_assign_slots(ch, c, slot_of);
                // inner children are numbered in SLOT order (what a kernel could recompute from the slot alone)
                int order[8], m = 0;
                for (int s = 0; s < 8; s++)
                    for (int j = 0; j < c; j++)
                        if (slot_of[j] == s && ch[j].inner) order[m++] = j;
                for (int j = 0; j < c; j++) ref_of[j] = ch[j].inner ? 0 : w8_leaf_ref(ch[j].first, ch[j].cnt);
                for (int k = 0; k < m; k++) {
                    ref_of[order[k]] = (int32_t)(next_base + offset[i] + (size_t)k);
                    next[offset[i] + (size_t)k] = (uint32_t)ch[order[k]].bnode;
                }
                w8_encode(ch, c, slot_of, ref_of, &out.words[(base + i) * kWide8Words]);
            }
        });
        base = next_base;
        level.swap(next);
        out.depth++;
        if (out.depth > 64) return RT_ERR_INVALID; // (a reference tree is at most 33 levels deep)
    }
    return RT_OK;
}

int build_wide4(const rt_bvh_node* bvh, uint32_t bvh_len, int leaf_max, std::vector<float>& nodes4, int* stack_need)
{
    nodes4.clear();
    if (stack_need) *stack_need = 0;
    if (!bvh || !bvh_len) return RT_ERR_INVALID;
    auto is_inner = [&](uint32_t b) { return bvh[b].tr_len == 0 && bvh[b].idx != 0; };
    if (!is_inner(0)) return RT_ERR_INVALID; // (a root that is a leaf is flatten.cpp's synthetic single node)
    auto put_node = [&](float* q, const W8Child* ch, int c, const int32_t* ref_of) {
        for (int i = 0; i < 12; i++) q[i] = INFINITY; // empty slot: centre +inf, half extent 0, ref NONE
        for (int i = 12; i < 24; i++) q[i] = 0.0f;
        const int32_t none = RT_REF_NONE_HOST;
        for (int i = 0; i < 4; i++) std::memcpy(&q[24 + i], &none, 4);
        for (int i = 28; i < 32; i++) q[i] = 0.0f;
        for (int j = 0; j < c; j++) {
            if (ref_of[j] == RT_REF_NONE_HOST) continue;
            for (int a = 0; a < 3; a++) {
                float ctr, half;
                box_center_half(ch[j].mn[a], ch[j].mx[a], ctr, half);
                q[4 * a + j] = ctr; q[12 + 4 * a + j] = half;
            }
            std::memcpy(&q[24 + j], &ref_of[j], 4);
        }
    };
    std::vector<uint32_t> level{0}, next, n_inner, offset;
    size_t base = 0;
    int depth = 0;
    while (!level.empty()) {
        const size_t n = level.size();
        n_inner.assign(n, 0);
        offset.assign(n + 1, 0);
        std::atomic<int> bad{0};
        par_for(n, [&](size_t lo, size_t hi) {
            W8Child ch[8];
            for (size_t i = lo; i < hi; i++) {
                const uint32_t r = level[i];
                if ((uint64_t)bvh[r].idx + 1 >= bvh_len || bvh[r].idx < 1) { bad.store(1); continue; }
                const int c = w8_expand(bvh, r, leaf_max, ch, 4);
                uint32_t k = 0;
                for (int j = 0; j < c; j++) k += (uint32_t)ch[j].inner;
                n_inner[i] = k;
            }
        });
        if (bad.load()) return RT_ERR_INVALID;
        for (size_t i = 0; i < n; i++) offset[i + 1] = offset[i] + n_inner[i];
        const size_t next_base = base + n;
        next.assign(offset[n], 0);
        nodes4.resize(next_base * 32);
        par_for(n, [&](size_t lo, size_t hi) {
            W8Child ch[8];
            int32_t ref_of[8];
            for (size_t i = lo; i < hi; i++) {
                const int c = w8_expand(bvh, level[i], leaf_max, ch, 4);
                uint32_t k = 0;
                for (int j = 0; j < c; j++) {
                    if (ch[j].inner) {
                        ref_of[j] = (int32_t)(next_base + offset[i] + k);
                        next[offset[i] + k] = (uint32_t)ch[j].bnode;
                        k++;
                    } else ref_of[j] = w8_leaf_ref(ch[j].first, ch[j].cnt);
                }
                put_node(&nodes4[(base + i) * 32], ch, c, ref_of);
            }
        });
        base = next_base;
        level.swap(next);
        if (++depth > 64) return RT_ERR_INVALID;
    }
    if (stack_need) { // bottom-up: children have larger indices than their parent
        const size_t n4 = nodes4.size() / 32;
        std::vector<int32_t> need(n4, 0);
        for (size_t k = n4; k-- > 0;) {
            const float* q = &nodes4[32 * k];
            int live = 0, deepest = 0;
            for (int i = 0; i < 4; i++) {
                int32_t ref;
                std::memcpy(&ref, &q[24 + i], 4);
                if (ref == RT_REF_NONE_HOST) continue;
                live++;
                if (ref >= 0) deepest = std::max(deepest, need[(size_t)ref]);
            }
            need[k] = std::max(live - 1, 0) + deepest;
        }
        *stack_need = need[0];
    }
    return RT_OK;
}

} // namespace rt

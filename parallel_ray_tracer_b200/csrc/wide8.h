// wide8.h — the compressed 8-wide collapse of the reference BVH (fast build only; device_layout.h: nodes8).
//
// One 96-byte record per 8-wide node, 32-byte aligned, read by the kernels as three 256-bit loads:
//
//   word 0..2   p.x p.y p.z      float: origin of the node's quantisation grid (just below the node's box minimum)
//   word 3      Ex | Ey<<8 | Ez<<16 | n_children<<24     E = biased exponent of s/128, s = 2^e the grid step of the axis
//   word 4..5   qlo_x[8]   word 6..7  qlo_y[8]   word 8..9   qlo_z[8]      one byte per child slot
//   word 10..11 qhi_x[8]   word 12..13 qhi_y[8]  word 14..15 qhi_z[8]
//   word 16..23 ref[8]     child references in the encoding of device_layout.h (>= 0: index of an 8-wide node,
//                          RT_REF_NONE: empty slot, otherwise ~((first_slot << 4) | count): a leaf of the REFERENCE tree)
//
// Child boxes are the reference's FP32 boxes rounded OUTWARD onto the node's 8-bit grid: lo = p + s*qlo <= true lo - s/16,
// hi = p + s*qhi >= true hi + s/16 (checked in double at build time).  The 1/16-step margin covers the rounding of the
// kernel's decode-and-test arithmetic (render_kernel.cuh: wide8_visit), so the test can only accept more than the exact
// FP32 box test would, never less — the opposite of the reference GPU program's round-to-nearest FP16 boxes
// (gpu/src/gpu.cu:176-185), which cull wrongly.  Empty slots hold the inverted range qlo = 255, qhi = 0 and are never hit.
//
// The tree is the reference tree (cpu/src/bvh.c:78-267) with levels removed: an 8-wide node is rooted at a reference inner
// node and holds the frontier reached by repeatedly replacing the inner child of largest surface area by its two
// children until eight children exist or none is inner; leaves of the reference tree stay leaves (same triangle slots).
// Children are assigned to slots so that slot ^ octant enumerates them roughly front to back for a ray of that direction
// octant (greedy assignment on centroid offsets): traversal needs no distance sort.  Nodes are numbered breadth-first,
// the inner children of a node consecutively — every level is a parallel pass over the previous one (flatten_gpu.cu runs
// the same passes on the device and produces the same bytes).
#pragma once
#include <cstdint>
#include <vector>

#include "rt_b200.h"

namespace rt {

constexpr int kWide8Words = 24;       // 96 bytes
constexpr int kWide8StackMax = 40;    // group-stack entries a kernel provides (sentinel + one group per level)

// Everything below is shared, source for source, by the host builder (wide8.cpp) and the device builder (flatten_gpu.cu).
#if defined(__CUDACC__)
#define RT_W8_HD __host__ __device__
#else
#define RT_W8_HD
#endif

struct W8Child {
    float mn[3], mx[3];
    int32_t bnode; // reference node index
    int32_t inner; // 1: becomes an 8-wide node, 0: leaf (a reference leaf, or a whole reference subtree of <= leaf_max triangles)
    int32_t first, cnt; // leaf: triangle slots [first, first + cnt)
};

// Reference subtrees of at most this many triangles become ONE leaf of the 8-wide tree (their slots are contiguous:
// bvh_split partitions tri_idx in place, cpu/src/bvh.c:244-259).  Testing 3-4 triangles costs less than visiting an
// 8-wide node whose slots would be three quarters empty; the image cannot change (the triangle test is the hit test).
#define RT_W8_LEAF_CAP 8                 /* largest value of the knob */
constexpr int kWide8LeafMaxDefault = 4;  /* rt_create default; RT_W8_LEAF_MAX in the environment overrides it (experiments) */

// surface area of a box, double (only used to rank the children of ONE node against each other)
RT_W8_HD inline double w8_area(const float* mn, const float* mx)
{
    const double dx = (double)mx[0] - mn[0], dy = (double)mx[1] - mn[1], dz = (double)mx[2] - mn[2];
    return dx * dy + dy * dz + dz * dx;
}

RT_W8_HD inline bool w8_is_inner(const rt_bvh_node& nd) { return nd.tr_len == 0 && nd.idx != 0; }

// Triangle slots below reference node b, if there are at most leaf_max (<= RT_W8_LEAF_CAP) of them (bounded walk, left to right).
RT_W8_HD inline bool w8_small_subtree(const rt_bvh_node* bvh, uint32_t b, int leaf_max, int32_t* first, int32_t* cnt)
{
    uint32_t stack[2 * RT_W8_LEAF_CAP + 34];
    int sp = 0, total = 0;
    int32_t f = -1;
    stack[sp++] = b;
    while (sp) {
        const rt_bvh_node& nd = bvh[stack[--sp]];
        if (w8_is_inner(nd)) {
            if (sp + 2 > (int)(sizeof stack / sizeof stack[0])) return false;
            stack[sp++] = (uint32_t)nd.idx + 1; // right below left: left is popped first
            stack[sp++] = (uint32_t)nd.idx;
        } else if (nd.tr_len > 0) {
            if (f < 0) f = nd.idx;
            total += nd.tr_len;
            if (total > leaf_max) return false;
        }
    }
    *first = f < 0 ? 0 : f;
    *cnt = total;
    return true;
}

// device reference of a leaf (device_layout.h): ~((first_slot << 4) | min(count, 15)); count >= 15 is looked up in leaf_cnt
RT_W8_HD inline int32_t w8_leaf_ref(int32_t first, int32_t cnt)
{
    return cnt <= 0 ? (int32_t)0x80000000 : ~((first << 4) | (cnt >= 15 ? 15 : cnt));
}

// The frontier of reference node `root` (an inner node): up to 8 children, in the order the expansion leaves them
// (left to right in the reference tree).  Empty leaves of the reference tree (tr_len == 0 && idx == 0) are dropped.
RT_W8_HD inline int w8_expand(const rt_bvh_node* bvh, uint32_t root, int leaf_max, W8Child* out, int width = 8)
{
    int n = 0;
    auto put = [&](int at, uint32_t b) {
        const rt_bvh_node& nd = bvh[b];
        W8Child& c = out[at];
        for (int a = 0; a < 3; a++) { c.mn[a] = nd.min[a]; c.mx[a] = nd.max[a]; }
        c.bnode = (int32_t)b;
        c.first = nd.idx; c.cnt = nd.tr_len;
        c.inner = 0;
        if (w8_is_inner(nd)) c.inner = w8_small_subtree(bvh, b, leaf_max, &c.first, &c.cnt) ? 0 : 1;
    };
    auto empty = [&](uint32_t b) { return bvh[b].tr_len == 0 && bvh[b].idx == 0; };
    const uint32_t l = (uint32_t)bvh[root].idx;
    if (!empty(l)) put(n++, l);
    if (!empty(l + 1)) put(n++, l + 1);
    for (;;) {
        if (n >= width) break;
        int best = -1;
        double best_a = -1.0;
        for (int i = 0; i < n; i++)
            if (out[i].inner) {
                const double a = w8_area(out[i].mn, out[i].mx);
                if (a > best_a) { best_a = a; best = i; } // first of equal areas: deterministic
            }
        if (best < 0) break;
        const uint32_t c = (uint32_t)bvh[out[best].bnode].idx;
        const bool e0 = empty(c), e1 = empty(c + 1);
        if (e0 && e1) { // (cannot happen in a reference tree: a split node has at least one triangle) drop the child
            for (int i = best; i + 1 < n; i++) out[i] = out[i + 1];
            n--;
            continue;
        }
        if (e0 || e1) { put(best, e0 ? c + 1 : c); continue; }
        for (int i = n; i > best + 1; i--) out[i] = out[i - 1]; // keep left-to-right order
        put(best, c);
        put(best + 1, c + 1);
        n++;
    }
    return n;
}

// Slot of every child: greedy assignment maximising sum of dot(child centroid - node centroid, diagonal of the slot),
// where slot s sits towards (s&1 ? +x : -x, s&2 ? +y : -y, s&4 ? +z : -z).  slot_of[i] in 0..7, all distinct.
RT_W8_HD inline void w8_assign_slots(const W8Child* ch, int n, int* slot_of)
{
    double cmn[3] = {1e300, 1e300, 1e300}, cmx[3] = {-1e300, -1e300, -1e300};
    for (int i = 0; i < n; i++)
        for (int a = 0; a < 3; a++) {
            if (ch[i].mn[a] < cmn[a]) cmn[a] = ch[i].mn[a];
            if (ch[i].mx[a] > cmx[a]) cmx[a] = ch[i].mx[a];
        }
    double cost[8][8];
    for (int i = 0; i < n; i++) {
        double off[3];
        for (int a = 0; a < 3; a++) off[a] = 0.5 * ((double)ch[i].mn[a] + ch[i].mx[a]) - 0.5 * (cmn[a] + cmx[a]);
        for (int s = 0; s < 8; s++)
            cost[i][s] = ((s & 1) ? off[0] : -off[0]) + ((s & 2) ? off[1] : -off[1]) + ((s & 4) ? off[2] : -off[2]);
    }
    bool child_done[8] = {false, false, false, false, false, false, false, false};
    bool slot_used[8] = {false, false, false, false, false, false, false, false};
    for (int round = 0; round < n; round++) {
        int bi = -1, bs = -1;
        double bc = -1e300;
        for (int i = 0; i < n; i++) {
            if (child_done[i]) continue;
            for (int s = 0; s < 8; s++) {
                if (slot_used[s]) continue;
                if (cost[i][s] > bc) { bc = cost[i][s]; bi = i; bs = s; } // first maximum: deterministic
            }
        }
        child_done[bi] = true;
        slot_used[bs] = true;
        slot_of[bi] = bs;
    }
}

// floor(log2(x)) for a positive finite double, without libm (bit-identical on host and device)
RT_W8_HD inline int w8_ilog2(double x)
{
    union { double d; uint64_t u; } v;
    v.d = x;
    const int e = (int)((v.u >> 52) & 0x7ff);
    return e == 0 ? -1023 : e - 1023;
}
RT_W8_HD inline double w8_pow2(int e)
{
    union { double d; uint64_t u; } v;
    v.u = (uint64_t)(e + 1023) << 52;
    return v.d;
}
// largest float <= x (x finite, |x| < 1e30)
RT_W8_HD inline float w8_float_down(double x)
{
    float f = (float)x;
    if ((double)f > x) {
        union { float f; uint32_t u; } v;
        v.f = f;
        if (f > 0.0f) v.u -= 1;
        else if (f < 0.0f) v.u += 1;
        else v.u = 0x80000001u; // below +0: the smallest negative denormal
        f = v.f;
    }
    return f;
}

// Grid of one axis for the range [lo, hi] with coordinates up to `amax` in magnitude: step 2^e and origin p such that
// every q = floor((x - p)/s - 1/16) is >= 0 and every q = ceil((x - p)/s + 1/16) is <= 255 for x in [lo, hi].
RT_W8_HD inline void w8_axis_grid(float lo, float hi, double amax, int* e_out, float* p_out)
{
    // resolution floor: a step is at least 2^-17 of the largest coordinate (a finer grid is below what FP32 ray arithmetic
    // resolves); and never outside [-100, 100] so that every derived scale stays a normal float
    int e_min = -100;
    if (amax > 0.0) { const int f = w8_ilog2(amax) - 17; if (f > e_min) e_min = f; }
    const double ext = (double)hi - (double)lo;
    int e = e_min;
    if (ext > 0.0) { const int g = w8_ilog2(ext / 254.0); if (g > e) e = g; }
    if (e > 100) e = 100;
    for (;; e++) {
        const double s = w8_pow2(e);
        const float p = w8_float_down((double)lo - s * 0.125);
        const double top = ((double)hi - (double)p) / s + 0.0625;
        if (top <= 255.0 || e >= 100) { *e_out = e; *p_out = p; return; }
    }
}

// One axis of one child on the node's grid (origin p, step 2^e): outward-rounded bytes with the 1/16-step margin.
RT_W8_HD inline void w8_quantize(float cmn, float cmx, float p, int e, unsigned char* qlo, unsigned char* qhi)
{
    const double st = w8_pow2(e);
    const double ql = ((double)cmn - (double)p) / st - 0.0625;
    const double qh = ((double)cmx - (double)p) / st + 0.0625;
    long long il = (long long)ql; if ((double)il > ql) il--;   // floor
    long long ih = (long long)qh; if ((double)ih < qh) ih++;   // ceil
    if (il < 0) il = 0;
    if (il > 255) il = 255;
    if (ih > 255) ih = 255;
    if (ih < 0) ih = 0;
    *qlo = (unsigned char)il;
    *qhi = (unsigned char)ih;
}
// The grid of one axis of a node whose children span [lo, hi] on it
RT_W8_HD inline void w8_node_axis(float lo, float hi, int* e, float* p)
{
    const double al = lo < 0 ? -(double)lo : (double)lo, ah = hi < 0 ? -(double)hi : (double)hi;
    w8_axis_grid(lo, hi, al > ah ? al : ah, e, p);
}
RT_W8_HD inline uint32_t w8_header_word(const int* e, int n)
{
    return (uint32_t)(e[0] - 7 + 127) | ((uint32_t)(e[1] - 7 + 127) << 8) | ((uint32_t)(e[2] - 7 + 127) << 16) | ((uint32_t)n << 24);
}

// Encode one node.  kids[0..n) with their slots; ref_of[i] = device reference of child i.  `w` receives 24 words.
RT_W8_HD inline void w8_encode(const W8Child* ch, int n, const int* slot_of, const int32_t* ref_of, uint32_t* w)
{
    float lo[3], hi[3];
    for (int a = 0; a < 3; a++) { lo[a] = ch[0].mn[a]; hi[a] = ch[0].mx[a]; }
    for (int i = 1; i < n; i++)
        for (int a = 0; a < 3; a++) {
            if (ch[i].mn[a] < lo[a]) lo[a] = ch[i].mn[a];
            if (ch[i].mx[a] > hi[a]) hi[a] = ch[i].mx[a];
        }
    int e[3];
    float p[3];
    for (int a = 0; a < 3; a++) w8_node_axis(lo[a], hi[a], &e[a], &p[a]);
    for (int k = 0; k < kWide8Words; k++) w[k] = 0;
    for (int a = 0; a < 3; a++) {
        union { float f; uint32_t u; } v;
        v.f = p[a];
        w[a] = v.u;
    }
    w[3] = w8_header_word(e, n); // E = biased exponent of s/128 per axis, child count
    unsigned char qlo[3][8], qhi[3][8];
    for (int a = 0; a < 3; a++)
        for (int s = 0; s < 8; s++) { qlo[a][s] = 255; qhi[a][s] = 0; } // empty slot: inverted, never hit
    for (int s = 0; s < 8; s++) w[16 + s] = 0x80000000u;               // RT_REF_NONE
    for (int i = 0; i < n; i++) {
        const int s = slot_of[i];
        for (int a = 0; a < 3; a++) w8_quantize(ch[i].mn[a], ch[i].mx[a], p[a], e[a], &qlo[a][s], &qhi[a][s]);
        w[16 + s] = (uint32_t)ref_of[i];
    }
    for (int a = 0; a < 3; a++)
        for (int s = 0; s < 8; s++) {
            w[4 + 2 * a + (s >> 2)] |= (uint32_t)qlo[a][s] << (8 * (s & 3));
            w[10 + 2 * a + (s >> 2)] |= (uint32_t)qhi[a][s] << (8 * (s & 3));
        }
}

// Host builder (wide8.cpp).
struct Wide8Tree {
    std::vector<uint32_t> words; // 24 per node
    int depth = 0;               // levels of 8-wide nodes (a ray's group stack needs depth + 1 entries)
    size_t n_nodes() const { return words.size() / kWide8Words; }
};
int build_wide8(const rt_bvh_node* bvh, uint32_t bvh_len, int leaf_max, Wide8Tree& out);

// The 4-wide FP32 tree of the fast build (device_layout.h: nodes4, 32 floats per node: centre and half extent of four child boxes,
// SoA, + four references) by the SAME expansion with width 4: a node rooted at a reference inner node holds the frontier reached by
// replacing the inner child of largest surface area by its two children until four children exist or none is inner — nodes are
// full wherever the subtree allows it, where the collapse of every other level left a third of the slots empty.  Reference
// subtrees of at most leaf_max triangles become one leaf (contiguous slots).  Nodes are numbered breadth-first; *stack_need is the
// number of traversal-stack entries a ray can need below the root (a node with c children enters one and leaves c - 1 pushed).
int build_wide4(const rt_bvh_node* bvh, uint32_t bvh_len, int leaf_max, std::vector<float>& nodes4, int* stack_need);
int wide4_leaf_max(); // 3, or RT_W4_LEAF_MAX from the environment, clamped to [1, RT_W8_LEAF_CAP] (sweep: profiles/r02_ab_w4_adaptive.log)
int wide8_leaf_max(); // the knob: kWide8LeafMaxDefault or RT_W8_LEAF_MAX from the environment, clamped to [2, RT_W8_LEAF_CAP]

} // namespace rt

// device_layout.h — the scene as it lives in HBM, and the per-frame kernel arguments.
//
// The reference GPU program keeps AoS triangles (48 B), split normals (32 B), a 24 B node with
// FP16 boxes and walks nodes one at a time (gpu/src/gpu.cu:129-201, gpu/src/bvh.cu:342-392).
// This layout is different by design:
//
//   nodes   one 64-byte record per INNER node holding BOTH child boxes in FP32 and both child
//           references, so an inner visit is one 64 B fetch (4 x LDG.128) instead of the
//           reference's pop fetch + two child fetches:
//             q0 = (L.min.x, L.min.y, L.min.z, L.max.x)
//             q1 = (L.max.y, L.max.z, R.min.x, R.min.y)
//             q2 = (R.min.z, R.max.x, R.max.y, R.max.z)
//             q3 = (left_ref, right_ref, 0, 0) as int bits
//           L/R are the reference's children (child, child+1; cpu/include/bvh.h:19-22) so the
//           reference's left-first tie rule (cpu/src/bvh.c:344) is preserved.
//   refs    ref >= 0            : inner node index (into `nodes`)
//           ref == RT_REF_NONE  : empty leaf / nothing (never pushed)
//           otherwise (ref < 0) : leaf, ~ref = (first_slot << 4) | min(count, 15); count == 15
//                                 means "look the count up in leaf_cnt[first_slot]"
//   tris    64 bytes per LEAF-ORDER slot j (slot j holds triangle tri_idx[j], so a leaf's
//           triangles are contiguous): q0 = (v0.xyz, e1.x) q1 = (e1.yz, e2.xy) q2 = (e2.z, n.xyz)
//           q3 = (original triangle index as int bits — the first-hit ID space —, 0, 0, 0)
//           with e1 = v1-v0, e2 = v2-v0, n = e1 x e2 computed in IEEE FP32 on the host in the
//           reference's operation order (cpu/src/raytracer.c:36-38) — bit-identical to
//           computing them per test, and never normalised.  A test reads the
//           first 48 bytes (one 256-bit + one 128-bit load); q3 is read once per ray, on a hit.
//           64-byte records keep every 256-bit load 32-byte aligned.
//   nodes8  the fast build's default tree: the reference tree collapsed to 8-wide nodes of 96 bytes with child boxes
//           quantised to 8 bits per plane, rounded outward (layout and rules: wide8.h)
//   shade[orig]   (unit normal norm[0].xyz, material index as int bits); norm[1] == -norm[0]
//   mats    3 x float4 per material: (ks, 0) (kd, 0) (kr, |kr| > 0 ? 1 : 0)
//   lights  2 x float4 per light: (pos, 0) (kl, 0)
#pragma once
#include <stdint.h>
#include "rt_b200.h"

#define RT_TILE_PIXELS (RT_TILE_W * RT_TILE_H)
#ifndef RT_MACRO_CHUNKS
#define RT_MACRO_CHUNKS 16u /* chunks (8x4 pixels) per macro tile = 4 tiles */
#endif
#define RT_MAX_SMS 256

#define RT_REF_NONE ((int)0x80000000)
#define RT_LEAF_CNT_ESC 15
#define RT_STACK_ENTRIES 40   /* sentinel + reference depth cap 32 (cpu/include/options.h:64) + postponed leaf */
#define RT_STACK_ENTRIES_WIDE 48 /* 4-wide tree: up to three siblings stay pushed per level */
#define RT_STACK8_ENTRIES 40     /* 8-wide tree: one (node, remaining-children mask) group per level, 8 bytes each */
#define RT_MAX_BOUNCES 8

struct RtDeviceScene {
    const float4* nodes;
    const float4* nodes4;   // 4-wide tree (fast build; wide8.h: build_wide4), 8 x float4 per node
    const uint4*  nodes8;   // compressed 8-wide collapse (fast build), 96 bytes = 6 x uint4 per node (wide8.h); may be null
    const float4* tris;
    const float4* shade;
    const float4* mats;
    const float4* lights;
    const int*    leaf_cnt;
    int   n_lights;
    float amb[3];
    // array lengths and an error word for the checked build (RT_DEBUG_BOUNDS, render_kernel.cuh); unused otherwise
    unsigned n_tris, n_inner, n_nodes4, n_nodes8;
    unsigned long long* err;
};

// A path handed from the per-lane render kernel to the cooperative drain kernel (render_kernel.cuh: drain_kernel): the
// pixel, what has been accumulated so far, and the ray that was in flight (it is traced again from its start).
struct RtPathRec {
    int   pix, sample, depth, kind;
    float acc[3], col[3], thr[3];
    float o[3], d[3];
    float ld2;
    float P[3], n[3], in[3], pend[3];
    int   mat, li, culled, pad;
};
#define RT_DRAIN_STACK 96 /* per-path stack of the drain kernel: (child ref, entry distance) pairs, up to 7 per level */

struct RtFrameArgs {
    // camera basis exactly as thread_render derives it (cpu/src/main.c:241-250)
    float pos[3], ul[3], inc_x[3], inc_y[3];
    int   width, height, spp, bounces;
    int   flip_y;                  // RT_FRAME_BOTTOM_UP: BGRA row y is stored at row height-1-y (BMP order)
    unsigned seed;
    int   tiles_x;                 // tiles per image row
    const unsigned* tile_list;     // tiles this device renders (tile id = ty * tiles_x + tx)
    int   n_tiles;
    unsigned* tile_counter;        // persistent-CTA work counter (device-local)
    unsigned long long* sm_cursor; // RT_OPT_SMQUEUE: one work cursor per SM (indexed by %smid), zeroed per frame
    unsigned n_sms;
    int   refill_threshold;
    int   cull;                    // fast build: test every chunk's ray pyramid against the top of nodes8 first (needs sc.nodes8)
    // fast build, heaviest tiles first (rt_api.cu: tile_class_kernel): receives every rendered pixel's traversal steps (u16,
    // saturating); the next frame's tile_list is ordered by them.  Null = not tracked.
    unsigned short* cost_out;
    // fast build, tail of the frame: once the chunk queue is empty, a warp left with <= drain_k live pixels writes their
    // paths to drain_queue and exits; drain_kernel finishes them with eight lanes per ray (0 = off)
    int   drain_k;
    unsigned drain_cap;            // records drain_queue can hold
    RtPathRec* drain_queue;
    unsigned* drain_count;         // records written (render kernel) / to process (drain kernel)
    unsigned* drain_next;          // drain kernel: next record to hand out
    // outputs (bgra may be a peer-mapped pointer into device 0's frame)
    uchar4* bgra;
    float*  rgb;                   // optional
    int*    tri_id;                // optional
    float*  depth;                 // optional
    unsigned long long* stats;     // [0] closest rays [1] shadow rays [2] inner visits [3] triangle tests
    unsigned long long* warp_trace; // optional, RT_AOV_WORK builds: 8 x u64 per warp (clock64 start / queue empty / exit, counts)
};

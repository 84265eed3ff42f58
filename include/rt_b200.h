/*
 * rt_b200.h — C-ABI of the B200-native render hot path (librt_b200.so).
 *
 * The reference (deluf/parallel-ray-tracer) has no plugin/FFI interface: its render call
 * is a set of free functions over process-wide globals.  The shape this header makes
 * explicit and re-entrant is the reference GPU program's triple
 *
 *      void  load_to_gpu();                                  gpu/include/gpu.cuh:24, gpu/src/gpu.cu:129-201
 *      float render_frame(bool is_metrics, int tx, int ty);  gpu/include/gpu.cuh:23, gpu/src/gpu.cu:98-127
 *      void  load_from_gpu();                                gpu/include/gpu.cuh:25, gpu/src/gpu.cu:203-228
 *
 * (CPU program: `void render_frame()` over the same globals, cpu/src/main.c:42,214-226.)
 * Each export below cites the reference interface it replaces.  Only plain pointers and
 * sizes cross the boundary; no CUDA, torch or C++ types.  Every function returns 0 on
 * success or a negative rt_status; nothing in the library calls exit() (the reference
 * does, cpu/src/triangle.c:28-31).  There is NO CPU fallback: without a CUDA device
 * rt_create fails with RT_ERR_NO_DEVICE.
 *
 * Threading: calls on one rt_ctx must not be concurrent; distinct contexts are independent
 * (the reference keeps device pointers in __constant__ symbols, one scene per process).
 */
#ifndef RT_B200_H
#define RT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_B200_ABI_VERSION 3
#define RT_MAX_DEVICES 16

typedef enum rt_status {
    RT_OK = 0,
    RT_ERR_INVALID = -1,    /* bad argument */
    RT_ERR_IO = -2,         /* file missing / malformed */
    RT_ERR_NO_DEVICE = -3,  /* no usable CUDA device (there is no CPU fallback) */
    RT_ERR_CUDA = -4,       /* a CUDA runtime call failed; see rt_last_error */
    RT_ERR_NOMEM = -5,
    RT_ERR_STATE = -6       /* call order (e.g. download before render) */
} rt_status;

/* ------------------------------------------------------------------------------------
 * Host scene — the reference's L2 "scene prep" results, as plain arrays.
 * ------------------------------------------------------------------------------------ */

/* One BVH node exactly as the reference stores it (cpu/include/bvh.h:9-23, 32 bytes):
 * leaf iff tr_len > 0 (triangles tri_idx[idx .. idx+tr_len)); inner iff tr_len == 0 and
 * idx != 0 (children idx, idx+1); tr_len == 0 && idx == 0 is an empty leaf. */
typedef struct rt_bvh_node {
    float   min[3];
    float   max[3];
    int32_t tr_len;
    int32_t idx;
} rt_bvh_node;

/* Borrowed view of a scene in the reference's host layout.  All pointers are caller-owned
 * and only read during the call they are passed to. */
typedef struct rt_scene_desc {
    const float*       tri_coords;  /* n_tris x 9: v0 v1 v2 (triangle_t.coords, cpu/include/triangle.h:9) */
    const uint32_t*    tri_mat;     /* n_tris material indices (mat_idx, gpu/include/triangle.cuh) */
    uint32_t           n_tris;      /* triangle index = OBJ face order = first-hit ID space */
    const float*       materials;   /* n_mats x 9: ks kd kr (cpu/include/triangle.h:11-13) */
    uint32_t           n_mats;
    const float*       lights;      /* n_lights x 6: pos kl (cpu/include/light.h:8-11) */
    uint32_t           n_lights;
    float              ambient[3];  /* amb_light, cpu/src/main.c:37 */
    const rt_bvh_node* bvh;         /* bvh_len nodes as built by bvh_build (cpu/src/bvh.c:360-388) */
    const int32_t*     tri_idx;     /* n_tris leaf-order permutation (cpu/src/bvh.c:17) */
    uint32_t           bvh_len;
} rt_scene_desc;

/* Library-owned host scene: loaders + host BVH build (C++), replacing
 * triangles_load/lights_load (cpu/src/triangle.c:74-126, cpu/src/light.c:6-29) and
 * bvh_build (cpu/src/bvh.c:360-388). */
typedef struct rt_scene rt_scene;

/* Parse OBJ/MTL/lights files with the reference loader's rules.  Material
 * fields a .mtl block does not set are zero (the reference leaves them uninitialised). */
int rt_scene_load_obj(const char* obj_path, const char* mtl_path, const char* lights_path, rt_scene** out);
/* Same, from DIR/triangles.obj, DIR/triangles.mtl, DIR/lights.obj (cpu/src/main.c:113-114). */
int rt_scene_load_dir(const char* dir, rt_scene** out);
/* Binary scene pack ".rtsc": "RTSC0001", u32 n_tris,n_mats,n_lights,0, f32 ambient[3],0,
 * f32 tri[n_tris][9], u32 mat_idx[n_tris], f32 mats[n_mats][9], f32 lights[n_lights][6]. */
int rt_scene_load_rtsc(const char* path, rt_scene** out);
int rt_scene_save_rtsc(const rt_scene* s, const char* path);
/* Copy caller arrays (desc->bvh may be NULL). */
int rt_scene_from_arrays(const rt_scene_desc* desc, rt_scene** out);
/* The reference's synthetic triangle soup (cpu/src/main.c:115-131) driven by libc
 * srand(seed)/rand(): a in [-5,5)^3, b = a+U[0,1)^3, c = b+U[0,1)^3, ks=1, kd=kr=0, no lights. */
int rt_scene_soup(uint32_t n_tris, uint32_t seed, rt_scene** out);
/* base instanced on an nx*ny*nz grid with the given pitch (scale kept: SURVEY.md §A.3b);
 * lights are replicated every `light_every` instances (0 = keep base lights only). */
int rt_scene_instance_grid(const rt_scene* base, uint32_t nx, uint32_t ny, uint32_t nz,
                           const float pitch[3], uint32_t light_every, rt_scene** out);
/* Host BVH build reproducing the reference tree node for node.  heuristic: 6 (binned
 * squared-diagonal "SAH", 32 bins; gpu/include/options.cuh:50), 0 or 1 (spatial median,
 * cpu/src/bvh.c:214-223).  Replaces any previous tree of the scene.
 * Arithmetic: by default the IEEE reading of the reference source — which is what the
 * reference GPU program's host code computes (nvcc passes no fast-math to the host compiler).
 * OR-ing RT_BVH_REFBIN selects the four contracted expressions gcc 13 emits for the reference
 * CPU program under its makefile flags (-O3 -ffast-math -march=native), reproducing THAT
 * binary's tree node for node (it differs from the IEEE tree in ~2 % of the nodes; the image
 * does not depend on which is used).  See csrc/bvh_build.cpp. */
#define RT_BVH_REFBIN 0x100
int rt_scene_build_bvh(rt_scene* s, int heuristic);
/* The same tree built on a GPU (csrc/bvh_build_gpu.cu): heuristic 6 only (optionally | RT_BVH_REFBIN), node for node
 * and tri_idx entry for entry the result of rt_scene_build_bvh.  Degenerate inputs that trip the reference's
 * bvh_len >= 2N guard (cpu/src/bvh.c:80-83) are handed to the host builder (stats->fell_back = 1). */
typedef struct rt_bvh_gpu_stats {
    float    total_ms, upload_ms, top_ms, subtree_ms, assemble_ms, download_ms;
    int32_t  levels;     /* level-synchronous passes over the top of the tree */
    int32_t  top_nodes;  /* nodes created by those passes */
    int32_t  subtrees;   /* subtrees finished by one thread each */
    int32_t  fell_back;
    uint32_t nodes;      /* bvh_len */
} rt_bvh_gpu_stats;
int rt_scene_build_bvh_gpu(rt_scene* s, int heuristic, int device, rt_bvh_gpu_stats* stats_out);
/* Fill a borrowed view (valid until the scene is changed or freed). */
int rt_scene_view(const rt_scene* s, rt_scene_desc* out);
void rt_scene_free(rt_scene* s);

/* ------------------------------------------------------------------------------------
 * Device context — replaces load_to_gpu / render_frame / load_from_gpu.
 * ------------------------------------------------------------------------------------ */
typedef struct rt_ctx rt_ctx;

/* Camera with cam_init semantics (cpu/src/cam.c:5-9, cpu/src/main.c:105-106): fov is the
 * full field-of-view angle in radians as passed to cam_init; rot is Euler Y->X->Z. */
typedef struct rt_camera {
    float pos[3];
    float rot[3];
    float fov;
} rt_camera;

enum {
    RT_MODE_FAST = 0,   /* FMA + reciprocal arithmetic, tolerance-checked against the oracle */
    RT_MODE_STRICT = 1  /* IEEE op-for-op restatement: bit-identical to oracle/rt_oracle.c */
};

enum { /* rt_render_params.aov_mask */
    RT_AOV_RGB_F32 = 1,  /* clamped float colour, 3 floats/pixel (the reference's pixels[]) */
    RT_AOV_TRI_ID = 2,   /* first-hit triangle index of sample 0, -1 = miss */
    RT_AOV_DEPTH = 4,    /* first-hit t of sample 0 (FLT_MAX = miss) */
    RT_AOV_WORK = 8      /* count inner-node visits and triangle tests (slower build of the kernel) */
};

enum { /* rt_render_params.traversal */
    RT_TRAVERSAL_DEFAULT = 0,
    RT_TRAVERSAL_PLAIN = 1,       /* while-while in the reference's visit order */
    RT_TRAVERSAL_SPECULATIVE = 2, /* 2-wide tree, postponed leaves: same image, more lanes busy */
    RT_TRAVERSAL_WIDE = 3,        /* 4-wide tree made from the reference tree + postponed leaves */
    RT_TRAVERSAL_WIDE8 = 4        /* compressed 8-wide collapse (96-byte nodes, 8-bit outward-rounded child boxes, csrc/wide8.h) */
};

enum { /* rt_render_params.gather */
    RT_GATHER_PEER_STORE = 0, /* fused: every device stores finished pixels straight into device 0's frame */
    RT_GATHER_PEER_COPY = 1   /* unfused: local frame, then packed tile copy device->device 0 + unpack */
};

/* Frame sequences (the reference's ITERATIONS loop, cpu/src/main.c:169-185, gpu/src/main.cu:111-114): a context
 * owns RT_FRAME_SLOTS device frames, so that frame k+1 can render while frame k is copied to the host. */
#define RT_FRAME_SLOTS 2
enum { /* rt_render_params.frame_flags */
    RT_FRAME_BOTTOM_UP = 1 /* store the BGRA rows bottom-up — the BMP row order (cpu/src/bmp_writer.c:122-146) — so that
                              the downloaded buffer IS the BMP pixel array and no host-side row flip is left.  AOVs stay
                              top-down.  Not available with RT_GATHER_PEER_COPY / packed tiles. */
};

typedef struct rt_render_params {
    rt_camera cam;
    int32_t   width, height;
    int32_t   spp;          /* >= 1; sample 0 is the reference's pixel-corner ray (include/rt_sampling.h) */
    uint32_t  seed;
    int32_t   bounces;      /* BOUNCES, cpu/include/options.h:52 (4) */
    int32_t   mode;         /* RT_MODE_* */
    int32_t   aov_mask;     /* RT_AOV_* ; the 8-bit BGRA frame is always produced */
    int32_t   gather;       /* RT_GATHER_* (only meaningful with > 1 device in the context) */
    /* Image partition for one-process-per-GPU use: this context renders only the tiles of
     * part `part_index` out of `part_count` (interleaved 16x8 tiles, see rt_tile_owner).
     * part_count <= 1 renders the whole image (split over the context's own devices). */
    int32_t   part_index, part_count;
    /* tuning knobs (0 = library default) */
    int32_t   block_threads;    /* threads per CTA */
    int32_t   ctas_per_sm;      /* persistent CTAs per SM */
    int32_t   refill_threshold; /* leave the traversal loop when fewer lanes than this are active; 0 = default (fast build: 14 / 16 on
                                   the 4- / 2-wide tree, 8 on chain-bound frames; strict build: 20) */
    int32_t   traversal;        /* RT_TRAVERSAL_* (fast mode only; strict always walks the reference order) */
    int32_t   frame_flags;      /* RT_FRAME_* */
    int32_t   frame_slot;       /* which of the context's RT_FRAME_SLOTS device frames to render into (frame sequences) */
    /* fast build on the compressed 8-wide tree (RT_TRAVERSAL_WIDE8), both off unless > 0: */
    int32_t   drain_k;          /* once the chunk queue is empty, a warp left with <= drain_k live pixels hands them to the
                                   cooperative drain kernel (eight lanes per ray) */
    int32_t   cull;             /* test every 8x4-pixel chunk's ray pyramid against the top of the tree before tracing it */
    /* fast build: 0 = heaviest tiles first — every frame records its per-pixel traversal cost, and the next frame of the same
     * shape (size, spp, partition) on this context renders its 16x8 tiles in the order of their most expensive pixel (frame
     * sequences are coherent; a wrong guess only costs time, the bytes of a pixel do not depend on when it is rendered);
     * < 0 = spatial tile order only */
    int32_t   schedule;
    int32_t   reserved;
} rt_render_params;

typedef struct rt_timing {
    float    kernel_ms[RT_MAX_DEVICES]; /* render kernel, CUDA events on the launching stream */
    float    gather_ms;                 /* frame assembly on device 0 (0 for fused peer stores) */
    float    total_ms;                  /* first launch -> frame complete on device 0 (events) */
    uint64_t rays_closest;              /* bvh_traverse calls (cpu/src/raytracer.c:113) */
    uint64_t rays_shadow;               /* bvh_light_traverse calls (cpu/src/raytracer.c:74) */
    uint64_t inner_visits;              /* RT_AOV_WORK only */
    uint64_t tri_tests;                 /* RT_AOV_WORK only */
    uint32_t launches;                  /* kernels launched by this call */
    uint32_t n_devices;
} rt_timing;

void rt_render_params_default(rt_render_params* p); /* reference defaults: 1920x1080, camera of main.c:105-106 */

/* load_to_gpu: flatten + compress the scene and upload it to every listed device.
 * devices == NULL / ndev == 0 means device 0 only.  desc->bvh must be present. */
int rt_create(const rt_scene_desc* desc, const int* devices, int ndev, rt_ctx** out);
/* The same context without the tree visiting the host: the heuristic-6 BVH (optionally | RT_BVH_REFBIN) is built on
 * devices[0], flattened there and fanned out to the other devices over NVLink.  download_tree != 0 also fills the host
 * scene's bvh / tri_idx (as rt_scene_build_bvh_gpu does); otherwise the host scene is left without a tree.  Renders are
 * identical to rt_scene_build_bvh + rt_create. */
int rt_create_gpu(rt_scene* s, int heuristic, const int* devices, int ndev, int download_tree, rt_ctx** out, rt_bvh_gpu_stats* stats_out);
/* render_frame: blocking; renders into device-resident frame(s). */
int rt_render(rt_ctx* ctx, const rt_render_params* params, rt_timing* timing_out);
/* load_from_gpu: copy the last rendered frame (of the slot rendered last) to host; blocking.  bgra: W*H*4 bytes, row 0 = top, bytes
 * B,G,R,255 exactly as vec_to_bgra (cpu/src/bmp_writer.c:88-95).  The other outputs are
 * optional (NULL) and require the matching RT_AOV_* bit in the last rt_render. */
int rt_download(rt_ctx* ctx, uint8_t* bgra, float* rgb, int32_t* tri_id, float* depth_t);
void rt_destroy(rt_ctx* ctx);
/* Message of the last failure on this context (ctx == NULL: of the calling thread). */
const char* rt_last_error(const rt_ctx* ctx);

/* -------- frame sequences: non-blocking render + overlapped device->host copy --------
 * rt_render_async queues a frame on params->frame_slot and returns without waiting for the device (with
 * RT_GATHER_PEER_COPY the gather still blocks).  rt_download_async queues the copy of that slot's BGRA frame to
 * `host_bgra` (W*H*4 bytes; pinned memory — rt_host_alloc — makes it a true asynchronous copy) on a second
 * stream, ordered after the slot's render on every device of the context.  rt_frame_wait blocks until everything
 * queued on the slot so far (render, then copy) is complete and returns the render's timing.  A slot must be
 * waited on before it is rendered into again; a copy still pending on it is ordered before the new render.
 *     rt_render(ctx, p, tm)  ==  rt_render_async(ctx, p); rt_frame_wait(ctx, p->frame_slot, tm)            */
int rt_render_async(rt_ctx* ctx, const rt_render_params* params);
int rt_download_async(rt_ctx* ctx, int slot, uint8_t* host_bgra);
int rt_frame_wait(rt_ctx* ctx, int slot, rt_timing* timing_out);
/* Page-locked host memory for rt_download / rt_download_async targets (cudaHostAlloc / cudaFreeHost). */
int rt_host_alloc(size_t bytes, void** out);
void rt_host_free(void* p);

/* -------- one-process-per-GPU frame assembly (torch.distributed / NCCL plumbing) -------- */
/* Tile geometry shared by the kernel, the pack/unpack kernels and the host. */
#define RT_TILE_W 16
#define RT_TILE_H 8
/* owner part of tile (tx,ty): diagonal interleave */
#if defined(__CUDACC__)
#define RT_INLINE_HD static inline __host__ __device__
#else
#define RT_INLINE_HD static inline
#endif
RT_INLINE_HD int rt_tile_owner(int tx, int ty, int part_count) { return part_count > 1 ? (tx + ty) % part_count : 0; }
/* Number of tiles part `part` owns in a width x height image. */
int rt_part_tile_count(int width, int height, int part_index, int part_count);
/* Device pointer (on the context's first device) of the packed BGRA tiles of the last
 * render: rt_part_tile_count(...) * RT_TILE_W*RT_TILE_H*4 bytes, tiles in owner order. */
int rt_packed_tiles(rt_ctx* ctx, void** dev_ptr, size_t* bytes);
/* Scatter `part_count` packed buffers (concatenated, each `stride_bytes` long, device
 * memory on the context's first device) into the context's BGRA frame. */
int rt_unpack_tiles(rt_ctx* ctx, const void* dev_gathered, size_t stride_bytes, int part_count);
/* CUDA IPC: export the BGRA frame of this context (64-byte handle) / make another
 * process's frame the peer-store target of this context's renders. */
int rt_frame_ipc_export(rt_ctx* ctx, int width, int height, void* handle64);
int rt_frame_ipc_import(rt_ctx* ctx, const void* handle64, int width, int height);
/* Same for frame slot `slot` (the two calls above are slot 0). */
int rt_frame_ipc_export_slot(rt_ctx* ctx, int slot, int width, int height, void* handle64);
int rt_frame_ipc_import_slot(rt_ctx* ctx, int slot, const void* handle64, int width, int height);
/* Diagnostics: per-warp timeline of RT_AOV_WORK renders (8 x u64 per warp, see csrc/rt_api.cu). */
int rt_debug_warp_trace(rt_ctx* ctx, int enable, unsigned long long* out, int max_warps);
/* Diagnostics: replace the first device's tile order (tile ids, ty * tiles_x + tx) for the current frame shape. */
int rt_debug_set_tile_order(rt_ctx* ctx, const unsigned* tiles, int n);
/* Diagnostics, host only (no device needed): the staging arrays rt_create would upload (csrc/device_layout.h).  which: 0 nodes
 * (64 B per inner node), 1 nodes4 (128 B), 2 tris (64 B per leaf-order slot), 3 shade (16 B), 4 leaf_cnt, 5 mats, 6 lights, 7 nodes8
 * (96 B per compressed 8-wide node, csrc/wide8.h).
 * out may be NULL to query the size; meta4 receives {max_depth, stack_need4, n_lights, depth8}. */
int rt_debug_flatten_host(const rt_scene_desc* desc, int which, void* out, size_t cap_bytes, size_t* bytes_out, int* meta4);
/* Diagnostics: copy of a scene array as it lies on the context's first device (same selectors; 0 nodes, 1 nodes4, 2 tris, 3 shade,
 * 7 nodes8): lets a test compare the device-side flatten of rt_create_gpu with the host staging arrays byte for byte. */
int rt_debug_device_array(rt_ctx* ctx, int which, void* out, size_t cap_bytes, size_t* bytes_out);
/* Roofline microbenchmark: GB/s of random 64-byte-record gathers (the shape of a node fetch, one record per
 * lane) from a working set of ws_bytes on `device` — L1-, L2- or HBM-resident depending on the size. */
int rt_debug_gather_bandwidth(int device, size_t ws_bytes, float* gbs_out);
/* Diagnostics: the per-pixel traversal-cost map behind rt_render_params.schedule (width*height u16 of the last fast frame on the
 * first device: traversal steps, saturating; pixels of other ranks' tiles are undefined) and hdr2 = {tiles ordered ahead of the
 * cheapest class, tiles of this rank}. */
int rt_debug_cost_map(rt_ctx* ctx, unsigned short* out, size_t n_pixels, unsigned* hdr2);
/* Diagnostics: the first device's tile list (tile = ty * tiles_x + tx of 16x8-pixel tiles) — which = 0: the partition's list in
 * its base order, 1: the order the next frame of the same shape will render in (heaviest tiles first; RT_ERR_STATE if there is
 * no cost history).  Writes min(n, cap) entries, *n_out = n. */
int rt_debug_tile_order(rt_ctx* ctx, int which, unsigned* out, int cap, int* n_out);
/* Diagnostics: host -> device copy rate (GB/s) of `bytes` of pageable memory: mode 0 plain cudaMemcpy (the reference's way,
 * gpu/src/gpu.cu:143-175), 1 the library's staged copy through its pinned ring (csrc/staged_copy.h), 2 cudaMemcpy from
 * page-locked memory (the ceiling on this box). */
int rt_debug_copy_bandwidth(int device, size_t bytes, int mode, float* gbs_out);
/* Raw device pointer of the BGRA frame (device 0 of the context). */
int rt_frame_device_ptr(rt_ctx* ctx, void** dev_ptr, size_t* bytes);

/* -------- image output (cpu/src/bmp_writer.c:177-211) -------- */
/* 54-byte header, 32 bpp, bottom-up rows; input is the top-down BGRA of rt_download. */
int rt_write_bmp(const char* path, const uint8_t* bgra_top_down, int width, int height);
/* Same file from a frame whose rows are already bottom-up (RT_FRAME_BOTTOM_UP): header + one write, no flip. */
int rt_write_bmp_bottom_up(const char* path, const uint8_t* bgra_bottom_up, int width, int height);

int rt_abi_version(void);
/* number of visible CUDA devices (0 when there is none / no driver) */
int rt_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* RT_B200_H */

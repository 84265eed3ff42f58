// rt_api.cu — the C-ABI device context (include/rt_b200.h): rt_create / rt_render / rt_download /
// rt_destroy and the frame-assembly helpers.  Host side of the reference GPU program's
// load_to_gpu / render_frame / load_from_gpu (gpu/src/gpu.cu:98-228), made re-entrant,
// error-checked (the reference checks one call, gpu.cu:121-124) and multi-device.
//
// Device work launched from here: the render kernel (render_fast.cu / render_strict.cu) and three
// small frame kernels below (fill, pack, unpack).  There is no CPU rendering path in this library.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "camera.h"
#include "device_layout.h"
#include "flatten.h"
#include "gpu_tree.h"
#include "host_scene.h"
#include "render_launch.h"
#include "staged_copy.h"

namespace {

constexpr size_t RT_CTRL_WORDS = 8 + RT_MAX_SMS;

struct Dev {
    int id = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    // second stream: the per-frame statistics copy and (device 0) the device->host copies of finished frames run here, so
    // that the next frame's kernel never queues behind a copy (one D2H engine serves both: a 32-byte statistics copy in
    // the render stream was measured to wait ~0.08 ms per frame behind the 8 MB frame copy of the previous frame)
    cudaStream_t aux = nullptr;
    // per frame slot: kernel start / kernel end / frame's device work complete (stats copied); ev2: gather (PEER_COPY)
    cudaEvent_t ev0[RT_FRAME_SLOTS] = {}, ev1[RT_FRAME_SLOTS] = {}, ev_done[RT_FRAME_SLOTS] = {}, ev2 = nullptr;
    // scene (replicated per device)
    float4 *nodes = nullptr, *nodes4 = nullptr, *tris = nullptr, *shade = nullptr, *mats = nullptr, *lights = nullptr;
    uint4* nodes8 = nullptr; // compressed 8-wide tree (wide8.h); null when the context has none
    size_t bytes[8] = {};    // sizes of the scene arrays, by rt_debug_flatten_host selector
    int* leaf_cnt = nullptr;
    // per-frame control block, one per frame slot (RT_CTRL_WORDS u64 each): [0..3] stats, [4] tile counter, [8..] SM cursors
    unsigned long long* ctrl = nullptr;
    unsigned long long* ctrl_host = nullptr; // pinned, 8 words per frame slot
    // tile list of this device for the current (w, h, parts), row-major: the order tiles are rendered in and
    // the layout of packed tile buffers
    unsigned* tile_list = nullptr;
    int n_tiles = 0;
    int tile_epoch = 0, tile_epoch_seen = -1; // bumped whenever tile_list changes: an order made from another list is void
    size_t tile_cap = 0;
    // local frame of devices > 0 in PEER_COPY mode (the assembled frames live in rt_ctx::slots, on device 0)
    uchar4* bgra = nullptr;
    size_t bgra_px = 0;
    uchar4* packed = nullptr; // packed tiles (gather paths)
    size_t packed_px = 0;
    bool peer_to_0 = false;
    // heaviest-tiles-first scheduling (fast build): per-pixel traversal cost of the last frame, the tile list ordered by it
    // (what the next frame of the same shape renders from), per-tile classes, two count/cursor headers (ping-pong), per-block
    // class counts of the ordering kernels, and what all of it belongs to
    unsigned short* cost = nullptr;
    size_t cost_px = 0;
    unsigned* tile_sorted = nullptr;
    unsigned char* tile_cls = nullptr;
    size_t sorted_cap = 0;
    unsigned* cost_hdr = nullptr;
    unsigned* cost_blk = nullptr;
    int cost_cur = 0;
    bool cost_valid = false;
    // heaviest pixel and total steps of the last finished frame that recorded them, for the occupancy of chain-bound frames
    bool slot_stat[RT_FRAME_SLOTS] = {};
    int slot_key[RT_FRAME_SLOTS][5] = {};
    unsigned long long stat_max = 0, stat_total = 0;
    int stat_key[5] = {0, 0, 0, 0, 0};
    int cost_key[5] = {0, 0, 0, 0, 0}; // width, height, spp, part_index, part_count
    RtPathRec* drain_queue[RT_FRAME_SLOTS] = {}; // tail hand-off queue per frame slot (render_kernel.cuh: drain_kernel)
    size_t drain_cap[RT_FRAME_SLOTS] = {};
    unsigned long long* warp_trace = nullptr; // diagnostics (rt_debug_warp_trace)
    size_t warp_trace_cap = 0;
    int warp_trace_n = 0;
};

} // namespace

namespace {
// One device frame of a frame sequence (RT_FRAME_SLOTS per context, on device 0) and the state of the render queued on it.
struct Slot {
    uchar4* bgra = nullptr;
    size_t bgra_px = 0;
    // AOVs of the frame rendered on this slot (device 0, peer-stored by the other devices like the BGRA frame): one set per
    // slot, so that a render queued on the other slot cannot overwrite what rt_download returns for this one
    float* rgb = nullptr;
    int* tri_id = nullptr;
    float* depth = nullptr;
    size_t aov_px[3] = {0, 0, 0};
    uchar4* ipc_frame = nullptr; // another process's frame (CUDA IPC): the peer-store target of this slot's renders
    int ipc_w = 0, ipc_h = 0;
    cudaEvent_t copy_done = nullptr;
    bool render_pending = false, copy_pending = false, rendered = false;
    int width = 0, height = 0, aov_mask = 0, spp = 0, part_index = 0, part_count = 1, flags = 0;
    float gather_ms = 0.f;
    unsigned launches = 0;
    rt_timing timing{};
};
} // namespace

struct rt_ctx {
    std::vector<Dev> devs;
    Slot slots[RT_FRAME_SLOTS];
    int last_slot = 0;
    cudaStream_t copy_stream = nullptr; // device 0: device->host copies of finished frames, overlapping the next render
    RtDeviceScene scene_host_view{}; // n_lights / amb only; pointers are per device
    std::string err;
    // state of the last finished render (rt_download, rt_packed_tiles, ... refer to it)
    int width = 0, height = 0, aov_mask = 0;
    bool rendered = false;
    int part_index = 0, part_count = 1;
    // tile-list cache key
    int tl_w = 0, tl_h = 0, tl_parts = 0, tl_index = -1;
    // device 0 helpers for unpack
    unsigned* local_index = nullptr; // per tile: index inside its owner's packed buffer
    size_t local_index_cap = 0;
    int li_w = 0, li_h = 0, li_parts = 0;
    uchar4* gather_buf = nullptr; // in-process PEER_COPY landing zone on device 0
    size_t gather_px = 0;
    size_t scene_bytes = 0;
    int max_depth = 0, stack_need4 = 0, depth8 = 0;
    bool want_trace = false;
};

namespace {

#define CK(ctx, call)                                                                                      \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess) {                                                                          \
            (ctx)->err = std::string(#call) + " failed: " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" + \
                         std::to_string(__LINE__) + ")";                                                   \
            rt::set_error((ctx)->err);                                                                     \
            return RT_ERR_CUDA;                                                                            \
        }                                                                                                  \
    } while (0)

int fail(rt_ctx* c, int code, const std::string& msg)
{
    if (c) c->err = msg;
    rt::set_error(msg);
    return code;
}

template <class T>
cudaError_t upload(T** dst, const void* src, size_t bytes, cudaStream_t st)
{
    if (!bytes) { *dst = nullptr; return cudaSuccess; }
    cudaError_t e = cudaMalloc((void**)dst, bytes);
    if (e != cudaSuccess) return e;
    return rt::staged_h2d(*dst, src, bytes, st); // pageable staging vectors -> pinned ring -> device (staged_copy.h)
}

// ------------------------------------------------------------------ frame kernels
__global__ void fill_bgra_kernel(uchar4* dst, size_t n, uchar4 v)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = v;
}

// one CTA of RT_TILE_PIXELS threads per owned tile: frame -> packed (tile-major, row-major inside)
__global__ void pack_tiles_kernel(const uchar4* __restrict__ frame, uchar4* __restrict__ packed,
                                  const unsigned* __restrict__ tile_list, int n_tiles, int tiles_x, int width, int height)
{
    const int i = blockIdx.x;
    if (i >= n_tiles) return;
    const unsigned t = tile_list[i];
    const int x = (int)(t % (unsigned)tiles_x) * RT_TILE_W + (threadIdx.x % RT_TILE_W);
    const int y = (int)(t / (unsigned)tiles_x) * RT_TILE_H + (threadIdx.x / RT_TILE_W);
    uchar4 v = make_uchar4(0, 0, 0, 0);
    if (x < width && y < height) v = frame[(size_t)y * width + x];
    packed[(size_t)i * RT_TILE_PIXELS + threadIdx.x] = v;
}

// one thread per pixel: gathered packed buffers -> frame
__global__ void unpack_tiles_kernel(uchar4* __restrict__ frame, const uchar4* __restrict__ gathered, size_t stride_px,
                                    const unsigned* __restrict__ local_index, int parts, int tiles_x, int width, int height)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= width || y >= height) return;
    const int tx = x / RT_TILE_W, ty = y / RT_TILE_H;
    const int owner = rt_tile_owner(tx, ty, parts);
    const unsigned li = local_index[ty * tiles_x + tx];
    const int p = (y % RT_TILE_H) * RT_TILE_W + (x % RT_TILE_W);
    frame[(size_t)y * width + x] = gathered[(size_t)owner * stride_px + (size_t)li * RT_TILE_PIXELS + p];
}

// ------------------------------------------------------------------ heaviest tiles first
// The frame time of a small frame is the dependent chain of its heaviest pixels (8 rays x hundreds of traversal steps) counted
// from the moment they START (profiles/r02_notes.md §6): a heavy pixel fetched when the chunk queue is nearly empty ends the frame
// a fifth of a millisecond later, alone on its SM.  Frame sequences are coherent, so the render kernel records every pixel's
// traversal steps (RtFrameArgs::cost_out) and the two kernels below order the NEXT frame's tile list by them: tiles sorted by the
// cost CLASS of their heaviest pixel (half octaves: < 32 steps, 32-47, 48-63, 64-95, ...), heaviest class first, original
// (spatial) order within a class — longest-processing-time-first scheduling by counting sort.  Tiles stay whole (16x8 pixels,
// four 8x4 chunks), so warps keep rendering neighbouring pixels.  Scheduling only: a pixel's bytes do not depend on when or
// where it is rendered.
//   header (RT_COST_HDR words, two of them, ping-pong): [k] tiles of class k.  Blocks take contiguous ranges of the tile
//   list; pass 1 leaves every tile's class in `cls`, every block's class counts in `blk` and the totals in the header; pass 2
//   places every block's tiles of a class behind those of the blocks before it — a stable counting sort.
constexpr int RT_COST_CLASSES = 21;
constexpr int RT_COST_HDR = 64;
constexpr int RT_COST_BLOCK = 1024; // few large blocks: short per-warp loops, and a short prefix over the blocks before

__device__ __forceinline__ int cost_class(unsigned v)
{
    if (v < 32u) return 0;
    const int e = 31 - __clz(v);                 // 5 .. 15
    return 1 + 2 * (e - 5) + (int)((v >> (e - 1)) & 1u); // <= 20 for v <= 49151; the map saturates at 65535
}

__global__ void __launch_bounds__(RT_COST_BLOCK) tile_class_kernel(const unsigned short* __restrict__ cost, int width, int height, int tiles_x,
                                                                    const unsigned* __restrict__ tile_list, int n_tiles, unsigned char* __restrict__ cls,
                                                                    unsigned* __restrict__ hdr, unsigned* __restrict__ hdr_old, unsigned* __restrict__ blk,
                                                                    unsigned long long* __restrict__ stat)
{
    __shared__ unsigned s_cnt[RT_COST_CLASSES];
    __shared__ unsigned s_max;
    __shared__ unsigned long long s_sum;
    // the header the placement kernel of the previous frame has finished with becomes the next frame's: zero it here (saves
    // a memset per frame; nothing touches it before the next frame's kernels)
    if (blockIdx.x == 0 && threadIdx.x < RT_COST_HDR) hdr_old[threadIdx.x] = 0u;
    if (threadIdx.x < RT_COST_CLASSES) s_cnt[threadIdx.x] = 0u;
    if (threadIdx.x == 0) { s_max = 0u; s_sum = 0ull; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, per = (n_tiles + gridDim.x - 1) / gridDim.x;
    const int k0 = blockIdx.x * per, k1 = min(k0 + per, n_tiles);
    unsigned long long w_sum = 0; // (lane 0 of each warp)
    unsigned w_max = 0;
    for (int k = k0 + warp; k < k1; k += RT_COST_BLOCK / 32) { // one 16x8 tile per warp: lane = row (lane >> 2), 4 columns from (lane & 3) * 4
        const unsigned tile = tile_list[k];
        const int y = (int)(tile / (unsigned)tiles_x) * RT_TILE_H + (lane >> 2), x0 = (int)(tile % (unsigned)tiles_x) * RT_TILE_W + ((lane & 3) << 2);
        unsigned m = 0, s = 0;
        if (y < height)
            for (int j = 0; j < 4; j++)
                if (x0 + j < width) { const unsigned v = cost[(size_t)y * width + x0 + j]; m = max(m, v); s += v; }
        m = __reduce_max_sync(0xffffffffu, m);
        s = __reduce_add_sync(0xffffffffu, s);
        w_sum += s; w_max = max(w_max, m);
        if (lane == 0) {
            const int c = min(cost_class(m), RT_COST_CLASSES - 1);
            cls[k] = (unsigned char)c;
            atomicAdd(&s_cnt[c], 1u);
        }
    }
    if (lane == 0 && w_max) { atomicMax(&s_max, w_max); atomicAdd(&s_sum, w_sum); }
    __syncthreads();
    if (threadIdx.x < RT_COST_CLASSES) {
        const unsigned n = s_cnt[threadIdx.x];
        blk[blockIdx.x * RT_COST_CLASSES + threadIdx.x] = n;
        if (n) atomicAdd(&hdr[threadIdx.x], n);
    }
    if (stat && threadIdx.x == 0 && s_max) { atomicMax(&stat[0], (unsigned long long)s_max); atomicAdd(&stat[1], s_sum); }
}

__global__ void __launch_bounds__(RT_COST_BLOCK) tile_place_kernel(const unsigned* __restrict__ tile_list, int n_tiles, const unsigned char* __restrict__ cls,
                                                                    const unsigned* __restrict__ hdr, const unsigned* __restrict__ blk, unsigned* __restrict__ sorted)
{
    __shared__ unsigned s_base[RT_COST_CLASSES];
    __shared__ unsigned s_wcnt[RT_COST_BLOCK / 32][RT_COST_CLASSES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, per = (n_tiles + gridDim.x - 1) / gridDim.x;
    const int k0 = blockIdx.x * per, k1 = min(k0 + per, n_tiles);
    // where this block's tiles of class c go: after all tiles of heavier classes and after this class's tiles of the blocks
    // before this one (their counts summed here: no cursor, the order is the list's order and the same on every run)
    if (threadIdx.x < RT_COST_CLASSES) s_base[threadIdx.x] = 0u;
    __syncthreads();
    if (lane < RT_COST_CLASSES) { // lane = class, warp w sums the rows w, w + 8, ... of the blocks before this one
        unsigned n = 0;
        for (int b = warp; b < (int)blockIdx.x; b += RT_COST_BLOCK / 32) n += blk[b * RT_COST_CLASSES + lane];
        if (n) atomicAdd(&s_base[lane], n);
    }
    __syncthreads();
    if (threadIdx.x < RT_COST_CLASSES) {
        unsigned above = 0;
        for (int j = threadIdx.x + 1; j < RT_COST_CLASSES; j++) above += hdr[j];   // heavier classes come first
        s_base[threadIdx.x] += above;
    }
    for (int b = k0; b < k1; b += RT_COST_BLOCK) { // one block-full of list entries at a time, kept in list order within a class
        for (int i = threadIdx.x; i < (RT_COST_BLOCK / 32) * RT_COST_CLASSES; i += RT_COST_BLOCK) (&s_wcnt[0][0])[i] = 0u;
        __syncthreads();
        const int k = b + threadIdx.x;
        const int c = k < k1 ? (int)cls[k] : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, c);
        if (c >= 0 && lane == __ffs(peers) - 1) s_wcnt[warp][c] = (unsigned)__popc(peers);
        __syncthreads();
        if (c >= 0) {
            unsigned pos = s_base[c] + (unsigned)__popc(peers & ((1u << lane) - 1u));
            for (int w = 0; w < warp; w++) pos += s_wcnt[w][c];
            sorted[pos] = tile_list[k];
        }
        __syncthreads();
        if (threadIdx.x < RT_COST_CLASSES) {
            unsigned n = 0;
            for (int w = 0; w < RT_COST_BLOCK / 32; w++) n += s_wcnt[w][threadIdx.x];
            s_base[threadIdx.x] += n;
        }
        __syncthreads();
    }
}

// Microbenchmark for the roofline: every lane gathers its own random 64-byte record (two 256-bit loads, the
// shape of an inner-node fetch) from a working set of `n_rec` records; four independent gathers in flight per lane.
__global__ void gather64_kernel(const float4* __restrict__ data, unsigned n_rec_mask, int iters, unsigned seed, float4* sink)
{
    unsigned x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + seed;
    float acc = 0.f;
    for (int i = 0; i < iters; i += 4) {
        float v[4][16];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            x ^= x << 13; x ^= x >> 17; x ^= x << 5; // xorshift32
            const float4* p = data + 4 * (size_t)(x & n_rec_mask);
            asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(v[u][0]), "=f"(v[u][1]), "=f"(v[u][2]), "=f"(v[u][3]), "=f"(v[u][4]), "=f"(v[u][5]), "=f"(v[u][6]), "=f"(v[u][7]) : "l"(p));
            asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(v[u][8]), "=f"(v[u][9]), "=f"(v[u][10]), "=f"(v[u][11]), "=f"(v[u][12]), "=f"(v[u][13]), "=f"(v[u][14]), "=f"(v[u][15]) : "l"(p + 2));
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int k = 0; k < 16; k++) acc += v[u][k];
    }
    if (acc == 123.456f) sink[0] = make_float4(acc, 0, 0, 0);
}

// ------------------------------------------------------------------ tiles
inline int tiles_x_of(int w) { return (w + RT_TILE_W - 1) / RT_TILE_W; }
inline int tiles_y_of(int h) { return (h + RT_TILE_H - 1) / RT_TILE_H; }

// Tiles of one part in rendering order: 2x2 blocks of tiles (32x16 pixels) in raster order, so that consecutive
// list entries are spatial neighbours.  The same order defines the layout of packed tile buffers.
void make_tile_list(int w, int h, int part, int parts, std::vector<unsigned>& out)
{
    out.clear();
    const int tx_n = tiles_x_of(w), ty_n = tiles_y_of(h);
    for (int by = 0; by < ty_n; by += 2)
        for (int bx = 0; bx < tx_n; bx += 2)
            for (int ty = by; ty < std::min(by + 2, ty_n); ty++)
                for (int tx = bx; tx < std::min(bx + 2, tx_n); tx++)
                    if (rt_tile_owner(tx, ty, parts) == part) out.push_back((unsigned)(ty * tx_n + tx));
}

template <class T>
int ensure(rt_ctx* c, T** p, size_t* cap, size_t need)
{
    if (*cap >= need && *p) return RT_OK;
    if (*p) CK(c, cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    CK(c, cudaMalloc((void**)p, need * sizeof(T)));
    *cap = need;
    return RT_OK;
}

int setup_tiles(rt_ctx* c, int w, int h, int part_index, int part_count)
{
    const int nd = (int)c->devs.size();
    const int parts = part_count * nd;
    if (c->tl_w == w && c->tl_h == h && c->tl_parts == parts && c->tl_index == part_index) return RT_OK;
    std::vector<unsigned> tl;
    for (int d = 0; d < nd; d++) {
        Dev& D = c->devs[d];
        make_tile_list(w, h, part_index * nd + d, parts, tl);
        CK(c, cudaSetDevice(D.id));
        int rc = ensure(c, &D.tile_list, &D.tile_cap, std::max<size_t>(tl.size(), 1));
        if (rc) return rc;
        if (!tl.empty()) CK(c, cudaMemcpyAsync(D.tile_list, tl.data(), tl.size() * 4, cudaMemcpyHostToDevice, D.stream));
        CK(c, cudaStreamSynchronize(D.stream)); // tl is reused by the next iteration
        D.n_tiles = (int)tl.size();
        D.tile_epoch++;
    }
    c->tl_w = w; c->tl_h = h; c->tl_parts = parts; c->tl_index = part_index;
    return RT_OK;
}

int setup_local_index(rt_ctx* c, int w, int h, int parts)
{
    if (c->li_w == w && c->li_h == h && c->li_parts == parts && c->local_index) return RT_OK;
    const int tx_n = tiles_x_of(w), ty_n = tiles_y_of(h);
    std::vector<unsigned> li((size_t)tx_n * ty_n), tl;
    for (int p = 0; p < parts; p++) {
        make_tile_list(w, h, p, parts, tl);
        for (size_t i = 0; i < tl.size(); i++) li[tl[i]] = (unsigned)i;
    }
    Dev& D0 = c->devs[0];
    CK(c, cudaSetDevice(D0.id));
    int rc = ensure(c, &c->local_index, &c->local_index_cap, li.size());
    if (rc) return rc;
    CK(c, cudaMemcpy(c->local_index, li.data(), li.size() * 4, cudaMemcpyHostToDevice));
    c->li_w = w; c->li_h = h; c->li_parts = parts;
    return RT_OK;
}

void free_dev(Dev& D)
{
    cudaSetDevice(D.id);
    cudaFree(D.nodes); cudaFree(D.nodes4); cudaFree(D.nodes8); cudaFree(D.tris); cudaFree(D.shade); cudaFree(D.mats); cudaFree(D.lights);
    cudaFree(D.leaf_cnt); cudaFree(D.ctrl); cudaFree(D.tile_list); cudaFree(D.warp_trace);
    cudaFree(D.bgra); cudaFree(D.packed);
    for (int s = 0; s < RT_FRAME_SLOTS; s++) cudaFree(D.drain_queue[s]);
    cudaFree(D.cost); cudaFree(D.tile_sorted); cudaFree(D.tile_cls); cudaFree(D.cost_hdr); cudaFree(D.cost_blk);
    if (D.ctrl_host) cudaFreeHost(D.ctrl_host);
    for (int s = 0; s < RT_FRAME_SLOTS; s++) {
        if (D.ev0[s]) cudaEventDestroy(D.ev0[s]);
        if (D.ev1[s]) cudaEventDestroy(D.ev1[s]);
        if (D.ev_done[s]) cudaEventDestroy(D.ev_done[s]);
    }
    if (D.ev2) cudaEventDestroy(D.ev2);
    if (D.aux) cudaStreamDestroy(D.aux);
    if (D.stream) cudaStreamDestroy(D.stream);
}

} // namespace

extern "C" {

int rt_abi_version(void) { return RT_B200_ABI_VERSION; }

int rt_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* rt_last_error(const rt_ctx* ctx) { return ctx ? ctx->err.c_str() : rt::get_error(); }

void rt_render_params_default(rt_render_params* p)
{
    if (!p) return;
    std::memset(p, 0, sizeof *p);
    p->cam.pos[0] = 0; p->cam.pos[1] = -9; p->cam.pos[2] = 3;      // cpu/src/main.c:105
    p->cam.rot[0] = (float)(-3.14159265358979323846 / 12);          // cpu/src/main.c:106
    p->cam.fov = (float)(3.14159265358979323846 / 3.2);             // cpu/src/main.c:105
    p->width = 1920; p->height = 1080;                              // cpu/include/options.h:6-7
    p->spp = 1; p->seed = 1;
    p->bounces = 4;                                                 // cpu/include/options.h:52
    p->mode = RT_MODE_FAST;
    p->gather = RT_GATHER_PEER_STORE;
    p->part_index = 0; p->part_count = 1;
}

int rt_part_tile_count(int width, int height, int part_index, int part_count)
{
    if (width < 1 || height < 1 || part_count < 1 || part_index < 0 || part_index >= part_count) return RT_ERR_INVALID;
    const int tx_n = tiles_x_of(width), ty_n = tiles_y_of(height);
    int n = 0;
    for (int ty = 0; ty < ty_n; ty++)
        for (int tx = 0; tx < tx_n; tx++) n += rt_tile_owner(tx, ty, part_count) == part_index;
    return n;
}

// Shared tail of rt_create / rt_create_gpu: per-device streams and events, the scene arrays on every device.  `flat` holds
// the host staging arrays (always mats / lights; nodes, triangles ... only when `pre` is null); `pre` = arrays that are
// already resident on devices[0] (device-side flatten) and are adopted by the context.
static int create_common(const rt::FlatScene& flat, rt::DeviceFlat* pre, const int* devices, int ndev, rt_ctx** out)
{
    rt_ctx* c = new rt_ctx();
    c->scene_bytes = pre ? pre->bytes() + 4 * (flat.mats.size() + flat.lights.size()) : flat.bytes();
    c->max_depth = pre ? pre->max_depth : flat.max_depth;
    c->stack_need4 = pre ? pre->stack_need4 : flat.stack_need4;
    c->depth8 = pre ? pre->depth8 : flat.depth8;
    c->scene_host_view.n_lights = (int)flat.n_lights;
    std::memcpy(c->scene_host_view.amb, flat.ambient, 12);
    c->devs.resize(ndev);
    auto bail = [&](int code) { std::string m = c->err; rt_destroy(c); rt::set_error(m); return code; };
#define CKC(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { c->err = std::string(#call) + " failed: " + cudaGetErrorString(e__); return bail(RT_ERR_CUDA); } } while (0)
    const size_t b_nodes = pre ? 64 * pre->n_inner : flat.nodes.size() * 4, b_nodes4 = pre ? 128 * pre->n_nodes4 : flat.nodes4.size() * 4;
    const size_t b_tris = pre ? 64 * pre->n_tris : flat.tris.size() * 4, b_shade = pre ? 16 * pre->n_tris : flat.shade.size() * 4;
    const size_t b_leaf = pre ? (pre->leaf_cnt ? 4 * pre->n_tris : 0) : flat.leaf_cnt.size() * 4;
    const size_t b_nodes8 = pre ? 96 * pre->n_nodes8 : flat.nodes8.size() * 4;
    for (int i = 0; i < ndev; i++) {
        Dev& D = c->devs[i];
        D.id = devices[i];
        CKC(cudaSetDevice(D.id));
        cudaDeviceProp prop;
        CKC(cudaGetDeviceProperties(&prop, D.id));
        D.sm_count = prop.multiProcessorCount;
        CKC(cudaStreamCreateWithFlags(&D.stream, cudaStreamNonBlocking));
        CKC(cudaStreamCreateWithFlags(&D.aux, cudaStreamNonBlocking));
        for (int s = 0; s < RT_FRAME_SLOTS; s++) { CKC(cudaEventCreate(&D.ev0[s])); CKC(cudaEventCreate(&D.ev1[s])); CKC(cudaEventCreate(&D.ev_done[s])); }
        CKC(cudaEventCreate(&D.ev2));
        if (i == 0) {
            c->copy_stream = D.aux;
            for (int s = 0; s < RT_FRAME_SLOTS; s++) CKC(cudaEventCreateWithFlags(&c->slots[s].copy_done, cudaEventDisableTiming));
        }
        if (i > 0) { // NVLink / NVSwitch peer mapping towards the frame owner (also makes the fan-out a direct peer copy)
            int can = 0;
            CKC(cudaDeviceCanAccessPeer(&can, D.id, c->devs[0].id));
            if (can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(c->devs[0].id, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
                CKC(e);
                D.peer_to_0 = true;
            }
        }
        // One host->device upload (device 0), then fan-out device 0 -> device i over NVLink / NVSwitch: the other devices
        // do not pull the scene through PCIe again (the reference uploads once to its single device, gpu/src/gpu.cu:143-175).
        auto put = [&](auto** dst, const void* host, auto* const* src0, size_t bytes) -> cudaError_t {
            if (i == 0) return upload(dst, host, bytes, D.stream);
            if (!bytes) { *dst = nullptr; return cudaSuccess; }
            cudaError_t e = cudaMalloc((void**)dst, bytes);
            if (e != cudaSuccess) return e;
            return cudaMemcpyPeerAsync(*dst, D.id, *src0, c->devs[0].id, bytes, D.stream);
        };
        Dev& Z = c->devs[0];
        if (i == 0 && pre) { // adopt the arrays the device-side flatten left on this device
            D.nodes = pre->nodes; D.nodes4 = pre->nodes4; D.tris = pre->tris; D.shade = pre->shade; D.leaf_cnt = pre->leaf_cnt;
            D.nodes8 = pre->nodes8;
            pre->nodes = pre->nodes4 = pre->tris = pre->shade = nullptr; pre->leaf_cnt = nullptr; pre->nodes8 = nullptr;
        } else {
            CKC(put(&D.nodes, flat.nodes.data(), &Z.nodes, b_nodes));
            CKC(put(&D.nodes4, flat.nodes4.data(), &Z.nodes4, b_nodes4));
            CKC(put(&D.nodes8, flat.nodes8.data(), &Z.nodes8, b_nodes8));
            CKC(put(&D.tris, flat.tris.data(), &Z.tris, b_tris));
            CKC(put(&D.shade, flat.shade.data(), &Z.shade, b_shade));
            CKC(put(&D.leaf_cnt, flat.leaf_cnt.data(), &Z.leaf_cnt, b_leaf));
        }
        CKC(put(&D.mats, flat.mats.data(), &Z.mats, flat.mats.size() * 4));
        CKC(put(&D.lights, flat.lights.data(), &Z.lights, flat.lights.size() * 4));
        D.bytes[0] = b_nodes; D.bytes[1] = b_nodes4; D.bytes[2] = b_tris; D.bytes[3] = b_shade; D.bytes[4] = b_leaf; D.bytes[7] = b_nodes8;
        CKC(cudaMalloc((void**)&D.ctrl, 8 * RT_CTRL_WORDS * RT_FRAME_SLOTS));
        CKC(cudaMallocHost((void**)&D.ctrl_host, 64 * RT_FRAME_SLOTS));
        if (i == 0) CKC(cudaStreamSynchronize(D.stream)); // the fan-out below reads device 0's arrays
    }
    // the fan-out copies of all devices run concurrently (NVSwitch: full bandwidth from device 0 to every peer)
    for (int i = 1; i < ndev; i++) {
        CKC(cudaSetDevice(c->devs[i].id));
        CKC(cudaStreamSynchronize(c->devs[i].stream));
    }
#undef CKC
    *out = c;
    return RT_OK;
}

static int check_devices(const char* who, const int*& devices, int& ndev, const int* dflt)
{
    int avail = rt_device_count();
    if (avail <= 0) return fail(nullptr, RT_ERR_NO_DEVICE, std::string(who) + ": no CUDA device available (this library has no CPU fallback)");
    if (!devices || ndev <= 0) { devices = dflt; ndev = 1; }
    if (ndev > RT_MAX_DEVICES) return fail(nullptr, RT_ERR_INVALID, std::string(who) + ": too many devices");
    for (int i = 0; i < ndev; i++)
        if (devices[i] < 0 || devices[i] >= avail) return fail(nullptr, RT_ERR_INVALID, std::string(who) + ": device index out of range");
    return RT_OK;
}

int rt_create(const rt_scene_desc* desc, const int* devices, int ndev, rt_ctx** out)
{
    return rt::guarded("rt_create", [&]() -> int {
    if (!desc || !out) return fail(nullptr, RT_ERR_INVALID, "rt_create: null argument");
    *out = nullptr;
    const int dflt = 0;
    int rc = check_devices("rt_create", devices, ndev, &dflt);
    if (rc) return rc;
    rt::FlatScene flat;
    std::string err;
    rc = rt::flatten_scene(*desc, flat, err);
    if (rc) return fail(nullptr, rc, "rt_create: " + err);
    return create_common(flat, nullptr, devices, ndev, out);
    });
}

// Triangles -> render-ready context without the tree visiting the host: heuristic-6 BVH built on devices[0]
// (bvh_build_gpu.cu), flattened there (flatten_gpu.cu), fanned out to the other devices over NVLink.
int rt_create_gpu(rt_scene* s, int heuristic, const int* devices, int ndev, int download_tree, rt_ctx** out, rt_bvh_gpu_stats* stats)
{
    return rt::guarded("rt_create_gpu", [&]() -> int {
    if (stats) std::memset(stats, 0, sizeof *stats);
    if (!s || !out) return fail(nullptr, RT_ERR_INVALID, "rt_create_gpu: null argument");
    *out = nullptr;
    if ((heuristic & ~RT_BVH_REFBIN) != 6) return fail(nullptr, RT_ERR_INVALID, "rt_create_gpu: only heuristic 6 is built on the GPU");
    const int dflt = 0;
    int rc = check_devices("rt_create_gpu", devices, ndev, &dflt);
    if (rc) return rc;
    rt_scene_desc d;
    rt::GpuTree tree;
    const bool tiny = s->n_tris() <= 2; // the root is a leaf: flatten.cpp's synthetic root node
    const bool timing = std::getenv("RT_TIMING") != nullptr;
    auto t_mark = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        const auto now = std::chrono::steady_clock::now();
        if (timing) std::fprintf(stderr, "[rt_create_gpu] %-28s %.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_mark).count());
        t_mark = now;
    };
    if (!tiny) {
        rc = rt::gpu_build_bvh(*s, (heuristic & RT_BVH_REFBIN) ? 1 : 0, devices[0], stats, download_tree ? nullptr : &tree);
        if (rc) return rc;
        lap("BVH build (incl. upload)");
    } else if ((rc = rt_scene_build_bvh(s, heuristic))) return rc;
    if (tiny || download_tree || tree.fell_back) { // the tree is on the host (asked for, or degenerate input): host flatten
        if ((rc = rt_scene_view(s, &d))) return rc;
        return rt_create(&d, devices, ndev, out);
    }
    s->bvh.clear(); s->tri_idx.clear(); // the host scene has no tree in this mode
    if ((rc = rt_scene_view(s, &d))) return rc;
    rt::FlatScene small;
    rt::flatten_small(d, small);
    rt::DeviceFlat df;
    std::string err;
    lap("host scene view");
    rc = rt::flatten_gpu(tree, s->tri_mat.empty() ? nullptr : s->tri_mat.data(), s->n_mats(), df, err);
    lap("device-side flatten");
    tree.release();
    lap("release of the build's arrays");
    if (rc) {
        cudaFree(df.nodes); cudaFree(df.nodes4); cudaFree(df.nodes8); cudaFree(df.tris); cudaFree(df.shade); cudaFree(df.leaf_cnt);
        return fail(nullptr, rc, "rt_create_gpu: " + err);
    }
    rc = create_common(small, &df, devices, ndev, out);
    cudaFree(df.nodes); cudaFree(df.nodes4); cudaFree(df.nodes8); cudaFree(df.tris); cudaFree(df.shade); cudaFree(df.leaf_cnt); // (null once adopted)
    lap("context (adopts the arrays)");
    return rc;
    });
}

void rt_destroy(rt_ctx* c)
{
    if (!c) return;
    if (!c->devs.empty()) {
        cudaSetDevice(c->devs[0].id);
        cudaDeviceSynchronize();
        for (Slot& S : c->slots) {
            if (S.ipc_frame) cudaIpcCloseMemHandle(S.ipc_frame);
            cudaFree(S.bgra); cudaFree(S.rgb); cudaFree(S.tri_id); cudaFree(S.depth);
            if (S.copy_done) cudaEventDestroy(S.copy_done);
        }
        cudaFree(c->local_index); cudaFree(c->gather_buf);
    }
    for (Dev& D : c->devs) free_dev(D);
    delete c;
}

// Queue one frame on slot p->frame_slot: validation, frame storage, launches on every device, stats copy.  Returns
// without waiting for the device, except for the unfused RT_GATHER_PEER_COPY assembly, which blocks as before.
static int enqueue_frame(rt_ctx* c, const rt_render_params* p)
{
    if (!c || !p) return fail(c, RT_ERR_INVALID, "rt_render: null argument");
    if (p->width < 1 || p->height < 1 || p->width > 65535 || p->height > 32767) return fail(c, RT_ERR_INVALID, "rt_render: bad resolution");
    if (p->spp < 1) return fail(c, RT_ERR_INVALID, "rt_render: spp must be >= 1");
    if (p->bounces > RT_MAX_BOUNCES) return fail(c, RT_ERR_INVALID, "rt_render: bounces > 8");
    if (p->mode != RT_MODE_FAST && p->mode != RT_MODE_STRICT) return fail(c, RT_ERR_INVALID, "rt_render: bad mode");
    if (p->frame_slot < 0 || p->frame_slot >= RT_FRAME_SLOTS) return fail(c, RT_ERR_INVALID, "rt_render: bad frame_slot");
    const int part_count = p->part_count < 1 ? 1 : p->part_count;
    if (p->part_index < 0 || p->part_index >= part_count) return fail(c, RT_ERR_INVALID, "rt_render: bad partition");
    const int nd = (int)c->devs.size();
    const int w = p->width, h = p->height;
    const size_t npx = (size_t)w * h;
    const int slot = p->frame_slot;
    Slot& S = c->slots[slot];
    if (S.render_pending) return fail(c, RT_ERR_STATE, "rt_render_async: a render is still queued on this frame slot (call rt_frame_wait first)");
    int gather = p->gather;
    for (int d = 1; d < nd; d++)
        if (!c->devs[d].peer_to_0) gather = RT_GATHER_PEER_COPY; // no NVLink mapping: staged copies
    if (nd > 1 && gather == RT_GATHER_PEER_COPY && (p->aov_mask & (RT_AOV_RGB_F32 | RT_AOV_TRI_ID | RT_AOV_DEPTH)))
        for (int d = 1; d < nd; d++)
            if (!c->devs[d].peer_to_0) return fail(c, RT_ERR_INVALID, "rt_render: AOVs on several devices need peer access");
    if (nd > 1 && gather == RT_GATHER_PEER_COPY && (p->frame_flags & RT_FRAME_BOTTOM_UP))
        return fail(c, RT_ERR_INVALID, "rt_render: RT_FRAME_BOTTOM_UP is not available with RT_GATHER_PEER_COPY");
    if (S.ipc_frame && (S.ipc_w != w || S.ipc_h != h)) return fail(c, RT_ERR_STATE, "rt_render: imported frame has another size");
    if (nd > 1 && gather == RT_GATHER_PEER_COPY && S.ipc_frame)
        return fail(c, RT_ERR_INVALID, "rt_render: RT_GATHER_PEER_COPY cannot assemble into an imported (CUDA IPC) frame");
    if (nd > 1 && part_count > 1) return fail(c, RT_ERR_INVALID, "rt_render: use either several devices per context or part_count > 1");

    int rc = setup_tiles(c, w, h, p->part_index, part_count);
    if (rc) return rc;

    // frame storage on device 0 (+ local frames for PEER_COPY)
    Dev& D0 = c->devs[0];
    CK(c, cudaSetDevice(D0.id));
    {   // Growing a buffer frees the old one: nothing queued on this slot may still touch it.  The slot has been waited on
        // (render_pending is false), which covers every device's render kernel; a device->host copy may still be in flight.
        const bool grow = (!S.ipc_frame && S.bgra_px < npx) || ((p->aov_mask & RT_AOV_RGB_F32) && S.aov_px[0] < 3 * npx) ||
                          ((p->aov_mask & RT_AOV_TRI_ID) && S.aov_px[1] < npx) || ((p->aov_mask & RT_AOV_DEPTH) && S.aov_px[2] < npx);
        if (grow) {
            if (S.copy_pending) { CK(c, cudaEventSynchronize(S.copy_done)); S.copy_pending = false; }
            for (Dev& D : c->devs) { CK(c, cudaSetDevice(D.id)); CK(c, cudaStreamSynchronize(D.stream)); }
            CK(c, cudaSetDevice(D0.id));
        }
    }
    if (!S.ipc_frame && (rc = ensure(c, &S.bgra, &S.bgra_px, npx))) return rc; // an imported frame is the target: no local one
    if (p->aov_mask & RT_AOV_RGB_F32) { if ((rc = ensure(c, &S.rgb, &S.aov_px[0], 3 * npx))) return rc; }
    if (p->aov_mask & RT_AOV_TRI_ID) { if ((rc = ensure(c, &S.tri_id, &S.aov_px[1], npx))) return rc; }
    if (p->aov_mask & RT_AOV_DEPTH) { if ((rc = ensure(c, &S.depth, &S.aov_px[2], npx))) return rc; }
    if (nd > 1 && gather == RT_GATHER_PEER_COPY) {
        for (int d = 1; d < nd; d++) {
            Dev& D = c->devs[d];
            CK(c, cudaSetDevice(D.id));
            if ((rc = ensure(c, &D.bgra, &D.bgra_px, npx))) return rc;
            if ((rc = ensure(c, &D.packed, &D.packed_px, std::max<size_t>((size_t)D.n_tiles * RT_TILE_PIXELS, 1)))) return rc;
        }
    }

    rt::CameraBasis cb;
    rt::camera_basis(p->cam, w, h, cb);

    RtFrameArgs fa;
    std::memset(&fa, 0, sizeof fa);
    std::memcpy(fa.pos, cb.pos, 12); std::memcpy(fa.ul, cb.ul, 12);
    std::memcpy(fa.inc_x, cb.inc_x, 12); std::memcpy(fa.inc_y, cb.inc_y, 12);
    fa.width = w; fa.height = h; fa.spp = p->spp; fa.bounces = p->bounces; fa.seed = p->seed;
    fa.flip_y = (p->frame_flags & RT_FRAME_BOTTOM_UP) ? 1 : 0;
    fa.tiles_x = tiles_x_of(w);

    RtLaunchCfg cfg;
    cfg.block_threads = p->block_threads == 64 ? 64 : 128;
    // Defaults: the 4-wide tree at every size (profiles/r02_ab_large.jsonl); frames above 4 M pixel samples per GPU draw their work
    // from per-SM cursors over macro tiles (throughput-bound: L1 sharing matters), smaller ones from one global counter.  Which
    // kernel variant runs is decided from the render parameters ONLY, never from what earlier frames did: two frames with equal
    // parameters run the same traversal code, on every rank of a partitioned render (history only orders tiles and sizes the grid).
    const double px_per_part = (double)w * h * p->spp / (double)(part_count * (int)c->devs.size());
    static const double smq_px = [] { const char* e = std::getenv("RT_SMQUEUE_PIXELS"); return e ? std::atof(e) : 4.0e6; }();
    const bool small_frame = px_per_part <= smq_px;
    // fast build: which tree to walk.  The compressed 8-wide tree when the context has one and its depth fits the group
    // stack; RT_TRAVERSAL_* pins a variant (the strict build always walks the reference's own 2-wide order).
    const bool have8 = c->devs[0].nodes8 != nullptr && c->depth8 + 2 <= RT_STACK8_ENTRIES;
    const bool have4 = c->stack_need4 <= RT_STACK_ENTRIES_WIDE;
    int wide = 0;
    if (p->traversal == RT_TRAVERSAL_WIDE8) wide = have8 ? 2 : (have4 ? 1 : 0);
    else if (p->traversal == RT_TRAVERSAL_WIDE) wide = have4 ? 1 : 0;
    else if (p->traversal == RT_TRAVERSAL_DEFAULT) wide = have4 ? 1 : 0; // (until the 4-wide tree was rebuilt in round 2, frames above 4 M pixel samples per GPU walked the 2-wide tree)
    cfg.min_ctas = p->ctas_per_sm > 0 ? p->ctas_per_sm : (cfg.block_threads == 64 ? 12 : (wide ? 6 : 8));
    cfg.work_counters = (p->aov_mask & RT_AOV_WORK) != 0;
    cfg.speculative = p->traversal != RT_TRAVERSAL_PLAIN;
    cfg.wide = wide;
    // on the compressed tree: chunk culling and the cooperative drain kernel, both opt-in (measured neutral / slower on the
    // shipped scenes, profiles/r02_notes.md)
    const bool tree8 = p->mode == RT_MODE_FAST && wide == 2;
    fa.cull = (tree8 && p->cull > 0) ? 1 : 0;
    int drain_k = (tree8 && p->drain_k > 0 && 7 * c->depth8 + 1 <= RT_DRAIN_STACK) ? std::min(p->drain_k, 32) : 0;
    if (cfg.work_counters && c->want_trace) drain_k = 0; // the per-warp timeline describes the per-lane kernel alone
    // lanes that have finished their ray wait for phase 1 until fewer than this many lanes of the warp are still tracing (sweep of the
    // final kernels: profiles/r02_ab_refill_final.log; 20 was the optimum of the round-1 kernels and stays for the strict build)
    fa.refill_threshold = p->refill_threshold > 0 ? std::min(p->refill_threshold, 32) : (p->mode == RT_MODE_FAST ? (wide ? 14 : 16) : 20);

    S.width = w; S.height = h; S.aov_mask = p->aov_mask; S.spp = p->spp; S.flags = p->frame_flags;
    S.part_index = p->part_index; S.part_count = part_count;
    S.gather_ms = 0.f; S.launches = 0;
    uchar4* const target = S.ipc_frame ? S.ipc_frame : S.bgra;

    if (p->bounces <= 0) {
        // BOUNCES == 0: raytrace returns black before any traversal (cpu/src/raytracer.c:104-105)
        CK(c, cudaSetDevice(D0.id));
        if (S.copy_pending) CK(c, cudaStreamWaitEvent(D0.stream, S.copy_done, 0));
        for (int d = 0; d < nd; d++) { // other devices have nothing to do: their events only mark "complete"
            Dev& D = c->devs[d];
            CK(c, cudaSetDevice(D.id));
            unsigned long long* const ctrl = D.ctrl + RT_CTRL_WORDS * slot;
            CK(c, cudaMemsetAsync(ctrl, 0, 8 * RT_CTRL_WORDS, D.stream));
            CK(c, cudaEventRecord(D.ev0[slot], D.stream));
            if (d == 0) {
                fill_bgra_kernel<<<D0.sm_count * 4, 256, 0, D0.stream>>>(target, npx, make_uchar4(0, 0, 0, 255));
                CK(c, cudaGetLastError());
                if (p->aov_mask & RT_AOV_RGB_F32) CK(c, cudaMemsetAsync(S.rgb, 0, 12 * npx, D0.stream));
                if (p->aov_mask & RT_AOV_TRI_ID) CK(c, cudaMemsetAsync(S.tri_id, 0xff, 4 * npx, D0.stream));
                if (p->aov_mask & RT_AOV_DEPTH) CK(c, cudaMemsetAsync(S.depth, 0, 4 * npx, D0.stream));
            }
            CK(c, cudaEventRecord(D.ev1[slot], D.stream));
            CK(c, cudaStreamWaitEvent(D.aux, D.ev1[slot], 0));
            CK(c, cudaMemcpyAsync(D.ctrl_host + 8 * slot, ctrl, 64, cudaMemcpyDeviceToHost, D.aux));
            CK(c, cudaEventRecord(D.ev_done[slot], D.aux));
        }
        S.launches = 1;
        S.render_pending = true;
        return RT_OK;
    }

    // ---- launch on every device ----
    unsigned launches = 0;
    for (int d = 0; d < nd; d++) {
        Dev& D = c->devs[d];
        CK(c, cudaSetDevice(D.id));
        RtDeviceScene sc = c->scene_host_view;
        sc.nodes = D.nodes; sc.nodes4 = D.nodes4; sc.nodes8 = D.nodes8; sc.tris = D.tris; sc.shade = D.shade;
        sc.mats = D.mats; sc.lights = D.lights; sc.leaf_cnt = D.leaf_cnt;
        sc.n_inner = (unsigned)(D.bytes[0] / 64); sc.n_nodes4 = (unsigned)(D.bytes[1] / 128); sc.n_tris = (unsigned)(D.bytes[2] / 64);
        sc.n_nodes8 = (unsigned)(D.bytes[7] / 96);
        sc.err = D.ctrl + RT_CTRL_WORDS * slot + 7; // written by the checked build only (RT_DEBUG_BOUNDS)
        RtFrameArgs f = fa;
        f.tile_list = D.tile_list; f.n_tiles = D.n_tiles;
        unsigned long long* const ctrl = D.ctrl + RT_CTRL_WORDS * slot;
        f.stats = ctrl; f.tile_counter = reinterpret_cast<unsigned*>(ctrl + 4);
        f.sm_cursor = small_frame ? nullptr : ctrl + 8; // SM-local work queues pay off on throughput-bound frames only
        f.n_sms = (unsigned)std::min(D.sm_count, RT_MAX_SMS);
        const bool local = (d > 0 && gather == RT_GATHER_PEER_COPY);
        f.bgra = local ? D.bgra : target; // peer-mapped for d > 0
        f.rgb = (p->aov_mask & RT_AOV_RGB_F32) ? S.rgb : nullptr;
        f.tri_id = (p->aov_mask & RT_AOV_TRI_ID) ? S.tri_id : nullptr;
        f.depth = (p->aov_mask & RT_AOV_DEPTH) ? S.depth : nullptr;

        RtLaunchCfg cf = cfg;
        static const bool cost_all = [] { const char* e = std::getenv("RT_COST_2WIDE"); return !e || std::atoi(e) != 0; }();
        const bool track_cost = p->mode == RT_MODE_FAST && (cfg.wide != 0 || cost_all) && p->schedule >= 0 && p->bounces > 0 && D.n_tiles > 0;
        const int key[5] = {w, h, p->spp, p->part_index, part_count};
        if (track_cost && cfg.wide != 0 && cfg.block_threads == 128 && D.stat_total > 0 && std::memcmp(key, D.stat_key, sizeof key) == 0) {
            // A frame whose heaviest pixel takes much longer than an even share of the frame's steps is bound by that pixel's
            // dependent chain, and the chain runs faster with fewer warps competing for the SM's issue slots (and with the
            // registers of the roomier kernel instance); a frame with work for every lane wants all the warps.
            // r = heaviest pixel / (total steps / lane slots of the default occupancy); thresholds from the sweep in
            // profiles/r02_ab_occupancy_ratio2.jsonl (full frames and partitions of three scenes, final kernels).  The bytes of a
            // frame do not depend on the grid.
            static const bool adaptive = [] { const char* e = std::getenv("RT_ADAPTIVE_CTAS"); return !e || std::atoi(e) != 0; }();
            const double even = (double)D.stat_total / ((double)D.sm_count * 6 * cfg.block_threads); // (6 = the wide kernels' default CTAs/SM)
            const double r = (double)D.stat_max / even;
            const int want = r < 1.45 ? 6 : r < 3.2 ? 5 : r < 3.55 ? 4 : r < 6.0 ? 3 : 2;
            if (adaptive && p->ctas_per_sm <= 0 && want < cf.min_ctas) cf.min_ctas = want;
            if (adaptive && r >= 1.45 && p->refill_threshold <= 0) f.refill_threshold = 8; // chain-bound: phase 1 as rarely as possible
            // work distribution: per-SM cursors over macro tiles keep the warps of an SM on neighbouring tiles (L1 sharing: -3 % on
            // throughput-bound frames of any size) but hand the heaviest tiles out four at a time per SM, which a chain-bound frame
            // cannot afford (+13 %): profiles/r02_ab_smq_px.log.  Without statistics the frame size decides (above).
            if (adaptive) f.sm_cursor = r < 1.45 ? ctrl + 8 : nullptr;
            // ... and the kernel instance: with the path state parked in shared memory the 4-wide kernel fits 8 CTAs per SM, which a
            // throughput-bound frame can use (car_boxed 1080p / 4K -4.4 %, 8K -5.8 %; a share with r = 1.2 loses 3 %: hence r < 1); a
            // chain-bound frame wants few warps and all registers
            static const bool park_ok = [] { const char* e = std::getenv("RT_PARK"); return !e || std::atoi(e) != 0; }();
            if (adaptive && park_ok && r < 1.0 && p->ctas_per_sm <= 0 && !cfg.work_counters) { cf.park = true; cf.min_ctas = 8; }
        }
        int occ = 0, regs = 0;
        cudaError_t e = (p->mode == RT_MODE_STRICT) ? rt_occupancy_strict(cf, &occ, &regs) : rt_occupancy_fast(cf, &occ, &regs);
        CK(c, e);
        if (occ < 1) return fail(c, RT_ERR_CUDA, "rt_render: kernel does not fit on an SM");
        // persistent: one resident wave; ctas_per_sm may ask for FEWER resident CTAs than fit (fewer warps per SM
        // run each warp faster, which shortens the tail of long paths at equal throughput)
        const int ctas = (p->ctas_per_sm > 0 || cf.min_ctas != cfg.min_ctas) ? std::min(occ, cf.min_ctas) : occ;
        cf.grid = D.sm_count * ctas;
        const int warps_needed = (D.n_tiles * (RT_TILE_PIXELS / 32) + (cf.block_threads / 32) - 1) / (cf.block_threads / 32);
        if (warps_needed < cf.grid) cf.grid = std::max(warps_needed, 1);

        // heaviest tiles first: this frame records per-pixel costs and, when the previous frame had this shape, renders the
        // tiles in the order made from that frame's costs
        f.cost_out = nullptr;
        if (track_cost) {
            if (D.cost_px < npx || D.sorted_cap < (size_t)D.n_tiles) { // (re)allocate
                CK(c, cudaStreamSynchronize(D.stream));
                cudaFree(D.cost); cudaFree(D.tile_sorted); cudaFree(D.tile_cls);
                D.cost = nullptr; D.tile_sorted = nullptr; D.tile_cls = nullptr; D.cost_px = 0; D.sorted_cap = 0; D.cost_valid = false;
                CK(c, cudaMalloc((void**)&D.cost, npx * 2));
                CK(c, cudaMalloc((void**)&D.tile_sorted, (size_t)D.n_tiles * 4)); CK(c, cudaMalloc((void**)&D.tile_cls, (size_t)D.n_tiles));
                if (!D.cost_hdr) {
                    CK(c, cudaMalloc((void**)&D.cost_hdr, 2 * RT_COST_HDR * 4));
                    CK(c, cudaMalloc((void**)&D.cost_blk, (size_t)D.sm_count * 4 * RT_COST_CLASSES * 4));
                }
                D.cost_px = npx; D.sorted_cap = (size_t)D.n_tiles;
                std::memset(D.cost_key, 0, sizeof D.cost_key);
            }
            if (std::memcmp(key, D.cost_key, sizeof key) != 0 || D.tile_epoch_seen != D.tile_epoch) {
                // another shape, partition or tile list: the history is void (pixels this rank does not render are never read)
                D.cost_valid = false; D.tile_epoch_seen = D.tile_epoch;
                CK(c, cudaMemsetAsync(D.cost_hdr, 0, 2 * RT_COST_HDR * 4, D.stream));
            }
            f.cost_out = D.cost;
            if (D.cost_valid) f.tile_list = D.tile_sorted;
        }
        f.drain_k = 0; f.drain_queue = nullptr; f.drain_cap = 0;
        f.drain_count = reinterpret_cast<unsigned*>(ctrl + 5); f.drain_next = reinterpret_cast<unsigned*>(ctrl + 6);
        if (drain_k > 0) {
            const size_t cap = (size_t)cf.grid * (cf.block_threads / 32) * (size_t)drain_k;
            if ((rc = ensure(c, &D.drain_queue[slot], &D.drain_cap[slot], cap))) return rc;
            f.drain_k = drain_k; f.drain_queue = D.drain_queue[slot]; f.drain_cap = (unsigned)cap;
        }
        f.warp_trace = nullptr;
        if (cf.work_counters && c->want_trace) {
            const size_t nw = (size_t)cf.grid * (cf.block_threads / 32);
            if ((rc = ensure(c, &D.warp_trace, &D.warp_trace_cap, nw * 8))) return rc;
            CK(c, cudaMemsetAsync(D.warp_trace, 0, nw * 64, D.stream));
            f.warp_trace = D.warp_trace;
            D.warp_trace_n = (int)nw;
        }
        // a copy of this slot's previous frame that is still in flight must finish before its pixels are overwritten
        if (S.copy_pending) CK(c, cudaStreamWaitEvent(D.stream, S.copy_done, 0));
        CK(c, cudaMemsetAsync(ctrl, 0, 8 * RT_CTRL_WORDS, D.stream));
        CK(c, cudaEventRecord(D.ev0[slot], D.stream));
        e = (p->mode == RT_MODE_STRICT) ? rt_launch_strict(sc, f, cf, D.stream) : rt_launch_fast(sc, f, cf, D.stream);
        CK(c, e);
        launches++;
        if (f.drain_k > 0) { // the paths the render kernel's warps handed off at the end of the chunk queue
            CK(c, rt_launch_drain(sc, f, cf.work_counters, D.sm_count, D.stream));
            launches++;
        }
        if (track_cost) { // order the tiles for the next frame of this shape (inside this frame's timed window)
            const int nb = std::min(D.sm_count, (D.n_tiles + RT_COST_BLOCK / 32 - 1) / (RT_COST_BLOCK / 32));
            unsigned* const hdr = D.cost_hdr + RT_COST_HDR * D.cost_cur;
            // (frame statistics for the occupancy choice below go into two words of the control block that only the opt-in drain
            // kernel uses otherwise; they reach the host with the ray counters)
            unsigned long long* const stat = f.drain_k > 0 ? nullptr : ctrl + 5;
            tile_class_kernel<<<nb, RT_COST_BLOCK, 0, D.stream>>>(D.cost, w, h, fa.tiles_x, D.tile_list, D.n_tiles, D.tile_cls, hdr,
                                                                  D.cost_hdr + RT_COST_HDR * (1 - D.cost_cur), D.cost_blk, stat);
            D.slot_stat[slot] = stat != nullptr;
            std::memcpy(D.slot_key[slot], key, sizeof key);
            tile_place_kernel<<<nb, RT_COST_BLOCK, 0, D.stream>>>(D.tile_list, D.n_tiles, D.tile_cls, hdr, D.cost_blk, D.tile_sorted);
            CK(c, cudaGetLastError());
            launches += 2;
            D.cost_cur = 1 - D.cost_cur; D.cost_valid = true;
            std::memcpy(D.cost_key, key, sizeof key);
        }
        CK(c, cudaEventRecord(D.ev1[slot], D.stream));
        // statistics -> pinned host memory on the second stream (this slot's control block is not touched again before
        // the slot has been waited on, so the next frame's kernel does not depend on this copy)
        CK(c, cudaStreamWaitEvent(D.aux, D.ev1[slot], 0));
        CK(c, cudaMemcpyAsync(D.ctrl_host + 8 * slot, ctrl, 64, cudaMemcpyDeviceToHost, D.aux));
        CK(c, cudaEventRecord(D.ev_done[slot], D.aux));
    }
    S.render_pending = true;

    // ---- unfused gather: pack on each device, copy to device 0, unpack (blocking) ----
    float gather_ms = 0.f;
    if (nd > 1 && gather == RT_GATHER_PEER_COPY) {
        const int parts = part_count * nd;
        if ((rc = setup_local_index(c, w, h, parts))) return rc;
        size_t stride = 0;
        for (int d = 0; d < nd; d++) stride = std::max(stride, (size_t)c->devs[d].n_tiles * RT_TILE_PIXELS);
        CK(c, cudaSetDevice(D0.id));
        if ((rc = ensure(c, &c->gather_buf, &c->gather_px, stride * parts))) return rc;
        if ((rc = ensure(c, &D0.packed, &D0.packed_px, std::max<size_t>((size_t)D0.n_tiles * RT_TILE_PIXELS, 1)))) return rc;
        for (int d = 1; d < nd; d++) {
            Dev& D = c->devs[d];
            CK(c, cudaSetDevice(D.id));
            if (D.n_tiles) {
                pack_tiles_kernel<<<D.n_tiles, RT_TILE_PIXELS, 0, D.stream>>>(D.bgra, D.packed, D.tile_list, D.n_tiles, fa.tiles_x, w, h);
                CK(c, cudaGetLastError());
                launches++;
                CK(c, cudaMemcpyPeerAsync(c->gather_buf + stride * (size_t)(p->part_index * nd + d), D0.id, D.packed, D.id,
                                          (size_t)D.n_tiles * RT_TILE_PIXELS * 4, D.stream));
            }
            CK(c, cudaEventRecord(D.ev2, D.stream));
        }
        CK(c, cudaSetDevice(D0.id));
        for (int d = 1; d < nd; d++) CK(c, cudaStreamWaitEvent(D0.stream, c->devs[d].ev2, 0));
        CK(c, cudaEventRecord(D0.ev2, D0.stream));
        // device 0's own tiles are already in place; scatter the others (device 0's slot is skipped by
        // packing its own tiles too, which keeps the unpack kernel branch-free)
        if (D0.n_tiles) {
            pack_tiles_kernel<<<D0.n_tiles, RT_TILE_PIXELS, 0, D0.stream>>>(S.bgra, D0.packed, D0.tile_list, D0.n_tiles, fa.tiles_x, w, h);
            CK(c, cudaGetLastError());
            launches++;
            CK(c, cudaMemcpyAsync(c->gather_buf + stride * (size_t)(p->part_index * nd), D0.packed, (size_t)D0.n_tiles * RT_TILE_PIXELS * 4,
                                  cudaMemcpyDeviceToDevice, D0.stream));
        }
        dim3 blk(32, 8), grd((w + 31) / 32, (h + 7) / 8);
        unpack_tiles_kernel<<<grd, blk, 0, D0.stream>>>(S.bgra, c->gather_buf, stride, c->local_index, parts, fa.tiles_x, w, h);
        CK(c, cudaGetLastError());
        launches++;
        cudaEvent_t done;
        CK(c, cudaEventCreate(&done));
        CK(c, cudaEventRecord(done, D0.stream));
        CK(c, cudaEventSynchronize(done));
        // gather time = slowest (pack + NVLink copy) of the other devices + (pack + unpack) on device 0
        CK(c, cudaEventElapsedTime(&gather_ms, D0.ev2, done));
        CK(c, cudaEventDestroy(done));
        float worst = 0.f;
        for (int d = 1; d < nd; d++) {
            float t_d = 0.f;
            CK(c, cudaSetDevice(c->devs[d].id));
            CK(c, cudaEventSynchronize(c->devs[d].ev2));
            CK(c, cudaEventElapsedTime(&t_d, c->devs[d].ev1[slot], c->devs[d].ev2));
            worst = std::max(worst, t_d);
        }
        gather_ms += worst;
    }
    S.gather_ms = gather_ms;
    S.launches = launches;
    return RT_OK;
}

// Wait for everything queued on the slot (render, then a device->host copy if one was queued); timing of its render.
static int finish_frame(rt_ctx* c, int slot, rt_timing* tm)
{
    if (!c) return fail(c, RT_ERR_INVALID, "rt_frame_wait: null context");
    if (slot < 0 || slot >= RT_FRAME_SLOTS) return fail(c, RT_ERR_INVALID, "rt_frame_wait: bad frame slot");
    Slot& S = c->slots[slot];
    if (!S.render_pending && !S.rendered) return fail(c, RT_ERR_STATE, "rt_frame_wait: nothing was rendered on this frame slot");
    const int nd = (int)c->devs.size();
    if (S.render_pending) {
        rt_timing t;
        std::memset(&t, 0, sizeof t);
        float kmax = 0.f;
        for (int d = 0; d < nd; d++) {
            Dev& D = c->devs[d];
            CK(c, cudaSetDevice(D.id));
            CK(c, cudaEventSynchronize(D.ev_done[slot]));
            CK(c, cudaEventElapsedTime(&t.kernel_ms[d], D.ev0[slot], D.ev1[slot]));
            kmax = std::max(kmax, t.kernel_ms[d]);
            const unsigned long long* st = D.ctrl_host + 8 * slot;
            t.rays_closest += st[0];
            t.rays_shadow += st[1];
            t.inner_visits += st[2];
            t.tri_tests += st[3];
            if (D.slot_stat[slot]) { D.stat_max = st[5]; D.stat_total = st[6]; std::memcpy(D.stat_key, D.slot_key[slot], sizeof D.stat_key); D.slot_stat[slot] = false; }
            if (st[7]) { // checked build: an index left its array (render_kernel.cuh: RT_BCHECK codes)
                S.render_pending = false;
                return fail(c, RT_ERR_STATE, "RT_DEBUG_BOUNDS: check " + std::to_string(st[7]) + " failed on device " + std::to_string(D.id));
            }
        }
        t.gather_ms = S.gather_ms;
        t.total_ms = kmax + S.gather_ms;
        t.launches = S.launches;
        t.n_devices = (uint32_t)nd;
        S.timing = t;
        S.render_pending = false;
        S.rendered = true;
        c->width = S.width; c->height = S.height; c->aov_mask = S.aov_mask; c->rendered = true;
        c->part_index = S.part_index; c->part_count = S.part_count;
        c->last_slot = slot;
    }
    if (S.copy_pending) {
        CK(c, cudaSetDevice(c->devs[0].id));
        CK(c, cudaEventSynchronize(S.copy_done));
        S.copy_pending = false;
    }
    if (tm) *tm = S.timing;
    return RT_OK;
}

int rt_render_async(rt_ctx* c, const rt_render_params* p) { return enqueue_frame(c, p); }

int rt_frame_wait(rt_ctx* c, int slot, rt_timing* tm) { return finish_frame(c, slot, tm); }

int rt_render(rt_ctx* c, const rt_render_params* p, rt_timing* tm)
{
    int rc = enqueue_frame(c, p);
    if (rc) return rc;
    return finish_frame(c, p->frame_slot, tm);
}

int rt_download_async(rt_ctx* c, int slot, uint8_t* host_bgra)
{
    if (!c || !host_bgra) return fail(c, RT_ERR_INVALID, "rt_download_async: null argument");
    if (slot < 0 || slot >= RT_FRAME_SLOTS) return fail(c, RT_ERR_INVALID, "rt_download_async: bad frame slot");
    Slot& S = c->slots[slot];
    if (!S.render_pending && !S.rendered) return fail(c, RT_ERR_STATE, "rt_download_async: nothing was rendered on this frame slot");
    Dev& D0 = c->devs[0];
    CK(c, cudaSetDevice(D0.id));
    // the copy is ordered after the slot's render kernel on EVERY device (peer stores land in this frame)
    for (Dev& D : c->devs) CK(c, cudaStreamWaitEvent(c->copy_stream, D.ev1[slot], 0));
    CK(c, cudaMemcpyAsync(host_bgra, S.ipc_frame ? S.ipc_frame : S.bgra, (size_t)S.width * S.height * 4, cudaMemcpyDeviceToHost, c->copy_stream));
    CK(c, cudaEventRecord(S.copy_done, c->copy_stream));
    S.copy_pending = true;
    return RT_OK;
}

int rt_host_alloc(size_t bytes, void** out)
{
    if (!out || !bytes) return fail(nullptr, RT_ERR_INVALID, "rt_host_alloc: bad argument");
    *out = nullptr;
    if (rt_device_count() <= 0) return fail(nullptr, RT_ERR_NO_DEVICE, "rt_host_alloc: no CUDA device");
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, RT_ERR_NOMEM, std::string("rt_host_alloc: ") + cudaGetErrorString(e)); }
    return RT_OK;
}

void rt_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

int rt_download(rt_ctx* c, uint8_t* bgra, float* rgb, int32_t* tri_id, float* depth_t)
{
    if (!c) return fail(c, RT_ERR_INVALID, "rt_download: null context");
    if (!c->rendered) return fail(c, RT_ERR_STATE, "rt_download: nothing rendered yet");
    Dev& D0 = c->devs[0];
    Slot& S = c->slots[c->last_slot];
    const size_t npx = (size_t)c->width * c->height;
    CK(c, cudaSetDevice(D0.id));
    if (rgb && !(c->aov_mask & RT_AOV_RGB_F32)) return fail(c, RT_ERR_STATE, "rt_download: RT_AOV_RGB_F32 was not rendered");
    if (tri_id && !(c->aov_mask & RT_AOV_TRI_ID)) return fail(c, RT_ERR_STATE, "rt_download: RT_AOV_TRI_ID was not rendered");
    if (depth_t && !(c->aov_mask & RT_AOV_DEPTH)) return fail(c, RT_ERR_STATE, "rt_download: RT_AOV_DEPTH was not rendered");
    if (bgra) CK(c, cudaMemcpyAsync(bgra, S.ipc_frame ? S.ipc_frame : S.bgra, npx * 4, cudaMemcpyDeviceToHost, D0.stream));
    if (rgb) CK(c, cudaMemcpyAsync(rgb, S.rgb, npx * 12, cudaMemcpyDeviceToHost, D0.stream));
    if (tri_id) CK(c, cudaMemcpyAsync(tri_id, S.tri_id, npx * 4, cudaMemcpyDeviceToHost, D0.stream));
    if (depth_t) CK(c, cudaMemcpyAsync(depth_t, S.depth, npx * 4, cudaMemcpyDeviceToHost, D0.stream));
    CK(c, cudaStreamSynchronize(D0.stream));
    return RT_OK;
}

int rt_frame_device_ptr(rt_ctx* c, void** dev_ptr, size_t* bytes)
{
    if (!c || !dev_ptr) return fail(c, RT_ERR_INVALID, "rt_frame_device_ptr: null argument");
    if (!c->rendered) return fail(c, RT_ERR_STATE, "rt_frame_device_ptr: nothing rendered yet");
    { Slot& S = c->slots[c->last_slot]; *dev_ptr = S.ipc_frame ? S.ipc_frame : S.bgra; }
    if (bytes) *bytes = (size_t)c->width * c->height * 4;
    return RT_OK;
}

int rt_packed_tiles(rt_ctx* c, void** dev_ptr, size_t* bytes)
{
    if (!c || !dev_ptr) return fail(c, RT_ERR_INVALID, "rt_packed_tiles: null argument");
    if (!c->rendered) return fail(c, RT_ERR_STATE, "rt_packed_tiles: nothing rendered yet");
    if (c->devs.size() != 1) return fail(c, RT_ERR_STATE, "rt_packed_tiles: one device per context expected");
    Dev& D = c->devs[0];
    Slot& S = c->slots[c->last_slot];
    if (S.flags & RT_FRAME_BOTTOM_UP) return fail(c, RT_ERR_STATE, "rt_packed_tiles: the last frame was rendered bottom-up");
    CK(c, cudaSetDevice(D.id));
    int rc = ensure(c, &D.packed, &D.packed_px, std::max<size_t>((size_t)D.n_tiles * RT_TILE_PIXELS, 1));
    if (rc) return rc;
    if (D.n_tiles) {
        pack_tiles_kernel<<<D.n_tiles, RT_TILE_PIXELS, 0, D.stream>>>(S.ipc_frame ? S.ipc_frame : S.bgra, D.packed, D.tile_list, D.n_tiles, tiles_x_of(c->width), c->width, c->height);
        CK(c, cudaGetLastError());
    }
    CK(c, cudaStreamSynchronize(D.stream));
    *dev_ptr = D.packed;
    if (bytes) *bytes = (size_t)D.n_tiles * RT_TILE_PIXELS * 4;
    return RT_OK;
}

int rt_unpack_tiles(rt_ctx* c, const void* dev_gathered, size_t stride_bytes, int part_count)
{
    if (!c || !dev_gathered || part_count < 1 || stride_bytes % 4) return fail(c, RT_ERR_INVALID, "rt_unpack_tiles: bad argument");
    if (!c->rendered) return fail(c, RT_ERR_STATE, "rt_unpack_tiles: nothing rendered yet");
    Dev& D0 = c->devs[0];
    const int w = c->width, h = c->height;
    if (!c->slots[c->last_slot].bgra) return fail(c, RT_ERR_STATE, "rt_unpack_tiles: this context renders into an imported frame");
    int rc = setup_local_index(c, w, h, part_count);
    if (rc) return rc;
    CK(c, cudaSetDevice(D0.id));
    dim3 blk(32, 8), grd((w + 31) / 32, (h + 7) / 8);
    unpack_tiles_kernel<<<grd, blk, 0, D0.stream>>>(c->slots[c->last_slot].bgra, static_cast<const uchar4*>(dev_gathered), stride_bytes / 4, c->local_index,
                                                    part_count, tiles_x_of(w), w, h);
    CK(c, cudaGetLastError());
    CK(c, cudaStreamSynchronize(D0.stream));
    return RT_OK;
}

// Diagnostics: per-warp timeline of the last RT_AOV_WORK render on the context's first device.  out holds
// 8 x u64 per warp: clock64 at start / when the tile queue ran dry / at exit, chunks fetched, traversal
// iterations, lanes served by inner steps, lanes served by triangle steps, SM id.  enable != 0 arms it.
int rt_debug_warp_trace(rt_ctx* c, int enable, unsigned long long* out, int max_warps)
{
    if (!c) return RT_ERR_INVALID;
    c->want_trace = enable != 0;
    if (!out) return 0;
    Dev& D = c->devs[0];
    const int n = std::min(max_warps, D.warp_trace_n);
    if (n > 0) {
        cudaSetDevice(D.id);
        if (cudaMemcpy(out, D.warp_trace, (size_t)n * 64, cudaMemcpyDeviceToHost) != cudaSuccess) return RT_ERR_CUDA;
    }
    return n;
}

// Diagnostics: replace device 0's tile order for the current frame shape (the tile list must already exist, i.e. a frame
// of that shape was rendered) — experiments with cost-ordered scheduling.
int rt_debug_set_tile_order(rt_ctx* c, const unsigned* tiles, int n)
{
    if (!c || !tiles) return RT_ERR_INVALID;
    Dev& D = c->devs[0];
    if (n != D.n_tiles) return fail(c, RT_ERR_INVALID, "rt_debug_set_tile_order: tile count differs from the current tile list");
    CK(c, cudaSetDevice(D.id));
    CK(c, cudaMemcpy(D.tile_list, tiles, (size_t)n * 4, cudaMemcpyHostToDevice));
    D.tile_epoch++;
    return RT_OK;
}

// Diagnostics: the per-pixel traversal-cost map of the last fast frame on the context's first device (W*H u16 steps,
// saturating; pixels another rank renders are undefined), and hdr2 = {tiles ordered ahead of the cheap class, tiles in all}.
int rt_debug_cost_map(rt_ctx* c, unsigned short* out, size_t n_pixels, unsigned* hdr2)
{
    if (!c || !out) return fail(c, RT_ERR_INVALID, "rt_debug_cost_map: null argument");
    Dev& D = c->devs[0];
    if (!D.cost_valid || n_pixels > D.cost_px) return fail(c, RT_ERR_STATE, "rt_debug_cost_map: no cost map of that size");
    CK(c, cudaSetDevice(D.id));
    CK(c, cudaStreamSynchronize(D.stream));
    CK(c, cudaMemcpy(out, D.cost, n_pixels * 2, cudaMemcpyDeviceToHost));
    if (hdr2) {
        unsigned h[RT_COST_CLASSES];
        CK(c, cudaMemcpy(h, D.cost_hdr + RT_COST_HDR * (1 - D.cost_cur), sizeof h, cudaMemcpyDeviceToHost)); // the last frame's header
        hdr2[0] = 0; hdr2[1] = 0;
        for (int k = 0; k < RT_COST_CLASSES; k++) { hdr2[1] += h[k]; if (k) hdr2[0] += h[k]; }
    }
    return RT_OK;
}

int rt_debug_tile_order(rt_ctx* c, int which, unsigned* out, int cap, int* n_out)
{
    if (!c || !out || !n_out || cap < 0 || which < 0 || which > 1) return fail(c, RT_ERR_INVALID, "rt_debug_tile_order: bad argument");
    Dev& D = c->devs[0];
    if (which == 1 && (!D.cost_valid || D.tile_epoch_seen != D.tile_epoch)) return fail(c, RT_ERR_STATE, "rt_debug_tile_order: no cost history");
    CK(c, cudaSetDevice(D.id));
    CK(c, cudaStreamSynchronize(D.stream));
    *n_out = D.n_tiles;
    const int n = std::min(cap, D.n_tiles);
    if (n > 0) CK(c, cudaMemcpy(out, which ? D.tile_sorted : D.tile_list, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return RT_OK;
}

int rt_debug_device_array(rt_ctx* c, int which, void* out, size_t cap_bytes, size_t* bytes_out)
{
    if (!c || !bytes_out || which < 0 || which > 7) return fail(c, RT_ERR_INVALID, "rt_debug_device_array: bad argument");
    Dev& D = c->devs[0];
    const void* src[8] = {D.nodes, D.nodes4, D.tris, D.shade, D.leaf_cnt, D.mats, D.lights, D.nodes8};
    *bytes_out = D.bytes[which];
    if (!out) return RT_OK;
    if (cap_bytes < D.bytes[which]) return fail(c, RT_ERR_INVALID, "rt_debug_device_array: buffer too small");
    CK(c, cudaSetDevice(D.id));
    if (D.bytes[which]) CK(c, cudaMemcpy(out, src[which], D.bytes[which], cudaMemcpyDeviceToHost));
    return RT_OK;
}

// Random 64-byte-record gather bandwidth of one device for a working set of ws_bytes (rounded down to a power of two):
// GB/s of records delivered to the lanes.  ws <= L1 size measures L1, a few MB L2 (the shipped scenes), GBs HBM.
int rt_debug_gather_bandwidth(int device, size_t ws_bytes, float* gbs_out)
{
    if (!gbs_out || ws_bytes < 4096) return fail(nullptr, RT_ERR_INVALID, "rt_debug_gather_bandwidth: bad argument");
    if (rt_device_count() <= device || device < 0) return fail(nullptr, RT_ERR_NO_DEVICE, "rt_debug_gather_bandwidth: no such device");
    size_t n_rec = 1;
    while (n_rec * 2 * 64 <= ws_bytes) n_rec *= 2;
    if (n_rec > (1ull << 31)) n_rec = 1ull << 31;
    float4 *data = nullptr, *sink = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    float best = 0.f;
    auto run = [&]() -> cudaError_t {
        cudaError_t e;
        if ((e = cudaSetDevice(device)) != cudaSuccess) return e;
        if ((e = cudaMalloc((void**)&data, n_rec * 64)) != cudaSuccess) return e;
        if ((e = cudaMalloc((void**)&sink, 64)) != cudaSuccess) return e;
        if ((e = cudaMemset(data, 0, n_rec * 64)) != cudaSuccess) return e;
        cudaDeviceProp prop;
        if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return e;
        const int grid = prop.multiProcessorCount * 8, block = 256, iters = 512;
        if ((e = cudaEventCreate(&e0)) != cudaSuccess || (e = cudaEventCreate(&e1)) != cudaSuccess) return e;
        for (int rep = 0; rep < 3; rep++) { // rep 0 warms the caches / clocks
            if ((e = cudaEventRecord(e0)) != cudaSuccess) return e;
            gather64_kernel<<<grid, block>>>(data, (unsigned)(n_rec - 1), iters, 12345u + rep, sink);
            if ((e = cudaEventRecord(e1)) != cudaSuccess || (e = cudaEventSynchronize(e1)) != cudaSuccess) return e;
            float ms = 0.f;
            if ((e = cudaEventElapsedTime(&ms, e0, e1)) != cudaSuccess) return e;
            const float gbs = (float)((double)grid * block * iters * 64.0 / (ms * 1e-3) / 1e9);
            if (rep > 0 && gbs > best) best = gbs;
        }
        return cudaSuccess;
    };
    const cudaError_t e = run();
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(data); cudaFree(sink);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, RT_ERR_CUDA, std::string("rt_debug_gather_bandwidth: ") + cudaGetErrorString(e)); }
    *gbs_out = best;
    return RT_OK;
}

// Host -> device copy rate of `bytes` from pageable memory: mode 0 = plain cudaMemcpy (what the reference does,
// gpu/src/gpu.cu:143-175), 1 = the library's staged copy through the pinned ring (staged_copy.h), 2 = cudaMemcpy from
// page-locked memory (the ceiling).  GB/s of the best of three repetitions.
int rt_debug_copy_bandwidth(int device, size_t bytes, int mode, float* gbs_out)
{
    if (!gbs_out || bytes < 4096 || mode < 0 || mode > 2) return fail(nullptr, RT_ERR_INVALID, "rt_debug_copy_bandwidth: bad argument");
    if (rt_device_count() <= device || device < 0) return fail(nullptr, RT_ERR_NO_DEVICE, "rt_debug_copy_bandwidth: no such device");
    void* dst = nullptr;
    void* pinned = nullptr;
    std::vector<char> src;
    float best = 0.f;
    auto run = [&]() -> cudaError_t {
        cudaError_t e;
        if ((e = cudaSetDevice(device)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&dst, bytes)) != cudaSuccess) return e;
        if (mode == 2) { if ((e = cudaHostAlloc(&pinned, bytes, cudaHostAllocDefault)) != cudaSuccess) return e; std::memset(pinned, 1, bytes); }
        else src.assign(bytes, 1);
        for (int rep = 0; rep < 3; rep++) {
            const auto t0 = std::chrono::steady_clock::now();
            if (mode == 0) e = cudaMemcpy(dst, src.data(), bytes, cudaMemcpyHostToDevice);
            else if (mode == 1) e = rt::staged_h2d(dst, src.data(), bytes, 0);
            else e = cudaMemcpy(dst, pinned, bytes, cudaMemcpyHostToDevice);
            if (e != cudaSuccess) return e;
            if ((e = cudaDeviceSynchronize()) != cudaSuccess) return e;
            const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            best = std::max(best, (float)(bytes / s / 1e9));
        }
        return cudaSuccess;
    };
    const cudaError_t e = run();
    cudaFree(dst);
    if (pinned) cudaFreeHost(pinned);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, RT_ERR_CUDA, std::string("rt_debug_copy_bandwidth: ") + cudaGetErrorString(e)); }
    *gbs_out = best;
    return RT_OK;
}

int rt_frame_ipc_export_slot(rt_ctx* c, int slot, int width, int height, void* handle64)
{
    if (!c || !handle64 || width < 1 || height < 1 || slot < 0 || slot >= RT_FRAME_SLOTS) return fail(c, RT_ERR_INVALID, "rt_frame_ipc_export: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    Dev& D0 = c->devs[0];
    Slot& S = c->slots[slot];
    CK(c, cudaSetDevice(D0.id));
    int rc = ensure(c, &S.bgra, &S.bgra_px, (size_t)width * height);
    if (rc) return rc;
    cudaIpcMemHandle_t hnd;
    CK(c, cudaIpcGetMemHandle(&hnd, S.bgra));
    std::memcpy(handle64, &hnd, 64);
    return RT_OK;
}

int rt_frame_ipc_import_slot(rt_ctx* c, int slot, const void* handle64, int width, int height)
{
    if (!c || !handle64 || width < 1 || height < 1 || slot < 0 || slot >= RT_FRAME_SLOTS) return fail(c, RT_ERR_INVALID, "rt_frame_ipc_import: bad argument");
    Dev& D0 = c->devs[0];
    Slot& S = c->slots[slot];
    CK(c, cudaSetDevice(D0.id));
    if (S.ipc_frame) { CK(c, cudaIpcCloseMemHandle(S.ipc_frame)); S.ipc_frame = nullptr; }
    cudaIpcMemHandle_t hnd;
    std::memcpy(&hnd, handle64, 64);
    void* p = nullptr;
    CK(c, cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess));
    S.ipc_frame = static_cast<uchar4*>(p);
    S.ipc_w = width; S.ipc_h = height;
    return RT_OK;
}

int rt_frame_ipc_export(rt_ctx* c, int width, int height, void* handle64) { return rt_frame_ipc_export_slot(c, 0, width, height, handle64); }
int rt_frame_ipc_import(rt_ctx* c, const void* handle64, int width, int height) { return rt_frame_ipc_import_slot(c, 0, handle64, width, height); }

} // extern "C"

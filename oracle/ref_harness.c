/*
 * ref_harness.c — TEST INFRASTRUCTURE.  Driver that links the UNMODIFIED reference
 * CPU renderer sources (compiled where they lie under /root/reference/cpu/src by
 * oracle/Makefile; nothing is copied into this repository) and exposes the
 * reference's per-pixel results to the parity tests and to bench.py's CPU arm.
 *
 * What is reference code and what is not:
 *   - bvh.c raytracer.c triangle.c light.c cam.c vec.c bmp_writer.c are compiled as
 *     they are.  Every intersection, traversal and shading result this program
 *     prints is computed by those files.
 *   - main.c cannot be linked (it owns main() and a static WIDTH*HEIGHT frame), so
 *     the ~35 lines of ray generation and row scheduling it contains are re-stated
 *     here with run-time width/height/camera:  thread_main() below follows
 *     cpu/src/main.c:241-264 (thread_render) and px_render() follows
 *     cpu/src/main.c:228-239 (render_pixel).  The process globals that main.c
 *     defines (cpu/src/main.c:27-37) are defined here under the same names.
 *   - options.h is shadowed through -I order by a generated header (oracle/Makefile)
 *     so that BVH_HEURISTIC can be 6 (the tree the GPU kernels use) or 3 (the
 *     reference CPU default).
 *
 * Outputs (raw little-endian, row 0 = top of the image, x fastest):
 *   PREFIX.id.i32   first-hit triangle index (OBJ face order), -1 = miss
 *   PREFIX.t.f32    first-hit ray parameter t (FLT_MAX on a miss), in units of the
 *                   un-normalised primary direction
 *   PREFIX.rgb.f32  clamped float colour, 3 floats / pixel
 *   PREFIX.bgra.u8  the 4 bytes/pixel the reference BMP writer would emit
 *   PREFIX.bmp      the reference's own bmp_write_file() output (optional)
 * and one JSON line on stdout with per-frame wall times (CLOCK_MONOTONIC around
 * thread create/join, as cpu/src/main.c:171-185 does).
 */
#define _GNU_SOURCE

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "bmp_writer.h"
#include "bvh.h"
#include "cam.h"
#include "light.h"
#include "raytracer.h"
#include "triangle.h"
#include "vec.h"

#include "rt_sampling.h"

/* ---- globals main.c would define (cpu/src/main.c:27-37) ---- */
cam_t cam;
size_t triangles_len;
triangle_t* triangles;
size_t lights_len;
light_t* lights;
vec_t amb_light = {.r = 0.5, .g = 0.5, .b = 0.5};

/* reference BVH arrays (non-static globals of cpu/src/bvh.c:16-18) */
extern bvh_t* bvh;
extern int* tri_idx;
extern int bvh_len;

/* ---- run-time frame state ---- */
static int g_w, g_h, g_spp;
static uint32_t g_seed;
static vec_t* g_rgb;     /* clamped colour */
static int32_t* g_id;
static float* g_t;
static int g_want_aov;
static atomic_int g_row_counter;

/* cpu/src/main.c:228-239 with a sub-pixel offset; offset (0,0) is the reference ray */
static vec_t px_dir(const vec_t* ul, const vec_t* inc_x, const vec_t* inc_y, float fx, float fy)
{
    vec_t dir = vec_sub(ul, &cam.pos);
    vec_t pos_x = vec_mul(inc_x, fx);
    vec_t pos_y = vec_mul(inc_y, fy);
    dir = vec_add(&dir, &pos_x);
    dir = vec_add(&dir, &pos_y);
    return dir;
}

static void px_render(const vec_t* ul, const vec_t* inc_x, const vec_t* inc_y, int x, int y)
{
    int idx = y * g_w + x;
    vec_t sum = {0, 0, 0};
    for (int s = 0; s < g_spp; s++) {
        float jx, jy;
        rt_sample_offset((uint32_t)x, (uint32_t)y, (uint32_t)s, g_seed, &jx, &jy);
        vec_t dir = px_dir(ul, inc_x, inc_y, (float)x + jx, (float)y + jy);
        vec_t col = raytrace(cam.pos, dir, 0);
        if (g_spp == 1) { sum = col; break; }
        sum = vec_add(&sum, &col);
    }
    if (g_spp > 1) sum = vec_div(&sum, (float)g_spp);
    const vec_t vec_0 = {0, 0, 0};
    const vec_t vec_1 = {1, 1, 1};
    vec_constrain(&sum, &vec_0, &vec_1);
    g_rgb[idx] = sum;

    if (g_want_aov) {
        vec_t dir = px_dir(ul, inc_x, inc_y, (float)x, (float)y);
        int nd = 0, id = -1;
        float t = FLT_MAX;
        bvh_traverse(0, &cam.pos, &dir, &nd, &t, &id);
        g_id[idx] = id;
        g_t[idx] = t;
    }
}

/* cpu/src/main.c:241-264: whole rows handed out through one atomic counter */
static void* thread_main(void* arg)
{
    (void)arg;
    vec_t sp[3];
    cam_calculate_screen_coords(&cam, sp, (float)g_w / g_h);
    vec_t ul = sp[0], ur = sp[1], dl = sp[2];
    vec_t inc_x = vec_sub(&ur, &ul);
    inc_x = vec_div(&inc_x, g_w);
    vec_t inc_y = vec_sub(&dl, &ul);
    inc_y = vec_div(&inc_y, g_h);
    for (;;) {
        int y = atomic_fetch_add(&g_row_counter, 1);
        if (y >= g_h) break;
        for (int x = 0; x < g_w; x++) px_render(&ul, &inc_x, &inc_y, x, y);
    }
    return NULL;
}

static double now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec / 1e6;
}

static double frame(int threads)
{
    pthread_t* th = malloc(sizeof(pthread_t) * threads);
    double t0 = now_ms();
    atomic_store(&g_row_counter, 0);
    for (int i = 0; i < threads; i++) pthread_create(&th[i], NULL, thread_main, NULL);
    for (int i = 0; i < threads; i++) pthread_join(th[i], NULL);
    double t1 = now_ms();
    free(th);
    return t1 - t0;
}

/* ---- scene pack (.rtsc) writer/reader: see include/rt_b200.h for the layout ---- */
typedef struct { float k[9]; } mat9_t; /* ks, kd, kr */

static int write_rtsc(const char* path)
{
    mat9_t* mats = NULL;
    uint32_t nm = 0;
    uint32_t* midx = malloc(sizeof(uint32_t) * triangles_len);
    for (size_t i = 0; i < triangles_len; i++) {
        mat9_t m;
        memcpy(&m.k[0], &triangles[i].ks, 12);
        memcpy(&m.k[3], &triangles[i].kd, 12);
        memcpy(&m.k[6], &triangles[i].kr, 12);
        uint32_t j;
        for (j = 0; j < nm; j++)
            if (!memcmp(&mats[j], &m, sizeof m)) break;
        if (j == nm) {
            mats = realloc(mats, sizeof(mat9_t) * (nm + 1));
            mats[nm++] = m;
        }
        midx[i] = j;
    }
    FILE* f = fopen(path, "wb");
    if (!f) return -1;
    uint32_t hdr[4] = {(uint32_t)triangles_len, nm, (uint32_t)lights_len, 0};
    float amb[4] = {amb_light.r, amb_light.g, amb_light.b, 0};
    fwrite("RTSC0001", 1, 8, f);
    fwrite(hdr, 4, 4, f);
    fwrite(amb, 4, 4, f);
    for (size_t i = 0; i < triangles_len; i++) fwrite(triangles[i].coords, 4, 9, f);
    fwrite(midx, 4, triangles_len, f);
    fwrite(mats, sizeof(mat9_t), nm, f);
    for (size_t i = 0; i < lights_len; i++) {
        fwrite(&lights[i].pos, 4, 3, f);
        fwrite(&lights[i].kl, 4, 3, f);
    }
    fclose(f);
    free(mats);
    free(midx);
    return 0;
}

static int read_rtsc(const char* path)
{
    FILE* f = fopen(path, "rb");
    if (!f) return -1;
    char magic[8];
    uint32_t hdr[4];
    float amb[4];
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "RTSC0001", 8)) return -2;
    if (fread(hdr, 4, 4, f) != 4 || fread(amb, 4, 4, f) != 4) return -2;
    uint32_t nt = hdr[0], nm = hdr[1], nl = hdr[2];
    float* tri = malloc(36u * (size_t)nt);
    uint32_t* midx = malloc(4u * (size_t)nt);
    mat9_t* mats = malloc(sizeof(mat9_t) * (nm ? nm : 1));
    if (fread(tri, 36, nt, f) != nt || fread(midx, 4, nt, f) != nt || fread(mats, sizeof(mat9_t), nm, f) != nm)
        return -2;
    triangles_len = nt;
    triangles = malloc(sizeof(triangle_t) * (nt ? nt : 1));
    for (uint32_t i = 0; i < nt; i++) {
        vec_t a, b, c, ks, kd, kr;
        memcpy(&a, tri + 9 * (size_t)i, 12);
        memcpy(&b, tri + 9 * (size_t)i + 3, 12);
        memcpy(&c, tri + 9 * (size_t)i + 6, 12);
        memcpy(&ks, &mats[midx[i]].k[0], 12);
        memcpy(&kd, &mats[midx[i]].k[3], 12);
        memcpy(&kr, &mats[midx[i]].k[6], 12);
        triangle_init(&triangles[i], &a, &b, &c, &ks, &kd, &kr);
    }
    lights_len = nl;
    lights = malloc(sizeof(light_t) * (nl ? nl : 1));
    for (uint32_t i = 0; i < nl; i++) {
        float l[6];
        if (fread(l, 4, 6, f) != 6) return -2;
        memcpy(&lights[i].pos, l, 12);
        memcpy(&lights[i].kl, l + 3, 12);
    }
    amb_light.r = amb[0]; amb_light.g = amb[1]; amb_light.b = amb[2];
    fclose(f);
    free(tri); free(midx); free(mats);
    return 0;
}

/* the reference's synthetic triangle soup, cpu/src/main.c:115-131 (caller did srand) */
static void make_soup(int n)
{
    triangles_len = n;
    triangles = (triangle_t*)malloc(sizeof(triangle_t) * triangles_len);
    for (int i = 0; i < n; i++) {
        vec_t vec0 = {0.0f, 0.0f, 0.0f};
        vec_t vec1 = {1.0f, 1.0f, 1.0f};
        vec_t r0 = {(float)rand() / RAND_MAX, (float)rand() / RAND_MAX, (float)rand() / RAND_MAX};
        vec_t r1 = {(float)rand() / RAND_MAX, (float)rand() / RAND_MAX, (float)rand() / RAND_MAX};
        vec_t r2 = {(float)rand() / RAND_MAX, (float)rand() / RAND_MAX, (float)rand() / RAND_MAX};
        vec_t a = vec_mul(&r0, 10);
        a.x -= 5; a.y -= 5; a.z -= 5;
        vec_t b = vec_add(&a, &r1);
        vec_t c = vec_add(&b, &r2);
        triangle_init(&triangles[i], &a, &b, &c, &vec1, &vec0, &vec0);
    }
    lights_len = 0;
    lights = malloc(sizeof(light_t));
}

static void dump(const char* prefix, const char* ext, const void* p, size_t bytes)
{
    char path[1024];
    snprintf(path, sizeof path, "%s.%s", prefix, ext);
    FILE* f = fopen(path, "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", path); exit(2); }
    fwrite(p, 1, bytes, f);
    fclose(f);
}

int main(int argc, char** argv)
{
    const char* scene_dir = NULL; const char* rtsc = NULL; const char* out = NULL;
    const char* dump_bvh = NULL; const char* dump_scene = NULL;
    int soup = 0, threads = 1, frames = 1, warmup = 0, want_bmp = 0;
    float cp[3] = {0, -9, 3}, cr[3] = {(float)(-M_PI / 12), 0, 0};
    double fov = M_PI / 3.2;
    int default_cam = 1;
    g_w = 1920; g_h = 1080; g_spp = 1; g_seed = 1; g_want_aov = 0;

    for (int i = 1; i < argc; i++) {
        const char* a = argv[i];
#define NEXT (i + 1 < argc ? argv[++i] : (fprintf(stderr, "missing value for %s\n", a), exit(2), ""))
        if (!strcmp(a, "--scene-dir")) scene_dir = NEXT;
        else if (!strcmp(a, "--rtsc")) rtsc = NEXT;
        else if (!strcmp(a, "--soup")) soup = atoi(NEXT);
        else if (!strcmp(a, "--width")) g_w = atoi(NEXT);
        else if (!strcmp(a, "--height")) g_h = atoi(NEXT);
        else if (!strcmp(a, "--spp")) g_spp = atoi(NEXT);
        else if (!strcmp(a, "--seed")) g_seed = (uint32_t)strtoul(NEXT, NULL, 10);
        else if (!strcmp(a, "--threads")) threads = atoi(NEXT);
        else if (!strcmp(a, "--frames")) frames = atoi(NEXT);
        else if (!strcmp(a, "--warmup")) warmup = atoi(NEXT);
        else if (!strcmp(a, "--out")) { out = NEXT; g_want_aov = 1; }
        else if (!strcmp(a, "--bmp")) want_bmp = 1;
        else if (!strcmp(a, "--dump-bvh")) dump_bvh = NEXT;
        else if (!strcmp(a, "--dump-scene")) dump_scene = NEXT;
        else if (!strcmp(a, "--cam")) {
            default_cam = 0;
            for (int k = 0; k < 3; k++) cp[k] = (float)atof(NEXT);
            for (int k = 0; k < 3; k++) cr[k] = (float)atof(NEXT);
            fov = atof(NEXT);
        } else { fprintf(stderr, "unknown option %s\n", a); return 2; }
    }
    if (threads < 1 || g_w < 1 || g_h < 1 || g_spp < 1) { fprintf(stderr, "bad arguments\n"); return 2; }

    srand(1); /* SEED 1, cpu/src/main.c:91-95 */

    /* cpu/src/main.c:105-106 */
    cam_init(&cam, &(vec_t){cp[0], cp[1], cp[2]}, fov);
    if (default_cam) cam.rot.x = -M_PI / 12;
    else { cam.rot.x = cr[0]; cam.rot.y = cr[1]; cam.rot.z = cr[2]; }

    if (scene_dir) {
        char o[1024], m[1024], l[1024];
        snprintf(o, sizeof o, "%s/triangles.obj", scene_dir);
        snprintf(m, sizeof m, "%s/triangles.mtl", scene_dir);
        snprintf(l, sizeof l, "%s/lights.obj", scene_dir);
        triangles = triangles_load(o, m, &triangles_len);
        lights = lights_load(l, &lights_len);
    } else if (rtsc) {
        if (read_rtsc(rtsc)) { fprintf(stderr, "cannot read %s\n", rtsc); return 2; }
    } else if (soup > 0) {
        make_soup(soup);
    } else { fprintf(stderr, "need --scene-dir, --rtsc or --soup\n"); return 2; }

    if (dump_scene && write_rtsc(dump_scene)) { fprintf(stderr, "cannot write %s\n", dump_scene); return 2; }

    fflush(stdout);
    FILE* saved = stdout; /* bvh_build prints statistics; keep stdout a single JSON line */
    double b0 = now_ms();
    stdout = stderr;
    bvh_build(triangles, triangles_len);
    stdout = saved;
    double build_ms = now_ms() - b0;

    if (dump_bvh) {
        FILE* f = fopen(dump_bvh, "wb");
        if (!f) { fprintf(stderr, "cannot write %s\n", dump_bvh); return 2; }
        int32_t hdr[2] = {bvh_len, (int32_t)triangles_len};
        fwrite(hdr, 4, 2, f);
        fwrite(bvh, sizeof(bvh_t), bvh_len, f); /* 32 B: min[3] max[3] tr_len idx */
        fwrite(tri_idx, 4, triangles_len, f);
        fclose(f);
    }

    size_t npx = (size_t)g_w * g_h;
    g_rgb = malloc(sizeof(vec_t) * npx);
    g_id = malloc(4 * npx);
    g_t = malloc(4 * npx);

    for (int i = 0; i < warmup; i++) frame(threads);
    double* ms = malloc(sizeof(double) * (frames > 0 ? frames : 1));
    for (int i = 0; i < frames; i++) ms[i] = frame(threads);

    if (out && frames + warmup > 0) {
        dump(out, "id.i32", g_id, 4 * npx);
        dump(out, "t.f32", g_t, 4 * npx);
        dump(out, "rgb.f32", g_rgb, 12 * npx);
        /* byte-for-byte what cpu/src/bmp_writer.c:88-95 (vec_to_bgra) stores, top-down */
        uint8_t* bgra = malloc(4 * npx);
        for (size_t i = 0; i < npx; i++) {
            bgra[4 * i + 0] = (uint8_t)(g_rgb[i].b * 255.0f);
            bgra[4 * i + 1] = (uint8_t)(g_rgb[i].g * 255.0f);
            bgra[4 * i + 2] = (uint8_t)(g_rgb[i].r * 255.0f);
            bgra[4 * i + 3] = 255;
        }
        dump(out, "bgra.u8", bgra, 4 * npx);
        free(bgra);
        if (want_bmp) {
            char path[1024];
            snprintf(path, sizeof path, "%s.bmp", out);
            bmp_write_file(g_rgb, g_w, g_h, path);
        }
    }

    printf("{\"triangles\": %zu, \"lights\": %zu, \"bvh_nodes\": %d, \"bvh_build_ms\": %.3f, \"width\": %d, \"height\": %d, "
           "\"spp\": %d, \"threads\": %d, \"frame_ms\": [",
           triangles_len, lights_len, bvh_len, build_ms, g_w, g_h, g_spp, threads);
    for (int i = 0; i < frames; i++) printf("%s%.3f", i ? ", " : "", ms[i]);
    printf("]}\n");
    return 0;
}

// rt_render_main.cpp — C++ host driver above the C-ABI, mirroring the reference programs' main()
// (cpu/src/main.c:90-212, gpu/src/main.cu:80-140): load scene, build BVH, upload, WARMUP +
// ITERATIONS frames with per-frame times, statistics (mean, median, stddev, 99 % CI, FPS),
// download, write BMP.  Everything that is a compile-time #define in the reference's options.h is
// a command-line flag here.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "rt_b200.h"

static void usage()
{
    std::printf("usage: rt_render (--scene-dir DIR | --rtsc FILE | --soup N) [--width W] [--height H] [--spp S] [--seed K]\n"
                "                 [--bounces B] [--heuristic 6|0|1] [--refbin-tree] [--strict] [--gpus N] [--iterations I] [--warmup W]\n"
                "                 [--cam px py pz rx ry rz fov] [--out FILE.bmp] [--ctas-per-sm C] [--block T] [--refill R] [--peer-copy]\n"
                "                 [--gpu-build]   build the BVH and lay the scene out on the GPU (rt_create_gpu) instead of the host\n"
                "                 [--sequence N [--spin DZ]]   N frames end to end (camera rot.z += DZ per frame), each copied to the host\n"
                "                                              while the next one renders; the last one is written bottom-up as the BMP\n");
}

int main(int argc, char** argv)
{
    const char *scene_dir = nullptr, *rtsc = nullptr, *out = "render.bmp";
    int soup = 0, heuristic = 6, gpus = 1, iterations = 100, warmup = 50; // gpu/include/options.cuh:25-26
    int sequence = 0;
    float spin = 0.0f;
    bool gpu_build = false;
    rt_render_params p;
    rt_render_params_default(&p);
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { if (i + 1 >= argc) { usage(); std::exit(2); } return argv[++i]; };
        if (a == "--scene-dir") scene_dir = next();
        else if (a == "--rtsc") rtsc = next();
        else if (a == "--soup") soup = std::atoi(next());
        else if (a == "--width") p.width = std::atoi(next());
        else if (a == "--height") p.height = std::atoi(next());
        else if (a == "--spp") p.spp = std::atoi(next());
        else if (a == "--seed") p.seed = (uint32_t)std::strtoul(next(), nullptr, 10);
        else if (a == "--bounces") p.bounces = std::atoi(next());
        else if (a == "--heuristic") heuristic = std::atoi(next());
        else if (a == "--refbin-tree") heuristic |= RT_BVH_REFBIN;
        else if (a == "--strict") p.mode = RT_MODE_STRICT;
        else if (a == "--gpus") gpus = std::atoi(next());
        else if (a == "--iterations") iterations = std::atoi(next());
        else if (a == "--warmup") warmup = std::atoi(next());
        else if (a == "--out") out = next();
        else if (a == "--ctas-per-sm") p.ctas_per_sm = std::atoi(next());
        else if (a == "--block") p.block_threads = std::atoi(next());
        else if (a == "--refill") p.refill_threshold = std::atoi(next());
        else if (a == "--peer-copy") p.gather = RT_GATHER_PEER_COPY;
        else if (a == "--gpu-build") gpu_build = true;
        else if (a == "--sequence") sequence = std::atoi(next());
        else if (a == "--spin") spin = (float)std::atof(next());
        else if (a == "--cam") {
            for (int k = 0; k < 3; k++) p.cam.pos[k] = (float)std::atof(next());
            for (int k = 0; k < 3; k++) p.cam.rot[k] = (float)std::atof(next());
            p.cam.fov = (float)std::atof(next());
        } else { usage(); return 2; }
    }
    rt_scene* sc = nullptr;
    int rc;
    std::printf("Loading scene...\n");
    if (scene_dir) rc = rt_scene_load_dir(scene_dir, &sc);
    else if (rtsc) rc = rt_scene_load_rtsc(rtsc, &sc);
    else if (soup > 0) rc = rt_scene_soup((uint32_t)soup, 1, &sc);
    else { usage(); return 2; }
    if (rc) { std::fprintf(stderr, "%s\n", rt_last_error(nullptr)); return 1; }
    std::vector<int> devs(gpus);
    for (int i = 0; i < gpus; i++) devs[i] = i;
    rt_ctx* ctx = nullptr;
    rt_scene_desc d;
    std::printf("Building BVH...\n");
    if (gpu_build) {
        // bvh_build + load_to_gpu in one step on the device (cpu/src/main.c:138, gpu/src/main.cu:98-110)
        rt_bvh_gpu_stats st;
        if ((rc = rt_create_gpu(sc, heuristic, devs.data(), gpus, 0, &ctx, &st))) { std::fprintf(stderr, "%s\n", rt_last_error(nullptr)); return 1; }
        rt_scene_view(sc, &d);
        std::printf("BVH built on the GPU in %.2f ms (%u nodes, %d level passes, %d subtrees%s)\n", st.total_ms, st.nodes, st.levels, st.subtrees,
                    st.fell_back ? ", host fallback" : "");
        d.bvh_len = st.nodes;
    } else {
        if ((rc = rt_scene_build_bvh(sc, heuristic))) { std::fprintf(stderr, "%s\n", rt_last_error(nullptr)); return 1; }
        rt_scene_view(sc, &d);
    }
    std::printf("\n# Scene complexity #\nResolution: %d x %d\nNumber of triangles: %u\nNumber of lights: %u\nNumber of ray bounces: %d\nBVH nodes: %u\n",
                p.width, p.height, d.n_tris, d.n_lights, p.bounces, d.bvh_len);
    if (!ctx && (rc = rt_create(&d, devs.data(), gpus, &ctx))) { std::fprintf(stderr, "%s\n", rt_last_error(nullptr)); return 1; }

    std::printf("\nRendering...\n");
    std::vector<double> times;
    rt_timing tm;
    for (int i = 0; i < warmup + iterations; i++) {
        if ((rc = rt_render(ctx, &p, &tm))) { std::fprintf(stderr, "%s\n", rt_last_error(ctx)); return 1; }
        if (i >= warmup) times.push_back(tm.total_ms);
    }
    std::vector<uint8_t> bgra((size_t)p.width * p.height * 4);
    if ((rc = rt_download(ctx, bgra.data(), nullptr, nullptr, nullptr))) { std::fprintf(stderr, "%s\n", rt_last_error(ctx)); return 1; }
    if (rt_write_bmp(out, bgra.data(), p.width, p.height)) { std::fprintf(stderr, "%s\n", rt_last_error(nullptr)); return 1; }

    if (!times.empty()) { // cpu/src/main.c:45-88, 193-209
        const int n = (int)times.size();
        double mean = 0;
        for (double t : times) mean += t;
        mean /= n;
        std::vector<double> s = times;
        std::sort(s.begin(), s.end());
        const double median = n % 2 ? s[n / 2] : (s[n / 2 - 1] + s[n / 2]) / 2;
        double var = 0;
        for (double t : times) var += (t - mean) * (t - mean);
        const double sd = std::sqrt(var / n), ci = 2.5758293035489004 * sd / std::sqrt((double)n);
        const double rays = (double)(tm.rays_closest + tm.rays_shadow);
        std::printf("\n# Metrics #\nTotal execution time of %d frames: %.3f ms\n", n, mean * n);
        if (n >= 30) std::printf("Frame time (mean +/- 99%% CI): %.3f +/- %.3f = [%.3f, %.3f] ms\n", mean, ci, mean - ci, mean + ci);
        else std::printf("Frame time (mean): %.3f ms\n", mean);
        std::printf("Frame time (median): %.3f ms\nFrame time (stddev): %.3f ms^2\nExpected FPS: %.3f\n", median, sd, 1000 / mean);
        std::printf("Rays per frame: %.0f (closest %llu, shadow %llu)\nMrays/s (median frame): %.1f\n", rays,
                    (unsigned long long)tm.rays_closest, (unsigned long long)tm.rays_shadow, rays / median / 1e3);
    }
    if (sequence > 0) {
        // Frame sequence, end to end (cpu/src/main.c:169-185 with the moving camera of :107): frame k renders into frame
        // slot k % 2 while frame k-1 is copied into pinned host memory; rows are stored bottom-up by the kernel, so the
        // last host buffer is written as the BMP pixel array without a flip.
        void* host[RT_FRAME_SLOTS] = {nullptr, nullptr};
        const size_t bytes = (size_t)p.width * p.height * 4;
        for (int s = 0; s < RT_FRAME_SLOTS; s++)
            if ((rc = rt_host_alloc(bytes, &host[s]))) { std::fprintf(stderr, "%s\n", rt_last_error(nullptr)); return 1; }
        rt_render_params q = p;
        q.frame_flags |= RT_FRAME_BOTTOM_UP;
        if (q.gather == RT_GATHER_PEER_COPY) { std::fprintf(stderr, "--sequence needs the fused gather\n"); return 2; }
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        for (int k = 0; k < sequence; k++) {
            const int s = k % RT_FRAME_SLOTS;
            if (k >= RT_FRAME_SLOTS && (rc = rt_frame_wait(ctx, s, nullptr))) { std::fprintf(stderr, "%s\n", rt_last_error(ctx)); return 1; }
            q.frame_slot = s;
            q.cam.rot[2] = p.cam.rot[2] + spin * (float)k;
            if ((rc = rt_render_async(ctx, &q)) || (rc = rt_download_async(ctx, s, (uint8_t*)host[s]))) { std::fprintf(stderr, "%s\n", rt_last_error(ctx)); return 1; }
        }
        for (int s = 0; s < RT_FRAME_SLOTS && s < sequence; s++)
            if ((rc = rt_frame_wait(ctx, s, nullptr))) { std::fprintf(stderr, "%s\n", rt_last_error(ctx)); return 1; }
        clock_gettime(CLOCK_MONOTONIC, &t1);
        const double ms = (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) / 1e6;
        std::printf("\n# Sequence #\n%d frames rendered and copied to the host in %.3f ms: %.3f ms per frame, %.1f FPS end to end\n", sequence, ms,
                    ms / sequence, 1000.0 * sequence / ms);
        std::string seq_out = std::string(out) + ".last.bmp";
        if (rt_write_bmp_bottom_up(seq_out.c_str(), (const uint8_t*)host[(sequence - 1) % RT_FRAME_SLOTS], p.width, p.height)) {
            std::fprintf(stderr, "%s\n", rt_last_error(nullptr));
            return 1;
        }
        for (int s = 0; s < RT_FRAME_SLOTS; s++) rt_host_free(host[s]);
    }
    rt_destroy(ctx);
    rt_scene_free(sc);
    return 0;
}

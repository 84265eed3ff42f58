// scene.cpp — host scene ingest: the reference's OBJ/MTL/lights formats, the binary scene pack,
// the reference's synthetic soup and grid instancing.  Replaces triangles_load / lights_load
// (cpu/src/triangle.c:26-126, cpu/src/light.c:6-29) with their parsing rules kept:
//   * lines are read with a 256-byte fgets buffer (longer lines split, as in load_strings);
//   * vertices: lines starting "v " -> sscanf("v %f %f %f")                      (triangle.c:83-84)
//   * materials: at each "newmtl", only the NEXT FIVE lines are scanned for Kd/Ks/Kr
//     (triangle.c:58-67); at most 128 materials (triangle.c:89-90)
//   * "usemtl NAME" selects the first material with that exact name; an unknown name keeps the
//     previous material (triangle.c:97-108)
//   * faces: any line starting with 'f' -> sscanf("f %d %d %d"), 1-based      (triangle.c:109-113)
//   * lights: sscanf("%f %f %f %f %f %f") per line                               (light.c:18-24)
// Deliberate differences, all on inputs the reference mishandles:
//   * material fields a block does not set are 0 (the reference reads uninitialised stack);
//   * a face with fewer than three parsed indices or an out-of-range index is an error
//     (RT_ERR_IO) instead of undefined behaviour; a light line that does not parse is skipped
//     instead of appending garbage; nothing calls exit().
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "host_scene.h"

namespace rt {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const char* get_error() { return g_err.c_str(); }
} // namespace rt

namespace {

bool read_lines(const char* path, std::vector<std::string>& out)
{
    FILE* f = std::fopen(path, "r");
    if (!f) return false;
    char buf[256]; // same granularity as load_strings (cpu/src/triangle.c:34-42)
    while (std::fgets(buf, sizeof buf, f)) out.emplace_back(buf);
    std::fclose(f);
    return true;
}

struct Material { char name[256]; float kd[3], ks[3], kr[3]; };

} // namespace

extern "C" {

static int load_obj_serial(const char* obj_path, const char* mtl_path, const char* lights_path, rt_scene** out)
{
    if (!obj_path || !mtl_path || !out) { rt::set_error("rt_scene_load_obj: null argument"); return RT_ERR_INVALID; }
    std::vector<std::string> obj, mtl;
    if (!read_lines(obj_path, obj)) { rt::set_error(std::string("cannot load ") + obj_path); return RT_ERR_IO; }
    if (!read_lines(mtl_path, mtl)) { rt::set_error(std::string("cannot load ") + mtl_path); return RT_ERR_IO; }

    std::vector<float> verts;
    for (const std::string& l : obj) {
        if (l.size() >= 2 && l[0] == 'v' && l[1] == ' ') {
            float v[3] = {0, 0, 0};
            std::sscanf(l.c_str(), "v %f %f %f", &v[0], &v[1], &v[2]);
            verts.insert(verts.end(), v, v + 3);
        }
    }

    std::vector<Material> mats;
    for (size_t i = 0; i < mtl.size(); i++) {
        if (std::strncmp(mtl[i].c_str(), "newmtl", 6) == 0 && mats.size() < 128) {
            Material m;
            std::memset(&m, 0, sizeof m);
            std::sscanf(mtl[i].c_str(), "newmtl %255s", m.name);
            for (size_t j = i + 1; j < i + 6 && j < mtl.size(); j++) {
                const char* s = mtl[j].c_str();
                if (std::strncmp(s, "Kd", 2) == 0) std::sscanf(s, "Kd %f %f %f", &m.kd[0], &m.kd[1], &m.kd[2]);
                else if (std::strncmp(s, "Ks", 2) == 0) std::sscanf(s, "Ks %f %f %f", &m.ks[0], &m.ks[1], &m.ks[2]);
                else if (std::strncmp(s, "Kr", 2) == 0) std::sscanf(s, "Kr %f %f %f", &m.kr[0], &m.kr[1], &m.kr[2]);
            }
            mats.push_back(m);
        }
    }

    rt_scene* sc = new rt_scene();
    // material 0 = the all-zero "no usemtl yet" material (current_ks/kd/kr = {0}, triangle.c:92)
    sc->mats.assign(9, 0.0f);
    for (const Material& m : mats) {
        sc->mats.insert(sc->mats.end(), m.ks, m.ks + 3);
        sc->mats.insert(sc->mats.end(), m.kd, m.kd + 3);
        sc->mats.insert(sc->mats.end(), m.kr, m.kr + 3);
    }
    uint32_t current = 0;
    const int nv = (int)(verts.size() / 3);
    for (const std::string& l : obj) {
        if (std::strncmp(l.c_str(), "usemtl", 6) == 0) {
            char name[256] = {0};
            std::sscanf(l.c_str(), "usemtl %255s", name);
            for (size_t m = 0; m < mats.size(); m++)
                if (std::strcmp(name, mats[m].name) == 0) { current = (uint32_t)m + 1; break; }
        } else if (!l.empty() && l[0] == 'f') {
            int v[3] = {0, 0, 0};
            int got = std::sscanf(l.c_str(), "f %d %d %d", &v[0], &v[1], &v[2]);
            if (got != 3 || v[0] < 1 || v[1] < 1 || v[2] < 1 || v[0] > nv || v[1] > nv || v[2] > nv) {
                rt::set_error(std::string(obj_path) + ": unsupported face line: " + l);
                delete sc;
                return RT_ERR_IO;
            }
            for (int k = 0; k < 3; k++) sc->tri.insert(sc->tri.end(), &verts[3 * (size_t)(v[k] - 1)], &verts[3 * (size_t)(v[k] - 1)] + 3);
            sc->tri_mat.push_back(current);
        }
    }

    if (lights_path) {
        std::vector<std::string> ll;
        if (!read_lines(lights_path, ll)) {
            rt::set_error(std::string("cannot open ") + lights_path);
            delete sc;
            return RT_ERR_IO;
        }
        for (const std::string& l : ll) {
            float v[6];
            if (std::sscanf(l.c_str(), "%f %f %f %f %f %f", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5]) == 6)
                sc->lights.insert(sc->lights.end(), v, v + 6);
        }
    }
    *out = sc;
    return RT_OK;
}

} // extern "C"

// ---- the same loader at memory speed: one read of the file, line index, lines parsed on all host threads ----
// Semantics are those of the line-by-line version above (which stays as the checker: RT_LOADER_SERIAL=1, and
// tests/test_scene_host.py compares the two): lines are what fgets(buf, 256) would return (at most 255 characters,
// cut after a newline, text after a NUL ignored), "v " lines give up to three floats (missing ones are 0), any line
// starting with 'f' must give three in-range vertex numbers, "usemtl" switches the current material by exact name and
// an unknown name keeps the previous one.  Vertices may be defined after the faces that use them (two passes).
namespace {

struct LineRef { size_t pos; uint32_t len; };

bool read_file(const char* path, std::string& out)
{
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out.resize(sz > 0 ? (size_t)sz : 0);
    const bool ok = out.empty() || std::fread(&out[0], 1, out.size(), f) == out.size();
    std::fclose(f);
    return ok;
}

void split_lines(const std::string& data, std::vector<LineRef>& lines)
{
    const char* base = data.data();
    size_t pos = 0;
    const size_t n = data.size();
    while (pos < n) {
        const size_t maxend = std::min(pos + 255, n);
        const void* nl = std::memchr(base + pos, '\n', maxend - pos);
        const size_t end = nl ? (size_t)((const char*)nl - base) + 1 : maxend;
        lines.push_back({pos, (uint32_t)(end - pos)});
        pos = end;
    }
}

// copy of one line as the C string the serial loader sees (NUL-terminated, cut at an embedded NUL)
inline void line_cstr(const std::string& data, const LineRef& l, char* buf)
{
    std::memcpy(buf, data.data() + l.pos, l.len);
    buf[l.len] = 0;
}

// sscanf(s, "%f %f %f") after a literal prefix: up to three floats, stop at the first that does not parse
inline int scan_floats(const char* s, float* v, int n)
{
    int got = 0;
    for (; got < n; got++) {
        char* end;
        const float x = std::strtof(s, &end);
        if (end == s) break;
        v[got] = x;
        s = end;
    }
    return got;
}

inline int scan_ints(const char* s, int* v, int n)
{
    int got = 0;
    for (; got < n; got++) {
        char* end;
        const long x = std::strtol(s, &end, 10);
        if (end == s) break;
        v[got] = (int)x;
        s = end;
    }
    return got;
}

struct ObjItem { uint32_t line; int kind; int v[3]; }; // kind 0 = face, 1 = usemtl (v[0] = index into names), 2 = bad face

template <class F>
void run_chunks(size_t count, size_t chunks, F f)
{
    if (chunks <= 1) { f(0, (size_t)0, count); return; }
    std::vector<std::thread> th;
    for (size_t i = 0; i < chunks; i++) th.emplace_back([=] { f(i, count * i / chunks, count * (i + 1) / chunks); });
    for (auto& x : th) x.join();
}

} // namespace

extern "C" {

int rt_scene_load_obj(const char* obj_path, const char* mtl_path, const char* lights_path, rt_scene** out)
{
    return rt::guarded("rt_scene_load_obj", [&]() -> int {
    if (!obj_path || !mtl_path || !out) { rt::set_error("rt_scene_load_obj: null argument"); return RT_ERR_INVALID; }
    if (const char* e = std::getenv("RT_LOADER_SERIAL")) if (e[0] == '1') return load_obj_serial(obj_path, mtl_path, lights_path, out);
    std::string data;
    if (!read_file(obj_path, data)) { rt::set_error(std::string("cannot load ") + obj_path); return RT_ERR_IO; }
    std::vector<std::string> mtl;
    if (!read_lines(mtl_path, mtl)) { rt::set_error(std::string("cannot load ") + mtl_path); return RT_ERR_IO; }
    std::vector<LineRef> lines;
    split_lines(data, lines);

    // materials: a handful of lines, as in the serial loader
    std::vector<Material> mats;
    for (size_t i = 0; i < mtl.size(); i++) {
        if (std::strncmp(mtl[i].c_str(), "newmtl", 6) == 0 && mats.size() < 128) {
            Material m;
            std::memset(&m, 0, sizeof m);
            std::sscanf(mtl[i].c_str(), "newmtl %255s", m.name);
            for (size_t j = i + 1; j < i + 6 && j < mtl.size(); j++) {
                const char* s = mtl[j].c_str();
                if (std::strncmp(s, "Kd", 2) == 0) std::sscanf(s, "Kd %f %f %f", &m.kd[0], &m.kd[1], &m.kd[2]);
                else if (std::strncmp(s, "Ks", 2) == 0) std::sscanf(s, "Ks %f %f %f", &m.ks[0], &m.ks[1], &m.ks[2]);
                else if (std::strncmp(s, "Kr", 2) == 0) std::sscanf(s, "Kr %f %f %f", &m.kr[0], &m.kr[1], &m.kr[2]);
            }
            mats.push_back(m);
        }
    }

    size_t chunks = 1;
    {
        const char* e = std::getenv("RT_LOADER_CHUNKS"); // tests force many small chunks
        const size_t hw = e ? (size_t)std::atoi(e) : (size_t)std::thread::hardware_concurrency();
        chunks = std::max<size_t>(1, std::min(hw ? hw : 1, e ? lines.size() : lines.size() / 4096));
    }
    std::vector<std::vector<float>> cverts(chunks);
    std::vector<std::vector<ObjItem>> citems(chunks);
    std::vector<std::vector<std::string>> cnames(chunks);
    run_chunks(lines.size(), chunks, [&](size_t c, size_t lo, size_t hi) {
        char buf[256];
        for (size_t i = lo; i < hi; i++) {
            const char* p = data.data() + lines[i].pos;
            const uint32_t len = lines[i].len;
            if (len >= 2 && p[0] == 'v' && p[1] == ' ') {
                line_cstr(data, lines[i], buf);
                float v[3] = {0, 0, 0};
                scan_floats(buf + 1, v, 3);
                cverts[c].insert(cverts[c].end(), v, v + 3);
            } else if (len >= 6 && std::memcmp(p, "usemtl", 6) == 0) {
                line_cstr(data, lines[i], buf);
                char name[256] = {0};
                std::sscanf(buf, "usemtl %255s", name);
                cnames[c].push_back(name);
                citems[c].push_back({(uint32_t)i, 1, {(int)cnames[c].size() - 1, 0, 0}});
            } else if (len >= 1 && p[0] == 'f') {
                line_cstr(data, lines[i], buf);
                ObjItem it{(uint32_t)i, 0, {0, 0, 0}};
                if (scan_ints(buf + 1, it.v, 3) != 3) it.kind = 2;
                citems[c].push_back(it);
            }
        }
    });
    std::vector<float> verts;
    {
        size_t total = 0;
        for (auto& v : cverts) total += v.size();
        verts.reserve(total);
        for (auto& v : cverts) verts.insert(verts.end(), v.begin(), v.end());
    }
    const int nv = (int)(verts.size() / 3);

    // current material at the start of every chunk (serial over the few usemtl lines), face offsets
    std::vector<uint32_t> carry(chunks + 1, 0);
    std::vector<size_t> face_off(chunks + 1, 0);
    auto resolve = [&](const std::string& name, uint32_t cur) {
        for (size_t m = 0; m < mats.size(); m++)
            if (std::strcmp(name.c_str(), mats[m].name) == 0) return (uint32_t)m + 1;
        return cur; // an unknown name keeps the previous material (triangle.c:97-108)
    };
    for (size_t c = 0; c < chunks; c++) {
        uint32_t cur = carry[c];
        size_t faces = 0;
        for (const ObjItem& it : citems[c]) {
            if (it.kind == 1) cur = resolve(cnames[c][(size_t)it.v[0]], cur);
            else faces++;
        }
        carry[c + 1] = cur;
        face_off[c + 1] = face_off[c] + faces;
    }

    rt_scene* sc = new rt_scene();
    // material 0 = the all-zero "no usemtl yet" material (current_ks/kd/kr = {0}, triangle.c:92)
    sc->mats.assign(9, 0.0f);
    for (const Material& m : mats) {
        sc->mats.insert(sc->mats.end(), m.ks, m.ks + 3);
        sc->mats.insert(sc->mats.end(), m.kd, m.kd + 3);
        sc->mats.insert(sc->mats.end(), m.kr, m.kr + 3);
    }
    sc->tri.resize(9 * face_off[chunks]);
    sc->tri_mat.resize(face_off[chunks]);
    std::vector<uint32_t> first_bad(chunks, UINT32_MAX);
    run_chunks(chunks, chunks, [&](size_t, size_t lo, size_t hi) {
        for (size_t c = lo; c < hi; c++) {
            uint32_t cur = carry[c];
            size_t f = face_off[c];
            for (const ObjItem& it : citems[c]) {
                if (it.kind == 1) { cur = resolve(cnames[c][(size_t)it.v[0]], cur); continue; }
                const int* v = it.v;
                if (it.kind == 2 || v[0] < 1 || v[1] < 1 || v[2] < 1 || v[0] > nv || v[1] > nv || v[2] > nv) { first_bad[c] = it.line; break; }
                for (int k = 0; k < 3; k++) std::memcpy(&sc->tri[9 * f + 3 * (size_t)k], &verts[3 * (size_t)(v[k] - 1)], 12);
                sc->tri_mat[f] = cur;
                f++;
            }
        }
    });
    for (size_t c = 0; c < chunks; c++) {
        if (first_bad[c] == UINT32_MAX) continue;
        char buf[256];
        line_cstr(data, lines[first_bad[c]], buf);
        rt::set_error(std::string(obj_path) + ": unsupported face line: " + buf);
        delete sc;
        return RT_ERR_IO;
    }

    if (lights_path) {
        std::vector<std::string> ll;
        if (!read_lines(lights_path, ll)) {
            rt::set_error(std::string("cannot open ") + lights_path);
            delete sc;
            return RT_ERR_IO;
        }
        for (const std::string& l : ll) {
            float v[6];
            if (std::sscanf(l.c_str(), "%f %f %f %f %f %f", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5]) == 6)
                sc->lights.insert(sc->lights.end(), v, v + 6);
        }
    }
    *out = sc;
    return RT_OK;
    });
}

int rt_scene_load_dir(const char* dir, rt_scene** out)
{
    if (!dir || !out) { rt::set_error("rt_scene_load_dir: null argument"); return RT_ERR_INVALID; }
    std::string d(dir);
    return rt_scene_load_obj((d + "/triangles.obj").c_str(), (d + "/triangles.mtl").c_str(), (d + "/lights.obj").c_str(), out);
}

int rt_scene_load_rtsc(const char* path, rt_scene** out)
{
    if (!path || !out) { rt::set_error("rt_scene_load_rtsc: null argument"); return RT_ERR_INVALID; }
    FILE* f = std::fopen(path, "rb");
    if (!f) { rt::set_error(std::string("cannot open ") + path); return RT_ERR_IO; }
    char magic[8];
    uint32_t hdr[4];
    float amb[4];
    bool ok = std::fread(magic, 1, 8, f) == 8 && std::memcmp(magic, "RTSC0001", 8) == 0 &&
              std::fread(hdr, 4, 4, f) == 4 && std::fread(amb, 4, 4, f) == 4;
    if (ok) { // header counts against the file size: a corrupt header must not turn into a multi-GB allocation
        const long here = std::ftell(f);
        std::fseek(f, 0, SEEK_END);
        const long end = std::ftell(f);
        std::fseek(f, here, SEEK_SET);
        const unsigned long long need = 40ull * hdr[0] + 36ull * hdr[1] + 24ull * hdr[2];
        ok = here >= 0 && end >= here && need == (unsigned long long)(end - here);
    }
    rt_scene* sc = new (std::nothrow) rt_scene();
    if (!sc) { std::fclose(f); rt::set_error("out of memory"); return RT_ERR_NOMEM; }
    if (ok) try {
        sc->tri.resize(9 * (size_t)hdr[0]);
        sc->tri_mat.resize(hdr[0]);
        sc->mats.resize(9 * (size_t)hdr[1]);
        sc->lights.resize(6 * (size_t)hdr[2]);
        ok = std::fread(sc->tri.data(), 4, sc->tri.size(), f) == sc->tri.size() &&
             std::fread(sc->tri_mat.data(), 4, sc->tri_mat.size(), f) == sc->tri_mat.size() &&
             std::fread(sc->mats.data(), 4, sc->mats.size(), f) == sc->mats.size() &&
             std::fread(sc->lights.data(), 4, sc->lights.size(), f) == sc->lights.size();
        for (uint32_t m : sc->tri_mat) ok = ok && m < hdr[1];
        std::memcpy(sc->ambient, amb, 12);
    } catch (const std::bad_alloc&) { std::fclose(f); delete sc; rt::set_error(std::string(path) + ": out of memory"); return RT_ERR_NOMEM; }
    std::fclose(f);
    if (!ok) { delete sc; rt::set_error(std::string(path) + ": not a valid RTSC0001 scene pack"); return RT_ERR_IO; }
    *out = sc;
    return RT_OK;
}

int rt_scene_save_rtsc(const rt_scene* s, const char* path)
{
    if (!s || !path) { rt::set_error("rt_scene_save_rtsc: null argument"); return RT_ERR_INVALID; }
    FILE* f = std::fopen(path, "wb");
    if (!f) { rt::set_error(std::string("cannot write ") + path); return RT_ERR_IO; }
    uint32_t hdr[4] = {s->n_tris(), s->n_mats(), s->n_lights(), 0};
    float amb[4] = {s->ambient[0], s->ambient[1], s->ambient[2], 0};
    std::fwrite("RTSC0001", 1, 8, f);
    std::fwrite(hdr, 4, 4, f);
    std::fwrite(amb, 4, 4, f);
    std::fwrite(s->tri.data(), 4, s->tri.size(), f);
    std::fwrite(s->tri_mat.data(), 4, s->tri_mat.size(), f);
    std::fwrite(s->mats.data(), 4, s->mats.size(), f);
    std::fwrite(s->lights.data(), 4, s->lights.size(), f);
    bool ok = std::ferror(f) == 0;
    std::fclose(f);
    if (!ok) { rt::set_error(std::string("write failed: ") + path); return RT_ERR_IO; }
    return RT_OK;
}

int rt_scene_from_arrays(const rt_scene_desc* d, rt_scene** out)
{
    return rt::guarded("rt_scene_from_arrays", [&]() -> int {
    if (!d || !out || (d->n_tris && (!d->tri_coords)) || (d->n_mats && !d->materials) || (d->n_lights && !d->lights)) {
        rt::set_error("rt_scene_from_arrays: null argument");
        return RT_ERR_INVALID;
    }
    rt_scene* sc = new rt_scene();
    sc->tri.assign(d->tri_coords, d->tri_coords + 9 * (size_t)d->n_tris);
    if (d->tri_mat) sc->tri_mat.assign(d->tri_mat, d->tri_mat + d->n_tris);
    else sc->tri_mat.assign(d->n_tris, 0u);
    if (d->n_mats) sc->mats.assign(d->materials, d->materials + 9 * (size_t)d->n_mats);
    else sc->mats.assign(9, 0.0f);
    for (uint32_t m : sc->tri_mat)
        if (m >= sc->n_mats()) { delete sc; rt::set_error("rt_scene_from_arrays: material index out of range"); return RT_ERR_INVALID; }
    if (d->n_lights) sc->lights.assign(d->lights, d->lights + 6 * (size_t)d->n_lights);
    std::memcpy(sc->ambient, d->ambient, 12);
    if (d->bvh && d->tri_idx && d->bvh_len) {
        sc->bvh.assign(d->bvh, d->bvh + d->bvh_len);
        sc->tri_idx.assign(d->tri_idx, d->tri_idx + d->n_tris);
    }
    *out = sc;
    return RT_OK;
    });
}

int rt_scene_soup(uint32_t n_tris, uint32_t seed, rt_scene** out)
{
    return rt::guarded("rt_scene_soup", [&]() -> int {
    if (!out || !n_tris) { rt::set_error("rt_scene_soup: bad argument"); return RT_ERR_INVALID; }
    rt_scene* sc = new rt_scene();
    sc->tri.resize(9 * (size_t)n_tris);
    sc->tri_mat.assign(n_tris, 0u);
    const float m[9] = {1, 1, 1, 0, 0, 0, 0, 0, 0}; // ks = 1, kd = kr = 0 (cpu/src/main.c:119-120,128)
    sc->mats.assign(m, m + 9);
    std::srand(seed);
    for (uint32_t i = 0; i < n_tris; i++) { // cpu/src/main.c:118-129, same rand() call order
        float r[9];
        for (int k = 0; k < 9; k++) r[k] = (float)std::rand() / RAND_MAX;
        float* t = &sc->tri[9 * (size_t)i];
        for (int k = 0; k < 3; k++) {
            float a = r[k] * 10;
            a -= 5;
            float b = a + r[3 + k];
            float c = b + r[6 + k];
            t[k] = a; t[3 + k] = b; t[6 + k] = c;
        }
    }
    *out = sc;
    return RT_OK;
    });
}

int rt_scene_instance_grid(const rt_scene* base, uint32_t nx, uint32_t ny, uint32_t nz, const float pitch[3],
                           uint32_t light_every, rt_scene** out)
{
    return rt::guarded("rt_scene_instance_grid", [&]() -> int {
    if (!base || !out || !pitch || !nx || !ny || !nz) { rt::set_error("rt_scene_instance_grid: bad argument"); return RT_ERR_INVALID; }
    const uint64_t copies = (uint64_t)nx * ny * nz;
    const uint64_t total = copies * base->n_tris();
    if (total >= (1ull << 27)) { rt::set_error("rt_scene_instance_grid: more than 2^27 triangles"); return RT_ERR_INVALID; }
    rt_scene* sc = new rt_scene();
    sc->mats = base->mats;
    std::memcpy(sc->ambient, base->ambient, 12);
    sc->tri.resize(9 * (size_t)total);
    sc->tri_mat.resize((size_t)total);
    sc->lights = base->lights;
    size_t o = 0, inst = 0;
    // the grid is centred on the base scene so that the reference camera still looks at it
    for (uint32_t iz = 0; iz < nz; iz++)
        for (uint32_t iy = 0; iy < ny; iy++)
            for (uint32_t ix = 0; ix < nx; ix++, inst++) {
                const float off[3] = {((float)ix - 0.5f * (float)(nx - 1)) * pitch[0],
                                      ((float)iy - 0.5f * (float)(ny - 1)) * pitch[1],
                                      ((float)iz - 0.5f * (float)(nz - 1)) * pitch[2]};
                for (uint32_t t = 0; t < base->n_tris(); t++, o++) {
                    for (int k = 0; k < 9; k++) sc->tri[9 * o + k] = base->tri[9 * (size_t)t + k] + off[k % 3];
                    sc->tri_mat[o] = base->tri_mat[t];
                }
                if (light_every && inst && inst % light_every == 0)
                    for (uint32_t l = 0; l < base->n_lights(); l++) {
                        float v[6];
                        std::memcpy(v, &base->lights[6 * (size_t)l], 24);
                        for (int k = 0; k < 3; k++) v[k] += off[k];
                        sc->lights.insert(sc->lights.end(), v, v + 6);
                    }
            }
    *out = sc;
    return RT_OK;
    });
}

int rt_scene_view(const rt_scene* s, rt_scene_desc* out)
{
    if (!s || !out) { rt::set_error("rt_scene_view: null argument"); return RT_ERR_INVALID; }
    std::memset(out, 0, sizeof *out);
    out->tri_coords = s->tri.data();
    out->tri_mat = s->tri_mat.data();
    out->n_tris = s->n_tris();
    out->materials = s->mats.data();
    out->n_mats = s->n_mats();
    out->lights = s->lights.data();
    out->n_lights = s->n_lights();
    std::memcpy(out->ambient, s->ambient, 12);
    out->bvh = s->bvh.empty() ? nullptr : s->bvh.data();
    out->tri_idx = s->tri_idx.empty() ? nullptr : s->tri_idx.data();
    out->bvh_len = (uint32_t)s->bvh.size();
    return RT_OK;
}

void rt_scene_free(rt_scene* s) { delete s; }

int rt_scene_build_bvh(rt_scene* s, int heuristic)
{
    return rt::guarded("rt_scene_build_bvh", [&]() -> int {
    if (!s) { rt::set_error("rt_scene_build_bvh: null scene"); return RT_ERR_INVALID; }
    const rt::BvhArith arith = (heuristic & RT_BVH_REFBIN) ? rt::BVH_REFBIN : rt::BVH_IEEE;
    return rt::build_bvh(*s, heuristic & ~RT_BVH_REFBIN, arith, 0);
    });
}

} // extern "C"

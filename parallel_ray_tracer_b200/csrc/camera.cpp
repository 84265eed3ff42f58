// camera.cpp — pin-hole camera of the reference (cpu/src/cam.c:5-48) and the per-frame ray
// basis thread_render derives from it (cpu/src/main.c:241-250).  Computed once per frame on the
// host (the reference GPU kernel recomputes it in every pixel thread, gpu/src/gpu.cu:81-89).
// Same libm calls, same operation order as the reference; compile with -ffp-contract=off.
#include <cmath>

#include "camera.h"

namespace rt {

namespace {
struct V { float x, y, z; };
// cam_rotateX/Y/Z, cam.c:17-33
void rotX(const rt_camera& c, V& p) { V t = p; p.y = t.y * cosf(c.rot[0]) - t.z * sinf(c.rot[0]); p.z = t.y * sinf(c.rot[0]) + t.z * cosf(c.rot[0]); }
void rotY(const rt_camera& c, V& p) { V t = p; p.x = t.x * cosf(c.rot[1]) + t.z * sinf(c.rot[1]); p.z = -t.x * sinf(c.rot[1]) + t.z * cosf(c.rot[1]); }
void rotZ(const rt_camera& c, V& p) { V t = p; p.x = t.x * cosf(c.rot[2]) - t.y * sinf(c.rot[2]); p.y = t.x * sinf(c.rot[2]) + t.y * cosf(c.rot[2]); }
} // namespace

void camera_basis(const rt_camera& cam, int width, int height, CameraBasis& out)
{
    const float focal = 1.0 / tanf(cam.fov / 2.0f);        // cam_init, cam.c:8 (double divide, float store)
    const float aspect = (float)width / height;            // main.c:243
    V p[3] = {{-1 * aspect, focal, +1}, {+1 * aspect, focal, +1}, {-1 * aspect, focal, -1}}; // cam.c:36-38
    for (int i = 0; i < 3; i++) {
        rotY(cam, p[i]); rotX(cam, p[i]); rotZ(cam, p[i]); // cam_rotate, cam.c:11-15
        p[i].x = p[i].x + cam.pos[0]; p[i].y = p[i].y + cam.pos[1]; p[i].z = p[i].z + cam.pos[2]; // cam.c:44-46
    }
    const V ul = p[0], ur = p[1], dl = p[2];
    out.pos[0] = cam.pos[0]; out.pos[1] = cam.pos[1]; out.pos[2] = cam.pos[2];
    out.ul[0] = ul.x; out.ul[1] = ul.y; out.ul[2] = ul.z;
    const float w = (float)width, h = (float)height;
    out.inc_x[0] = (ur.x - ul.x) / w; out.inc_x[1] = (ur.y - ul.y) / w; out.inc_x[2] = (ur.z - ul.z) / w; // main.c:247-248
    out.inc_y[0] = (dl.x - ul.x) / h; out.inc_y[1] = (dl.y - ul.y) / h; out.inc_y[2] = (dl.z - ul.z) / h; // main.c:249-250
}

} // namespace rt

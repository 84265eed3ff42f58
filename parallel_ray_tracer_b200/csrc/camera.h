#pragma once
#include "rt_b200.h"
namespace rt {
struct CameraBasis { float pos[3], ul[3], inc_x[3], inc_y[3]; };
void camera_basis(const rt_camera& cam, int width, int height, CameraBasis& out);
}

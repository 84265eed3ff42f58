// bmp.cpp — BMP output with the reference's file layout (cpu/src/bmp_writer.c:97-175): 14-byte
// file header + 40-byte BITMAPINFOHEADER, 32 bpp, no compression, rows bottom-up, file size
// 54 + 4*W*H.  The float -> BGRA conversion (vec_to_bgra, :88-95) is not here: the render kernel
// fuses it into its writeback, so this writer only flips rows.
#include <cstdio>
#include <cstring>
#include <vector>

#include "host_scene.h"

static int write_bmp_rows(const char* path, const uint8_t* bgra, int width, int height, bool rows_bottom_up)
{
    if (!path || !bgra || width <= 0 || height <= 0) { rt::set_error("rt_write_bmp: bad parameters"); return RT_ERR_INVALID; }
    const int row = width * 4, header = 14 + 40;
    const int file_size = header + row * height;
    uint8_t b[14 + 40];
    std::memset(b, 0, sizeof b);
    b[0] = 'B'; b[1] = 'M';
    std::memcpy(b + 0x02, &file_size, 4);
    std::memcpy(b + 0x0A, &header, 4);
    const int dib = 40;
    const uint16_t planes = 1, bpp = 32;
    std::memcpy(b + 0x0E, &dib, 4);
    std::memcpy(b + 0x12, &width, 4);
    std::memcpy(b + 0x16, &height, 4);
    std::memcpy(b + 0x1A, &planes, 2);
    std::memcpy(b + 0x1C, &bpp, 2);
    FILE* f = std::fopen(path, "wb");
    if (!f) { rt::set_error(std::string("Unable to open the BMP file ") + path); return RT_ERR_IO; }
    bool ok = std::fwrite(b, 1, sizeof b, f) == sizeof b;
    if (rows_bottom_up) {
        // the frame was stored in BMP row order by the kernel (RT_FRAME_BOTTOM_UP): the pixel array is written as is
        ok = ok && std::fwrite(bgra, 1, (size_t)row * height, f) == (size_t)row * height;
    } else {
        for (int y = 0; y < height && ok; y++) // bottom-up
            ok = std::fwrite(bgra + (size_t)(height - 1 - y) * row, 1, (size_t)row, f) == (size_t)row;
    }
    std::fclose(f);
    if (!ok) { rt::set_error("Unable to save BMP buffer to disk"); return RT_ERR_IO; }
    return RT_OK;
}

extern "C" int rt_write_bmp(const char* path, const uint8_t* bgra, int width, int height)
{
    return write_bmp_rows(path, bgra, width, height, false);
}

extern "C" int rt_write_bmp_bottom_up(const char* path, const uint8_t* bgra, int width, int height)
{
    return write_bmp_rows(path, bgra, width, height, true);
}

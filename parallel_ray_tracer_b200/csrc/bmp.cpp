// bmp.cpp — BMP output with the reference's file layout (cpu/src/bmp_writer.c:97-175): 14-byte
// file header + 40-byte BITMAPINFOHEADER, 32 bpp, no compression, rows bottom-up, file size
// 54 + 4*W*H.  The float -> BGRA conversion (vec_to_bgra, :88-95) is not here: the render kernel
// fuses it into its writeback, so this writer only flips rows.
#include <cstdio>
#include <cstring>
#include <vector>

#include "host_scene.h"

extern "C" int rt_write_bmp(const char* path, const uint8_t* bgra, int width, int height)
{
    if (!path || !bgra || width <= 0 || height <= 0) { rt::set_error("rt_write_bmp: bad parameters"); return RT_ERR_INVALID; }
    const int row = width * 4, header = 14 + 40;
    const int file_size = header + row * height;
    std::vector<uint8_t> buf((size_t)file_size, 0);
    uint8_t* b = buf.data();
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
    for (int y = 0; y < height; y++) // bottom-up
        std::memcpy(b + header + (size_t)y * row, bgra + (size_t)(height - 1 - y) * row, (size_t)row);
    FILE* f = std::fopen(path, "wb");
    if (!f) { rt::set_error(std::string("Unable to open the BMP file ") + path); return RT_ERR_IO; }
    const bool ok = std::fwrite(b, 1, buf.size(), f) == buf.size();
    std::fclose(f);
    if (!ok) { rt::set_error("Unable to save BMP buffer to disk"); return RT_ERR_IO; }
    return RT_OK;
}

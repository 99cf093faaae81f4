// TEST HARNESS: decodes an image file through include/fountain_imageio.hpp and dumps (w, h, rgb f32) for
// tests/test_imageio.py to compare with the Python host's decoder.
//   imageio_dump <file> <out.bin> [scale gamma(-1|0|1) flip_y(0|1)]
#include <cstdio>
#include <cstdlib>
#include "fountain_imageio.hpp"

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s <file> <out.bin> [scale gamma flip_y]\n", argv[0]); return 2; }
    try {
        fountain::imageio::ImageTexInfo info;
        info.filename = argv[1];
        fountain::imageio::Image img;
        if (argc >= 6) { info.scale = (float)std::atof(argv[3]); info.gamma = std::atoi(argv[4]); info.flip_y = std::atoi(argv[5]) != 0; img = fountain::imageio::load_texels(info); }
        else img = fountain::imageio::load_image(info.filename);
        FILE* f = std::fopen(argv[2], "wb");
        if (!f) return 3;
        const int32_t wh[2] = {img.width, img.height};
        std::fwrite(wh, 4, 2, f);
        std::fwrite(img.rgb.data(), 4, img.rgb.size(), f);
        std::fclose(f);
    } catch (const fountain::Error& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}

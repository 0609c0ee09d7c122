// TEST INFRASTRUCTURE (authoring container only): pins nbvh_write_exr_host by EXECUTION of the reference's own EXR dependency.
// Drives the vendored tinyexr (/root/reference/ext/tinyexr/tinyexr.h, included where it lies) with the call sequence of the
// reference's write_exr (/root/reference/include/base/bitmap.h:97-182: channels A,B,G,R, FLOAT, no compression, every value divided
// by rgb_renormalisation), on a seeded image, and (mode "load") decodes a file written by the product so the test can compare pixels.
//   exr_golden save <out.exr> <W> <H> <renorm> <seed>      -> writes the reference's file + <out.exr>.rgba (the float input)
//   exr_golden load <in.exr> <out.rgba>                    -> LoadEXR (RGBA float, as the reference's load path reads images)
#define TINYEXR_IMPLEMENTATION
#define TINYEXR_USE_MINIZ 0
#define TINYEXR_USE_STB_ZLIB 0
#include <zlib.h>
#include "tinyexr.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

int main(int argc, char **argv)
{
    if (argc >= 7 && !strcmp(argv[1], "save")) {
        const size_t W = atoi(argv[3]), H = atoi(argv[4]);
        const float renorm = (float)atof(argv[5]);
        uint32_t s = (uint32_t)atoi(argv[6]);
        std::vector<float> px(4 * W * H);
        for (auto &v : px) {
            s = s * 1664525u + 1013904223u;
            v = (float)(s >> 8) * (1.0f / 16777216.0f) * 37.5f - 1.25f;
        }
        FILE *f = fopen((std::string(argv[2]) + ".rgba").c_str(), "wb");
        fwrite(px.data(), 4, px.size(), f);
        fclose(f);
        const size_t n = W * H;
        std::vector<float> sep(4 * n);
        for (size_t c = 0; c < 4; ++c)
            for (size_t i = 0; i < n; ++i) {
                sep[c * n + i] = px[4 * i + c];
                sep[c * n + i] /= renorm;
            }
        float *ptrs[4] = {sep.data() + 3 * n, sep.data() + 2 * n, sep.data() + n, sep.data()};
        EXRChannelInfo ci[4];
        memset(ci, 0, sizeof(ci));
        ci[0].name[0] = 'A', ci[1].name[0] = 'B', ci[2].name[0] = 'G', ci[3].name[0] = 'R';
        int types[4] = {TINYEXR_PIXELTYPE_FLOAT, TINYEXR_PIXELTYPE_FLOAT, TINYEXR_PIXELTYPE_FLOAT, TINYEXR_PIXELTYPE_FLOAT};
        EXRHeader h;
        InitEXRHeader(&h);
        h.num_channels = 4, h.channels = ci, h.pixel_types = types, h.requested_pixel_types = types;
        h.compression_type = TINYEXR_COMPRESSIONTYPE_NONE;
        EXRImage img;
        InitEXRImage(&img);
        img.num_channels = 4, img.images = reinterpret_cast<unsigned char **>(ptrs), img.width = (int)W, img.height = (int)H;
        return SaveEXRImageToFile(&img, &h, argv[2], nullptr) == TINYEXR_SUCCESS ? 0 : 1;
    }
    if (argc >= 4 && !strcmp(argv[1], "load")) {
        float *rgba = nullptr;
        int w = 0, h = 0;
        const char *err = nullptr;
        if (LoadEXR(&rgba, &w, &h, argv[2], &err) != TINYEXR_SUCCESS) {
            fprintf(stderr, "LoadEXR: %s\n", err ? err : "?");
            return 1;
        }
        FILE *f = fopen(argv[3], "wb");
        fwrite(rgba, 4, (size_t)4 * w * h, f);
        fclose(f);
        printf("%d %d\n", w, h);
        return 0;
    }
    fprintf(stderr, "usage: exr_golden save|load ...\n");
    return 2;
}

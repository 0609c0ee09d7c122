// TEST INFRASTRUCTURE (built in the authoring container, runs on the GPU box): pins nbvh_flip_error by EXECUTION of the reference's own
// vendored flip-cuda (/root/reference/ext/flip-cuda/FLIP.h, included where it lies), driven like
// NeuralBVHRenderer::compute_and_display_flip_error (/root/reference/src/cuda/neural/neural_bvh_renderer.cu:40-96): both images go
// LinearRGB -> sRGB, then image<float>::FLIP(reference, test, PPD) with PPD = 0.7 * (3840 / 0.7) * pi / 180, and the mean is pooled.
//   flip_harness <W> <H> <reference.rgb32f> <test.rgb32f> <out_error_map.f32>     (linear RGB, 3 floats per pixel, row major)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "FLIP.h" // (the header relies on its includer for <vector> etc., as the reference's renderer provides them)

int main(int argc, char **argv)
{
    if (argc < 6) {
        fprintf(stderr, "usage: flip_harness W H reference.rgb32f test.rgb32f out.f32\n");
        return 2;
    }
    const int W = atoi(argv[1]), H = atoi(argv[2]);
    std::vector<float> a((size_t)3 * W * H), b(a.size());
    FILE *f = fopen(argv[3], "rb");
    if (!f || fread(a.data(), 4, a.size(), f) != a.size())
        return 1;
    fclose(f);
    f = fopen(argv[4], "rb");
    if (!f || fread(b.data(), 4, b.size(), f) != b.size())
        return 1;
    fclose(f);
    FLIP::image<FLIP::color3> reference(W, H), test(W, H);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const size_t i = 3 * ((size_t)y * W + x);
            reference.set(x, y, FLIP::color3(a[i], a[i + 1], a[i + 2]));
            test.set(x, y, FLIP::color3(b[i], b[i + 1], b[i + 2]));
        }
    reference.LinearRGB2sRGB();
    test.LinearRGB2sRGB();
    const float ppd = 0.7f * (3840.0f / 0.7f) * (float(FLIP::PI) / 180.0f);
    FLIP::image<float> err(W, H, 0.0f);
    err.FLIP(reference, test, ppd);
    std::vector<float> out((size_t)W * H);
    double sum = 0;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
            sum += (out[(size_t)y * W + x] = err.get(x, y));
    f = fopen(argv[5], "wb");
    fwrite(out.data(), 4, out.size(), f);
    fclose(f);
    printf("{\"W\": %d, \"H\": %d, \"ppd\": %.6f, \"mean_flip\": %.9g}\n", W, H, ppd, sum / out.size());
    return 0;
}

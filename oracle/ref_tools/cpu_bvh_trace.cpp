// TEST / BENCH INFRASTRUCTURE -- the reference's CPU path for BASELINE.json config 1, timed by
// bench.py's `cpu_baseline` and `--impl reference` legs.  Never linked into the product.
//
// Compiled against the reference's OWN vendored ext/bvh headers where they lie under /root/reference
// (recipe: oracle/Makefile.ref); only the binary (oracle/_ref/cpu_bvh_trace) travels to the GPU box.
//
// What it does: loads the flattened Chess meshes written by tools/make_fixtures.py, builds one
// BVH per mesh with bvh::v2::DefaultBuilder (Quality::High, leaf 1..2 -- /root/reference/src/cuda/cuda_backend.cpp:270-310),
// then traces closest-hit primary rays of the shipped camera through bvh::v2::Bvh::intersect<false,false> with
// PrecomputedTri and permuted primitives (usage pattern /root/reference/ext/bvh/test/simple_example.cpp:57-92),
// image rows spread over the host threads by bvh::v2::ParallelExecutor.  The neural mesh is traced as ordinary
// triangles (== ignore_neural_blas).
//
// usage: cpu_bvh_trace <fixture_dir> <width> <height> <first_pixel> <n_pixels> <repeats> [dump_file]
// prints one JSON line: {"mrays_s":..,"best_s":..,"threads":..,"build_s":..,"hits":..,"n_rays":..}
#include <bvh/v2/bvh.h>
#include <bvh/v2/default_builder.h>
#include <bvh/v2/executor.h>
#include <bvh/v2/node.h>
#include <bvh/v2/ray.h>
#include <bvh/v2/stack.h>
#include <bvh/v2/thread_pool.h>
#include <bvh/v2/tri.h>
#include <bvh/v2/vec.h>

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

using Scalar         = float;
using Vec3           = bvh::v2::Vec<Scalar, 3>;
using BBox           = bvh::v2::BBox<Scalar, 3>;
using Tri            = bvh::v2::Tri<Scalar, 3>;
using Node           = bvh::v2::Node<Scalar, 3>;
using Bvh            = bvh::v2::Bvh<Node>;
using Ray            = bvh::v2::Ray<Scalar, 3>;
using PrecomputedTri = bvh::v2::PrecomputedTri<Scalar>;

template <typename T>
static std::vector<T> slurp(const std::string &path)
{
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) {
        fprintf(stderr, "cannot read %s\n", path.c_str());
        exit(2);
    }
    const size_t bytes = (size_t)f.tellg();
    std::vector<T> v(bytes / sizeof(T));
    f.seekg(0);
    f.read(reinterpret_cast<char *>(v.data()), bytes);
    return v;
}

// RNG of the reference's raygen (include/cuda/utils/lcg_rng.cuh:12-72)
static uint32_t murmur_mix(uint32_t hash, uint32_t k)
{
    k *= 0xcc9e2d51u;
    k = (k << 15) | (k >> 17);
    k *= 0x1b873593u;
    hash ^= k;
    return ((hash << 13) | (hash >> 19)) * 5u + 0xe6546b64u;
}
static uint32_t rng_init(uint32_t index, uint32_t frame, uint32_t linear)
{
    uint32_t h = murmur_mix(0u, linear);
    h          = murmur_mix(h, 512u * frame + index);
    h ^= h >> 16, h *= 0x85ebca6bu, h ^= h >> 13, h *= 0xc2b2ae35u, h ^= h >> 16;
    return h;
}
static float rng_next(uint32_t &s)
{
    s = s * 1664525u + 1013904223u;
    return (float)s * 0x1p-32f;
}

struct MeshAccel {
    Bvh bvh;
    std::vector<PrecomputedTri> tris;
};

int main(int argc, char **argv)
{
    if (argc < 7) {
        fprintf(stderr, "usage: cpu_bvh_trace <fixture_dir> <w> <h> <first_pixel> <n_pixels> <repeats> [dump]\n");
        return 2;
    }
    const std::string root = argv[1];
    const uint32_t W = atoi(argv[2]), H = atoi(argv[3]);
    const uint32_t first = (uint32_t)atoll(argv[4]);
    uint32_t n = (uint32_t)atoll(argv[5]);
    const int repeats = atoi(argv[6]);
    if (n == 0 || first + n > W * H)
        n = W * H - first;
    const auto cam = slurp<float>(root + "/camera_params.f32"); // origin, du, dv, top_left

    bvh::v2::ThreadPool pool;
    bvh::v2::ParallelExecutor executor(pool);
    const unsigned threads = std::thread::hardware_concurrency();

    const auto t_build0 = std::chrono::steady_clock::now();
    std::vector<MeshAccel> meshes;
    for (int mi = 0;; mi++) {
        const std::string p = root + "/mesh" + std::to_string(mi);
        if (!std::ifstream(p + "_vertices.f32"))
            break;
        const auto verts = slurp<float>(p + "_vertices.f32");
        const auto idx   = slurp<uint32_t>(p + "_indices.u32");
        const size_t nt  = idx.size() / 3;
        std::vector<Tri> tris(nt);
        std::vector<BBox> bboxes(nt);
        std::vector<Vec3> centers(nt);
        executor.for_each(0, nt, [&](size_t b, size_t e) {
            for (size_t i = b; i < e; ++i) {
                const float *a = &verts[3 * idx[3 * i]], *bb = &verts[3 * idx[3 * i + 1]], *c = &verts[3 * idx[3 * i + 2]];
                tris[i]    = Tri(Vec3(a[0], a[1], a[2]), Vec3(bb[0], bb[1], bb[2]), Vec3(c[0], c[1], c[2]));
                bboxes[i]  = tris[i].get_bbox();
                centers[i] = tris[i].get_center();
            }
        });
        typename bvh::v2::DefaultBuilder<Node>::Config cfg;
        cfg.quality       = bvh::v2::DefaultBuilder<Node>::Quality::High;
        cfg.min_leaf_size = 1;
        cfg.max_leaf_size = 2;
        MeshAccel acc;
        acc.bvh = bvh::v2::DefaultBuilder<Node>::build(pool, bboxes, centers, cfg);
        acc.tris.resize(nt);
        executor.for_each(0, nt, [&](size_t b, size_t e) {
            for (size_t i = b; i < e; ++i)
                acc.tris[i] = tris[acc.bvh.prim_ids[i]];
        });
        meshes.push_back(std::move(acc));
    }
    const double build_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_build0).count();

    std::vector<float> out_t(n);
    std::vector<uint32_t> out_prim(n);
    double best = 1e30;
    std::vector<double> rep_s;
    for (int rep = 0; rep < repeats; rep++) {
        const auto t0 = std::chrono::steady_clock::now();
        executor.for_each(0, n, [&](size_t b, size_t e) {
            bvh::v2::SmallStack<Bvh::Index, 64> stack;
            for (size_t i = b; i < e; ++i) {
                const uint32_t pixel = first + (uint32_t)i;
                const uint32_t px = pixel % W, py = pixel / W;
                uint32_t rng   = rng_init(0u, 0u, pixel);
                const float xf = (rng_next(rng) + (float)px) / (float)W;
                const float yf = (rng_next(rng) + (float)py) / (float)H;
                float d[3];
                for (int k = 0; k < 3; k++)
                    d[k] = xf * cam[3 + k] + yf * cam[6 + k] + cam[9 + k];
                const float inv = 1.f / std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                Ray ray{Vec3(cam[0], cam[1], cam[2]), Vec3(d[0] * inv, d[1] * inv, d[2] * inv), 1e-4f,
                        std::numeric_limits<float>::max()};
                uint32_t hit_id = 0xFFFFFFFFu;
                for (size_t mi = 0; mi < meshes.size(); mi++) {
                    const MeshAccel &m = meshes[mi];
                    stack.size         = 0;
                    m.bvh.intersect<false, false>(ray, m.bvh.get_root().index, stack, [&](size_t begin, size_t end) {
                        bool any = false;
                        for (size_t j = begin; j < end; ++j) {
                            if (auto hit = m.tris[j].intersect(ray)) { // shrinks ray.tmax
                                hit_id = ((uint32_t)mi << 28) | (uint32_t)m.bvh.prim_ids[j];
                                any    = true;
                            }
                        }
                        return any;
                    });
                }
                out_t[i]    = hit_id == 0xFFFFFFFFu ? std::numeric_limits<float>::max() : ray.tmax;
                out_prim[i] = hit_id;
            }
        });
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (s < best)
            best = s;
        rep_s.push_back(s);
    }
    size_t hits = 0;
    for (uint32_t i = 0; i < n; i++)
        hits += out_prim[i] != 0xFFFFFFFFu;
    if (argc > 7) {
        std::ofstream f(argv[7], std::ios::binary);
        f.write(reinterpret_cast<const char *>(out_t.data()), sizeof(float) * n);
        f.write(reinterpret_cast<const char *>(out_prim.data()), sizeof(uint32_t) * n);
    }
    printf("{\"mrays_s\": %.4f, \"best_s\": %.6f, \"threads\": %u, \"build_s\": %.3f, \"hits\": %zu, \"n_rays\": %u, \"rep_s\": [",
           n / best / 1e6, best, threads, build_s, hits, n);
    for (size_t i = 0; i < rep_s.size(); i++)
        printf("%s%.6f", i ? ", " : "", rep_s[i]);
    printf("]}\n");
    return 0;
}

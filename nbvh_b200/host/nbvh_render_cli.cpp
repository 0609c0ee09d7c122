// nbvh_render_cli.cpp -- a C++ host of libnbvh_b200.so: what the reference's CudaBackend + NeuralBVHRenderer do around the module,
// reduced to a command line tool.  Loads a scene directory in the layout tools/make_fixtures.py writes (config JSON, .nbvh tree-cut,
// network_models JSON unchanged from the reference; TLAS / BLAS arrays as raw files), uploads it (CudaBackend::build_and_upload,
// /root/reference/src/cuda/cuda_backend.cpp:186-441), renders `spp` samples through nbvh_b200::NeuralPathModule::neural_render
// (== NeuralBVHRenderer::render -> neural_render, neural_bvh_renderer.cu:292-312) and writes the 8-bit frame as a PPM, or (out.exr) the accumulation / spp
// as the reference's save_image does.
//
// usage: nbvh_render_cli <scene_dir> <base_config_name> <out.ppm | out.exr> [W H] [spp] [max_depth]
// build: g++ -std=c++17 -O2 -Iinclude nbvh_b200/host/nbvh_render_cli.cpp -Lnbvh_b200 -lnbvh_b200 -L/usr/local/cuda/lib64 -lcudart
#include "nbvh_b200.hpp"

#include <cuda_runtime_api.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

static std::vector<uint8_t> slurp(const std::string &path, bool required = true)
{
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) {
        if (required)
            throw std::runtime_error("cannot read " + path);
        return {};
    }
    std::vector<uint8_t> v((size_t)f.tellg());
    f.seekg(0);
    f.read(reinterpret_cast<char *>(v.data()), (std::streamsize)v.size());
    return v;
}

template <typename T = void>
static T *upload(const void *host, size_t bytes)
{
    void *d = nullptr;
    if (cudaMalloc(&d, bytes ? bytes : 16) != cudaSuccess || (bytes && cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice) != cudaSuccess))
        throw std::runtime_error("device upload failed");
    return reinterpret_cast<T *>(d);
}

int main(int argc, char **argv)
{
    if (argc < 4) {
        fprintf(stderr, "usage: nbvh_render_cli <scene_dir> <base_config_name> <out.ppm> [W H] [spp] [max_depth]\n");
        return 2;
    }
    try {
        const std::string dir = argv[1], base = argv[2], out = argv[3];
        const auto cfg = slurp(dir + "/" + base + "_neural.json");
        nbvh_b200::NeuralPathModule mod(0, std::string(cfg.begin(), cfg.end()));
        mod.load_neural_bvh(dir + "/" + base + ".nbvh");
        mod.load_network_model(dir + "/" + base + "_network.json");
        // scene: TLAS + per-mesh BLAS (the neural mesh's own BLAS is not needed for inference: a one-node placeholder)
        // TLAS: the host-built files when present (ext/bvh, like CudaBackend::build_tlas), otherwise built on the device below
        const auto tlas = slurp(dir + "/tlas_nodes.bin", false), tlas_idx = slurp(dir + "/tlas_indices.u32", false);
        std::vector<nbvh_cuda_blas> meshes;
        for (int m = 0;; ++m) {
            const std::string p = dir + "/mesh" + std::to_string(m);
            auto nodes = slurp(p + "_bvh_nodes.bin", false);
            if (nodes.empty())
                break;
            nbvh_cuda_blas b{};
            if (m == 0) { // neural mesh (manifest.json: neural_mesh 0)
                // one-node placeholder that keeps the real root box (the device TLAS build reads it) and holds one degenerate triangle
                std::vector<uint8_t> one(nodes.begin(), nodes.begin() + 32), tri(36, 0), idx(12, 0), prim(8, 0);
                reinterpret_cast<nbvh_node32 *>(one.data())->prim_count_or_left = 1;
                reinterpret_cast<nbvh_node32 *>(one.data())->left_or_prim_offset = 0;
                b.num_nodes = 1, b.nodes = upload<nbvh_node32>(one.data(), 32), b.vertices = upload<float>(tri.data(), 36);
                b.indices = upload<uint32_t>(idx.data(), 12), b.prim_indices = upload<uint64_t>(prim.data(), 8), b.num_prim_indices = 1;
            } else {
                const auto prim32 = slurp(p + "_prim_ids.u32"), verts = slurp(p + "_vertices.f32"), idx = slurp(p + "_indices.u32");
                std::vector<uint64_t> prim64(prim32.size() / 4); // CudaBVH2::prim_indices is size_t*
                for (size_t i = 0; i < prim64.size(); ++i)
                    prim64[i] = reinterpret_cast<const uint32_t *>(prim32.data())[i];
                b.num_nodes = (uint32_t)(nodes.size() / 32), b.nodes = upload<nbvh_node32>(nodes.data(), nodes.size());
                b.prim_indices = upload<uint64_t>(prim64.data(), prim64.size() * 8), b.num_prim_indices = (uint32_t)prim64.size();
                b.vertices = upload<float>(verts.data(), verts.size()), b.indices = upload<uint32_t>(idx.data(), idx.size());
            }
            meshes.push_back(b);
        }
        nbvh_scene_data scene{};
        scene.mesh_bvhs = upload<nbvh_cuda_blas>(meshes.data(), meshes.size() * sizeof(nbvh_cuda_blas)), scene.num_meshes = (uint32_t)meshes.size();
        if (!tlas.empty() && !tlas_idx.empty()) {
            scene.tlas_nodes = upload<nbvh_node32>(tlas.data(), tlas.size()), scene.num_tlas_nodes = (uint32_t)(tlas.size() / 32);
            scene.blas_indices = upload<uint32_t>(tlas_idx.data(), tlas_idx.size());
        } else { // nbvh_build_tlas over the meshes' root boxes; mesh 0 is the mesh the N-BVH stands for (manifest.json)
            const uint32_t n = (uint32_t)meshes.size();
            nbvh_node32 *nodes_dev = nullptr;
            cudaMalloc(reinterpret_cast<void **>(&nodes_dev), (size_t)(2 * n - 1) * sizeof(nbvh_node32));
            uint32_t *idx_dev = nullptr;
            cudaMalloc(reinterpret_cast<void **>(&idx_dev), n * sizeof(uint32_t));
            mod.build_tlas(scene.mesh_bvhs, n, {0u}, nodes_dev, idx_dev);
            cudaDeviceSynchronize();
            scene.tlas_nodes = nodes_dev, scene.num_tlas_nodes = 2 * n - 1, scene.blas_indices = idx_dev;
        }
        mod.set_scene(scene);
        uint32_t W = 1920, H = 1080, spp = 1;
        if (argc >= 6)
            W = (uint32_t)atoi(argv[4]), H = (uint32_t)atoi(argv[5]);
        if (argc >= 7)
            spp = (uint32_t)std::max(1, atoi(argv[6]));
        mod.set_max_depth(argc >= 8 ? atoi(argv[7]) : 1);
        auto sc        = mod.get_constants<nbvh_scene_constants>(NBVH_CONST_SCENE);
        sc.envmap_eval = true, sc.skylike_background = true; // no HDRI in this tool: the reference's sky-like background
        mod.set_constants(NBVH_CONST_SCENE, sc);
        mod.reinitialise_inference_data(W, H);
        std::vector<uint8_t> ldr((size_t)W * H * 4);
        mod.neural_render_host(0, 0, ldr.data()); // warm-up (sample 0 again below: accumulated_spp 0 restarts the accumulation)
        const auto t0 = std::chrono::steady_clock::now();
        for (uint32_t s = 0; s + 1 < spp; ++s)
            mod.neural_render(s, 0);
        mod.neural_render_host(spp - 1, 0, ldr.data()); // last sample + LDR resolve + copy to the host
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (out.size() > 4 && out.compare(out.size() - 4, 4, ".exr") == 0) {
            mod.save_image(out, spp); // == NeuralBVHRenderer::save_image: accumulation / spp as an A,B,G,R float EXR
        } else {
            FILE *f = fopen(out.c_str(), "wb");
            if (!f)
                throw std::runtime_error("cannot write " + out);
            fprintf(f, "P6\n%u %u\n255\n", W, H);
            for (size_t i = 0; i < (size_t)W * H; ++i)
                fwrite(&ldr[4 * i], 1, 3, f);
            fclose(f);
        }
        const auto c = mod.counters();
        printf("{\"module\": \"%s\", \"W\": %u, \"H\": %u, \"spp\": %u, \"max_depth\": %d, \"tree_cut_nodes\": %u, \"n_params\": %llu, "
               "\"rays\": %llu, \"queries\": %llu, \"ms_all_samples\": %.3f}\n",
               mod.name().c_str(), W, H, spp, mod.max_depth(), mod.neural_bvh_num_nodes(), (unsigned long long)mod.n_params(),
               (unsigned long long)c[NBVH_CNT_RAYS], (unsigned long long)c[NBVH_CNT_QUERIES], ms);
        return 0;
    } catch (const std::exception &e) {
        fprintf(stderr, "nbvh_render_cli: %s\n", e.what());
        return 1;
    }
}

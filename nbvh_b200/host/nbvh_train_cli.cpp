// nbvh_train_cli.cpp -- a C++ host of libnbvh_b200.so for the ONLINE OPTIMISATION path: the reference's validation mode "Full"
// (NeuralBVHRenderer::render + process_validation_config_end, /root/reference/src/cuda/neural/neural_bvh_renderer.cu:113-312,
// neural_bvh_renderer_init.cu:260-383) without the window: from-scratch optimisation with tree-cut learning, then `spp` frames, and
// the reference's result files (results/*.csv, results/*.exr, nbvhs/*.nbvh, network_models/*_network.json, configs/*.json).
// Same flow as nbvh_b200/validation.py, written against include/nbvh_b200.hpp.
//
// usage: nbvh_train_cli <scene_dir> <base_config_name> <out_dir> [steps] [spp] [W H]
#include "nbvh_b200.hpp"

#include "../csrc/json_min.hpp" // the library's own minimal JSON reader (header-only), for the renderer-owned keys of the config

#include <cuda_runtime_api.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sys/stat.h>

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

template <typename T>
static T *upload(const void *host, size_t bytes)
{
    void *d = nullptr;
    if (cudaMalloc(&d, bytes ? bytes : 16) != cudaSuccess || (bytes && cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice) != cudaSuccess))
        throw std::runtime_error("device upload failed");
    return reinterpret_cast<T *>(d);
}

template <typename T>
static T *device_alloc(size_t count)
{
    void *d = nullptr;
    if (cudaMalloc(&d, count * sizeof(T)) != cudaSuccess || cudaMemset(d, 0, count * sizeof(T)) != cudaSuccess)
        throw std::runtime_error("device allocation failed");
    return reinterpret_cast<T *>(d);
}

// == BvhSplitScheduler::get_bvh_splits (include/cuda/neural/utils/bvh_split_scheduler.h:73-95); in the reference application the
// renderer keeps using its own scheduler class unchanged
struct SplitScheduler {
    int splitting_interval = 50, last_split_iteration = 2000, start_split_num = 1;
    float split_scaling = 2.f, split_interval_scaling = 2.f;
    int current_num_splits = 1, current_splitting_interval = 50;
    int get_bvh_splits(int it)
    {
        if (it == 0)
            current_num_splits = start_split_num, current_splitting_interval = splitting_interval;
        if (it > last_split_iteration)
            return 0;
        int n = 0;
        if (it % splitting_interval == 0)
            n = current_num_splits;
        if (it % current_splitting_interval == 0) {
            current_num_splits *= split_scaling; // int *= float truncates, like the reference
            current_splitting_interval *= split_interval_scaling;
        }
        return n;
    }
};

int main(int argc, char **argv)
{
    if (argc < 4) {
        fprintf(stderr, "usage: nbvh_train_cli <scene_dir> <base_config_name> <out_dir> [steps] [spp] [W H]\n");
        return 2;
    }
    try {
        const std::string dir = argv[1], base = argv[2], out = argv[3], variant = "cpp";
        const auto cfg_bytes = slurp(dir + "/" + base + "_neural.json");
        const std::string cfg_text(cfg_bytes.begin(), cfg_bytes.end());
        const nbvh::Json cfg  = nbvh::json_parse(cfg_text);
        const nbvh::Json &rend = cfg.object("neural_bvh_renderer");
        const uint32_t steps = argc >= 5 ? (uint32_t)atoi(argv[4]) : (uint32_t)rend.value("max_training_steps", 10000.0);
        const uint32_t spp   = argc >= 6 ? (uint32_t)std::max(1, atoi(argv[5])) : (uint32_t)rend.value("max_accumulated_spp", 64.0);
        const uint32_t B     = (uint32_t)rend.value("batch_size", 262144.0);
        SplitScheduler sched;
        {
            const nbvh::Json &j = rend.object("bvh_split_scheduler");
            sched.splitting_interval     = (int)j.value("splitting_interval", 50.0);
            sched.last_split_iteration   = (int)j.value("last_split_iteration", 2000.0);
            sched.start_split_num        = (int)j.value("start_split_num", 1.0);
            sched.split_scaling          = (float)j.value("split_scaling", 2.0);
            sched.split_interval_scaling = (float)j.value("split_interval_scaling", 2.0);
        }
        const bool tree_cut_learning = rend.value("tree_cut_learning", true);

        nbvh_b200::NeuralPathModule mod(0, cfg_text);
        uint32_t W = 0, H = 0;
        if (argc >= 8) { // another frame size: framebuffer first, then the camera again (its aspect ratio follows the framebuffer)
            W = (uint32_t)atoi(argv[6]), H = (uint32_t)atoi(argv[7]);
            mod.reinitialise_inference_data(W, H);
            const nbvh::Json &cam = cfg.object("cuda_scene").object("camera");
            if (const nbvh::Json *m = cam.find("matrix")) {
                float mtx[16];
                for (int i = 0; i < 16; i++)
                    mtx[i] = (float)m->at_num(i);
                mod.set_camera(mtx, (float)cam.value("yfov", 0.7853981633974483));
            }
        }
        // scene with the neural mesh's own BLAS and the shading attributes the sample generator reads
        const auto tlas = slurp(dir + "/tlas_nodes.bin"), tlas_idx = slurp(dir + "/tlas_indices.u32");
        std::vector<nbvh_cuda_blas> meshes;
        for (int m = 0;; ++m) {
            const std::string p = dir + "/mesh" + std::to_string(m);
            const auto nodes = slurp(p + "_bvh_nodes.bin", false);
            if (nodes.empty())
                break;
            const auto prim32 = slurp(p + "_prim_ids.u32"), verts = slurp(p + "_vertices.f32"), idx = slurp(p + "_indices.u32");
            const auto normals = slurp(p + "_normals.f32"), uvs = slurp(p + "_texcoords.f32");
            const auto mat_ids = slurp(p + "_material_ids.i32"), mat_map = slurp(p + "_material_map.u32");
            std::vector<uint64_t> prim64(prim32.size() / 4); // CudaBVH2::prim_indices is size_t*
            for (size_t i = 0; i < prim64.size(); ++i)
                prim64[i] = reinterpret_cast<const uint32_t *>(prim32.data())[i];
            nbvh_cuda_blas b{};
            b.num_nodes = (uint32_t)(nodes.size() / 32), b.nodes = upload<nbvh_node32>(nodes.data(), nodes.size());
            b.prim_indices = upload<uint64_t>(prim64.data(), prim64.size() * 8), b.num_prim_indices = (uint32_t)prim64.size();
            b.vertices = upload<float>(verts.data(), verts.size()), b.indices = upload<uint32_t>(idx.data(), idx.size());
            b.normals = upload<float>(normals.data(), normals.size()), b.texcoords = upload<float>(uvs.data(), uvs.size());
            b.material_ids = upload<int32_t>(mat_ids.data(), mat_ids.size());
            b.material_map = upload<uint32_t>(mat_map.data(), mat_map.size()), b.num_materials = (uint32_t)(mat_map.size() / 4);
            meshes.push_back(b);
        }
        const auto mats = slurp(dir + "/materials.bin");
        std::vector<uint64_t> tex_handles;
        for (int k = 0;; ++k) {
            const auto px = slurp(dir + "/tex" + std::to_string(k) + ".rgba8", false);
            if (px.empty())
                break;
            const auto dims = slurp(dir + "/tex" + std::to_string(k) + ".dims");
            const uint32_t *wh = reinterpret_cast<const uint32_t *>(dims.data());
            uint64_t h = 0;
            if (nbvh_create_texture_rgba8(mod.handle(), px.data(), wh[0], wh[1], &h) != NBVH_OK)
                throw std::runtime_error(std::string("texture: ") + nbvh_last_error(mod.handle()));
            tex_handles.push_back(h);
        }
        nbvh_scene_data scene{};
        scene.tlas_nodes = upload<nbvh_node32>(tlas.data(), tlas.size()), scene.num_tlas_nodes = (uint32_t)(tlas.size() / 32);
        scene.blas_indices = upload<uint32_t>(tlas_idx.data(), tlas_idx.size());
        scene.mesh_bvhs = upload<nbvh_cuda_blas>(meshes.data(), meshes.size() * sizeof(nbvh_cuda_blas)), scene.num_meshes = (uint32_t)meshes.size();
        scene.materials = upload<uint8_t>(mats.data(), mats.size());
        scene.textures  = upload<uint64_t>(tex_handles.data(), tex_handles.size() * 8);
        mod.set_scene(scene);
        mod.set_max_depth(1);

        for (const char *d : {"", "/results", "/nbvhs", "/network_models", "/configs"})
            mkdir((out + d).c_str(), 0755);
        auto result_path = [&](int kind, uint32_t n) {
            char buf[4096];
            if (nbvh_result_path(out.c_str(), base.c_str(), variant.c_str(), kind, n, buf, sizeof(buf)) != NBVH_OK)
                throw std::runtime_error("result path too long");
            return std::string(buf);
        };

        // ---- optimisation: reinitialise_neural_bvh_learning_states + one step per frame (neural_bvh_renderer.cu:140-290)
        mod.init_params(1337);
        mod.tree_cut_init((uint32_t)rend.value("neural_bvh_min_depth", 0.0));
        float *inputs = device_alloc<float>((size_t)B * 9), *targets = device_alloc<float>((size_t)B * 8), *L = device_alloc<float>((size_t)B * 16);
        uint32_t *sampled = device_alloc<uint32_t>(B);
        nbvh_b200::FrameTimer timer;
        std::vector<float> losses;
        float max_node_loss = 0.f;
        for (uint32_t s = 0; s < steps; ++s) {
            timer.start_frame(s);
            mod.optimisation_data_generation(s, B, 0, sampled, s == 0 ? 0.f : max_node_loss, mod.tree_cut_data(), inputs, targets);
            losses.push_back(mod.training_step(inputs, targets, B, L));
            max_node_loss = mod.tree_cut_step(L, sampled, B, tree_cut_learning ? (uint32_t)sched.get_bvh_splits((int)s) : 0u);
            timer.end_frame();
        }
        timer.write_to_file(result_path(1, 0), &losses);
        mod.save_neural_bvh(out + "/nbvhs/" + base + ".nbvh");
        mod.save_network_model(out + "/network_models/" + base + "_network.json");

        // ---- inference: one neural_render per frame
        nbvh_b200::FrameTimer itimer;
        for (uint32_t s = 0; s < spp; ++s) {
            itimer.start_frame(s);
            mod.neural_render(s, 0);
            cudaDeviceSynchronize();
            itimer.end_frame();
        }
        mod.save_image(result_path(0, spp), spp);
        itimer.write_to_file(result_path(2, spp));
        std::ofstream(out + "/configs/" + base + "_" + variant + ".json") << mod.save_config() << "\n";

        double last = 0.0;
        const size_t tail = std::min<size_t>(10, losses.size());
        for (size_t i = losses.size() - tail; i < losses.size(); ++i)
            last += losses[i] / (double)tail;
        printf("{\"steps\": %u, \"spp\": %u, \"batch\": %u, \"tree_cut_nodes\": %u, \"loss_first\": %.9g, \"loss_last_mean_10\": %.9g, "
               "\"image\": \"%s\"}\n",
               steps, spp, B, mod.neural_bvh_num_nodes(), losses.empty() ? 0.0 : (double)losses[0], last, result_path(0, spp).c_str());
        return 0;
    } catch (const std::exception &e) {
        fprintf(stderr, "nbvh_train_cli: %s\n", e.what());
        return 1;
    }
}

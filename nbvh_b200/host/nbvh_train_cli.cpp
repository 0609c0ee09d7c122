// nbvh_train_cli.cpp -- a C++ host of libnbvh_b200.so for the ONLINE OPTIMISATION path: the reference's validation mode "Full"
// (NeuralBVHRenderer::render + process_validation_config_end, /root/reference/src/cuda/neural/neural_bvh_renderer.cu:113-312,
// neural_bvh_renderer_init.cu:260-383) without the window: from-scratch optimisation with tree-cut learning, then `spp` frames, and
// the reference's result files (results/*.csv, results/*.exr, nbvhs/*.nbvh, network_models/*_network.json, configs/*.json).
// Same flow as nbvh_b200/validation.py, written against include/nbvh_b200.hpp.
//
// usage: nbvh_train_cli <scene_dir> <base_config_name> <out_dir> [steps] [spp] [W H]
//
// Multi-GPU (one process per GPU, e.g. `python -m torch.distributed.run --no-python --nproc-per-node N nbvh_train_cli ...`, or any
// launcher that sets RANK / WORLD_SIZE / LOCAL_RANK): data-parallel optimisation inside the library -- the config's batch_size is the
// GLOBAL batch, every rank samples and trains batch_size / WORLD_SIZE slots, gradients, loss and tree-cut statistics are summed by
// NCCL (nbvh_allreduce_*), so all ranks hold the same weights and tree-cut; rank 0 renders and writes the result files.  The NCCL
// unique id travels through the file NBVH_NCCL_ID_FILE (default /tmp/nbvh_nccl_id_<MASTER_PORT>).
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
    int current_num_splits = 1, current_splitting_interval = 50, current_total_splits = 0;
    int total_splits = 0; // BvhSplitStats::total_splits of update_stats_and_reset
    void reset() { current_num_splits = start_split_num, current_splitting_interval = splitting_interval, current_total_splits = 0; }
    // == update_stats_and_reset (:38-71): the splits the schedule will hand out, clamped by the leaves available at each cut
    void update_stats_and_reset(int bvh_start_num_nodes, int num_iterations)
    {
        reset();
        total_splits = 0;
        int expected_nodes = bvh_start_num_nodes;
        for (int i = 0; i < num_iterations; i++) {
            const int leaves = (expected_nodes + 1) / 2;
            int s            = get_bvh_splits(i);
            if (s > 0) {
                s = std::min(s, leaves);
                total_splits += s, expected_nodes += 2 * s;
            }
        }
        reset();
    }
    // == should_assign_new_lod (:97-125), EqualIterations (the reference's default LoD decision type)
    bool should_assign_new_lod(int it, int current_learned_lod, int num_lods_learned, int max_bvh_lods = 7) const
    {
        const int next = (max_bvh_lods - current_learned_lod) * last_split_iteration / num_lods_learned;
        return next <= it || it >= last_split_iteration;
    }
    int get_bvh_splits(int it)
    {
        if (it == 0)
            reset();
        if (it > last_split_iteration)
            return 0;
        int n = 0;
        if (it % splitting_interval == 0) {
            n = current_num_splits;
            current_total_splits += n;
        }
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
        bool tree_cut_learning = rend.value("tree_cut_learning", true);
        auto env_int = [](const char *name, int fallback) {
            const char *v = getenv(name);
            return v && *v ? atoi(v) : fallback;
        };
        const int rank = env_int("RANK", 0), world = std::max(1, env_int("WORLD_SIZE", 1)), local_rank = env_int("LOCAL_RANK", rank);
        if (B % (uint32_t)world != 0 || ((B / (uint32_t)world) & (B / (uint32_t)world - 1u)) != 0)
            throw std::runtime_error("batch_size / WORLD_SIZE must be a power of two");
        const uint32_t B_local = B / (uint32_t)world;

        nbvh_b200::NeuralPathModule mod(local_rank, cfg_text);
        cudaSetDevice(local_rank);
        if (world > 1) {
            const char *idf        = getenv("NBVH_NCCL_ID_FILE");
            const std::string path = idf && *idf ? idf : "/tmp/nbvh_nccl_id_" + std::to_string(env_int("MASTER_PORT", 0));
            mod.setup_data_parallel(nbvh_b200::exchange_unique_id_via_file(path, rank), rank, world);
            if (rank == 0)
                std::remove(path.c_str()); // every rank has joined once the communicator exists
        }
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
        // rows are strided by the config's dims (3 * input_segment_points inputs, output_network_dims targets)
        const uint32_t in_dims  = mod.get_constants<nbvh_module_constants>(NBVH_CONST_MODULE).input_network_dims;
        const uint32_t out_dims = mod.network_output_dims();
        float *inputs = device_alloc<float>((size_t)B_local * in_dims), *targets = device_alloc<float>((size_t)B_local * out_dims),
              *L = device_alloc<float>((size_t)B_local * 16);
        uint32_t *sampled = device_alloc<uint32_t>(B_local);
        const nbvh_renderer_constants bvh_c = mod.get_constants<nbvh_renderer_constants>(NBVH_CONST_RENDERER);
        const uint32_t log2_trav            = (uint32_t)rend.object("bvh_constants").value("log2_bvh_max_traversals", 0.0);
        const int num_lods_learned          = bvh_c.num_lods_learned;
        sched.update_stats_and_reset((int)mod.neural_bvh_num_nodes(), (int)rend.value("max_training_steps", (double)steps));
        // data parallel: node-indexed scratch for the statistics allreduce (the schedule bounds the node count) and the loss scalar
        const size_t node_cap   = (size_t)mod.neural_bvh_num_nodes() + 2u * (size_t)sched.total_splits + 2u;
        float *loss_scratch     = world > 1 ? device_alloc<float>(node_cap) : nullptr;
        uint32_t *samp_scratch  = world > 1 ? device_alloc<uint32_t>(node_cap) : nullptr;
        float *loss_sum_dev     = world > 1 ? device_alloc<float>(1) : nullptr;
        nbvh_b200::FrameTimer timer;
        std::vector<float> losses;
        float max_node_loss = 0.f;
        bool all_lods_assigned = false;
        for (uint32_t s = 0; s < steps; ++s) {
            timer.start_frame(s);
            // walk-ahead of the sample generator: up to 16 steps per walk launch, not past the next scheduled split
            uint32_t ahead = 16;
            if (tree_cut_learning && (int)s <= sched.last_split_iteration) {
                const uint32_t iv = (uint32_t)std::max(1, sched.splitting_interval);
                ahead = std::min(ahead, ((s + iv - 1u) / iv) * iv - s + 1u);
            }
            mod.sampler_walk_ahead(ahead, (uint32_t)world);
            // the sampler's RNG is keyed by training_step * batch + slot: step * world + rank gives every rank its own global slots
            mod.optimisation_data_generation(s * (uint32_t)world + (uint32_t)rank, B_local, log2_trav, sampled, s == 0 ? 0.f : max_node_loss,
                                             mod.tree_cut_data(), inputs, targets);
            const uint32_t n_splits = tree_cut_learning ? (uint32_t)sched.get_bvh_splits((int)s) : 0u;
            if (world == 1) {
                losses.push_back(mod.training_step(inputs, targets, B_local, L));
                max_node_loss = mod.tree_cut_step(L, sampled, B_local, n_splits);
            } else {
                mod.training_forward_backward(inputs, targets, B_local, B, L, loss_sum_dev);
                mod.optimizer_step();
                float loss = 0.f;
                cudaMemcpy(&loss, loss_sum_dev, sizeof(float), cudaMemcpyDeviceToHost);
                losses.push_back(loss);
                max_node_loss = mod.tree_cut_step_data_parallel(L, sampled, B_local, n_splits, loss_scratch, samp_scratch);
            }
            // assign_lod_to_current_neural_leaf_nodes when the schedule says so (neural_bvh_renderer.cu:262-282)
            if (tree_cut_learning && num_lods_learned > 1 && !all_lods_assigned) {
                const int lod = mod.get_constants<nbvh_renderer_constants>(NBVH_CONST_RENDERER).current_learned_lod;
                if (sched.should_assign_new_lod((int)s, lod, num_lods_learned)) {
                    all_lods_assigned = mod.tree_cut_assign_lod();
                    if (all_lods_assigned)
                        tree_cut_learning = false;
                }
            }
            timer.end_frame();
        }
        if (rank != 0) { // replicas hold the same weights and tree-cut: rank 0 renders and writes
            printf("{\"rank\": %d, \"steps\": %u, \"tree_cut_nodes\": %u, \"loss_first\": %.9g}\n", rank, steps, mod.neural_bvh_num_nodes(),
                   losses.empty() ? 0.0 : (double)losses[0]);
            return 0;
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
        printf("{\"world\": %d, \"steps\": %u, \"spp\": %u, \"batch\": %u, \"tree_cut_nodes\": %u, \"loss_first\": %.9g, \"loss_last_mean_10\": %.9g, "
               "\"image\": \"%s\"}\n",
               world, steps, spp, B, mod.neural_bvh_num_nodes(), losses.empty() ? 0.0 : (double)losses[0], last, result_path(0, spp).c_str());
        return 0;
    } catch (const std::exception &e) {
        fprintf(stderr, "nbvh_train_cli: %s\n", e.what());
        return 1;
    }
}

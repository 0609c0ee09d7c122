// nbvh_b200.hpp -- C++ host side above the C ABI (include/nbvh_b200.h): a header-only RAII class with the method names, argument
// meaning and error behaviour of the reference's NeuralBVHModule / NeuralPathModule
// (/root/reference/include/cuda/neural/neural_bvh_module.h:144-241, /root/reference/include/cuda/neural/path/path_module.h:256-345)
// and of the tcnn objects the renderer reaches through it (trainer->training_step, network->inference, serialize / deserialize).
// Errors become std::runtime_error, like the reference's CUDA_CHECK / tcnn CHECK_THROW (/root/reference/include/ntwr.h:30-54).
// No reference header is needed: the scene, node and constant structs are the layout-identical PODs of nbvh_b200.h.
#pragma once
#include "nbvh_b200.h"

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace nbvh_b200 {

class NeuralPathModule {
public:
    explicit NeuralPathModule(int device = 0, const std::string &config_json = std::string())
    {
        const int rc = nbvh_create(device, config_json.empty() ? nullptr : config_json.c_str(), &m_ctx);
        if (rc != NBVH_OK) {
            const std::string msg = m_ctx ? nbvh_last_error(m_ctx) : "no usable sm_100 CUDA device (nbvh_b200 has no CPU fallback)";
            if (m_ctx)
                nbvh_destroy(m_ctx);
            m_ctx = nullptr;
            throw std::runtime_error("nbvh_create: " + msg);
        }
    }
    ~NeuralPathModule()
    {
        if (m_ctx)
            nbvh_destroy(m_ctx);
    }
    NeuralPathModule(const NeuralPathModule &)            = delete;
    NeuralPathModule &operator=(const NeuralPathModule &) = delete;

    std::string name() const { return "Path Tracing Module (B200)"; } // NeuralPathModule::name

    // ---- configuration: load_config_from_json / setup_neural_module_constants (neural_bvh_renderer_init.cu:93-103,566-749)
    void load_config_from_json(const std::string &json_text) { check(nbvh_load_config_json(m_ctx, json_text.c_str())); }
    template <typename Pod>
    void set_constants(nbvh_constant_block which, const Pod &pod) { check(nbvh_set_constants(m_ctx, which, &pod, sizeof(Pod))); }
    template <typename Pod>
    Pod get_constants(nbvh_constant_block which) const
    {
        Pod pod;
        check(nbvh_get_constants(m_ctx, which, &pod, sizeof(Pod)));
        return pod;
    }
    void set_camera(const float matrix16[16], float yfov_rad) { check(nbvh_set_camera(m_ctx, matrix16, yfov_rad)); }
    void set_max_depth(int max_depth) { check(nbvh_set_max_depth(m_ctx, max_depth)); }
    int max_depth() const { return nbvh_get_max_depth(m_ctx); }
    uint32_t network_output_dims() const { return get_constants<nbvh_pt_output_constants>(NBVH_CONST_PT_OUTPUT).output_network_dims; }

    // ---- scene, tree-cut, weights (CudaSceneData / NeuralBVH are passed as they are)
    void set_scene(const nbvh_scene_data &scene) { check(nbvh_set_scene(m_ctx, &scene)); }
    void set_envmap(uint64_t envmap_texture) { check(nbvh_set_envmap(m_ctx, envmap_texture)); }
    void set_neural_bvh(const nbvh_node32 *nodes_device, uint32_t n) { check(nbvh_set_neural_bvh(m_ctx, nodes_device, n)); }
    void load_neural_bvh(const std::string &path) { check(nbvh_load_neural_bvh_file(m_ctx, path.c_str())); }
    void save_neural_bvh(const std::string &path) { check(nbvh_save_neural_bvh_file(m_ctx, path.c_str())); }
    void load_network_model(const std::string &path) { check(nbvh_load_network_file(m_ctx, path.c_str())); }
    void save_network_model(const std::string &path) { check(nbvh_save_network_file(m_ctx, path.c_str())); }
    void init_params(uint64_t seed = 1337) { check(nbvh_init_params(m_ctx, seed)); } // tcnn's Trainer initialisation (trainer.h:50-57)
    uint64_t n_params() const { return nbvh_n_params(m_ctx); }
    uint32_t neural_bvh_num_nodes() const { return nbvh_neural_bvh_num_nodes(m_ctx); }

    // ---- inference: reinitialise_inference_data / neural_render (neural_bvh_module.h:150,188; path_module.cu:139-245)
    void reinitialise_inference_data(uint32_t fb_w, uint32_t fb_h, float *accum_rgba32f = nullptr, uint8_t *ldr_rgba8 = nullptr)
    {
        check(nbvh_set_framebuffer(m_ctx, accum_rgba32f, ldr_rgba8, fb_w, fb_h));
    }
    void neural_render(uint32_t accumulated_spp, uint32_t sample_offset, void *stream = nullptr)
    {
        check(nbvh_render(m_ctx, accumulated_spp, sample_offset, 0, 0, stream));
    }
    void neural_render_samples(uint32_t first_spp, uint32_t n_spp, uint32_t sample_offset = 0, uint32_t shard = 0, uint32_t n_shards = 1,
                               void *stream = nullptr)
    {
        check(nbvh_render_samples(m_ctx, first_spp, n_spp, sample_offset, shard, n_shards, stream));
    }
    // drain order of the neural ray queue by expected query-chain length (scheduling only; all limits 0 = arrival order)
    void set_queue_order(uint32_t t0 = 4, uint32_t t1 = 3, uint32_t t2 = 2, uint32_t t3 = 2, uint32_t flags = 1)
    {
        check(nbvh_set_queue_order(m_ctx, t0, t1, t2, t3, flags));
    }
    void neural_render_host(uint32_t accumulated_spp, uint32_t sample_offset, uint8_t *ldr_host, const nbvh_camera_params *camera = nullptr)
    {
        check(nbvh_render_host(m_ctx, accumulated_spp, sample_offset, camera, ldr_host, nullptr));
    }
    void inference(const float *in_device, float *out_device, uint32_t n, void *stream = nullptr) // == network->inference
    {
        check(nbvh_query(m_ctx, in_device, out_device, n, stream));
    }
    void trace_rays(const float *o, const float *d, uint32_t n, uint32_t ray_depth, nbvh_hit_record *hits, void *stream = nullptr)
    {
        check(nbvh_trace_rays(m_ctx, o, d, n, ray_depth, hits, stream));
    }

    // ---- online optimisation: optimisation_data_generation / trainer->training_step / tree-cut learning
    void optimisation_data_generation(uint32_t training_step, uint32_t batch_size, uint32_t log2_bvh_max_traversals,
                                      uint32_t *sampled_node_indices, float current_max_loss, const nbvh_tree_cut_data *tree_cut_data,
                                      float *inputs, float *targets, void *stream = nullptr)
    {
        check(nbvh_optimisation_data_generation(m_ctx, training_step, batch_size, log2_bvh_max_traversals, sampled_node_indices,
                                                current_max_loss, tree_cut_data, inputs, targets, stream));
    }
    void sampler_walk_ahead(uint32_t steps, uint32_t step_stride = 1) { check(nbvh_sampler_walk_ahead(m_ctx, steps, step_stride)); }
    float training_step(const float *in, const float *target, uint32_t batch, float *per_sample_loss = nullptr, void *stream = nullptr)
    {
        float loss = 0.f;
        check(nbvh_training_step(m_ctx, in, target, batch, per_sample_loss, &loss, stream));
        return loss;
    }
    // ---- data-parallel optimisation inside the library (nbvh_allreduce_*): one process per GPU, gradients and tree-cut statistics
    //      summed by NCCL; the host's launcher only carries the 128-byte unique id from rank 0 to the other ranks
    static std::vector<uint8_t> nccl_unique_id()
    {
        std::vector<uint8_t> id(NBVH_NCCL_UNIQUE_ID_BYTES);
        if (nbvh_nccl_unique_id(id.data()) != NBVH_OK)
            throw std::runtime_error("nbvh_b200: NCCL is not loadable (libnccl.so.2)");
        return id;
    }
    void setup_data_parallel(const std::vector<uint8_t> &unique_id, int rank, int world)
    {
        if (unique_id.size() != NBVH_NCCL_UNIQUE_ID_BYTES)
            throw std::runtime_error("nbvh_b200: bad NCCL unique id");
        check(nbvh_allreduce_setup_unique_id(m_ctx, unique_id.data(), rank, world));
    }
    void setup_data_parallel(void *nccl_comm) { check(nbvh_allreduce_setup(m_ctx, nccl_comm)); } // the caller's ncclComm_t
    int data_parallel_world(int *rank = nullptr) const { return nbvh_data_parallel_world(m_ctx, rank); }
    void allreduce(void *device_ptr, uint64_t count, int dtype, void *stream = nullptr)
    {
        check(nbvh_allreduce_buffer(m_ctx, device_ptr, count, dtype, stream));
    }
    // this rank's share of the batch, the GLOBAL batch size; on return (stream order) the gradients and *loss_sum_dev are global sums
    void training_forward_backward(const float *in, const float *target, uint32_t batch, uint32_t global_batch, float *per_sample_loss,
                                   float *loss_sum_dev, void *stream = nullptr)
    {
        check(nbvh_training_forward_backward(m_ctx, in, target, batch, global_batch, per_sample_loss, loss_sum_dev, stream));
    }
    void optimizer_step(void *stream = nullptr) { check(nbvh_optimizer_step(m_ctx, stream)); }
    // per-rank statistics -> allreduce -> the same deterministic decisions and splits on every rank; scratch: one float and one
    // uint32 per tree-cut node (capacity: the node count the schedule will reach)
    float tree_cut_step_data_parallel(const float *per_sample_loss, const uint32_t *sampled_node_indices, uint32_t batch, uint32_t num_splits,
                                      float *loss_scratch, uint32_t *samples_scratch, uint32_t *n_split = nullptr, void *stream = nullptr)
    {
        const uint32_t n = nbvh_neural_bvh_num_nodes(m_ctx);
        check(nbvh_tree_cut_accumulate(m_ctx, per_sample_loss, 16, network_output_dims(), sampled_node_indices, batch, loss_scratch,
                                       samples_scratch, stream));
        allreduce(loss_scratch, n, NBVH_DTYPE_F32, stream);
        allreduce(samples_scratch, n, NBVH_DTYPE_U32, stream);
        float max_loss = 0.f;
        check(nbvh_tree_cut_decide(m_ctx, loss_scratch, samples_scratch, num_splits, &max_loss, n_split, stream));
        return max_loss;
    }
    bool tree_cut_assign_lod() // == assign_lod_to_current_neural_leaf_nodes; true when every LoD has been assigned
    {
        int done = 0;
        check(nbvh_tree_cut_assign_lod(m_ctx, &done));
        return done != 0;
    }
    void tree_cut_init(uint32_t min_depth) { check(nbvh_tree_cut_init(m_ctx, min_depth)); }
    const nbvh_tree_cut_data *tree_cut_data() const { return nbvh_tree_cut_data_device(m_ctx); }
    float tree_cut_step(const float *per_sample_loss, const uint32_t *sampled_node_indices, uint32_t batch, uint32_t num_splits,
                        uint32_t *n_split = nullptr, void *stream = nullptr)
    {
        float max_loss = 0.f;
        check(nbvh_tree_cut_step(m_ctx, per_sample_loss, 16, network_output_dims(), sampled_node_indices, batch, num_splits, &max_loss, n_split,
                                 stream));
        return max_loss;
    }

    // ---- validation metrics: compute_and_display_render_error / compute_and_display_flip_error (neural_bvh_renderer.cu:34-96,346-358)
    float image_error(const float *reference_rgba32f, uint32_t accumulated_spp, int postprocess_mode, bool write_error_image = false,
                      void *stream = nullptr)
    {
        float mean = 0.f;
        check(nbvh_image_error(m_ctx, reference_rgba32f, accumulated_spp, postprocess_mode, write_error_image ? 1 : 0, &mean, stream));
        return mean;
    }
    float flip_error(const float *reference_rgba32f, uint32_t accumulated_spp, float *error_map_device = nullptr, void *stream = nullptr)
    {
        float mean = 0.f;
        check(nbvh_flip_error(m_ctx, reference_rgba32f, accumulated_spp, 0.f, error_map_device, &mean, stream));
        return mean;
    }
    // ---- scene build on the device: CudaBackend's BLAS / TLAS builds (cuda_backend.cpp:186-385)
    uint32_t build_blas(const float *vertices, const uint32_t *indices, uint32_t n_tris, nbvh_node32 *nodes_out, uint64_t *prim_indices_out,
                        void *stream = nullptr)
    {
        uint32_t depth = 0;
        check(nbvh_build_blas(m_ctx, vertices, indices, n_tris, nodes_out, prim_indices_out, &depth, stream));
        return depth;
    }
    void build_tlas(const nbvh_cuda_blas *mesh_bvhs_device, uint32_t num_meshes, const std::vector<uint32_t> &neural_meshes,
                    nbvh_node32 *tlas_nodes_out, uint32_t *blas_indices_out, void *stream = nullptr)
    {
        check(nbvh_build_tlas(m_ctx, mesh_bvhs_device, num_meshes, neural_meshes.data(), (uint32_t)neural_meshes.size(), tlas_nodes_out,
                              blas_indices_out, stream));
    }

    // ---- results (f4): save_image / save_config / save_performance_metrics (neural_bvh_renderer_init.cu:328-383,406-509)
    void save_image(const std::string &path, uint32_t accumulated_spp, void *stream = nullptr)
    {
        check(nbvh_save_image_exr(m_ctx, path.c_str(), accumulated_spp, stream));
    }
    std::string save_config() const
    {
        size_t needed = 0;
        nbvh_save_config_json(m_ctx, nullptr, 0, &needed);
        std::string text(needed, '\0');
        check(nbvh_save_config_json(m_ctx, text.data(), needed, nullptr));
        text.resize(needed - 1);
        return text;
    }

    std::vector<uint64_t> counters() const
    {
        std::vector<uint64_t> c(NBVH_CNT_COUNT);
        check(nbvh_get_counters(m_ctx, c.data(), NBVH_CNT_COUNT));
        return c;
    }
    nbvh_ctx *handle() const { return m_ctx; }

private:
    void check(int rc) const
    {
        if (rc != NBVH_OK)
            throw std::runtime_error(std::string("nbvh_b200: ") + nbvh_last_error(m_ctx));
    }
    nbvh_ctx *m_ctx = nullptr;
};

// Carries the NCCL unique id from rank 0 to the other ranks of a one-process-per-GPU launch through a file (any launcher that sets
// RANK / WORLD_SIZE works: `python -m torch.distributed.run --no-python`, mpirun with a wrapper, a shell loop).  Rank 0 writes
// `path` atomically (rename), the others poll for it.
inline std::vector<uint8_t> exchange_unique_id_via_file(const std::string &path, int rank, int timeout_s = 120)
{
    if (rank == 0) {
        const std::vector<uint8_t> id = NeuralPathModule::nccl_unique_id();
        const std::string tmp         = path + ".tmp";
        std::ofstream(tmp, std::ios::binary).write(reinterpret_cast<const char *>(id.data()), (std::streamsize)id.size());
        if (std::rename(tmp.c_str(), path.c_str()) != 0)
            throw std::runtime_error("nbvh_b200: cannot publish the NCCL unique id at " + path);
        return id;
    }
    for (int waited_ms = 0; waited_ms < timeout_s * 1000; waited_ms += 20) {
        std::ifstream f(path, std::ios::binary | std::ios::ate);
        if (f && (size_t)f.tellg() == NBVH_NCCL_UNIQUE_ID_BYTES) {
            std::vector<uint8_t> id(NBVH_NCCL_UNIQUE_ID_BYTES);
            f.seekg(0);
            f.read(reinterpret_cast<char *>(id.data()), (std::streamsize)id.size());
            return id;
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(20));
    }
    throw std::runtime_error("nbvh_b200: timed out waiting for the NCCL unique id at " + path);
}

// == FrameTimer (include/utils/frame_timer.h): start_frame / end_frame / write_to_file with the reference's CSV layout
class FrameTimer {
public:
    void start_frame(uint32_t frame_idx)
    {
        if (frame_idx == 0)
            m_frame_times.clear();
        m_start = std::chrono::high_resolution_clock::now();
    }
    void end_frame()
    {
        const auto now = std::chrono::high_resolution_clock::now();
        m_frame_times.push_back(std::chrono::duration<double, std::milli>(now - m_start).count());
        m_start = now;
    }
    bool write_to_file(const std::string &filename, const std::vector<float> *losses = nullptr) const
    {
        if (losses && losses->size() < m_frame_times.size())
            return false;
        return nbvh_write_frame_times_csv_host(filename.c_str(), m_frame_times.data(), losses ? losses->data() : nullptr,
                                               (uint32_t)m_frame_times.size()) == NBVH_OK;
    }
    const std::vector<double> &frame_times() const { return m_frame_times; }

private:
    std::chrono::time_point<std::chrono::high_resolution_clock> m_start;
    std::vector<double> m_frame_times;
};

} // namespace nbvh_b200

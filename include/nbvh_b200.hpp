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
#include <stdexcept>
#include <string>
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
    float training_step(const float *in, const float *target, uint32_t batch, float *per_sample_loss = nullptr, void *stream = nullptr)
    {
        float loss = 0.f;
        check(nbvh_training_step(m_ctx, in, target, batch, per_sample_loss, &loss, stream));
        return loss;
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

// context.cu -- host side of libnbvh_b200.so: context object, the reference's on-disk formats
// (config JSON, .nbvh, network_models JSON), constant blocks, parameter layout, and the C ABI glue.
// Mirrors what NeuralBVHRenderer / NeuralBVHModule do on the host around the hot path
// (/root/reference/src/cuda/neural/neural_bvh_renderer_init.cu:385-404,566-749,
//  /root/reference/src/cuda/neural/neural_bvh_renderer_accel.cpp:170-254,
//  /root/reference/src/cuda/neural/neural_bvh_module.cu:12-263).
#include "json_min.hpp"
#include "kernels.cuh"
#include "sampler.cuh"
#include "path.cuh"
#include "treecut.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

using namespace nbvh;

#define CU_TRY(ctx, expr)                                                                                   \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess) {                                                                            \
            (ctx)->error = std::string(#expr) + ": " + cudaGetErrorString(_e);                              \
            return NBVH_ERR_CUDA;                                                                           \
        }                                                                                                   \
    } while (0)

namespace {

const char *const kLossTypes[] = {"RelativeL2", "L2", "RelativeL1", "L1", "RelativeL2Luminance", "BinaryCrossEntropy",
                                  "LogL2", "LHalf"};
const char *const kDepthModes[]           = {"LocalT", "WorldT", "Position", "LocalPosition"};
const char *const kSegmentDistributions[] = {"EquidistantInterior", "EquidistantExterior", "Uniform"};
const char *const kOutputModes[]          = {"NeuralPathTracing"};

template <size_t N>
int find_index(const std::string &s, const char *const (&table)[N])
{
    for (size_t i = 0; i < N; i++)
        if (s == table[i])
            return (int)i;
    return -1;
}

// ---- defaults of the reference structs ---------------------------------------------------------
nbvh_renderer_constants default_renderer_constants()
{
    nbvh_renderer_constants c;
    memset(&c, 0, sizeof(c));
    c.neural_bvh_idx = 0, c.num_lods_learned = 1, c.current_learned_lod = 0, c.rendered_lod = 0;
    c.ignore_neural_blas = false, c.path_lod_switch = false, c.lod_switch = 0;
    c.leaf_boundary_sampling_ratio = 0.f;
    c.inference_mint_epsilon = 1e-4f, c.inference_maxt_epsilon = 1e-3f;
    c.learning_mint_epsilon = 1e-3f, c.learning_maxt_epsilon = 1e-3f;
    c.sampling_bbox_scaling = 1.5f, c.fixed_node_dilation = 1e-5f, c.node_extent_dilation = 1e-4f;
    c.loss_exposure = 2e-7f, c.show_loss_heuristic = false, c.ellipsoidal_bvh = false, c.rescale_ellipsoid = false;
    c.log2_bvh_max_traversals = 0, c.sample_cosine_weighted = true, c.backface_bounce_training = false;
    c.max_training_bounces = 1, c.training_node_survival = true, c.survival_probability_scale = 1.f;
    c.min_rr_survival_prob = 0.005f, c.view_dependent_sampling = false, c.view_dependent_ratio = 0.f;
    return c;
}

nbvh_module_constants default_module_constants()
{
    nbvh_module_constants c;
    memset(&c, 0, sizeof(c));
    c.input_segment_point_dim = 3, c.input_segment_points = 3, c.input_segment_points_dim = 9;
    c.input_segment_direction_dim = 0, c.input_network_dims = 9;
    c.segment_distribution = 0; // EquidistantInterior
    return c;
}

nbvh_pt_output_constants default_pt_constants()
{
    nbvh_pt_output_constants c;
    memset(&c, 0, sizeof(c));
    c.output_visibility_dim = 1, c.output_depth_dim = 1, c.output_normal_dim = 0, c.output_albedo_dim = 0;
    c.output_network_dims = 2;
    c.visibility_loss = 5, c.visibility_loss_scale = 1.f, c.depth_loss = 3, c.depth_loss_scale = 1.f;
    c.normal_loss = 3, c.normal_loss_scale = 1.f, c.albedo_loss = 0, c.albedo_loss_scale = 1.f;
    c.depth_learning_mode = 0;
    c.output_normal_linear_activation = true, c.output_depth_linear_activation = false;
    c.output_albedo_linear_activation = false, c.visibility_masking_loss = true;
    c.primary_visibility_threshold = 0.5f, c.secondary_visibility_threshold = 0.5f;
    c.depth_epsilon = 0.2f, c.fixed_depth_epsilon = 1e-5f, c.interpolate_depth = false, c.depth_interpolation_eps = 1e-3f;
    return c;
}

nbvh_path_constants default_path_constants()
{
    nbvh_path_constants c;
    memset(&c, 0, sizeof(c));
    c.output_mode = 0, c.normal_epsilon_offset = 1e-4f, c.apply_russian_roulette = true, c.rr_start_depth = 2;
    c.neural_roughness = 1.f, c.neural_specular = 0.f, c.neural_metallic = 0.f;
    return c;
}

nbvh_scene_constants default_scene_constants()
{
    nbvh_scene_constants c;
    memset(&c, 0, sizeof(c));
    c.constant_roughness = 1.f, c.constant_specular = 0.f, c.constant_metallic = 0.f;
    c.envmap_strength = 1.f, c.envmap_rotation = 0.f, c.envmap_eval = true, c.emission_eval = true;
    c.skylike_background = false, c.sample_light = false, c.ema_weight = 0.05f, c.ema_accumulate = true;
    c.force_constant_material = false;
    return c;
}

// ---- camera (host float arithmetic, no contraction: compiled with -fmad=false, and x86 baseline has no FMA)
// glm::inverse cofactor expansion + Camera::origin/dir/up + CudaBackend::update_camera
// (/root/reference/src/base/camera/camera.cpp:39-82, /root/reference/src/cuda/cuda_backend.cpp:595-611)
void invert4(const float *a, float *out)
{
    auto M = [&](int c, int r) { return a[c * 4 + r]; };
    const float c00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3), c02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3);
    const float c03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3), c04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
    const float c06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3), c07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
    const float c08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2), c10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2);
    const float c11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2), c12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3);
    const float c14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3), c15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
    const float c16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2), c18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2);
    const float c19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2), c20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1);
    const float c22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1), c23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);
    const float f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
    const float f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
    const float v0[4] = {M(1, 0), M(0, 0), M(0, 0), M(0, 0)}, v1[4] = {M(1, 1), M(0, 1), M(0, 1), M(0, 1)};
    const float v2[4] = {M(1, 2), M(0, 2), M(0, 2), M(0, 2)}, v3[4] = {M(1, 3), M(0, 3), M(0, 3), M(0, 3)};
    float inv[16];
    for (int i = 0; i < 4; i++) {
        const float sa = (i & 1) ? -1.f : 1.f, sb = -sa;
        inv[0 + i]  = (v1[i] * f0[i] - v2[i] * f1[i] + v3[i] * f2[i]) * sa;
        inv[4 + i]  = (v0[i] * f0[i] - v2[i] * f3[i] + v3[i] * f4[i]) * sb;
        inv[8 + i]  = (v0[i] * f1[i] - v1[i] * f3[i] + v3[i] * f5[i]) * sa;
        inv[12 + i] = (v0[i] * f2[i] - v1[i] * f4[i] + v2[i] * f5[i]) * sb;
    }
    const float det = (M(0, 0) * inv[0] + M(0, 1) * inv[4]) + (M(0, 2) * inv[8] + M(0, 3) * inv[12]);
    const float ood = 1.f / det;
    for (int i = 0; i < 16; i++)
        out[i] = inv[i] * ood;
}

void normalize3(const float *v, float *o)
{
    const float dot = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const float inv = 1.f / sqrtf(dot);
    o[0] = v[0] * inv, o[1] = v[1] * inv, o[2] = v[2] * inv;
}

void cross3(const float *x, const float *y, float *o)
{
    o[0] = x[1] * y[2] - y[1] * x[2];
    o[1] = x[2] * y[0] - y[2] * x[0];
    o[2] = x[0] * y[1] - y[0] * x[1];
}

nbvh_camera_params camera_from_matrix(const float *m, float yfov_rad, uint32_t w, uint32_t h)
{
    float inv[16];
    invert4(m, inv);
    nbvh_camera_params cp;
    float dir_raw[3], up_raw[3], dir[3], up[3];
    for (int r = 0; r < 3; r++) {
        cp.origin[r] = (inv[0 + r] * 0.f + inv[4 + r] * 0.f) + (inv[8 + r] * 0.f + inv[12 + r] * 1.f);
        dir_raw[r]   = (inv[0 + r] * 0.f + inv[4 + r] * 0.f) + (inv[8 + r] * -1.f + inv[12 + r] * 0.f);
        up_raw[r]    = (inv[0 + r] * 0.f + inv[4 + r] * 1.f) + (inv[8 + r] * 0.f + inv[12 + r] * 0.f);
    }
    normalize3(dir_raw, dir);
    normalize3(up_raw, up);
    const float fovy_deg = yfov_rad * 57.295779513082320876798154814105f;        // glm::degrees
    const float half_rad = (0.5f * fovy_deg) * 0.01745329251994329576923690768489f; // glm::radians
    const float plane_y  = 2.f * tanf(half_rad);
    const float aspect   = (float)w / (float)h;
    const float plane_x  = plane_y * aspect;
    float c1[3], c1n[3], c2[3], c2n[3];
    cross3(dir, up, c1);
    normalize3(c1, c1n);
    for (int i = 0; i < 3; i++)
        cp.du[i] = c1n[i] * plane_x;
    cross3(cp.du, dir, c2);
    normalize3(c2, c2n);
    for (int i = 0; i < 3; i++)
        cp.dv[i] = -c2n[i] * plane_y;
    for (int i = 0; i < 3; i++)
        cp.dir_top_left[i] = dir[i] - 0.5f * cp.du[i] - 0.5f * cp.dv[i];
    return cp;
}

bool read_file(const std::string &path, std::string &out)
{
    std::ifstream f(path, std::ios::binary);
    if (!f)
        return false;
    std::ostringstream ss;
    ss << f.rdbuf();
    out = ss.str();
    return true;
}

uint16_t f32_to_f16_host(float f) { return __half_as_ushort(__float2half_rn(f)); }

} // namespace

// ================================================================================================ context
constexpr int kPipeSlots = NBVH_MAX_FRAMES_IN_FLIGHT; // frames in flight of nbvh_render_host_submit

struct nbvh_ctx {
    int device = 0;
    int num_sms = 148;
    std::string error;

    nbvh_renderer_constants bvh_const = default_renderer_constants();
    nbvh_module_constants module_const = default_module_constants();
    nbvh_pt_output_constants pt_const = default_pt_constants();
    nbvh_path_constants path_const = default_path_constants();
    nbvh_scene_constants scene_const = default_scene_constants();
    nbvh_camera_params camera{};
    float cam_matrix[16] = {};  // as given to nbvh_set_camera (save_config writes it back)
    float cam_yfov = 0.f;
    bool cam_set = false;
    std::string dir_enc_type = "Identity"; // direction_encoding.type: kept for save_config (unused while input_segment_direction_dim == 0)
    nbvh_bbox3f scene_bbox{};
    bool scene_bbox_set = false;
    nbvh_network_desc net_desc{8, 4, 18, 8, 2.0f, 64, 4};
    nbvh_optimizer_desc opt_desc{1e-2f, 10000, 2000, 0.33f, 0.9f, 0.999f, 1e-8f, 1e-8f};
    int max_depth = 2; // NeuralPathModule::m_max_depth default (path_module.h:313)
    uint32_t fb_w = 0, fb_h = 0;
    uint32_t training_step = 0;

    // network parameter layout
    uint64_t n_params = 0, n_mlp_params = 0;
    std::vector<uint32_t> level_offsets; // entries
    uint32_t in_k = 0, n_layers = 0;
    uint16_t *d_params = nullptr; // fp16: MLP then grid
    uint8_t *d_wimg    = nullptr;
    uint32_t wimg_bytes = 0;
    std::vector<uint32_t> layer_off;
    bool params_loaded = false;

    // tree-cut
    float4 *d_nbvh = nullptr, *d_nbvh_dilated = nullptr;
    uint32_t nbvh_nodes = 0;
    uint32_t nbvh_capacity = 0; // allocated nodes (tree-cut learning appends two nodes per split)
    // tree-cut learning state (the reference keeps it in NeuralBVHRenderer: m_neural_tree_cut_data, m_neural_node_map)
    nbvh_tree_cut_data *d_tree_cut = nullptr;
    uint32_t *d_node_map = nullptr;       // tree-cut node -> node of the neural mesh's classical BLAS
    // frames in flight (nbvh_render_host_submit / _wait): two 8-bit frame buffers, a render and a copy stream
    // slot 1 renders on its own stream with its own neural ray queue and counters, so that its k_primary overlaps the drain of
    // slot 0's k_neural (60 % of a 1080p launch leaves the SMs mostly idle)
    int render_slot = 0; // which queue / counter set render_impl uses (set by nbvh_render_host_submit)
    // slots 1 .. kPipeSlots-1 each own a queue, a counter block and a render stream (index slot - 1); slot 0 uses the main ones
    uint32_t *d_next_chunk_alt[kPipeSlots - 1] = {};
    NeuralQueueEntry *d_queue_alt[kPipeSlots - 1] = {};
    size_t queue_capacity_alt[kPipeSlots - 1] = {};
    cudaStream_t pipe_render_alt[kPipeSlots - 1] = {};
    uchar4 *d_ldr_slot[kPipeSlots] = {};
    size_t ldr_slot_pixels = 0;
    cudaStream_t pipe_render = nullptr, pipe_copy = nullptr;
    cudaEvent_t pipe_rendered[kPipeSlots] = {}, pipe_copied[kPipeSlots] = {};
    void *d_build_ws = nullptr; // nbvh_build_blas / nbvh_flip_images scratch (grow-only)
    float flip_ppd = 0.f;       // pixels per degree the FLIP filters in constant memory were built for
    size_t build_ws_bytes = 0;
    void *d_tc_workspace = nullptr;
    size_t tc_workspace_bytes = 0;
    TreeCutResult *d_tc_result = nullptr;
    bool tree_cut_ready = false;
    uint32_t tree_cut_max_depth = 0;
    float *d_node_keep = nullptr;    // per-node survival probability of the sampler's roulette (grow-only)
    uint32_t node_keep_capacity = 0;
    void *d_sampler_items = nullptr; // two-pass sample generator: 48-byte items handed from k_sample_walk to k_sample_targets (grow-only)
    uint32_t sampler_item_capacity = 0;
    float dilated_with  = -1.f;
    bool nbvh_safe = false;            // node arrays passed k_check_nodes (hoisted IEEE division precondition)
    bool blas_safe[kMaxBlas] = {};
    bool tlas_safe = false;

    // scene
    SceneConst scene{};
    bool scene_set = false;
    std::vector<nbvh_cuda_blas> blas_desc; // host copy of the caller's descriptors (shading attributes for the sampler)
    const void *materials = nullptr, *textures = nullptr;
    std::vector<std::pair<cudaTextureObject_t, cudaArray_t>> own_textures; // nbvh_create_texture_rgba8 / nbvh_create_envmap_rgba32f
    cudaTextureObject_t envmap = 0;                                        // envmap_texture (cuda_backend_constants.cuh); 0 = white

    // multi-bounce path state (max_depth > 1): one PathState + PathHit per pixel, 2 queue counters per bounce
    PathState *d_paths[2] = {nullptr, nullptr}; // ping-pong path lists (survivors are compacted at every bounce)
    PathHit *d_path_hits[2] = {nullptr, nullptr};
    size_t path_capacity = 0;
    uint32_t *d_bounce_counters = nullptr;

    // framebuffer
    float4 *accum = nullptr;
    uchar4 *ldr   = nullptr;
    bool own_accum = false, own_ldr = false;
    nbvh_hit_record *hits = nullptr;
    nbvh_hit_record *own_hits = nullptr;

    uint32_t *d_next_chunk = nullptr; // [0] chunk counter (mega kernel), [1] queue count, [2] queue cursor, [3] node-check flag, [4] sampler slot counter, [5] image-error sum, [6] BLAS depth, [7] sampler item count
    NeuralQueueEntry *d_queue = nullptr;
    size_t queue_capacity = 0;
    bool l2_prefetch = false; // NBVH_L2_PREFETCH=1 enables a per-frame L2 warm-up of the gather tables (A/B: no effect measured)
    bool mega_kernel = false; // NBVH_RENDER_MODE=mega: the single persistent kernel of the first design (kept for A/B timing)
    unsigned long long *d_counters = nullptr;

    // kernel timing
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float total_ms = 0.f;
    uint32_t launches = 0;
    // stage timing of the split frame pipeline: event triples (start, between k_primary and k_neural, end) recorded
    // without host synchronisation; read back by nbvh_get_stage_times
    std::vector<cudaEvent_t> stage_ev;
    uint32_t stage_used = 0; // frames recorded

    // training state (allocated lazily): tcnn::Trainer's buffers (trainer.h:68-87) + Adam moments and step counters
    float *d_grads = nullptr, *d_master = nullptr, *d_m = nullptr, *d_v = nullptr;
    uint32_t *d_steps = nullptr;
    uint8_t *d_feat = nullptr, *d_dfeat = nullptr;
    uint32_t feat_capacity = 0; // samples
    float *d_loss = nullptr;
    float *own_grads = nullptr; // internal gradient buffer (d_grads may point at a host-owned one instead)
    bool train_state_valid = false;
    uint32_t opt_step = 0;    // optimiser steps since the state was (re)initialised (tcnn: optimizer->step())
    float lr_factor   = 1.f;  // ExponentialDecay's running factor
};

namespace {

int fail(nbvh_ctx *ctx, int code, const std::string &msg)
{
    if (ctx)
        ctx->error = msg;
    return code;
}

uint32_t grid_resolution(float scale) { return (uint32_t)ceilf(scale) + 1u; }
float grid_scale(uint32_t level, float log2_pls, uint32_t base) { return exp2f((float)level * log2_pls) * (float)base - 1.0f; }

// Parameter layout: FullyFusedMLP matrices (row-major W[out][in]) then the grid levels
// (network_with_input_encoding.h:115-122, fully_fused_mlp.cu:636-678, grid.h:692-723)
int rebuild_layout(nbvh_ctx *ctx)
{
    const nbvh_network_desc &d = ctx->net_desc;
    if (d.n_features_per_level != 4)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "only n_features_per_level == 4 is implemented");
    if (d.n_neurons != kNeurons)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "only n_neurons == 64 is implemented");
    if (d.n_levels < 1 || d.n_levels > kMaxLevels || d.n_hidden_layers < 1 || d.n_hidden_layers + 1 > kMaxLayers)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "n_levels / n_hidden_layers out of range");
    if (ctx->module_const.input_segment_direction_dim != 0)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "direction encoding input is not implemented");
    const uint32_t n_points = ctx->module_const.input_segment_points;
    if (n_points < 1 || n_points > 4)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "1..4 segment points supported");
    ctx->in_k = n_points * d.n_levels * d.n_features_per_level;
    if (ctx->in_k % 16u != 0u || ctx->in_k > kMaxInputK)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "encoded input width must be a multiple of 16 and <= 128");
    if (ctx->pt_const.output_network_dims > 8)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "output_network_dims > 8");
    ctx->n_layers = (uint32_t)d.n_hidden_layers + 1u;

    ctx->level_offsets.assign(d.n_levels + 1, 0);
    uint32_t offset    = 0;
    const float l2     = log2f(d.per_level_scale);
    for (int l = 0; l < d.n_levels; l++) {
        const uint32_t res = grid_resolution(grid_scale(l, l2, d.base_resolution));
        const uint32_t max_params = 0xFFFFFFFFu / 2u;
        uint32_t in_level = powf((float)res, 3.f) > (float)max_params ? max_params : res * res * res;
        in_level          = (in_level + 7u) / 8u * 8u;
        const uint32_t hs = 1u << d.log2_hashmap_size;
        if (in_level > hs)
            in_level = hs;
        ctx->level_offsets[l] = offset;
        offset += in_level;
    }
    ctx->level_offsets[d.n_levels] = offset;
    ctx->n_mlp_params = (uint64_t)kNeurons * ctx->in_k + (uint64_t)(d.n_hidden_layers - 1) * kNeurons * kNeurons +
                        (uint64_t)kOutPadded * kNeurons;
    ctx->n_params = ctx->n_mlp_params + (uint64_t)offset * d.n_features_per_level;

    // weights image: one canonical K-major operand block per layer
    ctx->layer_off.assign(ctx->n_layers, 0);
    uint32_t bytes = 0;
    for (uint32_t l = 0; l < ctx->n_layers; l++) {
        const uint32_t k = l == 0 ? ctx->in_k : (uint32_t)kNeurons;
        const uint32_t n = l + 1 == ctx->n_layers ? (uint32_t)kOutPadded : (uint32_t)kNeurons;
        ctx->layer_off[l] = bytes;
        bytes += n * k * 2u;
    }
    ctx->wimg_bytes = bytes;
    if (ctx->d_params)
        cudaFree(ctx->d_params);
    if (ctx->d_wimg)
        cudaFree(ctx->d_wimg);
    ctx->d_params = nullptr, ctx->d_wimg = nullptr;
    ctx->params_loaded = false;
    cudaFree(ctx->own_grads), cudaFree(ctx->d_master), cudaFree(ctx->d_m), cudaFree(ctx->d_v), cudaFree(ctx->d_steps);
    ctx->d_grads = ctx->own_grads = ctx->d_master = ctx->d_m = ctx->d_v = nullptr;
    ctx->d_steps = nullptr;
    ctx->train_state_valid = false;
    CU_TRY(ctx, cudaMalloc(&ctx->d_params, ctx->n_params * sizeof(uint16_t)));
    CU_TRY(ctx, cudaMemset(ctx->d_params, 0, ctx->n_params * sizeof(uint16_t)));
    CU_TRY(ctx, cudaMalloc(&ctx->d_wimg, ctx->wimg_bytes));
    CU_TRY(ctx, configure_kernels(ctx->wimg_bytes, ctx->in_k));
    CU_TRY(ctx, configure_render(ctx->wimg_bytes, ctx->in_k));
    CU_TRY(ctx, configure_training(ctx->wimg_bytes, ctx->in_k, ctx->n_layers));
    return NBVH_OK;
}

// Re-packs the MLP matrices into the UMMA canonical layout and uploads them
int upload_weight_image(nbvh_ctx *ctx, const uint16_t *mlp_host)
{
    std::vector<uint8_t> img(ctx->wimg_bytes, 0);
    const uint16_t *w = mlp_host;
    for (uint32_t l = 0; l < ctx->n_layers; l++) {
        const uint32_t k = l == 0 ? ctx->in_k : (uint32_t)kNeurons;
        const uint32_t n = l + 1 == ctx->n_layers ? (uint32_t)kOutPadded : (uint32_t)kNeurons;
        uint8_t *dst     = img.data() + ctx->layer_off[l];
        for (uint32_t r = 0; r < n; r++)
            for (uint32_t c = 0; c < k; c++) {
                const uint32_t off = (r >> 3) * ((k >> 3) * 128u) + (c >> 3) * 128u + (r & 7u) * 16u + (c & 7u) * 2u;
                memcpy(dst + off, &w[(size_t)r * k + c], 2);
            }
        w += (size_t)n * k;
    }
    CU_TRY(ctx, cudaMemcpy(ctx->d_wimg, img.data(), img.size(), cudaMemcpyHostToDevice));
    return NBVH_OK;
}

NetConst make_net_const(const nbvh_ctx *ctx)
{
    NetConst nc;
    memset(&nc, 0, sizeof(nc));
    const nbvh_network_desc &d = ctx->net_desc;
    nc.grid     = reinterpret_cast<const uint2 *>(ctx->d_params + ctx->n_mlp_params);
    nc.n_levels = d.n_levels;
    nc.n_points = ctx->module_const.input_segment_points;
    nc.in_k     = ctx->in_k;
    nc.n_layers = ctx->n_layers;
    nc.n_out    = ctx->pt_const.output_network_dims;
    nc.wimg     = ctx->d_wimg;
    nc.wimg_bytes = ctx->wimg_bytes;
    for (uint32_t l = 0; l < ctx->n_layers; l++)
        nc.layer_off[l] = ctx->layer_off[l];
    const float l2 = log2f(d.per_level_scale);
    for (int l = 0; l < d.n_levels; l++) {
        GridLevel &L = nc.lv[l];
        L.offset     = ctx->level_offsets[l];
        L.size       = ctx->level_offsets[l + 1] - ctx->level_offsets[l];
        L.mask       = (L.size > 1u && (L.size & (L.size - 1u)) == 0u) ? L.size - 1u : 0u;
        L.scale      = grid_scale(l, l2, d.base_resolution);
        L.res        = grid_resolution(L.scale);
        // tcnn grid_index: dense indexing unless the dense stride product exceeds the level size
        const uint64_t r = L.res;
        L.dense          = (r <= L.size && r * r <= L.size && r * r * r <= (uint64_t)L.size) ? 1u : 0u;
    }
    return nc;
}

NbvhConst make_nbvh_const(const nbvh_ctx *ctx, bool inference)
{
    NbvhConst bc;
    bc.nodes        = ctx->d_nbvh_dilated;
    bc.num_nodes    = ctx->nbvh_nodes;
    bc.leaf_lod_max = 2u + (uint32_t)ctx->bvh_const.rendered_lod;
    bc.mint_eps     = inference ? ctx->bvh_const.inference_mint_epsilon : ctx->bvh_const.learning_mint_epsilon;
    bc.maxt_eps     = inference ? ctx->bvh_const.inference_maxt_epsilon : ctx->bvh_const.learning_maxt_epsilon;
    bc.dilation     = ctx->bvh_const.fixed_node_dilation;
    bc.predilated   = 1u;
    return bc;
}

// hoisted-division precondition on a node array (device_core.cuh): synchronous, setup time only
int check_nodes(nbvh_ctx *ctx, const float4 *nodes, uint32_t n, bool &ok)
{
    uint32_t one = 1u, result = 0u;
    uint32_t *flag = ctx->d_next_chunk + 3;
    CU_TRY(ctx, cudaMemcpy(flag, &one, 4, cudaMemcpyHostToDevice));
    CU_TRY(ctx, launch_check_nodes(nodes, n, flag, 0));
    CU_TRY(ctx, cudaMemcpy(&result, flag, 4, cudaMemcpyDeviceToHost));
    ok = result != 0u;
    return NBVH_OK;
}

// keep the pre-dilated node copy in sync with fixed_node_dilation
int refresh_dilated(nbvh_ctx *ctx)
{
    if (!ctx->d_nbvh)
        return fail(ctx, NBVH_ERR_STATE, "no neural BVH set");
    if (ctx->bvh_const.ellipsoidal_bvh)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "ellipsoidal_bvh is not implemented");
    if (ctx->dilated_with == ctx->bvh_const.fixed_node_dilation)
        return NBVH_OK;
    if (ctx->bvh_const.fixed_node_dilation == 0.f) {
        CU_TRY(ctx, cudaMemcpy(ctx->d_nbvh_dilated, ctx->d_nbvh, (size_t)ctx->nbvh_nodes * 32, cudaMemcpyDeviceToDevice));
    } else {
        CU_TRY(ctx, launch_predilate(ctx->d_nbvh, ctx->d_nbvh_dilated, ctx->nbvh_nodes, ctx->bvh_const.fixed_node_dilation, 0));
        CU_TRY(ctx, cudaDeviceSynchronize());
    }
    ctx->dilated_with = ctx->bvh_const.fixed_node_dilation;
    return check_nodes(ctx, ctx->d_nbvh_dilated, ctx->nbvh_nodes, ctx->nbvh_safe);
}

EncodeConst make_encode_const(const nbvh_ctx *ctx)
{
    EncodeConst ec;
    memset(&ec, 0, sizeof(ec));
    // neural_bvh_module_kernels.cu:39-43 (host evaluation of the same float expressions, no contraction)
    const float dilation = (ctx->bvh_const.node_extent_dilation + 1.f) + ctx->bvh_const.fixed_node_dilation;
    for (int i = 0; i < 3; i++) {
        ec.extent[i] = dilation * (ctx->scene_bbox.max[i] - ctx->scene_bbox.min[i]);
        ec.center[i] = (ctx->scene_bbox.max[i] + ctx->scene_bbox.min[i]) * 0.5f;
    }
    const uint32_t n = ctx->module_const.input_segment_points;
    ec.n_points      = n;
    float delta, offset;
    if (ctx->module_const.segment_distribution == 1) { // EquidistantExterior
        delta  = n == 1 ? 1.f : 1.f / (float)(n - 1);
        offset = 0.f;
    } else {
        delta  = 1.f / (float)n;
        offset = 0.5f;
    }
    for (uint32_t k = 0; k < n && k < 4; k++)
        ec.point_s[k] = delta * ((float)k + offset);
    return ec;
}

int check_ready(nbvh_ctx *ctx, bool need_scene, bool need_nbvh, bool need_params)
{
    if (need_params && !ctx->params_loaded)
        return fail(ctx, NBVH_ERR_STATE, "network parameters not loaded");
    if (need_nbvh && !ctx->d_nbvh)
        return fail(ctx, NBVH_ERR_STATE, "neural BVH not set");
    if (need_scene && !ctx->scene_set)
        return fail(ctx, NBVH_ERR_STATE, "scene not set");
    if (ctx->module_const.segment_distribution == 2)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "Uniform segment distribution is not implemented");
    if (ctx->pt_const.depth_learning_mode != 0)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "only depth_learning_mode LocalT is implemented");
    return NBVH_OK;
}

struct KernelTimer {
    nbvh_ctx *ctx;
    cudaStream_t st;
    KernelTimer(nbvh_ctx *c, cudaStream_t s) : ctx(c), st(s)
    {
        if (ctx->timing)
            cudaEventRecord(ctx->ev0, st);
    }
    void stop()
    {
        if (!ctx->timing)
            return;
        cudaEventRecord(ctx->ev1, st);
        cudaEventSynchronize(ctx->ev1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        ctx->total_ms += ms;
        ctx->launches++;
    }
};

} // namespace

// ================================================================================================ C ABI
extern "C" {

uint32_t nbvh_abi_version(void) { return NBVH_ABI_VERSION; }

const char *nbvh_last_error(const nbvh_ctx *ctx) { return ctx ? ctx->error.c_str() : "null context"; }

int nbvh_create(int device, const char *config_json, nbvh_ctx **out)
{
    if (!out)
        return NBVH_ERR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count)
        return NBVH_ERR_CUDA; // no CPU fallback
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess)
        return NBVH_ERR_CUDA;
    if (prop.major != 10)
        return NBVH_ERR_CUDA; // the kernels exist for sm_100a only
    nbvh_ctx *ctx = new nbvh_ctx();
    ctx->device   = device;
    ctx->num_sms  = prop.multiProcessorCount;
    if (const char *m = getenv("NBVH_RENDER_MODE"))
        ctx->mega_kernel = std::string(m) == "mega";
    if (const char *m = getenv("NBVH_L2_PREFETCH"))
        ctx->l2_prefetch = std::string(m) != "0";
    if (cudaMalloc(&ctx->d_next_chunk, 8 * sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc(&ctx->d_counters, sizeof(unsigned long long) * (CNT_COUNT + 1)) != cudaSuccess) {
        delete ctx;
        return NBVH_ERR_CUDA;
    }
    cudaMemset(ctx->d_counters, 0, sizeof(unsigned long long) * CNT_COUNT);
    cudaEventCreate(&ctx->ev0);
    cudaEventCreate(&ctx->ev1);
    int rc = config_json ? nbvh_load_config_json(ctx, config_json) : rebuild_layout(ctx);
    *out   = ctx; // returned even on a config error so the caller can read nbvh_last_error
    return rc;
}

void nbvh_destroy(nbvh_ctx *ctx)
{
    if (!ctx)
        return;
    cudaSetDevice(ctx->device);
    nbvh_destroy_textures(ctx);
    cudaFree(ctx->d_params);
    cudaFree(ctx->d_wimg);
    cudaFree(ctx->d_nbvh);
    cudaFree(ctx->d_nbvh_dilated);
    if (ctx->own_accum)
        cudaFree(ctx->accum);
    if (ctx->own_ldr)
        cudaFree(ctx->ldr);
    cudaFree(ctx->own_hits);
    cudaFree(ctx->d_next_chunk);
    cudaFree(ctx->d_sampler_items);
    cudaFree(ctx->d_node_keep);
    cudaFree(ctx->d_queue);
    for (int k = 0; k < 2; k++)
        cudaFree(ctx->d_paths[k]), cudaFree(ctx->d_path_hits[k]);
    cudaFree(ctx->d_build_ws);
    for (int k = 0; k < kPipeSlots - 1; k++) {
        cudaFree(ctx->d_next_chunk_alt[k]);
        cudaFree(ctx->d_queue_alt[k]);
        if (ctx->pipe_render_alt[k])
            cudaStreamDestroy(ctx->pipe_render_alt[k]);
    }
    for (int k = 0; k < kPipeSlots; k++) {
        cudaFree(ctx->d_ldr_slot[k]);
        if (ctx->pipe_rendered[k])
            cudaEventDestroy(ctx->pipe_rendered[k]);
        if (ctx->pipe_copied[k])
            cudaEventDestroy(ctx->pipe_copied[k]);
    }
    if (ctx->pipe_render)
        cudaStreamDestroy(ctx->pipe_render);
    if (ctx->pipe_copy)
        cudaStreamDestroy(ctx->pipe_copy);
    cudaFree(ctx->d_tree_cut);
    cudaFree(ctx->d_node_map);
    cudaFree(ctx->d_tc_workspace);
    cudaFree(ctx->d_tc_result);
    cudaFree(ctx->d_bounce_counters);
    cudaFree(ctx->d_counters);
    cudaFree(ctx->own_grads);
    cudaFree(ctx->d_master);
    cudaFree(ctx->d_m);
    cudaFree(ctx->d_v);
    cudaFree(ctx->d_steps);
    cudaFree(ctx->d_feat);
    cudaFree(ctx->d_dfeat);
    cudaFree(ctx->d_loss);
    if (ctx->ev0)
        cudaEventDestroy(ctx->ev0);
    if (ctx->ev1)
        cudaEventDestroy(ctx->ev1);
    for (cudaEvent_t e : ctx->stage_ev)
        cudaEventDestroy(e);
    delete ctx;
}

// ---- configuration ---------------------------------------------------------------------------------
int nbvh_load_config_json(nbvh_ctx *ctx, const char *json_text)
{
    if (!ctx || !json_text)
        return NBVH_ERR_INVALID_ARG;
    Json root;
    try {
        root = json_parse(json_text);
    } catch (const std::exception &e) {
        return fail(ctx, NBVH_ERR_PARSE, e.what());
    }
    const Json &scene = root.object("cuda_scene");
    const Json &rend  = root.object("neural_bvh_renderer");
    const Json &mod   = rend.object("neural_module");

    // scene constants (cuda_backend.h:93-107)
    {
        const Json &j           = scene.object("scene_constants");
        nbvh_scene_constants d  = default_scene_constants(), &c = ctx->scene_const;
        c.force_constant_material = j.value("force_constant_material", d.force_constant_material);
        c.constant_roughness    = (float)j.value("constant_roughness", (double)d.constant_roughness);
        c.constant_specular     = (float)j.value("constant_specular", (double)d.constant_specular);
        c.constant_metallic     = (float)j.value("constant_metallic", (double)d.constant_metallic);
        c.envmap_rotation       = (float)j.value("envmap_rotation", (double)d.envmap_rotation);
        c.envmap_strength       = (float)j.value("envmap_strength", (double)d.envmap_strength);
        c.ema_accumulate        = j.value("ema_accumulate", d.ema_accumulate);
        c.ema_weight            = (float)j.value("ema_weight", (double)d.ema_weight);
        c.envmap_eval           = j.value("envmap_eval", d.envmap_eval);
        c.emission_eval         = j.value("emission_eval", d.emission_eval);
        c.skylike_background    = j.value("skylike_background", d.skylike_background);
    }
    // bvh constants (neural_bvh_renderer_constants.cuh:77-106)
    {
        const Json &j = rend.object("bvh_constants");
        nbvh_renderer_constants d = default_renderer_constants(), &c = ctx->bvh_const;
#define GETI(name) c.name = (int32_t)j.value(#name, (double)d.name)
#define GETF(name) c.name = (float)j.value(#name, (double)d.name)
#define GETB(name) c.name = j.value(#name, (bool)d.name)
        GETI(num_lods_learned);
        GETI(current_learned_lod);
        GETI(rendered_lod);
        GETB(ignore_neural_blas);
        GETF(leaf_boundary_sampling_ratio);
        GETF(inference_mint_epsilon);
        GETF(inference_maxt_epsilon);
        GETF(learning_mint_epsilon);
        GETF(learning_maxt_epsilon);
        GETF(sampling_bbox_scaling);
        GETF(fixed_node_dilation);
        GETF(node_extent_dilation);
        GETF(loss_exposure);
        GETB(show_loss_heuristic);
        GETB(ellipsoidal_bvh);
        GETB(rescale_ellipsoid);
        GETI(log2_bvh_max_traversals);
        GETB(sample_cosine_weighted);
        GETB(backface_bounce_training);
        GETI(max_training_bounces);
        GETB(training_node_survival);
        GETF(survival_probability_scale);
        GETF(min_rr_survival_prob);
#undef GETI
#undef GETF
#undef GETB
    }
    // module constants (neural_bvh_module.h:112-130)
    {
        const Json &j = mod.object("bvh_module_constants");
        nbvh_module_constants d = default_module_constants(), &c = ctx->module_const;
        c.input_segment_point_dim     = (uint32_t)j.value("input_segment_point_dim", (double)d.input_segment_point_dim);
        c.input_segment_points        = (uint32_t)j.value("input_segment_points", (double)d.input_segment_points);
        c.input_segment_points_dim    = (uint32_t)j.value("input_segment_points_dim", (double)d.input_segment_points_dim);
        c.input_segment_direction_dim = (uint32_t)j.value("input_segment_direction_dim", (double)d.input_segment_direction_dim);
        c.input_network_dims          = (uint32_t)j.value("input_network_dims", (double)d.input_network_dims);
        const int sd = find_index(j.value("segment_distribution", std::string(kSegmentDistributions[d.segment_distribution])),
                                  kSegmentDistributions);
        if (sd < 0)
            return fail(ctx, NBVH_ERR_PARSE, "unknown segment_distribution");
        c.segment_distribution = sd;
        c.antialias_direction  = j.value("antialias_direction", d.antialias_direction);
        c.merge_operator       = j.value("merge_operator", d.merge_operator);
        c.apply_encoding_trafo = j.value("apply_encoding_trafo", d.apply_encoding_trafo);
        // derived, whatever the file says: NeuralBVHModule::reinitialise_neural_model (neural_bvh_module.cu:126-131) -- the shipped
        // partial configs state only input_segment_points / input_segment_direction_dim
        c.input_segment_points_dim = c.input_segment_points * c.input_segment_point_dim;
        c.input_network_dims       = c.input_segment_points_dim + c.input_segment_direction_dim;
        if (c.apply_encoding_trafo)
            return fail(ctx, NBVH_ERR_UNSUPPORTED, "apply_encoding_trafo (asserted off in the reference) is not implemented");
    }
    // path-tracing output constants (neural_bvh_module.h:305-351)
    {
        const Json &j = mod.object("pt_output_constants");
        nbvh_pt_output_constants d = default_pt_constants(), &c = ctx->pt_const;
        c.output_visibility_dim = (uint32_t)j.value("output_visibility_dim", (double)d.output_visibility_dim);
        c.output_depth_dim      = (uint32_t)j.value("output_depth_dim", (double)d.output_depth_dim);
        c.output_normal_dim     = (uint32_t)j.value("output_normal_dim", (double)d.output_normal_dim);
        c.output_albedo_dim     = (uint32_t)j.value("output_albedo_dim", (double)d.output_albedo_dim);
        // derived: NeuralPathModule::update_neural_model_module (path_module_init.cu:259-265) overrides the file's value with the sum
        c.output_network_dims   = c.output_visibility_dim + c.output_depth_dim + c.output_normal_dim + c.output_albedo_dim;
        struct {
            const char *key;
            uint8_t *dst;
            uint8_t def;
        } losses[] = {{"visibility_loss", &c.visibility_loss, d.visibility_loss},
                      {"depth_loss", &c.depth_loss, d.depth_loss},
                      {"normal_loss", &c.normal_loss, d.normal_loss},
                      {"albedo_loss", &c.albedo_loss, d.albedo_loss}};
        for (auto &l : losses) {
            const int idx = find_index(j.value(l.key, std::string(kLossTypes[l.def])), kLossTypes);
            if (idx < 0)
                return fail(ctx, NBVH_ERR_PARSE, std::string("unknown loss type for ") + l.key);
            *l.dst = (uint8_t)idx;
        }
        c.visibility_loss_scale = (float)j.value("visibility_loss_scale", (double)d.visibility_loss_scale);
        c.depth_loss_scale      = (float)j.value("depth_loss_scale", (double)d.depth_loss_scale);
        c.normal_loss_scale     = (float)j.value("normal_loss_scale", (double)d.normal_loss_scale);
        c.albedo_loss_scale     = (float)j.value("albedo_loss_scale", (double)d.albedo_loss_scale);
        const int dm = find_index(j.value("depth_learning_mode", std::string(kDepthModes[d.depth_learning_mode])), kDepthModes);
        if (dm < 0)
            return fail(ctx, NBVH_ERR_PARSE, "unknown depth_learning_mode");
        c.depth_learning_mode             = (uint8_t)dm;
        c.output_normal_linear_activation = j.value("output_normal_linear_activation", d.output_normal_linear_activation);
        c.output_depth_linear_activation  = j.value("output_depth_linear_activation", d.output_depth_linear_activation);
        c.output_albedo_linear_activation = j.value("output_albedo_linear_activation", d.output_albedo_linear_activation);
        c.visibility_masking_loss         = j.value("visibility_masking_loss", d.visibility_masking_loss);
        c.primary_visibility_threshold    = (float)j.value("primary_visibility_threshold", (double)d.primary_visibility_threshold);
        c.secondary_visibility_threshold  = (float)j.value("secondary_visibility_threshold", (double)d.secondary_visibility_threshold);
        c.depth_epsilon                   = (float)j.value("depth_epsilon", (double)d.depth_epsilon);
        c.fixed_depth_epsilon             = (float)j.value("fixed_depth_epsilon", (double)d.fixed_depth_epsilon);
    }
    // path module constants (path_module.h:242-253)
    {
        const Json &j = mod.object("path_module_constants");
        nbvh_path_constants d = default_path_constants(), &c = ctx->path_const;
        const int om = find_index(j.value("output_mode", std::string(kOutputModes[d.output_mode])), kOutputModes);
        if (om < 0)
            return fail(ctx, NBVH_ERR_PARSE, "unknown output_mode");
        c.output_mode            = (uint8_t)om;
        c.normal_epsilon_offset  = (float)j.value("normal_epsilon_offset", (double)d.normal_epsilon_offset);
        c.apply_russian_roulette = j.value("apply_russian_roulette", d.apply_russian_roulette);
        c.rr_start_depth         = (int32_t)j.value("rr_start_depth", (double)d.rr_start_depth);
        c.neural_roughness       = (float)j.value("neural_roughness", (double)d.neural_roughness);
        c.neural_specular        = (float)j.value("neural_specular", (double)d.neural_specular);
        c.neural_metallic        = (float)j.value("neural_metallic", (double)d.neural_metallic);
    }
    ctx->max_depth = (int)mod.value("max_depth", (double)ctx->max_depth);
    // network + encoding + optimizer (tcnn_config.h)
    {
        const Json &nw = mod.object("network");
        const Json &hg = mod.object("segment_encoding").object("hashgrid_options");
        nbvh_network_desc &d = ctx->net_desc;
        const std::string net_type = nw.value("network", std::string("FullyFusedMLP"));
        if (net_type != "FullyFusedMLP" && net_type != "CutlassMLP")
            return fail(ctx, NBVH_ERR_PARSE, "unknown network type");
        if (nw.value("activation", std::string("ReLU")) != "ReLU" ||
            nw.value("output_activation", std::string("Sigmoid")) != "Sigmoid")
            return fail(ctx, NBVH_ERR_UNSUPPORTED, "only ReLU hidden / Sigmoid output activations are implemented");
        const std::string enc_type = mod.object("segment_encoding").value("type", std::string("HashGrid"));
        if (enc_type != "HashGrid")
            return fail(ctx, NBVH_ERR_UNSUPPORTED, "only the HashGrid segment encoding is implemented");
        if (mod.object("segment_encoding").value("reduction", std::string("Concatenation")) != "Concatenation")
            return fail(ctx, NBVH_ERR_UNSUPPORTED, "only the Concatenation reduction is implemented");
        if (hg.value("hash", std::string("CoherentPrime")) != "CoherentPrime" ||
            hg.value("interpolation", std::string("Linear")) != "Linear" || hg.value("stochastic_interpolation", false))
            return fail(ctx, NBVH_ERR_UNSUPPORTED, "only CoherentPrime hash with Linear interpolation is implemented");
        d.n_neurons            = (int32_t)nw.value("n_neurons", 64.0);
        d.n_hidden_layers      = (int32_t)nw.value("n_hidden_layers", 4.0);
        d.n_levels             = (int32_t)hg.value("n_levels", 8.0);
        d.n_features_per_level = (int32_t)hg.value("n_features_per_level", 4.0);
        d.log2_hashmap_size    = (int32_t)hg.value("log2_hashmap_size", 18.0);
        d.base_resolution      = (int32_t)hg.value("base_resolution", 8.0);
        d.per_level_scale      = (float)hg.value("per_level_scale", 2.0);
        const Json &op         = mod.object("optimizer");
        nbvh_optimizer_desc &o = ctx->opt_desc;
        if (op.value("type", std::string("Adam")) != "Adam")
            return fail(ctx, NBVH_ERR_UNSUPPORTED, "only the Adam optimizer is implemented");
        o.learning_rate  = (float)op.value("learning_rate", 1e-2);
        o.decay_start    = (int32_t)op.value("decay_start", 10000.0);
        o.decay_interval = (int32_t)op.value("decay_interval", 2000.0);
        o.decay_base     = (float)op.value("decay_base", (double)0.33f);
        o.beta1          = (float)op.value("beta1", (double)0.9f);
        o.beta2          = (float)op.value("beta2", (double)0.999f);
        o.epsilon        = (float)op.value("epsilon", (double)1e-8f);
        o.l2_reg         = (float)op.value("l2_reg", (double)1e-8f);
    }
    ctx->dir_enc_type  = mod.object("direction_encoding").value("type", std::string("Identity"));
    ctx->training_step = (uint32_t)rend.value("training_step", 0.0);
    // framebuffer size + camera
    if (const Json *fb = scene.find("fb_size")) {
        if (fb->size() == 2) {
            int rc = nbvh_set_framebuffer(ctx, nullptr, nullptr, (uint32_t)fb->at_num(0), (uint32_t)fb->at_num(1));
            if (rc != NBVH_OK)
                return rc;
        }
    }
    const Json &cam = scene.object("camera");
    if (const Json *m = cam.find("matrix")) {
        if (m->size() != 16)
            return fail(ctx, NBVH_ERR_PARSE, "camera.matrix must have 16 entries");
        float mtx[16];
        for (int i = 0; i < 16; i++)
            mtx[i] = (float)m->at_num(i);
        const float yfov = (float)cam.value("yfov", (double)(45.0f * 0.01745329251994329576923690768489f));
        int rc           = nbvh_set_camera(ctx, mtx, yfov);
        if (rc != NBVH_OK)
            return rc;
    }
    ctx->dilated_with = -1.f;
    return rebuild_layout(ctx);
}

static void *block_ptr(nbvh_ctx *ctx, int which, size_t &bytes)
{
    switch (which) {
    case NBVH_CONST_RENDERER: bytes = sizeof(ctx->bvh_const); return &ctx->bvh_const;
    case NBVH_CONST_MODULE: bytes = sizeof(ctx->module_const); return &ctx->module_const;
    case NBVH_CONST_PT_OUTPUT: bytes = sizeof(ctx->pt_const); return &ctx->pt_const;
    case NBVH_CONST_PATH: bytes = sizeof(ctx->path_const); return &ctx->path_const;
    case NBVH_CONST_SCENE: bytes = sizeof(ctx->scene_const); return &ctx->scene_const;
    case NBVH_CONST_CAMERA: bytes = sizeof(ctx->camera); return &ctx->camera;
    case NBVH_CONST_SCENE_BBOX: bytes = sizeof(ctx->scene_bbox); return &ctx->scene_bbox;
    case NBVH_CONST_NETWORK: bytes = sizeof(ctx->net_desc); return &ctx->net_desc;
    case NBVH_CONST_OPTIMIZER: bytes = sizeof(ctx->opt_desc); return &ctx->opt_desc;
    default: bytes = 0; return nullptr;
    }
}

int nbvh_set_constants(nbvh_ctx *ctx, int which, const void *pod_host, size_t bytes)
{
    if (!ctx || !pod_host)
        return NBVH_ERR_INVALID_ARG;
    size_t want = 0;
    void *dst   = block_ptr(ctx, which, want);
    if (!dst || bytes != want)
        return fail(ctx, NBVH_ERR_INVALID_ARG, "nbvh_set_constants: unknown block or size mismatch");
    if (which == NBVH_CONST_NETWORK || which == NBVH_CONST_MODULE) {
        // these change the parameter layout: compare first so per-frame re-uploads stay cheap
        if (memcmp(dst, pod_host, bytes) == 0)
            return NBVH_OK;
        memcpy(dst, pod_host, bytes);
        return rebuild_layout(ctx);
    }
    memcpy(dst, pod_host, bytes);
    if (which == NBVH_CONST_SCENE_BBOX)
        ctx->scene_bbox_set = true;
    return NBVH_OK;
}

int nbvh_get_constants(const nbvh_ctx *ctx, int which, void *pod_host, size_t bytes)
{
    if (!ctx || !pod_host)
        return NBVH_ERR_INVALID_ARG;
    size_t want     = 0;
    const void *src = block_ptr(const_cast<nbvh_ctx *>(ctx), which, want);
    if (!src || bytes != want)
        return NBVH_ERR_INVALID_ARG;
    memcpy(pod_host, src, bytes);
    return NBVH_OK;
}

int nbvh_set_camera(nbvh_ctx *ctx, const float matrix16_host[16], float yfov_rad)
{
    if (!ctx || !matrix16_host)
        return NBVH_ERR_INVALID_ARG;
    if (ctx->fb_w == 0 || ctx->fb_h == 0)
        return fail(ctx, NBVH_ERR_STATE, "set the framebuffer size before the camera (aspect ratio)");
    ctx->camera = camera_from_matrix(matrix16_host, yfov_rad, ctx->fb_w, ctx->fb_h);
    memcpy(ctx->cam_matrix, matrix16_host, sizeof(ctx->cam_matrix));
    ctx->cam_yfov = yfov_rad, ctx->cam_set = true;
    return NBVH_OK;
}

int nbvh_get_max_depth(const nbvh_ctx *ctx) { return ctx ? ctx->max_depth : NBVH_ERR_INVALID_ARG; }
int nbvh_set_max_depth(nbvh_ctx *ctx, int max_depth)
{
    if (!ctx || max_depth < 1)
        return NBVH_ERR_INVALID_ARG;
    ctx->max_depth = max_depth;
    return NBVH_OK;
}

// ---- scene, tree-cut, weights --------------------------------------------------------------------------
int nbvh_set_scene(nbvh_ctx *ctx, const nbvh_scene_data *s)
{
    if (!ctx || !s)
        return NBVH_ERR_INVALID_ARG;
    if (s->num_meshes > (uint32_t)kMaxBlas)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "more than 16 BLAS");
    std::vector<nbvh_cuda_blas> blas(s->num_meshes);
    if (s->num_meshes)
        CU_TRY(ctx, cudaMemcpy(blas.data(), s->mesh_bvhs, sizeof(nbvh_cuda_blas) * s->num_meshes, cudaMemcpyDeviceToHost));
    memset(&ctx->scene, 0, sizeof(ctx->scene));
    ctx->scene.tlas.nodes        = reinterpret_cast<const float4 *>(s->tlas_nodes);
    ctx->scene.tlas.blas_indices = s->blas_indices;
    ctx->scene.tlas.num_meshes   = s->num_meshes;
    ctx->scene.n_blas            = s->num_meshes;
    for (uint32_t i = 0; i < s->num_meshes; i++) {
        ctx->scene.blas[i].nodes        = reinterpret_cast<const float4 *>(blas[i].nodes);
        ctx->scene.blas[i].prim_indices = reinterpret_cast<const unsigned long long *>(blas[i].prim_indices);
        ctx->scene.blas[i].vertices     = blas[i].vertices;
        ctx->scene.blas[i].indices      = blas[i].indices;
        int rc = check_nodes(ctx, ctx->scene.blas[i].nodes, blas[i].num_nodes, ctx->blas_safe[i]);
        if (rc != NBVH_OK)
            return rc;
    }
    {
        int rc = check_nodes(ctx, ctx->scene.tlas.nodes, s->num_tlas_nodes, ctx->tlas_safe);
        if (rc != NBVH_OK)
            return rc;
    }
    ctx->blas_desc = blas;
    ctx->materials = s->materials;
    ctx->textures  = s->textures;
    ctx->scene_set = true;
    return NBVH_OK;
}

static int adopt_nbvh(nbvh_ctx *ctx, const void *nodes, uint32_t n, cudaMemcpyKind kind)
{
    if (n == 0)
        return fail(ctx, NBVH_ERR_INVALID_ARG, "empty neural BVH");
    if (ctx->nbvh_capacity < n) {
        cudaFree(ctx->d_nbvh);
        cudaFree(ctx->d_nbvh_dilated);
        cudaFree(ctx->d_tree_cut), cudaFree(ctx->d_node_map); // node-indexed arrays share one capacity
        ctx->d_tree_cut = nullptr, ctx->d_node_map = nullptr;
        ctx->d_nbvh = ctx->d_nbvh_dilated = nullptr;
        ctx->nbvh_capacity = 0;
        CU_TRY(ctx, cudaMalloc(&ctx->d_nbvh, (size_t)n * 32));
        CU_TRY(ctx, cudaMalloc(&ctx->d_nbvh_dilated, (size_t)n * 32));
        ctx->nbvh_capacity = n;
    }
    ctx->nbvh_nodes     = n;
    ctx->tree_cut_ready = false; // a new tree-cut invalidates the learning state (reinitialise_neural_bvh_learning_states)
    CU_TRY(ctx, cudaMemcpy(ctx->d_nbvh, nodes, (size_t)n * 32, kind));
    ctx->dilated_with = -1.f;
    if (!ctx->scene_bbox_set) {
        // load_neural_bvh takes the CPU BLAS root box, which is the box of tree-cut node 0
        // (neural_bvh_renderer_accel.cpp:40,248-251)
        nbvh_node32 root;
        CU_TRY(ctx, cudaMemcpy(&root, ctx->d_nbvh, 32, cudaMemcpyDeviceToHost));
        memcpy(ctx->scene_bbox.min, root.aabb_min, 12);
        memcpy(ctx->scene_bbox.max, root.aabb_max, 12);
    }
    return refresh_dilated(ctx);
}

int nbvh_set_neural_bvh(nbvh_ctx *ctx, const nbvh_node32 *nodes, uint32_t n)
{
    if (!ctx || !nodes)
        return NBVH_ERR_INVALID_ARG;
    return adopt_nbvh(ctx, nodes, n, cudaMemcpyDeviceToDevice);
}

// .nbvh = bvh::v2::Bvh<Node<float,3,64,32>>::serialize: u64 n_nodes, u64 n_prim_ids, n_nodes x {6 f32 bounds
// (minx,maxx,miny,maxy,minz,maxz), u64 first_id, u64 prim_count}, n_prim_ids x u64
// (/root/reference/ext/bvh/src/bvh/v2/bvh.h:133-153, node.h:119-133)
int nbvh_load_neural_bvh_file(nbvh_ctx *ctx, const char *path)
{
    if (!ctx || !path)
        return NBVH_ERR_INVALID_ARG;
    std::string data;
    if (!read_file(path, data))
        return fail(ctx, NBVH_ERR_IO, std::string("cannot read ") + path);
    if (data.size() < 16)
        return fail(ctx, NBVH_ERR_PARSE, ".nbvh too short");
    uint64_t n_nodes, n_prims;
    memcpy(&n_nodes, data.data(), 8);
    memcpy(&n_prims, data.data() + 8, 8);
    if (n_nodes == 0 || n_nodes > 0x7FFFFFFFull || n_prims > data.size() / 8 || data.size() != 16 + n_nodes * 40 + n_prims * 8)
        return fail(ctx, NBVH_ERR_PARSE, ".nbvh size does not match its header");
    std::vector<nbvh_node32> nodes(n_nodes);
    const char *p = data.data() + 16;
    for (uint64_t i = 0; i < n_nodes; i++, p += 40) {
        float b[6];
        uint64_t first, count;
        memcpy(b, p, 24);
        memcpy(&first, p + 24, 8);
        memcpy(&count, p + 32, 8);
        nodes[i].aabb_min[0] = b[0], nodes[i].aabb_max[0] = b[1];
        nodes[i].aabb_min[1] = b[2], nodes[i].aabb_max[1] = b[3];
        nodes[i].aabb_min[2] = b[4], nodes[i].aabb_max[2] = b[5];
        nodes[i].prim_count_or_left  = (uint32_t)count;
        nodes[i].left_or_prim_offset = (uint32_t)first;
    }
    return adopt_nbvh(ctx, nodes.data(), (uint32_t)n_nodes, cudaMemcpyHostToDevice);
}

int nbvh_save_neural_bvh_file(nbvh_ctx *ctx, const char *path)
{
    if (!ctx || !path)
        return NBVH_ERR_INVALID_ARG;
    if (!ctx->d_nbvh)
        return fail(ctx, NBVH_ERR_STATE, "no neural BVH set");
    std::vector<nbvh_node32> nodes(ctx->nbvh_nodes);
    CU_TRY(ctx, cudaMemcpy(nodes.data(), ctx->d_nbvh, (size_t)ctx->nbvh_nodes * 32, cudaMemcpyDeviceToHost));
    std::ofstream f(path, std::ios::binary);
    if (!f)
        return fail(ctx, NBVH_ERR_IO, std::string("cannot write ") + path);
    const uint64_t n = nodes.size(), one = 1, zero = 0;
    f.write((const char *)&n, 8);
    f.write((const char *)&one, 8); // one dummy prim id (neural_bvh_renderer_accel.cpp:186-187)
    for (auto &nd : nodes) {
        const float b[6] = {nd.aabb_min[0], nd.aabb_max[0], nd.aabb_min[1], nd.aabb_max[1], nd.aabb_min[2], nd.aabb_max[2]};
        const uint64_t first = nd.left_or_prim_offset, count = nd.prim_count_or_left;
        f.write((const char *)b, 24);
        f.write((const char *)&first, 8);
        f.write((const char *)&count, 8);
    }
    f.write((const char *)&zero, 8);
    return f ? NBVH_OK : fail(ctx, NBVH_ERR_IO, "write failed");
}

uint32_t nbvh_neural_bvh_num_nodes(const nbvh_ctx *ctx) { return ctx ? ctx->nbvh_nodes : 0; }

int nbvh_get_neural_bvh_host(const nbvh_ctx *ctx, nbvh_node32 *nodes_host, uint32_t n)
{
    if (!ctx || !nodes_host || n != ctx->nbvh_nodes || !ctx->d_nbvh)
        return NBVH_ERR_INVALID_ARG;
    return cudaMemcpy(nodes_host, ctx->d_nbvh, (size_t)n * 32, cudaMemcpyDeviceToHost) == cudaSuccess ? NBVH_OK : NBVH_ERR_CUDA;
}

uint64_t nbvh_n_params(const nbvh_ctx *ctx) { return ctx ? ctx->n_params : 0; }

int nbvh_load_params_host(nbvh_ctx *ctx, const uint16_t *fp16, uint64_t n)
{
    if (!ctx || !fp16)
        return NBVH_ERR_INVALID_ARG;
    if (n != ctx->n_params)
        return fail(ctx, NBVH_ERR_INVALID_ARG,
                    "parameter count mismatch: file has " + std::to_string(n) + ", network needs " +
                        std::to_string(ctx->n_params));
    CU_TRY(ctx, cudaMemcpy(ctx->d_params, fp16, n * sizeof(uint16_t), cudaMemcpyHostToDevice));
    int rc = upload_weight_image(ctx, fp16);
    if (rc != NBVH_OK)
        return rc;
    ctx->params_loaded = true;
    ctx->train_state_valid = false; // fp32 master copy, Adam moments and step counters restart from these parameters
    return NBVH_OK;
}

int nbvh_get_params_host(const nbvh_ctx *ctx, uint16_t *fp16, uint64_t n)
{
    if (!ctx || !fp16 || n != ctx->n_params)
        return NBVH_ERR_INVALID_ARG;
    return cudaMemcpy(fp16, ctx->d_params, n * sizeof(uint16_t), cudaMemcpyDeviceToHost) == cudaSuccess ? NBVH_OK : NBVH_ERR_CUDA;
}

// network_models/*_network.json: {"n_params":N,"params_binary":{"bytes":[...],"subtype":null},"params_type":"__half"}
// (ext/tiny-cuda-nn trainer.h:275-320, gpu_memory_json.h:36-60); "float" payloads are accepted too (trainer.h:292-295)
int nbvh_load_network_file(nbvh_ctx *ctx, const char *path)
{
    if (!ctx || !path)
        return NBVH_ERR_INVALID_ARG;
    std::string text;
    if (!read_file(path, text))
        return fail(ctx, NBVH_ERR_IO, std::string("cannot read ") + path);
    Json root;
    try {
        root = json_parse(text);
    } catch (const std::exception &e) {
        return fail(ctx, NBVH_ERR_PARSE, e.what());
    }
    const Json *bin = root.find("params_binary");
    const Json *bytes = bin ? bin->find("bytes") : nullptr;
    if (!bytes || (bytes->type != Json::NumArray && bytes->type != Json::Array))
        return fail(ctx, NBVH_ERR_PARSE, "params_binary.bytes missing");
    const uint64_t n         = (uint64_t)root.value("n_params", 0.0);
    const std::string type   = root.value("params_type", std::string("__half"));
    const size_t elem        = type == "float" ? 4 : 2;
    if (type != "float" && type != "__half")
        return fail(ctx, NBVH_ERR_PARSE, "unknown params_type " + type);
    if (bytes->size() != n * elem)
        return fail(ctx, NBVH_ERR_PARSE, "params_binary.bytes length does not match n_params");
    std::vector<uint8_t> raw(bytes->size());
    for (size_t i = 0; i < raw.size(); i++)
        raw[i] = (uint8_t)bytes->at_num(i);
    std::vector<uint16_t> half(n);
    if (elem == 2) {
        memcpy(half.data(), raw.data(), raw.size());
    } else {
        for (uint64_t i = 0; i < n; i++) {
            float f;
            memcpy(&f, raw.data() + 4 * i, 4);
            half[i] = f32_to_f16_host(f);
        }
    }
    return nbvh_load_params_host(ctx, half.data(), n);
}

int nbvh_save_network_file(nbvh_ctx *ctx, const char *path)
{
    if (!ctx || !path)
        return NBVH_ERR_INVALID_ARG;
    std::vector<uint16_t> half(ctx->n_params);
    CU_TRY(ctx, cudaMemcpy(half.data(), ctx->d_params, half.size() * 2, cudaMemcpyDeviceToHost));
    FILE *f = fopen(path, "wb");
    if (!f)
        return fail(ctx, NBVH_ERR_IO, std::string("cannot write ") + path);
    fprintf(f, "{\"n_params\":%llu,\"params_binary\":{\"bytes\":[", (unsigned long long)ctx->n_params);
    const uint8_t *b = reinterpret_cast<const uint8_t *>(half.data());
    std::string buf;
    buf.reserve(1 << 20);
    for (size_t i = 0; i < half.size() * 2; i++) {
        if (i)
            buf += ',';
        buf += std::to_string((unsigned)b[i]);
        if (buf.size() > (1 << 20) - 8) {
            fwrite(buf.data(), 1, buf.size(), f);
            buf.clear();
        }
    }
    fwrite(buf.data(), 1, buf.size(), f);
    fprintf(f, "],\"subtype\":null},\"params_type\":\"__half\"}");
    fclose(f);
    return NBVH_OK;
}

// tcnn initial values: grid U(-1e-4, 1e-4) (grid.h:1059-1062), MLP Xavier uniform (fully_fused_mlp.cu:866-891).
// Own splitmix64 stream (tcnn uses pcg32 seed 1337; initial weights are statistically, not bitwise, equivalent).
int nbvh_init_params(nbvh_ctx *ctx, uint64_t seed)
{
    if (!ctx)
        return NBVH_ERR_INVALID_ARG;
    std::vector<uint16_t> p(ctx->n_params);
    uint64_t s = seed;
    auto next  = [&]() {
        s += 0x9E3779B97F4A7C15ull;
        uint64_t z = s;
        z          = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z          = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        return (float)((z >> 40) * (1.0 / 16777216.0)); // [0,1)
    };
    size_t idx = 0;
    for (uint32_t l = 0; l < ctx->n_layers; l++) {
        const uint32_t k = l == 0 ? ctx->in_k : (uint32_t)kNeurons;
        const uint32_t n = l + 1 == ctx->n_layers ? (uint32_t)kOutPadded : (uint32_t)kNeurons;
        const float sc   = sqrtf(6.f / (float)(k + n));
        for (uint32_t i = 0; i < n * k; i++)
            p[idx++] = f32_to_f16_host((next() * 2.f - 1.f) * sc);
    }
    for (; idx < p.size(); idx++)
        p[idx] = f32_to_f16_host((next() * 2.f - 1.f) * 1e-4f);
    return nbvh_load_params_host(ctx, p.data(), p.size());
}

// ---- inference ------------------------------------------------------------------------------------------
int nbvh_set_framebuffer(nbvh_ctx *ctx, float *accum, uint8_t *ldr, uint32_t w, uint32_t h)
{
    if (!ctx || w == 0 || h == 0)
        return NBVH_ERR_INVALID_ARG;
    if (ctx->own_accum)
        cudaFree(ctx->accum);
    if (ctx->own_ldr)
        cudaFree(ctx->ldr);
    ctx->accum = nullptr, ctx->ldr = nullptr, ctx->own_accum = ctx->own_ldr = false;
    ctx->fb_w = w, ctx->fb_h = h;
    const size_t n = (size_t)w * h;
    if (accum) {
        ctx->accum = reinterpret_cast<float4 *>(accum);
    } else {
        CU_TRY(ctx, cudaMalloc(&ctx->accum, n * sizeof(float4)));
        CU_TRY(ctx, cudaMemset(ctx->accum, 0, n * sizeof(float4)));
        ctx->own_accum = true;
    }
    if (ldr) {
        ctx->ldr = reinterpret_cast<uchar4 *>(ldr);
    } else {
        CU_TRY(ctx, cudaMalloc(&ctx->ldr, n * sizeof(uchar4)));
        CU_TRY(ctx, cudaMemset(ctx->ldr, 0, n * sizeof(uchar4)));
        ctx->own_ldr = true;
    }
    cudaFree(ctx->own_hits);
    ctx->own_hits = nullptr;
    ctx->hits     = nullptr;
    return NBVH_OK;
}

int nbvh_set_hit_buffer(nbvh_ctx *ctx, nbvh_hit_record *hits)
{
    if (!ctx)
        return NBVH_ERR_INVALID_ARG;
    ctx->hits = hits;
    return NBVH_OK;
}

int nbvh_query(nbvh_ctx *ctx, const float *in, float *out, uint32_t n, void *stream)
{
    if (!ctx || (n && (!in || !out)))
        return NBVH_ERR_INVALID_ARG;
    int rc = check_ready(ctx, false, false, true);
    if (rc != NBVH_OK)
        return rc;
    const NetConst nc = make_net_const(ctx);
    KernelTimer kt(ctx, (cudaStream_t)stream);
    CU_TRY(ctx, launch_query(nc, in, out, n, ctx->module_const.input_network_dims, ctx->d_counters, ctx->num_sms,
                             (cudaStream_t)stream));
    kt.stop();
    return NBVH_OK;
}

int nbvh_trace_intervals(nbvh_ctx *ctx, const float *o, const float *d, uint32_t n, uint32_t max_intervals, uint32_t *counts,
                         uint32_t *node_idx, float *mint, float *maxt, void *stream)
{
    if (!ctx || (n && (!o || !d || !counts || !node_idx || !mint || !maxt)))
        return NBVH_ERR_INVALID_ARG;
    int rc = check_ready(ctx, false, true, false);
    if (rc == NBVH_OK)
        rc = refresh_dilated(ctx);
    if (rc != NBVH_OK)
        return rc;
    KernelTimer kt(ctx, (cudaStream_t)stream);
    CU_TRY(ctx, launch_trace_intervals(make_nbvh_const(ctx, true), ctx->nbvh_safe ? 1u : 0u, o, d, n, max_intervals, counts, node_idx, mint, maxt,
                                       ctx->d_counters, (cudaStream_t)stream));
    kt.stop();
    return NBVH_OK;
}

int nbvh_blas_intersect(nbvh_ctx *ctx, uint32_t blas_idx, const float *o, const float *d, const float *maxt, uint32_t n,
                        float *t, float *u, float *v, uint32_t *prim, void *stream)
{
    if (!ctx || (n && (!o || !d || !t || !u || !v || !prim)))
        return NBVH_ERR_INVALID_ARG;
    if (!ctx->scene_set || blas_idx >= ctx->scene.n_blas)
        return fail(ctx, NBVH_ERR_STATE, "scene not set or BLAS index out of range");
    KernelTimer kt(ctx, (cudaStream_t)stream);
    CU_TRY(ctx, launch_blas_intersect(ctx->scene.blas[blas_idx], ctx->blas_safe[blas_idx] ? 1u : 0u, o, d, maxt, n, t, u, v, prim, ctx->d_counters,
                                      (cudaStream_t)stream));
    kt.stop();
    return NBVH_OK;
}

static CameraConst make_camera_const(const nbvh_ctx *ctx)
{
    CameraConst cc;
    memcpy(cc.origin, ctx->camera.origin, 12);
    memcpy(cc.du, ctx->camera.du, 12);
    memcpy(cc.dv, ctx->camera.dv, 12);
    memcpy(cc.top_left, ctx->camera.dir_top_left, 12);
    cc.fb_w = ctx->fb_w, cc.fb_h = ctx->fb_h;
    bool safe = ctx->nbvh_safe && ctx->tlas_safe;
    for (uint32_t i = 0; i < ctx->scene.n_blas; i++)
        safe = safe && ctx->blas_safe[i];
    cc.nodes_safe = safe ? 1u : 0u;
    return cc;
}

int nbvh_raygen(nbvh_ctx *ctx, uint32_t first_pixel, uint32_t n, int sample_idx, int sample_offset, float *o, float *d,
                void *stream)
{
    if (!ctx || (n && (!o || !d)))
        return NBVH_ERR_INVALID_ARG;
    if (ctx->fb_w == 0)
        return fail(ctx, NBVH_ERR_STATE, "framebuffer not set");
    CU_TRY(ctx, launch_raygen(make_camera_const(ctx), first_pixel, n, sample_idx, sample_offset, o, d, (cudaStream_t)stream));
    return NBVH_OK;
}

int nbvh_encode_segments(nbvh_ctx *ctx, const float *o, const float *d, const float *mint, const float *maxt, uint32_t n, float *out,
                         void *stream)
{
    if (!ctx || (n && (!o || !d || !mint || !maxt || !out)))
        return NBVH_ERR_INVALID_ARG;
    if (ctx->module_const.input_network_dims != 3u * ctx->module_const.input_segment_points)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "encode_segments: directional inputs are not supported");
    CU_TRY(ctx, launch_encode_segments(make_encode_const(ctx), o, d, mint, maxt, n, out, (cudaStream_t)stream));
    return NBVH_OK;
}

int nbvh_optimisation_data_generation(nbvh_ctx *ctx, uint32_t training_step, uint32_t batch_size, uint32_t log2_bvh_max_traversals,
                                      uint32_t *sampled_node_indices, float current_max_loss, const nbvh_tree_cut_data *tree_cut_data,
                                      float *inputs, float *targets, void *stream)
{
    if (!ctx || !sampled_node_indices || !inputs || !targets)
        return NBVH_ERR_INVALID_ARG;
    if (!ctx->d_nbvh || !ctx->scene_set)
        return fail(ctx, NBVH_ERR_STATE, "optimisation_data_generation: neural BVH and scene must be set");
    const nbvh_renderer_constants &bc = ctx->bvh_const;
    if (batch_size == 0 || (batch_size & (batch_size - 1)) || log2_bvh_max_traversals > 31 || (batch_size >> log2_bvh_max_traversals) == 0)
        return fail(ctx, NBVH_ERR_INVALID_ARG, "batch_size must be a power of two >= 2^log2_bvh_max_traversals");
    if (bc.ellipsoidal_bvh || bc.view_dependent_sampling || bc.leaf_boundary_sampling_ratio != 0.f)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "sampler: ellipsoidal_bvh / view_dependent_sampling / leaf_boundary_sampling_ratio are not implemented");
    if (ctx->pt_const.depth_learning_mode != 0 /* LocalT */ || ctx->module_const.input_segment_direction_dim != 0 ||
        ctx->module_const.segment_distribution == 2 /* Uniform */ || ctx->module_const.apply_encoding_trafo)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "sampler: only LocalT depth targets and equidistant segment points are implemented");
    if (bc.training_node_survival && !tree_cut_data)
        return fail(ctx, NBVH_ERR_INVALID_ARG, "training_node_survival needs tree_cut_data");
    if (bc.neural_bvh_idx >= ctx->blas_desc.size() || ctx->blas_desc[bc.neural_bvh_idx].num_nodes < 1 || !ctx->blas_desc[bc.neural_bvh_idx].nodes)
        return fail(ctx, NBVH_ERR_STATE, "sampler: the scene has no classical BLAS for the neural mesh");
    if ((uint32_t)bc.current_learned_lod >= 7u)
        return fail(ctx, NBVH_ERR_INVALID_ARG, "current_learned_lod out of range");
    const nbvh_cuda_blas &nb = ctx->blas_desc[bc.neural_bvh_idx];
    SamplerParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.nbvh.nodes      = ctx->d_nbvh;
    sp.nbvh.num_nodes  = ctx->nbvh_nodes;
    sp.nbvh.mint_eps   = bc.learning_mint_epsilon;
    sp.nbvh.maxt_eps   = bc.learning_maxt_epsilon;
    sp.nbvh.predilated = 0u;
    sp.blas            = ctx->scene.blas[bc.neural_bvh_idx];
    sp.attr.normals       = nb.normals;
    sp.attr.texcoords     = nb.texcoords;
    sp.attr.material_ids  = nb.material_ids;
    sp.attr.material_map  = nb.material_map;
    sp.attr.num_materials = nb.num_materials;
    if (!nb.material_map || nb.num_materials == 0)
        sp.attr.material_ids = nullptr; // no material table: every hit takes the default material
    sp.nodes_safe = 0u; // the raw tree-cut is dilated per sample: every quotient goes through div.rn
    sp.materials  = reinterpret_cast<const MaterialPod *>(ctx->materials);
    sp.textures   = reinterpret_cast<const cudaTextureObject_t *>(ctx->textures);
    sp.enc        = make_encode_const(ctx);
    for (int i = 0; i < 3; i++) {
        sp.bbox_min[i] = ctx->scene_bbox.min[i];
        sp.bbox_ext[i] = ctx->scene_bbox.max[i] - ctx->scene_bbox.min[i];
    }
    sp.sampling_bbox_scaling      = bc.sampling_bbox_scaling;
    sp.fixed_node_dilation        = bc.fixed_node_dilation;
    sp.node_extent_dilation       = bc.node_extent_dilation;
    sp.current_learned_lod        = (uint32_t)bc.current_learned_lod;
    sp.num_bvh_lods               = 7u; // NUM_BVH_LODS, neural_bvh2.cuh:14
    sp.training_node_survival     = bc.training_node_survival ? 1u : 0u;
    sp.survival_probability_scale = bc.survival_probability_scale;
    sp.min_rr_survival_prob       = bc.min_rr_survival_prob;
    sp.max_training_bounces       = bc.max_training_bounces;
    sp.backface_bounce_training   = bc.backface_bounce_training ? 1u : 0u;
    sp.sample_cosine_weighted     = bc.sample_cosine_weighted ? 1u : 0u;
    sp.n_in        = ctx->module_const.input_network_dims;
    sp.n_out       = ctx->pt_const.output_network_dims;
    sp.vis_dim     = ctx->pt_const.output_visibility_dim;
    sp.depth_dim   = ctx->pt_const.output_depth_dim;
    sp.normal_dim  = ctx->pt_const.output_normal_dim;
    sp.albedo_dim  = ctx->pt_const.output_albedo_dim;
    sp.depth_linear = ctx->pt_const.output_depth_linear_activation ? 1u : 0u;
    if (sp.depth_dim > 1u || (sp.normal_dim != 0u && sp.normal_dim != 3u) || (sp.albedo_dim != 0u && sp.albedo_dim != 3u) ||
        sp.vis_dim + sp.depth_dim + sp.normal_dim + sp.albedo_dim != sp.n_out)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "sampler: unsupported output layout");
    sp.sub_batch_size = (int)(batch_size >> log2_bvh_max_traversals);
    sp.n_sub_batch    = (int)(1u << log2_bvh_max_traversals);
    sp.rng_index      = training_step * batch_size;
    sp.sampled_node_indices = sampled_node_indices;
    sp.inputs           = inputs;
    sp.targets          = targets;
    sp.current_max_loss = current_max_loss;
    sp.tree_cut         = tree_cut_data;
    sp.counters         = ctx->d_counters;
    if (sp.n_sub_batch == 1) {
        if (ctx->sampler_item_capacity < batch_size) {
            CU_TRY(ctx, cudaStreamSynchronize((cudaStream_t)stream)); // an earlier generation may still read the old list
            cudaFree(ctx->d_sampler_items);
            ctx->d_sampler_items = nullptr, ctx->sampler_item_capacity = 0;
            CU_TRY(ctx, cudaMalloc(&ctx->d_sampler_items, (size_t)batch_size * 48u));
            ctx->sampler_item_capacity = batch_size;
        }
        if (ctx->node_keep_capacity < ctx->nbvh_nodes) {
            CU_TRY(ctx, cudaStreamSynchronize((cudaStream_t)stream));
            cudaFree(ctx->d_node_keep);
            ctx->d_node_keep = nullptr, ctx->node_keep_capacity = 0;
            const uint32_t cap = ctx->nbvh_capacity > ctx->nbvh_nodes ? ctx->nbvh_capacity : ctx->nbvh_nodes;
            CU_TRY(ctx, cudaMalloc(&ctx->d_node_keep, (size_t)cap * sizeof(float)));
            ctx->node_keep_capacity = cap;
        }
        sp.node_keep  = ctx->d_node_keep;
        sp.items      = ctx->d_sampler_items;
        sp.item_count = ctx->d_next_chunk ? ctx->d_next_chunk + 7 : nullptr;
    }
    CU_TRY(ctx, launch_generate_batch(sp, ctx->d_next_chunk ? ctx->d_next_chunk + 4 : nullptr, ctx->num_sms, (cudaStream_t)stream));
    return NBVH_OK;
}

int nbvh_create_texture_rgba8(nbvh_ctx *ctx, const uint8_t *pixels_host, uint32_t w, uint32_t h, uint64_t *tex_out)
{
    if (!ctx || !pixels_host || !tex_out || w == 0 || h == 0)
        return NBVH_ERR_INVALID_ARG;
    cudaChannelFormatDesc cd = cudaCreateChannelDesc<uchar4>();
    cudaArray_t arr = nullptr;
    CU_TRY(ctx, cudaMallocArray(&arr, &cd, w, h));
    CU_TRY(ctx, cudaMemcpy2DToArray(arr, 0, 0, pixels_host, (size_t)w * 4, (size_t)w * 4, h, cudaMemcpyHostToDevice));
    cudaResourceDesc rd = {};
    rd.resType          = cudaResourceTypeArray;
    rd.res.array.array  = arr;
    cudaTextureDesc td  = {};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeWrap;
    td.filterMode          = cudaFilterModeLinear;
    td.readMode            = cudaReadModeNormalizedFloat;
    td.normalizedCoords    = 1;
    td.maxAnisotropy       = 1;
    td.maxMipmapLevelClamp = 99;
    td.minMipmapLevelClamp = 0;
    td.mipmapFilterMode    = cudaFilterModePoint;
    td.borderColor[0]      = 1.0f;
    td.sRGB                = 0;
    cudaTextureObject_t tex = 0;
    CU_TRY(ctx, cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    ctx->own_textures.emplace_back(tex, arr);
    *tex_out = (uint64_t)tex;
    return NBVH_OK;
}

int nbvh_create_envmap_rgba32f(nbvh_ctx *ctx, const float *pixels_host, uint32_t w, uint32_t h, uint64_t *tex_out)
{
    if (!ctx || !pixels_host || !tex_out || w == 0 || h == 0)
        return NBVH_ERR_INVALID_ARG;
    // CudaBackend::load_envmap (cuda_backend.cpp:474-500): float4 array, linear filter, normalized coordinates, clamp
    cudaChannelFormatDesc cd = cudaCreateChannelDesc(32, 32, 32, 32, cudaChannelFormatKindFloat);
    cudaArray_t arr = nullptr;
    CU_TRY(ctx, cudaMallocArray(&arr, &cd, w, h));
    CU_TRY(ctx, cudaMemcpy2DToArray(arr, 0, 0, pixels_host, (size_t)w * 16, (size_t)w * 16, h, cudaMemcpyHostToDevice));
    cudaResourceDesc rd = {};
    rd.resType          = cudaResourceTypeArray;
    rd.res.array.array  = arr;
    cudaTextureDesc td  = {};
    td.filterMode       = cudaFilterModeLinear;
    td.normalizedCoords = 1;
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex = 0;
    CU_TRY(ctx, cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    ctx->own_textures.emplace_back(tex, arr);
    *tex_out = (uint64_t)tex;
    return NBVH_OK;
}

int nbvh_set_envmap(nbvh_ctx *ctx, uint64_t envmap_texture)
{
    if (!ctx)
        return NBVH_ERR_INVALID_ARG;
    ctx->envmap = (cudaTextureObject_t)envmap_texture;
    return NBVH_OK;
}

int nbvh_destroy_textures(nbvh_ctx *ctx)
{
    if (!ctx)
        return NBVH_ERR_INVALID_ARG;
    for (auto &t : ctx->own_textures) {
        cudaDestroyTextureObject(t.first);
        cudaFreeArray(t.second);
    }
    ctx->own_textures.clear();
    ctx->envmap = 0;
    return NBVH_OK;
}

static int render_impl(nbvh_ctx *ctx, uint32_t accumulated_spp, uint32_t sample_offset, uint32_t first_pixel, uint32_t pixel_count,
                       uint32_t chunk_stride, uint32_t chunk_offset, void *stream, uint32_t spp_count = 1u);

int nbvh_render(nbvh_ctx *ctx, uint32_t accumulated_spp, uint32_t sample_offset, uint32_t first_pixel, uint32_t pixel_count,
                void *stream)
{
    return render_impl(ctx, accumulated_spp, sample_offset, first_pixel, pixel_count, 1u, 0u, stream);
}

int nbvh_render_interleaved(nbvh_ctx *ctx, uint32_t accumulated_spp, uint32_t sample_offset, uint32_t shard_index, uint32_t shard_count,
                            void *stream)
{
    if (!ctx || shard_count == 0 || shard_index >= shard_count)
        return NBVH_ERR_INVALID_ARG;
    const uint32_t total = ctx->fb_w * ctx->fb_h;
    if ((total & 31u) != 0u)
        return fail(ctx, NBVH_ERR_INVALID_ARG, "interleaved shards need a pixel count that is a multiple of 32");
    const uint32_t n_chunks = total >> 5;
    const uint32_t mine     = n_chunks / shard_count + (shard_index < n_chunks % shard_count ? 1u : 0u);
    if (mine == 0u)
        return NBVH_OK;
    return render_impl(ctx, accumulated_spp, sample_offset, 0u, mine * 32u, shard_count, shard_index, stream);
}

// Several samples per pixel in ONE launch sequence (multi-bounce mode): n_spp samples starting at first_spp for the interleaved
// shard (shard_count 1 = the whole frame).  The path lists hold pixels x n_spp slots, so small shards / frames still fill the
// persistent kernels; samples of a pixel accumulate with 16-byte reductions (fp32 sum order differs from sample-by-sample
// accumulation), the 8-bit frame is resolved once at the end.  first_spp == 0 clears the shard's accumulation first.
int nbvh_render_samples(nbvh_ctx *ctx, uint32_t first_spp, uint32_t n_spp, uint32_t sample_offset, uint32_t shard_index,
                        uint32_t shard_count, void *stream)
{
    if (!ctx || shard_count == 0 || shard_index >= shard_count || n_spp == 0)
        return NBVH_ERR_INVALID_ARG;
    if (n_spp == 1u)
        return nbvh_render_interleaved(ctx, first_spp, sample_offset, shard_index, shard_count, stream);
    if (ctx->mega_kernel)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "nbvh_render_samples: not available in NBVH_RENDER_MODE=mega");
    if (ctx->scene_const.ema_accumulate)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "nbvh_render_samples: ema_accumulate needs sample-by-sample accumulation");
    const uint32_t total = ctx->fb_w * ctx->fb_h;
    if ((total & 31u) != 0u)
        return fail(ctx, NBVH_ERR_INVALID_ARG, "multi-sample launches need a pixel count that is a multiple of 32");
    const uint32_t n_chunks = total >> 5;
    const uint32_t mine     = n_chunks / shard_count + (shard_index < n_chunks % shard_count ? 1u : 0u);
    if (mine == 0u)
        return NBVH_OK;
    if ((uint64_t)mine * 32u * n_spp > 0x7FFFFFFFull)
        return fail(ctx, NBVH_ERR_INVALID_ARG, "nbvh_render_samples: pixels x samples exceeds 2^31 path slots");
    return render_impl(ctx, first_spp, sample_offset, 0u, mine * 32u, shard_count, shard_index, stream, n_spp);
}

static int render_impl(nbvh_ctx *ctx, uint32_t accumulated_spp, uint32_t sample_offset, uint32_t first_pixel, uint32_t pixel_count,
                       uint32_t chunk_stride, uint32_t chunk_offset, void *stream, uint32_t spp_count)
{
    if (!ctx)
        return NBVH_ERR_INVALID_ARG;
    int rc = check_ready(ctx, true, true, true);
    if (rc == NBVH_OK)
        rc = refresh_dilated(ctx);
    if (rc != NBVH_OK)
        return rc;
    if (!ctx->accum)
        return fail(ctx, NBVH_ERR_STATE, "framebuffer not set");
    if (ctx->max_depth > kMaxPathDepth)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "max_depth > 64");
    if (ctx->path_const.output_mode != 0)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "only OutputMode::NeuralPathTracing is implemented");
    const uint32_t total = ctx->fb_w * ctx->fb_h;
    if (pixel_count == 0) {
        first_pixel = 0;
        pixel_count = total;
    }
    if (first_pixel >= total || pixel_count > total - first_pixel)
        return fail(ctx, NBVH_ERR_INVALID_ARG, "pixel range outside the framebuffer");
    if ((first_pixel & 31u) != 0u)
        return fail(ctx, NBVH_ERR_INVALID_ARG, "first_pixel must be a multiple of 32");
    cudaStream_t st = (cudaStream_t)stream;
    RenderParams rp;
    memset(&rp, 0, sizeof(rp));
    rp.cam   = make_camera_const(ctx);
    rp.nbvh  = make_nbvh_const(ctx, true);
    rp.scene = ctx->scene;
    rp.enc   = make_encode_const(ctx);
    rp.dec.primary_thr         = ctx->pt_const.primary_visibility_threshold;
    rp.dec.secondary_thr       = ctx->pt_const.secondary_visibility_threshold;
    rp.dec.depth_epsilon       = ctx->pt_const.depth_epsilon;
    rp.dec.fixed_depth_epsilon = ctx->pt_const.fixed_depth_epsilon;
    rp.ignore_neural_blas      = ctx->bvh_const.ignore_neural_blas;
    rp.max_depth               = (uint32_t)ctx->max_depth;
    rp.bg.envmap_eval          = ctx->scene_const.envmap_eval;
    rp.bg.skylike_background   = ctx->scene_const.skylike_background;
    {   // glm::rotate(mat4(1), radians(envmap_rotation), +Y) (path_inference_kernels.cu:45-47)
        const float a  = ctx->scene_const.envmap_rotation * 0.01745329251994329576923690768489f;
        rp.bg.rot_cos  = cosf(a);
        rp.bg.rot_sin  = sinf(a);
        rp.bg.strength = ctx->scene_const.envmap_strength;
        rp.bg.envmap   = ctx->envmap;
    }
    rp.first_pixel = first_pixel, rp.pixel_count = pixel_count;
    rp.chunk_stride = chunk_stride, rp.chunk_offset = chunk_offset;
    rp.spp_count     = spp_count;
    const bool alt = ctx->render_slot >= 1 && ctx->render_slot < kPipeSlots;
    const int ai   = alt ? ctx->render_slot - 1 : 0;
    if (alt && !ctx->d_next_chunk_alt[ai]) {
        CU_TRY(ctx, cudaMalloc(&ctx->d_next_chunk_alt[ai], 8 * sizeof(uint32_t)));
        CU_TRY(ctx, cudaMemset(ctx->d_next_chunk_alt[ai], 0, 8 * sizeof(uint32_t)));
    }
    uint32_t *chunk_ctr        = alt ? ctx->d_next_chunk_alt[ai] : ctx->d_next_chunk;
    NeuralQueueEntry *&queue   = alt ? ctx->d_queue_alt[ai] : ctx->d_queue;
    size_t &queue_capacity     = alt ? ctx->queue_capacity_alt[ai] : ctx->queue_capacity;
    rp.next_chunk    = chunk_ctr;
    const size_t slots = spp_count > 1u ? (size_t)pixel_count * spp_count : (size_t)total; // path / queue slots this launch may use
    if (queue_capacity < slots) {
        CU_TRY(ctx, cudaDeviceSynchronize());
        cudaFree(queue);
        queue = nullptr, queue_capacity = 0;
        CU_TRY(ctx, cudaMalloc(&queue, slots * sizeof(NeuralQueueEntry)));
        queue_capacity = slots;
    }
    rp.queue = queue, rp.queue_count = chunk_ctr + 1, rp.queue_cursor = chunk_ctr + 2;
    rp.queue_count_back = chunk_ctr + 3; // zeroed with the other counters below
    rp.queue_capacity   = (uint32_t)queue_capacity;
    rp.sample_idx    = (int)accumulated_spp;
    rp.sample_offset = (int)sample_offset;
    rp.accum = ctx->accum, rp.ldr = ctx->ldr, rp.hits = ctx->hits;
    rp.ema_accumulate = ctx->scene_const.ema_accumulate;
    rp.ema_weight     = ctx->scene_const.ema_weight;
    rp.counters       = ctx->d_counters;
    const NetConst nc = make_net_const(ctx);
    CU_TRY(ctx, cudaMemsetAsync(chunk_ctr, 0, 4 * sizeof(uint32_t), st));
    KernelTimer kt(ctx, st);
    if (ctx->l2_prefetch) {
        // warm L2 with the tables the scattered gathers of this frame touch (see k_prefetch_l2)
        const void *ptrs[12];
        size_t bytes[12];
        int n = 0;
        ptrs[n] = ctx->d_params, bytes[n++] = (size_t)ctx->n_params * 2;
        ptrs[n] = ctx->d_nbvh_dilated, bytes[n++] = (size_t)ctx->nbvh_nodes * 32;
        for (size_t i = 0; i < ctx->blas_desc.size() && n + 3 <= 12; i++) {
            const nbvh_cuda_blas &b = ctx->blas_desc[i];
            if ((i == ctx->bvh_const.neural_bvh_idx && !ctx->bvh_const.ignore_neural_blas) || b.num_nodes < 2)
                continue; // the neural mesh's classical BLAS is not traversed at inference
            if ((size_t)b.num_nodes * 32 > (32u << 20))
                continue; // large meshes stream through L2 anyway
            ptrs[n] = b.nodes, bytes[n++] = (size_t)b.num_nodes * 32;
            ptrs[n] = b.prim_indices, bytes[n++] = (size_t)b.num_prim_indices * 8;
            ptrs[n] = b.indices, bytes[n++] = (size_t)b.num_prim_indices * 12;
        }
        CU_TRY(ctx, launch_prefetch_l2(ptrs, bytes, n, ctx->num_sms, st));
    }
    if (ctx->max_depth > 1) {
        // hybrid neural path tracing (path.cu): shading + re-queue per bounce, no host synchronisation
        if (ctx->path_capacity < slots) {
            ctx->path_capacity = 0;
            for (int k = 0; k < 2; k++) {
                cudaFree(ctx->d_paths[k]), cudaFree(ctx->d_path_hits[k]);
                ctx->d_paths[k] = nullptr, ctx->d_path_hits[k] = nullptr;
                CU_TRY(ctx, cudaMalloc(&ctx->d_paths[k], slots * sizeof(PathState)));
                CU_TRY(ctx, cudaMalloc(&ctx->d_path_hits[k], slots * sizeof(PathHit)));
            }
            ctx->path_capacity = slots;
        }
        const size_t counter_bytes = kPathCountersPerDepth * (kMaxPathDepth + 1) * sizeof(uint32_t);
        if (!ctx->d_bounce_counters)
            CU_TRY(ctx, cudaMalloc(&ctx->d_bounce_counters, counter_bytes));
        CU_TRY(ctx, cudaMemsetAsync(ctx->d_bounce_counters, 0, counter_bytes, st));
        PathConst pc;
        memset(&pc, 0, sizeof(pc));
        for (size_t i = 0; i < ctx->blas_desc.size() && i < (size_t)kMaxBlas; i++) {
            const nbvh_cuda_blas &b = ctx->blas_desc[i];
            pc.attr[i].normals       = b.normals;
            pc.attr[i].texcoords     = b.texcoords;
            pc.attr[i].material_ids  = (b.material_map && b.num_materials) ? b.material_ids : nullptr;
            pc.attr[i].material_map  = b.material_map;
            pc.attr[i].num_materials = b.material_map ? b.num_materials : 0u;
        }
        pc.materials = reinterpret_cast<const MaterialPod *>(ctx->materials);
        pc.textures  = reinterpret_cast<const cudaTextureObject_t *>(ctx->textures);
        pc.neural_bvh_idx          = ctx->bvh_const.neural_bvh_idx;
        pc.normal_epsilon_offset   = ctx->path_const.normal_epsilon_offset;
        pc.apply_russian_roulette  = ctx->path_const.apply_russian_roulette ? 1u : 0u;
        pc.rr_start_depth          = (uint32_t)ctx->path_const.rr_start_depth;
        pc.neural_roughness        = ctx->path_const.neural_roughness;
        pc.neural_specular         = ctx->path_const.neural_specular;
        pc.neural_metallic         = ctx->path_const.neural_metallic;
        pc.force_constant_material = ctx->scene_const.force_constant_material ? 1u : 0u;
        pc.emission_eval           = ctx->scene_const.emission_eval ? 1u : 0u;
        pc.constant_roughness      = ctx->scene_const.constant_roughness;
        pc.constant_specular       = ctx->scene_const.constant_specular;
        pc.constant_metallic       = ctx->scene_const.constant_metallic;
        NbvhConst nbvh_bounce = rp.nbvh;
        if (ctx->bvh_const.path_lod_switch && ctx->bvh_const.num_lods_learned > 0)
            nbvh_bounce.leaf_lod_max = 2u + (uint32_t)ctx->bvh_const.lod_switch;
        CU_TRY(ctx, launch_render_path(rp, pc, nc, nbvh_bounce, ctx->d_paths, ctx->d_path_hits, ctx->d_bounce_counters, ctx->num_sms, st));
    } else if (ctx->mega_kernel) {
        CU_TRY(ctx, launch_render(rp, nc, ctx->num_sms, st));
    } else if (3u * (ctx->stage_used + 1u) <= ctx->stage_ev.size()) {
        cudaEvent_t *e = ctx->stage_ev.data() + 3u * ctx->stage_used;
        CU_TRY(ctx, cudaEventRecord(e[0], st));
        CU_TRY(ctx, launch_render_split(rp, nc, ctx->num_sms, st, e[1]));
        CU_TRY(ctx, cudaEventRecord(e[2], st));
        ctx->stage_used++;
    } else {
        CU_TRY(ctx, launch_render_split(rp, nc, ctx->num_sms, st, nullptr));
    }
    kt.stop();
    return NBVH_OK;
}

int nbvh_trace_rays(nbvh_ctx *ctx, const float *o, const float *d, uint32_t n, uint32_t ray_depth, nbvh_hit_record *hits, void *stream)
{
    if (!ctx || (n && (!o || !d || !hits)) || ray_depth == 0)
        return NBVH_ERR_INVALID_ARG;
    int rc = check_ready(ctx, true, true, true);
    if (rc == NBVH_OK)
        rc = refresh_dilated(ctx);
    if (rc != NBVH_OK)
        return rc;
    if (n == 0)
        return NBVH_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (ctx->queue_capacity < (size_t)n) {
        cudaFree(ctx->d_queue);
        ctx->d_queue = nullptr, ctx->queue_capacity = 0;
        CU_TRY(ctx, cudaMalloc(&ctx->d_queue, (size_t)n * sizeof(NeuralQueueEntry)));
        ctx->queue_capacity = n;
    }
    if (ctx->path_capacity < (size_t)n) {
        ctx->path_capacity = 0;
        for (int k = 0; k < 2; k++) {
            cudaFree(ctx->d_paths[k]), cudaFree(ctx->d_path_hits[k]);
            ctx->d_paths[k] = nullptr, ctx->d_path_hits[k] = nullptr;
            CU_TRY(ctx, cudaMalloc(&ctx->d_paths[k], (size_t)n * sizeof(PathState)));
            CU_TRY(ctx, cudaMalloc(&ctx->d_path_hits[k], (size_t)n * sizeof(PathHit)));
        }
        ctx->path_capacity = n;
    }
    const size_t counter_bytes = kPathCountersPerDepth * (kMaxPathDepth + 1) * sizeof(uint32_t);
    if (!ctx->d_bounce_counters)
        CU_TRY(ctx, cudaMalloc(&ctx->d_bounce_counters, counter_bytes));
    RenderParams rp;
    memset(&rp, 0, sizeof(rp));
    rp.cam   = make_camera_const(ctx); // only nodes_safe is read
    rp.nbvh  = make_nbvh_const(ctx, true);
    if (ray_depth > 1u && ctx->bvh_const.path_lod_switch && ctx->bvh_const.num_lods_learned > 0)
        rp.nbvh.leaf_lod_max = 2u + (uint32_t)ctx->bvh_const.lod_switch;
    rp.scene = ctx->scene;
    rp.enc   = make_encode_const(ctx);
    rp.dec.primary_thr         = ctx->pt_const.primary_visibility_threshold;
    rp.dec.secondary_thr       = ctx->pt_const.secondary_visibility_threshold;
    rp.dec.depth_epsilon       = ctx->pt_const.depth_epsilon;
    rp.dec.fixed_depth_epsilon = ctx->pt_const.fixed_depth_epsilon;
    rp.ignore_neural_blas      = ctx->bvh_const.ignore_neural_blas;
    rp.max_depth               = ray_depth;
    rp.queue                   = ctx->d_queue;
    rp.counters                = ctx->d_counters;
    const NetConst nc = make_net_const(ctx);
    KernelTimer kt(ctx, st);
    CU_TRY(ctx, launch_trace_rays(rp, nc, o, d, n, ray_depth, ctx->d_paths[0], ctx->d_path_hits[0], ctx->d_bounce_counters, hits, ctx->num_sms, st));
    kt.stop();
    return NBVH_OK;
}

// ---- tree-cut learning ---------------------------------------------------------------------
namespace {

// grow the node-indexed arrays (nodes, dilated nodes, tree-cut data, node map) to `capacity`, keeping their contents
int grow_tree_cut(nbvh_ctx *ctx, uint32_t capacity)
{
    if (capacity > ctx->nbvh_capacity || !ctx->d_tree_cut) {
        const uint32_t cap = capacity > ctx->nbvh_capacity ? capacity : ctx->nbvh_capacity;
        float4 *nodes = nullptr, *dil = nullptr;
        nbvh_tree_cut_data *cut = nullptr;
        uint32_t *map = nullptr;
        CU_TRY(ctx, cudaMalloc(&nodes, (size_t)cap * 32));
        CU_TRY(ctx, cudaMalloc(&dil, (size_t)cap * 32));
        CU_TRY(ctx, cudaMalloc(&cut, (size_t)cap * sizeof(nbvh_tree_cut_data)));
        CU_TRY(ctx, cudaMalloc(&map, (size_t)cap * 4));
        if (ctx->d_nbvh)
            CU_TRY(ctx, cudaMemcpy(nodes, ctx->d_nbvh, (size_t)ctx->nbvh_nodes * 32, cudaMemcpyDeviceToDevice));
        if (ctx->d_tree_cut)
            CU_TRY(ctx, cudaMemcpy(cut, ctx->d_tree_cut, (size_t)ctx->nbvh_nodes * sizeof(nbvh_tree_cut_data), cudaMemcpyDeviceToDevice));
        if (ctx->d_node_map)
            CU_TRY(ctx, cudaMemcpy(map, ctx->d_node_map, (size_t)ctx->nbvh_nodes * 4, cudaMemcpyDeviceToDevice));
        cudaFree(ctx->d_nbvh), cudaFree(ctx->d_nbvh_dilated), cudaFree(ctx->d_tree_cut), cudaFree(ctx->d_node_map);
        ctx->d_nbvh = nodes, ctx->d_nbvh_dilated = dil, ctx->d_tree_cut = cut, ctx->d_node_map = map;
        ctx->nbvh_capacity = cap;
        ctx->dilated_with  = -1.f;
    }
    const size_t need = tree_cut_workspace_bytes(ctx->nbvh_capacity);
    if (need > ctx->tc_workspace_bytes) {
        cudaFree(ctx->d_tc_workspace);
        ctx->d_tc_workspace = nullptr, ctx->tc_workspace_bytes = 0;
        CU_TRY(ctx, cudaMalloc(&ctx->d_tc_workspace, need));
        ctx->tc_workspace_bytes = need;
    }
    if (!ctx->d_tc_result) {
        CU_TRY(ctx, cudaMalloc(&ctx->d_tc_result, sizeof(TreeCutResult)));
        CU_TRY(ctx, cudaMemset(ctx->d_tc_result, 0, sizeof(TreeCutResult)));
    }
    return NBVH_OK;
}

int neural_blas_nodes_host(nbvh_ctx *ctx, std::vector<nbvh_node32> &out)
{
    const uint32_t idx = ctx->bvh_const.neural_bvh_idx;
    if (!ctx->scene_set || idx >= ctx->blas_desc.size() || ctx->blas_desc[idx].num_nodes < 1 || !ctx->blas_desc[idx].nodes)
        return fail(ctx, NBVH_ERR_STATE, "tree-cut learning needs the classical BLAS of the neural mesh in the scene");
    out.resize(ctx->blas_desc[idx].num_nodes);
    CU_TRY(ctx, cudaMemcpy(out.data(), ctx->blas_desc[idx].nodes, out.size() * sizeof(nbvh_node32), cudaMemcpyDeviceToHost));
    return NBVH_OK;
}

int upload_tree_cut_state(nbvh_ctx *ctx, const std::vector<nbvh_node32> &nodes, const std::vector<nbvh_tree_cut_data> &cut,
                          const std::vector<uint32_t> &map, uint32_t max_depth)
{
    const uint32_t n = (uint32_t)nodes.size();
    ctx->nbvh_nodes  = ctx->nbvh_nodes < n ? ctx->nbvh_nodes : n; // grow_tree_cut copies the first nbvh_nodes entries
    int rc = grow_tree_cut(ctx, n > 4u ? n : 4u);
    if (rc != NBVH_OK)
        return rc;
    CU_TRY(ctx, cudaMemcpy(ctx->d_nbvh, nodes.data(), (size_t)n * 32, cudaMemcpyHostToDevice));
    CU_TRY(ctx, cudaMemcpy(ctx->d_tree_cut, cut.data(), (size_t)n * sizeof(nbvh_tree_cut_data), cudaMemcpyHostToDevice));
    CU_TRY(ctx, cudaMemcpy(ctx->d_node_map, map.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
    TreeCutResult r{0.f, 0u, max_depth, 0u};
    CU_TRY(ctx, cudaMemcpy(ctx->d_tc_result, &r, sizeof(r), cudaMemcpyHostToDevice));
    ctx->nbvh_nodes         = n;
    ctx->tree_cut_max_depth = max_depth;
    ctx->dilated_with       = -1.f;
    ctx->tree_cut_ready     = true;
    return refresh_dilated(ctx);
}

} // namespace

// build_and_upload_neural_bvh + traverse_and_build_neural_bvh (neural_bvh_renderer_accel.cpp:24-74): the initial cut of the
// neural mesh's BLAS at `min_depth` (neural_bvh_min_depth; 0 = the root alone), node numbering as the reference's recursion
int nbvh_tree_cut_init(nbvh_ctx *ctx, uint32_t min_depth)
{
    if (!ctx)
        return NBVH_ERR_INVALID_ARG;
    std::vector<nbvh_node32> blas;
    int rc = neural_blas_nodes_host(ctx, blas);
    if (rc != NBVH_OK)
        return rc;
    const uint32_t first_lod = (uint32_t)ctx->bvh_const.current_learned_lod;
    std::vector<nbvh_node32> nodes(1);
    std::vector<nbvh_tree_cut_data> cut(1);
    std::vector<uint32_t> map(1);
    uint32_t max_depth = 0;
    struct Item { uint32_t bvh, neural, depth; };
    // the reference recurses left subtree first; an explicit stack that pushes right before left visits in the same order
    std::vector<Item> stack{{0u, 0u, 0u}};
    while (!stack.empty()) {
        const Item it = stack.back();
        stack.pop_back();
        const nbvh_node32 &b = blas[it.bvh];
        map[it.neural]       = it.bvh;
        nodes[it.neural]     = b;
        cut[it.neural]       = nbvh_tree_cut_data{0.f, 1u, it.depth, 0u};
        max_depth            = it.depth > max_depth ? it.depth : max_depth;
        const bool blas_leaf = b.prim_count_or_left > 0u;
        if (blas_leaf || it.depth == min_depth) {
            nodes[it.neural].prim_count_or_left  = blas_leaf ? 1u : 2u + first_lod;
            nodes[it.neural].left_or_prim_offset = 0xFFFFFFFFu;
            continue;
        }
        const uint32_t left = (uint32_t)nodes.size();
        nodes.resize(left + 2), cut.resize(left + 2), map.resize(left + 2);
        nodes[it.neural].prim_count_or_left  = 0u;
        nodes[it.neural].left_or_prim_offset = left;
        stack.push_back({b.left_or_prim_offset + 1u, left + 1u, it.depth + 1u});
        stack.push_back({b.left_or_prim_offset, left, it.depth + 1u});
    }
    // NOTE: the reference allocates the children of a node before descending, but descends depth-first, so the children of
    // the LEFT subtree are numbered before those of the right one -- the stack order above reproduces exactly that.
    memcpy(ctx->scene_bbox.min, blas[0].aabb_min, 12);
    memcpy(ctx->scene_bbox.max, blas[0].aabb_max, 12);
    return upload_tree_cut_state(ctx, nodes, cut, map, max_depth);
}

// reinitialise_neural_bvh_learning_states + traverse_neural_bvh_and_update_states (neural_bvh_renderer_accel.cpp:76-150):
// learning state for the tree-cut already in the context (loaded from a .nbvh file): node map, depths, zeroed statistics
int nbvh_tree_cut_attach(nbvh_ctx *ctx)
{
    if (!ctx || !ctx->d_nbvh)
        return ctx ? fail(ctx, NBVH_ERR_STATE, "no neural BVH set") : NBVH_ERR_INVALID_ARG;
    std::vector<nbvh_node32> blas;
    int rc = neural_blas_nodes_host(ctx, blas);
    if (rc != NBVH_OK)
        return rc;
    const uint32_t n = ctx->nbvh_nodes;
    std::vector<nbvh_node32> nodes(n);
    CU_TRY(ctx, cudaMemcpy(nodes.data(), ctx->d_nbvh, (size_t)n * 32, cudaMemcpyDeviceToHost));
    std::vector<nbvh_tree_cut_data> cut(n, nbvh_tree_cut_data{0.f, 1u, 0u, 0u});
    std::vector<uint32_t> map(n, 0u);
    uint32_t max_depth = 0;
    struct Item { uint32_t bvh, neural, depth; };
    std::vector<Item> stack{{0u, 0u, 0u}};
    while (!stack.empty()) {
        const Item it = stack.back();
        stack.pop_back();
        if (it.neural >= n || it.bvh >= blas.size())
            return fail(ctx, NBVH_ERR_STATE, "the tree-cut is not a cut of the neural mesh's BLAS");
        map[it.neural]       = it.bvh;
        cut[it.neural].depth = it.depth;
        max_depth            = it.depth > max_depth ? it.depth : max_depth;
        if (nodes[it.neural].left_or_prim_offset == 0xFFFFFFFFu)
            continue;
        if (blas[it.bvh].prim_count_or_left > 0u)
            return fail(ctx, NBVH_ERR_STATE, "the tree-cut is not a cut of the neural mesh's BLAS");
        stack.push_back({blas[it.bvh].left_or_prim_offset + 1u, nodes[it.neural].left_or_prim_offset + 1u, it.depth + 1u});
        stack.push_back({blas[it.bvh].left_or_prim_offset, nodes[it.neural].left_or_prim_offset, it.depth + 1u});
    }
    return upload_tree_cut_state(ctx, nodes, cut, map, max_depth);
}

const nbvh_tree_cut_data *nbvh_tree_cut_data_device(const nbvh_ctx *ctx) { return ctx && ctx->tree_cut_ready ? ctx->d_tree_cut : nullptr; }
const nbvh_node32 *nbvh_neural_bvh_device(const nbvh_ctx *ctx) { return ctx ? reinterpret_cast<const nbvh_node32 *>(ctx->d_nbvh) : nullptr; }

// test / restore hook: overwrite the per-node statistics (the reference restores them when it reloads a training state)
int nbvh_set_tree_cut_host(nbvh_ctx *ctx, const nbvh_tree_cut_data *cut_host, uint32_t n)
{
    if (!ctx || !cut_host || !ctx->tree_cut_ready || n != ctx->nbvh_nodes)
        return NBVH_ERR_INVALID_ARG;
    CU_TRY(ctx, cudaMemcpy(ctx->d_tree_cut, cut_host, (size_t)n * sizeof(nbvh_tree_cut_data), cudaMemcpyHostToDevice));
    return NBVH_OK;
}

int nbvh_get_tree_cut_host(const nbvh_ctx *ctx, nbvh_tree_cut_data *cut_host, uint32_t *node_map_host, uint32_t n)
{
    if (!ctx || !ctx->tree_cut_ready || n > ctx->nbvh_nodes)
        return NBVH_ERR_INVALID_ARG;
    if (cut_host && cudaMemcpy(cut_host, ctx->d_tree_cut, (size_t)n * sizeof(nbvh_tree_cut_data), cudaMemcpyDeviceToHost) != cudaSuccess)
        return NBVH_ERR_CUDA;
    if (node_map_host && cudaMemcpy(node_map_host, ctx->d_node_map, (size_t)n * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
        return NBVH_ERR_CUDA;
    return NBVH_OK;
}

// The tree-cut part of one optimisation step (neural_bvh_renderer.cu:199-283): scatter the per-element losses of the batch
// into the nodes' statistics, age the leaves, find the num_splits largest loss heuristics, split those leaves.
static int tree_cut_decide(nbvh_ctx *ctx, uint32_t num_splits, float *max_node_loss_host, uint32_t *n_split_host, cudaStream_t st)
{
    const uint32_t n        = ctx->nbvh_nodes;
    const uint32_t blas_idx = ctx->bvh_const.neural_bvh_idx;
    CU_TRY(ctx, launch_tree_cut_decide_and_split(ctx->d_nbvh, n, ctx->nbvh_capacity, ctx->d_tree_cut, ctx->d_node_map,
                                                 ctx->scene.blas[blas_idx].nodes, num_splits, (uint32_t)ctx->bvh_const.current_learned_lod,
                                                 ctx->d_tc_workspace, ctx->tc_workspace_bytes, ctx->d_tc_result, st));
    TreeCutResult r;
    CU_TRY(ctx, cudaMemcpyAsync(&r, ctx->d_tc_result, sizeof(r), cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    if (r.n_split) {
        ctx->nbvh_nodes   = n + 2u * r.n_split;
        ctx->dilated_with = -1.f; // the pre-dilated copy is rebuilt before the next traversal
        int rc = refresh_dilated(ctx);
        if (rc != NBVH_OK)
            return rc;
    }
    ctx->tree_cut_max_depth = r.max_depth;
    if (max_node_loss_host)
        *max_node_loss_host = r.max_node_loss;
    if (n_split_host)
        *n_split_host = r.n_split;
    return NBVH_OK;
}

// room for the splits of one step: m_neural_nodes_alloc_count doubles in the reference; here grow to what is needed, x2
static int tree_cut_reserve(nbvh_ctx *ctx, uint32_t num_splits, cudaStream_t st)
{
    if (!ctx->tree_cut_ready)
        return fail(ctx, NBVH_ERR_STATE, "tree-cut learning state missing: call nbvh_tree_cut_init or nbvh_tree_cut_attach");
    const uint32_t n     = ctx->nbvh_nodes;
    const uint32_t limit = num_splits < (n + 1u) / 2u ? num_splits : (n + 1u) / 2u;
    if (n + 2u * limit > ctx->nbvh_capacity) {
        CU_TRY(ctx, cudaStreamSynchronize(st));
        return grow_tree_cut(ctx, 2u * (n + 2u * limit));
    }
    if (tree_cut_workspace_bytes(ctx->nbvh_capacity) > ctx->tc_workspace_bytes)
        return grow_tree_cut(ctx, ctx->nbvh_capacity);
    return NBVH_OK;
}

int nbvh_tree_cut_step(nbvh_ctx *ctx, const float *per_sample_loss, uint32_t stride, uint32_t dims, const uint32_t *sampled_node_indices,
                       uint32_t batch, uint32_t num_splits, float *max_node_loss_host, uint32_t *n_split_host, void *stream)
{
    if (!ctx || !per_sample_loss || !sampled_node_indices || dims > stride)
        return NBVH_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = tree_cut_reserve(ctx, num_splits, st);
    if (rc != NBVH_OK)
        return rc;
    CU_TRY(ctx, launch_tree_cut_accumulate(per_sample_loss, stride, dims, sampled_node_indices, batch, ctx->d_tree_cut, st));
    return tree_cut_decide(ctx, num_splits, max_node_loss_host, n_split_host, st);
}

// Data-parallel form of the same step: every rank scatters the losses of ITS batch shard into two node-indexed
// delta arrays (caller-owned, num_nodes entries each), the host allreduces them (ncclAllReduce sum), and every rank applies the
// summed deltas and runs the deterministic decision + split -> identical tree-cuts on all ranks.
int nbvh_tree_cut_accumulate(nbvh_ctx *ctx, const float *per_sample_loss, uint32_t stride, uint32_t dims, const uint32_t *sampled_node_indices,
                             uint32_t batch, float *loss_delta, uint32_t *samples_delta, void *stream)
{
    if (!ctx || !per_sample_loss || !sampled_node_indices || !loss_delta || !samples_delta || dims > stride)
        return NBVH_ERR_INVALID_ARG;
    if (!ctx->tree_cut_ready)
        return fail(ctx, NBVH_ERR_STATE, "tree-cut learning state missing: call nbvh_tree_cut_init or nbvh_tree_cut_attach");
    CU_TRY(ctx, launch_tree_cut_accumulate_delta(per_sample_loss, stride, dims, sampled_node_indices, batch, ctx->nbvh_nodes, loss_delta,
                                                 samples_delta, (cudaStream_t)stream));
    return NBVH_OK;
}

int nbvh_tree_cut_decide(nbvh_ctx *ctx, const float *loss_delta, const uint32_t *samples_delta, uint32_t num_splits, float *max_node_loss_host,
                         uint32_t *n_split_host, void *stream)
{
    if (!ctx || !loss_delta || !samples_delta)
        return NBVH_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = tree_cut_reserve(ctx, num_splits, st);
    if (rc != NBVH_OK)
        return rc;
    CU_TRY(ctx, launch_tree_cut_apply_delta(ctx->d_tree_cut, ctx->nbvh_nodes, loss_delta, samples_delta, st));
    return tree_cut_decide(ctx, num_splits, max_node_loss_host, n_split_host, st);
}

// assign_lod_to_current_neural_leaf_nodes (neural_bvh_renderer_accel.cpp:394-441): tags the current leaves with the level
// being learned and moves on to the next finer level.  Returns the new current_learned_lod through the constant block.
int nbvh_tree_cut_assign_lod(nbvh_ctx *ctx, int *all_lods_assigned)
{
    if (!ctx || !ctx->d_nbvh)
        return NBVH_ERR_INVALID_ARG;
    const int level = ctx->bvh_const.current_learned_lod;
    if (level < 0 || level >= 7)
        return fail(ctx, NBVH_ERR_INVALID_ARG, "current_learned_lod out of range");
    if (level != 6) // at the coarsest level the leaves already carry LEAF_LOD(6) (NUM_BVH_LODS - 1)
        CU_TRY(ctx, launch_tree_cut_assign_lod(ctx->d_nbvh, ctx->nbvh_nodes, (uint32_t)level, 0));
    const bool done = level == 0;
    if (!done)
        ctx->bvh_const.current_learned_lod--;
    if (all_lods_assigned)
        *all_lods_assigned = done ? 1 : 0;
    ctx->dilated_with = -1.f;
    return refresh_dilated(ctx);
}

uint32_t nbvh_tree_cut_max_depth(const nbvh_ctx *ctx) { return ctx ? ctx->tree_cut_max_depth : 0; }

// == compute_and_display_render_error + the mean (neural_bvh_renderer.cu:346-358): error of the accumulated image (accum / spp)
// against a reference RGBA32F image on the device.  mode = PostprocessMode: 3 AbsError, 4 RelAbsError, 5 MSEError, 6 RelMSEError.
int nbvh_image_error(nbvh_ctx *ctx, const float *reference_rgba32f, uint32_t accumulated_spp, int mode, int write_error_image,
                     float *mean_error_host, void *stream)
{
    if (!ctx || !reference_rgba32f || !mean_error_host || mode < 3 || mode > 6 || accumulated_spp == 0)
        return NBVH_ERR_INVALID_ARG;
    if (!ctx->accum)
        return fail(ctx, NBVH_ERR_STATE, "framebuffer not set");
    const uint32_t n = ctx->fb_w * ctx->fb_h;
    float *sum_dev   = reinterpret_cast<float *>(ctx->d_next_chunk + 5);
    cudaStream_t st  = (cudaStream_t)stream;
    CU_TRY(ctx, launch_image_error(ctx->accum, reinterpret_cast<const float4 *>(reference_rgba32f), n, accumulated_spp, mode, sum_dev,
                                   write_error_image ? ctx->ldr : nullptr, st));
    float sum = 0.f;
    CU_TRY(ctx, cudaMemcpyAsync(&sum, sum_dev, sizeof(float), cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    *mean_error_host = sum / (float)n;
    return NBVH_OK;
}

// Device-side BLAS build (f3): LBVH over n_tris triangles into caller-owned arrays: nodes_out (2 * n_tris - 1 BVH2Node records),
// prim_indices_out (n_tris size_t entries).  Replaces the host ext/bvh build of CudaBackend (cuda_backend.cpp:186-385) for scenes
// that do not need the ext/bvh tree (shipped .nbvh files are cuts of that tree).  The traversal kernels keep a 32-entry stack
// like the reference (bvh2.cuh:13-88): a tree deeper than that is refused.
int nbvh_build_blas(nbvh_ctx *ctx, const float *vertices, const uint32_t *indices, uint32_t n_tris, nbvh_node32 *nodes_out,
                    uint64_t *prim_indices_out, uint32_t *max_depth_host, void *stream)
{
    if (!ctx || !vertices || !indices || !nodes_out || !prim_indices_out || n_tris == 0 || n_tris > (1u << 30))
        return NBVH_ERR_INVALID_ARG;
    cudaStream_t st    = (cudaStream_t)stream;
    const size_t bytes = build_blas_workspace_bytes(n_tris);
    if (ctx->build_ws_bytes < bytes) { // grow-only scratch: cudaMalloc / cudaFree of ~60 B per triangle cost far more than the build
        cudaFree(ctx->d_build_ws);
        ctx->d_build_ws = nullptr, ctx->build_ws_bytes = 0;
        CU_TRY(ctx, cudaMalloc(&ctx->d_build_ws, bytes));
        ctx->build_ws_bytes = bytes;
    }
    void *ws            = ctx->d_build_ws;
    uint32_t *depth_dev = ctx->d_next_chunk + 6;
    cudaError_t e = launch_build_blas(vertices, indices, n_tris, reinterpret_cast<float4 *>(nodes_out),
                                      reinterpret_cast<unsigned long long *>(prim_indices_out), ws, bytes, depth_dev, st);
    uint32_t depth = 0;
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(&depth, depth_dev, 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(st);
    if (e != cudaSuccess)
        return fail(ctx, NBVH_ERR_CUDA, std::string("nbvh_build_blas: ") + cudaGetErrorString(e));
    if (max_depth_host)
        *max_depth_host = depth;
    if (depth > (uint32_t)kBlasStackSize)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "nbvh_build_blas: tree deeper than the traversal stack (32)");
    return NBVH_OK;
}

int nbvh_render_host(nbvh_ctx *ctx, uint32_t accumulated_spp, uint32_t sample_offset, const nbvh_camera_params *camera_host,
                     uint8_t *ldr_host, nbvh_hit_record *hits_host)
{
    if (!ctx)
        return NBVH_ERR_INVALID_ARG;
    if (camera_host)
        ctx->camera = *camera_host; // the per-frame constant upload of the reference (setup_constants)
    const size_t n = (size_t)ctx->fb_w * ctx->fb_h;
    if (hits_host && !ctx->hits) {
        if (!ctx->own_hits)
            CU_TRY(ctx, cudaMalloc(&ctx->own_hits, n * sizeof(nbvh_hit_record)));
        ctx->hits = ctx->own_hits;
    }
    int rc = nbvh_render(ctx, accumulated_spp, sample_offset, 0, 0, nullptr);
    if (rc != NBVH_OK)
        return rc;
    if (ldr_host)
        CU_TRY(ctx, cudaMemcpyAsync(ldr_host, ctx->ldr, n * sizeof(uchar4), cudaMemcpyDeviceToHost, 0));
    if (hits_host)
        CU_TRY(ctx, cudaMemcpyAsync(hits_host, ctx->hits, n * sizeof(nbvh_hit_record), cudaMemcpyDeviceToHost, 0));
    CU_TRY(ctx, cudaStreamSynchronize(0));
    return NBVH_OK;
}

// Frames in flight: the same work as nbvh_render_host, split into submit and wait so that the device-to-host copy of frame i
// overlaps the kernels of frame i + 1 (two frame buffers, a render stream and a copy stream, events in between).  ldr_host
// must be pinned memory for the copy to be asynchronous.  slot in {0, 1}; a slot must be waited for before it is reused.
int nbvh_render_host_submit(nbvh_ctx *ctx, uint32_t accumulated_spp, uint32_t sample_offset, const nbvh_camera_params *camera_host,
                            uint8_t *ldr_host, uint32_t slot)
{
    if (!ctx || !ldr_host || slot >= (uint32_t)kPipeSlots)
        return NBVH_ERR_INVALID_ARG;
    const size_t n = (size_t)ctx->fb_w * ctx->fb_h;
    if (n == 0)
        return fail(ctx, NBVH_ERR_STATE, "framebuffer not set");
    if (!ctx->pipe_render) {
        CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->pipe_render, cudaStreamNonBlocking));
        for (int k = 0; k < kPipeSlots - 1; k++)
            CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->pipe_render_alt[k], cudaStreamNonBlocking));
        CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->pipe_copy, cudaStreamNonBlocking));
        for (int k = 0; k < kPipeSlots; k++) {
            CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->pipe_rendered[k], cudaEventDisableTiming));
            CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->pipe_copied[k], cudaEventDisableTiming));
        }
    }
    if (ctx->ldr_slot_pixels != n) {
        CU_TRY(ctx, cudaDeviceSynchronize());
        for (int k = 0; k < kPipeSlots; k++) {
            cudaFree(ctx->d_ldr_slot[k]);
            ctx->d_ldr_slot[k] = nullptr;
            CU_TRY(ctx, cudaMalloc(&ctx->d_ldr_slot[k], n * sizeof(uchar4)));
        }
        ctx->ldr_slot_pixels = n;
    }
    if (camera_host)
        ctx->camera = *camera_host;
    // Independent frames (accumulated_spp == 0: the accumulation is overwritten, never read) at max_depth 1 render on one stream
    // per slot with their own queue and counters, so the kernels of consecutive frames overlap; accumulating samples and the
    // multi-bounce pipeline (one set of path lists) stay on a single stream.
    const bool concurrent  = accumulated_spp == 0 && ctx->max_depth == 1 && !ctx->mega_kernel && !ctx->scene_const.ema_accumulate && !ctx->hits;
    cudaStream_t rstream   = (concurrent && slot >= 1u) ? ctx->pipe_render_alt[slot - 1u] : ctx->pipe_render;
    if (!concurrent) { // order after whatever the other slots' streams still run
        for (uint32_t k = 0; k < (uint32_t)kPipeSlots; k++)
            if (k != slot)
                CU_TRY(ctx, cudaStreamWaitEvent(rstream, ctx->pipe_rendered[k], 0));
    }
    // the slot's previous copy must have drained before the kernels overwrite its frame buffer
    CU_TRY(ctx, cudaStreamWaitEvent(rstream, ctx->pipe_copied[slot], 0));
    uchar4 *saved     = ctx->ldr;
    ctx->ldr          = ctx->d_ldr_slot[slot];
    ctx->render_slot  = concurrent ? (int)slot : 0;
    const int rc      = nbvh_render(ctx, accumulated_spp, sample_offset, 0, 0, rstream);
    ctx->render_slot  = 0;
    ctx->ldr          = saved;
    if (rc != NBVH_OK)
        return rc;
    CU_TRY(ctx, cudaEventRecord(ctx->pipe_rendered[slot], rstream));
    CU_TRY(ctx, cudaStreamWaitEvent(ctx->pipe_copy, ctx->pipe_rendered[slot], 0));
    CU_TRY(ctx, cudaMemcpyAsync(ldr_host, ctx->d_ldr_slot[slot], n * sizeof(uchar4), cudaMemcpyDeviceToHost, ctx->pipe_copy));
    CU_TRY(ctx, cudaEventRecord(ctx->pipe_copied[slot], ctx->pipe_copy));
    return NBVH_OK;
}

int nbvh_render_host_wait(nbvh_ctx *ctx, uint32_t slot)
{
    if (!ctx || slot >= (uint32_t)kPipeSlots || !ctx->pipe_copied[slot])
        return NBVH_ERR_INVALID_ARG;
    CU_TRY(ctx, cudaEventSynchronize(ctx->pipe_copied[slot]));
    return NBVH_OK;
}

int nbvh_selftest_division(nbvh_ctx *ctx, uint64_t seed, uint64_t n_trials, uint64_t *mismatches_host)
{
    if (!ctx || !mismatches_host)
        return NBVH_ERR_INVALID_ARG;
    unsigned long long *d = ctx->d_counters + CNT_COUNT; // spare slot
    CU_TRY(ctx, cudaMemset(d, 0, sizeof(unsigned long long)));
    const uint32_t iters  = 1024u;
    const uint32_t blocks = (uint32_t)((n_trials + 256ull * iters - 1) / (256ull * iters));
    CU_TRY(ctx, launch_selftest_division(seed, blocks ? blocks : 1u, iters, d, 0));
    CU_TRY(ctx, cudaMemcpy(mismatches_host, d, sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return NBVH_OK;
}

int nbvh_get_counters(nbvh_ctx *ctx, uint64_t *out_host, uint32_t n)
{
    if (!ctx || !out_host || n > CNT_COUNT)
        return NBVH_ERR_INVALID_ARG;
    CU_TRY(ctx, cudaMemcpy(out_host, ctx->d_counters, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost));
    return NBVH_OK;
}

int nbvh_reset_counters(nbvh_ctx *ctx)
{
    if (!ctx)
        return NBVH_ERR_INVALID_ARG;
    CU_TRY(ctx, cudaMemset(ctx->d_counters, 0, sizeof(unsigned long long) * CNT_COUNT));
    ctx->total_ms = 0.f, ctx->launches = 0;
    return NBVH_OK;
}

int nbvh_enable_kernel_timing(nbvh_ctx *ctx, int enable)
{
    if (!ctx)
        return NBVH_ERR_INVALID_ARG;
    ctx->timing = enable != 0;
    return NBVH_OK;
}

int nbvh_get_kernel_time(nbvh_ctx *ctx, float *total_ms, uint32_t *launches)
{
    if (!ctx)
        return NBVH_ERR_INVALID_ARG;
    if (total_ms)
        *total_ms = ctx->total_ms;
    if (launches)
        *launches = ctx->launches;
    return NBVH_OK;
}

int nbvh_enable_stage_timing(nbvh_ctx *ctx, uint32_t max_frames)
{
    if (!ctx)
        return NBVH_ERR_INVALID_ARG;
    for (cudaEvent_t e : ctx->stage_ev)
        cudaEventDestroy(e);
    ctx->stage_ev.clear();
    ctx->stage_used = 0;
    for (uint32_t i = 0; i < 3u * max_frames; i++) {
        cudaEvent_t e;
        CU_TRY(ctx, cudaEventCreate(&e));
        ctx->stage_ev.push_back(e);
    }
    return NBVH_OK;
}

int nbvh_get_stage_times(nbvh_ctx *ctx, float *primary_ms_total, float *neural_ms_total, uint32_t *frames)
{
    if (!ctx)
        return NBVH_ERR_INVALID_ARG;
    float a = 0.f, b = 0.f;
    for (uint32_t i = 0; i < ctx->stage_used; i++) {
        cudaEvent_t *e = ctx->stage_ev.data() + 3u * i;
        CU_TRY(ctx, cudaEventSynchronize(e[2]));
        float m0 = 0.f, m1 = 0.f;
        CU_TRY(ctx, cudaEventElapsedTime(&m0, e[0], e[1]));
        CU_TRY(ctx, cudaEventElapsedTime(&m1, e[1], e[2]));
        a += m0, b += m1;
    }
    if (primary_ms_total)
        *primary_ms_total = a;
    if (neural_ms_total)
        *neural_ms_total = b;
    if (frames)
        *frames = ctx->stage_used;
    ctx->stage_used = 0;
    return NBVH_OK;
}

// ---- online optimisation -------------------------------------------------------------------------------------
static int ensure_training_state(nbvh_ctx *ctx, uint32_t batch, cudaStream_t st)
{
    if (!ctx->params_loaded)
        return fail(ctx, NBVH_ERR_STATE, "network parameters not loaded (nbvh_init_params / nbvh_load_network_file)");
    // TMEM budget of k_train_mlp: working accumulators + one fp32 accumulator per weight matrix
    if (192u + ctx->in_k + (ctx->n_layers - 2u) * kNeurons + kOutPadded > 512u)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "network too deep/wide for the TMEM-resident weight-gradient accumulators");
    const size_t n = ctx->n_params;
    if (!ctx->d_grads) {
        CU_TRY(ctx, cudaMalloc(&ctx->own_grads, n * sizeof(float)));
        ctx->d_grads = ctx->own_grads;
    }
    if (!ctx->d_master) {
        CU_TRY(ctx, cudaMalloc(&ctx->d_master, n * sizeof(float)));
        CU_TRY(ctx, cudaMalloc(&ctx->d_m, n * sizeof(float)));
        CU_TRY(ctx, cudaMalloc(&ctx->d_v, n * sizeof(float)));
        CU_TRY(ctx, cudaMalloc(&ctx->d_steps, n * sizeof(uint32_t)));
        ctx->train_state_valid = false;
    }
    if (!ctx->d_loss)
        CU_TRY(ctx, cudaMalloc(&ctx->d_loss, sizeof(float)));
    if (!ctx->train_state_valid) {
        CU_TRY(ctx, launch_half_to_float(ctx->d_params, ctx->d_master, n, st));
        CU_TRY(ctx, cudaMemsetAsync(ctx->d_m, 0, n * sizeof(float), st));
        CU_TRY(ctx, cudaMemsetAsync(ctx->d_v, 0, n * sizeof(float), st));
        CU_TRY(ctx, cudaMemsetAsync(ctx->d_steps, 0, n * sizeof(uint32_t), st));
        CU_TRY(ctx, cudaMemsetAsync(ctx->d_grads, 0, n * sizeof(float), st));
        ctx->opt_step = 0, ctx->lr_factor = 1.f;
        ctx->train_state_valid = true;
    }
    const uint32_t padded = (batch + kTileRows - 1u) / kTileRows * kTileRows;
    if (padded > ctx->feat_capacity) {
        cudaFree(ctx->d_feat), cudaFree(ctx->d_dfeat);
        ctx->d_feat = ctx->d_dfeat = nullptr, ctx->feat_capacity = 0;
        CU_TRY(ctx, cudaMalloc(&ctx->d_feat, (size_t)padded * ctx->in_k * 2u));
        CU_TRY(ctx, cudaMalloc(&ctx->d_dfeat, (size_t)padded * ctx->in_k * 2u));
        ctx->feat_capacity = padded;
    }
    return NBVH_OK;
}

int nbvh_training_forward_backward(nbvh_ctx *ctx, const float *in, const float *target, uint32_t batch, uint32_t global_batch,
                                   float *per_sample_loss, float *loss_sum_dev, void *stream)
{
    if (!ctx || !in || !target || batch == 0)
        return ctx ? fail(ctx, NBVH_ERR_INVALID_ARG, "nbvh_training_forward_backward: null input or empty batch") : NBVH_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_ready(ctx, false, false, true);
    if (rc == NBVH_OK)
        rc = ensure_training_state(ctx, batch, st);
    if (rc != NBVH_OK)
        return rc;
    if (global_batch < batch)
        global_batch = batch;
    const NetConst nc = make_net_const(ctx);
    float *loss_dev   = loss_sum_dev ? loss_sum_dev : ctx->d_loss;
    // GradientMode::Overwrite (trainer.h:171): every call starts from zero gradients
    CU_TRY(ctx, cudaMemsetAsync(ctx->d_grads, 0, ctx->n_params * sizeof(float), st));
    CU_TRY(ctx, cudaMemsetAsync(loss_dev, 0, sizeof(float), st));
    KernelTimer kt(ctx, st);
    CU_TRY(ctx, launch_encode_features(nc, in, batch, ctx->module_const.input_network_dims, ctx->d_feat, ctx->num_sms, st));
    TrainParams tp;
    memset(&tp, 0, sizeof(tp));
    tp.feat = ctx->d_feat, tp.dfeat = ctx->d_dfeat, tp.target = target, tp.per_sample_loss = per_sample_loss;
    tp.wgrad = ctx->d_grads, tp.loss_sum = loss_dev, tp.batch = batch;
    tp.n_total    = global_batch * ctx->pt_const.output_network_dims;
    tp.loss_scale = 128.f; // default_loss_scale<__half>() (ext/tiny-cuda-nn common.h:232)
    const nbvh_pt_output_constants &pc = ctx->pt_const;
    tp.loss.vis_dim = pc.output_visibility_dim, tp.loss.depth_dim = pc.output_depth_dim;
    tp.loss.normal_dim = pc.output_normal_dim, tp.loss.albedo_dim = pc.output_albedo_dim;
    tp.loss.vis_loss = pc.visibility_loss, tp.loss.depth_loss = pc.depth_loss;
    tp.loss.normal_loss = pc.normal_loss, tp.loss.albedo_loss = pc.albedo_loss;
    tp.loss.vis_scale = pc.visibility_loss_scale, tp.loss.depth_scale = pc.depth_loss_scale;
    tp.loss.normal_scale = pc.normal_loss_scale, tp.loss.albedo_scale = pc.albedo_loss_scale;
    tp.loss.normal_linear = pc.output_normal_linear_activation, tp.loss.depth_linear = pc.output_depth_linear_activation;
    tp.loss.albedo_linear = pc.output_albedo_linear_activation, tp.loss.vis_masking = pc.visibility_masking_loss;
    CU_TRY(ctx, launch_train_mlp(nc, tp, ctx->num_sms, st));
    CU_TRY(ctx, launch_grid_backward(nc, in, ctx->d_dfeat, batch, ctx->module_const.input_network_dims, ctx->d_grads + ctx->n_mlp_params,
                                     ctx->num_sms, st));
    kt.stop();
    return NBVH_OK;
}

float *nbvh_gradients(nbvh_ctx *ctx) { return ctx ? ctx->d_grads : nullptr; }

int nbvh_set_gradient_buffer(nbvh_ctx *ctx, float *grads)
{
    if (!ctx)
        return NBVH_ERR_INVALID_ARG;
    if (grads) {
        ctx->d_grads = grads;
    } else {
        if (!ctx->own_grads)
            CU_TRY(ctx, cudaMalloc(&ctx->own_grads, ctx->n_params * sizeof(float)));
        ctx->d_grads = ctx->own_grads;
    }
    return NBVH_OK;
}

int nbvh_optimizer_step(nbvh_ctx *ctx, void *stream)
{
    if (!ctx)
        return NBVH_ERR_INVALID_ARG;
    if (!ctx->train_state_valid || !ctx->d_grads)
        return fail(ctx, NBVH_ERR_STATE, "nbvh_optimizer_step before nbvh_training_forward_backward");
    cudaStream_t st = (cudaStream_t)stream;
    // ExponentialDecayOptimizer::step (exponential_decay.h:60-71)
    const nbvh_optimizer_desc &o = ctx->opt_desc;
    if (ctx->opt_step == 0)
        ctx->lr_factor = 1.f;
    if ((int64_t)ctx->opt_step >= o.decay_start && o.decay_interval > 0 &&
        ((int64_t)ctx->opt_step - o.decay_start) % o.decay_interval == 0)
        ctx->lr_factor *= o.decay_base;
    AdamParams ap;
    memset(&ap, 0, sizeof(ap));
    ap.n_params = ctx->n_params, ap.n_matrix = ctx->n_mlp_params;
    ap.loss_scale = 128.f, ap.lr = o.learning_rate * ctx->lr_factor;
    ap.beta1 = o.beta1, ap.beta2 = o.beta2, ap.epsilon = o.epsilon, ap.l2_reg = o.l2_reg;
    ap.grads = ctx->d_grads, ap.master = ctx->d_master, ap.m = ctx->d_m, ap.v = ctx->d_v, ap.steps = ctx->d_steps;
    ap.params = reinterpret_cast<__half *>(ctx->d_params), ap.wimg = ctx->d_wimg;
    ap.in_k = ctx->in_k, ap.n_layers = ctx->n_layers;
    for (uint32_t l = 0; l < ctx->n_layers; l++)
        ap.layer_off[l] = ctx->layer_off[l];
    CU_TRY(ctx, launch_adam(ap, st));
    ctx->opt_step++;
    ctx->training_step++;
    return NBVH_OK;
}

int nbvh_training_step(nbvh_ctx *ctx, const float *in, const float *target, uint32_t batch, float *per_sample_loss, float *loss_host,
                       void *stream)
{
    int rc = nbvh_training_forward_backward(ctx, in, target, batch, batch, per_sample_loss, nullptr, stream);
    if (rc != NBVH_OK)
        return rc;
    rc = nbvh_optimizer_step(ctx, stream);
    if (rc != NBVH_OK)
        return rc;
    if (loss_host) { // == trainer->loss(stream, *ctx): one scalar back to the host
        CU_TRY(ctx, cudaMemcpyAsync(loss_host, ctx->d_loss, sizeof(float), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
        CU_TRY(ctx, cudaStreamSynchronize((cudaStream_t)stream));
    }
    return NBVH_OK;
}

uint32_t nbvh_training_step_count(const nbvh_ctx *ctx) { return ctx ? ctx->training_step : 0; }

} // extern "C"

// Device-side TLAS build (f3): == CudaBackend::build_tlas (cuda_backend.cpp:312-383).  mesh_bvhs_device: the DEVICE array of descriptors
// that nbvh_scene_data.mesh_bvhs will point to (each BLAS already built: its root box is read from nodes[0]); neural_meshes_host lists
// the meshes represented by the N-BVH (the reference's m_neural_bvh_indices).  Outputs (device, caller-owned): 2 * num_meshes - 1
// TLASNode records and num_meshes blas_indices.
int nbvh_build_tlas(nbvh_ctx *ctx, const nbvh_cuda_blas *mesh_bvhs_device, uint32_t num_meshes, const uint32_t *neural_meshes_host,
                    uint32_t n_neural, nbvh_node32 *tlas_nodes_out, uint32_t *blas_indices_out, void *stream)
{
    if (!ctx || !mesh_bvhs_device || !tlas_nodes_out || !blas_indices_out || num_meshes == 0 || (n_neural && !neural_meshes_host))
        return NBVH_ERR_INVALID_ARG;
    if (num_meshes > (uint32_t)kMaxBlas)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "more than 16 BLAS");
    uint32_t mask = 0;
    for (uint32_t i = 0; i < n_neural; i++) {
        if (neural_meshes_host[i] >= num_meshes)
            return fail(ctx, NBVH_ERR_INVALID_ARG, "neural mesh index out of range");
        mask |= 1u << neural_meshes_host[i];
    }
    CU_TRY(ctx, launch_build_tlas(mesh_bvhs_device, num_meshes, mask, reinterpret_cast<float4 *>(tlas_nodes_out), blas_indices_out,
                                  (cudaStream_t)stream));
    return NBVH_OK;
}

// == NeuralBVHRenderer::compute_and_display_flip_error (neural_bvh_renderer.cu:34-96) without the colour-mapped display: mean LDR-FLIP of
// the accumulated image (accum / accumulated_spp) against a device RGBA32F reference image of the framebuffer's size; ppd <= 0 selects the
// reference's viewing conditions (0.7 m from a 0.7 m wide 3840-pixel monitor: 67.02 pixels per degree).  error_map_out: optional device
// array of W x H floats (the per-pixel FLIP error the reference colour-maps with magma).
int nbvh_flip_error(nbvh_ctx *ctx, const float *reference_rgba32f, uint32_t accumulated_spp, float ppd, float *error_map_out,
                    float *mean_flip_host, void *stream)
{
    if (!ctx || !reference_rgba32f || !mean_flip_host || accumulated_spp == 0)
        return NBVH_ERR_INVALID_ARG;
    if (!ctx->accum)
        return fail(ctx, NBVH_ERR_STATE, "framebuffer not set");
    return nbvh_flip_images(ctx, reference_rgba32f, reinterpret_cast<const float *>(ctx->accum), (float)accumulated_spp, ctx->fb_w, ctx->fb_h,
                            ppd, error_map_out, mean_flip_host, stream);
}

// The same for two arbitrary device RGBA32F images (linear RGB; alpha ignored); the test image is divided by test_divisor first.
int nbvh_flip_images(nbvh_ctx *ctx, const float *reference_rgba32f, const float *test_rgba32f, float test_divisor, uint32_t width,
                     uint32_t height, float ppd, float *error_map_out, float *mean_flip_host, void *stream)
{
    if (!ctx || !reference_rgba32f || !test_rgba32f || !mean_flip_host || width == 0 || height == 0 || !(test_divisor > 0.f))
        return NBVH_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (!(ppd > 0.f))
        ppd = 0.7f * (3840.0f / 0.7f) * (3.14159265358979f / 180.0f);
    if (ctx->flip_ppd != ppd) {
        if (flip_set_ppd(ppd, st) != cudaSuccess)
            return fail(ctx, NBVH_ERR_INVALID_ARG, "FLIP: pixels-per-degree value out of range (filter radius 1..64)");
        ctx->flip_ppd = ppd;
    }
    const size_t bytes = flip_workspace_bytes(width, height);
    if (ctx->build_ws_bytes < bytes) { // shares the grow-only scratch of nbvh_build_blas
        CU_TRY(ctx, cudaStreamSynchronize(st));
        cudaFree(ctx->d_build_ws);
        ctx->d_build_ws = nullptr, ctx->build_ws_bytes = 0;
        CU_TRY(ctx, cudaMalloc(&ctx->d_build_ws, bytes));
        ctx->build_ws_bytes = bytes;
    }
    double *sum_dev = reinterpret_cast<double *>(ctx->d_counters + CNT_COUNT); // spare 8-byte slot
    CU_TRY(ctx, launch_flip(reinterpret_cast<const float4 *>(reference_rgba32f), reinterpret_cast<const float4 *>(test_rgba32f), test_divisor,
                            width, height, ctx->d_build_ws, error_map_out, sum_dev, st));
    double sum = 0.0;
    CU_TRY(ctx, cudaMemcpyAsync(&sum, sum_dev, sizeof(double), cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    *mean_flip_host = (float)(sum / ((double)width * height));
    return NBVH_OK;
}

// ---- f4: result writers -------------------------------------------------------------------------------------------
// == write_exr (include/base/bitmap.h:97-182) as the reference calls it from RenderApp::save_framebuffer (render_app.cpp:132-165):
// scanline OpenEXR 2.0, channels A, B, G, R as 32-bit float, no compression (one scanline per chunk), every value -- alpha too --
// divided by rgb_renormalisation (the accumulated sample count).  Byte-identical to the file the reference's tinyexr writes
// (tests/golden/ref_exr_7x5.exr, generated by oracle/ref_tools/exr_golden.cpp).
namespace {
void exr_attr(std::string &h, const char *name, const char *type, const void *data, int32_t size)
{
    h.append(name, strlen(name) + 1);
    h.append(type, strlen(type) + 1);
    h.append(reinterpret_cast<const char *>(&size), 4);
    h.append(reinterpret_cast<const char *>(data), (size_t)size);
}
} // namespace

int nbvh_write_exr_host(const char *path, uint32_t width, uint32_t height, const float *rgba_host, float rgb_renormalisation)
{
    if (!path || !rgba_host || width == 0 || height == 0)
        return NBVH_ERR_INVALID_ARG;
    std::string h;
    const unsigned char magic[8] = {0x76, 0x2f, 0x31, 0x01, 2, 0, 0, 0};
    h.append(reinterpret_cast<const char *>(magic), 8);
    std::string ch;
    for (const char name : {'A', 'B', 'G', 'R'}) {
        ch.push_back(name), ch.push_back('\0');
        const int32_t v[4] = {2 /* FLOAT */, 0 /* pLinear + 3 reserved bytes */, 1, 1};
        ch.append(reinterpret_cast<const char *>(v), 16);
    }
    ch.push_back('\0');
    exr_attr(h, "channels", "chlist", ch.data(), (int32_t)ch.size());
    const unsigned char zero = 0;
    exr_attr(h, "compression", "compression", &zero, 1);
    const int32_t win[4] = {0, 0, (int32_t)width - 1, (int32_t)height - 1};
    exr_attr(h, "dataWindow", "box2i", win, 16);
    exr_attr(h, "displayWindow", "box2i", win, 16);
    exr_attr(h, "lineOrder", "lineOrder", &zero, 1);
    const float one = 1.f, centre[2] = {0.f, 0.f};
    exr_attr(h, "pixelAspectRatio", "float", &one, 4);
    exr_attr(h, "screenWindowCenter", "v2f", centre, 8);
    exr_attr(h, "screenWindowWidth", "float", &one, 4);
    h.push_back('\0');
    FILE *f = fopen(path, "wb");
    if (!f)
        return NBVH_ERR_IO;
    const size_t row_bytes = (size_t)width * 16, chunk = 8 + row_bytes;
    std::vector<uint64_t> offsets(height);
    for (uint32_t y = 0; y < height; y++)
        offsets[y] = h.size() + (size_t)height * 8 + (size_t)y * chunk;
    bool ok = fwrite(h.data(), 1, h.size(), f) == h.size() && fwrite(offsets.data(), 8, height, f) == height;
    std::vector<float> row((size_t)width * 4);
    for (uint32_t y = 0; y < height && ok; y++) {
        const float *src = rgba_host + (size_t)y * width * 4;
        for (int c = 0; c < 4; c++) // file channel c = A, B, G, R = source component 3 - c
            for (uint32_t x = 0; x < width; x++)
                row[(size_t)c * width + x] = src[4 * x + (3 - c)] / rgb_renormalisation;
        const int32_t hdr[2] = {(int32_t)y, (int32_t)row_bytes};
        ok = fwrite(hdr, 4, 2, f) == 2 && fwrite(row.data(), 4, row.size(), f) == row.size();
    }
    ok = (fclose(f) == 0) && ok;
    return ok ? NBVH_OK : NBVH_ERR_IO;
}

// == NeuralBVHRenderer::save_image (neural_bvh_renderer_init.cu:361-383): reads the accumulation buffer back and writes it divided by
// accumulated_spp; refuses while ema_accumulate is on, like the reference ("would just save incorrect images").
int nbvh_save_image_exr(nbvh_ctx *ctx, const char *path, uint32_t accumulated_spp, void *stream)
{
    if (!ctx || !path || accumulated_spp == 0)
        return NBVH_ERR_INVALID_ARG;
    if (!ctx->accum)
        return fail(ctx, NBVH_ERR_STATE, "framebuffer not set");
    if (ctx->scene_const.ema_accumulate)
        return fail(ctx, NBVH_ERR_STATE, "saving images while ema_accumulate is on is not supported");
    std::vector<float> host((size_t)ctx->fb_w * ctx->fb_h * 4);
    cudaStream_t st = (cudaStream_t)stream;
    CU_TRY(ctx, cudaMemcpyAsync(host.data(), ctx->accum, host.size() * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    const int rc = nbvh_write_exr_host(path, ctx->fb_w, ctx->fb_h, host.data(), (float)accumulated_spp);
    return rc == NBVH_OK ? rc : fail(ctx, rc, std::string("cannot write ") + path);
}

// == FrameTimer::write_to_file (include/utils/frame_timer.h:33-58): "frame index,frame time (ms)[, loss]" and one line per frame with
// the default ostream formatting (6 significant digits).
int nbvh_write_frame_times_csv_host(const char *path, const double *frame_ms, const float *losses_or_null, uint32_t n)
{
    if (!path || (!frame_ms && n))
        return NBVH_ERR_INVALID_ARG;
    FILE *f = fopen(path, "w");
    if (!f)
        return NBVH_ERR_IO;
    fputs(losses_or_null ? "frame index,frame time (ms), loss\n" : "frame index,frame time (ms)\n", f);
    for (uint32_t i = 0; i < n; i++) {
        if (losses_or_null)
            fprintf(f, "%u,%g,%g\n", i, frame_ms[i], (double)losses_or_null[i]);
        else
            fprintf(f, "%u,%g\n", i, frame_ms[i]);
    }
    return fclose(f) == 0 ? NBVH_OK : NBVH_ERR_IO;
}

// Result file names of the reference's validation mode (neural_bvh_renderer_init.cu:328-372):
// kind 0: <dir>/results/<base>_<variant>_spp_<n>.exr, 1: ..._learning_spp_<n>.csv, 2: ..._inference_spp_<n>.csv
int nbvh_result_path(const char *scene_dir, const char *base_config_name, const char *variant_config_name, int kind, uint32_t accumulated_spp,
                     char *out, size_t out_capacity)
{
    if (!scene_dir || !base_config_name || !variant_config_name || !out || kind < 0 || kind > 2)
        return NBVH_ERR_INVALID_ARG;
    const char *tag[] = {"", "_learning", "_inference"};
    const std::string p = std::string(scene_dir) + "/results/" + base_config_name + "_" + variant_config_name + tag[kind] + "_spp_" +
                          std::to_string(accumulated_spp) + (kind == 0 ? ".exr" : ".csv");
    if (p.size() + 1 > out_capacity)
        return NBVH_ERR_INVALID_ARG;
    memcpy(out, p.c_str(), p.size() + 1);
    return NBVH_OK;
}

// == the module-owned part of NeuralBVHRenderer::save_config (neural_bvh_renderer_init.cu:406-509): cuda_scene {fb_size, camera,
// scene_constants} and neural_bvh_renderer {neural_module {max_depth, pt_output_constants, path_module_constants, bvh_module_constants,
// network, optimizer, segment_encoding, direction_encoding}, training_step, bvh_constants} with the reference's key names, floats
// printed with 9 significant digits so that nbvh_load_config_json(save(x)) restores every constant block bit for bit.  The renderer
// adds the keys it owns (batch_size, tree_cut_learning, bvh_split_scheduler, ...).  Writes at most `capacity` bytes (NUL-terminated)
// and the needed size (including the NUL) into *needed.
int nbvh_save_config_json(const nbvh_ctx *ctx, char *out, size_t capacity, size_t *needed)
{
    if (!ctx || (!out && capacity))
        return NBVH_ERR_INVALID_ARG;
    std::string s;
    char buf[64];
    bool first = true;
    auto key = [&](const char *k) {
        if (!first)
            s += ", ";
        first = false;
        s += "\"", s += k, s += "\": ";
    };
    auto open  = [&](const char *k) { key(k), s += "{", first = true; };
    auto close = [&]() { s += "}", first = false; };
    auto F = [&](const char *k, float v) { key(k), snprintf(buf, sizeof(buf), "%.9g", (double)v), s += buf; };
    auto I = [&](const char *k, long long v) { key(k), s += std::to_string(v); };
    auto B = [&](const char *k, bool v) { key(k), s += v ? "true" : "false"; };
    auto S = [&](const char *k, const char *v) { key(k), s += "\"", s += v, s += "\""; };
    s += "{", first = true;
    open("cuda_scene");
    key("fb_size"), s += "[" + std::to_string(ctx->fb_w) + ", " + std::to_string(ctx->fb_h) + "]";
    if (ctx->cam_set) {
        open("camera");
        S("type", "perspective");
        F("yfov", ctx->cam_yfov);
        key("matrix"), s += "[";
        for (int i = 0; i < 16; i++)
            snprintf(buf, sizeof(buf), i ? ", %.9g" : "%.9g", (double)ctx->cam_matrix[i]), s += buf;
        s += "]";
        close();
    }
    {
        const nbvh_scene_constants &c = ctx->scene_const;
        open("scene_constants");
        B("force_constant_material", c.force_constant_material), F("constant_roughness", c.constant_roughness);
        F("constant_specular", c.constant_specular), F("constant_metallic", c.constant_metallic);
        F("envmap_rotation", c.envmap_rotation), F("envmap_strength", c.envmap_strength);
        B("ema_accumulate", c.ema_accumulate), F("ema_weight", c.ema_weight), B("envmap_eval", c.envmap_eval);
        B("emission_eval", c.emission_eval), B("skylike_background", c.skylike_background);
        close();
    }
    close();
    open("neural_bvh_renderer");
    S("bvh_module_type", "Path");
    open("neural_module");
    I("max_depth", ctx->max_depth);
    {
        const nbvh_pt_output_constants &c = ctx->pt_const;
        open("pt_output_constants");
        I("output_visibility_dim", c.output_visibility_dim), I("output_depth_dim", c.output_depth_dim);
        I("output_normal_dim", c.output_normal_dim), I("output_albedo_dim", c.output_albedo_dim);
        I("output_network_dims", c.output_network_dims);
        S("visibility_loss", kLossTypes[c.visibility_loss]), S("depth_loss", kLossTypes[c.depth_loss]);
        S("normal_loss", kLossTypes[c.normal_loss]), S("albedo_loss", kLossTypes[c.albedo_loss]);
        F("visibility_loss_scale", c.visibility_loss_scale), F("depth_loss_scale", c.depth_loss_scale);
        F("normal_loss_scale", c.normal_loss_scale), F("albedo_loss_scale", c.albedo_loss_scale);
        S("depth_learning_mode", kDepthModes[c.depth_learning_mode]);
        B("output_normal_linear_activation", c.output_normal_linear_activation);
        B("output_depth_linear_activation", c.output_depth_linear_activation);
        B("output_albedo_linear_activation", c.output_albedo_linear_activation);
        B("visibility_masking_loss", c.visibility_masking_loss);
        F("primary_visibility_threshold", c.primary_visibility_threshold);
        F("secondary_visibility_threshold", c.secondary_visibility_threshold);
        F("depth_epsilon", c.depth_epsilon), F("fixed_depth_epsilon", c.fixed_depth_epsilon);
        close();
    }
    {
        const nbvh_path_constants &c = ctx->path_const;
        open("path_module_constants");
        S("output_mode", kOutputModes[c.output_mode]), F("normal_epsilon_offset", c.normal_epsilon_offset);
        B("apply_russian_roulette", c.apply_russian_roulette), I("rr_start_depth", c.rr_start_depth);
        F("neural_roughness", c.neural_roughness), F("neural_specular", c.neural_specular), F("neural_metallic", c.neural_metallic);
        close();
    }
    {
        const nbvh_module_constants &c = ctx->module_const;
        open("bvh_module_constants");
        I("input_segment_point_dim", c.input_segment_point_dim), I("input_segment_points", c.input_segment_points);
        I("input_segment_points_dim", c.input_segment_points_dim), I("input_segment_direction_dim", c.input_segment_direction_dim);
        I("input_network_dims", c.input_network_dims), S("segment_distribution", kSegmentDistributions[c.segment_distribution]);
        B("antialias_direction", c.antialias_direction), B("merge_operator", c.merge_operator);
        B("apply_encoding_trafo", c.apply_encoding_trafo);
        close();
    }
    {
        const nbvh_network_desc &d = ctx->net_desc;
        open("network");
        S("network", "FullyFusedMLP"), I("n_neurons", d.n_neurons), I("n_hidden_layers", d.n_hidden_layers);
        S("activation", "ReLU"), S("output_activation", "Sigmoid");
        close();
        const nbvh_optimizer_desc &o = ctx->opt_desc;
        open("optimizer");
        S("type", "Adam"), F("learning_rate", o.learning_rate), I("decay_start", o.decay_start), I("decay_interval", o.decay_interval);
        F("decay_base", o.decay_base), F("beta1", o.beta1), F("beta2", o.beta2), F("epsilon", o.epsilon), F("l2_reg", o.l2_reg);
        F("relative_decay", 0.f), F("absolute_decay", 0.f), B("adabound", false);
        close();
        open("segment_encoding");
        S("type", "HashGrid");
        open("hashgrid_options");
        S("otype", "HashGrid"), I("n_levels", d.n_levels), I("n_features_per_level", d.n_features_per_level);
        I("log2_hashmap_size", d.log2_hashmap_size), I("base_resolution", d.base_resolution), F("per_level_scale", d.per_level_scale);
        S("hash", "CoherentPrime"), S("interpolation", "Linear"), B("stochastic_interpolation", false);
        close();
        S("reduction", "Concatenation");
        close();
        open("direction_encoding");
        S("type", ctx->dir_enc_type.c_str());
        close();
    }
    close(); // neural_module
    I("training_step", ctx->training_step);
    {
        const nbvh_renderer_constants &c = ctx->bvh_const;
        open("bvh_constants");
        I("num_lods_learned", c.num_lods_learned), I("current_learned_lod", c.current_learned_lod), I("rendered_lod", c.rendered_lod);
        B("ignore_neural_blas", c.ignore_neural_blas), F("leaf_boundary_sampling_ratio", c.leaf_boundary_sampling_ratio);
        F("inference_mint_epsilon", c.inference_mint_epsilon), F("inference_maxt_epsilon", c.inference_maxt_epsilon);
        F("learning_mint_epsilon", c.learning_mint_epsilon), F("learning_maxt_epsilon", c.learning_maxt_epsilon);
        F("sampling_bbox_scaling", c.sampling_bbox_scaling), F("fixed_node_dilation", c.fixed_node_dilation);
        F("node_extent_dilation", c.node_extent_dilation), F("loss_exposure", c.loss_exposure);
        B("show_loss_heuristic", c.show_loss_heuristic), B("ellipsoidal_bvh", c.ellipsoidal_bvh), B("rescale_ellipsoid", c.rescale_ellipsoid);
        I("log2_bvh_max_traversals", c.log2_bvh_max_traversals), B("sample_cosine_weighted", c.sample_cosine_weighted);
        B("backface_bounce_training", c.backface_bounce_training), I("max_training_bounces", c.max_training_bounces);
        B("training_node_survival", c.training_node_survival), F("survival_probability_scale", c.survival_probability_scale);
        F("min_rr_survival_prob", c.min_rr_survival_prob);
        close();
    }
    close(); // neural_bvh_renderer
    s += "}";
    if (needed)
        *needed = s.size() + 1;
    if (capacity) {
        const size_t n = s.size() < capacity - 1 ? s.size() : capacity - 1;
        memcpy(out, s.data(), n);
        out[n] = '\0';
    }
    return s.size() + 1 <= capacity || !capacity ? NBVH_OK : NBVH_ERR_INVALID_ARG;
}

// host_context.hpp -- the context object of libnbvh_b200.so and the host-side helpers shared by its translation units:
//   context.cu     life cycle, constant blocks, scene / tree-cut / parameter upload, framebuffer, counters and timing
//   config.cu      the reference's config JSON (defaults, load, save)
//   io.cu          on-disk formats: .nbvh, network_models JSON, EXR, timing CSV, result paths
//   abi_render.cu  inference entry points (queries, parity hooks, frames, ray batches, textures, metrics, BLAS / TLAS build)
//   abi_train.cu   online optimisation (sample generation, tree-cut learning, training step, data-parallel gradient exchange)
#pragma once
#include "json_min.hpp"
#include "kernels.cuh"
#include "sampler.cuh"
#include "path.cuh"
#include "treecut.cuh"
#include "shading.cuh"

#include <algorithm>
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

namespace nbvh_host {

static const char *const kLossTypes[] = {"RelativeL2", "L2", "RelativeL1", "L1", "RelativeL2Luminance", "BinaryCrossEntropy",
                                  "LogL2", "LHalf"};
static const char *const kDepthModes[]           = {"LocalT", "WorldT", "Position", "LocalPosition"};
static const char *const kSegmentDistributions[] = {"EquidistantInterior", "EquidistantExterior", "Uniform"};
static const char *const kOutputModes[]          = {"NeuralPathTracing"};

template <size_t N>
inline int find_index(const std::string &s, const char *const (&table)[N])
{
    for (size_t i = 0; i < N; i++)
        if (s == table[i])
            return (int)i;
    return -1;
}

nbvh_renderer_constants default_renderer_constants();
nbvh_module_constants default_module_constants();
nbvh_pt_output_constants default_pt_constants();
nbvh_path_constants default_path_constants();
nbvh_scene_constants default_scene_constants();
nbvh_camera_params camera_from_matrix(const float *m, float yfov_rad, uint32_t w, uint32_t h);
bool read_file(const std::string &path, std::string &out);
uint16_t f32_to_f16_host(float f);

} // namespace nbvh_host
using namespace nbvh_host;

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
    // data-parallel optimisation (comm.cu): NCCL communicator (the caller's, or one the library created from a unique id),
    // communication stream and the events that order the gradient buckets against the compute stream
    void *nccl_comm = nullptr;
    bool nccl_owned = false;
    int dp_rank = 0, dp_world = 1;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_bucket[2] = {nullptr, nullptr}, ev_reduced = nullptr;
    uint32_t *d_sampler_counters = nullptr;  // cursor / flags of one walk launch (sampler.cu)
    // walk-ahead (nbvh_sampler_walk_ahead): one k_sample_walk launch walks the slots of `walk_ahead_steps` consecutive optimisation
    // steps (training_step, training_step + stride, ...); the items stay valid while the tree-cut and the sampler constants do
    uint32_t walk_ahead_steps = 1, walk_ahead_stride = 1;
    uint32_t walk_version = 0; // bumped by everything the walk depends on (node array, constants)
    bool ahead_valid = false;
    uint32_t ahead_first_step = 0, ahead_stride = 1, ahead_steps = 0, ahead_batch = 0, ahead_version = 0;
    cudaEvent_t ahead_done = nullptr; // the walk launch has finished (orders target passes issued on other streams)
    void *d_sampler_items = nullptr; // two-pass sample generator: 48-byte items handed from k_sample_walk to k_sample_targets (grow-only)
    uint64_t sampler_item_capacity = 0;
    float dilated_with  = -1.f;
    bool nbvh_safe = false;            // node arrays passed k_check_nodes (hoisted IEEE division precondition)
    float nbvh_poke = -1.f;            // dilated tree-cut: largest overhang of a descendant box over an ancestor's (< 0: walk-stack skipping off)
    std::vector<uint8_t> blas_safe;   // per mesh: its BLAS nodes passed k_check_nodes
    BlasConst *d_blas_ext = nullptr;  // device tables of the BLAS descriptors / shading attributes of scenes with more than kMaxBlas meshes
    BlasAttr *d_attr_ext = nullptr;
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
    // Drain order of the neural ray queue by expected query-chain length (kernels.cuh: RenderParams::class_lists).  d_chain_hist:
    // per-pixel chain length of the last frame (k_neural writes it, k_primary reads it); one header + index-list block per frame
    // slot.  Scheduling only: results never depend on it.  NBVH_CHAIN_CLASSES="t0,t1,t2,t3" sets the class limits ("0" disables
    // the ordering: arrival order), NBVH_CHAIN_PRED the predictor flags (bit 0: 3x3 maximum, bit 1: decaying maximum)
    uint8_t *d_chain_hist = nullptr;
    size_t chain_hist_pixels = 0;
    uint32_t *d_class_lists[kPipeSlots] = {};
    size_t class_lists_capacity[kPipeSlots] = {};
    uint32_t long_chain_thr = (4u << 24) | (3u << 16) | (2u << 8) | 2u, chain_pred = 1; // classes: >= 4, 3, 2, -, 1 (s45 sweep)
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
    // the same for the optimisation step: per recorded step 7 events (before encode | mlp | grid backward | gradient exchange |
    // end of forward_backward; before | after Adam), read back by nbvh_get_training_stage_times
    std::vector<cudaEvent_t> train_ev;
    uint32_t train_ev_steps = 0, train_ev_fb = 0, train_ev_opt = 0;

    // training state (allocated lazily): tcnn::Trainer's buffers (trainer.h:68-87) + Adam moments and step counters
    float *d_grads = nullptr, *d_master = nullptr, *d_m = nullptr, *d_v = nullptr;
    uint32_t *d_steps = nullptr;
    float *d_wgrad_partial = nullptr; // num_sms x n_mlp_params: per-CTA weight-gradient sums of k_train_mlp (training.cu)
    uint8_t *d_feat = nullptr, *d_dfeat = nullptr;
    uint32_t feat_capacity = 0; // samples
    float *d_loss = nullptr;
    float *own_grads = nullptr; // internal gradient buffer (d_grads may point at a host-owned one instead)
    bool train_state_valid = false;
    uint32_t opt_step = 0;    // optimiser steps since the state was (re)initialised (tcnn: optimizer->step())
    float lr_factor   = 1.f;  // ExponentialDecay's running factor
};

namespace nbvh_host {

// Every entry point runs on the context's device: several contexts (one per GPU) may live in one process, and the caller's
// current device is restored on return.
struct DeviceGuard {
    int prev      = -1;
    bool switched = false;
    explicit DeviceGuard(const nbvh_ctx *ctx)
    {
        if (ctx && cudaGetDevice(&prev) == cudaSuccess && prev != ctx->device)
            switched = cudaSetDevice(ctx->device) == cudaSuccess;
    }
    ~DeviceGuard()
    {
        if (switched)
            cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard &)            = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};

int fail(nbvh_ctx *ctx, int code, const std::string &msg);
std::string layout_error(const nbvh_ctx *ctx);
int rebuild_layout(nbvh_ctx *ctx, bool keep_parameters_if_unchanged = false);
int upload_weight_image(nbvh_ctx *ctx, const uint16_t *mlp_host);
NetConst make_net_const(const nbvh_ctx *ctx);
NbvhConst make_nbvh_const(const nbvh_ctx *ctx, bool inference);
int check_nodes(nbvh_ctx *ctx, const float4 *nodes, uint32_t n, bool &ok);
int refresh_dilated(nbvh_ctx *ctx);
EncodeConst make_encode_const(const nbvh_ctx *ctx);
DecodeConst make_decode_const(const nbvh_ctx *ctx);
int check_wavefront_modes(nbvh_ctx *ctx);
int check_ready(nbvh_ctx *ctx, bool need_scene, bool need_nbvh, bool need_params);
CameraConst make_camera_const(const nbvh_ctx *ctx);
int adopt_nbvh(nbvh_ctx *ctx, const void *nodes, uint32_t n, cudaMemcpyKind kind);
BlasConst host_blas(const nbvh_ctx *ctx, uint32_t i); // descriptor of mesh i (host copy: works for any mesh count)
// comm.cu
int dp_allreduce_bucket(nbvh_ctx *ctx, int bucket, float *ptr, size_t count, cudaStream_t st);
int dp_join(nbvh_ctx *ctx, cudaStream_t st);
void dp_release(nbvh_ctx *ctx);

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

} // namespace nbvh_host

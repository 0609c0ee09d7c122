// context.cu -- host side of libnbvh_b200.so: context object, constant blocks, parameter layout, scene / tree-cut / weight upload,
// framebuffer, counters and timing.  Mirrors what NeuralBVHRenderer / NeuralBVHModule do on the host around the hot path
// (/root/reference/src/cuda/neural/neural_bvh_renderer_init.cu:385-404,566-749,
//  /root/reference/src/cuda/neural/neural_bvh_renderer_accel.cpp:170-254,
//  /root/reference/src/cuda/neural/neural_bvh_module.cu:12-263).
#include "host_context.hpp"

namespace nbvh_host {

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
// What the kernels cannot run, as a message (empty = fine).  Pure check: callers validate BEFORE they commit a new description.
std::string layout_error(const nbvh_ctx *ctx)
{
    const nbvh_network_desc &d = ctx->net_desc;
    if (d.n_features_per_level != 4)
        return "only n_features_per_level == 4 is implemented";
    if (d.n_neurons != kNeurons)
        return "only n_neurons == 64 is implemented";
    if (d.n_levels < 1 || d.n_levels > kMaxLevels || d.n_hidden_layers < 1 || d.n_hidden_layers + 1 > kMaxLayers)
        return "n_levels / n_hidden_layers out of range";
    if (ctx->module_const.input_segment_direction_dim != 0)
        return "direction encoding input is not implemented";
    const uint32_t n_points = ctx->module_const.input_segment_points;
    if (n_points < 1 || n_points > 4)
        return "1..4 segment points supported";
    const uint32_t in_k = n_points * d.n_levels * d.n_features_per_level;
    if (in_k % 16u != 0u || in_k > kMaxInputK)
        return "encoded input width must be a multiple of 16 and <= 128";
    if (ctx->pt_const.output_network_dims > 8)
        return "output_network_dims > 8";
    return std::string();
}

int rebuild_layout(nbvh_ctx *ctx, bool keep_parameters_if_unchanged)
{
    const nbvh_network_desc &d = ctx->net_desc;
    const std::string err      = layout_error(ctx);
    if (!err.empty())
        return fail(ctx, NBVH_ERR_UNSUPPORTED, err);
    const uint32_t old_in_k = ctx->in_k, old_n_layers = ctx->n_layers, old_wimg_bytes = ctx->wimg_bytes;
    const uint64_t old_n_params = ctx->n_params;
    const std::vector<uint32_t> old_level_offsets = ctx->level_offsets, old_layer_off = ctx->layer_off;
    ctx->in_k = ctx->module_const.input_segment_points * d.n_levels * d.n_features_per_level;
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
    // A constant block that leaves the parameter layout as it is (segment distribution, epsilons, ...) keeps the loaded parameters
    // and the optimisation state: the reference's update of bvh_module_const does not touch the model either
    if (keep_parameters_if_unchanged && ctx->d_params && ctx->d_wimg && ctx->in_k == old_in_k && ctx->n_layers == old_n_layers && ctx->n_params == old_n_params &&
        ctx->wimg_bytes == old_wimg_bytes && ctx->level_offsets == old_level_offsets && ctx->layer_off == old_layer_off)
        return NBVH_OK;
    if (ctx->d_params)
        cudaFree(ctx->d_params);
    if (ctx->d_wimg)
        cudaFree(ctx->d_wimg);
    ctx->d_params = nullptr, ctx->d_wimg = nullptr;
    ctx->params_loaded = false;
    cudaFree(ctx->own_grads), cudaFree(ctx->d_master), cudaFree(ctx->d_m), cudaFree(ctx->d_v), cudaFree(ctx->d_steps), cudaFree(ctx->d_wgrad_partial);
    ctx->d_grads = ctx->own_grads = ctx->d_master = ctx->d_m = ctx->d_v = ctx->d_wgrad_partial = nullptr;
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
    // node index bits | truncated entry distance (device_core.cuh, nbvh_pop); at least 9 bits (sign, exponent) must remain for the distance
    bc.idx_mask = 0x7FFFFFFFu; // bit 31 of a stack entry is the children flag
    static const bool no_skip = [] { const char *e = getenv("NBVH_NO_STACK_SKIP"); return e && e[0] == '1'; }(); // A/B measurements
    bc.poke = 0.f;
    if (ctx->nbvh_poke >= 0.f && !no_skip) {
        bc.poke = ctx->nbvh_poke;
        uint32_t bits = 1u;
        while (bits < 32u && (1u << bits) < ctx->nbvh_nodes)
            ++bits;
        if (bits <= 20u)
            bc.idx_mask = (1u << bits) - 1u;
    }
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
    {
        // Input of the walk's distance-skipped stack entries (device_core.cuh, nbvh_pop): the largest distance by which a box REACHABLE
        // below a node sticks out of that node's box (0 for an exactly nested array).  Setup time only: one depth-first pass that
        // compares every node with all its ancestors.  Malformed arrays (child indices outside the array, cycles, NaN) disable skipping.
        std::vector<float> h((size_t)ctx->nbvh_nodes * 8u);
        CU_TRY(ctx, cudaMemcpy(h.data(), ctx->d_nbvh_dilated, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
        struct Item { uint32_t node, depth; };
        std::vector<Item> todo;
        std::vector<uint32_t> path;
        todo.push_back({0u, 0u});
        double poke    = 0.0;
        bool ok        = ctx->nbvh_nodes > 0;
        size_t visited = 0;
        while (ok && !todo.empty()) {
            const Item it = todo.back();
            todo.pop_back();
            if (++visited > (size_t)ctx->nbvh_nodes) { // a cycle: not a tree
                ok = false;
                break;
            }
            path.resize(it.depth);
            const float *q = h.data() + (size_t)it.node * 8u;
            for (uint32_t anc : path) {
                const float *p = h.data() + (size_t)anc * 8u;
                for (int k = 0; k < 3; k++) {
                    const double lo = (double)p[k] - (double)q[k], hi = (double)q[4 + k] - (double)p[4 + k];
                    if (!(lo == lo) || !(hi == hi))
                        ok = false;
                    poke = std::max(poke, std::max(lo, hi));
                }
            }
            uint32_t a, b;
            memcpy(&a, q + 3, 4);
            memcpy(&b, q + 7, 4);
            if (b == 0xFFFFFFFFu || a == 1u) // no neural child / LEAF_FINAL (neural_bvh2.cuh:16,29)
                continue;
            if (b >= ctx->nbvh_nodes || b + 1u >= ctx->nbvh_nodes) {
                ok = false;
                break;
            }
            path.push_back(it.node);
            todo.push_back({b, it.depth + 1u});
            todo.push_back({b + 1u, it.depth + 1u});
        }
        ctx->nbvh_poke = (ok && std::isfinite(poke)) ? (float)(poke * 1.0001) + 1e-30f : -1.f;
    }
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
    ec.uniform       = ctx->module_const.segment_distribution == 2 ? 1u : 0u;
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

DecodeConst make_decode_const(const nbvh_ctx *ctx)
{
    DecodeConst dc;
    memset(&dc, 0, sizeof(dc));
    dc.primary_thr         = ctx->pt_const.primary_visibility_threshold;
    dc.secondary_thr       = ctx->pt_const.secondary_visibility_threshold;
    dc.depth_epsilon       = ctx->pt_const.depth_epsilon;
    dc.fixed_depth_epsilon = ctx->pt_const.fixed_depth_epsilon;
    dc.depth_mode          = (uint32_t)ctx->pt_const.depth_learning_mode;
    for (int i = 0; i < 3; i++)
        dc.bbox_min[i] = ctx->scene_bbox.min[i], dc.bbox_max[i] = ctx->scene_bbox.max[i];
    return dc;
}

// the wavefront kernels (segment encoding + output decode) implement every segment distribution and the LocalT / WorldT depth modes
int check_wavefront_modes(nbvh_ctx *ctx)
{
    // SegmentDistribution::Uniform: k_neural<., true> (render.cu) draws the segment parameters from the ray's generator; the single
    // mega-kernel development mode has no such variant
    if (ctx->module_const.segment_distribution == 2 && ctx->mega_kernel)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "Uniform segment distribution is not available in NBVH_RENDER_MODE=mega");
    // Position / LocalPosition: the training sampler produces their targets (sampler.cu); the inference-side decode is not
    // implemented (the reference's LocalPosition branch prints every query, neural_bvh_module_common.cuh:131-134)
    if (ctx->pt_const.depth_learning_mode > 1)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "inference supports depth_learning_mode LocalT and WorldT only");

    return NBVH_OK;
}

int check_ready(nbvh_ctx *ctx, bool need_scene, bool need_nbvh, bool need_params)
{
    if (need_params && !ctx->params_loaded)
        return fail(ctx, NBVH_ERR_STATE, "network parameters not loaded");
    if (need_nbvh && !ctx->d_nbvh)
        return fail(ctx, NBVH_ERR_STATE, "neural BVH not set");
    if (need_scene && !ctx->scene_set)
        return fail(ctx, NBVH_ERR_STATE, "scene not set");
    return NBVH_OK;
}

} // namespace nbvh_host

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
    if (const char *m = getenv("NBVH_CHAIN_CLASSES")) { // A/B measurements: "t0,t1,t2,t3" (descending), "0" = arrival order
        unsigned t[4] = {0u, 0u, 0u, 0u};
        const int n = sscanf(m, "%u,%u,%u,%u", &t[0], &t[1], &t[2], &t[3]);
        ctx->long_chain_thr = n <= 0 ? 0u : ((t[0] & 255u) << 24) | ((t[1] & 255u) << 16) | ((t[2] & 255u) << 8) | (t[3] & 255u);
    }
    if (const char *m = getenv("NBVH_CHAIN_PRED"))
        ctx->chain_pred = (uint32_t)atoi(m) & 7u;
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
    dp_release(ctx);
    cudaFree(ctx->d_blas_ext), cudaFree(ctx->d_attr_ext);
    cudaFree(ctx->d_sampler_items), cudaFree(ctx->d_sampler_counters);
    if (ctx->ahead_done)
        cudaEventDestroy(ctx->ahead_done);
    cudaFree(ctx->d_node_keep);
    cudaFree(ctx->d_queue);
    cudaFree(ctx->d_chain_hist);
    for (int k = 0; k < kPipeSlots; k++)
        cudaFree(ctx->d_class_lists[k]);
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
    cudaFree(ctx->d_wgrad_partial);
    cudaFree(ctx->d_feat);
    cudaFree(ctx->d_dfeat);
    cudaFree(ctx->d_loss);
    if (ctx->ev0)
        cudaEventDestroy(ctx->ev0);
    if (ctx->ev1)
        cudaEventDestroy(ctx->ev1);
    for (cudaEvent_t e : ctx->stage_ev)
        cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->train_ev)
        cudaEventDestroy(e);
    delete ctx;
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
    DeviceGuard device_guard(ctx);
    if (!ctx || !pod_host)
        return NBVH_ERR_INVALID_ARG;
    size_t want = 0;
    void *dst   = block_ptr(ctx, which, want);
    if (!dst || bytes != want)
        return fail(ctx, NBVH_ERR_INVALID_ARG, "nbvh_set_constants: unknown block or size mismatch");
    if (which == NBVH_CONST_NETWORK || which == NBVH_CONST_MODULE || which == NBVH_CONST_PT_OUTPUT) {
        // these decide the parameter layout / the output width: compare first so per-frame re-uploads stay cheap, and validate the
        // new description before committing it -- a rejected block leaves the context (parameters included) untouched
        if (memcmp(dst, pod_host, bytes) == 0)
            return NBVH_OK;
        std::vector<uint8_t> old((const uint8_t *)dst, (const uint8_t *)dst + bytes);
        memcpy(dst, pod_host, bytes);
        const std::string err = layout_error(ctx);
        if (!err.empty()) {
            memcpy(dst, old.data(), bytes);
            return fail(ctx, NBVH_ERR_UNSUPPORTED, err);
        }
        ctx->walk_version++;
        return which == NBVH_CONST_PT_OUTPUT ? NBVH_OK : rebuild_layout(ctx, which == NBVH_CONST_MODULE);
    }
    memcpy(dst, pod_host, bytes);
    ctx->walk_version++; // sampler constants may have changed: walked-ahead items are stale
    if (which == NBVH_CONST_SCENE_BBOX)
        ctx->scene_bbox_set = true;
    return NBVH_OK;
}

int nbvh_get_constants(const nbvh_ctx *ctx, int which, void *pod_host, size_t bytes)
{
    DeviceGuard device_guard(ctx);
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
    DeviceGuard device_guard(ctx);
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
    DeviceGuard device_guard(ctx);
    if (!ctx || max_depth < 1)
        return NBVH_ERR_INVALID_ARG;
    ctx->max_depth = max_depth;
    return NBVH_OK;
}

// ---- scene, tree-cut, weights --------------------------------------------------------------------------
int nbvh_set_scene(nbvh_ctx *ctx, const nbvh_scene_data *s)
{
    DeviceGuard device_guard(ctx);
    if (!ctx || !s)
        return NBVH_ERR_INVALID_ARG;
    std::vector<nbvh_cuda_blas> blas(s->num_meshes);
    if (s->num_meshes)
        CU_TRY(ctx, cudaMemcpy(blas.data(), s->mesh_bvhs, sizeof(nbvh_cuda_blas) * s->num_meshes, cudaMemcpyDeviceToHost));
    memset(&ctx->scene, 0, sizeof(ctx->scene));
    ctx->scene.tlas.nodes        = reinterpret_cast<const float4 *>(s->tlas_nodes);
    ctx->scene.tlas.blas_indices = s->blas_indices;
    ctx->scene.tlas.num_meshes   = s->num_meshes;
    ctx->scene.n_blas            = s->num_meshes;
    std::vector<BlasConst> table(s->num_meshes);
    std::vector<BlasAttr> attrs(s->num_meshes);
    ctx->blas_safe.assign(s->num_meshes, 0);
    for (uint32_t i = 0; i < s->num_meshes; i++) {
        table[i].nodes        = reinterpret_cast<const float4 *>(blas[i].nodes);
        table[i].prim_indices = reinterpret_cast<const unsigned long long *>(blas[i].prim_indices);
        table[i].vertices     = blas[i].vertices;
        table[i].indices      = blas[i].indices;
        attrs[i].normals       = blas[i].normals;
        attrs[i].texcoords     = blas[i].texcoords;
        attrs[i].material_ids  = (blas[i].material_map && blas[i].num_materials) ? blas[i].material_ids : nullptr;
        attrs[i].material_map  = blas[i].material_map;
        attrs[i].num_materials = blas[i].material_map ? blas[i].num_materials : 0u;
        if (i < (uint32_t)kMaxBlas)
            ctx->scene.blas[i] = table[i];
        bool safe = false;
        int rc    = check_nodes(ctx, table[i].nodes, blas[i].num_nodes, safe);
        if (rc != NBVH_OK)
            return rc;
        ctx->blas_safe[i] = safe ? 1 : 0;
    }
    // more meshes than fit the kernel parameters: the descriptors and shading attributes move to device tables
    cudaFree(ctx->d_blas_ext), cudaFree(ctx->d_attr_ext);
    ctx->d_blas_ext = nullptr, ctx->d_attr_ext = nullptr;
    if (s->num_meshes > (uint32_t)kMaxBlas) {
        CU_TRY(ctx, cudaMalloc(&ctx->d_blas_ext, table.size() * sizeof(BlasConst)));
        CU_TRY(ctx, cudaMemcpy(ctx->d_blas_ext, table.data(), table.size() * sizeof(BlasConst), cudaMemcpyHostToDevice));
        CU_TRY(ctx, cudaMalloc(&ctx->d_attr_ext, attrs.size() * sizeof(BlasAttr)));
        CU_TRY(ctx, cudaMemcpy(ctx->d_attr_ext, attrs.data(), attrs.size() * sizeof(BlasAttr), cudaMemcpyHostToDevice));
        ctx->scene.blas_ext = ctx->d_blas_ext;
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

} // extern "C"

BlasConst nbvh_host::host_blas(const nbvh_ctx *ctx, uint32_t i)
{
    BlasConst b;
    memset(&b, 0, sizeof(b));
    if (i < ctx->blas_desc.size()) {
        b.nodes        = reinterpret_cast<const float4 *>(ctx->blas_desc[i].nodes);
        b.prim_indices = reinterpret_cast<const unsigned long long *>(ctx->blas_desc[i].prim_indices);
        b.vertices     = ctx->blas_desc[i].vertices;
        b.indices      = ctx->blas_desc[i].indices;
    }
    return b;
}

int nbvh_host::adopt_nbvh(nbvh_ctx *ctx, const void *nodes, uint32_t n, cudaMemcpyKind kind)
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
    ctx->dilated_with = -1.f, ctx->walk_version++;
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

extern "C" {

int nbvh_set_neural_bvh(nbvh_ctx *ctx, const nbvh_node32 *nodes, uint32_t n)
{
    DeviceGuard device_guard(ctx);
    if (!ctx || !nodes)
        return NBVH_ERR_INVALID_ARG;
    return adopt_nbvh(ctx, nodes, n, cudaMemcpyDeviceToDevice);
}

uint32_t nbvh_neural_bvh_num_nodes(const nbvh_ctx *ctx) { return ctx ? ctx->nbvh_nodes : 0; }

int nbvh_get_neural_bvh_host(const nbvh_ctx *ctx, nbvh_node32 *nodes_host, uint32_t n)
{
    DeviceGuard device_guard(ctx);
    if (!ctx || !nodes_host || n != ctx->nbvh_nodes || !ctx->d_nbvh)
        return NBVH_ERR_INVALID_ARG;
    return cudaMemcpy(nodes_host, ctx->d_nbvh, (size_t)n * 32, cudaMemcpyDeviceToHost) == cudaSuccess ? NBVH_OK : NBVH_ERR_CUDA;
}

uint64_t nbvh_n_params(const nbvh_ctx *ctx) { return ctx ? ctx->n_params : 0; }

int nbvh_load_params_host(nbvh_ctx *ctx, const uint16_t *fp16, uint64_t n)
{
    DeviceGuard device_guard(ctx);
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
    DeviceGuard device_guard(ctx);
    if (!ctx || !fp16 || n != ctx->n_params)
        return NBVH_ERR_INVALID_ARG;
    return cudaMemcpy(fp16, ctx->d_params, n * sizeof(uint16_t), cudaMemcpyDeviceToHost) == cudaSuccess ? NBVH_OK : NBVH_ERR_CUDA;
}

// tcnn initial values: grid U(-1e-4, 1e-4) (grid.h:1059-1062), MLP Xavier uniform (fully_fused_mlp.cu:866-891).
// Own splitmix64 stream (tcnn uses pcg32 seed 1337; initial weights are statistically, not bitwise, equivalent).
int nbvh_init_params(nbvh_ctx *ctx, uint64_t seed)
{
    DeviceGuard device_guard(ctx);
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
    DeviceGuard device_guard(ctx);
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
    DeviceGuard device_guard(ctx);
    if (!ctx)
        return NBVH_ERR_INVALID_ARG;
    ctx->hits = hits;
    return NBVH_OK;
}

int nbvh_selftest_division(nbvh_ctx *ctx, uint64_t seed, uint64_t n_trials, uint64_t *mismatches_host)
{
    DeviceGuard device_guard(ctx);
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
    DeviceGuard device_guard(ctx);
    if (!ctx || !out_host || n > CNT_COUNT)
        return NBVH_ERR_INVALID_ARG;
    CU_TRY(ctx, cudaMemcpy(out_host, ctx->d_counters, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost));
    return NBVH_OK;
}

int nbvh_reset_counters(nbvh_ctx *ctx)
{
    DeviceGuard device_guard(ctx);
    if (!ctx)
        return NBVH_ERR_INVALID_ARG;
    CU_TRY(ctx, cudaMemset(ctx->d_counters, 0, sizeof(unsigned long long) * CNT_COUNT));
    ctx->total_ms = 0.f, ctx->launches = 0;
    return NBVH_OK;
}

int nbvh_enable_kernel_timing(nbvh_ctx *ctx, int enable)
{
    DeviceGuard device_guard(ctx);
    if (!ctx)
        return NBVH_ERR_INVALID_ARG;
    ctx->timing = enable != 0;
    return NBVH_OK;
}

int nbvh_get_kernel_time(nbvh_ctx *ctx, float *total_ms, uint32_t *launches)
{
    DeviceGuard device_guard(ctx);
    if (!ctx)
        return NBVH_ERR_INVALID_ARG;
    if (total_ms)
        *total_ms = ctx->total_ms;
    if (launches)
        *launches = ctx->launches;
    return NBVH_OK;
}

int nbvh_set_queue_order(nbvh_ctx *ctx, uint32_t t0, uint32_t t1, uint32_t t2, uint32_t t3, uint32_t flags)
{
    if (!ctx)
        return NBVH_ERR_INVALID_ARG;
    if (t0 > 255u || t0 < t1 || t1 < t2 || t2 < t3)
        return fail(ctx, NBVH_ERR_INVALID_ARG, "nbvh_set_queue_order: class limits must be descending and below 256");
    ctx->long_chain_thr = (t0 << 24) | (t1 << 16) | (t2 << 8) | t3;
    ctx->chain_pred     = flags & 7u;
    return NBVH_OK;
}

int nbvh_enable_stage_timing(nbvh_ctx *ctx, uint32_t max_frames)
{
    DeviceGuard device_guard(ctx);
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
    DeviceGuard device_guard(ctx);
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

} // extern "C"

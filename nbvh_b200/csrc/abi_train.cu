// abi_train.cu -- online-optimisation entry points of the C ABI: training-sample generation, tree-cut learning, the training
// step (forward / loss / backward / Adam).
#include "host_context.hpp"

extern "C" {

int nbvh_sampler_walk_ahead(nbvh_ctx *ctx, uint32_t steps, uint32_t step_stride)
{
    if (!ctx || steps == 0 || step_stride == 0)
        return ctx ? fail(ctx, NBVH_ERR_INVALID_ARG, "nbvh_sampler_walk_ahead: steps and stride must be >= 1") : NBVH_ERR_INVALID_ARG;
    ctx->walk_ahead_steps = steps, ctx->walk_ahead_stride = step_stride;
    return NBVH_OK;
}

int nbvh_optimisation_data_generation(nbvh_ctx *ctx, uint32_t training_step, uint32_t batch_size, uint32_t log2_bvh_max_traversals,
                                      uint32_t *sampled_node_indices, float current_max_loss, const nbvh_tree_cut_data *tree_cut_data,
                                      float *inputs, float *targets, void *stream)
{
    DeviceGuard device_guard(ctx);
    if (!ctx || !sampled_node_indices || !inputs || !targets)
        return NBVH_ERR_INVALID_ARG;
    if (!ctx->d_nbvh || !ctx->scene_set)
        return fail(ctx, NBVH_ERR_STATE, "optimisation_data_generation: neural BVH and scene must be set");
    const nbvh_renderer_constants &bc = ctx->bvh_const;
    if (batch_size == 0 || (batch_size & (batch_size - 1)) || log2_bvh_max_traversals > 31 || (batch_size >> log2_bvh_max_traversals) == 0)
        return fail(ctx, NBVH_ERR_INVALID_ARG, "batch_size must be a power of two >= 2^log2_bvh_max_traversals");
    if (bc.ellipsoidal_bvh)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "sampler: ellipsoidal_bvh is not implemented");
    if (ctx->module_const.input_segment_direction_dim != 0 || ctx->module_const.apply_encoding_trafo)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "sampler: directional inputs / apply_encoding_trafo are not implemented");
    if ((uint32_t)ctx->pt_const.depth_learning_mode > 3u)
        return fail(ctx, NBVH_ERR_INVALID_ARG, "depth_learning_mode out of range");
    if (bc.leaf_boundary_sampling_ratio != 0.f && log2_bvh_max_traversals != 0)
        return fail(ctx, NBVH_ERR_INVALID_ARG, "leaf_boundary_sampling_ratio needs log2_bvh_max_traversals == 0 (neural_bvh_module_kernels.cu:149)");
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
    sp.nbvh.idx_mask   = 0x7FFFFFFFu; // plain stack entries (device_core.cuh)
    sp.blas            = host_blas(ctx, bc.neural_bvh_idx);
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
        sp.bbox_max[i] = ctx->scene_bbox.max[i];
        sp.bbox_ext[i] = ctx->scene_bbox.max[i] - ctx->scene_bbox.min[i];
    }
    sp.uniform_points               = ctx->module_const.segment_distribution == 2 ? 1u : 0u;
    sp.depth_mode                   = (uint32_t)ctx->pt_const.depth_learning_mode;
    sp.view_dependent_sampling      = bc.view_dependent_sampling ? 1u : 0u;
    sp.view_dependent_ratio         = bc.view_dependent_ratio;
    sp.leaf_boundary_sampling_ratio = bc.leaf_boundary_sampling_ratio;
    sp.cam                          = make_camera_const(ctx);
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
    const uint32_t want_depth_dim = sp.depth_mode >= 2u ? 3u : 1u; // Position / LocalPosition targets are points
    if ((sp.depth_dim != 0u && sp.depth_dim != want_depth_dim) || (sp.normal_dim != 0u && sp.normal_dim != 3u) ||
        (sp.albedo_dim != 0u && sp.albedo_dim != 3u) ||
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
    // view-dependent and leaf-boundary sampling change the number of draws per origin / per slot: the general kernel follows the
    // reference's loop literally; everything else takes the two-pass generator
    if (sp.n_sub_batch == 1 && !sp.view_dependent_sampling && sp.leaf_boundary_sampling_ratio == 0.f) {
        // two-pass generator (sampler.cu).  The walk depends on the tree-cut, the sampler constants and the RNG only, so it may run for
        // several steps at once (nbvh_sampler_walk_ahead); this step's statistics only enter the roulette of the target pass.
        cudaStream_t st = (cudaStream_t)stream;
        const uint32_t stride = ctx->walk_ahead_stride ? ctx->walk_ahead_stride : 1u;
        uint32_t k = 0;
        const bool hit = ctx->ahead_valid && ctx->ahead_version == ctx->walk_version && ctx->ahead_batch == batch_size &&
                         ctx->ahead_stride == stride && training_step >= ctx->ahead_first_step &&
                         (training_step - ctx->ahead_first_step) % stride == 0u &&
                         (k = (training_step - ctx->ahead_first_step) / stride) < ctx->ahead_steps;
        if (!hit) {
            const uint32_t n_steps = ctx->walk_ahead_steps ? ctx->walk_ahead_steps : 1u;
            const uint64_t n_items = (uint64_t)batch_size * n_steps;
            if (n_items > 0xFFFFFFFFull)
                return fail(ctx, NBVH_ERR_INVALID_ARG, "walk-ahead: batch_size x steps exceeds 2^32 slots");
            if (ctx->sampler_item_capacity < n_items) {
                CU_TRY(ctx, cudaDeviceSynchronize()); // an earlier generation (any stream) may still read the old list
                cudaFree(ctx->d_sampler_items);
                ctx->d_sampler_items = nullptr, ctx->sampler_item_capacity = 0;
                CU_TRY(ctx, cudaMalloc(&ctx->d_sampler_items, (size_t)n_items * 48u));
                ctx->sampler_item_capacity = n_items;
            }
            if (!ctx->d_sampler_counters)
                CU_TRY(ctx, cudaMalloc(&ctx->d_sampler_counters, 8 * sizeof(uint32_t)));
            if (!ctx->ahead_done)
                CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->ahead_done, cudaEventDisableTiming));
            sp.items          = ctx->d_sampler_items;
            sp.n_steps        = n_steps;
            sp.rng_index_step = stride * batch_size;
            CU_TRY(ctx, launch_sample_walk(sp, ctx->d_sampler_counters, ctx->num_sms, st));
            CU_TRY(ctx, cudaEventRecord(ctx->ahead_done, st));
            ctx->ahead_valid = true, ctx->ahead_first_step = training_step, ctx->ahead_stride = stride, ctx->ahead_steps = n_steps;
            ctx->ahead_batch = batch_size, ctx->ahead_version = ctx->walk_version;
            k = 0;
        }
        if (ctx->node_keep_capacity < ctx->nbvh_nodes) {
            CU_TRY(ctx, cudaDeviceSynchronize());
            cudaFree(ctx->d_node_keep);
            ctx->d_node_keep = nullptr, ctx->node_keep_capacity = 0;
            const uint32_t cap = ctx->nbvh_capacity > ctx->nbvh_nodes ? ctx->nbvh_capacity : ctx->nbvh_nodes;
            CU_TRY(ctx, cudaMalloc(&ctx->d_node_keep, (size_t)cap * sizeof(float)));
            ctx->node_keep_capacity = cap;
        }
        sp.node_keep = ctx->d_node_keep;
        sp.items     = ctx->d_sampler_items;
        sp.n_steps   = 1;
        CU_TRY(ctx, cudaStreamWaitEvent(st, ctx->ahead_done, 0));
        CU_TRY(ctx, launch_sample_targets(sp, k * batch_size, st));
        return NBVH_OK;
    }
    CU_TRY(ctx, launch_generate_batch(sp, ctx->num_sms, (cudaStream_t)stream));
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
    DeviceGuard device_guard(ctx);
    if (!ctx)
        return NBVH_ERR_INVALID_ARG;
    std::vector<nbvh_node32> blas;
    int rc = neural_blas_nodes_host(ctx, blas);
    if (rc != NBVH_OK)
        return rc;
    // build_and_upload_neural_bvh resets the level unconditionally (neural_bvh_renderer_accel.cpp:520): learning starts at the
    // coarsest LoD whatever the config file said (34 of the 35 shipped Chess configs omit the key)
    ctx->bvh_const.current_learned_lod = 6; // NUM_BVH_LODS - 1
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
    DeviceGuard device_guard(ctx);
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
    // traverse_neural_bvh_and_update_states also tracks the LoD the walk is in: an interior node tagged as a leaf of levels
    // t-2 .. current moves the walk to level t-3 (accel.cpp:126-134); the coarsest level any childless node is reached at
    // decides where learning resumes (accel.cpp:496-501)
    int min_traversed_lod = 6; // NUM_BVH_LODS - 1
    struct Item { uint32_t bvh, neural, depth; int lod; };
    std::vector<Item> stack{{0u, 0u, 0u, 6}};
    while (!stack.empty()) {
        const Item it = stack.back();
        stack.pop_back();
        if (it.neural >= n || it.bvh >= blas.size())
            return fail(ctx, NBVH_ERR_STATE, "the tree-cut is not a cut of the neural mesh's BLAS");
        map[it.neural]       = it.bvh;
        cut[it.neural].depth = it.depth;
        max_depth            = it.depth > max_depth ? it.depth : max_depth;
        if (nodes[it.neural].left_or_prim_offset == 0xFFFFFFFFu) {
            min_traversed_lod = it.lod < min_traversed_lod ? it.lod : min_traversed_lod;
            continue;
        }
        if (blas[it.bvh].prim_count_or_left > 0u)
            return fail(ctx, NBVH_ERR_STATE, "the tree-cut is not a cut of the neural mesh's BLAS");
        int lod = it.lod;
        const uint32_t tag = nodes[it.neural].prim_count_or_left;
        for (int l = it.lod; l >= 0; --l)
            if (tag > 0u && tag <= 2u + (uint32_t)l) // is_leaf_at_lod(l)
                --lod;
        if (lod < 0)
            return fail(ctx, NBVH_ERR_STATE, "the tree-cut nests more LoD levels than NUM_BVH_LODS");
        stack.push_back({blas[it.bvh].left_or_prim_offset + 1u, nodes[it.neural].left_or_prim_offset + 1u, it.depth + 1u, lod});
        stack.push_back({blas[it.bvh].left_or_prim_offset, nodes[it.neural].left_or_prim_offset, it.depth + 1u, lod});
    }
    const int nl = ctx->bvh_const.num_lods_learned;
    ctx->bvh_const.current_learned_lod = (min_traversed_lod == 7 - nl && nl > 1) ? 0 : min_traversed_lod;
    return upload_tree_cut_state(ctx, nodes, cut, map, max_depth);
}

const nbvh_tree_cut_data *nbvh_tree_cut_data_device(const nbvh_ctx *ctx) { return ctx && ctx->tree_cut_ready ? ctx->d_tree_cut : nullptr; }
const nbvh_node32 *nbvh_neural_bvh_device(const nbvh_ctx *ctx) { return ctx ? reinterpret_cast<const nbvh_node32 *>(ctx->d_nbvh) : nullptr; }

// test / restore hook: overwrite the per-node statistics (the reference restores them when it reloads a training state)
int nbvh_set_tree_cut_host(nbvh_ctx *ctx, const nbvh_tree_cut_data *cut_host, uint32_t n)
{
    DeviceGuard device_guard(ctx);
    if (!ctx || !cut_host || !ctx->tree_cut_ready || n != ctx->nbvh_nodes)
        return NBVH_ERR_INVALID_ARG;
    CU_TRY(ctx, cudaMemcpy(ctx->d_tree_cut, cut_host, (size_t)n * sizeof(nbvh_tree_cut_data), cudaMemcpyHostToDevice));
    return NBVH_OK;
}

int nbvh_get_tree_cut_host(const nbvh_ctx *ctx, nbvh_tree_cut_data *cut_host, uint32_t *node_map_host, uint32_t n)
{
    DeviceGuard device_guard(ctx);
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
                                                 host_blas(ctx, blas_idx).nodes, num_splits, (uint32_t)ctx->bvh_const.current_learned_lod,
                                                 ctx->d_tc_workspace, ctx->tc_workspace_bytes, ctx->d_tc_result, st));
    TreeCutResult r;
    CU_TRY(ctx, cudaMemcpyAsync(&r, ctx->d_tc_result, sizeof(r), cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    if (r.n_split) {
        ctx->nbvh_nodes   = n + 2u * r.n_split;
        ctx->dilated_with = -1.f, ctx->walk_version++; // the pre-dilated copy is rebuilt before the next traversal
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
    DeviceGuard device_guard(ctx);
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
    DeviceGuard device_guard(ctx);
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
    DeviceGuard device_guard(ctx);
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
    DeviceGuard device_guard(ctx);
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
    ctx->dilated_with = -1.f, ctx->walk_version++;
    return refresh_dilated(ctx);
}

uint32_t nbvh_tree_cut_max_depth(const nbvh_ctx *ctx) { return ctx ? ctx->tree_cut_max_depth : 0; }

// ---- online optimisation -------------------------------------------------------------------------------------
static int ensure_training_state(nbvh_ctx *ctx, uint32_t batch, cudaStream_t st)
{
    if (!ctx->params_loaded)
        return fail(ctx, NBVH_ERR_STATE, "network parameters not loaded (nbvh_init_params / nbvh_load_network_file)");
    // TMEM budget of k_train_mlp: working accumulators + one fp32 accumulator per weight matrix
    if (train_tmem_cols(ctx->in_k, ctx->n_layers, 1u) > 512u)
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
    if (!ctx->d_wgrad_partial)
        CU_TRY(ctx, cudaMalloc(&ctx->d_wgrad_partial, (size_t)ctx->num_sms * ctx->n_mlp_params * sizeof(float)));
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
    DeviceGuard device_guard(ctx);
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
    cudaEvent_t *tev = ctx->train_ev_fb < ctx->train_ev_steps ? ctx->train_ev.data() + 7u * ctx->train_ev_fb : nullptr;
    if (tev)
        CU_TRY(ctx, cudaEventRecord(tev[0], st));
    CU_TRY(ctx, launch_encode_features(nc, in, batch, ctx->module_const.input_network_dims, ctx->d_feat, ctx->num_sms, st));
    if (tev)
        CU_TRY(ctx, cudaEventRecord(tev[1], st));
    TrainParams tp;
    memset(&tp, 0, sizeof(tp));
    tp.feat = ctx->d_feat, tp.dfeat = ctx->d_dfeat, tp.target = target, tp.per_sample_loss = per_sample_loss;
    tp.wgrad = ctx->d_grads, tp.loss_sum = loss_dev, tp.batch = batch;
    tp.wgrad_partial = ctx->d_wgrad_partial, tp.n_matrix = (uint32_t)ctx->n_mlp_params;
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
    if (tev)
        CU_TRY(ctx, cudaEventRecord(tev[2], st));
    // data parallel (comm.cu): the MLP gradients (and the loss scalar) are summed across the ranks while the grid backward runs,
    // the grid table afterwards; `st` resumes when both buckets are reduced, so the caller (and Adam) only ever see global sums
    if ((rc = dp_allreduce_bucket(ctx, 0, ctx->d_grads, ctx->n_mlp_params, st)) != NBVH_OK)
        return rc;
    CU_TRY(ctx, launch_grid_backward(nc, in, ctx->d_dfeat, batch, ctx->module_const.input_network_dims, ctx->d_grads + ctx->n_mlp_params,
                                     ctx->num_sms, st));
    if (tev)
        CU_TRY(ctx, cudaEventRecord(tev[3], st));
    if ((rc = dp_allreduce_bucket(ctx, 1, ctx->d_grads + ctx->n_mlp_params, ctx->n_params - ctx->n_mlp_params, st)) != NBVH_OK)
        return rc;
    if (ctx->nccl_comm && ctx->dp_world > 1 && (rc = nbvh_allreduce_buffer(ctx, loss_dev, 1, NBVH_DTYPE_F32, ctx->comm_stream)) != NBVH_OK)
        return rc;
    if ((rc = dp_join(ctx, st)) != NBVH_OK)
        return rc;
    if (tev) {
        CU_TRY(ctx, cudaEventRecord(tev[4], st));
        ctx->train_ev_fb++;
    }
    kt.stop();
    return NBVH_OK;
}

float *nbvh_gradients(nbvh_ctx *ctx) { return ctx ? ctx->d_grads : nullptr; }

int nbvh_set_gradient_buffer(nbvh_ctx *ctx, float *grads)
{
    DeviceGuard device_guard(ctx);
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
    DeviceGuard device_guard(ctx);
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
    cudaEvent_t *tev = ctx->train_ev_opt < ctx->train_ev_steps ? ctx->train_ev.data() + 7u * ctx->train_ev_opt : nullptr;
    if (tev)
        CU_TRY(ctx, cudaEventRecord(tev[5], st));
    CU_TRY(ctx, launch_adam(ap, st));
    if (tev) {
        CU_TRY(ctx, cudaEventRecord(tev[6], st));
        ctx->train_ev_opt++;
    }
    ctx->opt_step++;
    ctx->training_step++;
    return NBVH_OK;
}

int nbvh_training_step(nbvh_ctx *ctx, const float *in, const float *target, uint32_t batch, float *per_sample_loss, float *loss_host,
                       void *stream)
{
    DeviceGuard device_guard(ctx);
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

int nbvh_enable_training_stage_timing(nbvh_ctx *ctx, uint32_t max_steps)
{
    DeviceGuard device_guard(ctx);
    if (!ctx)
        return NBVH_ERR_INVALID_ARG;
    for (cudaEvent_t e : ctx->train_ev)
        cudaEventDestroy(e);
    ctx->train_ev.clear();
    ctx->train_ev_steps = ctx->train_ev_fb = ctx->train_ev_opt = 0;
    for (uint32_t i = 0; i < 7u * max_steps; i++) {
        cudaEvent_t e;
        CU_TRY(ctx, cudaEventCreate(&e));
        ctx->train_ev.push_back(e);
    }
    ctx->train_ev_steps = max_steps;
    return NBVH_OK;
}

int nbvh_get_training_stage_times(nbvh_ctx *ctx, float ms_total_out[5], uint32_t *steps)
{
    DeviceGuard device_guard(ctx);
    if (!ctx || !ms_total_out)
        return NBVH_ERR_INVALID_ARG;
    const uint32_t n = ctx->train_ev_fb < ctx->train_ev_opt ? ctx->train_ev_fb : ctx->train_ev_opt;
    float sum[5]     = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (uint32_t i = 0; i < n; i++) {
        cudaEvent_t *e = ctx->train_ev.data() + 7u * i;
        CU_TRY(ctx, cudaEventSynchronize(e[6]));
        float ms = 0.f;
        for (int k = 0; k < 4; k++) { // encode | mlp | grid backward | wait for the gradient exchange
            CU_TRY(ctx, cudaEventElapsedTime(&ms, e[k], e[k + 1]));
            sum[k] += ms;
        }
        CU_TRY(ctx, cudaEventElapsedTime(&ms, e[5], e[6])); // Adam
        sum[4] += ms;
    }
    for (int k = 0; k < 5; k++)
        ms_total_out[k] = sum[k];
    if (steps)
        *steps = n;
    ctx->train_ev_fb = ctx->train_ev_opt = 0;
    return NBVH_OK;
}

} // extern "C"

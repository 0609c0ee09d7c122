// abi_render.cu -- inference entry points of the C ABI: network queries and parity hooks, frames (nbvh_render*), ray batches,
// textures / environment map, image metrics, BLAS / TLAS builds.
#include "host_context.hpp"

extern "C" {

int nbvh_query(nbvh_ctx *ctx, const float *in, float *out, uint32_t n, void *stream)
{
    DeviceGuard device_guard(ctx);
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
    DeviceGuard device_guard(ctx);
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
    DeviceGuard device_guard(ctx);
    if (!ctx || (n && (!o || !d || !t || !u || !v || !prim)))
        return NBVH_ERR_INVALID_ARG;
    if (!ctx->scene_set || blas_idx >= ctx->scene.n_blas)
        return fail(ctx, NBVH_ERR_STATE, "scene not set or BLAS index out of range");
    KernelTimer kt(ctx, (cudaStream_t)stream);
    CU_TRY(ctx, launch_blas_intersect(host_blas(ctx, blas_idx), ctx->blas_safe[blas_idx] ? 1u : 0u, o, d, maxt, n, t, u, v, prim, ctx->d_counters,
                                      (cudaStream_t)stream));
    kt.stop();
    return NBVH_OK;
}

} // extern "C"

CameraConst nbvh_host::make_camera_const(const nbvh_ctx *ctx)
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

extern "C" {

int nbvh_raygen(nbvh_ctx *ctx, uint32_t first_pixel, uint32_t n, int sample_idx, int sample_offset, float *o, float *d,
                void *stream)
{
    DeviceGuard device_guard(ctx);
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
    DeviceGuard device_guard(ctx);
    if (!ctx || (n && (!o || !d || !mint || !maxt || !out)))
        return NBVH_ERR_INVALID_ARG;
    if (ctx->module_const.input_network_dims != 3u * ctx->module_const.input_segment_points)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "encode_segments: directional inputs are not supported");
    if (ctx->module_const.segment_distribution == 2)
        return fail(ctx, NBVH_ERR_UNSUPPORTED, "encode_segments: Uniform segment points depend on each ray's generator state (render / trace_rays implement them)");
    CU_TRY(ctx, launch_encode_segments(make_encode_const(ctx), o, d, mint, maxt, n, out, (cudaStream_t)stream));
    return NBVH_OK;
}

int nbvh_create_texture_rgba8(nbvh_ctx *ctx, const uint8_t *pixels_host, uint32_t w, uint32_t h, uint64_t *tex_out)
{
    DeviceGuard device_guard(ctx);
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
    DeviceGuard device_guard(ctx);
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
    DeviceGuard device_guard(ctx);
    if (!ctx)
        return NBVH_ERR_INVALID_ARG;
    ctx->envmap = (cudaTextureObject_t)envmap_texture;
    return NBVH_OK;
}

int nbvh_destroy_textures(nbvh_ctx *ctx)
{
    DeviceGuard device_guard(ctx);
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
    DeviceGuard device_guard(ctx);
    return render_impl(ctx, accumulated_spp, sample_offset, first_pixel, pixel_count, 1u, 0u, stream);
}

int nbvh_render_interleaved(nbvh_ctx *ctx, uint32_t accumulated_spp, uint32_t sample_offset, uint32_t shard_index, uint32_t shard_count,
                            void *stream)
{
    DeviceGuard device_guard(ctx);
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
    DeviceGuard device_guard(ctx);
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
        rc = check_wavefront_modes(ctx);
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
    rp.dec                     = make_decode_const(ctx);
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
    // drain order by expected chain length (kernels.cuh: class_lists): a 16-word header (render counters [0..3], class counts
    // [4..8]) followed by kChainClasses index lists of queue_capacity entries; camera rays only (bounce rays have no per-pixel history)
    uint32_t memset_words = 4;
    if (ctx->long_chain_thr != 0u && !ctx->mega_kernel) {
        uint32_t *&lists   = ctx->d_class_lists[alt ? ai + 1 : 0];
        size_t &lists_cap  = ctx->class_lists_capacity[alt ? ai + 1 : 0];
        if (lists_cap != queue_capacity) {
            CU_TRY(ctx, cudaDeviceSynchronize());
            cudaFree(lists);
            lists = nullptr, lists_cap = 0;
            CU_TRY(ctx, cudaMalloc(&lists, (16 + (size_t)kChainClasses * queue_capacity) * sizeof(uint32_t)));
            lists_cap = queue_capacity;
        }
        if (ctx->chain_hist_pixels != (size_t)total) {
            CU_TRY(ctx, cudaDeviceSynchronize());
            cudaFree(ctx->d_chain_hist);
            ctx->d_chain_hist = nullptr, ctx->chain_hist_pixels = 0;
            CU_TRY(ctx, cudaMalloc(&ctx->d_chain_hist, (size_t)total));
            CU_TRY(ctx, cudaMemset(ctx->d_chain_hist, 0, (size_t)total));
            ctx->chain_hist_pixels = (size_t)total;
        }
        chunk_ctr      = lists;
        rp.next_chunk  = chunk_ctr;
        rp.class_count = lists + 4, rp.class_lists = lists + 16;
        rp.chain_hist  = ctx->d_chain_hist;
        rp.chain_thr   = ctx->long_chain_thr, rp.chain_flags = ctx->chain_pred;
        memset_words   = 4 + kChainClasses;
    }
    rp.queue = queue, rp.queue_count = chunk_ctr + 1, rp.queue_cursor = chunk_ctr + 2;
    rp.queue_capacity   = (uint32_t)queue_capacity;
    rp.sample_idx    = (int)accumulated_spp;
    rp.sample_offset = (int)sample_offset;
    rp.accum = ctx->accum, rp.ldr = ctx->ldr, rp.hits = ctx->hits;
    rp.ema_accumulate = ctx->scene_const.ema_accumulate;
    rp.ema_weight     = ctx->scene_const.ema_weight;
    rp.counters       = ctx->d_counters;
    const NetConst nc = make_net_const(ctx);
    CU_TRY(ctx, cudaMemsetAsync(chunk_ctr, 0, memset_words * sizeof(uint32_t), st));
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
        pc.attr_ext  = ctx->d_attr_ext;
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
    DeviceGuard device_guard(ctx);
    if (!ctx || (n && (!o || !d || !hits)) || ray_depth == 0)
        return NBVH_ERR_INVALID_ARG;
    int rc = check_ready(ctx, true, true, true);
    if (rc == NBVH_OK)
        rc = check_wavefront_modes(ctx);
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
    rp.dec                     = make_decode_const(ctx);
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

// == compute_and_display_render_error + the mean (neural_bvh_renderer.cu:346-358): error of the accumulated image (accum / spp)
// against a reference RGBA32F image on the device.  mode = PostprocessMode: 3 AbsError, 4 RelAbsError, 5 MSEError, 6 RelMSEError.
int nbvh_image_error(nbvh_ctx *ctx, const float *reference_rgba32f, uint32_t accumulated_spp, int mode, int write_error_image,
                     float *mean_error_host, void *stream)
{
    DeviceGuard device_guard(ctx);
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
    DeviceGuard device_guard(ctx);
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
    DeviceGuard device_guard(ctx);
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
    DeviceGuard device_guard(ctx);
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
    DeviceGuard device_guard(ctx);
    if (!ctx || slot >= (uint32_t)kPipeSlots || !ctx->pipe_copied[slot])
        return NBVH_ERR_INVALID_ARG;
    CU_TRY(ctx, cudaEventSynchronize(ctx->pipe_copied[slot]));
    return NBVH_OK;
}

} // extern "C"

// Device-side TLAS build (f3): == CudaBackend::build_tlas (cuda_backend.cpp:312-383).  mesh_bvhs_device: the DEVICE array of descriptors
// that nbvh_scene_data.mesh_bvhs will point to (each BLAS already built: its root box is read from nodes[0]); neural_meshes_host lists
// the meshes represented by the N-BVH (the reference's m_neural_bvh_indices).  Outputs (device, caller-owned): 2 * num_meshes - 1
// TLASNode records and num_meshes blas_indices.
int nbvh_build_tlas(nbvh_ctx *ctx, const nbvh_cuda_blas *mesh_bvhs_device, uint32_t num_meshes, const uint32_t *neural_meshes_host,
                    uint32_t n_neural, nbvh_node32 *tlas_nodes_out, uint32_t *blas_indices_out, void *stream)
{
    DeviceGuard device_guard(ctx);
    if (!ctx || !mesh_bvhs_device || !tlas_nodes_out || !blas_indices_out || num_meshes == 0 || (n_neural && !neural_meshes_host))
        return NBVH_ERR_INVALID_ARG;
    std::vector<uint32_t> flags(num_meshes, 0u);
    for (uint32_t i = 0; i < n_neural; i++) {
        if (neural_meshes_host[i] >= num_meshes)
            return fail(ctx, NBVH_ERR_INVALID_ARG, "neural mesh index out of range");
        flags[neural_meshes_host[i]] = 1u;
    }
    // scratch: the sweep's per-mesh arrays + the neural flags (shares the grow-only scratch of nbvh_build_blas)
    const size_t ws_bytes = build_tlas_workspace_bytes(num_meshes), bytes = ws_bytes + (size_t)num_meshes * sizeof(uint32_t);
    if (ctx->build_ws_bytes < bytes) {
        CU_TRY(ctx, cudaDeviceSynchronize());
        cudaFree(ctx->d_build_ws);
        ctx->d_build_ws = nullptr, ctx->build_ws_bytes = 0;
        CU_TRY(ctx, cudaMalloc(&ctx->d_build_ws, bytes));
        ctx->build_ws_bytes = bytes;
    }
    uint32_t *flags_dev = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(ctx->d_build_ws) + ws_bytes);
    CU_TRY(ctx, cudaMemcpyAsync(flags_dev, flags.data(), flags.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    CU_TRY(ctx, cudaStreamSynchronize((cudaStream_t)stream)); // `flags` leaves scope
    CU_TRY(ctx, launch_build_tlas(mesh_bvhs_device, num_meshes, flags_dev, ctx->d_build_ws, reinterpret_cast<float4 *>(tlas_nodes_out),
                                  blas_indices_out, (cudaStream_t)stream));
    return NBVH_OK;
}

// == NeuralBVHRenderer::compute_and_display_flip_error (neural_bvh_renderer.cu:34-96) without the colour-mapped display: mean LDR-FLIP of
// the accumulated image (accum / accumulated_spp) against a device RGBA32F reference image of the framebuffer's size; ppd <= 0 selects the
// reference's viewing conditions (0.7 m from a 0.7 m wide 3840-pixel monitor: 67.02 pixels per degree).  error_map_out: optional device
// array of W x H floats (the per-pixel FLIP error the reference colour-maps with magma).
int nbvh_flip_error(nbvh_ctx *ctx, const float *reference_rgba32f, uint32_t accumulated_spp, float ppd, float *error_map_out,
                    float *mean_flip_host, void *stream)
{
    DeviceGuard device_guard(ctx);
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
    DeviceGuard device_guard(ctx);
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

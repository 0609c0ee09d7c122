// kernels.cuh -- kernel parameter blocks shared between kernels.cu (device) and context.cu (host)
#pragma once
#include "device_core.cuh"
#include "mlp_tile.cuh"
#include "../../include/nbvh_b200.h"

namespace nbvh {

constexpr int kMaxBlas = 16;
constexpr int kMaxPathDepth = 64; // max_depth bound of the multi-bounce path (queue counters per bounce)

enum Counter {
    CNT_RAYS = 0,
    CNT_QUERIES,
    CNT_NBVH_STEPS,
    CNT_BLAS_STEPS,
    CNT_TRI_TESTS,
    CNT_TILE_ITERS,
    CNT_QUEUED,
    CNT_NBVH_STEPS_NEURAL,
    // k_neural timeline (globaltimer ns): start = ~min, first queue exhaustion = ~min, end = max, summed CTA lifetimes
    CNT_T_START_INV,
    CNT_T_EXHAUST_INV,
    CNT_T_END,
    CNT_T_CTA_SUM,
    // k_neural phase clocks (NBVH_PHASE_TIMING builds only; SM cycles summed over tiles): walk / tile barrier / encode / MLP while the
    // queue still has entries, then the same four during the drain, and the tile iterations of the drain
    CNT_PH_WALK, CNT_PH_SYNC, CNT_PH_ENCODE, CNT_PH_MLP,
    CNT_PH_WALK_D, CNT_PH_SYNC_D, CNT_PH_ENCODE_D, CNT_PH_MLP_D,
    CNT_ITERS_DRAIN,
    CNT_COUNT
};

struct SceneConst {
    TlasConst tlas;
    BlasConst blas[kMaxBlas];  // the descriptors of scenes with up to kMaxBlas meshes travel in the kernel parameters ...
    const BlasConst *blas_ext; // ... larger scenes (the reference's TLAS has no bound, cuda_backend.cpp:312-385) read a device table
    uint32_t n_blas;
};

__device__ __forceinline__ BlasConst scene_blas(const SceneConst &s, uint32_t i)
{
    if (s.blas_ext) {
        const uint4 *q = reinterpret_cast<const uint4 *>(s.blas_ext + i);
        const uint4 a = __ldg(q), b = __ldg(q + 1);
        BlasConst r;
        r.nodes        = reinterpret_cast<const float4 *>(((unsigned long long)a.y << 32) | a.x);
        r.prim_indices = reinterpret_cast<const unsigned long long *>(((unsigned long long)a.w << 32) | a.z);
        r.vertices     = reinterpret_cast<const float *>(((unsigned long long)b.y << 32) | b.x);
        r.indices      = reinterpret_cast<const uint32_t *>(((unsigned long long)b.w << 32) | b.z);
        return r;
    }
    return s.blas[i];
}

// A ray with a pending neural leaf interval: pixel id, closest classical hit so far, the interval and the N-BVH stack
// (a: pixel, best_t, valid_hit, u | b: v, blas_idx, prim, stack size | c: seg_mint, seg_maxt, -, - | st: 24 stack entries)
struct NeuralQueueEntry {
    uint4 a, b, c;
    uint4 st[kNeuralStackSize / 4];
};

// Multi-bounce path state (max_depth > 1), one per pixel slot of the shard:
//   s0 = (o.x, o.y, o.z, d.x)  s1 = (d.y, d.z, throughput.x, throughput.y)  s2 = (throughput.z, illumination.xyz)
//   s3 = (rng state, ray depth -- 0 = path finished, pixel, -)
struct PathState {
    float4 s0, s1, s2;
    uint4 s3;
};
// Closest hit of the path's current ray == HybridIntersectionResult (path_module.h:104-210):
//   h0 = (t, valid_hit bits, normal_or_uv.xy)  h1 = (normal_or_uv.z, albedo_or_ids.xyz)
struct PathHit {
    float4 h0, h1;
};

// background_hit_eval constants (scene_constants + envmap_texture, path_inference_kernels.cu:37-59)
struct BackgroundConst {
    uint32_t envmap_eval, skylike_background;
    float rot_cos, rot_sin; // glm::rotate(mat4(1), radians(envmap_rotation), +Y): x' = c x + s z, z' = -s x + c z
    float strength;
    cudaTextureObject_t envmap; // RGBA32F lat-long map, linear filter, clamp; 0 = white (the reference's fallback image)
};

constexpr int kChainClasses = 5;

struct RenderParams {
    CameraConst cam;
    NbvhConst nbvh;
    SceneConst scene;
    EncodeConst enc;
    DecodeConst dec;
    uint32_t ignore_neural_blas;
    uint32_t max_depth;
    BackgroundConst bg; // radiance of rays that escape
    // work distribution
    uint32_t first_pixel, pixel_count; // shard of the frame rendered by this launch
    uint32_t spp_count; // samples per launch (multi-sample path launches: slot = sample * pixel_count + pixel slot; 0/1 = one)
    uint32_t chunk_stride, chunk_offset; // interleaved shard: every chunk_stride-th 32-pixel chunk from chunk_offset (0/1 = contiguous)
    uint32_t *next_chunk;               // global work counter (32-pixel chunks)
    NeuralQueueEntry *queue;            // neural ray queue filled by k_primary, drained by k_neural
    uint32_t *queue_count, *queue_cursor;
    uint32_t queue_capacity;
    // Drain order of the queue (scheduling only -- results never depend on it).  class_lists != null: k_primary files every queued
    // ray under one of kChainClasses classes by the query-chain length it is EXPECTED to need, k_neural drains class 0 (longest)
    // first.  class_lists[k * queue_capacity + i] = queue slot of the i-th ray of class k; class_count[k] = rays in class k.
    // The expectation is the chain the pixel's camera ray needed in the last frame: chain_hist[pixel] (0 = unknown), written by
    // k_neural; chain_thr packs the class limits t0 >= t1 >= t2 >= t3 (one byte each, t0 in the top byte).
    uint32_t *class_lists, *class_count;
    uint8_t *chain_hist;
    uint32_t chain_thr, chain_flags; // flags: bit 0 = predict from the 3x3 neighbourhood's maximum, bit 1 = decaying maximum
    int sample_idx, sample_offset;
    // outputs
    float4 *accum;          // RGBA32F accumulation buffer (linear memory replaces the GL surface)
    uchar4 *ldr;            // sRGB8 framebuffer
    nbvh_hit_record *hits;  // optional per-pixel primary-hit records
    uint32_t ema_accumulate;
    float ema_weight;
    unsigned long long *counters;
    // multi-bounce mode (path.cu): queue entries carry the path slot instead of the pixel, rays live in `paths`
    PathState *paths;
    PathHit *path_hits;
    uint32_t ray_depth; // depth of the rays in flight (1 = camera rays): selects the visibility threshold
};

// smem carve-up of the tile kernels
struct SmemLayout {
    uint32_t w_off, a_off, bar_off, tmem_off, total;
};

__host__ __device__ inline SmemLayout smem_layout(uint32_t wimg_bytes, uint32_t in_k, uint32_t tiles)
{
    SmemLayout s;
    s.w_off    = 0;
    s.a_off    = (wimg_bytes + 127u) & ~127u;
    const uint32_t a_bytes = kTileRows * (in_k > (uint32_t)kNeurons ? in_k : (uint32_t)kNeurons) * 2u;
    s.bar_off  = s.a_off + tiles * a_bytes;
    s.tmem_off = s.bar_off + 8u * (tiles + 1u);
    s.total    = s.tmem_off + 16u;
    return s;
}

// The tile kernels take their dynamic shared-memory size from the network description of EACH launch (contexts with different
// network shapes may live in one process, on any device): the per-function limit is raised to the device's opt-in maximum once
// per context creation, and nothing about the size is kept in process-global state.
inline cudaError_t allow_max_dynamic_smem(const void *func)
{
    int dev = 0, max_optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return e;
    if ((e = cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess)
        return e;
    cudaFuncAttributes fa;
    if ((e = cudaFuncGetAttributes(&fa, func)) != cudaSuccess)
        return e;
    // the opt-in limit covers static + dynamic shared memory of the kernel
    return cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin - (int)fa.sharedSizeBytes);
}

// ---- online optimisation (training.cu)
struct LossConst { // PathTracingOutputConstants subset read by the composite loss
    uint32_t vis_dim, depth_dim, normal_dim, albedo_dim;
    uint32_t vis_loss, depth_loss, normal_loss, albedo_loss;
    float vis_scale, depth_scale, normal_scale, albedo_scale;
    uint32_t normal_linear, depth_linear, albedo_linear, vis_masking;
};

struct TrainParams {
    const uint8_t *feat;    // fp16 feature tiles, tile-canonical layout (one 128 x in_k K-major operand block per tile)
    uint8_t *dfeat;         // fp16 dL/dfeatures, same layout
    const float *target;    // batch x n_out
    float *per_sample_loss; // batch x 16 (ctx.L of tcnn) or nullptr
    float *wgrad;           // fp32 gradients of the MLP matrices
    float *wgrad_partial;   // gridDim.x x n_matrix: every CTA stores its TMEM-resident sums here, k_reduce_wgrad adds them up in a fixed
                            // order (148 CTAs adding atomically to the same 19 k addresses cost a quarter of the kernel); null: atomics
    uint32_t n_matrix;      // number of MLP matrix weights
    float *loss_sum;        // device accumulator of the batch loss
    uint32_t batch, n_total;
    float loss_scale;
    LossConst loss;
};

struct AdamParams {
    uint64_t n_params, n_matrix;
    float loss_scale, lr, beta1, beta2, epsilon, l2_reg;
    const float *grads;
    float *master, *m, *v;
    uint32_t *steps;
    __half *params;
    uint8_t *wimg;
    uint32_t in_k, n_layers;
    uint32_t layer_off[kMaxLayers];
};

cudaError_t configure_training(uint32_t wimg_bytes, uint32_t in_k, uint32_t n_layers);
cudaError_t launch_encode_features(const NetConst &nc, const float *in, uint32_t batch, uint32_t n_in_dims, uint8_t *feat, int num_sms,
                                   cudaStream_t stream);
// TMEM columns of k_train_mlp (training.cu): per tile in flight one working accumulator, then one fp32 accumulator per weight matrix
__host__ __device__ constexpr uint32_t train_slot_cols(uint32_t in_k) { return in_k > 64u ? (in_k + 31u) & ~31u : 64u; }
__host__ __device__ constexpr uint32_t train_tmem_cols(uint32_t in_k, uint32_t n_layers, uint32_t slots)
{
    return slots * train_slot_cols(in_k) + in_k + (n_layers - 2u) * kNeurons + kOutPadded;
}
cudaError_t launch_train_mlp(const NetConst &nc, const TrainParams &tp, int num_sms, cudaStream_t stream);
cudaError_t launch_grid_backward(const NetConst &nc, const float *in, const uint8_t *dfeat, uint32_t batch, uint32_t n_in_dims,
                                 float *grid_grads, int num_sms, cudaStream_t stream);
cudaError_t launch_adam(const AdamParams &ap, cudaStream_t stream);
cudaError_t launch_half_to_float(const void *src, float *dst, uint64_t n, cudaStream_t stream);

cudaError_t launch_query(const NetConst &nc, const float *in, float *out, uint32_t n, uint32_t n_in_dims,
                         unsigned long long *counters, int num_sms, cudaStream_t stream);
cudaError_t launch_trace_intervals(const NbvhConst &bc, uint32_t nodes_safe, const float *o, const float *d, uint32_t n, uint32_t max_intervals,
                                   uint32_t *counts, uint32_t *node_idx, float *mint, float *maxt,
                                   unsigned long long *counters, cudaStream_t stream);
cudaError_t launch_blas_intersect(const BlasConst &b, uint32_t nodes_safe, const float *o, const float *d, const float *maxt, uint32_t n, float *t,
                                  float *u, float *v, uint32_t *prim, unsigned long long *counters, cudaStream_t stream);
cudaError_t launch_raygen(const CameraConst &cam, uint32_t first, uint32_t n, int sample_idx, int sample_offset, float *o,
                          float *d, cudaStream_t stream);
cudaError_t launch_encode_segments(const EncodeConst &ec, const float *o, const float *d, const float *mint, const float *maxt, uint32_t n,
                                   float *out, cudaStream_t stream);
cudaError_t launch_check_nodes(const float4 *nodes, uint32_t n, uint32_t *ok_dev, cudaStream_t stream);
cudaError_t launch_selftest_division(uint64_t seed, uint32_t blocks, uint32_t iters, unsigned long long *mismatches, cudaStream_t stream);
cudaError_t launch_predilate(const float4 *src, float4 *dst, uint32_t n, float dilation, cudaStream_t stream);
cudaError_t launch_render(const RenderParams &rp, const NetConst &nc, int num_sms, cudaStream_t stream);
cudaError_t launch_render_split(const RenderParams &rp, const NetConst &nc, int num_sms, cudaStream_t stream, cudaEvent_t between);
cudaError_t configure_render(uint32_t wimg_bytes, uint32_t in_k);
cudaError_t launch_neural(const RenderParams &rp, const NetConst &nc, int num_sms, bool path_mode, cudaStream_t stream);
cudaError_t configure_kernels(uint32_t wimg_bytes, uint32_t in_k);
cudaError_t launch_prefetch_l2(const void *const *ptrs, const size_t *bytes, int n, int num_sms, cudaStream_t stream);
size_t build_blas_workspace_bytes(uint32_t n_tris);
cudaError_t launch_build_blas(const float *vertices, const uint32_t *indices, uint32_t n_tris, float4 *nodes, unsigned long long *prim_indices,
                              void *workspace, size_t workspace_bytes, uint32_t *max_depth_dev, cudaStream_t stream);
size_t build_tlas_workspace_bytes(uint32_t num_meshes);
cudaError_t launch_build_tlas(const void *mesh_bvhs_device, uint32_t num_meshes, const uint32_t *neural_flags, void *workspace, float4 *tlas_nodes,
                              uint32_t *blas_indices, cudaStream_t stream);
// LDR-FLIP (flip.cu)
size_t flip_workspace_bytes(uint32_t w, uint32_t h);
cudaError_t flip_set_ppd(float ppd, cudaStream_t stream);
cudaError_t launch_flip(const float4 *reference, const float4 *test, float test_divisor, uint32_t w, uint32_t h, void *workspace,
                        float *error_map, double *sum_dev, cudaStream_t stream);

} // namespace nbvh

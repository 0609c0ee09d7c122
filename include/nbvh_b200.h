/*
 * nbvh_b200.h -- C ABI of libnbvh_b200.so: the B200-native (sm_100a) neural ray-query hot path of N-BVH.
 *
 * This is the drop-in boundary for the reference's `NeuralBVHModule` / `NeuralPathModule`
 * (/root/reference/include/cuda/neural/neural_bvh_module.h:144-241,
 *  /root/reference/include/cuda/neural/path/path_module.h:256-345) and for the tiny-cuda-nn surface the
 * renderer touches through it (`network->inference`, `trainer->training_step`, `serialize/deserialize`,
 * `n_params`; /root/reference/src/cuda/neural/path/path_module.cu:92,
 * /root/reference/src/cuda/neural/neural_bvh_renderer.cu:172-175,
 * /root/reference/src/cuda/neural/neural_bvh_renderer_init.cu:385-404).
 *
 * Conventions: opaque context; every call returns 0 on success or a negative nbvh_status, and
 * nbvh_last_error() gives the text (no C++ exception crosses the boundary -- the reference throws
 * std::runtime_error from CUDA_CHECK, include/ntwr.h:30-54).  Pointers are DEVICE pointers unless the
 * parameter name ends in `_host`.  `stream` is a cudaStream_t passed as void* (0 = default stream).
 * There is no CPU fallback: every compute entry point fails with NBVH_ERR_CUDA when no sm_100 device
 * is usable.
 *
 * The POD constant blocks below are layout-identical to the reference's __constant__ structs so the
 * renderer can hand over the very bytes it uploads today with cudaMemcpyToSymbol
 * (neural_bvh_renderer_init.cu:93-103).
 */
#ifndef NBVH_B200_H
#define NBVH_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBVH_ABI_VERSION 1

typedef enum {
    NBVH_OK              = 0,
    NBVH_ERR_INVALID_ARG = -1,
    NBVH_ERR_CUDA        = -2,
    NBVH_ERR_IO          = -3,
    NBVH_ERR_PARSE       = -4,
    NBVH_ERR_UNSUPPORTED = -5,
    NBVH_ERR_STATE       = -6
} nbvh_status;

typedef struct nbvh_ctx nbvh_ctx;

/* ---- POD constant blocks (layout == reference) ------------------------------------------------ */

/* NeuralBVHRendererConstants, include/cuda/neural/neural_bvh_renderer_constants.cuh:13-47 */
typedef struct __attribute__((aligned(8))) {
    uint32_t neural_bvh_idx;
    int32_t num_lods_learned;
    int32_t current_learned_lod;
    int32_t rendered_lod;
    bool ignore_neural_blas;
    bool path_lod_switch;
    int32_t lod_switch;
    float leaf_boundary_sampling_ratio;
    float inference_mint_epsilon;
    float inference_maxt_epsilon;
    float learning_mint_epsilon;
    float learning_maxt_epsilon;
    float sampling_bbox_scaling;
    float fixed_node_dilation;
    float node_extent_dilation;
    float loss_exposure;
    bool show_loss_heuristic;
    bool ellipsoidal_bvh;
    bool rescale_ellipsoid;
    int32_t log2_bvh_max_traversals;
    bool sample_cosine_weighted;
    bool backface_bounce_training;
    int32_t max_training_bounces;
    bool training_node_survival;
    float survival_probability_scale;
    float min_rr_survival_prob;
    bool view_dependent_sampling;
    float view_dependent_ratio;
} nbvh_renderer_constants;

/* NeuralBVHModuleConstants, include/cuda/neural/neural_bvh_module.h:79-94 */
typedef struct __attribute__((aligned(8))) {
    uint32_t input_segment_point_dim;
    uint32_t input_segment_points;
    uint32_t input_segment_points_dim;
    uint32_t input_segment_direction_dim;
    uint32_t input_network_dims;
    int32_t segment_distribution; /* 0 EquidistantInterior, 1 EquidistantExterior, 2 Uniform */
    bool antialias_direction;
    bool merge_operator;
    bool apply_encoding_trafo;
} nbvh_module_constants;

/* PathTracingOutputConstants, include/cuda/neural/neural_bvh_module.h:243-274 */
typedef struct __attribute__((aligned(8))) {
    uint32_t output_visibility_dim;
    uint32_t output_depth_dim;
    uint32_t output_normal_dim;
    uint32_t output_albedo_dim;
    uint32_t output_network_dims;
    uint8_t visibility_loss; /* LossType, include/cuda/neural/utils/tcnn_config.h:26-35 */
    float visibility_loss_scale;
    uint8_t depth_loss;
    float depth_loss_scale;
    uint8_t normal_loss;
    float normal_loss_scale;
    uint8_t albedo_loss;
    float albedo_loss_scale;
    uint8_t depth_learning_mode; /* 0 LocalT, 1 WorldT, 2 Position, 3 LocalPosition */
    bool output_normal_linear_activation;
    bool output_depth_linear_activation;
    bool output_albedo_linear_activation;
    bool visibility_masking_loss;
    float primary_visibility_threshold;
    float secondary_visibility_threshold;
    float depth_epsilon;
    float fixed_depth_epsilon;
    bool interpolate_depth;
    float depth_interpolation_eps;
} nbvh_pt_output_constants;

/* NeuralPathModuleConstants, include/cuda/neural/path/path_module.h:217-254 */
typedef struct __attribute__((aligned(8))) {
    uint8_t output_mode; /* 0 NeuralPathTracing */
    float normal_epsilon_offset;
    bool apply_russian_roulette;
    int32_t rr_start_depth;
    float neural_roughness;
    float neural_specular;
    float neural_metallic;
} nbvh_path_constants;

/* CudaSceneConstants, include/cuda/cuda_backend.h:58-72 */
typedef struct __attribute__((aligned(8))) {
    float constant_roughness;
    float constant_specular;
    float constant_metallic;
    float envmap_strength;
    float envmap_rotation;
    bool envmap_eval;
    bool emission_eval;
    bool skylike_background;
    bool sample_light;
    float ema_weight;
    bool ema_accumulate;
    bool force_constant_material;
} nbvh_scene_constants;

/* CameraParams (origin, du, dv, dir_top_left as 4 packed vec3), src/cuda/cuda_backend.cpp:595-611 */
typedef struct {
    float origin[3];
    float du[3];
    float dv[3];
    float dir_top_left[3];
} nbvh_camera_params;

/* Bbox3f neural_scene_bbox, include/base/bbox.h:8-11 */
typedef struct {
    float min[3];
    float max[3];
} nbvh_bbox3f;

/* HashGridOptions / NetworkOptions / OptimizerOptions, include/cuda/neural/utils/tcnn_config.h:69-77,153-166,228-233 */
typedef struct {
    int32_t n_levels;
    int32_t n_features_per_level;
    int32_t log2_hashmap_size;
    int32_t base_resolution;
    float per_level_scale;
    int32_t n_neurons;
    int32_t n_hidden_layers;
} nbvh_network_desc;

typedef struct {
    float learning_rate;
    int32_t decay_start;
    int32_t decay_interval;
    float decay_base;
    float beta1, beta2, epsilon, l2_reg;
} nbvh_optimizer_desc;

typedef enum {
    NBVH_CONST_RENDERER   = 0, /* nbvh_renderer_constants   (symbol bvh_const)         */
    NBVH_CONST_MODULE     = 1, /* nbvh_module_constants     (symbol bvh_module_const)  */
    NBVH_CONST_PT_OUTPUT  = 2, /* nbvh_pt_output_constants  (symbol pt_target_const)   */
    NBVH_CONST_PATH       = 3, /* nbvh_path_constants       (symbol path_module_const) */
    NBVH_CONST_SCENE      = 4, /* nbvh_scene_constants      (symbol scene_constants)   */
    NBVH_CONST_CAMERA     = 5, /* nbvh_camera_params        (symbol camera_params)     */
    NBVH_CONST_SCENE_BBOX = 6, /* nbvh_bbox3f               (symbol neural_scene_bbox) */
    NBVH_CONST_NETWORK    = 7, /* nbvh_network_desc  (rebuilds the parameter layout)   */
    NBVH_CONST_OPTIMIZER  = 8  /* nbvh_optimizer_desc                                  */
} nbvh_constant_block;

/* ---- scene structs (layout == reference; device pointers inside) ------------------------------- */

/* 32-byte node: BVH2Node / NeuralBVH2Node (include/base/accel/bvh.h:66-71) and TLASNode (tlas.cuh:64-68) */
typedef struct {
    float aabb_min[3];
    uint32_t prim_count_or_left; /* BVH2: prim_count (N-BVH: LoD tag); TLAS: left_or_blas_offset */
    float aabb_max[3];
    uint32_t left_or_prim_offset; /* BVH2: left child / first prim; TLAS: leaf_type (0 interior, 1 leaf, 2 neural) */
} nbvh_node32;

/* CudaBVH2 : BaseCudaBLAS (include/cuda/accel/bvh.cuh:11-23, bvh2.cuh:9-10) */
typedef struct {
    uint32_t num_nodes;
    const uint64_t *prim_indices; /* size_t* in the reference */
    uint32_t num_prim_indices;
    const float *vertices;   /* vec3 */
    const uint32_t *indices; /* uvec3 */
    const float *normals;    /* vec3 or NULL */
    const float *texcoords;  /* vec2 or NULL */
    const int32_t *material_ids;
    const uint32_t *material_map;
    uint32_t num_materials;
    const nbvh_node32 *nodes;
} nbvh_cuda_blas;

/* CudaSceneData { CudaTLAS, PackedMaterial*, cudaTextureObject_t* } (include/cuda/cuda_backend.h:109-113,
 * include/cuda/accel/tlas.cuh:124-131) */
typedef struct {
    const nbvh_node32 *tlas_nodes;
    uint32_t num_tlas_nodes;
    const uint32_t *blas_indices;
    const nbvh_cuda_blas *mesh_bvhs; /* DEVICE array of num_meshes descriptors */
    uint32_t num_meshes;
    const void *materials; /* PackedMaterial* (96-byte records; read by the sample generator and the shading kernel) */
    const void *textures;  /* cudaTextureObject_t* */
} nbvh_scene_data;

/* Primary-hit record written per pixel by nbvh_render when a hit buffer is attached:
 * the content of NeuralTraversalResult / HybridIntersectionResult for the first bounce
 * (include/cuda/neural/path/path_module.h:19-45). */
typedef struct {
    float t;                /* FLT_MAX = miss */
    uint32_t valid_hit;     /* 0 none, 1 NEURAL_TRAVERSAL_HIT, 2 NORMAL_TRAVERSAL_HIT */
    float normal_or_uv[3];  /* neural: unit normal; normal: (u, v, 0) */
    float albedo_or_ids[3]; /* neural: albedo; normal: (mesh_idx, prim_idx, 0) bit-cast */
} nbvh_hit_record;

/* NeuralTreeCutData (include/cuda/neural/neural_bvh_renderer.h:36-47): per tree-cut node loss statistics, owned by
 * the renderer, read by the sample generator's node-survival roulette */
typedef struct {
    float loss;
    uint32_t iter_alive;
    uint32_t depth;
    uint32_t samples;
} nbvh_tree_cut_data;

/* Counter slots of nbvh_get_counters */
enum {
    NBVH_CNT_RAYS = 0,          /* rays generated                            */
    NBVH_CNT_QUERIES,           /* network queries (ray x leaf)              */
    NBVH_CNT_NBVH_STEPS,        /* N-BVH interior steps (child pairs tested) */
    NBVH_CNT_BLAS_STEPS,        /* classical BLAS interior steps             */
    NBVH_CNT_TRI_TESTS,         /* triangle tests                            */
    NBVH_CNT_TILE_ITERATIONS,   /* 128-row MLP tiles executed                */
    NBVH_CNT_QUEUED_RAYS,       /* rays handed from k_primary to k_neural    */
    NBVH_CNT_NBVH_STEPS_NEURAL, /* the part of NBVH_STEPS taken in k_neural  */
    /* timeline of the LAST k_neural launches since the reset (%globaltimer, ns): bitwise-inverted minimum start time, inverted
     * time of the first queue exhaustion, maximum end time, sum of the CTA lifetimes */
    NBVH_CNT_T_START_INV,
    NBVH_CNT_T_EXHAUST_INV,
    NBVH_CNT_T_END,
    NBVH_CNT_T_CTA_SUM,
    /* development builds (-DNBVH_PHASE_TIMING=1) only, zero otherwise: SM cycles k_neural's tiles spent in the walk / tile barrier /
     * encode / MLP phases while the queue had entries, the same four during the drain, and the drain's tile iterations */
    NBVH_CNT_PH_WALK, NBVH_CNT_PH_SYNC, NBVH_CNT_PH_ENCODE, NBVH_CNT_PH_MLP,
    NBVH_CNT_PH_WALK_D, NBVH_CNT_PH_SYNC_D, NBVH_CNT_PH_ENCODE_D, NBVH_CNT_PH_MLP_D,
    NBVH_CNT_ITERS_DRAIN,
    NBVH_CNT_COUNT
};

/* ---- life cycle ------------------------------------------------------------------------------ */
uint32_t nbvh_abi_version(void);
/* Creates a context on CUDA device `device`.  config_json may be NULL (reference defaults). */
int nbvh_create(int device, const char *config_json, nbvh_ctx **out);
void nbvh_destroy(nbvh_ctx *ctx);
const char *nbvh_last_error(const nbvh_ctx *ctx);

/* ---- configuration: replaces load_config_from_json + setup_constants --------------------------- */
/* (neural_bvh_renderer_init.cu:566-749,93-103).  Partial configs are legal; missing keys take the
 * reference defaults; unknown enum strings are an error (the reference silently stores -1). */
int nbvh_load_config_json(nbvh_ctx *ctx, const char *json_text);
int nbvh_set_constants(nbvh_ctx *ctx, int which, const void *pod_host, size_t bytes);
int nbvh_get_constants(const nbvh_ctx *ctx, int which, void *pod_host, size_t bytes);
/* camera.cpp:39-82 + cuda_backend.cpp:595-611: 4x4 column-major camera matrix + yfov (radians) -> camera_params */
int nbvh_set_camera(nbvh_ctx *ctx, const float matrix16_host[16], float yfov_rad);
int nbvh_get_max_depth(const nbvh_ctx *ctx);
int nbvh_set_max_depth(nbvh_ctx *ctx, int max_depth);

/* ---- scene, tree-cut, weights ------------------------------------------------------------------ */
int nbvh_set_scene(nbvh_ctx *ctx, const nbvh_scene_data *scene_host);
/* replaces build_and_upload_neural_bvh's upload / load_neural_bvh (neural_bvh_renderer_accel.cpp:208-254).
 * set_neural_bvh copies n 32-byte nodes from DEVICE memory; the root box becomes neural_scene_bbox
 * unless NBVH_CONST_SCENE_BBOX is set afterwards. */
int nbvh_set_neural_bvh(nbvh_ctx *ctx, const nbvh_node32 *nodes, uint32_t n);
int nbvh_load_neural_bvh_file(nbvh_ctx *ctx, const char *path_host);
int nbvh_save_neural_bvh_file(nbvh_ctx *ctx, const char *path_host);
uint32_t nbvh_neural_bvh_num_nodes(const nbvh_ctx *ctx);
int nbvh_get_neural_bvh_host(const nbvh_ctx *ctx, nbvh_node32 *nodes_host, uint32_t n);
/* replaces load_network_model / save_network_model (neural_bvh_renderer_init.cu:385-404) */
int nbvh_load_network_file(nbvh_ctx *ctx, const char *path_host);
int nbvh_save_network_file(nbvh_ctx *ctx, const char *path_host);
uint64_t nbvh_n_params(const nbvh_ctx *ctx); /* == network->n_params() */
int nbvh_load_params_host(nbvh_ctx *ctx, const uint16_t *fp16_host, uint64_t n);
int nbvh_get_params_host(const nbvh_ctx *ctx, uint16_t *fp16_host, uint64_t n);
/* tcnn initialisation ranges (grid U(-1e-4,1e-4), MLP Xavier uniform), own counter-based RNG */
int nbvh_init_params(nbvh_ctx *ctx, uint64_t seed);

/* ---- inference --------------------------------------------------------------------------------- */
/* reinitialise_inference_data (neural_bvh_module.cu:12-116) + GL surfaces replaced by linear memory:
 * accum = float4[w*h], ldr = uchar4[w*h] (either may be NULL -> owned internally). */
int nbvh_set_framebuffer(nbvh_ctx *ctx, float *accum_rgba32f, uint8_t *ldr_rgba8, uint32_t w, uint32_t h);
int nbvh_set_hit_buffer(nbvh_ctx *ctx, nbvh_hit_record *hits /* w*h, or NULL to detach */);
/* == network->inference(stream, in 9xN col-major f32, out 8xN col-major f32) (path_module.cu:92) */
int nbvh_query(nbvh_ctx *ctx, const float *in, float *out, uint32_t n, void *stream);
/* parity hook: unconditional ordered leaf-interval list of each ray (NeuralBVH2::intersect driven with
 * maxt = FLT_MAX, neural_bvh2.cuh:170-292).  o,d: n x 3; outputs n x max_intervals (+ counts[n]). */
int nbvh_trace_intervals(nbvh_ctx *ctx, const float *o, const float *d, uint32_t n, uint32_t max_intervals,
                         uint32_t *counts, uint32_t *node_idx, float *mint, float *maxt, void *stream);
/* parity hook: CudaBVH2::intersect<false> on BLAS `blas_idx` (bvh2.cuh:13-88); maxt may be NULL */
int nbvh_blas_intersect(nbvh_ctx *ctx, uint32_t blas_idx, const float *o, const float *d, const float *maxt,
                        uint32_t n, float *t, float *u, float *v, uint32_t *prim, void *stream);
/* parity hook: camera_raygen for pixels [first, first+n) (neural_bvh_module_common.cuh:440-446) */
int nbvh_raygen(nbvh_ctx *ctx, uint32_t first_pixel, uint32_t n, int sample_idx, int sample_offset, float *o,
                float *d, void *stream);
/* parity hook: encode_bvh_network_input_data (neural_bvh_module_kernels.cu:266-303) for n segments
 * (o,d: n x 3; mint,maxt: n) -> out: n x input_network_dims (the column-major 9xN matrix of the reference) */
int nbvh_encode_segments(nbvh_ctx *ctx, const float *o, const float *d, const float *mint, const float *maxt,
                         uint32_t n, float *out, void *stream);
/* == neural_render(accumulated_spp, sample_offset, neural_bvh) (path_module.cu:139-245): one sample per
 * pixel of the attached framebuffer.  pixel range [first, first+count) restricts the launch to a
 * shard of the image (count = 0 -> whole frame). */
int nbvh_render(nbvh_ctx *ctx, uint32_t accumulated_spp, uint32_t sample_offset, uint32_t first_pixel,
                uint32_t pixel_count, void *stream);
/* Closest-hit neural ray query of an arbitrary ray batch: what one wavefront iteration of the reference computes for the rays
 * of one depth (tlas_traversal_and_fill_next_hybrid_traversal + non_neural_bvh_intersection + neural_bvh_traversal,
 * path_module.cu:54-137, neural_bvh_module_common.cuh:291-420): TLAS walk, classical BLAS closest hits, N-BVH leaf intervals,
 * hash grid + MLP per interval, closest-hit bookkeeping.  o, d: n x 3 (d normalised); ray_depth selects the visibility
 * threshold and LoD (1 = camera rays); hits: n records, same content as the primary-hit records of nbvh_render. */
int nbvh_trace_rays(nbvh_ctx *ctx, const float *o, const float *d, uint32_t n, uint32_t ray_depth, nbvh_hit_record *hits,
                    void *stream);
/* Multi-GPU tile partition of neural_render: shard `shard_index` of `shard_count` renders every
 * shard_count-th 32-pixel block (8x4 pixels) of the frame -- an interleaved tile partition that balances the per-pixel cost skew.
 * The RNG is keyed by the global pixel id, so the union of the shards is bit-identical to the full frame. */
int nbvh_render_interleaved(nbvh_ctx *ctx, uint32_t accumulated_spp, uint32_t sample_offset, uint32_t shard_index,
                            uint32_t shard_count, void *stream);
/* == compute_and_display_render_error + its mean (src/cuda/neural/neural_bvh_renderer_kernels.cu:91-160, neural_bvh_renderer.cu:346-358):
 * error of the accumulated image (accum / accumulated_spp) against a device RGBA32F reference image of the framebuffer's size.
 * mode = PostprocessMode (neural_bvh_renderer.h:83): 3 AbsError, 4 RelAbsError |d|/(ref+0.1), 5 MSEError, 6 RelMSEError d^2/(ref+0.1).
 * write_error_image != 0 also writes the sRGB8 error image into the 8-bit framebuffer, as the reference's validation mode does. */
int nbvh_image_error(nbvh_ctx *ctx, const float *reference_rgba32f, uint32_t accumulated_spp, int mode, int write_error_image,
                     float *mean_error_host, void *stream);
/* Device-side BLAS build: an LBVH (27-bit Morton codes, Karras hierarchy, bottom-up fit) over n_tris triangles, written in the
 * reference's BVH2Node layout (include/base/accel/bvh.h:66-71; children adjacent, root 0, one triangle per leaf) into caller-owned
 * device arrays: nodes_out[2 * n_tris - 1], prim_indices_out[n_tris] (size_t, as CudaBVH2::prim_indices).  Stands in for the host
 * ext/bvh build of CudaBackend (src/cuda/cuda_backend.cpp:186-385) when the ext/bvh tree itself is not needed (shipped .nbvh files
 * are cuts of that tree).  max_depth_host receives the tree depth; trees deeper than the 32-entry traversal stack are refused. */
int nbvh_build_blas(nbvh_ctx *ctx, const float *vertices, const uint32_t *indices, uint32_t n_tris, nbvh_node32 *nodes_out,
                    uint64_t *prim_indices_out, uint32_t *max_depth_host, void *stream);
/* n_spp samples per pixel (first_spp .. first_spp + n_spp - 1) of an interleaved shard in ONE launch sequence (max_depth > 1): the
 * path lists hold pixels x n_spp slots, which keeps the persistent kernels full on small shards (8-GPU tile partitions) and small
 * frames.  Samples of a pixel accumulate with 16-byte reductions, so the fp32 sum order differs from calling nbvh_render once per
 * sample; the 8-bit frame is resolved at the end.  first_spp == 0 clears the shard's accumulation first. */
int nbvh_render_samples(nbvh_ctx *ctx, uint32_t first_spp, uint32_t n_spp, uint32_t sample_offset, uint32_t shard_index,
                        uint32_t shard_count, void *stream);
/* End-to-end variant with HOST buffers: uploads `camera_host` (may be NULL), renders, copies the
 * 8-bit framebuffer (and hit records if hits_host != NULL) back.  Synchronous. */
int nbvh_render_host(nbvh_ctx *ctx, uint32_t accumulated_spp, uint32_t sample_offset,
                     const nbvh_camera_params *camera_host, uint8_t *ldr_rgba8_host, nbvh_hit_record *hits_host);
/* Frames in flight: nbvh_render_host split into submit and wait (one frame buffer, neural ray queue and render stream per slot, one
 * copy stream), so that the device-to-host copy of one frame and the drain of its k_neural overlap the kernels of the next frames.
 * ldr_rgba8_host must be pinned; slot < NBVH_MAX_FRAMES_IN_FLIGHT; wait for a slot before reusing it (and before reading its host
 * buffer). */
#define NBVH_MAX_FRAMES_IN_FLIGHT 4
int nbvh_render_host_submit(nbvh_ctx *ctx, uint32_t accumulated_spp, uint32_t sample_offset, const nbvh_camera_params *camera_host,
                            uint8_t *ldr_rgba8_host, uint32_t slot);
int nbvh_render_host_wait(nbvh_ctx *ctx, uint32_t slot);
/* device self-test of the hoisted IEEE division used by the slab tests (csrc/device_core.cuh): compares n_trials random
 * quotients inside the documented operand range against div.rn and returns the number of differing bit patterns */
int nbvh_selftest_division(nbvh_ctx *ctx, uint64_t seed, uint64_t n_trials, uint64_t *mismatches_host);
int nbvh_get_counters(nbvh_ctx *ctx, uint64_t *out_host, uint32_t n);
int nbvh_reset_counters(nbvh_ctx *ctx);
/* average duration (ms) and launch count of the dominant kernel since the last reset, measured with
 * CUDA events on the launching stream */
int nbvh_get_kernel_time(nbvh_ctx *ctx, float *total_ms, uint32_t *launches);
int nbvh_enable_kernel_timing(nbvh_ctx *ctx, int enable);
/* per-kernel timing of the frame pipeline without host synchronisation: after enabling, each of the next `max_frames`
 * nbvh_render calls records CUDA events before k_primary, between k_primary and k_neural, and after k_neural on the
 * launching stream; nbvh_get_stage_times synchronises on them, returns the two summed durations and re-arms. */
int nbvh_enable_stage_timing(nbvh_ctx *ctx, uint32_t max_frames);
int nbvh_get_stage_times(nbvh_ctx *ctx, float *primary_ms_total, float *neural_ms_total, uint32_t *frames);
/* Drain order of the neural ray queue (scheduling only: frames, hit records and counters do not depend on it).  The reference's
 * wavefront compaction keeps arrival order (path_module.cu:109-137); here the persistent tile kernel pulls the queued camera rays
 * in the order of their EXPECTED query-chain length, longest first: the expectation is the longest chain the camera rays of the
 * pixel's 3x3 neighbourhood needed in the last frame (a one-byte-per-pixel history the kernel maintains), and rays are filed
 * under five classes with the limits t0 >= t1 >= t2 >= t3 (class = number of limits the expectation stays below; default
 * 4, 3, 2, 2).  All limits 0: arrival order.  flags: bit 0 = 3x3 neighbourhood maximum (default), bit 1 = decaying maximum,
 * bit 2 = maximum over the last two frames. */
int nbvh_set_queue_order(nbvh_ctx *ctx, uint32_t t0, uint32_t t1, uint32_t t2, uint32_t t3, uint32_t flags);
/* the same for the optimisation step: the next `max_steps` nbvh_training_forward_backward + nbvh_optimizer_step pairs record events
 * between their kernels; nbvh_get_training_stage_times returns the summed milliseconds of
 * {k_encode_features, k_train_mlp, k_grid_backward (both launches), wait for the gradient exchange, k_adam} and re-arms. */
int nbvh_enable_training_stage_timing(nbvh_ctx *ctx, uint32_t max_steps);
int nbvh_get_training_stage_times(nbvh_ctx *ctx, float ms_total_out[5], uint32_t *steps);

/* ---- online optimisation ------------------------------------------------------------------------ */
/* == trainer->training_step(stream, in 9xB, target 8xB) -> ctx.L (16xB) + trainer->loss()
 * (ext/tiny-cuda-nn trainer.h:163-190; loss: include/cuda/neural/utils/pt_composite_loss.cuh:19-123).
 * per_sample_loss (16 x B floats, may be NULL); loss_host receives the mean loss if non-NULL. */
int nbvh_training_step(nbvh_ctx *ctx, const float *in, const float *target, uint32_t batch, float *per_sample_loss,
                       float *loss_host, void *stream);
/* Split form for data-parallel training: backward leaves fp32 gradients in the context's gradient
 * buffer (nbvh_gradients() -> device pointer, n_params floats) for an external allreduce (NCCL over
 * NVLink, driven by the host process), then optimizer_step applies Adam with the given gradient scale. */
int nbvh_training_forward_backward(nbvh_ctx *ctx, const float *in, const float *target, uint32_t batch,
                                   uint32_t global_batch, float *per_sample_loss, float *loss_sum_dev, void *stream);
float *nbvh_gradients(nbvh_ctx *ctx);

/* ---- data-parallel optimisation inside the library (NCCL over NVLink / NVSwitch; SURVEY.md 8b "nbvh_allreduce_setup(ctx, ncclComm_t)") --
 * The reference is single-GPU (tcnn::Trainer steps on its own batch, ext/tiny-cuda-nn trainer.h:163-190).  With a communicator attached,
 * nbvh_training_forward_backward -- and nbvh_training_step -- sum the gradients of all ranks before returning control to `stream`: the MLP
 * matrices in one allreduce that overlaps the hash-grid backward kernels, the grid table and the loss scalar in a second one, both on the
 * library's communication stream.  Every rank passes its own share of the batch and the GLOBAL batch size; the optimiser step is
 * replicated and leaves bit-identical parameters on all ranks.  NCCL is bound with dlopen("libnccl.so.2") (NBVH_NCCL_LIB overrides the
 * name): single-GPU hosts need no NCCL at all.
 *   nbvh_allreduce_setup            attaches the CALLER's communicator (an ncclComm_t, passed as void *; stays the caller's)
 *   nbvh_nccl_unique_id             == ncclGetUniqueId: 128 bytes that rank 0 hands to the other ranks (any channel: MPI, a file, a socket)
 *   nbvh_allreduce_setup_unique_id  == ncclCommInitRank on the context's device; the library owns the communicator
 *   nbvh_allreduce_buffer           sums a device array over the ranks in place on `stream` (tree-cut statistics, loss sums, image tiles)
 *   nbvh_nccl_comm                  the attached communicator, to share it with another context of the same process
 *   nbvh_data_parallel_world        -> number of ranks (1 without a communicator), optionally this context's rank */
#define NBVH_NCCL_UNIQUE_ID_BYTES 128
enum { NBVH_DTYPE_F32 = 0, NBVH_DTYPE_U32 = 1, NBVH_DTYPE_I32 = 2, NBVH_DTYPE_U64 = 3, NBVH_DTYPE_F64 = 4 };
int nbvh_nccl_unique_id(void *id_out_128_bytes);
int nbvh_allreduce_setup(nbvh_ctx *ctx, void *nccl_comm);
int nbvh_allreduce_setup_unique_id(nbvh_ctx *ctx, const void *unique_id_128_bytes, int rank, int world);
int nbvh_allreduce_teardown(nbvh_ctx *ctx);
int nbvh_allreduce_buffer(nbvh_ctx *ctx, void *device_ptr, uint64_t count, int dtype, void *stream);
int nbvh_data_parallel_world(const nbvh_ctx *ctx, int *rank_out);
/* the attached communicator (ncclComm_t as void *, NULL without one): lets several contexts of one process share it through
 * nbvh_allreduce_setup (the context that created it must outlive the others) */
void *nbvh_nccl_comm(const nbvh_ctx *ctx);
/* Lets the host own the fp32 gradient vector (n_params floats of device memory; NULL -> back to the internal buffer), so
 * the process that drives NCCL can allreduce it in place without a copy. */
/* == NeuralBVHModule::optimisation_data_generation -> generate_and_fill_bvh_output_reference (path_module.cu:20-52,
 * kernel granular_sample_bvh_network_input_data, neural_bvh_module_kernels.cu:119-264): fills `inputs`
 * (batch x input_network_dims) and `targets` (batch x output_network_dims) and the sampled node index of every slot.
 * The scene set with nbvh_set_scene must carry the neural mesh's classical BLAS (mesh bvh_const.neural_bvh_idx) with its
 * shading attributes, plus materials / textures when targets need them.  All pointers are device pointers;
 * batch_size must be a power of two >= 2^log2_bvh_max_traversals. */
int nbvh_optimisation_data_generation(nbvh_ctx *ctx, uint32_t training_step, uint32_t batch_size,
                                      uint32_t log2_bvh_max_traversals, uint32_t *sampled_node_indices,
                                      float current_max_loss, const nbvh_tree_cut_data *tree_cut_data, float *inputs,
                                      float *targets, void *stream);
/* Walk-ahead of the sample generator.  The rejection walk of a batch slot (random ray until it meets a tree-cut leaf: 85 % of the
 * generator's work) depends on the tree-cut, the sampler constants and the RNG only -- not on the network, nor on the per-node loss
 * statistics, which enter the node-survival roulette of the target pass.  After nbvh_sampler_walk_ahead(ctx, K, stride), a call of
 * nbvh_optimisation_data_generation for training step s that finds no walked items walks the slots of the K steps s, s + stride, ...,
 * s + (K-1) stride in ONE launch (a 2^18-slot batch alone gives every lane of a B200 1.5 slots -- no steady state; four steps' worth
 * give it 6), and the following calls for those steps only run their target pass.  Outputs are bit-identical to K = 1 (the default);
 * items are discarded when the tree-cut or a constant block changes.  stride = number of data-parallel ranks when training_step is
 * step * world + rank. */
int nbvh_sampler_walk_ahead(nbvh_ctx *ctx, uint32_t steps, uint32_t step_stride);
/* harness helper mirroring CudaBackend::create_cuda_textures_buffer (cuda_backend.cpp:387-441): RGBA8 array texture,
 * wrap addressing, linear filtering, normalized float reads.  Returns a cudaTextureObject_t in *tex_out. */
int nbvh_create_texture_rgba8(nbvh_ctx *ctx, const uint8_t *pixels_host, uint32_t w, uint32_t h, uint64_t *tex_out);
int nbvh_destroy_textures(nbvh_ctx *ctx);
/* Environment map == `envmap_texture` (include/cuda/cuda_backend_constants.cuh, CudaBackend::load_envmap
 * src/cuda/cuda_backend.cpp:455-500): an RGBA32F lat-long cudaTextureObject_t (linear filter, normalized coordinates,
 * clamp).  nbvh_set_envmap takes the caller's object (0 = the reference's white fallback image);
 * nbvh_create_envmap_rgba32f builds one from host pixels (w*h*4 floats) for hosts without their own. */
int nbvh_create_envmap_rgba32f(nbvh_ctx *ctx, const float *pixels_host, uint32_t w, uint32_t h, uint64_t *tex_out);
int nbvh_set_envmap(nbvh_ctx *ctx, uint64_t envmap_texture);
int nbvh_set_gradient_buffer(nbvh_ctx *ctx, float *grads);

/* ---- tree-cut learning (tree_cut_learning: true in the shipped configs) ----------------------------------------------
 * The reference keeps this state in NeuralBVHRenderer (m_neural_bvh_nodes, m_neural_tree_cut_data, m_neural_node_map,
 * include/cuda/neural/neural_bvh_renderer.h); here the context owns it and every step stays on the device.
 * The scene given to nbvh_set_scene must carry the classical BLAS of the neural mesh (bvh_const.neural_bvh_idx). */
/* == build_and_upload_neural_bvh(neural_bvh_min_depth) (neural_bvh_renderer_accel.cpp:24-74,152-206): the initial cut */
int nbvh_tree_cut_init(nbvh_ctx *ctx, uint32_t min_depth);
/* == reinitialise_neural_bvh_learning_states (neural_bvh_renderer_accel.cpp:76-150) for the tree-cut already loaded */
int nbvh_tree_cut_attach(nbvh_ctx *ctx);
/* == update_neural_tree_cut_loss + (k+1) x update_neural_tree_cut_decisions + split_neural_bvh_at (neural_bvh_renderer.cu:199-268,
 * neural_bvh_renderer_kernels.cu:202-287): per_sample_loss = ctx->L of the training step (stride x batch, first `dims` rows used).
 * Writes m_max_node_loss_host and the number of leaves split. num_splits = BvhSplitScheduler::get_bvh_splits(step) (0: statistics only). */
int nbvh_tree_cut_step(nbvh_ctx *ctx, const float *per_sample_loss, uint32_t stride, uint32_t dims, const uint32_t *sampled_node_indices,
                       uint32_t batch, uint32_t num_splits, float *max_node_loss_host, uint32_t *n_split_host, void *stream);
/* Data-parallel form of nbvh_tree_cut_step (one process per GPU): nbvh_tree_cut_accumulate scatters the losses of this rank's batch
 * shard into two caller-owned node-indexed delta arrays (nbvh_neural_bvh_num_nodes entries each, zeroed by the call); the host
 * allreduces both (ncclAllReduce sum); nbvh_tree_cut_decide applies the sums and runs the deterministic decision + split, so every
 * rank ends the step with the same tree-cut. */
int nbvh_tree_cut_accumulate(nbvh_ctx *ctx, const float *per_sample_loss, uint32_t stride, uint32_t dims, const uint32_t *sampled_node_indices,
                             uint32_t batch, float *loss_delta, uint32_t *samples_delta, void *stream);
int nbvh_tree_cut_decide(nbvh_ctx *ctx, const float *loss_delta, const uint32_t *samples_delta, uint32_t num_splits,
                         float *max_node_loss_host, uint32_t *n_split_host, void *stream);
/* == assign_lod_to_current_neural_leaf_nodes (neural_bvh_renderer_accel.cpp:394-441) */
int nbvh_tree_cut_assign_lod(nbvh_ctx *ctx, int *all_lods_assigned);
/* device pointers for nbvh_optimisation_data_generation (tree_cut_data) and inspection; host copies for tests / save */
const nbvh_tree_cut_data *nbvh_tree_cut_data_device(const nbvh_ctx *ctx);
const nbvh_node32 *nbvh_neural_bvh_device(const nbvh_ctx *ctx);
int nbvh_get_tree_cut_host(const nbvh_ctx *ctx, nbvh_tree_cut_data *cut_host, uint32_t *node_map_host, uint32_t n);
int nbvh_set_tree_cut_host(nbvh_ctx *ctx, const nbvh_tree_cut_data *cut_host, uint32_t n);
uint32_t nbvh_tree_cut_max_depth(const nbvh_ctx *ctx);
int nbvh_optimizer_step(nbvh_ctx *ctx, void *stream);
uint32_t nbvh_training_step_count(const nbvh_ctx *ctx);

/* Device-side TLAS build (f3) == CudaBackend::build_tlas (src/cuda/cuda_backend.cpp:312-383): a sweep-SAH binary tree over the root
 * boxes of the per-mesh BLAS in the TLASNode layout (include/cuda/accel/tlas.cuh:64-68), one mesh per leaf.  mesh_bvhs_device is the
 * DEVICE descriptor array later passed as nbvh_scene_data.mesh_bvhs; neural_meshes_host lists the meshes the N-BVH stands for.
 * Outputs are caller-owned device arrays: 2 * num_meshes - 1 nodes, num_meshes blas_indices. */
int nbvh_build_tlas(nbvh_ctx *ctx, const nbvh_cuda_blas *mesh_bvhs_device, uint32_t num_meshes, const uint32_t *neural_meshes_host,
                    uint32_t n_neural, nbvh_node32 *tlas_nodes_out, uint32_t *blas_indices_out, void *stream);

/* ---- f4: result writers and configuration round trip ------------------------------------------------------------ */
/* == NeuralBVHRenderer::compute_and_display_flip_error (src/cuda/neural/neural_bvh_renderer.cu:34-96; replaces ext/flip-cuda): mean
 * LDR-FLIP of accum / accumulated_spp against a device RGBA32F reference image of the framebuffer's size.  ppd <= 0: the reference's
 * viewing conditions (67.02 pixels per degree).  error_map_out: optional device array of W x H floats (per-pixel FLIP error).
 * Like the reference, the images are NOT clamped to [0, 1] first: LDR-FLIP is defined on that range, and HDR highlights can drive the
 * feature exponent past 1 (per-pixel errors above 1 or infinite, in flip-cuda as well); clamp or tone-map beforehand for such images. */
int nbvh_flip_error(nbvh_ctx *ctx, const float *reference_rgba32f, uint32_t accumulated_spp, float ppd, float *error_map_out,
                    float *mean_flip_host, void *stream);
/* the same for two arbitrary device RGBA32F images (linear RGB, alpha ignored); test is divided by test_divisor first */
int nbvh_flip_images(nbvh_ctx *ctx, const float *reference_rgba32f, const float *test_rgba32f, float test_divisor, uint32_t width,
                     uint32_t height, float ppd, float *error_map_out, float *mean_flip_host, void *stream);
/* == write_exr (include/base/bitmap.h:97-182) as called by RenderApp::save_framebuffer (src/app/render_app.cpp:132-165): scanline
 * OpenEXR, channels A,B,G,R as 32-bit float, no compression, every value divided by rgb_renormalisation.  Host memory only; the file
 * is byte-identical to the one the reference's tinyexr writes. */
int nbvh_write_exr_host(const char *path, uint32_t width, uint32_t height, const float *rgba_host, float rgb_renormalisation);
/* == NeuralBVHRenderer::save_image (src/cuda/neural/neural_bvh_renderer_init.cu:361-383): accumulation buffer / accumulated_spp -> EXR;
 * NBVH_ERR_STATE while scene_constants.ema_accumulate is on (the reference refuses as well). */
int nbvh_save_image_exr(nbvh_ctx *ctx, const char *path, uint32_t accumulated_spp, void *stream);
/* == FrameTimer::write_to_file (include/utils/frame_timer.h:33-58): learning (with losses) / inference (losses NULL) timing CSV */
int nbvh_write_frame_times_csv_host(const char *path, const double *frame_ms, const float *losses_or_null, uint32_t n);
/* Result file names of validation mode (neural_bvh_renderer_init.cu:328-372).  kind 0: results/<base>_<variant>_spp_<n>.exr,
 * 1: results/<base>_<variant>_learning_spp_<n>.csv, 2: results/<base>_<variant>_inference_spp_<n>.csv */
int nbvh_result_path(const char *scene_dir, const char *base_config_name, const char *variant_config_name, int kind,
                     uint32_t accumulated_spp, char *out, size_t out_capacity);
/* == the module-owned part of NeuralBVHRenderer::save_config (neural_bvh_renderer_init.cu:406-509), reference key names; loading the
 * result with nbvh_load_config_json restores every constant block bit for bit.  *needed = bytes required including the NUL. */
int nbvh_save_config_json(const nbvh_ctx *ctx, char *out, size_t capacity, size_t *needed);

#ifdef __cplusplus
}
#endif
#endif /* NBVH_B200_H */

/*
 * nbvh_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's neural ray-query hot path (WeiPhil/nbvh), used as the
 * parity checker for the CUDA kernels in nbvh_b200/csrc.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.  The product path never does.
 *
 * Parity status: the reference ships NO golden vectors or known-answer tests for this path
 * (SURVEY.md section 4 / 8c).  The oracle is pinned instead by
 *   (1) the reference's own traversal/encode headers compiled with nvcc for sm_100 and run on the GPU
 *       box through oracle/ref_tools/ref_harness.cu (bit-exact comparison of interval lists, closest
 *       hits, encoded inputs and decoded hits; fixtures committed under tests/golden/), and
 *   (2) the floating-point contraction pattern read from the PTX nvcc 12.9 emits for the reference's
 *       kernels (-arch=sm_100, default -fmad=true); every fmaf() below cites the expression it mirrors.
 * The tcnn network itself (fp16-accumulate wmma MLP) is restated by published algorithm; parity with
 * a tcnn build is tolerance-based and, until the tcnn oracle is built, "parity unpinned" for the MLP
 * values (the hash-grid encoding is restated operation-for-operation in fp16 and is expected bit-exact).
 *
 * All citations are into /root/reference.
 */
#ifndef NBVH_ORACLE_H
#define NBVH_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 32-byte node shared by BVH2Node / NeuralBVH2Node / TLASNode (include/base/accel/bvh.h:66-71,
 * include/cuda/accel/tlas.cuh:64-68).  a/b are: BVH2: prim_count / left_or_prim_offset,
 * TLAS: left_or_blas_offset / leaf_type. */
typedef struct {
    float mn[3];
    uint32_t a;
    float mx[3];
    uint32_t b;
} onode_t;

typedef struct {
    const onode_t *nodes;
    uint32_t n_nodes;
    const uint32_t *prim_ids; /* reference stores size_t; values identical */
    const float *vertices;    /* xyz */
    const uint32_t *indices;  /* 3 per triangle */
} oblas_t;

typedef struct {
    const onode_t *tlas_nodes;
    uint32_t n_tlas_nodes;
    const uint32_t *blas_indices;
    const oblas_t *blas;
    uint32_t n_blas;
} oscene_t;

/* Constants in force on the path (defaults: neural_bvh_renderer_constants.cuh:13-47,
 * neural_bvh_module.h:243-274). */
typedef struct {
    uint32_t rendered_lod;
    float fixed_node_dilation, node_extent_dilation;
    float inference_mint_epsilon, inference_maxt_epsilon;
    float learning_mint_epsilon, learning_maxt_epsilon;
    float primary_visibility_threshold, secondary_visibility_threshold;
    float depth_epsilon, fixed_depth_epsilon;
    uint32_t ignore_neural_blas;
    uint32_t n_segment_points;      /* 3 */
    uint32_t segment_distribution;  /* 0 EquidistantInterior 1 EquidistantExterior 2 Uniform (camera / path rays: points drawn from the ray's generator) */
    float bbox_min[3], bbox_max[3]; /* neural_scene_bbox */
    /* camera (cuda_backend.cpp:595-611) */
    float cam_origin[3], cam_du[3], cam_dv[3], cam_top_left[3];
    uint32_t fb_w, fb_h;
} oconst_t;

typedef struct {
    uint32_t n_levels, n_features, base_resolution, log2_hashmap_size;
    float per_level_scale;
    uint32_t n_hidden_layers; /* 4 -> 5 matrices */
    uint32_t n_neurons;       /* 64 */
    uint32_t n_input;         /* 96 */
    uint32_t n_output_padded; /* 16 */
    uint32_t n_output;        /* 8 */
    const uint16_t *params;   /* fp16 bits: MLP matrices then grid (network_with_input_encoding.h:115-122) */
    uint64_t n_params;
} onet_t;

/* --- small building blocks (exported for unit tests) --- */
uint16_t onbvh_f2h(float f);
float onbvh_h2f(uint16_t h);
uint32_t onbvh_lcg_rng_init(uint32_t index, uint32_t frame, uint32_t linear);
float onbvh_lcg_randomf(uint32_t *state);
void onbvh_camera_params(const float matrix[16], float yfov_rad, uint32_t w, uint32_t h, float origin[3], float du[3],
                         float dv[3], float top_left[3]);
void onbvh_raygen(const oconst_t *c, uint32_t pixel, int sample_idx, int rnd_offset, float o[3], float d[3],
                  uint32_t *rng_state);
int onbvh_slab(const float mn[3], const float mx[3], const float o[3], const float d[3], float ray_mint,
               float ray_maxt, int check, float *mint, float *maxt);
uint64_t onbvh_grid_n_params(const onet_t *net, uint32_t *offsets /* n_levels+1 */);

/* --- traversal --- */
/* a3: one resumable N-BVH step.  Returns 1 and (mint,maxt,node) or 0 when the walk ended. */
int onbvh_neural_intersect(const onode_t *nodes, const oconst_t *c, uint32_t lod, const float o[3],
                           const float d[3], uint32_t *stack, int *stack_size, int inference, float current_maxt,
                           float dilation, float *mint, float *maxt, uint32_t *node_idx, uint64_t *interior_steps);
/* unconditional ordered leaf-interval list per ray (maxt_ray = FLT_MAX throughout) */
void onbvh_trace_intervals(const onode_t *nodes, const oconst_t *c, uint32_t n_rays, const float *o, const float *d,
                           uint32_t max_intervals, uint32_t *counts, uint32_t *node_idx, float *mint, float *maxt,
                           uint64_t *interior_steps);
/* a14: BVH2 closest hit.  Returns t (FLT_MAX on miss). */
float onbvh_blas_intersect(const oblas_t *blas, const float o[3], const float d[3], float ray_mint, float ray_maxt,
                           float *u, float *v, uint32_t *prim_idx, uint64_t *interior_steps, uint64_t *tri_tests);
void onbvh_blas_intersect_batch(const oblas_t *blas, uint32_t n, const float *o, const float *d, const float *maxt,
                                float *t, float *u, float *v, uint32_t *prim, uint64_t *counters);
/* a4 */
int onbvh_tlas_intersect_only(const oscene_t *s, const float o[3], const float d[3], uint32_t *stack,
                              int *stack_size, uint32_t *blas_idx, int *is_neural);

/* --- network --- */
/* a8: leaf interval -> 3x3 encoded points */
void onbvh_encode_segment(const oconst_t *c, const float o[3], const float d[3], float mint, float maxt,
                          float *out /* 3*n_points */);
/* the same with the ray's generator: SegmentDistribution::Uniform draws one parameter per point from it (rng may be NULL) */
void onbvh_encode_segment_rng(const oconst_t *c, const float o[3], const float d[3], float mint, float maxt, uint32_t *rng,
                              float *out);
/* a9: hash grid, fp16 arithmetic as tcnn; writes n_points*n_levels*n_features fp16 bit patterns */
void onbvh_grid_encode(const onet_t *net, const float *points, uint32_t n_points, uint16_t *features);
/* a10: mode 0 = fp16 in / fp32 accumulate (the B200 kernel's arithmetic), 1 = fp16-accumulate emulation of
 * tcnn's wmma path (chunked by 16 along K), 2 = full fp32 on fp16-valued parameters. */
void onbvh_mlp_forward(const onet_t *net, const uint16_t *features, int mode, float *out /* n_output */,
                       float *hidden /* optional: n_hidden_layers*n_neurons post-activation */);
/* K2+K3+K4+K5 for a batch of column-major 9xN inputs -> 8xN outputs */
void onbvh_query(const onet_t *net, const float *in, uint32_t n, int mode, float *out);
/* a11 */
int onbvh_decode(const oconst_t *c, const float *y, float seg_mint, float seg_maxt, uint32_t depth, float *t_out);


/* --- online optimisation (a18-a21) --- */
/* PathTracingOutputConstants subset driving the composite loss (neural_bvh_module.h:243-274). Loss types:
 * 0 RelativeL2, 1 L2, 2 RelativeL1, 3 L1, 4 RelativeL2Luminance, 5 BinaryCrossEntropy, 6 LogL2, 7 LHalf */
typedef struct {
    uint32_t vis_dim, depth_dim, normal_dim, albedo_dim; /* 1,1,3,3 */
    uint32_t vis_loss, depth_loss, normal_loss, albedo_loss;
    float vis_scale, depth_scale, normal_scale, albedo_scale;
    uint32_t normal_linear, depth_linear, albedo_linear, vis_masking;
} oloss_t;

typedef struct {
    float learning_rate;
    int32_t decay_start, decay_interval;
    float decay_base, beta1, beta2, epsilon, l2_reg;
} oopt_t;

/* a18: composite_pt_loss (pt_composite_loss.cuh:19-123) for one output element.  Returns the loss value (already
 * divided by n_total) and writes the loss-scaled gradient (128 * g / n_total), before any fp16 rounding. */
float onbvh_loss_element(const oloss_t *lc, const float *target8, float prediction, uint32_t elem, uint32_t n_total,
                         float loss_scale, float *grad);
/* a9/a10 forward with stored activations + a18 + a19 + a20: gradients of the loss-scaled objective.
 * in: B x 9, target: B x 8 (row per sample).  L: B x 16 per-output loss values (may be NULL).
 * grads: n_params floats, OVERWRITTEN (MLP matrices then grid, same order as the parameters).
 * Rounding points follow the B200 kernels (fp16 operands, fp32 accumulation): features, hidden activations,
 * network output, dL/dy, every dZ tile and dL/dfeatures are rounded to fp16; sums are fp32/double.
 * n_total = global_batch * n_output (data-parallel shards pass the global batch). */
void onbvh_train_forward_backward(const onet_t *net, const oloss_t *lc, const float *in, const float *target,
                                  uint32_t batch, uint32_t global_batch, float loss_scale, float *L, float *grads,
                                  double *loss_sum);
/* a21: adam_step (optimizers/adam.h:48-123) under ExponentialDecay (exponential_decay.h:60-71).  `step` is the
 * number of optimiser steps already taken.  Gradients are first rounded to fp16 (tcnn's gradient type) so that the
 * "gradient == 0 -> skip" rule sees the same underflow. */
void onbvh_adam_step(const oopt_t *oc, uint32_t step, uint64_t n_params, uint64_t n_matrix, float loss_scale,
                     const float *grads, float *master, uint16_t *params, float *m, float *v, uint32_t *param_steps);
float onbvh_learning_rate(const oopt_t *oc, uint32_t step);

/* --- online-optimisation sample generator (a17, K10) --- */
/* PackedMaterial (include/base/material.h:52-87): vtable pointer + Disney parameters, 96 bytes */
typedef struct {
    uint64_t vptr;
    float base_color[3], metallic, specular, roughness, specular_tint, anisotropy, sheen, sheen_tint, clearcoat,
        clearcoat_gloss, base_emission[3], emission_scale, normal[3], ior, specular_transmission, pad;
} omaterial_t;
typedef struct {
    const uint8_t *rgba; /* w*h*4, row-major */
    uint32_t w, h;
} otexture_t;
typedef struct {
    float loss;
    uint32_t iter_alive, depth, samples;
} otreecut_t;
typedef struct {
    oblas_t blas;                 /* classical BLAS of the neural mesh */
    const float *normals;         /* vec3 per vertex or NULL */
    const float *texcoords;       /* vec2 per vertex or NULL */
    const int32_t *material_ids;  /* per material segment */
    const uint32_t *material_map; /* first primitive of each segment */
    uint32_t num_materials;
    const omaterial_t *materials;
    const otexture_t *textures;
    /* bvh_const fields the sampler reads (neural_bvh_renderer_constants.cuh:13-47) */
    float sampling_bbox_scaling;
    uint32_t current_learned_lod;
    uint32_t training_node_survival;
    float survival_probability_scale, min_rr_survival_prob;
    int32_t max_training_bounces;
    uint32_t backface_bounce_training, sample_cosine_weighted;
    uint32_t depth_linear_activation;
} osampler_t;
/* granular_sample_bvh_network_input_data (neural_bvh_module_kernels.cu:119-264) for the whole batch.  inputs: batch x 9,
 * targets: batch x 8, sampled_nodes: batch -- all read-modify-write (rejected slots keep their contents). */
void onbvh_generate_batch(const onode_t *nbvh, const oconst_t *c, const osampler_t *sp, uint32_t training_step,
                          uint32_t batch_size, uint32_t log2_bvh_max_traversals, uint32_t *sampled_nodes,
                          float current_max_loss, const otreecut_t *tree_cut, float *inputs, float *targets);
/* CUDA bilinear fetch (wrap addressing, normalized coordinates, 8-bit fractional weights) of an RGBA8 texture */
void onbvh_tex2d(const otexture_t *t, float u, float v, float out[4]);

/* --- whole primary-ray pipeline (section 9 of SURVEY.md), max_depth = 1 --- */
typedef struct {
    uint8_t valid_hit; /* 0 none, 1 neural, 2 normal */
    float t;
    float normal_or_uv[3];
    float albedo_or_ids[3]; /* neural: albedo; normal: (mesh, prim) bit-cast */
} ohit_t;
void onbvh_render_primary(const oscene_t *s, const onode_t *nbvh, const oconst_t *c, const onet_t *net,
                          uint32_t pixel_begin, uint32_t pixel_end, uint32_t pixel_stride, int sample_idx,
                          int rnd_offset, int mlp_mode, ohit_t *hits /* indexed by (pixel-begin)/stride */,
                          uint64_t *counters /* [0]=queries [1]=nbvh interior steps [2]=blas interior [3]=tri tests */,
                          int n_threads);

/* closest hit of arbitrary rays (o, d: n x 3) through TLAS, classical BLAS and the neural BLAS; lod = rendered_lod normally */
void onbvh_trace_rays(const oscene_t *s, const onode_t *nbvh, const oconst_t *c, const onet_t *net, uint32_t n, const float *o,
                      const float *d, uint32_t ray_depth, uint32_t lod, int mlp_mode, ohit_t *hits, uint64_t *counters, int n_threads);

/* --- hybrid neural path tracing, max_depth >= 1 (a15 / f1: shading, BSDF sampling, Russian roulette, re-queue) --- */
typedef struct {
    const float *normals, *texcoords;
    const int32_t *material_ids;
    const uint32_t *material_map;
    uint32_t num_materials;
} oattr_t;
typedef struct {
    const oattr_t *attr; /* one per BLAS */
    const omaterial_t *materials;
    const otexture_t *textures;
    uint32_t neural_bvh_idx, max_depth;
    uint32_t path_lod_switch, lod_switch;
    int32_t num_lods_learned;
    /* path_module_const (path_module.h:217-254) */
    float normal_epsilon_offset;
    uint32_t apply_russian_roulette, rr_start_depth;
    float neural_roughness, neural_specular, neural_metallic;
    /* scene_constants (cuda_backend.h:58-72) */
    uint32_t force_constant_material, emission_eval;
    float constant_roughness, constant_specular, constant_metallic;
    uint32_t envmap_eval, skylike_background;
    float envmap_rotation_deg, envmap_strength;
    const float *envmap; /* RGBA32F lat-long map or NULL (white) */
    uint32_t env_w, env_h;
} opath_t;
/* radiance: 3 floats per pixel (the value accumulate_pixel_value receives); counters[4] = rays traced */
void onbvh_render_path(const oscene_t *s, const onode_t *nbvh, const oconst_t *c, const onet_t *net, const opath_t *pc,
                       uint32_t pixel_begin, uint32_t pixel_end, uint32_t pixel_stride, int sample_idx, int rnd_offset,
                       int mlp_mode, float *radiance, ohit_t *primary_hits, uint64_t *counters, int n_threads);

#ifdef __cplusplus
}
#endif
#endif

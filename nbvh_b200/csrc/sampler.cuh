// sampler.cuh -- parameter block of the online-optimisation sample generator (sampler.cu), shared with context.cu
#pragma once
#include "shading.cuh"

namespace nbvh {

struct SamplerParams {
    NbvhConst nbvh; // raw (undilated) node array, learning epsilons; leaf_lod_max / dilation are set per slot
    BlasConst blas; // the neural mesh's classical BLAS (bvh_const.neural_bvh_idx)
    BlasAttr attr;
    uint32_t nodes_safe;
    const MaterialPod *materials;
    const cudaTextureObject_t *textures;
    EncodeConst enc;
    float bbox_min[3], bbox_ext[3];
    float sampling_bbox_scaling, fixed_node_dilation, node_extent_dilation;
    uint32_t current_learned_lod, num_bvh_lods;
    uint32_t training_node_survival;
    float survival_probability_scale, min_rr_survival_prob;
    int max_training_bounces;
    uint32_t backface_bounce_training, sample_cosine_weighted;
    uint32_t n_in, n_out, vis_dim, depth_dim, normal_dim, albedo_dim, depth_linear;
    int sub_batch_size, n_sub_batch;
    uint32_t rng_index; // training_step * batch_size
    uint32_t *sampled_node_indices;
    float *inputs, *targets;
    float current_max_loss;
    const nbvh_tree_cut_data *tree_cut;
    unsigned long long *counters;
    // two-pass generator (n_sub_batch == 1): items handed from k_sample_walk to k_sample_targets (48 bytes each, capacity = batch)
    void *items;
    uint32_t *item_count;
    uint32_t hard_origins, group, next_a, next_c; // hard-slot search: threshold, lanes per slot, LCG map of 3 (group - 1) draws
    uint2 *hard_slots;       // (slot, origins consumed) of slots handed to k_sample_walk_hard, capacity = batch
    uint32_t *hard_count;    // their number
    uint32_t *dilation_unsafe; // set when a tree-cut node fails k_check_dilation_safe (-> guarded hoisted division in the walk)
    float *node_keep; // survival_probability_scale * keep per tree-cut node (k_node_keep), capacity = tree-cut nodes
};

// sampler_counters: 8 device words owned by the context (cleared by every launch)
cudaError_t launch_generate_batch(const SamplerParams &p, uint32_t *sampler_counters, int num_sms, cudaStream_t stream);

} // namespace nbvh

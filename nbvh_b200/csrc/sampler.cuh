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
    float bbox_min[3], bbox_ext[3], bbox_max[3];
    uint32_t uniform_points;            // SegmentDistribution::Uniform: random segment points
    uint32_t depth_mode;                // DepthLearningMode: 0 LocalT, 1 WorldT, 2 Position, 3 LocalPosition
    uint32_t view_dependent_sampling;   // bvh_const.view_dependent_sampling / view_dependent_ratio (general kernel only)
    float view_dependent_ratio;
    float leaf_boundary_sampling_ratio; // bvh_const.leaf_boundary_sampling_ratio (general kernel only)
    CameraConst cam;                    // camera_params for view-dependent sampling
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
    // two-pass generator (n_sub_batch == 1): items handed from k_sample_walk to k_sample_targets (48 bytes per slot and step)
    void *items;
    uint32_t n_steps;          // optimisation steps walked by one k_sample_walk launch (walk-ahead), >= 1
    uint32_t rng_index_step;   // rng_index of step k = rng_index + k * rng_index_step  (step stride * batch size)
    uint32_t *dilation_unsafe; // set when a tree-cut node fails k_check_dilation_safe (-> guarded hoisted division in the walk)
    float *node_keep;          // survival_probability_scale * keep per tree-cut node (k_node_keep), capacity = tree-cut nodes
};

// n_sub_batch > 1 (rays continuing through the tree-cut): one thread per sub-batch
cudaError_t launch_generate_batch(const SamplerParams &p, int num_sms, cudaStream_t stream);
// n_sub_batch == 1: the walk of p.n_steps steps (sampler_counters: 8 device words owned by the context), then per step the targets
cudaError_t launch_sample_walk(const SamplerParams &p, uint32_t *sampler_counters, int num_sms, cudaStream_t stream);
cudaError_t launch_sample_targets(const SamplerParams &p, uint32_t step_first_item, cudaStream_t stream);

} // namespace nbvh

// treecut.cuh -- tree-cut learning kernels (treecut.cu), shared with context.cu
#pragma once
#include "kernels.cuh"

namespace nbvh {

struct TreeCutResult {
    float max_node_loss; // m_max_node_loss_host: largest loss heuristic of a splittable leaf (0 when none is positive)
    uint32_t n_split;    // leaves split in this step (the tree-cut grew by 2 * n_split nodes)
    uint32_t max_depth;  // m_neural_tree_cut_max_depth (running maximum)
    uint32_t pad;
};

size_t tree_cut_workspace_bytes(uint32_t capacity);
cudaError_t launch_tree_cut_accumulate(const float *L, uint32_t stride, uint32_t dims, const uint32_t *sampled_nodes, uint32_t batch,
                                       nbvh_tree_cut_data *cut, cudaStream_t stream);
cudaError_t launch_tree_cut_accumulate_delta(const float *L, uint32_t stride, uint32_t dims, const uint32_t *sampled_nodes, uint32_t batch,
                                             uint32_t n_nodes, float *loss_delta, uint32_t *samples_delta, cudaStream_t stream);
cudaError_t launch_tree_cut_apply_delta(nbvh_tree_cut_data *cut, uint32_t n, const float *loss_delta, const uint32_t *samples_delta, cudaStream_t stream);
cudaError_t launch_tree_cut_decide_and_split(float4 *nodes, uint32_t n, uint32_t capacity, nbvh_tree_cut_data *cut, uint32_t *node_map,
                                             const float4 *blas_nodes, uint32_t num_splits, uint32_t current_learned_lod, void *workspace,
                                             size_t workspace_bytes, TreeCutResult *result_dev, cudaStream_t stream);
cudaError_t launch_tree_cut_assign_lod(float4 *nodes, uint32_t n, uint32_t level, cudaStream_t stream);

} // namespace nbvh

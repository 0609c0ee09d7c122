// treecut.cu -- tree-cut learning of the online optimisation: replaces
//   update_neural_tree_cut_loss / update_neural_tree_cut_decisions  /root/reference/src/cuda/neural/neural_bvh_renderer_kernels.cu:202-287
//   the per-step driver                                              /root/reference/src/cuda/neural/neural_bvh_renderer.cu:160-283
//   NeuralBVHRenderer::split_neural_bvh_at / assign_lod_to_current_neural_leaf_nodes
//                                                                    /root/reference/src/cuda/neural/neural_bvh_renderer_accel.cpp:256-441
//   NeuralTreeCutData::compute_loss_heuristic                        /root/reference/include/cuda/neural/neural_bvh_renderer.h:36-47
//
// The reference scatters the per-element losses with two atomics per element, finds the k largest node losses with k + 1
// full passes over the node array (one atomicMax pass per split), copies the indices to the host, downloads the whole
// node array, splits on the CPU and uploads the changes.  Here one step is: one atomic pair per SAMPLE, one pass that ages
// the leaves and writes their loss heuristic, one device radix sort, one flag/scan/scatter split -- the node array never
// leaves the device and the only read-back is (max loss, number of splits), which the next step's sampler call needs.
//
// Results equal the reference's up to the order of fp32 additions in the loss sums (atomics, unordered in both): the k
// distinct largest positive heuristics are selected, limited to the number of leaves, and split in ascending node-index
// order (std::set order), so new node indices are assigned exactly as the reference assigns them.
#include "treecut.cuh"

#include <cub/cub.cuh>

namespace nbvh {

constexpr uint32_t kLeafFinal   = 1u;          // LEAF_FINAL
constexpr uint32_t kNoChild     = 0xFFFFFFFFu; // NO_NEURAL_CHILD
__host__ __device__ inline uint32_t leaf_lod(uint32_t lod) { return 2u + lod; } // LEAF_LOD

// update_neural_tree_cut_loss: L is stride x batch (tcnn ctx.L, 16 x B), the first `dims` rows of a column are one sample's losses
__global__ void k_tc_accumulate(const float *__restrict__ L, uint32_t stride, uint32_t dims, const uint32_t *__restrict__ sampled_nodes,
                                uint32_t batch, nbvh_tree_cut_data *cut)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch)
        return;
    float s = 0.f;
    for (uint32_t c = 0; c < dims; ++c)
        s += L[(size_t)i * stride + c];
    nbvh_tree_cut_data *nd = cut + sampled_nodes[i];
    atomicAdd(&nd->loss, s);
    atomicAdd(&nd->samples, dims); // the reference counts every loss element
}

// data-parallel variant: the batch shard's statistics go to delta arrays, which the host allreduces (NCCL) before they are applied
__global__ void k_tc_accumulate_delta(const float *__restrict__ L, uint32_t stride, uint32_t dims, const uint32_t *__restrict__ sampled_nodes,
                                      uint32_t batch, float *loss_delta, uint32_t *samples_delta)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch)
        return;
    float s = 0.f;
    for (uint32_t c = 0; c < dims; ++c)
        s += L[(size_t)i * stride + c];
    const uint32_t node = sampled_nodes[i];
    atomicAdd(loss_delta + node, s);
    atomicAdd(samples_delta + node, dims);
}

__global__ void k_tc_apply_delta(nbvh_tree_cut_data *cut, uint32_t n, const float *__restrict__ loss_delta, const uint32_t *__restrict__ samples_delta)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const uint32_t ds = samples_delta[i];
    if (ds) {
        cut[i].loss += loss_delta[i];
        cut[i].samples += ds;
    }
}

// update_neural_tree_cut_decisions, first pass: age the leaves, emit their loss heuristic as sort key
__global__ void k_tc_candidates(const float4 *__restrict__ nodes, uint32_t n, nbvh_tree_cut_data *cut, float *keys, uint32_t *vals)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const uint32_t prim_count = __float_as_uint(nodes[2u * i].w), left = __float_as_uint(nodes[2u * i + 1u].w);
    float key = -FLT_MAX;
    if (left == kNoChild) {
        nbvh_tree_cut_data nd = cut[i];
        nd.iter_alive += 1u;
        cut[i].iter_alive = nd.iter_alive;
        if (prim_count != kLeafFinal && nd.samples != 0u) {
            const float ia = (float)nd.iter_alive;
            key = nd.loss / (ia * sqrtf((float)nd.samples / ia));
        }
    }
    keys[i] = key;
    vals[i] = i;
}

// sorted descending: flag the first entry of every run of equal positive keys (the reference finds DISTINCT maxima)
__global__ void k_tc_flag_distinct(const float *__restrict__ keys, uint32_t n, uint32_t *flags)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n)
        return;
    const float k = keys[j];
    flags[j] = (k > 0.f && (j == 0u || keys[j - 1u] != k)) ? 1u : 0u;
}

// mark the nodes to split: the `limit` best distinct candidates whose BLAS node is interior
__global__ void k_tc_mark(const float *__restrict__ keys, const uint32_t *__restrict__ vals, const uint32_t *__restrict__ flags,
                          const uint32_t *__restrict__ rank, uint32_t n, uint32_t limit, const float4 *__restrict__ blas_nodes,
                          const uint32_t *__restrict__ node_map, uint32_t *split_flag, float *max_loss)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n)
        return;
    if (j == 0u)
        *max_loss = keys[0] > 0.f ? keys[0] : 0.f; // m_max_node_losses[0]
    if (flags[j] && rank[j] < limit) {
        const uint32_t node = vals[j];
        const uint32_t blas_prim_count = __float_as_uint(blas_nodes[2u * node_map[node]].w);
        if (blas_prim_count == 0u) // BVH2Node::is_leaf() == prim_count > 0: "Leaf node reached, no split"
            split_flag[node] = 1u;
    }
}

// split_neural_bvh_at: node -> interior, two new leaves appended at n + 2 * (rank among the split nodes, ascending index)
__global__ void k_tc_split(float4 *nodes, uint32_t n, const uint32_t *__restrict__ split_flag, const uint32_t *__restrict__ split_rank,
                           const float4 *__restrict__ blas_nodes, uint32_t *node_map, nbvh_tree_cut_data *cut, uint32_t current_learned_lod,
                           uint32_t *n_split_out, uint32_t *max_depth_out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    if (i == n - 1u)
        *n_split_out = split_rank[i] + split_flag[i];
    if (!split_flag[i])
        return;
    const uint32_t bi    = node_map[i];
    const float4 p0 = blas_nodes[2u * bi], p1 = blas_nodes[2u * bi + 1u];
    const uint32_t bl    = __float_as_uint(p1.w); // left child in the BLAS, right = bl + 1
    const uint32_t left  = n + 2u * split_rank[i];
    uint32_t prim_count  = __float_as_uint(nodes[2u * i].w);
    if (prim_count == leaf_lod(current_learned_lod))
        prim_count = 0u; // a leaf assigned to the level being learned becomes a plain interior node
    nodes[2u * i]      = make_float4(p0.x, p0.y, p0.z, __uint_as_float(prim_count));
    nodes[2u * i + 1u] = make_float4(p1.x, p1.y, p1.z, __uint_as_float(left));
    const uint32_t depth = cut[i].depth + 1u;
    for (uint32_t c = 0; c < 2u; ++c) {
        const float4 c0 = blas_nodes[2u * (bl + c)], c1 = blas_nodes[2u * (bl + c) + 1u];
        const bool blas_leaf = __float_as_uint(c0.w) > 0u;
        nodes[2u * (left + c)]      = make_float4(c0.x, c0.y, c0.z, __uint_as_float(blas_leaf ? kLeafFinal : leaf_lod(current_learned_lod)));
        nodes[2u * (left + c) + 1u] = make_float4(c1.x, c1.y, c1.z, __uint_as_float(kNoChild));
        node_map[left + c] = bl + c;
        nbvh_tree_cut_data nd;
        nd.loss = 0.f, nd.iter_alive = 1u, nd.depth = depth, nd.samples = 0u;
        cut[left + c] = nd;
    }
    atomicMax(max_depth_out, depth);
}

// assign_lod_to_current_neural_leaf_nodes
__global__ void k_tc_assign_lod(float4 *nodes, uint32_t n, uint32_t level)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    if (__float_as_uint(nodes[2u * i + 1u].w) == kNoChild) {
        const uint32_t pc = __float_as_uint(nodes[2u * i].w);
        nodes[2u * i].w   = __uint_as_float(pc == kLeafFinal ? kLeafFinal : leaf_lod(level));
    }
}

size_t tree_cut_workspace_bytes(uint32_t capacity)
{
    size_t sort_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, sort_bytes, (const float *)nullptr, (float *)nullptr, (const uint32_t *)nullptr,
                                              (uint32_t *)nullptr, (int)capacity);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)capacity);
    const size_t tmp = ((sort_bytes > scan_bytes ? sort_bytes : scan_bytes) + 255) & ~(size_t)255;
    // keys in/out, vals in/out, flags, rank, split_flag, split_rank: 8 arrays of `capacity` words + 4 result words
    return tmp + 8 * (size_t)capacity * 4 + 256;
}

cudaError_t launch_tree_cut_accumulate(const float *L, uint32_t stride, uint32_t dims, const uint32_t *sampled_nodes, uint32_t batch,
                                       nbvh_tree_cut_data *cut, cudaStream_t stream)
{
    if (batch == 0)
        return cudaSuccess;
    k_tc_accumulate<<<(batch + 255u) / 256u, 256, 0, stream>>>(L, stride, dims, sampled_nodes, batch, cut);
    return cudaGetLastError();
}

cudaError_t launch_tree_cut_accumulate_delta(const float *L, uint32_t stride, uint32_t dims, const uint32_t *sampled_nodes, uint32_t batch,
                                             uint32_t n_nodes, float *loss_delta, uint32_t *samples_delta, cudaStream_t stream)
{
    cudaError_t e;
    if ((e = cudaMemsetAsync(loss_delta, 0, (size_t)n_nodes * 4, stream)) != cudaSuccess)
        return e;
    if ((e = cudaMemsetAsync(samples_delta, 0, (size_t)n_nodes * 4, stream)) != cudaSuccess)
        return e;
    if (batch == 0)
        return cudaSuccess;
    k_tc_accumulate_delta<<<(batch + 255u) / 256u, 256, 0, stream>>>(L, stride, dims, sampled_nodes, batch, loss_delta, samples_delta);
    return cudaGetLastError();
}

cudaError_t launch_tree_cut_apply_delta(nbvh_tree_cut_data *cut, uint32_t n, const float *loss_delta, const uint32_t *samples_delta, cudaStream_t stream)
{
    if (n == 0)
        return cudaSuccess;
    k_tc_apply_delta<<<(n + 255u) / 256u, 256, 0, stream>>>(cut, n, loss_delta, samples_delta);
    return cudaGetLastError();
}

cudaError_t launch_tree_cut_decide_and_split(float4 *nodes, uint32_t n, uint32_t capacity, nbvh_tree_cut_data *cut, uint32_t *node_map,
                                             const float4 *blas_nodes, uint32_t num_splits, uint32_t current_learned_lod, void *workspace,
                                             size_t workspace_bytes, TreeCutResult *result_dev, cudaStream_t stream)
{
    size_t sort_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, sort_bytes, (const float *)nullptr, (float *)nullptr, (const uint32_t *)nullptr,
                                              (uint32_t *)nullptr, (int)capacity);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)capacity);
    size_t tmp_bytes = ((sort_bytes > scan_bytes ? sort_bytes : scan_bytes) + 255) & ~(size_t)255;
    if (tree_cut_workspace_bytes(capacity) > workspace_bytes || n > capacity)
        return cudaErrorInvalidValue;
    uint8_t *w       = (uint8_t *)workspace;
    void *tmp        = w;
    uint32_t *words  = (uint32_t *)(w + tmp_bytes);
    float *keys_in   = (float *)words, *keys_out = (float *)(words + capacity);
    uint32_t *vals_in = words + 2 * (size_t)capacity, *vals_out = words + 3 * (size_t)capacity;
    uint32_t *flags = words + 4 * (size_t)capacity, *rank = words + 5 * (size_t)capacity;
    uint32_t *split_flag = words + 6 * (size_t)capacity, *split_rank = words + 7 * (size_t)capacity;
    const uint32_t blocks = (n + 255u) / 256u;
    cudaError_t e;
    k_tc_candidates<<<blocks, 256, 0, stream>>>(nodes, n, cut, keys_in, vals_in);
    if ((e = cub::DeviceRadixSort::SortPairsDescending(tmp, sort_bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0, 32, stream)) != cudaSuccess)
        return e;
    k_tc_flag_distinct<<<blocks, 256, 0, stream>>>(keys_out, n, flags);
    if ((e = cub::DeviceScan::ExclusiveSum(tmp, scan_bytes, flags, rank, (int)n, stream)) != cudaSuccess)
        return e;
    if ((e = cudaMemsetAsync(split_flag, 0, (size_t)n * 4, stream)) != cudaSuccess)
        return e;
    // "Make sure we don't split more nodes than possible at low depth": at most (#nodes + 1) / 2 of the k candidates
    uint32_t limit = num_splits < (n + 1u) / 2u ? num_splits : (n + 1u) / 2u;
    if (2u * (size_t)limit + n > capacity)
        limit = (capacity - n) / 2u;
    k_tc_mark<<<blocks, 256, 0, stream>>>(keys_out, vals_out, flags, rank, n, limit, blas_nodes, node_map, split_flag, &result_dev->max_node_loss);
    if ((e = cub::DeviceScan::ExclusiveSum(tmp, scan_bytes, split_flag, split_rank, (int)n, stream)) != cudaSuccess)
        return e;
    k_tc_split<<<blocks, 256, 0, stream>>>(nodes, n, split_flag, split_rank, blas_nodes, node_map, cut, current_learned_lod, &result_dev->n_split,
                                           &result_dev->max_depth);
    return cudaGetLastError();
}

cudaError_t launch_tree_cut_assign_lod(float4 *nodes, uint32_t n, uint32_t level, cudaStream_t stream)
{
    if (n == 0)
        return cudaSuccess;
    k_tc_assign_lod<<<(n + 255u) / 256u, 256, 0, stream>>>(nodes, n, level);
    return cudaGetLastError();
}

} // namespace nbvh

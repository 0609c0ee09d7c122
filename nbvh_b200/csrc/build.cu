// build.cu -- device-side BLAS construction: an LBVH over the triangles of one mesh in the reference's
// 32-byte BVH2Node layout (include/base/accel/bvh.h:66-71: aabb_min, prim_count | aabb_max, left_or_prim_offset; the two
// children of an interior node are adjacent, root = node 0), replacing the host-side ext/bvh build of
// CudaBackend::build_blas (/root/reference/src/cuda/cuda_backend.cpp:186-385) when no ext/bvh tree is needed for
// compatibility with shipped .nbvh files (those are cuts of the ext/bvh tree and keep using it).
//
//   k_tri_bounds      triangle AABBs (flat triangles padded like the reference's bbox regularisation, 1e-4) + scene box
//   k_morton          27-bit Morton code of each centroid inside the scene box
//   cub radix sort    (code, triangle)
//   k_hierarchy       Karras 2012: every interior node finds its range and split from the common-prefix function
//                     (ties broken by the index, so duplicate codes form balanced subtrees)
//   k_fit             bottom-up AABB fit, the second child to arrive continues upward; also records the depth
//
// Node numbering: interior node k (Karras index, root k = 0) stores its children in the adjacent slots 1 + 2k and 2 + 2k;
// leaves hold one triangle (prim_count 1, left_or_prim_offset = position in the sorted primitive array).  2n - 1 nodes.
#include "kernels.cuh"

#include <cub/cub.cuh>

namespace nbvh {

__device__ __forceinline__ uint32_t expand_bits9(uint32_t v) // 9 bits -> every third bit
{
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

__device__ __forceinline__ void atomic_min_float(float *addr, float v)
{
    // monotone mapping float -> int for non-NaN values
    if (v >= 0.f)
        atomicMin(reinterpret_cast<int *>(addr), __float_as_int(v));
    else
        atomicMax(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float *addr, float v)
{
    if (v >= 0.f)
        atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
    else
        atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

__global__ void k_tri_bounds(const float *__restrict__ vertices, const uint32_t *__restrict__ indices, uint32_t n, float4 *tri_lo, float4 *tri_hi,
                             float *scene /* min xyz, max xyz */)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < n) {
        for (int k = 0; k < 3; ++k) {
            const uint32_t v = indices[3u * i + k];
            for (int a = 0; a < 3; ++a) {
                const float x = vertices[3u * v + a];
                lo[a] = fminf(lo[a], x), hi[a] = fmaxf(hi[a], x);
            }
        }
        for (int a = 0; a < 3; ++a) { // bbox regularisation of flat boxes (cuda_backend.cpp: 1e-4)
            if (hi[a] - lo[a] <= 1e-4f)
                lo[a] -= 5e-5f, hi[a] += 5e-5f;
        }
        tri_lo[i] = make_float4(lo[0], lo[1], lo[2], 0.f);
        tri_hi[i] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
    // warp + block reduction, one atomic set per block
    __shared__ float s_lo[8][3], s_hi[8][3];
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xFFFFFFFFu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xFFFFFFFFu, hi[a], o));
        }
        if ((threadIdx.x & 31u) == 0u)
            s_lo[threadIdx.x >> 5][a] = lo[a], s_hi[threadIdx.x >> 5][a] = hi[a];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int a = threadIdx.x;
        float l = s_lo[0][a], h = s_hi[0][a];
        for (int w = 1; w < 8; ++w)
            l = fminf(l, s_lo[w][a]), h = fmaxf(h, s_hi[w][a]);
        if (l <= h) {
            atomic_min_float(scene + a, l);
            atomic_max_float(scene + 3 + a, h);
        }
    }
}

__global__ void k_morton(const float4 *__restrict__ tri_lo, const float4 *__restrict__ tri_hi, uint32_t n, const float *__restrict__ scene,
                         uint32_t *codes, uint32_t *ids)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const float4 lo = tri_lo[i], hi = tri_hi[i];
    const float c[3] = {0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z)};
    uint32_t q[3];
    for (int a = 0; a < 3; ++a) {
        const float ext = scene[3 + a] - scene[a];
        const float u   = ext > 0.f ? (c[a] - scene[a]) / ext : 0.f;
        q[a]            = (uint32_t)fminf(fmaxf(u * 512.f, 0.f), 511.f);
    }
    codes[i] = (expand_bits9(q[0]) << 2) | (expand_bits9(q[1]) << 1) | expand_bits9(q[2]);
    ids[i]   = i;
}

// common-prefix length of the keys at i and j (key = 27-bit code, then the index as tie-breaker); -1 outside the array
__device__ __forceinline__ int delta(const uint32_t *__restrict__ codes, int n, int i, int j)
{
    if (j < 0 || j >= n)
        return -1;
    const uint32_t a = codes[i], b = codes[j];
    if (a != b)
        return __clz(a ^ b);
    return 32 + __clz((uint32_t)i ^ (uint32_t)j);
}

// Karras, "Maximizing Parallelism in the Construction of BVHs, Octrees, and k-d Trees" (HPG 2012), Fig. 4
__global__ void k_hierarchy(const uint32_t *__restrict__ codes, const uint32_t *__restrict__ ids, int n, float4 *nodes, uint32_t *parent_slot,
                            unsigned long long *prim_indices, const float4 *__restrict__ tri_lo, const float4 *__restrict__ tri_hi)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) // sorted primitive array (size_t in the reference)
        prim_indices[i] = ids[i];
    if (i >= n - 1)
        return;
    const int d = (delta(codes, n, i, i + 1) - delta(codes, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(codes, n, i, i - d);
    int lmax = 2;
    while (delta(codes, n, i, i + lmax * d) > dmin)
        lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (delta(codes, n, i, i + (l + t) * d) > dmin)
            l += t;
    const int j     = i + l * d;
    const int dnode = delta(codes, n, i, j);
    int s = 0;
    for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
        if (delta(codes, n, i, i + (s + t) * d) > dnode)
            s += t;
        if (t == 1)
            break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    // children: leaf when the child's range is a single primitive
    const bool left_leaf = lo == gamma, right_leaf = hi == gamma + 1;
    // The children of interior node i live in the adjacent slots 1 + 2i and 2 + 2i (known without the parent).  An interior
    // child records which slot it got in the scratch array slot_of_interior (= parent_slot + 2n - 1); k_link then writes
    // the interior headers and the parent links.
    const uint32_t ls = 1u + 2u * (uint32_t)i, rs = ls + 1u;
    uint32_t *slot_of_interior = parent_slot + (2 * n - 1);
    if (i == 0)
        slot_of_interior[0] = 0u;
    if (!left_leaf)
        slot_of_interior[gamma] = ls;
    if (!right_leaf)
        slot_of_interior[gamma + 1] = rs;
    if (left_leaf) {
        const uint32_t t = ids[gamma];
        const float4 a = tri_lo[t], b = tri_hi[t];
        nodes[2u * ls]      = make_float4(a.x, a.y, a.z, __uint_as_float(1u));
        nodes[2u * ls + 1u] = make_float4(b.x, b.y, b.z, __uint_as_float((uint32_t)gamma));
    }
    if (right_leaf) {
        const uint32_t t = ids[gamma + 1];
        const float4 a = tri_lo[t], b = tri_hi[t];
        nodes[2u * rs]      = make_float4(a.x, a.y, a.z, __uint_as_float(1u));
        nodes[2u * rs + 1u] = make_float4(b.x, b.y, b.z, __uint_as_float((uint32_t)(gamma + 1)));
    }
}

// second pass: now that every interior node knows its slot, write interior headers (left child slot) and parent links
__global__ void k_link(int n, float4 *nodes, uint32_t *parent_slot)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1)
        return;
    const uint32_t *slot_of_interior = parent_slot + (2 * n - 1);
    const uint32_t me = slot_of_interior[i];
    const uint32_t ls = 1u + 2u * (uint32_t)i;
    nodes[2u * me].w      = __uint_as_float(0u); // prim_count 0 = interior
    nodes[2u * me + 1u].w = __uint_as_float(ls);
    parent_slot[ls]       = me;
    parent_slot[ls + 1u]  = me;
    if (i == 0)
        parent_slot[0] = 0xFFFFFFFFu;
}

// bottom-up fit: one thread per leaf slot climbs; the second arrival at a node merges the children boxes
__global__ void k_fit(int n, float4 *nodes, const uint32_t *__restrict__ parent_slot, uint32_t *arrivals)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= 2 * n - 1)
        return;
    if (__float_as_uint(nodes[2u * s].w) == 0u)
        return; // interior: reached from below
    uint32_t cur = (uint32_t)s;
    for (;;) {
        const uint32_t p = parent_slot[cur];
        if (p == 0xFFFFFFFFu)
            break;
        __threadfence();
        if (atomicAdd(arrivals + p, 1u) == 0u)
            return; // first child: the sibling will finish this node
        const uint32_t l = __float_as_uint(__ldcg(&nodes[2u * p + 1u]).w);
        // L1 is not coherent across SMs: read the children (one of them written by another thread) through L2
        const float4 a0 = __ldcg(&nodes[2u * l]), a1 = __ldcg(&nodes[2u * l + 1u]), b0 = __ldcg(&nodes[2u * (l + 1u)]),
                     b1 = __ldcg(&nodes[2u * (l + 1u) + 1u]);
        nodes[2u * p]      = make_float4(fminf(a0.x, b0.x), fminf(a0.y, b0.y), fminf(a0.z, b0.z), __uint_as_float(0u));
        nodes[2u * p + 1u] = make_float4(fmaxf(a1.x, b1.x), fmaxf(a1.y, b1.y), fmaxf(a1.z, b1.z), __uint_as_float(l));
        cur = p;
    }
}

// depth of the tree (top-down over the finished links): leaf depth = number of parent hops
__global__ void k_depth(int n, const float4 *__restrict__ nodes, const uint32_t *__restrict__ parent_slot, uint32_t *max_depth)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= 2 * n - 1 || __float_as_uint(nodes[2u * s].w) == 0u)
        return;
    uint32_t d = 0;
    for (uint32_t cur = (uint32_t)s; parent_slot[cur] != 0xFFFFFFFFu; cur = parent_slot[cur])
        ++d;
    atomicMax(max_depth, d);
}

size_t build_blas_workspace_bytes(uint32_t n_tris)
{
    size_t sort_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                    (int)n_tris, 0, 27);
    const size_t n = n_tris;
    // tri_lo, tri_hi (float4), codes/ids in+out, parent_slot (2n-1) + slot_of_interior (n), arrivals (2n-1), scene box + depth
    return ((sort_bytes + 255) & ~(size_t)255) + 2 * n * 16 + 4 * n * 4 + (3 * n) * 4 + (2 * n) * 4 + 256;
}

cudaError_t launch_build_blas(const float *vertices, const uint32_t *indices, uint32_t n_tris, float4 *nodes, unsigned long long *prim_indices,
                              void *workspace, size_t workspace_bytes, uint32_t *max_depth_dev, cudaStream_t stream)
{
    if (n_tris == 0 || build_blas_workspace_bytes(n_tris) > workspace_bytes)
        return cudaErrorInvalidValue;
    const size_t n = n_tris;
    size_t sort_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                    (int)n_tris, 0, 27);
    uint8_t *w = (uint8_t *)workspace;
    void *tmp  = w;
    w += (sort_bytes + 255) & ~(size_t)255;
    float4 *tri_lo = (float4 *)w, *tri_hi = tri_lo + n;
    uint32_t *codes_in = (uint32_t *)(tri_hi + n), *codes = codes_in + n, *ids_in = codes + n, *ids = ids_in + n;
    uint32_t *parent_slot = ids + n;          // 2n-1 parent links + n interior slots
    uint32_t *arrivals    = parent_slot + 3 * n; // 2n-1
    float *scene          = (float *)(arrivals + 2 * n);
    cudaError_t e;
    const float init[6] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
    if ((e = cudaMemcpyAsync(scene, init, sizeof(init), cudaMemcpyHostToDevice, stream)) != cudaSuccess)
        return e;
    if ((e = cudaMemsetAsync(arrivals, 0, 2 * n * 4, stream)) != cudaSuccess)
        return e;
    if ((e = cudaMemsetAsync(max_depth_dev, 0, 4, stream)) != cudaSuccess)
        return e;
    const uint32_t blocks = (n_tris + 255u) / 256u;
    k_tri_bounds<<<blocks, 256, 0, stream>>>(vertices, indices, n_tris, tri_lo, tri_hi, scene);
    if (n_tris == 1) { // a single leaf is the whole tree
        k_morton<<<1, 32, 0, stream>>>(tri_lo, tri_hi, 1, scene, codes, ids);
        cudaMemcpyAsync(nodes, tri_lo, 16, cudaMemcpyDeviceToDevice, stream);
        cudaMemcpyAsync(nodes + 1, tri_hi, 16, cudaMemcpyDeviceToDevice, stream);
        const uint32_t one = 1u, zero = 0u;
        cudaMemcpyAsync(&nodes[0].w, &one, 4, cudaMemcpyHostToDevice, stream);
        cudaMemcpyAsync(&nodes[1].w, &zero, 4, cudaMemcpyHostToDevice, stream);
        return cudaMemsetAsync(prim_indices, 0, 8, stream);
    }
    k_morton<<<blocks, 256, 0, stream>>>(tri_lo, tri_hi, n_tris, scene, codes_in, ids_in);
    if ((e = cub::DeviceRadixSort::SortPairs(tmp, sort_bytes, codes_in, codes, ids_in, ids, (int)n_tris, 0, 27, stream)) != cudaSuccess)
        return e;
    k_hierarchy<<<blocks, 256, 0, stream>>>(codes, ids, (int)n_tris, nodes, parent_slot, prim_indices, tri_lo, tri_hi);
    k_link<<<blocks, 256, 0, stream>>>((int)n_tris, nodes, parent_slot);
    const uint32_t blocks2 = (2u * n_tris - 1u + 255u) / 256u;
    k_fit<<<blocks2, 256, 0, stream>>>((int)n_tris, nodes, parent_slot, arrivals);
    k_depth<<<blocks2, 256, 0, stream>>>((int)n_tris, nodes, parent_slot, max_depth_dev);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ TLAS build (f3)
// == CudaBackend::build_tlas (cuda_backend.cpp:312-383) on the device: a binary tree over the root boxes of the per-mesh BLAS, one mesh
// per leaf, in the TLASNode layout the traversal reads (tlas.cuh:64-68: interior = {box, left child, 0} with the children adjacent,
// leaf = {box, offset into blas_indices, 1 | 2 for a neural mesh}).  A scene has a handful of meshes (the traversal kernels cap them at
// 16), so one thread runs an exact sweep-SAH build: per node the best of the (k - 1) x 3 centroid-ordered splits.
struct BlasDescDev { // layout of nbvh_cuda_blas (include/nbvh_b200.h)
    uint32_t num_nodes;
    const unsigned long long *prim_indices;
    uint32_t num_prim_indices;
    const float *vertices;
    const uint32_t *indices;
    const float *normals;
    const float *texcoords;
    const int *material_ids;
    const uint32_t *material_map;
    uint32_t num_materials;
    const float4 *nodes;
};

__device__ inline float half_area(const float *lo, const float *hi)
{
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return dx * dy + dy * dz + dz * dx;
}

// workspace per mesh: lo[3], hi[3], right_area (7 floats), order, neural flag, 3 words of task stack (5 uint32)
size_t build_tlas_workspace_bytes(uint32_t n) { return (size_t)n * (7u * sizeof(float) + 5u * sizeof(uint32_t)); }

__global__ void k_build_tlas(const BlasDescDev *meshes, uint32_t n, const uint32_t *neural_flags, float *ws_f, uint32_t *ws_u, float4 *nodes,
                             uint32_t *blas_indices)
{
    if (threadIdx.x != 0 || blockIdx.x != 0)
        return;
    float *lo = ws_f, *hi = ws_f + 3u * n, *right_area = ws_f + 6u * n;
    uint32_t *order = ws_u;
    for (uint32_t i = 0; i < n; i++) {
        const float4 a = meshes[i].nodes[0], b = meshes[i].nodes[1]; // root_bbox() of the mesh's BLAS
        lo[3 * i] = a.x, lo[3 * i + 1] = a.y, lo[3 * i + 2] = a.z, hi[3 * i] = b.x, hi[3 * i + 1] = b.y, hi[3 * i + 2] = b.z;
        order[i] = i;
    }
    struct Task {
        uint32_t node, begin, end;
    };
    Task *stack     = reinterpret_cast<Task *>(ws_u + 2u * n); // depth-first with the smaller... at most n pending tasks
    int ss          = 0;
    uint32_t n_nodes = 1;
    stack[ss++]     = {0u, 0u, n};
    while (ss > 0) {
        const Task t     = stack[--ss];
        const uint32_t k = t.end - t.begin;
        float blo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, bhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (uint32_t i = t.begin; i < t.end; i++)
            for (int c = 0; c < 3; c++)
                blo[c] = fminf(blo[c], lo[3 * order[i] + c]), bhi[c] = fmaxf(bhi[c], hi[3 * order[i] + c]);
        if (k == 1) {
            const uint32_t mesh = order[t.begin];
            nodes[2 * t.node]     = make_float4(blo[0], blo[1], blo[2], __uint_as_float(t.begin));
            nodes[2 * t.node + 1] = make_float4(bhi[0], bhi[1], bhi[2], __uint_as_float(neural_flags[mesh] ? 2u : 1u));
            continue;
        }
        int best_axis = 0;
        uint32_t best_split = 1;
        float best_cost = FLT_MAX;
        for (int pass = 0; pass < 4; pass++) { // 3 evaluation passes, then re-sort along the winning axis
            const int axis = pass < 3 ? pass : best_axis;
            for (uint32_t i = t.begin + 1; i < t.end; i++) { // insertion sort by centroid (stable: ties keep the mesh order)
                const uint32_t m = order[i];
                const float c    = lo[3 * m + axis] + hi[3 * m + axis];
                uint32_t j       = i;
                while (j > t.begin && lo[3 * order[j - 1] + axis] + hi[3 * order[j - 1] + axis] > c)
                    order[j] = order[j - 1], --j;
                order[j] = m;
            }
            if (pass == 3)
                break;
            float rlo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, rhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
            for (uint32_t i = t.end; i-- > t.begin + 1;) {
                for (int c = 0; c < 3; c++)
                    rlo[c] = fminf(rlo[c], lo[3 * order[i] + c]), rhi[c] = fmaxf(rhi[c], hi[3 * order[i] + c]);
                right_area[i - t.begin] = half_area(rlo, rhi);
            }
            float llo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, lhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
            for (uint32_t s = 1; s < k; s++) { // s meshes on the left
                for (int c = 0; c < 3; c++)
                    llo[c] = fminf(llo[c], lo[3 * order[t.begin + s - 1] + c]), lhi[c] = fmaxf(lhi[c], hi[3 * order[t.begin + s - 1] + c]);
                const float cost = half_area(llo, lhi) * (float)s + right_area[s] * (float)(k - s);
                if (cost < best_cost)
                    best_cost = cost, best_axis = axis, best_split = s;
            }
        }
        const uint32_t left = n_nodes;
        n_nodes += 2;
        nodes[2 * t.node]     = make_float4(blo[0], blo[1], blo[2], __uint_as_float(left));
        nodes[2 * t.node + 1] = make_float4(bhi[0], bhi[1], bhi[2], __uint_as_float(0u));
        stack[ss++] = {left + 1u, t.begin + best_split, t.end};
        stack[ss++] = {left, t.begin, t.begin + best_split};
    }
    for (uint32_t i = 0; i < n; i++)
        blas_indices[i] = order[i];
}

// workspace: build_tlas_workspace_bytes(num_meshes) of device memory; neural_flags: num_meshes device words (non-zero = neural mesh)
cudaError_t launch_build_tlas(const void *mesh_bvhs_device, uint32_t num_meshes, const uint32_t *neural_flags, void *workspace, float4 *tlas_nodes,
                              uint32_t *blas_indices, cudaStream_t stream)
{
    if (num_meshes == 0)
        return cudaErrorInvalidValue;
    float *ws_f    = reinterpret_cast<float *>(workspace);
    uint32_t *ws_u = reinterpret_cast<uint32_t *>(ws_f + 7u * (size_t)num_meshes);
    k_build_tlas<<<1, 32, 0, stream>>>(reinterpret_cast<const BlasDescDev *>(mesh_bvhs_device), num_meshes, neural_flags, ws_f, ws_u, tlas_nodes,
                                       blas_indices);
    return cudaGetLastError();
}

} // namespace nbvh

// sampler.cu -- online-optimisation sample generator (kernel K10): replaces
// granular_sample_bvh_network_input_data (/root/reference/src/cuda/neural/neural_bvh_module_kernels.cu:119-264) and the
// device functions it calls: stack_based_sample_uniform_bvh_density, generate_single_bvh_network_output_reference,
// compute_depth_or_position_output (include/cuda/neural/neural_bvh_module_common.cuh:586-861), BaseCudaBLAS::compute_its_data
// (include/cuda/accel/bvh.cuh:53-96), unpack_albedo / unpack_normal (include/cuda/base/material.cuh:66-89),
// shading_frame / ortho_basis / sample_lambertian_dir / square_to_uniform_sphere.
//
// One thread owns n_sub_batch consecutive batch slots (exactly like the reference: a ray that misses the mesh inside a
// leaf CONTINUES through the tree-cut and its next leaf fills the thread's next slot).  Per slot: sample (or continue) a ray,
// walk the tree-cut at a random LoD with a random node dilation (learning epsilons), node-survival roulette, encode the leaf
// segment, trace the leaf segment against the neural mesh's classical BLAS for the targets
// [hit, local depth, (n+1)/2, albedo].  Slots rejected by the roulette keep their previous contents (the reference's
// `continue` also skips the slot increment; replicated).
//
// Arithmetic: same contract as device_core.cuh (-fmad=false, explicit fma where the reference's SASS fuses).  The
// traversal, segment encoding and triangle test are the bit-exact building blocks pinned against the reference kernels;
// the sampling arithmetic (sphere sampling, barycentric interpolation, normal mapping) follows the source expression by
// expression with the contraction that reproduces the reference kernel's outputs BIT FOR BIT on the shipped Chess scene
// (node ids, inputs and all 8 targets of every slot; tests/test_sampler.py).  The optional training bounce
// (max_training_bounces > 1, off in every shipped config) is checked against the CPU restatement with a tolerance only.
#include "render_common.cuh"
#include "sampler.cuh"

#include <cstdlib>

namespace nbvh {

// include/cuda/utils/sampling.cuh:8-14
__device__ __forceinline__ V3 square_to_uniform_sphere(float sx, float sy)
{
    const float phi       = __fmul_rn(6.2831854820251465f, sx); // 2.0f * PI folded at compile time
    const float cos_theta = __fmaf_rn(sy, 2.0f, -1.0f);
    const float sin_theta = __fsqrt_rn(__fmaf_rn(-cos_theta, cos_theta, 1.0f));
    return V3{__fmul_rn(sin_theta, cosf(phi)), __fmul_rn(sin_theta, sinf(phi)), cos_theta};
}

// node survival roulette (neural_bvh_module_kernels.cu:181-198); consumes one random number
__device__ __forceinline__ bool node_survives(const SamplerParams &p, uint32_t node_idx, uint32_t &rng)
{
    const nbvh_tree_cut_data nd = p.tree_cut[node_idx];
    float est = -FLT_MAX;
    if (nd.samples != 0u) {
        const float ia = __uint2float_rn(nd.iter_alive);
        est = __fdiv_rn(nd.loss, __fmul_rn(ia, __fsqrt_rn(__fdiv_rn(__uint2float_rn(nd.samples), ia))));
    }
    float keep = fmaxf(p.min_rr_survival_prob, __fdiv_rn(est, p.current_max_loss));
    if (!(est <= p.current_max_loss))
        keep = 1.f;
    return !(__fmul_rn(p.survival_probability_scale, keep) < lcg_randomf(rng));
}

// Inputs (encode_segment_from_world_po_pi, neural_bvh_module_kernels.cu:10-97) and targets
// (generate_single_bvh_network_output_reference, neural_bvh_module_common.cuh:783-861) of one slot.  Returns the hit flag.
__device__ __forceinline__ bool emit_slot(const SamplerParams &p, uint32_t slot, const Ray &ray, float leaf_mint, float leaf_maxt,
                                          bool swap_dir, ItsData &its, uint32_t &cnt_blas, uint32_t &cnt_tri)
{
    {
        const float t_o = swap_dir ? leaf_maxt : leaf_mint, t_i = swap_dir ? leaf_mint : leaf_maxt;
        float eo[3], ei[3];
        encode_endpoints_raw(p.enc, ray, t_o, t_i, eo, ei);
        float *in = p.inputs + (size_t)slot * p.n_in;
        for (uint32_t k = 0; k < p.enc.n_points; ++k) {
            in[3u * k]      = lerp_point(eo[0], ei[0], p.enc.point_s[k]);
            in[3u * k + 1u] = lerp_point(eo[1], ei[1], p.enc.point_s[k]);
            in[3u * k + 2u] = lerp_point(eo[2], ei[2], p.enc.point_s[k]);
        }
    }
    Ray tr = ray;
    float seg_mint = leaf_mint, seg_maxt = leaf_maxt;
    if (swap_dir) {
        tr.dx = -ray.dx, tr.dy = -ray.dy, tr.dz = -ray.dz;
        ray_prepare_guarded(tr);
        seg_mint = -leaf_maxt, seg_maxt = -leaf_mint;
    }
    const TriHit hit = blas_closest_hit(p.blas, tr, seg_mint, seg_maxt, cnt_blas, cnt_tri);
    const bool is_hit = hit.prim != 0xFFFFFFFFu;
    float *tg = p.targets + (size_t)slot * p.n_out;
    uint32_t o = 0;
    tg[o] = is_hit ? 1.f : 0.f;
    o += p.vis_dim;
    if (is_hit && (p.normal_dim || p.albedo_dim || p.max_training_bounces > 0))
        its = compute_its_data(p.blas, p.attr, p.materials, p.textures, hit);
    if (p.depth_dim) { // LocalT: (t - mint) / ((double)len + 1e-3), evaluated in double like the reference's literal
        float depth = 0.f;
        if (is_hit) {
            const float len = __fsub_rn(seg_maxt, seg_mint);
            depth = __double2float_rn(__ddiv_rn((double)__fsub_rn(hit.t, seg_mint), __dadd_rn((double)len, 1e-3)));
            if (p.depth_linear)
                depth = __fmaf_rn(depth, 0.1f, 0.45f);
        }
        tg[o] = depth;
        o += p.depth_dim;
    }
    if (p.normal_dim) {
        tg[o]      = is_hit ? __fmul_rn(__fadd_rn(its.sh_normal.x, 1.f), 0.5f) : 0.f;
        tg[o + 1u] = is_hit ? __fmul_rn(__fadd_rn(its.sh_normal.y, 1.f), 0.5f) : 0.f;
        tg[o + 2u] = is_hit ? __fmul_rn(__fadd_rn(its.sh_normal.z, 1.f), 0.5f) : 0.f;
        o += p.normal_dim;
    }
    if (p.albedo_dim) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        if (is_hit) {
            a0 = a1 = a2 = 0.9f; // CudaMaterial{}.base_color
            if (its.material_id != -1 && p.materials) { // unpack_albedo (material.cuh:80-89)
                const MaterialPod &m = p.materials[its.material_id];
                const uint32_t mask  = __float_as_uint(m.base_color[0]);
                if (mask & 0x80000000u) {
                    const float4 tx = tex_fetch(p.textures, mask & 0x1fffffffu, its.u, its.v);
                    a0 = tx.x, a1 = tx.y, a2 = tx.z;
                } else {
                    a0 = m.base_color[0], a1 = m.base_color[1], a2 = m.base_color[2];
                }
            }
        }
        tg[o] = a0, tg[o + 1u] = a1, tg[o + 2u] = a2;
        o += p.albedo_dim;
    }
    return is_hit;
}

__global__ void __launch_bounds__(128)
k_generate_batch(const __grid_constant__ SamplerParams p)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.sub_batch_size)
        return;
    NbvhConst bc = p.nbvh;
    uint32_t nstack[kNeuralStackSize];
    int nss = 0;
    float leaf_mint = 0.f, leaf_maxt = 0.f;
    Ray ray;
    ray.ox = ray.oy = ray.oz = 0.f, ray.dx = ray.dy = ray.dz = 1.f;
    ray_prepare(ray, false);
    bool continue_traversal = false, swap_dir = false;
    int training_bounce = 0;
    uint32_t cnt_nbvh = 0, cnt_blas = 0, cnt_tri = 0, cnt_written = 0;

    uint32_t slot = (uint32_t)i * (uint32_t)p.n_sub_batch;
    for (int inner = 0; inner < p.n_sub_batch; ++inner) {
        uint32_t rng = lcg_rng_init(p.rng_index, 0u, slot);
        // ---------------- stack_based_sample_uniform_bvh_density (neural_bvh_module_common.cuh:703-781)
        const float s0 = lcg_randomf(rng), s1 = lcg_randomf(rng);
        const V3 dir   = square_to_uniform_sphere(s0, s1);
        bool valid     = continue_traversal && nss != 0;
        bool restarted = false;
        const uint32_t n_lods = p.num_bvh_lods - p.current_learned_lod;
        const uint32_t lod    = p.current_learned_lod + min((uint32_t)__fmul_rn(lcg_randomf(rng), __uint2float_rn(n_lods)), n_lods - 1u);
        bc.leaf_lod_max       = 2u + lod;
        bc.dilation           = __fmaf_rn(lcg_randomf(rng), p.node_extent_dilation, p.fixed_node_dilation);
        uint32_t node_idx     = 0u;
        // Phase-uniform retry loop: every iteration, lanes without a live walk draw a new origin (the direction is kept) and
        // every lane then visits exactly one node.  Same draws, same visiting order as the reference's
        // do { if (!valid) resample; valid = intersect(...) } while (!valid).
        {
            WalkState w;
            bool walking = valid;
            if (walking)
                walk_pop(bc, w, nstack, nss);
            for (;;) {
                if (!walking) {
                    nstack[0] = 0u, nss = 1;
                    const float r0 = lcg_randomf(rng), r1 = lcg_randomf(rng), r2 = lcg_randomf(rng);
                    const float off = __fmul_rn(__fsub_rn(p.sampling_bbox_scaling, 1.f), 0.5f);
                    const float lx = __fmaf_rn(p.sampling_bbox_scaling, r0, -off), ly = __fmaf_rn(p.sampling_bbox_scaling, r1, -off),
                                lz = __fmaf_rn(p.sampling_bbox_scaling, r2, -off);
                    ray.ox = __fmaf_rn(lx, p.bbox_ext[0], p.bbox_min[0]); // Bbox::local_to_bbox
                    ray.oy = __fmaf_rn(ly, p.bbox_ext[1], p.bbox_min[1]);
                    ray.oz = __fmaf_rn(lz, p.bbox_ext[2], p.bbox_min[2]);
                    ray.dx = dir.x, ray.dy = dir.y, ray.dz = dir.z;
                    ray_prepare_guarded(ray);
                    restarted = true;
                    walk_pop(bc, w, nstack, nss);
                    walking = true;
                }
                const WalkResult r = nbvh_walk_step(bc, ray, w, nstack, nss, FLT_MAX, leaf_mint, leaf_maxt, node_idx, cnt_nbvh);
                if (r == WALK_FOUND)
                    break;
                if (r == WALK_EXHAUSTED)
                    walking = false;
            }
        }
        p.sampled_node_indices[slot] = node_idx;
        if (restarted) {
            training_bounce = 0;
            swap_dir        = false;
        }
        if (swap_dir && leaf_mint >= 0.f)
            swap_dir = false;
        // ---------------- node survival roulette (neural_bvh_module_kernels.cu:181-198)
        if (p.training_node_survival && !node_survives(p, node_idx, rng))
            continue; // the slot is NOT advanced and keeps its previous contents (reference quirk)
        ItsData its;
        const bool is_hit = emit_slot(p, slot, ray, leaf_mint, leaf_maxt, swap_dir, its, cnt_blas, cnt_tri);
        ++cnt_written;
        swap_dir           = false;
        continue_traversal = !is_hit;
        // ---------------- optional training bounce (neural_bvh_module_kernels.cu:223-259)
        if (is_hit && p.max_training_bounces > 0) {
            ++training_bounce;
            V3 wo = v3(-ray.dx, -ray.dy, -ray.dz), shn = its.sh_normal, tangent, bitangent;
            bool backface = false;
            if (dot3(wo, shn) < 0.f) { // shading_frame (rendering.cuh:11-22)
                shn      = v3(-shn.x, -shn.y, -shn.z);
                backface = true;
            }
            ortho_basis(tangent, bitangent, shn);
            if ((backface && !p.backface_bounce_training) || training_bounce >= p.max_training_bounces) {
                training_bounce    = 0;
                continue_traversal = false;
            } else {
                const float u0 = lcg_randomf(rng), u1 = lcg_randomf(rng);
                V3 d;
                if (p.sample_cosine_weighted) {
                    d = sample_lambertian_dir(shn, tangent, bitangent, u0, u1);
                } else {
                    d = square_to_uniform_sphere(u0, u1);
                    if (!(__fmul_rn(dot3(wo, shn), dot3(d, shn)) > 0.f))
                        d = v3(-d.x, -d.y, -d.z);
                }
                ray.ox = its.position.x, ray.oy = its.position.y, ray.oz = its.position.z;
                ray.dx = d.x, ray.dy = d.y, ray.dz = d.z;
                ray_prepare_guarded(ray);
                continue_traversal = true;
                swap_dir           = true;
                nstack[0] = 0u, nss = 1;
            }
        }
        ++slot;
    }
    warp_add_counter(p.counters, CNT_NBVH_STEPS, cnt_nbvh);
    warp_add_counter(p.counters, CNT_BLAS_STEPS, cnt_blas);
    warp_add_counter(p.counters, CNT_TRI_TESTS, cnt_tri);
    warp_add_counter(p.counters, CNT_RAYS, cnt_written);
}

// ------------------------------------------------------------------------------------------------ n_sub_batch == 1
// One slot per reference thread means slots are independent (the RNG is keyed by the slot id), so a persistent warp may
// process them in any order.  Each lane runs the retry walk as a state machine; a lane whose ray found a surviving leaf
// PARKS, and when enough lanes are parked the warp traces their leaf segments together (emit_slot).  Idle lanes are
// refilled from a global slot counter.  Per slot the draws and the arithmetic are exactly those of k_generate_batch,
// so the outputs are bit-identical; only the lane <-> slot assignment changes.
constexpr uint32_t kParkThreshold   = 20u; // parked lanes that trigger the target pass
constexpr uint32_t kRefillThreshold = 8u;  // idle lanes that trigger a refill

__global__ void __launch_bounds__(128)
k_generate_batch_wq(const __grid_constant__ SamplerParams p, uint32_t *__restrict__ slot_counter)
{
    const uint32_t lane  = threadIdx.x & 31u;
    const uint32_t total = (uint32_t)p.sub_batch_size;
    NbvhConst bc = p.nbvh;
    uint32_t nstack[kNeuralStackSize];
    int nss = 0;
    float leaf_mint = 0.f, leaf_maxt = 0.f;
    Ray ray;
    ray.ox = ray.oy = ray.oz = 0.f, ray.dx = ray.dy = ray.dz = 1.f;
    ray_prepare(ray, false);
    WalkState w;
    w.node_idx = 0u, w.have_raw = false, w.cmax = w.cmin = 0.f;
    w.node.a = w.node.b = 0u;
    uint32_t rng = 0u, slot = 0u, node_idx = 0u;
    bool have_slot = false, walking = false, parked = false, exhausted = false;
    uint32_t cnt_nbvh = 0, cnt_blas = 0, cnt_tri = 0, cnt_written = 0;
    const float off = __fmul_rn(__fsub_rn(p.sampling_bbox_scaling, 1.f), 0.5f);
    const uint32_t n_lods = p.num_bvh_lods - p.current_learned_lod;

    for (;;) {
        // ---------------- refill idle lanes from the global counter (one atomic per refill)
        const uint32_t idle_mask = __ballot_sync(0xFFFFFFFFu, !have_slot);
        const uint32_t busy      = 32u - (uint32_t)__popc(idle_mask);
        if (!exhausted && ((uint32_t)__popc(idle_mask) >= kRefillThreshold || busy == 0u)) {
            const uint32_t want = (uint32_t)__popc(idle_mask);
            uint32_t base = 0u;
            if (lane == 0u)
                base = atomicAdd(slot_counter, want);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (base + want >= total)
                exhausted = true;
            if (!have_slot) {
                const uint32_t sl = base + (uint32_t)__popc(idle_mask & ((1u << lane) - 1u));
                if (sl < total) {
                    slot = sl;
                    rng  = lcg_rng_init(p.rng_index, 0u, slot);
                    const float s0 = lcg_randomf(rng), s1 = lcg_randomf(rng);
                    const V3 dir   = square_to_uniform_sphere(s0, s1);
                    ray.dx = dir.x, ray.dy = dir.y, ray.dz = dir.z;
                    const uint32_t lod = p.current_learned_lod + min((uint32_t)__fmul_rn(lcg_randomf(rng), __uint2float_rn(n_lods)), n_lods - 1u);
                    bc.leaf_lod_max = 2u + lod;
                    bc.dilation     = __fmaf_rn(lcg_randomf(rng), p.node_extent_dilation, p.fixed_node_dilation);
                    have_slot = true, walking = false, parked = false;
                }
            }
        }
        const bool active = have_slot && !parked;
        if (!__any_sync(0xFFFFFFFFu, have_slot))
            break; // nothing in flight and the counter is exhausted
        // ---------------- one node visit for every lane with a live slot
        if (active) {
            if (!walking) {
                nstack[0] = 0u, nss = 1;
                const float r0 = lcg_randomf(rng), r1 = lcg_randomf(rng), r2 = lcg_randomf(rng);
                const float lx = __fmaf_rn(p.sampling_bbox_scaling, r0, -off), ly = __fmaf_rn(p.sampling_bbox_scaling, r1, -off),
                            lz = __fmaf_rn(p.sampling_bbox_scaling, r2, -off);
                ray.ox = __fmaf_rn(lx, p.bbox_ext[0], p.bbox_min[0]);
                ray.oy = __fmaf_rn(ly, p.bbox_ext[1], p.bbox_min[1]);
                ray.oz = __fmaf_rn(lz, p.bbox_ext[2], p.bbox_min[2]);
                ray_prepare_guarded(ray);
                walk_pop(bc, w, nstack, nss);
                walking = true;
            }
            const WalkResult r = nbvh_walk_step(bc, ray, w, nstack, nss, FLT_MAX, leaf_mint, leaf_maxt, node_idx, cnt_nbvh);
            if (r == WALK_EXHAUSTED) {
                walking = false;
            } else if (r == WALK_FOUND) {
                p.sampled_node_indices[slot] = node_idx;
                if (p.training_node_survival && !node_survives(p, node_idx, rng))
                    have_slot = false; // rejected: the slot keeps its previous contents
                else
                    parked = true;
            }
        }
        // ---------------- target pass once enough lanes are parked (or nothing else is left to walk)
        const uint32_t parked_mask = __ballot_sync(0xFFFFFFFFu, parked);
        const bool any_walking     = __any_sync(0xFFFFFFFFu, have_slot && !parked);
        if (parked_mask && ((uint32_t)__popc(parked_mask) >= kParkThreshold || !any_walking)) {
            if (parked) {
                ItsData its;
                emit_slot(p, slot, ray, leaf_mint, leaf_maxt, false, its, cnt_blas, cnt_tri);
                ++cnt_written;
                parked = false, have_slot = false;
            }
        }
    }
    warp_add_counter(p.counters, CNT_NBVH_STEPS, cnt_nbvh);
    warp_add_counter(p.counters, CNT_BLAS_STEPS, cnt_blas);
    warp_add_counter(p.counters, CNT_TRI_TESTS, cnt_tri);
    warp_add_counter(p.counters, CNT_RAYS, cnt_written);
}

cudaError_t launch_generate_batch(const SamplerParams &p, uint32_t *slot_counter, int num_sms, cudaStream_t stream)
{
    if (p.sub_batch_size <= 0)
        return cudaSuccess;
    if (p.n_sub_batch == 1 && slot_counter && !getenv("NBVH_SAMPLER_SIMPLE")) {
        cudaError_t e = cudaMemsetAsync(slot_counter, 0, sizeof(uint32_t), stream);
        if (e != cudaSuccess)
            return e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_generate_batch_wq, 128, 0);
        if (e != cudaSuccess)
            return e;
        const uint32_t want = ((uint32_t)p.sub_batch_size + 127u) / 128u;
        uint32_t grid = (uint32_t)(num_sms * (per_sm > 0 ? per_sm : 1));
        if (grid > want)
            grid = want;
        k_generate_batch_wq<<<grid, 128, 0, stream>>>(p, slot_counter);
        return cudaGetLastError();
    }
    k_generate_batch<<<((uint32_t)p.sub_batch_size + 127u) / 128u, 128, 0, stream>>>(p);
    return cudaGetLastError();
}

} // namespace nbvh

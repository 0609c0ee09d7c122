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
// slab test of the scene box with the plain IEEE divisions of Ray3f::intersect_bbox_min_max (ray.cuh:116-132): WorldT depth targets
__device__ __forceinline__ void scene_box_interval(const SamplerParams &p, const Ray &r, float &mint, float &maxt)
{
    box_interval_ieee(p.bbox_min, p.bbox_max, r, mint, maxt);
}

// rng: the slot's generator after the roulette draw (SegmentDistribution::Uniform draws one number per segment point from it)
__device__ __forceinline__ bool emit_slot(const SamplerParams &p, uint32_t slot, const Ray &ray, float leaf_mint, float leaf_maxt,
                                          bool swap_dir, uint32_t node_idx, uint32_t &rng, ItsData &its, uint32_t &cnt_blas, uint32_t &cnt_tri)
{
    {
        const float t_o = swap_dir ? leaf_maxt : leaf_mint, t_i = swap_dir ? leaf_mint : leaf_maxt;
        float eo[3], ei[3];
        encode_endpoints_raw(p.enc, ray, t_o, t_i, eo, ei);
        float *in = p.inputs + (size_t)slot * p.n_in;
        for (uint32_t k = 0; k < p.enc.n_points; ++k) {
            // Uniform (neural_bvh_module_kernels.cu:53-69): t = lcg_randomf(rng) per point, same blend expression
            const float sk  = p.uniform_points ? lcg_randomf(rng) : p.enc.point_s[k];
            in[3u * k]      = lerp_point(eo[0], ei[0], sk);
            in[3u * k + 1u] = lerp_point(eo[1], ei[1], sk);
            in[3u * k + 2u] = lerp_point(eo[2], ei[2], sk);
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
    if (p.depth_dim) { // compute_depth_or_position_output (neural_bvh_module_common.cuh:586-629)
        float d3[3] = {0.f, 0.f, 0.f};
        if (is_hit) {
            if (p.depth_mode <= 1u) {
                // LocalT / WorldT: (t - mint) / ((double)len + 1e-3), evaluated in double like the reference's literal; the interval is
                // the leaf's (LocalT) or the scene box's along the traced ray (WorldT)
                float lo = seg_mint, hi = seg_maxt;
                if (p.depth_mode == 1u)
                    scene_box_interval(p, tr, lo, hi);
                const float len = __fsub_rn(hi, lo);
                d3[0] = __double2float_rn(__ddiv_rn((double)__fsub_rn(hit.t, lo), __dadd_rn((double)len, 1e-3)));
                d3[1] = d3[2] = d3[0];
            } else {
                // Position / LocalPosition: bbox_to_local(ray(t)) of the scene box, or of the node with the sampled node's INDEX in the
                // neural mesh's BLAS array (the reference indexes neural_blas.nodes with the tree-cut node id: replicated)
                const float px = __fmaf_rn(hit.t, tr.dx, tr.ox), py = __fmaf_rn(hit.t, tr.dy, tr.oy), pz = __fmaf_rn(hit.t, tr.dz, tr.oz);
                float mn[3] = {p.bbox_min[0], p.bbox_min[1], p.bbox_min[2]}, mx[3] = {p.bbox_max[0], p.bbox_max[1], p.bbox_max[2]};
                if (p.depth_mode == 3u) {
                    const Node n = load_node(p.blas.nodes, node_idx);
                    mn[0] = n.mnx, mn[1] = n.mny, mn[2] = n.mnz, mx[0] = n.mxx, mx[1] = n.mxy, mx[2] = n.mxz;
                }
                d3[0] = __fdiv_rn(__fsub_rn(px, mn[0]), __fsub_rn(mx[0], mn[0]));
                d3[1] = __fdiv_rn(__fsub_rn(py, mn[1]), __fsub_rn(mx[1], mn[1]));
                d3[2] = __fdiv_rn(__fsub_rn(pz, mn[2]), __fsub_rn(mx[2], mn[2]));
            }
            if (p.depth_linear)
                d3[0] = __fmaf_rn(d3[0], 0.1f, 0.45f), d3[1] = __fmaf_rn(d3[1], 0.1f, 0.45f), d3[2] = __fmaf_rn(d3[2], 0.1f, 0.45f);
        }
        tg[o] = d3[0];
        if (p.depth_mode >= 2u)
            tg[o + 1u] = d3[1], tg[o + 2u] = d3[2];
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
            int failed_attempts = 0;
            for (;;) {
                if (!walking) {
                    nstack[0] = 0u, nss = 1;
                    // view_dependent_sampling (neural_bvh_module_common.cuh:734-737): the first five attempts of a slot may take a camera
                    // ray through a random image position instead of a random origin along the slot's direction
                    bool camera = false;
                    if (p.view_dependent_sampling && failed_attempts < 5)
                        camera = lcg_randomf(rng) < p.view_dependent_ratio;
                    if (camera) {
                        // camera_raygen(lcg_randomf(rng), lcg_randomf(rng)): the device compiler evaluates the arguments left to right
                        // (pinned by the reference kernel's outputs, tests/test_sampler.py)
                        const float xf = lcg_randomf(rng), yf = lcg_randomf(rng);
                        const float vx = __fadd_rn(__fmaf_rn(xf, p.cam.du[0], __fmul_rn(yf, p.cam.dv[0])), p.cam.top_left[0]);
                        const float vy = __fadd_rn(__fmaf_rn(xf, p.cam.du[1], __fmul_rn(yf, p.cam.dv[1])), p.cam.top_left[1]);
                        const float vz = __fadd_rn(__fmaf_rn(xf, p.cam.du[2], __fmul_rn(yf, p.cam.dv[2])), p.cam.top_left[2]);
                        const float inv = __frcp_rn(__fsqrt_rn(__fmaf_rn(vz, vz, __fmaf_rn(vx, vx, __fmul_rn(vy, vy)))));
                        ray.ox = p.cam.origin[0], ray.oy = p.cam.origin[1], ray.oz = p.cam.origin[2];
                        ray.dx = __fmul_rn(vx, inv), ray.dy = __fmul_rn(vy, inv), ray.dz = __fmul_rn(vz, inv);
                    } else {
                        const float r0 = lcg_randomf(rng), r1 = lcg_randomf(rng), r2 = lcg_randomf(rng);
                        const float off = __fmul_rn(__fsub_rn(p.sampling_bbox_scaling, 1.f), 0.5f);
                        const float lx = __fmaf_rn(p.sampling_bbox_scaling, r0, -off), ly = __fmaf_rn(p.sampling_bbox_scaling, r1, -off),
                                    lz = __fmaf_rn(p.sampling_bbox_scaling, r2, -off);
                        ray.ox = __fmaf_rn(lx, p.bbox_ext[0], p.bbox_min[0]); // Bbox::local_to_bbox
                        ray.oy = __fmaf_rn(ly, p.bbox_ext[1], p.bbox_min[1]);
                        ray.oz = __fmaf_rn(lz, p.bbox_ext[2], p.bbox_min[2]);
                        ray.dx = dir.x, ray.dy = dir.y, ray.dz = dir.z;
                    }
                    ray_prepare_guarded(ray);
                    restarted = true;
                    walk_pop(bc, w, nstack, nss);
                    walking = true;
                }
                const WalkResult r = nbvh_walk_step(bc, ray, w, nstack, nss, FLT_MAX, leaf_mint, leaf_maxt, node_idx, cnt_nbvh);
                if (r == WALK_FOUND)
                    break;
                if (r == WALK_EXHAUSTED) {
                    walking = false;
                    ++failed_attempts;
                }
            }
        }
        // leaf_boundary_sampling_ratio (neural_bvh_module_common.cuh:765-777, sample_bbox_position :636-660): the segment becomes a chord
        // between two random points on two different faces of the sampled node's dilated box
        if (p.leaf_boundary_sampling_ratio != 0.f && lcg_randomf(rng) <= p.leaf_boundary_sampling_ratio) {
            const int face1 = min((int)__fmul_rn(lcg_randomf(rng), 6.f), 5);
            int face2;
            do {
                face2 = min((int)__fmul_rn(lcg_randomf(rng), 6.f), 5);
            } while (face1 == face2);
            float l1[3], l2[3];
            {   // get_cube_pos_on_face(face, lcg_randomf(rng), lcg_randomf(rng)): arguments evaluated left to right
                const float u1 = lcg_randomf(rng), v1 = lcg_randomf(rng);
                const float u2 = lcg_randomf(rng), v2 = lcg_randomf(rng);
                const float fixed1 = (face1 & 1) ? 0.f : 1.f, fixed2 = (face2 & 1) ? 0.f : 1.f;
                const int ax1 = face1 >> 1, ax2 = face2 >> 1; // faces 0,1: x fixed; 2,3: y fixed; 4,5: z fixed
                l1[0] = ax1 == 0 ? fixed1 : u1, l1[1] = ax1 == 1 ? fixed1 : (ax1 == 0 ? u1 : v1), l1[2] = ax1 == 2 ? fixed1 : v1;
                l2[0] = ax2 == 0 ? fixed2 : u2, l2[1] = ax2 == 1 ? fixed2 : (ax2 == 0 ? u2 : v2), l2[2] = ax2 == 2 ? fixed2 : v2;
            }
            const Node nd = load_node(p.nbvh.nodes, node_idx);
            const float sc = __fadd_rn(1.f, bc.dilation);
            const float mn[3] = {nd.mnx, nd.mny, nd.mnz}, mx[3] = {nd.mxx, nd.mxy, nd.mxz};
            float en[3], ex[3];
            for (int k = 0; k < 3; ++k) {
                const float e = __fsub_rn(mx[k], mn[k]);
                const float ext = __fmul_rn(e, sc);                                                       // bbox_extent() * (1 + dilation)
                const float nmin = __fsub_rn(__fadd_rn(mn[k], __fmul_rn(e, 0.5f)), __fmul_rn(ext, 0.5f));  // bbox_center() - ext * 0.5
                en[k] = __fmaf_rn(l1[k], ext, nmin);
                ex[k] = __fmaf_rn(l2[k], ext, nmin);
            }
            const float dx = __fsub_rn(ex[0], en[0]), dy = __fsub_rn(ex[1], en[1]), dz = __fsub_rn(ex[2], en[2]);
            const float l2n = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
            const float inv = __frcp_rn(__fsqrt_rn(l2n));
            ray.ox = en[0], ray.oy = en[1], ray.oz = en[2];
            ray.dx = __fmul_rn(dx, inv), ray.dy = __fmul_rn(dy, inv), ray.dz = __fmul_rn(dz, inv);
            ray_prepare_guarded(ray);
            leaf_mint = 0.f, leaf_maxt = __fsqrt_rn(l2n);
            restarted = true;
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
        const bool is_hit = emit_slot(p, slot, ray, leaf_mint, leaf_maxt, swap_dir, node_idx, rng, its, cnt_blas, cnt_tri);
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
// One slot per reference thread means slots are independent (the RNG is keyed by the slot id), so they may be processed in any
// order, in two passes, and ahead of time.  Per slot the draws and the arithmetic are exactly those of k_generate_batch, so the
// outputs are bit-identical; only the lane <-> slot assignment changes.
//
//   k_sample_walk     persistent warps, nothing but the retry walk (85 % of a slot's work on Chess: 8 origins and 70 interior steps
//                     per slot on average, heavy-tailed): every lane visits ONE node per iteration (per-node state machine); a lane
//                     whose ray met a leaf writes the 48-byte item {node, rng, ray, interval} of its slot and refills at once from the
//                     global slot counter -- no lane waits for others.  The stack carries the raw slab result of every pushed (far)
//                     child, so a popped leaf needs no second slab test: the leaf branch is a handful of compares and the warp
//                     executes one instruction stream, the interior step.  (Same ray, same dilated box, same operations -> the
//                     stored values are the bits the reference's re-test computes.)  The direction of a slot never changes, so its
//                     reciprocals are refined once per slot, and whether the hoisted division needs its per-box guard is decided once
//                     per launch for the tree-cut (k_check_dilation_safe) and once per origin.  When the slot counter is exhausted
//                     and a warp is down to a few live slots, its lanes search the remaining origins of those slots in parallel
//                     (LCG skip-ahead; the lowest origin that meets a leaf is the one the sequential loop stops at).
//                     The walk depends on the tree-cut and the RNG only -- not on the network, nor on the per-node loss statistics --
//                     so ONE launch may walk the slots of several consecutive optimisation steps (SamplerParams::n_steps): a batch of
//                     2^18 slots gives every lane of the machine 1.5 slots (no steady state, warps run at 12 of 32 lanes); four steps'
//                     worth give it 6.
//   k_sample_targets  one thread per slot of ONE step: node id, node-survival roulette (per-node table of this step's statistics,
//                     k_node_keep), block-wide compaction of the survivors, then segment encoding, leaf-segment trace against the
//                     neural mesh's BLAS, surface attributes, texture fetches, targets (emit_slot).
struct SampleItem {
    uint4 a; // node, rng, ox, oy
    uint4 b; // oz, dx, dy, dz
    uint4 c; // leaf_mint, leaf_maxt, -, -
};

constexpr uint32_t kRefillThreshold = 4u; // idle lanes that trigger a refill
constexpr uint32_t kRestartVote     = 3u; // lanes waiting for a new origin that trigger the restart block
constexpr uint32_t kCoopLanes       = 8u; // live slots per warp at or below which the drained warp searches origins in parallel

// Per generation, over the current tree-cut: may the retry walk divide with the hoisted reciprocal WITHOUT the per-box guard of
// slab_raw?  The walk dilates every box with a per-sample factor s = 1 + delta, delta in [0, 1): per axis e = max - min,
// ctr = min + e/2, h = (e s)/2, box' = ctr -+ h.  If e and ctr are each 0 or of magnitude >= 2^-36 (then h is 0 or >= 2^-37), a
// coordinate of box' is 0 or >= 2^-60 (the difference of two floats of magnitude >= 2^-37 is a multiple of 2^-60), and with an
// origin whose coordinates are 0 or >= 2^-61 every numerator (coordinate - origin) is 0 or >= 2^-84: the precondition of
// hoisted_div (device_core.cuh) holds for every dilation, so the guard is decided once per origin instead of once per box.
__global__ void __launch_bounds__(256)
k_check_dilation_safe(const float4 *__restrict__ nodes, uint32_t n, uint32_t *unsafe)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const float4 lo = nodes[2u * i], hi = nodes[2u * i + 1u];
    const float mn[3] = {lo.x, lo.y, lo.z}, mx[3] = {hi.x, hi.y, hi.z};
    bool good = true;
    for (int k = 0; k < 3; ++k) {
        const float e = __fsub_rn(mx[k], mn[k]), ctr = __fadd_rn(mn[k], __fmul_rn(e, 0.5f));
        const float tiny = 1.4551915228366852e-11f /* 2^-36 */, big = 274877906944.0f /* 2^38 */;
        good = good && (e == 0.f || e >= tiny) && (ctr == 0.f || fabsf(ctr) >= tiny) && fabsf(mn[k]) <= big && fabsf(mx[k]) <= big;
    }
    if (!good) // NaN coordinates fail the comparisons above
        *unsafe = 1u;
}

// scale * keep of node_survives() for every tree-cut node, once per generation (the roulette then costs one load per found leaf)
__global__ void __launch_bounds__(256)
k_node_keep(const __grid_constant__ SamplerParams p, uint32_t n_nodes)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes)
        return;
    const nbvh_tree_cut_data nd = p.tree_cut[i];
    float est = -FLT_MAX;
    if (nd.samples != 0u) {
        const float ia = __uint2float_rn(nd.iter_alive);
        est = __fdiv_rn(nd.loss, __fmul_rn(ia, __fsqrt_rn(__fdiv_rn(__uint2float_rn(nd.samples), ia))));
    }
    float keep = fmaxf(p.min_rr_survival_prob, __fdiv_rn(est, p.current_max_loss));
    if (!(est <= p.current_max_loss))
        keep = 1.f;
    p.node_keep[i] = __fmul_rn(p.survival_probability_scale, keep);
}

// lcg_randomf's generator advanced by n draws: x -> a x + c composed n times (binary powers of the affine map)
__device__ __forceinline__ uint32_t lcg_skip(uint32_t state, uint32_t n)
{
    uint32_t A = 1664525u, C = 1013904223u, acc_a = 1u, acc_c = 0u;
    while (n) {
        if (n & 1u) {
            acc_a = acc_a * A;
            acc_c = acc_c * A + C;
        }
        C = C * (A + 1u);
        A = A * A;
        n >>= 1;
    }
    return acc_a * state + acc_c;
}

// The retry walk of one ray as a per-node state machine.  The stack carries the raw slab result of every pushed (far) child.
struct CutWalk {
    uint32_t st_idx[kNeuralStackSize];
    float st_cmax[kNeuralStackSize], st_cmin[kNeuralStackSize];
    int nss;
    Node node; // only the tags a / b are maintained after the root
    uint32_t node_idx;
    float cmax, cmin;
    bool have_raw;
};

__device__ __forceinline__ void cut_walk_start(const NbvhConst &bc, CutWalk &w)
{
    w.nss      = 0; // the root is the current node (the reference pops it from the stack [0])
    w.node_idx = 0u;
    w.node     = load_node(bc.nodes, 0u);
    w.have_raw = false;
}

__device__ __forceinline__ bool cut_walk_pop(const NbvhConst &bc, CutWalk &w)
{
    if (w.nss == 0)
        return false;
    --w.nss;
    w.node_idx = w.st_idx[w.nss], w.cmax = w.st_cmax[w.nss], w.cmin = w.st_cmin[w.nss];
    // only the tags are needed: the box was tested when the node was pushed
    w.node.a   = __float_as_uint(__ldg(reinterpret_cast<const float *>(bc.nodes + 2u * w.node_idx) + 3));
    w.node.b   = __float_as_uint(__ldg(reinterpret_cast<const float *>(bc.nodes + 2u * w.node_idx + 1u) + 3));
    w.have_raw = true;
    return true;
}

// one node visit; returns WALK_FOUND (leaf interval in w.cmax / w.cmin, node w.node_idx), WALK_EXHAUSTED or WALK_CONTINUE
__device__ __forceinline__ WalkResult cut_walk_step(const NbvhConst &bc, const Ray &ray, CutWalk &w, uint32_t &steps)
{
    const bool is_leaf = (w.node.a > 0u && w.node.a <= bc.leaf_lod_max);
    if (is_leaf || w.node.b == kNoNeuralChild) {
        if (!w.have_raw) { // only a root that is a leaf itself arrives here without its slab result
            dilate_node(w.node, bc.dilation);
            slab_raw(w.node, ray, w.cmax, w.cmin);
        }
        // cur_maxt = FLT_MAX: the reference's `cmax < ray.maxt` only rejects cmax = +inf / NaN
        if (w.cmax <= w.cmin && (w.cmax < FLT_MAX) && ((w.cmax < 0.f) ? (w.cmin > bc.maxt_eps) : (w.cmax > bc.mint_eps)))
            return WALK_FOUND;
        return cut_walk_pop(bc, w) ? WALK_CONTINUE : WALK_EXHAUSTED;
    }
    ++steps;
    uint32_t c1 = w.node.b, c2 = w.node.b + 1u;
    Node n1 = load_node(bc.nodes, c1), n2 = load_node(bc.nodes, c2);
    dilate_node(n1, bc.dilation);
    dilate_node(n2, bc.dilation);
    float cmax1, cmin1, cmax2, cmin2;
    slab_raw(n1, ray, cmax1, cmin1);
    slab_raw(n2, ray, cmax2, cmin2);
    const float e1 = gmax(-FLT_MAX, cmax1), x1 = gmin(FLT_MAX, cmin1);
    const float e2 = gmax(-FLT_MAX, cmax2), x2 = gmin(FLT_MAX, cmin2);
    float m1 = (e1 <= x1) ? e1 : FLT_MAX;
    float m2 = (e2 <= x2) ? e2 : FLT_MAX;
    const bool swap = m1 > m2;
    if (swap) {
        const float t = m1;
        m1 = m2, m2 = t;
        const uint32_t ti = c1;
        c1 = c2, c2 = ti;
    }
    if (m1 == FLT_MAX)
        return cut_walk_pop(bc, w) ? WALK_CONTINUE : WALK_EXHAUSTED;
    w.node_idx = c1;
    w.node.a   = swap ? n2.a : n1.a;
    w.node.b   = swap ? n2.b : n1.b;
    w.cmax     = swap ? cmax2 : cmax1;
    w.cmin     = swap ? cmin2 : cmin1;
    w.have_raw = true;
    if (m2 != FLT_MAX) {
        w.st_idx[w.nss] = c2, w.st_cmax[w.nss] = swap ? cmax1 : cmax2, w.st_cmin[w.nss] = swap ? cmin1 : cmin2;
        ++w.nss;
        if (w.nss >= kNeuralStackSize)
            return WALK_EXHAUSTED; // reference: printf("Out of bound stack!") and the walk returns "no intersection"
    }
    return WALK_CONTINUE;
}

// sample a new origin in the scaled scene box (three draws; the direction of the slot is kept)
__device__ __forceinline__ void draw_origin(const SamplerParams &p, float off, uint32_t &rng, Ray &ray)
{
    const float r0 = lcg_randomf(rng), r1 = lcg_randomf(rng), r2 = lcg_randomf(rng);
    const float lx = __fmaf_rn(p.sampling_bbox_scaling, r0, -off), ly = __fmaf_rn(p.sampling_bbox_scaling, r1, -off),
                lz = __fmaf_rn(p.sampling_bbox_scaling, r2, -off);
    ray.ox = __fmaf_rn(lx, p.bbox_ext[0], p.bbox_min[0]); // Bbox::local_to_bbox
    ray.oy = __fmaf_rn(ly, p.bbox_ext[1], p.bbox_min[1]);
    ray.oz = __fmaf_rn(lz, p.bbox_ext[2], p.bbox_min[2]);
}

// ray.fast for a new origin: 1 = every slab test of this walk may use the hoisted reciprocal unguarded (k_check_dilation_safe said
// so for the tree-cut, the direction passed ray_prepare_guarded, and the origin is in range), 2 = guarded per box, 0 = div.rn
__device__ __forceinline__ void set_walk_division_mode(Ray &ray, bool dir_ok, bool cut_safe)
{
    const float lo = 4.336808689942018e-19f /* 2^-61 */, hi = 549755813888.0f /* 2^39 */;
    const bool origin_ok = in_range_or_zero(ray.ox, lo, hi) && in_range_or_zero(ray.oy, lo, hi) && in_range_or_zero(ray.oz, lo, hi);
    ray.fast = !dir_ok ? 0u : (cut_safe && origin_ok) ? 1u : 2u;
}

// the item of a slot whose walk met a leaf (position = step * batch + slot)
__device__ __forceinline__ void write_item(const SamplerParams &p, uint32_t idx, const CutWalk &w, uint32_t rng, const Ray &ray)
{
    uint4 *it = reinterpret_cast<uint4 *>(p.items) + 3u * (size_t)idx;
    it[0] = make_uint4(w.node_idx, rng, __float_as_uint(ray.ox), __float_as_uint(ray.oy));
    it[1] = make_uint4(__float_as_uint(ray.oz), __float_as_uint(ray.dx), __float_as_uint(ray.dy), __float_as_uint(ray.dz));
    it[2] = make_uint4(__float_as_uint(w.cmax), __float_as_uint(w.cmin), 0u, 0u);
}

// Everything a slot draws before its first origin: direction (fixed for the slot), LoD, node dilation.  rng ends at the state the
// origin draws continue from (origin k of the slot uses draws 3k+1 .. 3k+3 after it).  idx = step * batch + slot of a walk-ahead launch.
__device__ __forceinline__ bool begin_slot(const SamplerParams &p, uint32_t idx, uint32_t n_lods, uint32_t &rng, Ray &ray, NbvhConst &bc)
{
    const uint32_t step = idx / (uint32_t)p.sub_batch_size, slot = idx - step * (uint32_t)p.sub_batch_size;
    rng            = lcg_rng_init(p.rng_index + step * p.rng_index_step, 0u, slot);
    const float s0 = lcg_randomf(rng), s1 = lcg_randomf(rng);
    const V3 dir   = square_to_uniform_sphere(s0, s1);
    ray.dx = dir.x, ray.dy = dir.y, ray.dz = dir.z;
    ray_prepare_guarded(ray); // the direction is fixed for the slot: only origins are redrawn
    const uint32_t lod = p.current_learned_lod + min((uint32_t)__fmul_rn(lcg_randomf(rng), __uint2float_rn(n_lods)), n_lods - 1u);
    bc.leaf_lod_max = 2u + lod;
    bc.dilation     = __fmaf_rn(lcg_randomf(rng), p.node_extent_dilation, p.fixed_node_dilation);
    return ray.fast != 0u;
}

__global__ void __launch_bounds__(128)
k_sample_walk(const __grid_constant__ SamplerParams p, uint32_t *__restrict__ slot_counter)
{
    const uint32_t lane  = threadIdx.x & 31u;
    const uint32_t total = (uint32_t)p.sub_batch_size * p.n_steps;
    const bool cut_safe  = *reinterpret_cast<const volatile uint32_t *>(p.dilation_unsafe) == 0u;
    NbvhConst bc = p.nbvh;
    CutWalk w;
    w.nss = 0, w.node_idx = 0u, w.cmax = w.cmin = 0.f, w.have_raw = false;
    w.node.mnx = w.node.mny = w.node.mnz = w.node.mxx = w.node.mxy = w.node.mxz = 0.f, w.node.a = w.node.b = 0u;
    Ray ray;
    ray.ox = ray.oy = ray.oz = 0.f, ray.dx = ray.dy = ray.dz = 1.f;
    ray_prepare(ray, false);
    uint32_t rng = 0u, rng0 = 0u, idx = 0u, n_origins = 0u;
    bool have_slot = false, walking = false, exhausted = false, dir_ok = false;
    uint32_t cnt_nbvh = 0;
    const float off = __fmul_rn(__fsub_rn(p.sampling_bbox_scaling, 1.f), 0.5f);
    const uint32_t n_lods = p.num_bvh_lods - p.current_learned_lod;

    for (;;) {
        // ---------------- refill idle lanes from the global counter (one atomic per refill)
        const uint32_t idle_mask = __ballot_sync(0xFFFFFFFFu, !have_slot);
        const uint32_t n_idle    = (uint32_t)__popc(idle_mask);
        if (!exhausted && (n_idle >= kRefillThreshold || n_idle == 32u)) {
            uint32_t base = 0u;
            if (lane == 0u)
                base = atomicAdd(slot_counter, n_idle);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (base + n_idle >= total)
                exhausted = true;
            if (!have_slot) {
                const uint32_t sl = base + (uint32_t)__popc(idle_mask & ((1u << lane) - 1u));
                if (sl < total) {
                    idx       = sl;
                    dir_ok    = begin_slot(p, idx, n_lods, rng, ray, bc);
                    rng0      = rng;
                    n_origins = 0u;
                    have_slot = true, walking = false;
                }
            }
        }
        const uint32_t live_mask = __ballot_sync(0xFFFFFFFFu, have_slot);
        if (live_mask == 0u)
            break; // nothing in flight and the counter is exhausted
        if (exhausted && (uint32_t)__popc(live_mask) <= kCoopLanes)
            break; // drained warp: the few slots left are finished cooperatively below
        // ---------------- new origin for lanes whose walk ended without a leaf (voted: the block costs the whole warp)
        {
            const uint32_t need = __ballot_sync(0xFFFFFFFFu, have_slot && !walking);
            const uint32_t busy = __ballot_sync(0xFFFFFFFFu, have_slot && walking);
            if (need && ((uint32_t)__popc(need) >= kRestartVote || busy == 0u || (uint32_t)__popc(need) * 4u >= (uint32_t)__popc(busy))) {
                if (have_slot && !walking) {
                    draw_origin(p, off, rng, ray);
                    set_walk_division_mode(ray, dir_ok, cut_safe);
                    ++n_origins;
                    cut_walk_start(bc, w);
                    walking = true;
                }
            }
        }
        // ---------------- one node visit for every walking lane; a lane that met a leaf writes its item and becomes idle
        if (have_slot && walking) {
            const WalkResult r = cut_walk_step(bc, ray, w, cnt_nbvh);
            walking = r == WALK_CONTINUE;
            if (r == WALK_FOUND) {
                write_item(p, idx, w, rng, ray);
                have_slot = false;
            }
        }
    }

    // ---------------- cooperative tail.  The number of origins a slot needs is heavy-tailed (a fixed direction may rarely meet a
    // leaf: hundreds of origins), and with the slot counter exhausted nothing refills the other lanes -- the longest slot would
    // set the kernel time.  The origins of one slot are independent draws at known offsets of its LCG stream (lcg_skip), so the
    // warp splits into groups of G lanes per live slot and walks G consecutive origins of that slot at once; the lowest origin
    // that meets a leaf is the one the sequential loop would have stopped at (same draws, same walk -> same item).
    for (;;) {
        const uint32_t live_mask = __ballot_sync(0xFFFFFFFFu, have_slot);
        const uint32_t n_live    = (uint32_t)__popc(live_mask);
        if (n_live == 0u)
            break;
        const uint32_t G     = n_live == 1u ? 32u : n_live == 2u ? 16u : n_live <= 4u ? 8u : 4u; // lanes per slot (n_live <= 8)
        const uint32_t group = lane / G, g = lane % G;
        const bool has       = group < n_live;
        const uint32_t owner = has ? (uint32_t)__fns(live_mask, 0u, group + 1u) : 0u;
        // the owner's in-flight origin (if any) is redone: origin index of the group's first lane
        const uint32_t first   = n_origins - ((have_slot && walking) ? 1u : 0u);
        const uint32_t o_idx   = __shfl_sync(0xFFFFFFFFu, idx, owner);
        const uint32_t o_rng0  = __shfl_sync(0xFFFFFFFFu, rng0, owner);
        const uint32_t o_first = __shfl_sync(0xFFFFFFFFu, first, owner);
        const bool o_dir_ok    = __shfl_sync(0xFFFFFFFFu, dir_ok ? 1u : 0u, owner) != 0u;
        Ray cray;
        cray.dx = __shfl_sync(0xFFFFFFFFu, ray.dx, owner), cray.dy = __shfl_sync(0xFFFFFFFFu, ray.dy, owner);
        cray.dz = __shfl_sync(0xFFFFFFFFu, ray.dz, owner);
        cray.ix = __shfl_sync(0xFFFFFFFFu, ray.ix, owner), cray.iy = __shfl_sync(0xFFFFFFFFu, ray.iy, owner);
        cray.iz = __shfl_sync(0xFFFFFFFFu, ray.iz, owner);
        cray.fast = 0u;
        cray.ox = cray.oy = cray.oz = 0.f;
        NbvhConst cbc    = p.nbvh;
        cbc.leaf_lod_max = __shfl_sync(0xFFFFFFFFu, bc.leaf_lod_max, owner);
        cbc.dilation     = __shfl_sync(0xFFFFFFFFu, bc.dilation, owner);
        uint32_t crng    = lcg_skip(o_rng0, 3u * (o_first + g));
        bool cwalking = false, cfound = false;
        CutWalk &cw = w; // the owner's own walk state is dead: its in-flight origin is redone by the group
        if (has) {
            draw_origin(p, off, crng, cray);
            set_walk_division_mode(cray, o_dir_ok, cut_safe);
            cut_walk_start(cbc, cw);
            cwalking = true;
        }
        while (__any_sync(0xFFFFFFFFu, cwalking)) {
            if (cwalking) {
                const WalkResult r = cut_walk_step(cbc, cray, cw, cnt_nbvh);
                cfound   = r == WALK_FOUND;
                cwalking = r == WALK_CONTINUE;
            }
        }
        // the lowest origin of every group that met a leaf wins
        const uint32_t found_mask = __ballot_sync(0xFFFFFFFFu, cfound);
        const uint32_t gmask      = G == 32u ? 0xFFFFFFFFu : (((1u << G) - 1u) << (group * G));
        const uint32_t gfound     = found_mask & gmask;
        if (cfound && ((uint32_t)__ffs(gfound) - 1u == lane))
            write_item(p, o_idx, cw, crng, cray);
        // owners: slot done, or G more origins consumed
        if (have_slot) {
            const uint32_t my_group = (uint32_t)__popc(live_mask & ((1u << lane) - 1u));
            const uint32_t mine     = G == 32u ? found_mask : ((found_mask >> (my_group * G)) & ((1u << G) - 1u));
            if (mine) {
                have_slot = false;
            } else {
                n_origins = first + G;
                walking   = false;
            }
        }
    }
    warp_add_counter(p.counters, CNT_NBVH_STEPS, cnt_nbvh);
}

// One thread per slot of one step (items of that step: p.items + step_first_item).  The roulette's survivors are compacted inside
// the block, so the trace against the neural mesh's BLAS runs on dense warps.
__global__ void __launch_bounds__(128)
k_sample_targets(const __grid_constant__ SamplerParams p, uint32_t step_first_item)
{
    __shared__ uint32_t s_src[128];
    __shared__ uint32_t s_warp_count[4];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t i   = blockIdx.x * blockDim.x + tid;
    bool keep = false;
    if (i < (uint32_t)p.sub_batch_size) {
        const uint4 a = reinterpret_cast<const uint4 *>(p.items)[3u * (size_t)(step_first_item + i)];
        uint32_t rng  = a.y;
        p.sampled_node_indices[i] = a.x;
        // rejected: the slot keeps its previous contents (reference quirk)
        keep = !p.training_node_survival || !(__ldg(p.node_keep + a.x) < lcg_randomf(rng));
    }
    const uint32_t keep_mask = __ballot_sync(0xFFFFFFFFu, keep);
    if (lane == 0u)
        s_warp_count[warp] = (uint32_t)__popc(keep_mask);
    __syncthreads();
    uint32_t before = 0u, n_keep = 0u;
#pragma unroll
    for (uint32_t k = 0; k < 4u; ++k) {
        before += k < warp ? s_warp_count[k] : 0u;
        n_keep += s_warp_count[k];
    }
    if (keep)
        s_src[before + (uint32_t)__popc(keep_mask & ((1u << lane) - 1u))] = i;
    __syncthreads();
    uint32_t cnt_blas = 0, cnt_tri = 0, cnt_written = 0;
    if (tid < n_keep) {
        const uint32_t slot = s_src[tid];
        const uint4 *it     = reinterpret_cast<const uint4 *>(p.items) + 3u * (size_t)(step_first_item + slot);
        const uint4 a = it[0], b = it[1], c = it[2];
        Ray ray;
        ray.ox = __uint_as_float(a.z), ray.oy = __uint_as_float(a.w), ray.oz = __uint_as_float(b.x);
        ray.dx = __uint_as_float(b.y), ray.dy = __uint_as_float(b.z), ray.dz = __uint_as_float(b.w);
        ray_prepare_guarded(ray);
        uint32_t rng = a.y;
        if (p.training_node_survival)
            lcg_randomf(rng); // the roulette's draw (above) precedes the segment points'
        ItsData its;
        emit_slot(p, slot, ray, __uint_as_float(c.x), __uint_as_float(c.y), false, a.x, rng, its, cnt_blas, cnt_tri);
        ++cnt_written;
    }
    warp_add_counter(p.counters, CNT_BLAS_STEPS, cnt_blas);
    warp_add_counter(p.counters, CNT_TRI_TESTS, cnt_tri);
    warp_add_counter(p.counters, CNT_RAYS, cnt_written);
}

// Walks the slots of p.n_steps consecutive optimisation steps into p.items (n_steps * batch items of 48 bytes).
cudaError_t launch_sample_walk(const SamplerParams &p, uint32_t *sampler_counters, int num_sms, cudaStream_t stream)
{
    // counter block: [0] slot cursor, [4] "a node failed k_check_dilation_safe"
    cudaError_t e = cudaMemsetAsync(sampler_counters, 0, 8 * sizeof(uint32_t), stream);
    if (e != cudaSuccess)
        return e;
    SamplerParams q   = p;
    q.dilation_unsafe = sampler_counters + 4;
    k_check_dilation_safe<<<(q.nbvh.num_nodes + 255u) / 256u, 256, 0, stream>>>(q.nbvh.nodes, q.nbvh.num_nodes, q.dilation_unsafe);
    int per_sm = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sample_walk, 128, 0)) != cudaSuccess)
        return e;
    const uint64_t total = (uint64_t)q.sub_batch_size * q.n_steps;
    const uint32_t want  = (uint32_t)((total + 127u) / 128u);
    uint32_t grid        = (uint32_t)(num_sms * (per_sm > 0 ? per_sm : 1));
    if (grid > want)
        grid = want;
    k_sample_walk<<<grid, 128, 0, stream>>>(q, sampler_counters);
    return cudaGetLastError();
}

// Roulette + targets of ONE step from its walked items (first item of the step: step_first_item).
cudaError_t launch_sample_targets(const SamplerParams &p, uint32_t step_first_item, cudaStream_t stream)
{
    if (p.training_node_survival)
        k_node_keep<<<(p.nbvh.num_nodes + 255u) / 256u, 256, 0, stream>>>(p, p.nbvh.num_nodes);
    k_sample_targets<<<((uint32_t)p.sub_batch_size + 127u) / 128u, 128, 0, stream>>>(p, step_first_item);
    return cudaGetLastError();
}

cudaError_t launch_generate_batch(const SamplerParams &p, int num_sms, cudaStream_t stream)
{
    if (p.sub_batch_size <= 0)
        return cudaSuccess;
    k_generate_batch<<<((uint32_t)p.sub_batch_size + 127u) / 128u, 128, 0, stream>>>(p);
    return cudaGetLastError();
}

} // namespace nbvh

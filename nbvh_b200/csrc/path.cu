// path.cu -- hybrid neural path tracing for max_depth > 1: replaces the host loop
// NeuralPathModule::neural_path_tracing (/root/reference/src/cuda/neural/path/path_module.cu:139-245) and the kernels
// hybrid_camera_raygen_and_hybrid_first_hit_and_traversal (K1), non_neural_bvh_intersection (K8) and
// hybrid_scatter_and_get_next_tlas_traversal_step (K9: shading, BSDF sampling, Russian roulette, re-queue)
// (/root/reference/src/cuda/neural/path/path_inference_kernels.cu:195-429).
//
// Per depth d = 1 .. max_depth:   k_path_trace  ->  k_neural<path>  ->  k_path_shade
//
//   k_path_trace    one thread per path of the current list (depth 1: per pixel, camera rays): whole TLAS walk with every
//                   classical BLAS closest hit, N-BVH walk to the first leaf interval, ballot-compacted append to the neural
//                   ray queue -- exactly what k_primary does.  Rays that need no network query leave their hit in path_hits.
//   k_neural<true>  the persistent tcgen05 tile kernel of render.cu; queue entries name list indices, rays come from the path
//                   state, the closest hit goes to path_hits.
//   k_path_shade    one thread per path: background for escaped rays, last-bounce splat, intersection data and material of
//                   classical hits (texture fetches), the network's normal/albedo with the module's constant
//                   metallic/roughness/specular for neural hits, emission, Disney BSDF sampling, throughput update, Russian
//                   roulette, and the bounce ray.  Survivors are COMPACTED into the next path list (one atomic per warp), so
//                   every later kernel runs on dense warps although most paths die early (Chess: 42 % of the camera rays
//                   escape, and only 25 % of the pixels still carry a path at depth 3).
//
// No host synchronisation: 3 * max_depth launches per frame, list lengths stay on the device (the reference: ~180 launches
// and ~30 host round trips per bounce).  The reference interleaves BLAS visits in TLAS order over several wavefront iterations;
// the closest hit is the minimum over the same candidate set, so the paths are the same (see render.cu's header).
// Random numbers are drawn in the reference's order (2 camera jitter, 3 per BSDF sample, 1 per roulette), so every pixel
// follows the same path as in the reference up to floating-point rounding of the shading arithmetic.
#include "render_common.cuh"
#include "path.cuh"

namespace nbvh {

__device__ __forceinline__ void path_splat(const RenderParams &p, uint32_t pixel, V3 c) { splat(p, pixel, c.x, c.y, c.z, 1.0f); }

// K9 after the TLAS walk of the current ray is complete (path_inference_kernels.cu:306-428).  Returns true when the path
// continues with `ray` at depth + 1; otherwise the pixel has been splatted.
__device__ __forceinline__ bool shade_and_scatter(const RenderParams &p, const PathConst &pc, uint32_t pixel, Ray &ray, uint32_t depth,
                                                  const PathHit &ph, V3 &throughput, V3 &illumination, uint32_t &rng)
{
    const float best_t       = ph.h0.x;
    const uint32_t valid_hit = __float_as_uint(ph.h0.y);
    if (valid_hit == 0u || depth == p.max_depth) {
        if (valid_hit == 0u)
            illumination = add3(illumination, mul3(throughput, background_eval(p.bg, ray.dx, ray.dy, ray.dz)));
        path_splat(p, pixel, illumination);
        return false;
    }
    Bsdf bsdf   = default_bsdf();
    const V3 wo = v3(-ray.dx, -ray.dy, -ray.dz);
    V3 its_normal, hit_pos;
    if (valid_hit == 1u) { // NEURAL_TRAVERSAL_HIT: the network's normal and albedo, constant specular parameters
        bsdf.base_color = v3(ph.h1.y, ph.h1.z, ph.h1.w);
        bsdf.metallic = pc.neural_metallic, bsdf.roughness = pc.neural_roughness, bsdf.specular = pc.neural_specular;
        its_normal = v3(ph.h0.z, ph.h0.w, ph.h1.x);
        hit_pos    = v3(ray.ox + best_t * ray.dx, ray.oy + best_t * ray.dy, ray.oz + best_t * ray.dz);
    } else {
        TriHit th;
        th.t = best_t, th.u = ph.h0.z, th.v = ph.h0.w, th.prim = __float_as_uint(ph.h1.z);
        const uint32_t mesh = __float_as_uint(ph.h1.y);
        const ItsData its   = compute_its_data(scene_blas(p.scene, mesh), path_attr(pc, mesh), pc.materials, pc.textures, th);
        hit_pos = its.position, its_normal = its.sh_normal;
        const bool has_material = its.material_id != -1 && pc.materials;
        if (p.ignore_neural_blas && mesh == pc.neural_bvh_idx) {
            if (has_material)
                bsdf.base_color = color_param(pc.materials[its.material_id].base_color, its.u, its.v, pc.textures);
            bsdf.metallic = pc.neural_metallic, bsdf.roughness = pc.neural_roughness, bsdf.specular = pc.neural_specular;
        } else if (pc.force_constant_material) {
            if (has_material)
                bsdf.base_color = color_param(pc.materials[its.material_id].base_color, its.u, its.v, pc.textures);
            bsdf.metallic = pc.constant_metallic, bsdf.roughness = pc.constant_roughness, bsdf.specular = pc.constant_specular;
        } else if (has_material) {
            bsdf = unpack_material(pc.materials[its.material_id], its.u, its.v, pc.textures);
        }
    }
    V3 sh_normal, tangent, bitangent;
    shading_frame(wo, its_normal, sh_normal, tangent, bitangent);
    if ((depth == 1u || pc.emission_eval) && !is_zero3(bsdf.base_emission))
        illumination = add3(illumination, mul3(throughput, scale3(bsdf.base_emission, bsdf.emission_scale)));
    float pdf;
    V3 wi;
    const V3 f      = disney_sample_eval(bsdf, sh_normal, wo, tangent, bitangent, rng, wi, pdf);
    const float cti = fabsf(dot3(wi, sh_normal));
    throughput      = v3(throughput.x * (f.x * cti / pdf), throughput.y * (f.y * cti / pdf), throughput.z * (f.z * cti / pdf));
    if (pdf == 0.f || is_zero3(f)) {
        path_splat(p, pixel, illumination);
        return false;
    }
    if (pc.apply_russian_roulette && rr_terminate(depth, throughput, pc.rr_start_depth, rng)) {
        path_splat(p, pixel, illumination);
        return false;
    }
    ray.ox = hit_pos.x + sh_normal.x * pc.normal_epsilon_offset;
    ray.oy = hit_pos.y + sh_normal.y * pc.normal_epsilon_offset;
    ray.oz = hit_pos.z + sh_normal.z * pc.normal_epsilon_offset;
    ray.dx = wi.x, ray.dy = wi.y, ray.dz = wi.z;
    ray_prepare(ray, p.cam.nodes_safe != 0u);
    return true;
}

// ------------------------------------------------------------------------------------------------ k_path_shade
// One thread per path of the current (compacted) list: shades the closest hit, splats finished paths, and appends the
// survivors -- with their bounce ray -- to the next list (warp ballot/prefix-sum compaction, one atomic per warp), so
// that the following trace and network kernels run on dense warps.
#ifndef NBVH_SHADE_THREADS
#define NBVH_SHADE_THREADS 256
#endif
__global__ void __launch_bounds__(NBVH_SHADE_THREADS)
k_path_shade(const __grid_constant__ RenderParams p, const __grid_constant__ PathConst pc, const __grid_constant__ PathIo io)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t i    = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_in = io.count_in ? *io.count_in : p.pixel_count * (p.spp_count > 1u ? p.spp_count : 1u);
    if (i - lane >= n_in)
        return; // whole warp beyond the list
    bool alive = false;
    PathState st;
    if (i < n_in) {
        st = io.in[i];
        const PathHit ph = io.hits_in[i];
        uint32_t rng = st.s3.x, depth = st.s3.y;
        const uint32_t pixel = st.s3.z;
        Ray ray;
        ray.ox = st.s0.x, ray.oy = st.s0.y, ray.oz = st.s0.z, ray.dx = st.s0.w, ray.dy = st.s1.x, ray.dz = st.s1.y;
        V3 throughput = v3(st.s1.z, st.s1.w, st.s2.x), illumination = v3(st.s2.y, st.s2.z, st.s2.w);
        alive = shade_and_scatter(p, pc, pixel, ray, depth, ph, throughput, illumination, rng);
        if (alive) {
            st.s0 = make_float4(ray.ox, ray.oy, ray.oz, ray.dx);
            st.s1 = make_float4(ray.dy, ray.dz, throughput.x, throughput.y);
            st.s2 = make_float4(throughput.z, illumination.x, illumination.y, illumination.z);
            st.s3 = make_uint4(rng, depth + 1u, pixel, 0u);
        }
    }
    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, alive);
    if (mask) {
        uint32_t base = 0;
        if (lane == (uint32_t)(__ffs(mask) - 1))
            base = atomicAdd(io.count_out, (uint32_t)__popc(mask));
        base = __shfl_sync(0xFFFFFFFFu, base, __ffs(mask) - 1);
        if (alive)
            io.out[base + __popc(mask & ((1u << lane) - 1u))] = st;
    }
}

// ------------------------------------------------------------------------------------------------ k_path_trace
// One thread per path of the list (kFirst: per pixel, camera rays): TLAS walk + classical BLAS closest hits + first
// N-BVH interval, exactly like k_primary; rays with a pending interval go to the neural ray queue (entry id = list index),
// all others leave their closest hit in path_hits.
#ifndef NBVH_TRACE_THREADS
#define NBVH_TRACE_THREADS 64 // measured on B200 (gpurun s71), config 2b frame: 256 / 128 / 64 threads -> 2.406 / 2.343 / 2.325 ms (72 registers: 24 -> 28 resident warps)
#endif
template <bool kFirst>
__global__ void __launch_bounds__(NBVH_TRACE_THREADS)
k_path_trace(const __grid_constant__ RenderParams p, const __grid_constant__ PathIo io)
{
    const uint32_t lane  = threadIdx.x & 31u;
    const uint32_t chunk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t sidx  = chunk * 32u + lane;
    const uint32_t n_spp = p.spp_count > 1u ? p.spp_count : 1u;
    const uint32_t n_in  = kFirst ? p.pixel_count * n_spp : *io.count_in;
    if (chunk * 32u >= n_in)
        return;
    const bool active = sidx < n_in;
    uint32_t cnt_blas = 0, cnt_tri = 0, cnt_nbvh = 0;
    uint32_t pixel = 0;
    Ray ray;
    if (active) {
        if constexpr (kFirst) {
            uint32_t rng;
            // multi-sample launch: slot = sample * pixel_count + pixel slot (pixel_count is a multiple of 32, so a warp stays in
            // one sample and on one 8x4-pixel block)
            const uint32_t smp = sidx / p.pixel_count, local = sidx - smp * p.pixel_count;
            pixel = swizzle_pixel(p, p.first_pixel + local);
            ray   = camera_ray(p.cam, pixel, p.sample_idx + (int)smp, p.sample_offset, rng);
            PathState st;
            st.s0 = make_float4(ray.ox, ray.oy, ray.oz, ray.dx);
            st.s1 = make_float4(ray.dy, ray.dz, 1.f, 1.f);
            st.s2 = make_float4(1.f, 0.f, 0.f, 0.f);
            st.s3 = make_uint4(rng, 1u, pixel, 0u);
            p.paths[sidx] = st;
        } else {
            const float4 s0 = p.paths[sidx].s0, s1 = p.paths[sidx].s1;
            ray.ox = s0.x, ray.oy = s0.y, ray.oz = s0.z, ray.dx = s0.w, ray.dy = s1.x, ray.dz = s1.y;
            ray_prepare(ray, p.cam.nodes_safe != 0u);
        }
    }
    bool neural = false;
    uint32_t nstack[kNeuralStackSize];
    int nss        = 0;
    float seg_mint = 0.f, seg_maxt = 0.f;
    HitState h;
    h.best_t = FLT_MAX, h.valid_hit = 0u;
    h.n_or_uv[0] = h.n_or_uv[1] = h.n_or_uv[2] = 0.f;
    h.alb_or_ids[0] = h.alb_or_ids[1] = h.alb_or_ids[2] = 0.f;
    if (active) {
        uint32_t tstack[kTlasStackSize];
        int tss   = 1;
        tstack[0] = 0u;
        while (tss > 0) {
            uint32_t blas_idx;
            bool is_neural;
            if (!tlas_next_blas(p.scene.tlas, ray, tstack, tss, blas_idx, is_neural))
                break;
            if (is_neural && !p.ignore_neural_blas) {
                neural = true;
            } else if (blas_idx < p.scene.n_blas) {
                const TriHit th = blas_closest_hit(scene_blas(p.scene, blas_idx), ray, 1e-4f, h.best_t, cnt_blas, cnt_tri);
                if (th.t < h.best_t) {
                    h.best_t        = th.t;
                    h.valid_hit     = 2u;
                    h.n_or_uv[0]    = th.u, h.n_or_uv[1] = th.v, h.n_or_uv[2] = 0.f;
                    h.alb_or_ids[0] = __uint_as_float(blas_idx), h.alb_or_ids[1] = __uint_as_float(th.prim);
                    h.alb_or_ids[2] = 0.f;
                }
            }
        }
        if (neural) {
            uint32_t node;
            nstack[0] = 0u, nss = 1;
            neural = nbvh_next_interval(p.nbvh, ray, nstack, nss, h.best_t, seg_mint, seg_maxt, node, cnt_nbvh);
        }
    }
    uint32_t cls = 0u;
    if constexpr (kFirst) {
        if (p.class_lists && active && neural)
            cls = chain_class(p, pixel, nss);
    }
    const uint32_t qmask = __ballot_sync(0xFFFFFFFFu, active && neural);
    if (qmask) {
        uint32_t base = 0;
        if (lane == (uint32_t)(__ffs(qmask) - 1))
            base = atomicAdd(p.queue_count, (uint32_t)__popc(qmask));
        base = __shfl_sync(0xFFFFFFFFu, base, __ffs(qmask) - 1);
        const uint32_t slot = base + __popc(qmask & ((1u << lane) - 1u));
        if constexpr (kFirst) {
            if (p.class_lists)
                file_in_class(p, qmask, cls, slot, lane);
        }
        if (active && neural) {
            NeuralQueueEntry *q = p.queue + slot;
            q->a = make_uint4(sidx, __float_as_uint(h.best_t), h.valid_hit, __float_as_uint(h.n_or_uv[0]));
            q->b = make_uint4(__float_as_uint(h.n_or_uv[1]), __float_as_uint(h.alb_or_ids[0]), __float_as_uint(h.alb_or_ids[1]), (uint32_t)nss);
            q->c = make_uint4(__float_as_uint(seg_mint), __float_as_uint(seg_maxt), 0u, 0u);
            for (int i = 0; i < nss; i += 4)
                q->st[i >> 2] = make_uint4(nstack[i], nstack[i + 1], nstack[i + 2], nstack[i + 3]);
        }
    }
    if (active && !neural) {
        PathHit ph;
        ph.h0 = make_float4(h.best_t, __uint_as_float(h.valid_hit), h.n_or_uv[0], h.n_or_uv[1]);
        ph.h1 = make_float4(h.n_or_uv[2], h.alb_or_ids[0], h.alb_or_ids[1], h.alb_or_ids[2]);
        p.path_hits[sidx] = ph;
        if (kFirst && p.hits) { // primary-hit records of rays that never reach the network
            nbvh_hit_record r;
            r.t = h.best_t, r.valid_hit = h.valid_hit;
            r.normal_or_uv[0] = h.n_or_uv[0], r.normal_or_uv[1] = h.n_or_uv[1], r.normal_or_uv[2] = h.n_or_uv[2];
            r.albedo_or_ids[0] = h.alb_or_ids[0], r.albedo_or_ids[1] = h.alb_or_ids[1], r.albedo_or_ids[2] = h.alb_or_ids[2];
            p.hits[pixel] = r;
        }
    }
    warp_add_counter(p.counters, CNT_RAYS, active ? 1u : 0u);
    warp_add_counter(p.counters, CNT_BLAS_STEPS, cnt_blas);
    warp_add_counter(p.counters, CNT_TRI_TESTS, cnt_tri);
    warp_add_counter(p.counters, CNT_NBVH_STEPS, cnt_nbvh);
    warp_add_counter(p.counters, CNT_QUEUED, active && neural ? 1u : 0u);
}

// ------------------------------------------------------------------------------------------------ ray batches
// Closest-hit query of an arbitrary ray batch (BASELINE config 5, and the general "neural ray query" entry point):
// rays -> path list -> k_path_trace -> k_neural<path> -> hit records.  Same kernels as one depth of the path tracer.
__global__ void __launch_bounds__(256)
k_rays_to_paths(const float *__restrict__ o, const float *__restrict__ d, uint32_t n, uint32_t ray_depth, PathState *paths, uint32_t *counters)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    if (i == 0u)
        counters[2] = n; // list length read by k_path_trace
    PathState st;
    st.s0 = make_float4(o[3u * i], o[3u * i + 1u], o[3u * i + 2u], d[3u * i]);
    st.s1 = make_float4(d[3u * i + 1u], d[3u * i + 2u], 1.f, 1.f);
    st.s2 = make_float4(1.f, 0.f, 0.f, 0.f);
    st.s3 = make_uint4(0u, ray_depth, i, 0u);
    paths[i] = st;
}

__global__ void __launch_bounds__(256)
k_hits_to_records(const PathHit *__restrict__ hits, uint32_t n, nbvh_hit_record *out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const PathHit ph = hits[i];
    nbvh_hit_record r;
    r.t = ph.h0.x, r.valid_hit = __float_as_uint(ph.h0.y);
    r.normal_or_uv[0] = ph.h0.z, r.normal_or_uv[1] = ph.h0.w, r.normal_or_uv[2] = ph.h1.x;
    r.albedo_or_ids[0] = ph.h1.y, r.albedo_or_ids[1] = ph.h1.z, r.albedo_or_ids[2] = ph.h1.w;
    out[i] = r;
}

// counters: [0] queue count, [1] queue cursor (zeroed here), [2] list length (written by k_rays_to_paths)
cudaError_t launch_trace_rays(const RenderParams &rp_in, const NetConst &nc, const float *o, const float *d, uint32_t n, uint32_t ray_depth,
                              PathState *paths, PathHit *hits, uint32_t *counters, nbvh_hit_record *out, int num_sms, cudaStream_t stream)
{
    if (n == 0)
        return cudaSuccess;
    RenderParams rp = rp_in;
    rp.pixel_count  = n;
    rp.first_pixel  = 0;
    rp.chunk_stride = 1, rp.chunk_offset = 0;
    rp.queue_count = counters, rp.queue_cursor = counters + 1;
    rp.class_lists = nullptr, rp.class_count = nullptr, rp.chain_hist = nullptr; // arrival order (no per-pixel history of a ray batch)
    rp.ray_depth = ray_depth;
    rp.paths = paths, rp.path_hits = hits;
    rp.hits = nullptr;
    PathIo io;
    io.in = paths, io.hits_in = hits, io.out = nullptr, io.count_in = counters + 2, io.count_out = nullptr;
    cudaError_t e;
    if ((e = cudaMemsetAsync(counters, 0, 2 * sizeof(uint32_t), stream)) != cudaSuccess)
        return e;
    const uint32_t blocks = (n + 255u) / 256u;
    k_rays_to_paths<<<blocks, 256, 0, stream>>>(o, d, n, ray_depth, paths, counters);
    k_path_trace<false><<<(n + NBVH_TRACE_THREADS - 1u) / NBVH_TRACE_THREADS, NBVH_TRACE_THREADS, 0, stream>>>(rp, io);
    if ((e = cudaGetLastError()) != cudaSuccess)
        return e;
    if ((e = launch_neural(rp, nc, num_sms, true, stream)) != cudaSuccess)
        return e;
    k_hits_to_records<<<blocks, 256, 0, stream>>>(hits, n, out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ image error metrics (f4)
// compute_and_display_render_error (neural_bvh_renderer_kernels.cu:91-160) on linear memory: per-pixel error of accum / spp
// against a reference RGBA32F image, block-reduced sum of the channel-mean error (the reference: one atomicAdd per pixel),
// optional sRGB8 error image.  mode: 3 AbsError, 4 RelAbsError, 5 MSEError, 6 RelMSEError (PostprocessMode).
__global__ void __launch_bounds__(256)
k_image_error(const float4 *__restrict__ accum, const float4 *__restrict__ reference, uint32_t n, float inv_spp, int mode, float *sum, uchar4 *ldr)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float e = 0.f;
    if (i < n) {
        const float4 a = accum[i], r = reference[i];
        const float c[3] = {a.x * inv_spp, a.y * inv_spp, a.z * inv_spp}, ref[3] = {r.x, r.y, r.z};
        float err[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float d = ref[k] - c[k];
            err[k] = mode == 3 ? fabsf(d) : mode == 4 ? fabsf(d) / (ref[k] + 0.1f) : mode == 5 ? d * d : d * d / (ref[k] + 0.1f);
        }
        e = (err[0] + err[1] + err[2]) / 3.f;
        if (ldr) {
            uchar4 px;
            px.x = (unsigned char)min(max(int(linear_to_srgb(err[0]) * 256.0f), 0), 255);
            px.y = (unsigned char)min(max(int(linear_to_srgb(err[1]) * 256.0f), 0), 255);
            px.z = (unsigned char)min(max(int(linear_to_srgb(err[2]) * 256.0f), 0), 255);
            px.w = 255;
            ldr[i] = px;
        }
    }
    __shared__ float s_part[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        e += __shfl_xor_sync(0xFFFFFFFFu, e, o);
    if ((threadIdx.x & 31u) == 0u)
        s_part[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w)
            t += s_part[w];
        atomicAdd(sum, t);
    }
}

cudaError_t launch_image_error(const float4 *accum, const float4 *reference, uint32_t n, uint32_t spp, int mode, float *sum_dev, uchar4 *ldr,
                               cudaStream_t stream)
{
    cudaError_t e = cudaMemsetAsync(sum_dev, 0, sizeof(float), stream);
    if (e != cudaSuccess || n == 0)
        return e;
    k_image_error<<<(n + 255u) / 256u, 256, 0, stream>>>(accum, reference, n, 1.f / (float)(spp ? spp : 1u), mode, sum_dev, ldr);
    return cudaGetLastError();
}

// One frame of NeuralPathModule::neural_path_tracing.  `counters`: kPathCountersPerDepth zeroed words per depth
// (neural queue count, queue cursor, length of the path list).  paths[2] / hits[2]: ping-pong path lists.
cudaError_t launch_render_path(const RenderParams &rp_in, const PathConst &pc, const NetConst &nc, const NbvhConst &nbvh_bounce,
                               PathState *const paths[2], PathHit *const hits[2], uint32_t *counters, int num_sms, cudaStream_t stream)
{
    if (rp_in.pixel_count == 0)
        return cudaSuccess;
    RenderParams rp       = rp_in;
    const uint32_t slots  = rp.pixel_count * (rp.spp_count > 1u ? rp.spp_count : 1u);
    const uint32_t blocks = ((slots + 31u) / 32u + 7u) / 8u;
    const uint32_t tblocks = (slots + NBVH_TRACE_THREADS - 1u) / NBVH_TRACE_THREADS;
    cudaError_t e;
    if (rp.spp_count > 1u && rp.sample_idx == 0) // sample 0 overwrites the accumulation (accumulate_pixel_value); here: clear, then reduce
        k_clear_accum<<<(rp.pixel_count + 255u) / 256u, 256, 0, stream>>>(rp);
    for (uint32_t depth = 1; depth <= rp.max_depth; ++depth) {
        uint32_t *c     = counters + kPathCountersPerDepth * (depth - 1u);
        const int cur   = (int)((depth - 1u) & 1u);
        rp.queue_count  = c;
        rp.queue_cursor = c + 1;
        rp.ray_depth    = depth;
        rp.paths = paths[cur], rp.path_hits = hits[cur];
        PathIo io;
        io.in = paths[cur], io.hits_in = hits[cur], io.out = paths[cur ^ 1];
        io.count_in  = depth == 1u ? nullptr : c + 2; // list length written by the previous shade pass
        io.count_out = c + kPathCountersPerDepth + 2;
        if (depth == 1u) {
            k_path_trace<true><<<tblocks, NBVH_TRACE_THREADS, 0, stream>>>(rp, io);
        } else {
            rp.class_lists = nullptr, rp.class_count = nullptr, rp.chain_hist = nullptr; // bounce rays: arrival order
            rp.nbvh = nbvh_bounce; // path_lod_switch: bounce rays may walk a coarser LoD (neural_bvh_module_common.cuh:333-335)
            k_path_trace<false><<<tblocks, NBVH_TRACE_THREADS, 0, stream>>>(rp, io);
        }
        if ((e = cudaGetLastError()) != cudaSuccess)
            return e;
        if ((e = launch_neural(rp, nc, num_sms, true, stream)) != cudaSuccess)
            return e;
        k_path_shade<<<(slots + NBVH_SHADE_THREADS - 1u) / NBVH_SHADE_THREADS, NBVH_SHADE_THREADS, 0, stream>>>(rp, pc, io);
        if ((e = cudaGetLastError()) != cudaSuccess)
            return e;
    }
    if (rp.spp_count > 1u && rp.ldr) {
        k_resolve_ldr<<<(rp.pixel_count + 255u) / 256u, 256, 0, stream>>>(rp, 1.f / (float)(rp.sample_idx + (int)rp.spp_count));
        if ((e = cudaGetLastError()) != cudaSuccess)
            return e;
    }
    return cudaSuccess;
}

} // namespace nbvh

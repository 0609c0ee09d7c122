// render.cu -- the frame pipeline of the neural ray-query path as two kernels (replaces the host loops of
// /root/reference/src/cuda/neural/path/path_module.cu:54-245 and kernels K1-K9 at max_depth 1):
//
//   k_primary   one thread per pixel, no tensor-core state, full occupancy: camera ray (a6), complete TLAS walk (a4) with the
//               classical BLAS closest hits (a14) taken on the way, then -- for rays whose TLAS walk reached the neural BLAS --
//               the N-BVH walk from the root to the FIRST leaf interval (a3; coherent primary rays, so the long root walk
//               runs at high SIMT efficiency here).  Rays with a pending interval are appended, stack included, to the
//               neural ray queue with a warp ballot/prefix-sum compaction (one atomic per warp); all others are finished
//               (background, splat, hit record).
//   k_neural    persistent tile kernel (tcgen05/TMEM): lanes pull queue entries (ray regenerated from the pixel id: same
//               arithmetic -> same bits), every 128-lane tile runs hash-grid encode + MLP + decode (a8-a11) on the pending
//               intervals, then each lane resumes its walk with ray.maxt = current closest t (a12) -- a short pop/descend --
//               and finished lanes splat and refill from the queue (a13).
//
// Order of BLAS visits: the reference visits the BLAS of a ray in TLAS front-to-back order; here every classical BLAS is
// intersected before the neural walk starts, which only *prunes* neural leaves behind the classical hit.  The final closest
// hit is the minimum over the same candidate set, so hit records are identical; the number of network queries can only shrink.
//
// Compiled with -fmad=false: see device_core.cuh for the arithmetic contract.
#include "render_common.cuh"
#include <cstdio>

namespace nbvh {

#ifndef NBVH_NEURAL_TILES
#define NBVH_NEURAL_TILES 2
#endif
#ifndef NBVH_NEURAL_CTAS_PER_SM
#define NBVH_NEURAL_CTAS_PER_SM 2
#endif
#ifndef NBVH_WALK_ROUNDS
#define NBVH_WALK_ROUNDS 2
#endif
constexpr int kNeuralTiles   = NBVH_NEURAL_TILES;        // 128-lane tiles per CTA
constexpr int kNeuralThreads = kNeuralTiles * kTileRows;
constexpr int kNeuralCtasPerSm = NBVH_NEURAL_CTAS_PER_SM;
#ifndef NBVH_PHASE_TIMING
#define NBVH_PHASE_TIMING 0 // development builds: per-phase SM-cycle sums of k_neural (tools/perf_r2.py prints them)
#endif
#ifndef NBVH_PHASE_SPARSE
#define NBVH_PHASE_SPARSE 0
#endif
#ifndef NBVH_LONG_STACK
#define NBVH_LONG_STACK 4
#endif
#ifndef NBVH_CHAIN_STATS
#define NBVH_CHAIN_STATS 0 // development builds: histogram of (predicted chain, actual chain) of every queued ray, printed per launch
#endif
#ifndef NBVH_WALK_BUDGET
#define NBVH_WALK_BUDGET 0 // > 0: node fetches a lane may spend per walk round before it parks its walk (0 = unbounded).  Measured on
                           // B200 (gpurun s31): 4 fetches 0.606 ms, 8 fetches 0.507 ms, unbounded 0.506 ms per 1080p launch -- parked
                           // rays pay a whole extra network pass; kept as a switch only
#endif
#ifndef NBVH_TILE_COOP
#define NBVH_TILE_COOP 1 // 1: the hash-grid gathers of a tile's active rows are dealt to all 128 lanes of the tile (0: per warp).
                         // Measured (s31, 1080p launch): per warp 0.506 ms, per tile 0.481 ms, per tile chunk-major 0.465 ms
#endif
#ifndef NBVH_NEURAL_SMEM_STACK
#define NBVH_NEURAL_SMEM_STACK 0
#endif
// k_neural keeps the per-lane N-BVH walk stacks in shared memory (entry i of thread t at word i * kNeuralThreads + t: conflict-free
// whatever the lanes' stack depths).  As a thread-local array the stack is dynamically indexed local memory, one 128-byte line per
// stack level and warp, and the hash-grid gathers of every tile iteration evict those lines from L1: each pop of the resumed walk --
// also each entry skipped as behind the closest hit -- then costs a round trip to L2 on the critical path of the ray's query chain.
constexpr uint32_t kNeuralStackBytes = NBVH_NEURAL_SMEM_STACK ? (uint32_t)kNeuralThreads * (uint32_t)kNeuralStackSize * 4u : 0u;
struct LaneStack {
    uint32_t *base;
    __device__ __forceinline__ uint32_t &operator[](int i) const { return base[i * kNeuralThreads]; }
};
__host__ __device__ inline uint32_t neural_stack_offset(uint32_t wimg_bytes, uint32_t in_k)
{
    return (smem_layout(wimg_bytes, in_k, kNeuralTiles).total + 127u) & ~127u;
}
constexpr int kWalkRounds    = NBVH_WALK_ROUNDS;         // refill/advance rounds before a tile runs the network

// ------------------------------------------------------------------------------------------------ shared helpers
// hybrid_scatter_and_get_next_tlas_traversal_step at ray_depth == max_depth (path_inference_kernels.cu:306-321)
__device__ __forceinline__ void finish_pixel(const RenderParams &p, uint32_t pixel, const Ray &ray, const HitState &h)
{
    float illum[3] = {0.f, 0.f, 0.f};
    if (h.valid_hit == 0u)
        background(p, ray, illum);
    splat(p, pixel, illum[0], illum[1], illum[2], 1.0f);
    if (p.hits) {
        nbvh_hit_record r;
        r.t = h.best_t, r.valid_hit = h.valid_hit;
        r.normal_or_uv[0] = h.n_or_uv[0], r.normal_or_uv[1] = h.n_or_uv[1], r.normal_or_uv[2] = h.n_or_uv[2];
        r.albedo_or_ids[0] = h.alb_or_ids[0], r.albedo_or_ids[1] = h.alb_or_ids[1], r.albedo_or_ids[2] = h.alb_or_ids[2];
        p.hits[pixel] = r;
    }
}

// ------------------------------------------------------------------------------------------------ multi-sample helpers
// 8-bit frame of the shard's pixels after a multi-sample launch sequence: sRGB(accum / samples), as accumulate_pixel_value writes it
__global__ void __launch_bounds__(256)
k_resolve_ldr(const __grid_constant__ RenderParams p, float inv_samples)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.pixel_count)
        return;
    const uint32_t pixel = swizzle_pixel(p, p.first_pixel + i);
    const float4 a = p.accum[pixel];
    uchar4 px;
    px.x = (unsigned char)min(max(int(linear_to_srgb(a.x * inv_samples) * 256.0f), 0), 255);
    px.y = (unsigned char)min(max(int(linear_to_srgb(a.y * inv_samples) * 256.0f), 0), 255);
    px.z = (unsigned char)min(max(int(linear_to_srgb(a.z * inv_samples) * 256.0f), 0), 255);
    px.w = 255;
    p.ldr[pixel] = px;
}

__global__ void __launch_bounds__(256)
k_clear_accum(const __grid_constant__ RenderParams p)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < p.pixel_count)
        p.accum[swizzle_pixel(p, p.first_pixel + i)] = make_float4(0.f, 0.f, 0.f, 0.f);
}

#if NBVH_CHAIN_STATS
__device__ unsigned long long g_chain_stats[16 * 16];
__device__ unsigned int g_chain_done;
__device__ __forceinline__ uint32_t g_pred_dbg(const RenderParams &p, uint32_t pixel)
{
    uint32_t pred = 0u;
    if (!p.chain_hist)
        return 0u;
    const uint32_t w = p.cam.fb_w, hgt = p.cam.fb_h;
    const uint32_t px = pixel % w, py = pixel / w;
    const uint32_t x0 = px > 0u ? px - 1u : 0u, x1 = px + 1u < w ? px + 1u : px;
    const uint32_t y0 = py > 0u ? py - 1u : 0u, y1 = py + 1u < hgt ? py + 1u : py;
    for (uint32_t y = y0; y <= y1; ++y)
        for (uint32_t x = x0; x <= x1; ++x)
            pred = max(pred, chain_hist_value(p, p.chain_hist[y * w + x]));
    return pred < 15u ? pred : 15u;
}
#endif
// ------------------------------------------------------------------------------------------------ k_primary
// 64 registers -> 32 resident warps per SM (measured: 0.386 -> 0.367 ms at 1080p against 80 registers; the kernel is issue-bound and takes the extra warps)
#ifndef NBVH_PRIMARY_THREADS
#define NBVH_PRIMARY_THREADS 128 // measured on B200 (gpurun s69): 64 / 128 / 256 / 512 threads -> 0.347 / 0.342 / 0.346 / 0.377 ms per 1080p launch
#endif
__global__ void __launch_bounds__(NBVH_PRIMARY_THREADS, 1024 / NBVH_PRIMARY_THREADS)
k_primary(const __grid_constant__ RenderParams p)
{
    const uint32_t lane  = threadIdx.x & 31u;
    const uint32_t chunk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t sidx  = chunk * 32u + lane;
    const uint32_t n_spp = p.spp_count > 1u ? p.spp_count : 1u;
    const bool active    = sidx < p.pixel_count * n_spp;
    // multi-sample launch: slot = sample * pixel_count + pixel slot (pixel_count is a multiple of 32: a warp stays in one sample)
    const uint32_t smp   = sidx / p.pixel_count;
    uint32_t cnt_blas = 0, cnt_tri = 0, cnt_nbvh = 0;
    bool neural       = false;
    uint32_t pixel    = 0;
    uint32_t nstack[kNeuralStackSize];
    int nss        = 0;
    float seg_mint = 0.f, seg_maxt = 0.f;
    Ray ray;
    HitState h;
    h.best_t = FLT_MAX, h.valid_hit = 0u;
    h.n_or_uv[0] = h.n_or_uv[1] = h.n_or_uv[2] = 0.f;
    h.alb_or_ids[0] = h.alb_or_ids[1] = h.alb_or_ids[2] = 0.f;
    if (active) {
        uint32_t rng;
        pixel = swizzle_pixel(p, p.first_pixel + (sidx - smp * p.pixel_count));
        ray   = camera_ray(p.cam, pixel, p.sample_idx + (int)smp, p.sample_offset, rng);
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
                // non_neural_bvh_intersection (path_inference_kernels.cu:195-222): ray.maxt = closest t so far
                const TriHit th = blas_closest_hit(scene_blas(p.scene, blas_idx), ray, 1e-4f, h.best_t, cnt_blas, cnt_tri);
                if (th.t < h.best_t) { // set_or_update_with_preliminary_its_data (path_module.h:150-175)
                    h.best_t        = th.t;
                    h.valid_hit     = 2u;
                    h.n_or_uv[0]    = th.u, h.n_or_uv[1] = th.v, h.n_or_uv[2] = 0.f;
                    h.alb_or_ids[0] = __uint_as_float(blas_idx), h.alb_or_ids[1] = __uint_as_float(th.prim);
                    h.alb_or_ids[2] = 0.f;
                }
            }
        }
        if (neural) {
            // first leaf interval of the tree-cut, ray.maxt = closest classical hit (only prunes: see the header)
            uint32_t node;
            nstack[0] = 0u, nss = 1;
            neural = nbvh_next_interval(p.nbvh, ray, nstack, nss, h.best_t, seg_mint, seg_maxt, node, cnt_nbvh);
        }
    }
    // ---- queue compaction: one atomic per warp for the slot, one per class present in the warp for the drain order.
    // Longest-processing-time-first: k_neural drains the classes in order, so the rays expected to need the longest query chains
    // start first and the persistent kernel's tail consists of short chains (k_neural header).  The chain length
    // is uncorrelated with anything known about the ray at queue time, but strongly correlated from sample to sample of one pixel.
    const uint32_t cls = (p.class_lists && active && neural) ? chain_class(p, pixel, nss) : 0u;
    const uint32_t qmask = __ballot_sync(0xFFFFFFFFu, active && neural);
    if (qmask) {
        const uint32_t below = (1u << lane) - 1u;
        const int leader     = __ffs(qmask) - 1;
        uint32_t base = 0;
        if (lane == (uint32_t)leader)
            base = atomicAdd(p.queue_count, (uint32_t)__popc(qmask));
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        const uint32_t slot = base + __popc(qmask & below);
        if (active && neural) {
            NeuralQueueEntry *q = p.queue + slot;
            q->a = make_uint4(pixel, __float_as_uint(h.best_t), h.valid_hit, __float_as_uint(h.n_or_uv[0]));
            q->b = make_uint4(__float_as_uint(h.n_or_uv[1]), __float_as_uint(h.alb_or_ids[0]), __float_as_uint(h.alb_or_ids[1]), (uint32_t)nss);
#if NBVH_CHAIN_STATS
            q->c = make_uint4(__float_as_uint(seg_mint), __float_as_uint(seg_maxt), smp, g_pred_dbg(p, pixel) | (cls << 8));
#else
            q->c = make_uint4(__float_as_uint(seg_mint), __float_as_uint(seg_maxt), smp, 0u); // c.z: sample of a multi-sample launch
#endif
            for (int i = 0; i < nss; i += 4)
                q->st[i >> 2] = make_uint4(nstack[i], nstack[i + 1], nstack[i + 2], nstack[i + 3]);
        }
        if (p.class_lists)
            file_in_class(p, qmask, cls, slot, lane);
    }
    if (active && !neural)
        finish_pixel(p, pixel, ray, h);
    warp_add_counter(p.counters, CNT_RAYS, active ? 1u : 0u);
    warp_add_counter(p.counters, CNT_BLAS_STEPS, cnt_blas);
    warp_add_counter(p.counters, CNT_TRI_TESTS, cnt_tri);
    warp_add_counter(p.counters, CNT_NBVH_STEPS, cnt_nbvh);
    warp_add_counter(p.counters, CNT_QUEUED, active && neural ? 1u : 0u);
}

// ------------------------------------------------------------------------------------------------ k_neural
#if NBVH_PHASE_TIMING
__device__ unsigned long long g_probe[24];
__device__ unsigned int g_probe_done;
#endif
struct NeuralLane {
    Ray ray;
    uint32_t pixel;
    HitState h;
    float seg_mint, seg_maxt;
    int nss;
    bool have_ray, have_query;
    uint32_t nq;  // queries this ray has been through (chain-length history)
    uint32_t rng; // kUniform only: the ray's generator (segment parameters are drawn from it; path mode hands it back to the path state)
};

// kPath: multi-bounce mode -- the queue entry names a path slot, the ray is read from the path state and the closest
// hit goes to path_hits[slot] for the shading kernel (path.cu) instead of being splatted
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <bool kPath, bool kUniform>
__device__ __forceinline__ void finish_lane(const RenderParams &p, uint32_t id, const Ray &ray, const HitState &h, uint32_t rng, uint32_t nq)
{
    if constexpr (!kPath) {
#if NBVH_CHAIN_STATS
        atomicAdd(&g_chain_stats[(nq >> 16) * 16u + min(nq & 0xFFFFu, 15u)], 1ull);
        nq &= 0xFFFFu;
#endif
        if (p.chain_hist)
            chain_hist_store(p, id, nq);
    }
    if constexpr (kPath) {
        if (p.chain_hist && p.ray_depth == 1u) // camera rays of the multi-bounce pipeline keep the same per-pixel history
            chain_hist_store(p, p.paths[id].s3.z, nq);
        if constexpr (kUniform)
            p.paths[id].s3.x = rng; // HybridIntersectionResult takes the traversal data's throughput_rng back (path_module.h:111,141)
        PathHit ph;
        ph.h0 = make_float4(h.best_t, __uint_as_float(h.valid_hit), h.n_or_uv[0], h.n_or_uv[1]);
        ph.h1 = make_float4(h.n_or_uv[2], h.alb_or_ids[0], h.alb_or_ids[1], h.alb_or_ids[2]);
        p.path_hits[id] = ph;
        if (p.hits && p.ray_depth == 1u) { // primary-hit record (parity hook)
            nbvh_hit_record r;
            r.t = h.best_t, r.valid_hit = h.valid_hit;
            r.normal_or_uv[0] = h.n_or_uv[0], r.normal_or_uv[1] = h.n_or_uv[1], r.normal_or_uv[2] = h.n_or_uv[2];
            r.albedo_or_ids[0] = h.alb_or_ids[0], r.albedo_or_ids[1] = h.alb_or_ids[1], r.albedo_or_ids[2] = h.alb_or_ids[2];
            p.hits[p.paths[id].s3.z] = r;
        }
    } else {
        finish_pixel(p, id, ray, h);
    }
}

// kUniform: SegmentDistribution::Uniform (neural_bvh_module_kernels.cu:53-69,292-297) -- every query draws one parameter per segment
// point from its ray's generator; rows then encode their own points (no shipped configuration uses it: the cooperative gather paths
// serve the equidistant distributions only)
template <bool kPath, bool kUniform>
__global__ void __launch_bounds__(kNeuralThreads, kNeuralCtasPerSm)
k_neural(const __grid_constant__ RenderParams p, const __grid_constant__ NetConst nc)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t s_flags[kNeuralThreads / 32];
    // nothing queued (frames without neural rays): leave before TMEM / weight set-up.
    // The counts are final: every producer launch precedes this kernel in stream order.
    if (*reinterpret_cast<const volatile uint32_t *>(p.queue_count) == 0u)
        return;
    __shared__ uint32_t s_cum[kChainClasses + 1]; // class k of the drain order = queue positions [s_cum[k], s_cum[k + 1])
    if (threadIdx.x == 0) {
        uint32_t cum = 0u;
        for (int k = 0; k < kChainClasses; ++k) {
            s_cum[k] = cum;
            if (p.class_lists)
                cum += reinterpret_cast<const volatile uint32_t *>(p.class_count)[k];
        }
        s_cum[kChainClasses] = cum;
    }
#if NBVH_TILE_COOP
    __shared__ uint32_t s_qmask[kNeuralThreads / 32];
    __shared__ float s_ep[kNeuralThreads * 6];
#endif
    const SmemLayout lay     = smem_layout(nc.wimg_bytes, nc.in_k, kNeuralTiles);
    const uint32_t tmem_base = cta_setup<kNeuralTiles>(nc, smem, lay);
    const uint32_t tile      = __shfl_sync(0xFFFFFFFFu, threadIdx.x / kTileRows, 0); // warp-uniform for ptxas (mlp_tile_forward)
    const uint32_t row       = threadIdx.x % kTileRows;
    const uint32_t lane      = threadIdx.x & 31u;
    const uint32_t warp      = threadIdx.x >> 5;
    TileCtx t                = make_tile(smem, lay, nc, tile, tmem_base);

    const uint32_t n_queued = *reinterpret_cast<const volatile uint32_t *>(p.queue_count);
#if NBVH_NEURAL_SMEM_STACK
    const LaneStack nstack{reinterpret_cast<uint32_t *>(smem + neural_stack_offset(nc.wimg_bytes, nc.in_k)) + threadIdx.x};
#else
    uint32_t nstack[kNeuralStackSize];
#endif
    NeuralLane s;
    s.have_ray = s.have_query = false;
    s.nss = 0, s.pixel = 0u, s.nq = 0u;

    uint32_t chunk_base = 0u, chunk_used = 32u;
    bool exhausted = false;
    uint32_t cnt_q = 0, cnt_nbvh = 0, cnt_iters = 0;
    unsigned long long t_cta0 = 0ull;
    if (threadIdx.x == 0 && p.counters) {
        t_cta0 = global_timer_ns();
        atomicMax(p.counters + CNT_T_START_INV, ~t_cta0);
    }

#if NBVH_PHASE_TIMING
    unsigned long long ph[8] = {0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull};
    uint32_t iters_drain = 0u;
    unsigned long long mp_s[6] = {0ull, 0ull, 0ull, 0ull, 0ull, 0ull}, mp_d[6] = {0ull, 0ull, 0ull, 0ull, 0ull, 0ull};
    uint32_t wprobe[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#endif
    for (;;) {
#if NBVH_PHASE_TIMING
        const long long pt0 = clock64();
#if !NBVH_PHASE_SPARSE
        const uint32_t pdr  = exhausted ? 4u : 0u;
#endif
#endif
        // ---------------- walk phase: every lane works until it owns a pending interval (or the queue is drained)
        for (int round = 0; round < kWalkRounds; ++round) {
            uint32_t need_mask = __ballot_sync(0xFFFFFFFFu, !s.have_ray);
            while (need_mask && !exhausted) {
                if (chunk_used == 32u) {
                    uint32_t c = 0;
                    if (lane == 0)
                        c = atomicAdd(p.queue_cursor, 32u);
                    c = __shfl_sync(0xFFFFFFFFu, c, 0);
                    if (c >= n_queued) {
                        if (lane == 0 && p.counters)
                            atomicMax(p.counters + CNT_T_EXHAUST_INV, ~global_timer_ns());
                        exhausted = true;
                        break;
                    }
                    chunk_base = c, chunk_used = 0u;
                }
                const uint32_t avail = 32u - chunk_used;
                const uint32_t rank  = __popc(need_mask & ((1u << lane) - 1u));
                if (!s.have_ray && rank < avail) {
                    const uint32_t qi = chunk_base + chunk_used + rank;
                    if (qi < n_queued) {
                        uint32_t ei = qi;
                        if (p.class_lists) { // position in the drain order -> class list -> queue slot
                            int k = 0;
                            while (k < kChainClasses - 1 && qi >= s_cum[k + 1])
                                ++k;
                            ei = p.class_lists[(size_t)k * p.queue_capacity + (qi - s_cum[k])];
                        }
                        const NeuralQueueEntry *q = p.queue + ei;
                        const uint4 a = q->a, b = q->b, c = q->c;
                        uint32_t rng;
                        s.pixel       = a.x;
                        if constexpr (kPath) {
                            const float4 s0 = p.paths[a.x].s0, s1 = p.paths[a.x].s1;
                            s.ray.ox = s0.x, s.ray.oy = s0.y, s.ray.oz = s0.z;
                            s.ray.dx = s0.w, s.ray.dy = s1.x, s.ray.dz = s1.y;
                            ray_prepare(s.ray, p.cam.nodes_safe != 0u);
                            if constexpr (kUniform)
                                rng = p.paths[a.x].s3.x;
                        } else {
                            s.ray = camera_ray(p.cam, s.pixel, p.sample_idx + (int)c.z, p.sample_offset, rng);
                        }
                        if constexpr (kUniform)
                            s.rng = rng;
                        s.h.best_t    = __uint_as_float(a.y);
                        s.h.valid_hit = a.z;
                        s.h.n_or_uv[0] = __uint_as_float(a.w), s.h.n_or_uv[1] = __uint_as_float(b.x), s.h.n_or_uv[2] = 0.f;
                        s.h.alb_or_ids[0] = __uint_as_float(b.y), s.h.alb_or_ids[1] = __uint_as_float(b.z), s.h.alb_or_ids[2] = 0.f;
                        s.nss      = (int)b.w;
                        s.seg_mint = __uint_as_float(c.x), s.seg_maxt = __uint_as_float(c.y);
                        for (int i = 0; i < s.nss; i += 4) {
                            const uint4 v = q->st[i >> 2];
                            nstack[i] = v.x, nstack[i + 1] = v.y, nstack[i + 2] = v.z, nstack[i + 3] = v.w;
                        }
                        s.have_ray   = true;
                        s.have_query = true;
                        s.nq         = 0u;
#if NBVH_CHAIN_STATS
                        s.nq = (c.w & 15u) << 16; // predicted chain rides in the upper half
#endif
                    }
                }
                const uint32_t want = __popc(need_mask);
                chunk_used += want < avail ? want : avail;
                need_mask = __ballot_sync(0xFFFFFFFFu, !s.have_ray && !(rank < avail));
            }
            const bool work = s.have_ray && !s.have_query;
            if (!__any_sync(0xFFFFFFFFu, work))
                break;
            if (work) {
                // NeuralBVH2::intersect resumed with ray.maxt = closest t so far (neural_bvh_module_common.cuh:529-545)
                uint32_t node;
#if NBVH_WALK_BUDGET > 0
                int budget = NBVH_WALK_BUDGET;
                if (s.nss > 0 && nbvh_next_interval(p.nbvh, s.ray, nstack, s.nss, s.h.best_t, s.seg_mint, s.seg_maxt, node, cnt_nbvh, nullptr, &budget)) {
                    s.have_query = true;
                } else if (budget < 0) {
                    // parked mid-walk: the lane keeps its ray and walks on in the next round / after the tile's next network pass
#elif NBVH_PHASE_TIMING
                if (s.nss > 0 && nbvh_next_interval(p.nbvh, s.ray, nstack, s.nss, s.h.best_t, s.seg_mint, s.seg_maxt, node, cnt_nbvh, wprobe)) {
                    s.have_query = true;
#else
                if (s.nss > 0 && nbvh_next_interval(p.nbvh, s.ray, nstack, s.nss, s.h.best_t, s.seg_mint, s.seg_maxt, node, cnt_nbvh)) {
                    s.have_query = true;
#endif
                } else {
                    finish_lane<kPath, kUniform>(p, s.pixel, s.ray, s.h, s.rng, s.nq);
                    s.have_ray = false;
                }
            }
            __syncwarp();
            if (exhausted && !__any_sync(0xFFFFFFFFu, s.have_ray && !s.have_query))
                break;
        }
#if NBVH_PHASE_TIMING
        const long long pt1 = clock64();
#endif
        // ---------------- tile-wide agreement: anything to infer?  anything left at all?
#if NBVH_TILE_COOP
        uint32_t tile_masks[kTileRows / 32];
#endif
        {
            const uint32_t qm   = __ballot_sync(0xFFFFFFFFu, s.have_query);
            const uint32_t q    = qm ? 1u : 0u;
            const uint32_t live = (__any_sync(0xFFFFFFFFu, s.have_ray) || !exhausted) ? 2u : 0u;
            if (lane == 0)
                s_flags[warp] = q | live | (exhausted ? 0u : 4u); // bit 2: this warp can still refill from the queue
#if NBVH_TILE_COOP
            if (lane == 0)
                s_qmask[warp] = qm;
            if (!kUniform && s.have_query) {
                float eo[3], ei[3];
                encode_endpoints(p.enc, s.ray, s.seg_mint, s.seg_maxt, eo, ei);
                float *e = s_ep + 6u * threadIdx.x;
                e[0] = eo[0], e[1] = eo[1], e[2] = eo[2], e[3] = ei[0], e[4] = ei[1], e[5] = ei[2];
            }
#endif
            named_bar_sync(t.bar_id, kTileRows);
            uint32_t f = 0;
#pragma unroll
            for (int w = 0; w < kTileRows / 32; ++w)
                f |= s_flags[tile * (kTileRows / 32) + w];
#if NBVH_TILE_COOP
            for (int w = 0; w < kTileRows / 32; ++w)
                tile_masks[w] = s_qmask[tile * (kTileRows / 32) + w];
#endif
            named_bar_sync(t.bar_id, kTileRows);
            if ((f & 3u) == 0u)
                break;
            if (!(f & 1u))
                continue;
        }
        // ---------------- segment encoding + hash grid -> A tile (a8, a9)
#if NBVH_PHASE_TIMING
        const long long pt2 = clock64();
#endif
        if constexpr (kUniform) {
            float pts[12] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (s.have_query) {
                float eo[3], ei[3];
                encode_endpoints(p.enc, s.ray, s.seg_mint, s.seg_maxt, eo, ei);
                for (uint32_t k = 0; k < p.enc.n_points && k < 4u; ++k) {
                    const float sk = lcg_randomf(s.rng);
                    pts[3u * k] = lerp_point(eo[0], ei[0], sk), pts[3u * k + 1u] = lerp_point(eo[1], ei[1], sk), pts[3u * k + 2u] = lerp_point(eo[2], ei[2], sk);
                }
            }
            encode_row_to_tile(nc, t.a_tile, row, pts, s.have_query);
        } else
#if NBVH_TILE_COOP
        encode_rows_tile(nc, p.enc, t.a_tile, row, tile_masks, s_ep + 6u * tile * (uint32_t)kTileRows);
#else
        {
            float eo[3] = {0.f, 0.f, 0.f}, ei[3] = {0.f, 0.f, 0.f};
            if (s.have_query)
                encode_endpoints(p.enc, s.ray, s.seg_mint, s.seg_maxt, eo, ei);
            encode_rows_coop(nc, p.enc, t.a_tile, row & ~31u, lane, eo, ei, s.have_query);
        }
#endif
        fence_proxy_async_smem();
#if NBVH_PHASE_TIMING
        const long long pt3 = clock64();
#endif
        // ---------------- MLP on tensor cores (a10), decode + closest-hit update (a11)
        float y[8];
#if NBVH_PHASE_TIMING
#if NBVH_PHASE_SPARSE // second category = sparse iterations (at most NBVH_PHASE_SPARSE active rows in the tile) instead of "queue exhausted"
        const bool ph_sparse = __popc(tile_masks[0]) + __popc(tile_masks[1]) + __popc(tile_masks[2]) + __popc(tile_masks[3]) <= NBVH_PHASE_SPARSE;
        mlp_tile_forward(nc, t, row, y, ph_sparse ? mp_d : mp_s);
#else
        mlp_tile_forward(nc, t, row, y, exhausted ? mp_d : mp_s);
#endif
#else
        mlp_tile_forward(nc, t, row, y);
#endif
        ++cnt_iters;
#if NBVH_PHASE_TIMING
        {
            const long long pt4 = clock64();
#if NBVH_PHASE_SPARSE
            const uint32_t pdr = ph_sparse ? 4u : 0u;
#endif
            ph[pdr + 0] += (unsigned long long)(pt1 - pt0), ph[pdr + 1] += (unsigned long long)(pt2 - pt1);
            ph[pdr + 2] += (unsigned long long)(pt3 - pt2), ph[pdr + 3] += (unsigned long long)(pt4 - pt3);
            iters_drain += pdr ? 1u : 0u;
        }
#endif
        if (s.have_query) {
            ++cnt_q;
            ++s.nq;
            float t_hit;
            if (decode_hit(p.dec, s.ray, y[0], y[1], s.seg_mint, s.seg_maxt, kPath ? p.ray_depth : 1u, t_hit) && t_hit < s.h.best_t) {
                s.h.best_t    = t_hit;
                s.h.valid_hit = 1u;
                decode_normal(y[2], y[3], y[4], s.h.n_or_uv);
                s.h.alb_or_ids[0] = y[5], s.h.alb_or_ids[1] = y[6], s.h.alb_or_ids[2] = y[7];
            }
            s.have_query = false;
        }
    }
    warp_add_counter(p.counters, CNT_QUERIES, cnt_q);
    warp_add_counter(p.counters, CNT_NBVH_STEPS, cnt_nbvh);
    warp_add_counter(p.counters, CNT_NBVH_STEPS_NEURAL, cnt_nbvh);
    if (row == 0 && p.counters && cnt_iters)
        atomicAdd(p.counters + CNT_TILE_ITERS, (unsigned long long)cnt_iters);
#if NBVH_PHASE_TIMING
    if (row == 0 && p.counters) {
        for (int k = 0; k < 8; ++k)
            atomicAdd(p.counters + CNT_PH_WALK + k, ph[k]);
        atomicAdd(p.counters + CNT_ITERS_DRAIN, (unsigned long long)iters_drain);
        for (int k = 0; k < 6; ++k)
            atomicAdd(&g_probe[k], mp_s[k]), atomicAdd(&g_probe[8 + k], mp_d[k]);
    }
    for (int k = 0; k < 8; ++k) {
        const uint32_t v = __reduce_add_sync(0xFFFFFFFFu, wprobe[k]);
        if (lane == 0)
            atomicAdd(&g_probe[16 + k], (unsigned long long)v);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&g_probe_done, 1u) == gridDim.x - 1u) {
            printf("[k_neural probe] mlp steady: bar0 %llu issue %llu mmawait %llu epi %llu bar %llu fence %llu | drain: bar0 %llu issue %llu mmawait %llu epi %llu bar %llu fence %llu\n",
                   g_probe[0], g_probe[1], g_probe[2], g_probe[3], g_probe[4], g_probe[5], g_probe[8], g_probe[9], g_probe[10], g_probe[11], g_probe[12], g_probe[13]);
            printf("[k_neural probe] walk: pops %llu popped-leaf %llu leaf-behind-hit %llu popped-interior %llu (both culled %llu) descent-leaf %llu descent-interior-culled %llu skipped-entries %llu\n",
                   g_probe[16], g_probe[17], g_probe[18], g_probe[21], g_probe[19], g_probe[20], g_probe[22], g_probe[23]);
            for (int k = 0; k < 24; ++k)
                g_probe[k] = 0ull;
            g_probe_done = 0u;
        }
    }
#endif
    __syncthreads();
    if (threadIdx.x == 0 && p.counters) {
        const unsigned long long t1 = global_timer_ns();
        atomicMax(p.counters + CNT_T_END, t1);
        atomicAdd(p.counters + CNT_T_CTA_SUM, t1 - t_cta0);
    }
#if NBVH_CHAIN_STATS
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&g_chain_done, 1u) == gridDim.x - 1u) {
            for (int a = 0; a < 16; ++a) {
                printf("[chain stats] pred %2d:", a);
                for (int b = 0; b < 16; ++b)
                    printf(" %llu", g_chain_stats[a * 16 + b]), g_chain_stats[a * 16 + b] = 0ull;
                printf("\n");
            }
            g_chain_done = 0u;
        }
    }
#endif
    cta_teardown<kNeuralTiles>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ launchers
cudaError_t configure_render(uint32_t wimg_bytes, uint32_t in_k)
{
    (void)wimg_bytes, (void)in_k;
    cudaError_t e = allow_max_dynamic_smem(reinterpret_cast<const void *>(k_neural<false, false>));
    if (e == cudaSuccess)
        e = allow_max_dynamic_smem(reinterpret_cast<const void *>(k_neural<true, false>));
    if (e == cudaSuccess)
        e = allow_max_dynamic_smem(reinterpret_cast<const void *>(k_neural<false, true>));
    if (e == cudaSuccess)
        e = allow_max_dynamic_smem(reinterpret_cast<const void *>(k_neural<true, true>));
    return e;
}

cudaError_t launch_neural(const RenderParams &rp, const NetConst &nc, int num_sms, bool path_mode, cudaStream_t stream)
{
    const uint32_t smem_bytes = neural_stack_offset(nc.wimg_bytes, nc.in_k) + kNeuralStackBytes;
    const uint32_t grid = (uint32_t)num_sms * (uint32_t)kNeuralCtasPerSm;
    if (rp.enc.uniform) {
        if (path_mode)
            k_neural<true, true><<<grid, kNeuralThreads, smem_bytes, stream>>>(rp, nc);
        else
            k_neural<false, true><<<grid, kNeuralThreads, smem_bytes, stream>>>(rp, nc);
    } else if (path_mode)
        k_neural<true, false><<<grid, kNeuralThreads, smem_bytes, stream>>>(rp, nc);
    else
        k_neural<false, false><<<grid, kNeuralThreads, smem_bytes, stream>>>(rp, nc);
    return cudaGetLastError();
}

cudaError_t launch_render_split(const RenderParams &rp, const NetConst &nc, int num_sms, cudaStream_t stream, cudaEvent_t between)
{
    if (rp.pixel_count == 0)
        return cudaSuccess;
    const uint32_t slots    = rp.pixel_count * (rp.spp_count > 1u ? rp.spp_count : 1u);
    const uint32_t n_chunks = (slots + 31u) >> 5;
    if (rp.spp_count > 1u && rp.sample_idx == 0) // sample 0 overwrites the accumulation; here: clear, then 16-byte reductions
        k_clear_accum<<<(rp.pixel_count + 255u) / 256u, 256, 0, stream>>>(rp);
    constexpr uint32_t kWarpsPerBlock = NBVH_PRIMARY_THREADS / 32;
    k_primary<<<(n_chunks + kWarpsPerBlock - 1u) / kWarpsPerBlock, NBVH_PRIMARY_THREADS, 0, stream>>>(rp);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return e;
    if (between && (e = cudaEventRecord(between, stream)) != cudaSuccess)
        return e;
    e = launch_neural(rp, nc, num_sms, false, stream);
    if (e != cudaSuccess)
        return e;
    if (rp.spp_count > 1u && rp.ldr) {
        k_resolve_ldr<<<(rp.pixel_count + 255u) / 256u, 256, 0, stream>>>(rp, 1.f / (float)(rp.sample_idx + (int)rp.spp_count));
        return cudaGetLastError();
    }
    return cudaSuccess;
}

} // namespace nbvh

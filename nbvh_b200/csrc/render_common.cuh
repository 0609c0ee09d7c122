// render_common.cuh -- device helpers shared by the tile kernels (k_query, k_render, k_neural) and the primary pass.
#pragma once
#include "shading.cuh"

#ifndef NBVH_COOP_JMAJOR
#define NBVH_COOP_JMAJOR 1 // cooperative gathers dealt chunk-major (1) or row-major (0)
#endif

namespace nbvh {

// closest hit of the ray in flight (NeuralTraversalResult / HybridIntersectionResult, path_module.h:19-45,104-210)
struct HitState {
    float best_t;
    uint32_t valid_hit;
    float n_or_uv[3];
    float alb_or_ids[3];
};

// ------------------------------------------------------------------------------------------------ CTA setup
__host__ __device__ constexpr uint32_t tmem_cols(int tiles) { return tiles <= 1 ? 64u : tiles <= 2 ? 128u : tiles <= 4 ? 256u : 512u; }

// Loads the packed weights with one TMA bulk copy, allocates TMEM, initialises the per-tile mbarriers.
// Returns the TMEM base address.  Must be called by all threads of the CTA.
template <int kTilesPerCta>
__device__ __forceinline__ uint32_t cta_setup(const NetConst &nc, uint8_t *smem, const SmemLayout &lay)
{
    const uint32_t tid      = threadIdx.x;
    const uint32_t bar_base = smem_u32(smem + lay.bar_off);
    const uint32_t w_bar    = bar_base + 8u * kTilesPerCta;
    if (tid == 0) {
        for (int t = 0; t < kTilesPerCta; ++t)
            mbar_init(bar_base + 8u * t, 1u);
        mbar_init(w_bar, 1u);
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(w_bar, nc.wimg_bytes);
        tma_bulk_g2s(smem_u32(smem + lay.w_off), nc.wimg, nc.wimg_bytes, w_bar);
    }
    if (tid < 32) {
        tmem_alloc(smem_u32(smem + lay.tmem_off), tmem_cols(kTilesPerCta)); // 64 fp32 accumulator columns per tile
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(smem + lay.tmem_off);
    mbar_wait(w_bar, 0u); // weights landed (async proxy writes are visible after the mbarrier completes)
    return tmem_base;
}

template <int kTilesPerCta>
__device__ __forceinline__ void cta_teardown(uint32_t tmem_base)
{
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, tmem_cols(kTilesPerCta));
    }
}

__device__ __forceinline__ TileCtx make_tile(uint8_t *smem, const SmemLayout &lay, const NetConst &nc, uint32_t tile,
                                             uint32_t tmem_base)
{
    const uint32_t a_bytes = kTileRows * (nc.in_k > (uint32_t)kNeurons ? nc.in_k : (uint32_t)kNeurons) * 2u;
    TileCtx t;
    t.a_tile   = smem + lay.a_off + tile * a_bytes;
    t.w_smem   = smem_u32(smem + lay.w_off);
    t.mma_bar  = smem_u32(smem + lay.bar_off) + 8u * tile;
    t.tmem_acc = tmem_base + 64u * tile;
    t.bar_id   = 1u + tile;
    t.phase    = 0u;
    return t;
}

// Warp-cooperative segment encoding for SPARSE tiles (k_neural: walk divergence and the drain of dependent query chains leave
// ~19 of 32 lanes with a pending query, far fewer in the tail).  The 16-byte chunks (two (point, level) pairs, 16 table gathers)
// of the warp's ACTIVE rows are dealt round-robin to all 32 lanes, so a warp with n active rows spends ceil(n * in_k/8 / 32)
// gather rounds instead of in_k/8 and idle lanes hide the gather latency of busy ones.  The segment end points travel by
// shuffle; every chunk is computed by the same arithmetic as encode_row_to_tile (lerp_point + grid_level_features), so the A tile
// holds the same bits.  Rows without a query are not written: rows of a GEMM are independent and their outputs are never read.
__device__ __forceinline__ void encode_rows_coop(const NetConst &nc, const EncodeConst &ec, uint8_t *a_tile, uint32_t warp_row0,
                                                 uint32_t lane, const float eo[3], const float ei[3], bool valid)
{
    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, valid);
    if (mask == 0u)
        return;
    const uint32_t pairs  = nc.in_k >> 3;
    const uint32_t items  = (uint32_t)__popc(mask) * pairs;
    const uint32_t my_src = __fns(mask, 0u, lane + 1u); // lane of the (lane+1)-th active row (garbage beyond the count: never used)
    const uint32_t n_rows = (uint32_t)__popc(mask);
    for (uint32_t base = 0; base < items; base += 32u) {
        const uint32_t it = base + lane;
        const bool on     = it < items;
#if NBVH_COOP_JMAJOR
        // chunk-major: neighbouring lanes gather the same (point, level) pairs of different rows (coherent rays share grid cells)
        const uint32_t j  = on ? it / n_rows : 0u;
        const uint32_t r  = on ? it - j * n_rows : 0u;
#else
        const uint32_t r  = on ? it / pairs : 0u;
        const uint32_t j  = it - r * pairs;
#endif
        const uint32_t src = __shfl_sync(0xFFFFFFFFu, my_src, r) & 31u;
        const float ox = __shfl_sync(0xFFFFFFFFu, eo[0], src), oy = __shfl_sync(0xFFFFFFFFu, eo[1], src),
                    oz = __shfl_sync(0xFFFFFFFFu, eo[2], src);
        const float ix = __shfl_sync(0xFFFFFFFFu, ei[0], src), iy = __shfl_sync(0xFFFFFFFFu, ei[1], src),
                    iz = __shfl_sync(0xFFFFFFFFu, ei[2], src);
        if (on) {
            const uint32_t q0 = 2u * j, q1 = 2u * j + 1u;
            const uint32_t p0 = q0 / nc.n_levels, l0 = q0 - p0 * nc.n_levels;
            const uint32_t p1 = q1 / nc.n_levels, l1 = q1 - p1 * nc.n_levels;
            const float s0 = ec.point_s[p0], s1 = ec.point_s[p1];
            const uint2 f0 = grid_level_features(nc.grid, nc.lv[l0], lerp_point(ox, ix, s0), lerp_point(oy, iy, s0), lerp_point(oz, iz, s0));
            const uint2 f1 = grid_level_features(nc.grid, nc.lv[l1], lerp_point(ox, ix, s1), lerp_point(oy, iy, s1), lerp_point(oz, iz, s1));
            *reinterpret_cast<uint4 *>(a_tile + tile_chunk_off(warp_row0 + src, j, nc.in_k)) = make_uint4(f0.x, f0.y, f1.x, f1.y);
        }
    }
}

// The same for the whole 128-row tile: the chunks of ALL active rows of the tile are dealt round-robin to its 128 lanes, so the four
// warps finish together whatever their share of the active rows (per warp, the tile waits for the warp with the most queries: 7 k of
// the 37 k cycles of a drain iteration).  The segment end points travel through shared memory (`ep`: 6 floats per row, written by the
// row's owner before the tile barrier); `masks` = the four warps' ballots of rows with a query.  Same arithmetic per chunk, same bits.
__device__ __forceinline__ void encode_rows_tile(const NetConst &nc, const EncodeConst &ec, uint8_t *a_tile, uint32_t row,
                                                 const uint32_t masks[4], const float *ep)
{
    const uint32_t c0 = (uint32_t)__popc(masks[0]), c1 = (uint32_t)__popc(masks[1]), c2 = (uint32_t)__popc(masks[2]),
                   c3 = (uint32_t)__popc(masks[3]);
    const uint32_t n_rows = c0 + c1 + c2 + c3;
    const uint32_t pairs  = nc.in_k >> 3;
    const uint32_t items  = n_rows * pairs;
    for (uint32_t it = row; it < items; it += (uint32_t)kTileRows) {
#if NBVH_COOP_JMAJOR
        const uint32_t j = it / n_rows;
        uint32_t r       = it - j * n_rows;
#else
        uint32_t r       = it / pairs;
        const uint32_t j = it - r * pairs;
#endif
        uint32_t w = 0u, m = masks[0];
        if (r >= c0) {
            r -= c0, w = 1u, m = masks[1];
            if (r >= c1) {
                r -= c1, w = 2u, m = masks[2];
                if (r >= c2)
                    r -= c2, w = 3u, m = masks[3];
            }
        }
        const uint32_t src = 32u * w + __fns(m, 0u, r + 1u);
        const float *e     = ep + 6u * src;
        const float ox = e[0], oy = e[1], oz = e[2], ix = e[3], iy = e[4], iz = e[5];
        const uint32_t q0 = 2u * j, q1 = 2u * j + 1u;
        const uint32_t p0 = q0 / nc.n_levels, l0 = q0 - p0 * nc.n_levels;
        const uint32_t p1 = q1 / nc.n_levels, l1 = q1 - p1 * nc.n_levels;
        const float s0 = ec.point_s[p0], s1 = ec.point_s[p1];
        const uint2 f0 = grid_level_features(nc.grid, nc.lv[l0], lerp_point(ox, ix, s0), lerp_point(oy, iy, s0), lerp_point(oz, iz, s0));
        const uint2 f1 = grid_level_features(nc.grid, nc.lv[l1], lerp_point(ox, ix, s1), lerp_point(oy, iy, s1), lerp_point(oz, iz, s1));
        *reinterpret_cast<uint4 *>(a_tile + tile_chunk_off(src, j, nc.in_k)) = make_uint4(f0.x, f0.y, f1.x, f1.y);
    }
}

__device__ __forceinline__ void warp_add_counter(unsigned long long *counters, int slot, uint32_t v)
{
    v = __reduce_add_sync(0xFFFFFFFFu, v);
    if ((threadIdx.x & 31u) == 0u && v != 0u && counters)
        atomicAdd(counters + slot, (unsigned long long)v);
}

// accumulate_pixel_value (include/cuda/neural/neural_bvh_renderer_kernels.cuh:81-127) on linear memory
__device__ __forceinline__ float linear_to_srgb(float x)
{
    return (x <= 0.0031308f) ? 12.92f * x : 1.055f * powf(fmaxf(x, 0.f), 1.f / 2.4f) - 0.055f;
}

__device__ __forceinline__ void splat(const RenderParams &p, uint32_t pixel, float r, float g, float b, float alpha)
{
    if (p.spp_count > 1u) {
        // several samples of one pixel are in flight in the same launch: accumulate with one 16-byte reduction
        // (RED.E.ADD.F32x4); the 8-bit frame is resolved by k_resolve_ldr when the launch sequence is complete
        atomicAdd(p.accum + pixel, make_float4(r, g, b, alpha));
        return;
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.sample_idx > 0)
        acc = p.accum[pixel];
    if (p.ema_accumulate) {
        if (alpha == 0.f && p.sample_idx > 0)
            return;
        if (p.sample_idx > 0) {
            const float w = p.ema_weight, iw = 1.f - p.ema_weight;
            acc = make_float4(r * w + acc.x * iw, g * w + acc.y * iw, b * w + acc.z * iw, alpha * w + acc.w * iw);
        } else {
            acc = make_float4(r, g, b, alpha);
        }
        p.accum[pixel] = acc;
    } else {
        acc            = make_float4(acc.x + r, acc.y + g, acc.z + b, acc.w + alpha);
        p.accum[pixel] = acc;
        const float inv = 1.f / float(p.sample_idx + 1);
        acc = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
    }
    if (p.ldr) {
        uchar4 px;
        px.x = (unsigned char)min(max(int(linear_to_srgb(acc.x) * 256.0f), 0), 255);
        px.y = (unsigned char)min(max(int(linear_to_srgb(acc.y) * 256.0f), 0), 255);
        px.z = (unsigned char)min(max(int(linear_to_srgb(acc.z) * 256.0f), 0), 255);
        px.w = 255;
        p.ldr[pixel] = px;
    }
}

// background_hit_eval (path_inference_kernels.cu:37-59)
__device__ __forceinline__ void background(const RenderParams &p, const Ray &ray, float bg[3])
{
    const V3 c = background_eval(p.bg, ray.dx, ray.dy, ray.dz);
    bg[0] = c.x, bg[1] = c.y, bg[2] = c.z;
}

// 8x4 pixel blocks per 32-pixel chunk keep the rays of a warp spatially coherent (hash-grid and node reuse in L1).
// Interleaved shards (multi-GPU tile partition): the launch owns every chunk_stride-th 32-pixel chunk starting at
// chunk_offset, so its slot s maps to the global slot ((s / 32) * stride + offset) * 32 + s % 32.
__device__ __forceinline__ uint32_t swizzle_pixel(const RenderParams &p, uint32_t s)
{
    if (p.chunk_stride > 1u)
        s = (((s - p.first_pixel) >> 5) * p.chunk_stride + p.chunk_offset) * 32u + (s & 31u);
    const uint32_t w = p.cam.fb_w, h = p.cam.fb_h;
    if ((w & 7u) || (h & 3u))
        return s;
    const uint32_t chunk = s >> 5, k = s & 31u;
    const uint32_t bpr = w >> 3;
    const uint32_t bx = chunk % bpr, by = chunk / bpr;
    return (by * 4u + (k >> 3)) * w + bx * 8u + (k & 7u);
}

#ifndef NBVH_LONG_STACK
#define NBVH_LONG_STACK 4
#endif
constexpr int kLongStack = NBVH_LONG_STACK; // pending-stack depth from which a queued ray without history counts as a long query chain

// ---- drain order of the neural ray queue (RenderParams::class_lists; k_primary / k_path_trace<first> file, k_neural drains).
// Longest-processing-time-first: the rays expected to need the longest query chains start first, so that the tail of the persistent
// kernel consists of short chains.  The chain length is uncorrelated with anything known about the ray at queue time (DESIGN
// section 4), but the longest chain among the last frame's camera rays of the 3x3 pixel neighbourhood bounds it: measured on Chess
// at 1080p, 99.2 % of the rays predicted 1 need 1 query and 0.03 % need more than 2 (profiles/r2_s46_chain_stats.txt).
// History byte: the chain length of the last frame; with chain_flags bit 2 two nibbles, the last frame's (low) and the one before
// (high), and the prediction is their maximum -- halves the mispredicted long chains among the rays drained last.
__device__ __forceinline__ uint32_t chain_hist_value(const RenderParams &p, uint8_t v)
{
    return (p.chain_flags & 4u) ? max((uint32_t)v & 15u, (uint32_t)v >> 4) : (uint32_t)v;
}
__device__ __forceinline__ void chain_hist_store(const RenderParams &p, uint32_t pixel, uint32_t nq)
{
    if (p.chain_flags & 4u) {
        const uint32_t old = p.chain_hist[pixel];
        p.chain_hist[pixel] = (uint8_t)(((old & 15u) << 4) | (nq < 15u ? nq : 15u));
        return;
    }
    uint32_t v = nq < 255u ? nq : 255u;
    if (p.chain_flags & 2u) { // slow decay: a pixel that needed a long chain recently stays suspicious for a few frames
        const uint32_t old = p.chain_hist[pixel];
        v = max(v, old > 0u ? old - 1u : 0u);
    }
    p.chain_hist[pixel] = (uint8_t)v;
}
__device__ __forceinline__ uint32_t chain_class(const RenderParams &p, uint32_t pixel, int nss)
{
    uint32_t pred = 0u;
    if (p.chain_hist) {
        if (p.chain_flags & 1u) { // the jittered ray may fall anywhere in the pixel: the longest chain of the 3x3 neighbourhood
            const uint32_t w = p.cam.fb_w, hgt = p.cam.fb_h;
            const uint32_t px = pixel % w, py = pixel / w;
            const uint32_t x0 = px > 0u ? px - 1u : 0u, x1 = px + 1u < w ? px + 1u : px;
            const uint32_t y0 = py > 0u ? py - 1u : 0u, y1 = py + 1u < hgt ? py + 1u : py;
            for (uint32_t y = y0; y <= y1; ++y)
                for (uint32_t x = x0; x <= x1; ++x)
                    pred = max(pred, chain_hist_value(p, p.chain_hist[y * w + x]));
        } else
            pred = chain_hist_value(p, p.chain_hist[pixel]);
    }
    if (pred == 0u) // no history (first frame, or the neighbourhood queued nothing): pending-stack depth
        return nss >= kLongStack ? 2u : 3u;
    return (pred < (p.chain_thr >> 24) ? 1u : 0u) + (pred < ((p.chain_thr >> 16) & 255u) ? 1u : 0u) +
           (pred < ((p.chain_thr >> 8) & 255u) ? 1u : 0u) + (pred < (p.chain_thr & 255u) ? 1u : 0u);
}

// one atomic per class present in the warp; qmask = lanes that queued a ray (warp-uniform), slot = this lane's queue slot
__device__ __forceinline__ void file_in_class(const RenderParams &p, uint32_t qmask, uint32_t cls, uint32_t slot, uint32_t lane)
{
    const uint32_t below = (1u << lane) - 1u;
    uint32_t rem = qmask;
    while (rem) {
        const int l0     = __ffs(rem) - 1;
        const uint32_t k = __shfl_sync(0xFFFFFFFFu, cls, l0);
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, ((qmask >> lane) & 1u) && cls == k);
        uint32_t cb = 0;
        if (lane == (uint32_t)l0)
            cb = atomicAdd(p.class_count + k, (uint32_t)__popc(m));
        cb = __shfl_sync(0xFFFFFFFFu, cb, l0);
        if ((m >> lane) & 1u)
            p.class_lists[(size_t)k * p.queue_capacity + cb + __popc(m & below)] = slot;
        rem &= ~m;
    }
}

// multi-sample launches (render.cu): clear / resolve the shard's pixels
__global__ void k_clear_accum(const __grid_constant__ RenderParams p);
__global__ void k_resolve_ldr(const __grid_constant__ RenderParams p, float inv_samples);

} // namespace nbvh

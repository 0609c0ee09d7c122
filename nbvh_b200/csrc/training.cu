// training.cu -- the online-optimisation step of the neural ray-query path for sm_100a.
//
// Replaces tcnn::Trainer::training_step on the reference's model (ext/tiny-cuda-nn trainer.h:163-190):
//   k_encode_features  grid forward for a training batch: 9 x B inputs -> 96-wide fp16 feature tiles          (K3 x3)
//   k_train_mlp        per 128-sample tile, on tcgen05/TMEM: forward with the hidden activations kept in shared
//                      memory, composite loss + gradient, dgrad chain, and the five weight-gradient GEMMs whose fp32
//                      accumulators stay resident in TMEM across all tiles of the CTA                     (K11,K12,K13)
//   k_grid_backward    scatter of dL/dfeatures into the fp32 grid-gradient table: shared-memory privatised for the
//                      small dense levels, vector reductions (REDG.E.ADD.F32x4) for the rest                  (K14 x3)
//   k_adam             Adam under ExponentialDecay with tcnn's per-parameter step counters and zero-gradient skip;
//                      also refreshes the fp16 parameters and the packed tensor-core weight image            (K15)
// (kernel numbers: SURVEY.md section 2.4; /root/reference/ext/tiny-cuda-nn/include/tiny-cuda-nn/encodings/grid.h:49-320,
//  src/fully_fused_mlp.cu:151-259,736-836, include/tiny-cuda-nn/optimizers/adam.h:48-123,
//  /root/reference/include/cuda/neural/utils/pt_composite_loss.cuh:19-123.)
//
// Numerics: fp16 operands, fp32 accumulation everywhere (tcnn accumulates in fp16); gradients carry tcnn's loss scale
// of 128 and are rounded to fp16 only at the point where tcnn stores them, so Adam's "gradient == 0 -> skip" rule sees
// the same underflow.
#include "kernels.cuh"

#include <cuda_fp16.h>
#include <cstdlib>
#include <cstdio>

namespace nbvh {

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ float h2f_round(float x) { return __half2float(__float2half_rn(x)); }

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// Positions, weights and entry indices of one (point, level): identical arithmetic to grid_level_features
struct GridCorner {
    uint32_t index[8];
    float weight[8];
};

__device__ __forceinline__ void grid_corners(const GridLevel &L, float x, float y, float z, GridCorner &gc)
{
    float px = __fmaf_rn(L.scale, x, 0.5f), py = __fmaf_rn(L.scale, y, 0.5f), pz = __fmaf_rn(L.scale, z, 0.5f);
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    const uint32_t gx = (uint32_t)(int)fx, gy = (uint32_t)(int)fy, gz = (uint32_t)(int)fz;
    px = __fsub_rn(px, fx), py = __fsub_rn(py, fy), pz = __fsub_rn(pz, fz);
    const float qx = __fsub_rn(1.f, px), qy = __fsub_rn(1.f, py), qz = __fsub_rn(1.f, pz);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint32_t cx = gx + (c & 1), cy = gy + ((c >> 1) & 1), cz = gz + ((c >> 2) & 1);
        uint32_t index;
        if (L.dense)
            index = cx + cy * L.res + cz * L.res * L.res;
        else
            index = cx ^ (cy * 2654435761u) ^ (cz * 805459861u);
        gc.index[c] = grid_wrap(index, L.mask, L.size);
        float w     = (c & 1) ? px : qx;
        w           = __fmul_rn(w, (c & 2) ? py : qy);
        gc.weight[c] = __fmul_rn(w, (c & 4) ? pz : qz);
    }
}

// byte offset of the 8-byte (point, level) slot j of sample s in the tile-canonical feature layout
__device__ __forceinline__ size_t feat_slot_offset(uint32_t s, uint32_t j, uint32_t in_k)
{
    const uint32_t tile = s >> 7, row = s & 127u;
    return (size_t)tile * (kTileRows * in_k * 2u) + tile_chunk_off(row, j >> 1, in_k) + (j & 1u) * 8u;
}

// ------------------------------------------------------------------------------------------------ k_encode_features
// One thread per (sample, point, level).  Lanes are ordered so that 16 consecutive lanes write one 128-byte core
// matrix of the tile (8 rows x {even, odd} slot): coalesced stores, 8 independent 8-byte gathers in flight per thread.
__global__ void __launch_bounds__(256)
k_encode_features(const __grid_constant__ NetConst nc, const float *__restrict__ in, uint32_t batch, uint32_t padded_batch,
                  uint32_t n_in_dims, uint8_t *__restrict__ feat)
{
    const uint32_t slots = nc.in_k >> 2; // (point, level) slots per sample
    const uint64_t total = (uint64_t)padded_batch * slots;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        // i -> (row group of 8 samples, slot pair, row in group, slot parity)
        const uint32_t pairs = slots >> 1;
        const uint32_t lo    = (uint32_t)(i & 15u);
        const uint64_t hi    = i >> 4;
        const uint32_t pair  = (uint32_t)(hi % pairs);
        const uint32_t rgrp  = (uint32_t)(hi / pairs);
        const uint32_t s     = rgrp * 8u + (lo >> 1);
        const uint32_t j     = pair * 2u + (lo & 1u);
        uint2 f              = make_uint2(0u, 0u);
        if (s < batch) {
            const uint32_t p = j / nc.n_levels, l = j % nc.n_levels;
            const float *x   = in + (size_t)s * n_in_dims + 3u * p;
            f                = grid_level_features(nc.grid, nc.lv[l], __ldg(x), __ldg(x + 1), __ldg(x + 2));
        }
        *reinterpret_cast<uint2 *>(feat + feat_slot_offset(s, j, nc.in_k)) = f;
    }
}

// One thread per (sample, level): the points of a segment that fall into the same cell of the level share ONE gather of its eight
// corner entries (coarse levels: mostly all of them), blended per point with grid_level_features' arithmetic in its order -- same bits,
// fewer lookups for a kernel that runs at 93 % of the L1TEX peak.  Consecutive lanes = consecutive samples of one level.
__global__ void __launch_bounds__(256)
k_encode_features_merge(const __grid_constant__ NetConst nc, const float *__restrict__ in, uint32_t batch, uint32_t padded_batch,
                        uint32_t n_in_dims, uint8_t *__restrict__ feat)
{
    const uint64_t total = (uint64_t)padded_batch * nc.n_levels;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t l = (uint32_t)(i / padded_batch), s = (uint32_t)(i % padded_batch);
        if (s >= batch) {
            for (uint32_t p = 0; p < nc.n_points; ++p)
                *reinterpret_cast<uint2 *>(feat + feat_slot_offset(s, p * nc.n_levels + l, nc.in_k)) = make_uint2(0u, 0u);
            continue;
        }
        const GridLevel &L = nc.lv[l];
        GridCell prev;
        prev.gx = prev.gy = prev.gz = 0u, prev.px = prev.py = prev.pz = 0.f;
        uint2 g[8];
        for (uint32_t p = 0; p < nc.n_points; ++p) {
            const float *x   = in + (size_t)s * n_in_dims + 3u * p;
            const GridCell c = grid_cell(L, __ldg(x), __ldg(x + 1), __ldg(x + 2));
            if (p == 0u || c.gx != prev.gx || c.gy != prev.gy || c.gz != prev.gz)
                grid_gather_cell(nc.grid, L, c, g);
            prev = c;
            *reinterpret_cast<uint2 *>(feat + feat_slot_offset(s, p * nc.n_levels + l, nc.in_k)) = grid_blend_cell(c, g);
        }
    }
}

// ------------------------------------------------------------------------------------------------ composite loss (a18)
__device__ __forceinline__ void loss_and_gradient(float target, float prediction, uint32_t type, float &loss, float &grad)
{
    const float diff = prediction - target;
    switch (type) {
    case 0: { // RelativeL2
        const float denom = prediction * prediction + 1e-2f;
        loss = diff * diff / denom, grad = 2.0f * diff / denom;
        break;
    }
    case 3: // L1
        loss = fabsf(diff), grad = copysignf(1.0f, diff);
        break;
    case 2: { // RelativeL1
        const float scale = 1.0f / (fabsf(prediction) + 1e-2f);
        loss = fabsf(diff) * scale, grad = copysignf(scale, diff);
        break;
    }
    case 5: { // BinaryCrossEntropy
        const float eps  = 1e-5f;
        const float pred = fminf(fmaxf(prediction, eps), 1.f - eps);
        loss = -(target * logf(pred) + (1.f - target) * logf(1.f - pred));
        grad = -(target - pred) / (pred - (pred * pred));
        break;
    }
    case 6: { // LogL2
        const float d = logf(1.f + prediction) - logf(1.f + target);
        loss = d * d, grad = 2.0f * d / (prediction + 1.f);
        break;
    }
    case 7: { // LHalf
        const float s = sqrtf(fabsf(diff));
        loss = s, grad = 1.f / (2.f * s + 1e-5f);
        break;
    }
    default: // L2
        loss = diff * diff, grad = 2.0f * diff;
        break;
    }
}

// pt_composite_loss.cuh:19-123 for output element `e` of one sample; returns loss / n_total, grad = scale*g/n_total
// `target` = tgt[e], `target0` = tgt[0] (the visibility target that masks the other terms).  Division by n_total: when n_total is a
// power of two (every shipped batch size) the product with its exact reciprocal `inv_n` has the same bits and no division sequence.
__device__ __forceinline__ float loss_element(const LossConst &lc, float target, float target0, float prediction, uint32_t e, float n_total,
                                              float inv_n, bool n_pow2, float loss_scale, float &grad_out)
{
    const uint32_t depth0 = lc.vis_dim, normal0 = depth0 + lc.depth_dim, albedo0 = normal0 + lc.normal_dim,
                   end = albedo0 + lc.albedo_dim;
    float loss = 0.f, grad = 0.f;
    if (e < end) {
        const bool ignore  = lc.vis_masking ? (target0 < 0.5f) : false;
        if (lc.vis_dim > 0 && e == 0) {
            loss_and_gradient(target, prediction, lc.vis_loss, loss, grad);
            loss *= lc.vis_scale, grad *= lc.vis_scale;
        } else if (!ignore && e >= depth0 && e < normal0) {
            loss_and_gradient(target, prediction, lc.depth_loss, loss, grad);
            if (lc.depth_linear)
                grad *= (lc.depth_loss == 1u || lc.depth_loss == 0u) ? 100.f : 10.f;
            loss *= lc.depth_scale, grad *= lc.depth_scale;
        } else if (!ignore && e >= normal0 && e < albedo0) {
            if (lc.normal_linear) {
                const float n_length = 0.1f;
                const float lin      = ((prediction * 2.f - 1.f) / n_length) * 0.5f + 0.5f;
                loss_and_gradient(target, lin, lc.normal_loss, loss, grad);
                grad *= 1.f / n_length;
            } else {
                loss_and_gradient(target, prediction, lc.normal_loss, loss, grad);
            }
            loss *= lc.normal_scale, grad *= lc.normal_scale;
        } else if (!ignore && e >= albedo0 && e < end) {
            if (lc.albedo_linear) {
                const float n_length = 0.1f;
                loss_and_gradient(target, prediction / n_length, lc.albedo_loss, loss, grad);
                grad *= 1.f / n_length;
            } else {
                loss_and_gradient(target, prediction, lc.albedo_loss, loss, grad);
            }
            loss *= lc.albedo_scale, grad *= lc.albedo_scale;
        }
    }
    if (n_pow2) {
        grad_out = (loss_scale * grad) * inv_n;
        return loss * inv_n;
    }
    grad_out = loss_scale * grad / n_total;
    return loss / n_total;
}

// ------------------------------------------------------------------------------------------------ k_train_mlp
// UMMA descriptor of an MN-major no-swizzle operand: 16-byte rows hold 8 MN-contiguous elements, a core matrix is 8
// such rows along K; SBO = byte stride between 8-element groups along MN, LBO = byte stride between 8-row groups along K.
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr, uint32_t k_group_stride, uint32_t mn_group_stride)
{
    return umma_desc(smem_addr, k_group_stride, mn_group_stride);
}

__host__ __device__ constexpr uint32_t umma_idesc(uint32_t m, uint32_t n, uint32_t a_mn, uint32_t b_mn)
{
    return (1u << 4) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// ---- k_train_mlp: TMEM column map and shared-memory layout
// kSlots tiles are in flight per CTA (two when shared memory and TMEM allow it).  Slot s owns TMEM columns
// [s * slot_cols, (s + 1) * slot_cols): the 128 x 64 working accumulator (forward pre-activations / dgrad), which the
// 128 x in_k dL/dfeatures accumulator of layer 0 overlays (the working accumulator is dead by then).  The weight-gradient
// accumulators -- layer 0 (64 x in_k), hidden (64 x 64 each), last (64 x 16, transposed) -- follow and are SHARED by the slots.
// (train_slot_cols / train_tmem_cols: kernels.cuh)

struct TrainSmem {
    uint32_t w_off, slot_off, slot_bytes, a0_off, h_off, dzo_off, bar_off, tmem_off, total; // a0/h/dzo: relative to the slot
};

// Per slot: feature tile, one 128 x 64 tile per hidden layer (the activations H_l, overwritten IN PLACE by dZ_l during the
// backward pass: dZ_l = dA_l . ReLU'(H_l) is computed row by row by the thread that owns the row, after every MMA that read
// H_l has completed), and the 128 x 16 dZ tile of the output layer.
__host__ __device__ inline TrainSmem train_smem_layout(uint32_t wimg_bytes, uint32_t in_k, uint32_t n_hidden, uint32_t slots)
{
    TrainSmem s;
    s.w_off      = 0;
    s.slot_off   = (wimg_bytes + 127u) & ~127u;
    s.a0_off     = 0;
    s.h_off      = kTileRows * in_k * 2u;
    s.dzo_off    = s.h_off + n_hidden * kTileRows * kNeurons * 2u;
    s.slot_bytes = s.dzo_off + kTileRows * kOutPadded * 2u;
    s.bar_off    = s.slot_off + slots * s.slot_bytes;
    s.tmem_off   = s.bar_off + 8u + 32u * slots; // mbarriers: weights, then {features, operands ready, MMAs done, weight-gradient MMAs done} per slot
    s.total      = s.tmem_off + 16u;
    return s;
}

// non-blocking: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                 "selp.u32 %0, 1, 0, p;\n"
                 "}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return ok != 0u;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Warp roles: warps [8 s, 8 s + 8) are the epilogue warps of slot s -- two groups of four: thread r of either group owns sample row
// r = TMEM lane r of the slot's accumulators, group g converts columns [32 g, 32 g + 32) (a warp may only read the TMEM lane quarter
// warp % 4, so the second group halves every epilogue's work per thread instead of its rows); the last warp issues every
// tcgen05.mma of the CTA -- converged, one elected lane per instruction --, serving whichever slot has its next step ready, so
// that one slot's epilogue (tcgen05.ld, ReLU / loss / ReLU', fp16 stores) runs under the other slot's MMAs.  Hand-over is by
// mbarriers: `ready[s]` (256 arrivals: operands of the slot's next step are in shared memory and its TMEM reads are done), `done[s]`
// (tcgen05.commit of the step's forward / data-gradient MMAs) and `wdone[s]` (commit of its weight-gradient MMAs).  A tile is 10
// steps for the shipped network: 5 forward layers, then per layer, last to first, the data-gradient + the weight-gradient MMAs.
constexpr uint32_t kTrainSlotThreads = 2u * kTileRows;
template <int kSlots>
__global__ void __launch_bounds__(kSlots *kTrainSlotThreads + 32, 1)
k_train_mlp(const __grid_constant__ NetConst nc, const __grid_constant__ TrainParams tp)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t n_hidden = nc.n_layers - 1u;
    const TrainSmem lay     = train_smem_layout(nc.wimg_bytes, nc.in_k, n_hidden, kSlots);
    const uint32_t tid      = threadIdx.x;
    const uint32_t warp     = tid >> 5, lane = tid & 31u;
    const bool is_issuer    = warp == (uint32_t)(kSlots * 8);
    const uint32_t slot     = is_issuer ? 0u : tid / kTrainSlotThreads;
    const uint32_t row      = tid % kTileRows;
    const uint32_t half     = (tid / kTileRows) & 1u; // column half of the epilogues
    const uint32_t bar0     = smem_u32(smem + lay.bar_off);
    const uint32_t bar_w    = bar0;
    auto bar_feat  = [&](uint32_t sl) { return bar0 + 8u + 32u * sl; };
    auto bar_ready = [&](uint32_t sl) { return bar0 + 16u + 32u * sl; };
    auto bar_done  = [&](uint32_t sl) { return bar0 + 24u + 32u * sl; };
    auto bar_wdone = [&](uint32_t sl) { return bar0 + 32u + 32u * sl; };

    if (tid == 0) {
        mbar_init(bar_w, 1u);
        for (uint32_t sl = 0; sl < (uint32_t)kSlots; ++sl) {
            mbar_init(bar_feat(sl), 1u);
            mbar_init(bar_ready(sl), kTrainSlotThreads);
            mbar_init(bar_done(sl), 1u);
            mbar_init(bar_wdone(sl), 1u);
        }
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar_w, nc.wimg_bytes);
        tma_bulk_g2s(smem_u32(smem + lay.w_off), nc.wimg, nc.wimg_bytes, bar_w);
    }
    if (warp == 0)
        tmem_alloc(smem_u32(smem + lay.tmem_off), 512u);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(smem + lay.tmem_off);
    mbar_wait(bar_w, 0u);

    const uint32_t w_smem     = smem_u32(smem + lay.w_off);
    const uint32_t tile_bytes = kTileRows * nc.in_k * 2u;
    const uint32_t h_bytes    = kTileRows * kNeurons * 2u;
    const uint32_t rg_a0      = (nc.in_k >> 3) * 128u;            // row-group stride of the feature tile
    const uint32_t rg_h       = (uint32_t)(kNeurons >> 3) * 128u; // ... of the 64-wide tiles
    const uint32_t rg_o       = (uint32_t)(kOutPadded >> 3) * 128u;
    const uint32_t slot_cols  = train_slot_cols(nc.in_k);
    const uint32_t col_dw     = (uint32_t)kSlots * slot_cols;
    const uint32_t n_tiles    = (tp.batch + kTileRows - 1u) / kTileRows;
    // tile of (round, slot): blockIdx.x + (round * kSlots + slot) * gridDim.x
    const uint32_t my_tiles   = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1u) / gridDim.x : 0u;

    if (is_issuer) {
        // ================================================================ MMA issuer (the whole warp runs the loop; one elected lane issues)
        {
            // Event-driven: whichever slot has its next step ready is served (the steps are not equally long -- the loss epilogue --, and
            // strict alternation would make the faster slot wait for the slower one)
            uint32_t ready_phase[kSlots], feat_phase[kSlots], step_of[kSlots], round_of[kSlots];
            uint32_t remaining = my_tiles * 2u * nc.n_layers;
            for (int sl = 0; sl < kSlots; ++sl)
                ready_phase[sl] = feat_phase[sl] = step_of[sl] = round_of[sl] = 0u;
            uint32_t dw_started = 0u; // bit `layer`: the layer's weight-gradient accumulator holds a value
#ifdef NBVH_TRAIN_TIMING
            long long t_iss = 0, t_all0 = clock64();
#endif
            while (remaining) {
                {
#pragma unroll
                    for (int sl = 0; sl < kSlots; ++sl) {
                        if (round_of[sl] * (uint32_t)kSlots + (uint32_t)sl >= my_tiles)
                            continue;
                        const uint32_t step = step_of[sl], round = round_of[sl];
                        // step 0: the tile's features landed (TMA) and -- after the first tile -- the previous tile's last epilogue
                        // released the slot's TMEM columns; other steps: the slot's operands are in shared memory
                        if (step == 0u) {
                            if (!mbar_test(bar_feat((uint32_t)sl), feat_phase[sl]) || (round > 0u && !mbar_test(bar_ready((uint32_t)sl), ready_phase[sl])))
                                continue;
                            feat_phase[sl] ^= 1u;
                            if (round > 0u)
                                ready_phase[sl] ^= 1u;
                        } else {
                            if (!mbar_test(bar_ready((uint32_t)sl), ready_phase[sl]))
                                continue;
                            ready_phase[sl] ^= 1u;
                        }
#ifdef NBVH_TRAIN_TIMING
                        const long long t_i0 = clock64();
#endif
                        --remaining;
                        if (++step_of[sl] == 2u * nc.n_layers)
                            step_of[sl] = 0u, ++round_of[sl];
                        const uint32_t base    = smem_u32(smem + lay.slot_off + (uint32_t)sl * lay.slot_bytes);
                        const uint32_t a0_addr = base + lay.a0_off, h_addr = base + lay.h_off, dzo_addr = base + lay.dzo_off;
                        const uint32_t col_z   = (uint32_t)sl * slot_cols;
                        tc_fence_after_sync();
                        if (elect_one_sync()) {
                        // Descriptors advance by constant increments (16-byte units in the low word), so one MMA costs two 64-bit adds:
                        // this thread's instruction stream paces the whole CTA.
                        if (step < nc.n_layers) {
                            // ---- forward layer `step`: Z = A . W^T
                            const uint32_t layer  = step;
                            const uint32_t k_in   = layer == 0u ? nc.in_k : (uint32_t)kNeurons;
                            const uint32_t a_addr = layer == 0u ? a0_addr : h_addr + (layer - 1u) * h_bytes;
                            const uint32_t n_out  = layer + 1u == nc.n_layers ? (uint32_t)kOutPadded : (uint32_t)kNeurons;
                            const uint32_t sbo    = (k_in >> 3) * 128u;
                            const uint32_t idesc  = umma_idesc(128u, n_out, 0u, 0u);
                            uint64_t ad = umma_desc(a_addr, kLBO, sbo), bd = umma_desc(w_smem + nc.layer_off[layer], kLBO, sbo);
                            umma_f16(tmem + col_z, ad, bd, idesc, 0u);
                            for (uint32_t k = 16u; k < k_in; k += 16u) {
                                ad += 16u, bd += 16u; // two core matrices (256 bytes) along K
                                umma_f16(tmem + col_z, ad, bd, idesc, 1u);
                            }
                        } else {
                            // ---- backward of layer `layer`: dW_layer += dZ^T A_{layer-1};  dA_{layer-1} = dZ W_layer
                            const uint32_t layer   = 2u * nc.n_layers - 1u - step;
                            const bool last        = layer + 1u == nc.n_layers;
                            const uint32_t kin     = layer == 0u ? nc.in_k : (uint32_t)kNeurons;
                            const uint32_t in_addr = layer == 0u ? a0_addr : h_addr + (layer - 1u) * h_bytes;
                            const uint32_t in_rg   = layer == 0u ? rg_a0 : rg_h;
                            const uint32_t dz_addr = last ? dzo_addr : h_addr + layer * h_bytes; // dZ_l lives in H_l's tile
                            const uint32_t dz_rg   = last ? rg_o : rg_h;
                            const uint32_t dz_n    = last ? (uint32_t)kOutPadded : (uint32_t)kNeurons;
                            const uint32_t dw_col  = col_dw + (layer == 0u ? 0u : nc.in_k + (layer - 1u) * kNeurons);
                            const uint32_t fresh   = (dw_started >> layer) & 1u ? 0u : 1u;
                            dw_started |= 1u << layer;
                            // data gradient: D[128 x kin] = dZ (K-major, K = dz_n) . W_layer viewed MN-major (N = kin, K = out)
                            {
                                const uint32_t idesc = umma_idesc(128u, kin, 0u, 1u);
                                const uint32_t w_rg  = (kin >> 3) * 128u; // row-group (8 outputs) stride of the forward image
                                uint64_t ad = umma_desc(dz_addr, kLBO, dz_rg), bd = umma_desc_mn(w_smem + nc.layer_off[layer], w_rg, 128u);
                                const uint32_t b_inc = w_rg >> 3;
                                umma_f16(tmem + col_z, ad, bd, idesc, 0u);
                                for (uint32_t k = 16u; k < dz_n; k += 16u) {
                                    ad += 16u, bd += b_inc;
                                    umma_f16(tmem + col_z, ad, bd, idesc, 1u);
                                }
                            }
                            // the epilogue only needs dA: it may load and convert it while the weight-gradient MMAs run; it stores dZ
                            // (in place over the tile those MMAs read) after their own commit
                            umma_commit(bar_done((uint32_t)sl));
                            // weight gradient, K = the 128 samples of the tile (8 steps of 16 rows = two row groups), both operands MN-major:
                            //   hidden / first: D[64 (out) x kin] += dZ^T (M = 64) . A_{in} (N = kin)
                            //   last (transposed): D[64 (in) x 16 (out)] += A_{in}^T (M = 64) . dZ_out (N = 16)
                            {
                                const uint32_t idesc = last ? umma_idesc(64u, kOutPadded, 1u, 1u) : umma_idesc(64u, kin, 1u, 1u);
                                const uint64_t dzd = umma_desc_mn(dz_addr, dz_rg, 128u), ind = umma_desc_mn(in_addr, in_rg, 128u);
                                uint64_t ad = last ? ind : dzd, bd = last ? dzd : ind;
                                const uint32_t a_inc = (last ? in_rg : dz_rg) >> 3, b_inc = (last ? dz_rg : in_rg) >> 3; // 2 row groups / 16 bytes
                                umma_f16(tmem + dw_col, ad, bd, idesc, fresh ? 0u : 1u);
#pragma unroll
                                for (uint32_t ks = 1; ks < (uint32_t)kTileRows / 16u; ++ks) {
                                    ad += a_inc, bd += b_inc;
                                    umma_f16(tmem + dw_col, ad, bd, idesc, 1u);
                                }
                            }
                            umma_commit(bar_wdone((uint32_t)sl));
                        }
                        if (step < nc.n_layers)
                            umma_commit(bar_done((uint32_t)sl)); // implies tcgen05.fence::before_thread_sync
                        }
                        __syncwarp();
#ifdef NBVH_TRAIN_TIMING
                        t_iss += clock64() - t_i0;
#endif
                    }
                }
            }
#ifdef NBVH_TRAIN_TIMING
            if (blockIdx.x == 0 && lane == 0)
                printf("[k_train_mlp] issuer: %lld cycles total, %lld issuing (%u tiles)\n", clock64() - t_all0, t_iss, my_tiles);
#endif
        }
    } else {
        // ================================================================ epilogue warps of slot `slot`
        uint8_t *base            = smem + lay.slot_off + slot * lay.slot_bytes;
        uint8_t *a0_tile         = base + lay.a0_off;
        uint8_t *h_tiles         = base + lay.h_off;
        uint8_t *dzo_tile        = base + lay.dzo_off;
        const uint32_t col_z     = slot * slot_cols;
        const uint32_t lane_base = (row & ~31u) << 16;
        const float n_total_f    = (float)tp.n_total;
        const bool n_total_pow2  = tp.n_total != 0u && (tp.n_total & (tp.n_total - 1u)) == 0u;
        const float inv_n_total  = 1.f / n_total_f;
        uint32_t done_phase      = 0u, wdone_phase = 0u;
#ifdef NBVH_TRAIN_TIMING
        long long t_wait = 0, t_wwait = 0, t_loss = 0, t_ld = 0, t_cvt = 0, t_fence = 0, t_e0 = clock64();
#define TT_BEGIN const long long tt0_ = clock64();
#define TT_END(acc) acc += clock64() - tt0_;
#else
#define TT_BEGIN
#define TT_END(acc)
#endif
        float loss_acc           = 0.f;
        const bool lead = row == 0 && half == 0u; // the slot's TMA-issuing thread
        if (lead && slot < my_tiles) { // first tile of the slot
            mbar_expect_tx(bar_feat(slot), tile_bytes);
            tma_bulk_g2s(smem_u32(a0_tile), tp.feat + (size_t)(blockIdx.x + slot * gridDim.x) * tile_bytes, tile_bytes, bar_feat(slot));
        }
        const __half2 zero2 = __float2half2_rn(0.f);
        const uint32_t col_h = 32u * half; // this thread's columns of a 64-wide accumulator: [col_h, col_h + 32) = 16-byte chunks 4 half ..
        for (uint32_t round = 0; round * (uint32_t)kSlots + slot < my_tiles; ++round) {
            const uint32_t tile = blockIdx.x + (round * (uint32_t)kSlots + slot) * gridDim.x;
            const uint32_t s    = tile * kTileRows + row;
            const bool valid    = s < tp.batch;
            // the loss epilogue is split by output element: group `half` owns elements [4 half, 4 half + 4) of the row
            float tgt[4];
            const float tgt0 = (valid && nc.n_out > 0u) ? __ldg(tp.target + (size_t)s * nc.n_out) : 0.f;
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j)
                tgt[j] = (valid && 4u * half + j < nc.n_out) ? __ldg(tp.target + (size_t)s * nc.n_out + 4u * half + j) : 0.f;
            // ================= forward: activations of every layer stay in shared memory
            for (uint32_t layer = 0; layer < nc.n_layers; ++layer) {
                const bool last = layer + 1u == nc.n_layers;
                {
                    TT_BEGIN
                    mbar_wait(bar_done(slot), done_phase);
                    TT_END(t_wait)
                }
                done_phase ^= 1u;
                tc_fence_after_sync();
                if (!last) {
                    // ReLU after the fp16 rounding (same values: rounding is monotonic and keeps zero), two activations per instruction
                    uint8_t *h_tile = h_tiles + layer * h_bytes;
                    uint32_t v[32];
                    {
                        TT_BEGIN
                        tmem_ld32(tmem + col_z + lane_base + col_h, v);
                        tmem_ld_wait();
                        TT_END(t_ld)
                    }
                    TT_BEGIN
#pragma unroll
                    for (uint32_t c = 0; c < 4u; ++c) {
                        uint32_t h[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const __half2 p = __hmax2(__floats2half2_rn(__uint_as_float(v[8 * c + 2 * i]), __uint_as_float(v[8 * c + 2 * i + 1])), zero2);
                            h[i]            = *reinterpret_cast<const uint32_t *>(&p);
                        }
                        *reinterpret_cast<uint4 *>(h_tile + tile_chunk_off(row, 4u * half + c, kNeurons)) = make_uint4(h[0], h[1], h[2], h[3]);
                    }
                    TT_END(t_cvt)
                } else {
                    TT_BEGIN
                    uint32_t v[16];
                    tmem_ld16(tmem + col_z + lane_base, v);
                    tmem_ld_wait();
                    const uint32_t e0 = 4u * half;
                    float y[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) // the training forward hands fp16 predictions to the loss (trainer.h:108-131)
                        y[i] = h2f_round(__fdividef(1.0f, 1.0f + __expf(-__uint_as_float(half ? v[4 + i] : v[i]))));
                    // ---- loss + output-activation backward -> dZ_out tile (128 x 16 fp16, loss-scaled)
                    float dz[4];
                    float lrow = 0.f;
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j) {
                        const uint32_t e = e0 + j;
                        float g          = 0.f;
                        const float l    = (valid && e < nc.n_out)
                                               ? loss_element(tp.loss, tgt[j], tgt0, y[j], e, n_total_f, inv_n_total, n_total_pow2, tp.loss_scale, g)
                                               : 0.f;
                        if (tp.per_sample_loss && valid)
                            tp.per_sample_loss[(size_t)s * kOutPadded + e] = l;
                        lrow += l;
                        g     = valid ? h2f_round(g) : 0.f;
                        dz[j] = g * (y[j] * (1.f - y[j])); // sigmoid' from the output value
                    }
                    if (tp.per_sample_loss && valid) {
#pragma unroll
                        for (uint32_t j = 0; j < 4; ++j)
                            tp.per_sample_loss[(size_t)s * kOutPadded + 8u + e0 + j] = 0.f;
                    }
                    loss_acc += lrow;
                    const __half2 p0 = __floats2half2_rn(dz[0], dz[1]), p1 = __floats2half2_rn(dz[2], dz[3]);
                    *reinterpret_cast<uint2 *>(dzo_tile + tile_chunk_off(row, 0u, kOutPadded) + 8u * half) =
                        make_uint2(*reinterpret_cast<const uint32_t *>(&p0), *reinterpret_cast<const uint32_t *>(&p1));
                    *reinterpret_cast<uint2 *>(dzo_tile + tile_chunk_off(row, 1u, kOutPadded) + 8u * half) = make_uint2(0u, 0u);
                    TT_END(t_loss)
                }
                {
                    TT_BEGIN
                    fence_proxy_async_smem();
                    tc_fence_before_sync();
                    mbar_arrive(bar_ready(slot));
                    TT_END(t_fence)
                }
            }
            // ================= backward, layer = last .. 0: the issuer runs dA_{layer-1} (commit: done) and dW_layer (commit: wdone);
            // here dZ_{layer-1}
            for (int layer = (int)nc.n_layers - 1; layer >= 0; --layer) {
                {
                    TT_BEGIN
                    mbar_wait(bar_done(slot), done_phase);
                    TT_END(t_wait)
                }
                done_phase ^= 1u;
                tc_fence_after_sync();
                if (layer > 0) {
                    // dZ_{layer-1} = dA . ReLU'(H_{layer-1}) -> fp16: computed while the weight-gradient MMAs (which read H_{layer-1}) run,
                    // stored in place over H_{layer-1} once they have completed
                    uint8_t *h_tile = h_tiles + (uint32_t)(layer - 1) * h_bytes;
                    uint32_t v[32];
                    tmem_ld32(tmem + col_z + lane_base + col_h, v);
                    tmem_ld_wait();
                    uint32_t h[4][4];
#pragma unroll
                    for (uint32_t c = 0; c < 4u; ++c) {
                        const uint4 m = *reinterpret_cast<const uint4 *>(h_tile + tile_chunk_off(row, 4u * half + c, kNeurons));
                        const uint32_t mk[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            // stored activation > 0 (per half: all-ones / zero mask) selects the rounded gradient
                            const __half2 p = __floats2half2_rn(__uint_as_float(v[8 * c + 2 * i]), __uint_as_float(v[8 * c + 2 * i + 1]));
                            h[c][i] = *reinterpret_cast<const uint32_t *>(&p) & __hgt2_mask(*reinterpret_cast<const __half2 *>(&mk[i]), zero2);
                        }
                    }
                    {
                        TT_BEGIN
                        mbar_wait(bar_wdone(slot), wdone_phase);
                        TT_END(t_wwait)
                    }
                    wdone_phase ^= 1u;
#pragma unroll
                    for (uint32_t c = 0; c < 4u; ++c)
                        *reinterpret_cast<uint4 *>(h_tile + tile_chunk_off(row, 4u * half + c, kNeurons)) = make_uint4(h[c][0], h[c][1], h[c][2], h[c][3]);
                    fence_proxy_async_smem();
                } else {
                    // the feature tile is free once dW_0 has completed: fetch the slot's next tile under this epilogue
                    const uint32_t next_idx = (round + 1u) * (uint32_t)kSlots + slot;
                    if (lead && next_idx < my_tiles) {
                        mbar_wait(bar_wdone(slot), wdone_phase);
                        mbar_expect_tx(bar_feat(slot), tile_bytes);
                        tma_bulk_g2s(smem_u32(a0_tile), tp.feat + (size_t)(blockIdx.x + next_idx * gridDim.x) * tile_bytes, tile_bytes, bar_feat(slot));
                    }
                    wdone_phase ^= 1u;
                    // dL/dfeatures -> fp16 -> global, same tile-canonical layout as the features; the 16-column chunks alternate between
                    // the two warp groups
                    uint8_t *dst = tp.dfeat + (size_t)tile * tile_bytes;
                    for (uint32_t c0 = 16u * half; c0 < nc.in_k; c0 += 32u) {
                        uint32_t v[16];
                        tmem_ld16(tmem + col_z + lane_base + c0, v);
                        tmem_ld_wait();
                        uint32_t h[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const __half2 p = __floats2half2_rn(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                            h[i]            = *reinterpret_cast<const uint32_t *>(&p);
                        }
                        *reinterpret_cast<uint4 *>(dst + tile_chunk_off(row, c0 >> 3, nc.in_k))        = make_uint4(h[0], h[1], h[2], h[3]);
                        *reinterpret_cast<uint4 *>(dst + tile_chunk_off(row, (c0 >> 3) + 1u, nc.in_k)) = make_uint4(h[4], h[5], h[6], h[7]);
                    }
                }
                tc_fence_before_sync();
                mbar_arrive(bar_ready(slot));
            }
        }
        if (my_tiles > slot) // the last tile's weight-gradient MMAs (flushed below)
            mbar_wait(bar_wdone(slot), wdone_phase ^ 1u);
#ifdef NBVH_TRAIN_TIMING
        if (blockIdx.x == 0 && lead)
            printf("[k_train_mlp] slot %u epilogue: %lld cycles total, %lld waiting for MMAs, %lld for weight-gradient MMAs; forward: tmem ld %lld, cvt+store %lld, loss %lld, fence+arrive %lld\n", slot, clock64() - t_e0,
                   t_wait, t_wwait, t_ld, t_cvt, t_loss, t_fence);
#endif
        loss_acc = warp_sum(loss_acc);
        if (lane == 0 && tp.loss_sum && loss_acc != 0.f)
            atomicAdd(tp.loss_sum, loss_acc);
    }

    // ================= flush the TMEM-resident weight gradients (every commit has been waited on by its slot's epilogue warps)
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (!is_issuer && my_tiles > 0u) {
        // M = 64 accumulators: row m lives in TMEM lane (m % 16) + 32 * (m / 16)  ->  warp w (mod 4), lanes 0..15 hold rows 16 w .. 16 w + 15;
        // the 16-column chunks are dealt round-robin to the warp groups
        const uint32_t lane_base = (row & ~31u) << 16;
        const uint32_t r         = 16u * (warp & 3u) + lane;
        uint32_t col             = col_dw;
        size_t goff              = 0;
        uint32_t chunk           = 0;
        for (uint32_t layer = 0; layer < nc.n_layers; ++layer) {
            const bool last    = layer + 1u == nc.n_layers;
            const uint32_t kin = layer == 0 ? nc.in_k : (uint32_t)kNeurons;
            const uint32_t nn  = last ? (uint32_t)kOutPadded : kin; // accumulator columns
            for (uint32_t c0 = 0; c0 < nn; c0 += 16u, ++chunk) {
                if (chunk % (2u * (uint32_t)kSlots) != 2u * slot + half)
                    continue;
                uint32_t v[16];
                tmem_ld16(tmem + col + lane_base + c0, v);
                tmem_ld_wait();
                if (lane < 16u) {
                    // hidden/first: D[out = r][in = c0 + i]; last (transposed): D[in = r][out = c0 + i]
                    if (tp.wgrad_partial) {
                        float *dst = tp.wgrad_partial + (size_t)blockIdx.x * tp.n_matrix;
                        if (last) {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                dst[goff + (size_t)(c0 + i) * kNeurons + r] = __uint_as_float(v[i]);
                        } else {
                            float4 *d4 = reinterpret_cast<float4 *>(dst + goff + (size_t)r * kin + c0);
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                d4[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                                    __uint_as_float(v[4 * i + 3]));
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float g = __uint_as_float(v[i]);
                            if (g != 0.f) {
                                const size_t idx = last ? goff + (size_t)(c0 + i) * kNeurons + r : goff + (size_t)r * kin + c0 + i;
                                atomicAdd(tp.wgrad + idx, g);
                            }
                        }
                    }
                }
            }
            col += nn;
            goff += (size_t)(last ? kOutPadded : kNeurons) * kin;
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after_sync();
        tmem_dealloc(tmem, 512u);
    }
}

// wgrad[i] = sum over the CTAs of k_train_mlp of their partial sums, in CTA order (deterministic)
// (64-thread blocks: 304 CTAs for the 19,456 weights of the shipped network instead of 76, and 16 independent loads in flight per
// thread before the adds -- the additions keep their order, so the sums keep their bits; 24.8 -> measured in profiles/r2_s60)
__global__ void __launch_bounds__(64)
k_reduce_wgrad(const float *__restrict__ partial, uint32_t n_parts, uint32_t n, float *__restrict__ wgrad)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    float acc  = 0.f;
    uint32_t p = 0;
    for (; p + 16u <= n_parts; p += 16u) {
        float v[16];
#pragma unroll
        for (uint32_t k = 0; k < 16u; ++k)
            v[k] = __ldcs(partial + (size_t)(p + k) * n + i);
#pragma unroll
        for (uint32_t k = 0; k < 16u; ++k)
            acc += v[k];
    }
    for (; p < n_parts; ++p)
        acc += __ldcs(partial + (size_t)p * n + i);
    wgrad[i] = acc;
}

// ------------------------------------------------------------------------------------------------ k_grid_backward
// One thread per (sample, point) of ONE level (blockIdx.y), so a block's scatter targets one level's table.
// use_smem: the level's gradient table (size x 4 floats) is privatised in shared memory and flushed once per block.
__global__ void __launch_bounds__(256)
k_grid_backward(const __grid_constant__ NetConst nc, const float *__restrict__ in, const uint8_t *__restrict__ dfeat, uint32_t batch,
                uint32_t n_in_dims, uint32_t first_level, uint32_t use_smem, float *__restrict__ grid_grads)
{
    extern __shared__ float sg[];
    const uint32_t level = first_level + blockIdx.y;
    const GridLevel &L   = nc.lv[level];
    if (use_smem) {
        for (uint32_t i = threadIdx.x; i < L.size * 4u; i += blockDim.x)
            sg[i] = 0.f;
        __syncthreads();
    }
    const uint64_t total = (uint64_t)batch * nc.n_points;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t p = (uint32_t)(i / batch), s = (uint32_t)(i % batch); // consecutive threads -> consecutive samples
        const uint32_t j = p * nc.n_levels + level;
        const uint2 raw  = __ldg(reinterpret_cast<const uint2 *>(dfeat + feat_slot_offset(s, j, nc.in_k)));
        if ((raw.x | raw.y) == 0u)
            continue; // masked sample (target not visible) or dead ReLU path: nothing to scatter
        const float2 g01 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x));
        const float2 g23 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
        const float *x   = in + (size_t)s * n_in_dims + 3u * p;
        GridCorner gc;
        grid_corners(L, __ldg(x), __ldg(x + 1), __ldg(x + 2), gc);
        if (use_smem) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float *dst = sg + 4u * gc.index[c];
                atomicAdd(dst + 0, gc.weight[c] * g01.x);
                atomicAdd(dst + 1, gc.weight[c] * g01.y);
                atomicAdd(dst + 2, gc.weight[c] * g23.x);
                atomicAdd(dst + 3, gc.weight[c] * g23.y);
            }
        } else {
            float *base = grid_grads + 4u * (size_t)L.offset;
#pragma unroll
            for (int c = 0; c < 8; ++c)
                red_add_v4(base + 4u * (size_t)gc.index[c], gc.weight[c] * g01.x, gc.weight[c] * g01.y, gc.weight[c] * g23.x,
                           gc.weight[c] * g23.y);
        }
    }
    if (use_smem) {
        __syncthreads();
        float *base = grid_grads + 4u * (size_t)L.offset;
        for (uint32_t e = threadIdx.x; e < L.size; e += blockDim.x) {
            const float4 v = *reinterpret_cast<const float4 *>(sg + 4u * e);
            if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f)
                red_add_v4(base + 4u * (size_t)e, v.x, v.y, v.z, v.w);
        }
    }
}

// The privatised (dense, coarse) levels with one thread per SAMPLE: the points of a segment mostly fall into the same cell of a coarse
// level (cell edge 1/8 or 1/16 of the scene box against a leaf-sized segment), so their 8 x 4 corner contributions are summed in
// registers and flushed once per distinct cell -- up to n_points times fewer shared-memory float atomics, which are compare-and-swap
// loops on sm_100 (ATOMS.CAST.SPIN) and bound these launches (85 % L1TEX, profiles/r2_s59_full_ncu.txt).
__global__ void __launch_bounds__(256)
k_grid_backward_merge(const __grid_constant__ NetConst nc, const float *__restrict__ in, const uint8_t *__restrict__ dfeat, uint32_t batch,
                      uint32_t n_in_dims, uint32_t first_level, uint32_t use_smem, float *__restrict__ grid_grads)
{
    extern __shared__ float sg[];
    const uint32_t level = first_level + blockIdx.y;
    const GridLevel &L   = nc.lv[level];
    float *gbase         = grid_grads + 4u * (size_t)L.offset;
    if (use_smem) {
        for (uint32_t i = threadIdx.x; i < L.size * 4u; i += blockDim.x)
            sg[i] = 0.f;
        __syncthreads();
    }
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < batch; s += gridDim.x * blockDim.x) {
        float acc[8][4];
        uint32_t idx[8];
        bool have = false;
        auto flush = [&]() {
            if (use_smem) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float *dst = sg + 4u * idx[c];
                    atomicAdd(dst + 0, acc[c][0]);
                    atomicAdd(dst + 1, acc[c][1]);
                    atomicAdd(dst + 2, acc[c][2]);
                    atomicAdd(dst + 3, acc[c][3]);
                }
            } else {
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    red_add_v4(gbase + 4u * (size_t)idx[c], acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
            }
        };
        for (uint32_t p = 0; p < nc.n_points; ++p) {
            const uint32_t j = p * nc.n_levels + level;
            const uint2 raw  = __ldg(reinterpret_cast<const uint2 *>(dfeat + feat_slot_offset(s, j, nc.in_k)));
            if ((raw.x | raw.y) == 0u)
                continue; // masked sample (target not visible) or dead ReLU path: nothing to scatter
            const float2 g01 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x));
            const float2 g23 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
            const float *x   = in + (size_t)s * n_in_dims + 3u * p;
            GridCorner gc;
            grid_corners(L, __ldg(x), __ldg(x + 1), __ldg(x + 2), gc);
            bool same = have;
#pragma unroll
            for (int c = 0; c < 8; ++c)
                same = same && gc.index[c] == idx[c];
            if (!same) {
                if (have)
                    flush();
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    idx[c]    = gc.index[c];
                    acc[c][0] = gc.weight[c] * g01.x, acc[c][1] = gc.weight[c] * g01.y;
                    acc[c][2] = gc.weight[c] * g23.x, acc[c][3] = gc.weight[c] * g23.y;
                }
                have = true;
            } else {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    acc[c][0] += gc.weight[c] * g01.x, acc[c][1] += gc.weight[c] * g01.y;
                    acc[c][2] += gc.weight[c] * g23.x, acc[c][3] += gc.weight[c] * g23.y;
                }
            }
        }
        if (have)
            flush();
    }
    if (!use_smem)
        return;
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < L.size; e += blockDim.x) {
        const float4 v = *reinterpret_cast<const float4 *>(sg + 4u * e);
        if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f)
            red_add_v4(gbase + 4u * (size_t)e, v.x, v.y, v.z, v.w);
    }
}

// ------------------------------------------------------------------------------------------------ k_adam
// optimizers/adam.h:48-123 (relative/absolute decay 0, no clipping, AdaBound off).  lr already carries the
// ExponentialDecay factor.  Also writes the fp16 parameter and, for matrix weights, its slot in the packed weight image.
__global__ void __launch_bounds__(256)
k_adam(const __grid_constant__ AdamParams ap)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ap.n_params)
        return;
    // tcnn keeps gradients in fp16: round here so underflow-to-zero (and hence the skip below) is reproduced
    float gradient = __half2float(__float2half_rn(ap.grads[i])) / ap.loss_scale;
    const bool matrix = i < ap.n_matrix;
    if (!matrix && gradient == 0.f)
        return;
    const float w = ap.master[i];
    if (matrix)
        gradient += ap.l2_reg * w;
    const float g2 = gradient * gradient;
    const float m  = ap.m[i] = ap.beta1 * ap.m[i] + (1.f - ap.beta1) * gradient;
    const float v  = ap.v[i] = ap.beta2 * ap.v[i] + (1.f - ap.beta2) * g2;
    const uint32_t t = ++ap.steps[i];
    const float lr   = ap.lr * (sqrtf(1.f - powf(ap.beta2, (float)t)) / (1.f - powf(ap.beta1, (float)t)));
    const float nw   = w - (lr / (sqrtf(v) + ap.epsilon)) * m;
    ap.master[i]     = nw;
    const __half hw  = __float2half_rn(nw);
    ap.params[i]     = hw;
    if (matrix) {
        // (layer, r, c) of matrix weight i -> canonical K-major slot of the tensor-core weight image
        uint32_t rem = (uint32_t)i, layer = 0, k = ap.in_k, n = kNeurons;
        for (;; ++layer) {
            k = layer == 0 ? ap.in_k : (uint32_t)kNeurons;
            n = layer + 1u == ap.n_layers ? (uint32_t)kOutPadded : (uint32_t)kNeurons;
            if (rem < n * k)
                break;
            rem -= n * k;
        }
        const uint32_t r = rem / k, c = rem % k;
        const uint32_t off = ap.layer_off[layer] + (r >> 3) * ((k >> 3) * 128u) + (c >> 3) * 128u + (r & 7u) * 16u + (c & 7u) * 2u;
        *reinterpret_cast<__half *>(ap.wimg + off) = hw;
    }
}

__global__ void k_half_to_float(const __half *__restrict__ src, float *__restrict__ dst, uint64_t n)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        dst[i] = __half2float(src[i]);
}

// ------------------------------------------------------------------------------------------------ launchers
cudaError_t configure_training(uint32_t wimg_bytes, uint32_t in_k, uint32_t n_layers)
{
    (void)wimg_bytes, (void)in_k, (void)n_layers;
    cudaError_t e = allow_max_dynamic_smem(reinterpret_cast<const void *>(k_train_mlp<1>));
    if (e != cudaSuccess)
        return e;
    if ((e = allow_max_dynamic_smem(reinterpret_cast<const void *>(k_train_mlp<2>))) != cudaSuccess)
        return e;
    if ((e = cudaFuncSetAttribute(k_grid_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)) != cudaSuccess)
        return e;
    return cudaFuncSetAttribute(k_grid_backward_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
}

cudaError_t launch_encode_features(const NetConst &nc, const float *in, uint32_t batch, uint32_t n_in_dims, uint8_t *feat, int num_sms,
                                   cudaStream_t stream)
{
    if (batch == 0)
        return cudaSuccess;
    const uint32_t padded = (batch + kTileRows - 1u) / kTileRows * kTileRows;
    const uint64_t cap    = (uint64_t)num_sms * 32u;
    // measured at 2^18 on B200 (profiles/r2_s65): 0.106 -> 0.101 ms; NBVH_ENCODE_MERGE=0 = one thread per (sample, point, level)
    static const bool merge = [] { const char *e = getenv("NBVH_ENCODE_MERGE"); return !(e && e[0] == '0'); }();
    if (merge && (nc.in_k >> 2) == nc.n_points * nc.n_levels) {
        uint64_t blocks = ((uint64_t)padded * nc.n_levels + 255u) / 256u;
        if (blocks > cap)
            blocks = cap;
        k_encode_features_merge<<<(uint32_t)blocks, 256, 0, stream>>>(nc, in, batch, padded, n_in_dims, feat);
        return cudaGetLastError();
    }
    const uint64_t total  = (uint64_t)padded * (nc.in_k >> 2);
    uint64_t blocks       = (total + 255u) / 256u;
    if (blocks > cap)
        blocks = cap;
    k_encode_features<<<(uint32_t)blocks, 256, 0, stream>>>(nc, in, batch, padded, n_in_dims, feat);
    return cudaGetLastError();
}

// Two tiles in flight per CTA when the network's tiles fit shared memory (227 KB) and TMEM (512 columns) twice, else one.
uint32_t train_mlp_slots(const NetConst &nc)
{
    const bool two = train_smem_layout(nc.wimg_bytes, nc.in_k, nc.n_layers - 1u, 2u).total <= 227u * 1024u &&
                     train_tmem_cols(nc.in_k, nc.n_layers, 2u) <= 512u;
    static const bool force_one = [] { const char *e = getenv("NBVH_TRAIN_SLOTS"); return e && e[0] == '1'; }(); // A/B measurements
    return two && !force_one ? 2u : 1u;
}

cudaError_t launch_train_mlp(const NetConst &nc, const TrainParams &tp, int num_sms, cudaStream_t stream)
{
    if (tp.batch == 0)
        return cudaSuccess;
    const uint32_t n_tiles = (tp.batch + kTileRows - 1u) / kTileRows;
    const uint32_t slots   = train_mlp_slots(nc);
    const uint32_t want    = (n_tiles + slots - 1u) / slots;
    const uint32_t grid    = want < (uint32_t)num_sms ? want : (uint32_t)num_sms;
    const uint32_t smem    = train_smem_layout(nc.wimg_bytes, nc.in_k, nc.n_layers - 1u, slots).total;
    if (slots == 2u)
        k_train_mlp<2><<<grid, 2 * kTrainSlotThreads + 32, smem, stream>>>(nc, tp);
    else
        k_train_mlp<1><<<grid, kTrainSlotThreads + 32, smem, stream>>>(nc, tp);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || !tp.wgrad_partial)
        return e;
    k_reduce_wgrad<<<(tp.n_matrix + 63u) / 64u, 64, 0, stream>>>(tp.wgrad_partial, grid, tp.n_matrix, tp.wgrad);
    return cudaGetLastError();
}

cudaError_t launch_grid_backward(const NetConst &nc, const float *in, const uint8_t *dfeat, uint32_t batch, uint32_t n_in_dims,
                                 float *grid_grads, int num_sms, cudaStream_t stream)
{
    if (batch == 0)
        return cudaSuccess;
    static const bool per_point_all = [] { const char *e = getenv("NBVH_GRID_BACKWARD_PER_POINT"); return e && e[0] == '1'; }(); // A/B measurements
    // leading levels small enough for a 64 KB shared-memory table
    uint32_t n_small = 0;
    while (n_small < nc.n_levels && nc.lv[n_small].size * 16u <= 64u * 1024u)
        ++n_small;
    const uint64_t total = (uint64_t)batch * nc.n_points;
    // one launch per privatised level with exactly its table in shared memory: the small tables leave room for many resident CTAs,
    // whose warps hide the latency of the shared-memory atomics (one flush of the table per CTA)
    for (uint32_t l = 0; l < n_small; ++l) {
        const uint32_t smem = nc.lv[l].size * 16u;
        uint32_t per_sm     = (200u * 1024u) / (smem + 1024u);
        per_sm              = per_sm > 8u ? 8u : per_sm < 1u ? 1u : per_sm;
        const uint32_t cap  = (uint32_t)num_sms * per_sm;
        if (!per_point_all) {
            uint32_t bx = (batch + 255u) / 256u;
            if (bx > cap)
                bx = cap;
            k_grid_backward_merge<<<bx, 256, smem, stream>>>(nc, in, dfeat, batch, n_in_dims, l, 1u, grid_grads);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess)
                return e;
            continue;
        }
        uint32_t bx         = (uint32_t)((total + 255u) / 256u);
        if (bx > cap)
            bx = cap;
        k_grid_backward<<<dim3(bx, 1), 256, smem, stream>>>(nc, in, dfeat, batch, n_in_dims, l, 1u, grid_grads);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess)
            return e;
    }
    // the other levels: the same per-sample kernel with 16-byte reductions to L2 (measured at 2^18 on B200, profiles/r2_s63: all six
    // levels through it 0.247 ms for the whole backward pass, none 0.287 ms -- merging still pays on levels 2 and 3, and on the hashed
    // levels one thread per sample issues its three points' reductions back to back); NBVH_GRID_BACKWARD_PER_POINT=1 = the old mapping
    if (!per_point_all && n_small < nc.n_levels) {
        const uint32_t k = nc.n_levels - n_small;
        uint32_t bx = (batch + 255u) / 256u;
        const uint32_t cap = (uint32_t)num_sms * 8u;
        if (bx > cap)
            bx = cap;
        k_grid_backward_merge<<<dim3(bx, k), 256, 0, stream>>>(nc, in, dfeat, batch, n_in_dims, n_small, 0u, grid_grads);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess)
            return e;
        n_small += k;
    }
    if (n_small < nc.n_levels) {
        uint32_t bx = (uint32_t)((total + 255u) / 256u);
        const uint32_t cap = (uint32_t)num_sms * 8u;
        if (bx > cap)
            bx = cap;
        k_grid_backward<<<dim3(bx, nc.n_levels - n_small), 256, 0, stream>>>(nc, in, dfeat, batch, n_in_dims, n_small, 0u, grid_grads);
        return cudaGetLastError();
    }
    return cudaSuccess;
}

cudaError_t launch_adam(const AdamParams &ap, cudaStream_t stream)
{
    if (ap.n_params == 0)
        return cudaSuccess;
    k_adam<<<(uint32_t)((ap.n_params + 255u) / 256u), 256, 0, stream>>>(ap);
    return cudaGetLastError();
}

cudaError_t launch_half_to_float(const void *src, float *dst, uint64_t n, cudaStream_t stream)
{
    if (n == 0)
        return cudaSuccess;
    k_half_to_float<<<(uint32_t)((n + 255u) / 256u), 256, 0, stream>>>(reinterpret_cast<const __half *>(src), dst, n);
    return cudaGetLastError();
}

} // namespace nbvh

// mlp_tile.cuh -- the network half of the hot path on Blackwell tensor cores.
//
// One "tile" is 128 queries (ray x leaf): thread r of the tile owns query r, which is row r of the A operand
// and TMEM lane r of the accumulator.  Per tile and per layer:
//     features / activations (fp16, K-major core-matrix layout in shared memory)  --tcgen05.mma-->
//     fp32 accumulators in TMEM  --tcgen05.ld-->  registers --ReLU, cvt fp16--> shared memory (next A)
// The weights are TMA-bulk-copied once per CTA into shared memory in the same canonical layout
// (pre-packed on the host).  Replaces tcnn's kernel_grid x3 + kernel_mlp_fused + trim_and_cast
// (/root/reference/ext/tiny-cuda-nn/include/tiny-cuda-nn/encodings/grid.h:49-212,
//  /root/reference/ext/tiny-cuda-nn/src/fully_fused_mlp.cu:500-557, src/object.cu:61-65).
//
// Canonical K-major no-swizzle operand layout (UMMA "INTERLEAVE"): 8 rows x 16 bytes core matrices;
//   byte(row, k) = (row/8)*SBO + (k/8)*LBO + (row%8)*16 + (k%8)*2,  LBO = 128, SBO = (K/8)*128.
#pragma once
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace nbvh {

constexpr int kTileRows     = 128;
constexpr int kMaxLevels    = 16;
constexpr int kMaxLayers    = 9;   // n_hidden_layers (<= 8) + output layer
constexpr int kNeurons      = 64;  // FullyFusedMLP width of every shipped config
constexpr int kOutPadded    = 16;  // padded output width (tcnn pads 8 -> 16)
constexpr int kMaxInputK    = 128; // widest first-layer K supported by the A tile
constexpr uint32_t kLBO     = 128;

struct GridLevel {
    uint32_t offset;   // first entry of the level (entries of n_features halves)
    uint32_t size;     // hashmap_size (entries)
    uint32_t mask;     // size-1 when size is a power of two, else 0
    uint32_t res;      // grid_resolution
    uint32_t dense;    // 1: x + y*res + z*res^2, 0: coherent prime hash
    float scale;       // grid_scale
};

struct NetConst {
    const uint2 *grid;          // fp16 x 4 per entry (n_features_per_level == 4)
    GridLevel lv[kMaxLevels];
    uint32_t n_levels;
    uint32_t n_points;          // segment points (3)
    uint32_t in_k;              // n_points * n_levels * 4 (multiple of 16)
    uint32_t n_layers;          // hidden layers + 1
    uint32_t n_out;             // 8
    const uint8_t *wimg;        // packed weights image (canonical layout per layer), 16-byte aligned
    uint32_t wimg_bytes;
    uint32_t layer_off[kMaxLayers]; // byte offset of each layer in the image
};

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n"
                 ".reg .pred P1;\n"
                 "LAB_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                 "@P1 bra DONE;\n"
                 "bra LAB_WAIT;\n"
                 "DONE:\n"
                 "}" ::"r"(bar),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// generic-proxy writes (st.shared) -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// TMA bulk copy global -> shared (UBLKCP), completion on an mbarrier
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// UMMA shared-memory descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);       // start address  [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;  // leading byte offset [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;  // stride byte offset [32,46)
    d |= (uint64_t)1 << 46;                             // descriptor version 1 (Blackwell)
    return d;                                           // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE (0)
}

// Instruction descriptor for kind::f16: A,B = fp16 K-major, D = fp32 (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t umma_idesc_f16_f32(uint32_t m, uint32_t n)
{
    return (1u << 4) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "setp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
                 "}" ::"r"(tmem_d),
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr)
                 : "memory");
}
// 32 accumulator columns of this thread's TMEM lane in one instruction
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr)
                 : "memory");
}
// elects one lane of a converged warp (the others skip the predicated instruction stream)
__device__ __forceinline__ bool elect_one_sync()
{
    uint32_t pred;
    asm volatile("{\n"
                 ".reg .pred P;\n"
                 "elect.sync _|P, 0xFFFFFFFF;\n"
                 "selp.u32 %0, 1, 0, P;\n"
                 "}"
                 : "=r"(pred));
    return pred != 0u;
}
// 64 accumulator columns of this thread's TMEM lane in one instruction (the epilogues of the training kernel)
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&v)[64])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
                 "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]),
                   "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]),
                   "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]),
                   "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------------------ hash grid (a9)
// index % size: every hashed level has a power-of-two table (mask); dense levels only wrap when a coordinate sits on the
// upper face of the unit cube, so the true modulo is a rare out-of-line call instead of ~25 instructions per corner
static __device__ __noinline__ uint32_t grid_wrap_slow(uint32_t index, uint32_t size) { return index % size; }
__device__ __forceinline__ uint32_t grid_wrap(uint32_t index, uint32_t mask, uint32_t size)
{
    if (mask)
        return index & mask;
    return index < size ? index : grid_wrap_slow(index, size);
}

// (Measured on B200, gpurun s32: `ld.global.nc.L1::no_allocate` for the hashed levels' entries costs 13 % of the 1080p k_neural launch,
// for all levels 16 % -- the gathers of coherent rays do hit L1 on every level; L1::evict_last on the tree-cut node loads: +-1 %.)
// One level of tcnn's kernel_grid for one point, arithmetic identical to the reference: positions through
// fma(scale,x,0.5), weights as products in fp32 rounded to fp16, accumulation with __hfma2 in corner order
// 0..7 (grid.h:94-169; pos_fract common_device.h:857-868; grid_index :691-707; coherent prime hash :652-655).
__device__ __forceinline__ uint2 grid_level_features(const uint2 *__restrict__ grid, const GridLevel &L, float x, float y,
                                                     float z)
{
    float px = __fmaf_rn(L.scale, x, 0.5f), py = __fmaf_rn(L.scale, y, 0.5f), pz = __fmaf_rn(L.scale, z, 0.5f);
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    const uint32_t gx = (uint32_t)(int)fx, gy = (uint32_t)(int)fy, gz = (uint32_t)(int)fz;
    px = __fsub_rn(px, fx), py = __fsub_rn(py, fy), pz = __fsub_rn(pz, fz);
    const float qx = __fsub_rn(1.f, px), qy = __fsub_rn(1.f, py), qz = __fsub_rn(1.f, pz);

    uint2 g[8];
#ifdef NBVH_GRID_PAIRED
    // The two x-neighbours of a (y,z) corner pair are adjacent table entries on dense levels (index+1) and, when gx is
    // even, on hashed levels too ((gx^h) and (gx^h)^1): fetch the aligned 16-byte entry pair with one LDG.128 and only
    // issue the second gather when the neighbour lives elsewhere -> ~25 % fewer L1 lookups per query.
    const uint4 *__restrict__ pairs = reinterpret_cast<const uint4 *>(grid + L.offset); // level offsets are multiples of 8 entries
#pragma unroll
    for (int c2 = 0; c2 < 4; ++c2) {
        const uint32_t cy = gy + (c2 & 1), cz = gz + (c2 >> 1);
        uint32_t i0, i1;
        if (L.dense) {
            i0 = gx + cy * L.res + cz * L.res * L.res;
            i1 = i0 + 1u;
        } else {
            const uint32_t h = (cy * 2654435761u) ^ (cz * 805459861u);
            i0 = gx ^ h, i1 = (gx + 1u) ^ h;
        }
        i0 = grid_wrap(i0, L.mask, L.size), i1 = grid_wrap(i1, L.mask, L.size);
        const uint4 pr  = __ldg(pairs + (i0 >> 1));
        const bool odd  = i0 & 1u;
        g[2 * c2]       = odd ? make_uint2(pr.z, pr.w) : make_uint2(pr.x, pr.y);
        if ((i0 ^ i1) == 1u)
            g[2 * c2 + 1] = odd ? make_uint2(pr.x, pr.y) : make_uint2(pr.z, pr.w);
        else
            g[2 * c2 + 1] = __ldg(grid + L.offset + i1);
    }
#else
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint32_t cx = gx + (c & 1), cy = gy + ((c >> 1) & 1), cz = gz + ((c >> 2) & 1);
        uint32_t index;
        if (L.dense)
            index = cx + cy * L.res + cz * L.res * L.res;
        else
            index = cx ^ (cy * 2654435761u) ^ (cz * 805459861u);
        index = grid_wrap(index, L.mask, L.size);
        g[c]  = __ldg(grid + L.offset + index);
    }
#endif
    __half2 r01 = __float2half2_rn(0.f), r23 = __float2half2_rn(0.f);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float w = (c & 1) ? px : qx;             // weight = 1 * w0 * w1 * w2, left to right
        w       = __fmul_rn(w, (c & 2) ? py : qy);
        w       = __fmul_rn(w, (c & 4) ? pz : qz);
        const __half2 w2 = __float2half2_rn(w);
        r01 = __hfma2(w2, *reinterpret_cast<const __half2 *>(&g[c].x), r01);
        r23 = __hfma2(w2, *reinterpret_cast<const __half2 *>(&g[c].y), r23);
    }
    uint2 out;
    out.x = *reinterpret_cast<uint32_t *>(&r01);
    out.y = *reinterpret_cast<uint32_t *>(&r23);
    return out;
}

// Byte address of the 16-byte chunk (row, k/8 = chunk) in an operand tile whose K extent is k_total
__device__ __forceinline__ uint32_t tile_chunk_off(uint32_t row, uint32_t chunk, uint32_t k_total)
{
    return (row >> 3) * ((k_total >> 3) * 128u) + chunk * 128u + (row & 7u) * 16u;
}

// Cell of one point on one level + the trilinear fractions (first half of grid_level_features)
struct GridCell {
    uint32_t gx, gy, gz;
    float px, py, pz;
};
__device__ __forceinline__ GridCell grid_cell(const GridLevel &L, float x, float y, float z)
{
    GridCell c;
    float px = __fmaf_rn(L.scale, x, 0.5f), py = __fmaf_rn(L.scale, y, 0.5f), pz = __fmaf_rn(L.scale, z, 0.5f);
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    c.gx = (uint32_t)(int)fx, c.gy = (uint32_t)(int)fy, c.gz = (uint32_t)(int)fz;
    c.px = __fsub_rn(px, fx), c.py = __fsub_rn(py, fy), c.pz = __fsub_rn(pz, fz);
    return c;
}
__device__ __forceinline__ void grid_gather_cell(const uint2 *__restrict__ grid, const GridLevel &L, const GridCell &c, uint2 g[8])
{
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t cx = c.gx + (k & 1), cy = c.gy + ((k >> 1) & 1), cz = c.gz + ((k >> 2) & 1);
        uint32_t index;
        if (L.dense)
            index = cx + cy * L.res + cz * L.res * L.res;
        else
            index = cx ^ (cy * 2654435761u) ^ (cz * 805459861u);
        index = grid_wrap(index, L.mask, L.size);
        g[k]  = __ldg(grid + L.offset + index);
    }
}
__device__ __forceinline__ uint2 grid_blend_cell(const GridCell &c, const uint2 g[8])
{
    const float qx = __fsub_rn(1.f, c.px), qy = __fsub_rn(1.f, c.py), qz = __fsub_rn(1.f, c.pz);
    __half2 r01 = __float2half2_rn(0.f), r23 = __float2half2_rn(0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float w = (k & 1) ? c.px : qx; // weight = 1 * w0 * w1 * w2, left to right
        w       = __fmul_rn(w, (k & 2) ? c.py : qy);
        w       = __fmul_rn(w, (k & 4) ? c.pz : qz);
        const __half2 w2 = __float2half2_rn(w);
        r01 = __hfma2(w2, *reinterpret_cast<const __half2 *>(&g[k].x), r01);
        r23 = __hfma2(w2, *reinterpret_cast<const __half2 *>(&g[k].y), r23);
    }
    uint2 out;
    out.x = *reinterpret_cast<uint32_t *>(&r01);
    out.y = *reinterpret_cast<uint32_t *>(&r23);
    return out;
}

// Writes the full fp16 feature row of query `row` (3 points x levels x 4 features) into the A tile.
// pts: 3*n_points floats (encoded positions).  valid=false writes zeros (padding rows of a partial tile).
// Default: (point, level) pairs, 16 independent gathers in flight per thread.  NBVH_GRID_DEDUP (A/B variant): level-major, the
// points of one segment share their grid cell on the coarse levels, so the eight corner entries are gathered once per distinct
// cell (same bits, ~15 % fewer L1 lookups) -- measured on B200: no gain on frames, -12 % on random queries (only 8 gathers in
// flight per thread): the gather phase is bound by memory-level parallelism, not by the lookup count (profiles/r02e_*).
__device__ __forceinline__ void encode_row_to_tile(const NetConst &nc, uint8_t *a_tile, uint32_t row, const float *pts,
                                                   bool valid)
{
#ifndef NBVH_GRID_DEDUP
    const uint32_t pairs = nc.in_k >> 3; // 16-byte chunks per row; chunk j holds flattened (point,level) 2j, 2j+1
#if defined(NBVH_ENCODE_UNROLL2)
#pragma unroll 2
#elif defined(NBVH_ENCODE_UNROLL3)
#pragma unroll 3
#endif
    for (uint32_t j = 0; j < pairs; ++j) {
        uint4 chunk = make_uint4(0u, 0u, 0u, 0u);
        if (valid) {
            const uint32_t q0 = 2u * j, q1 = 2u * j + 1u;
            const uint32_t p0 = q0 / nc.n_levels, l0 = q0 % nc.n_levels;
            const uint32_t p1 = q1 / nc.n_levels, l1 = q1 % nc.n_levels;
            const uint2 f0 = grid_level_features(nc.grid, nc.lv[l0], pts[3 * p0], pts[3 * p0 + 1], pts[3 * p0 + 2]);
            const uint2 f1 = grid_level_features(nc.grid, nc.lv[l1], pts[3 * p1], pts[3 * p1 + 1], pts[3 * p1 + 2]);
            chunk          = make_uint4(f0.x, f0.y, f1.x, f1.y);
        }
        *reinterpret_cast<uint4 *>(a_tile + tile_chunk_off(row, j, nc.in_k)) = chunk;
    }
#else
    const uint32_t n_points = nc.in_k / (nc.n_levels * 4u);
    for (uint32_t l = 0; l < nc.n_levels; ++l) {
        uint2 g[8];
        uint32_t cgx = 0xFFFFFFFFu, cgy = 0u, cgz = 0u; // no cell cached (grid coordinates are far below 2^32 - 1)
        for (uint32_t p = 0; p < n_points; ++p) {
            uint2 f = make_uint2(0u, 0u);
            if (valid) {
                const GridCell c = grid_cell(nc.lv[l], pts[3 * p], pts[3 * p + 1], pts[3 * p + 2]);
                if (c.gx != cgx || c.gy != cgy || c.gz != cgz) {
                    grid_gather_cell(nc.grid, nc.lv[l], c, g);
                    cgx = c.gx, cgy = c.gy, cgz = c.gz;
                }
                f = grid_blend_cell(c, g);
            }
            const uint32_t q = p * nc.n_levels + l; // flattened (point, level): 8 bytes at chunk q / 2, half q & 1
            *reinterpret_cast<uint2 *>(a_tile + tile_chunk_off(row, q >> 1, nc.in_k) + (q & 1u) * 8u) = f;
        }
    }
#endif
}

// ------------------------------------------------------------------------------------------------ MLP (a10)
struct TileCtx {
    uint8_t *a_tile;     // 128 x in_k fp16 operand buffer of this tile (also reused for the 64-wide activations)
    uint32_t w_smem;     // shared address of the packed weights image
    uint32_t mma_bar;    // shared address of this tile's MMA-completion mbarrier
    uint32_t tmem_acc;   // TMEM address (lane 0, first accumulator column) of this tile
    uint32_t bar_id;     // named barrier id of this tile (128 threads)
    uint32_t phase;      // mbarrier phase parity (thread-local copy, identical in all threads of the tile)
};

// Runs all layers on the 128 rows currently in t.a_tile.  All 128 threads of the tile must call it.
// `row` = thread index within the tile.  On return out[0..n_out) holds the post-sigmoid outputs of this
// thread's row (fp32).  The caller must have executed fence_proxy_async_smem() after writing the A tile.
__device__ __forceinline__ void mlp_tile_forward(const NetConst &nc, TileCtx &t, uint32_t row, float *out, unsigned long long *mp = nullptr)
{
    long long mp_t = mp ? clock64() : 0ll; // development builds: [0] barrier, [1] MMA issue, [2] MMA wait, [3] epilogue
    auto mp_lap = [&](int i) {
        if (mp) {
            const long long now = clock64();
            mp[i] += (unsigned long long)(now - mp_t);
            mp_t = now;
        }
    };
    const uint32_t a_addr    = smem_u32(t.a_tile);
    const uint32_t lane_base = (row & ~31u) << 16; // TMEM lane quadrant of this warp
    uint32_t k_in            = nc.in_k;
    for (uint32_t layer = 0; layer < nc.n_layers; ++layer) {
        const bool last      = layer + 1 == nc.n_layers;
        const uint32_t n_out = last ? (uint32_t)kOutPadded : (uint32_t)kNeurons;
        // ---- all rows of A are in shared memory and visible to the async proxy
        tc_fence_before_sync();
        named_bar_sync(t.bar_id, kTileRows);
        mp_lap(layer == 0 ? 0 : 4);
        // The MMAs of the layer are issued by ONE elected lane of the tile's first warp, which enters the block converged: every operand of
        // the instruction stream (tile index via shuffle in the callers, shared-memory addresses, descriptors) is warp-uniform, so ptxas keeps
        // the descriptors in uniform registers and emits UTCHMMA back to back (a `row == 0` branch is divergent: ptxas then wraps every MMA
        // in an ELECT / 5 x R2UR.BROADCAST / BRA.U.ANY loop, measured 1.1 k cycles per layer on the tile's critical path)
        if (row < 32u) {
            if (elect_one_sync()) {
                tc_fence_after_sync();
                const uint32_t sbo_a = (k_in >> 3) * 128u;
                const uint32_t idesc = umma_idesc_f16_f32(128u, n_out);
                // one K step = two 128-byte core-matrix columns: both descriptors advance by 256 B = 16 address units
                uint64_t ad = umma_desc(a_addr, kLBO, sbo_a), bd = umma_desc(t.w_smem + nc.layer_off[layer], kLBO, sbo_a);
                umma_f16(t.tmem_acc, ad, bd, idesc, 0u);
                for (uint32_t k = 16u; k < k_in; k += 16u) {
                    ad += 16u, bd += 16u;
                    umma_f16(t.tmem_acc, ad, bd, idesc, 1u);
                }
                umma_commit(t.mma_bar); // implies tcgen05.fence::before_thread_sync
            }
            __syncwarp();
        }
        mp_lap(1);
        mbar_wait(t.mma_bar, t.phase);
        t.phase ^= 1u;
        tc_fence_after_sync();
        mp_lap(2);
        // ---- epilogue: this thread's accumulator row -> registers
        if (!last) {
#ifdef NBVH_EPI_WIDE
            // A/B variant: all four 16-column loads in flight before one wait
            uint32_t vv[4][16];
#pragma unroll
            for (uint32_t c0 = 0; c0 < (uint32_t)kNeurons; c0 += 16u)
                tmem_ld16(t.tmem_acc + lane_base + c0, vv[c0 >> 4]);
            tmem_ld_wait();
#endif
#pragma unroll
            for (uint32_t c0 = 0; c0 < (uint32_t)kNeurons; c0 += 16u) {
#ifdef NBVH_EPI_WIDE
                uint32_t(&v)[16] = vv[c0 >> 4];
#else
                uint32_t v[16];
                tmem_ld16(t.tmem_acc + lane_base + c0, v);
                tmem_ld_wait();
#endif
                uint32_t h[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float a = fmaxf(__uint_as_float(v[2 * i]), 0.f);
                    const float b = fmaxf(__uint_as_float(v[2 * i + 1]), 0.f);
                    const __half2 p = __floats2half2_rn(a, b);
                    h[i]            = *reinterpret_cast<const uint32_t *>(&p);
                }
                // the activations become the next layer's A operand (K = 64): chunks c0/8 and c0/8+1
                *reinterpret_cast<uint4 *>(t.a_tile + tile_chunk_off(row, c0 >> 3, kNeurons)) =
                    make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4 *>(t.a_tile + tile_chunk_off(row, (c0 >> 3) + 1u, kNeurons)) =
                    make_uint4(h[4], h[5], h[6], h[7]);
            }
            mp_lap(3);
            fence_proxy_async_smem();
            mp_lap(5);
            k_in = kNeurons;
        } else {
            uint32_t v[16];
            tmem_ld16(t.tmem_acc + lane_base, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if ((uint32_t)i < nc.n_out) {
                    const float x = __uint_as_float(v[i]);
                    out[i]        = __fdividef(1.0f, 1.0f + __expf(-x)); // logistic, common_device.h:129-133
                }
            }
        }
    }
    // make sure every thread's TMEM reads are done before the next tile iteration overwrites the accumulators
    tc_fence_before_sync();
}

} // namespace nbvh

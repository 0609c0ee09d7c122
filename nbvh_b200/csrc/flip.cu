// flip.cu -- LDR-FLIP of the accumulated image against a reference image on the device.
// Replaces the reference's flip-cuda dependency as NeuralBVHRenderer::compute_and_display_flip_error drives it
// (/root/reference/src/cuda/neural/neural_bvh_renderer.cu:34-96; algorithm: Andersson et al., "FLIP: A Difference Evaluator for
// Alternating Images", HPG 2020, as implemented in ext/flip-cuda/cudaKernels.cuh:248-470 and color.cuh):
//   linear RGB -> sRGB (no clamp, like the reference) -> linear -> XYZ -> YCxCz;
//   colour pipeline: separable contrast-sensitivity filters (Y and Cx one Gaussian each, Cz the sum of two -> two partial filters),
//   back to clamped linear RGB -> CIELab -> Hunt adjustment -> HyAB distance -> ^qc -> two-segment remap with pc, pt, cmax;
//   feature pipeline: Gaussian / first / second derivative filters of width gw on the normalised Y channel -> edge and point
//   magnitudes -> (|difference| / sqrt 2)^qf;   error = colour^(1 - feature).
// Three kernels instead of flip-cuda's seven (both images and both pipelines share each pass): k_flip_prepare (per pixel),
// k_flip_rows (horizontal pass: 14 partial sums per pixel), k_flip_cols (vertical pass + differences + error + block-reduced sum).
// Edge handling = clamp to the border, summation order = -r ... +r, both like the reference; the filter sums use explicit fmaf (the
// library is built with -fmad=false for the traversal kernels; flip-cuda is built with FMA contraction).  Floating point, not
// bit-exact against flip-cuda: the tests state 2e-4 on the error map.  Intermediate images are planar (one plane per image / per partial
// sum), so both passes read and write contiguous rows per warp.
#include "kernels.cuh"

#include <cfloat>
#include <cmath>
#include <vector>

namespace nbvh {

namespace {

constexpr int kMaxFlipRadius = 64;

struct FlipFilters {
    int r_spatial, r_feature;
    float ycx[2 * kMaxFlipRadius + 1][2];  // Y, Cx weights
    float cz[2 * kMaxFlipRadius + 1][2];   // the two partial Cz filters
    float feat[2 * kMaxFlipRadius + 1][3]; // Gaussian, first derivative, second derivative
    float cmax, pccmax;
};

__constant__ FlipFilters c_flip;

constexpr float kQc = 0.7f, kPc = 0.4f, kPt = 0.95f, kW = 0.082f, kQf = 0.5f; // sharedflip.h:10-16
constexpr float kPi = 3.14159265358979f;

struct C3 {
    float x, y, z;
};

__host__ __device__ inline float srgb_to_linear(float s) { return s <= 0.04045f ? s / 12.92f : powf((s + 0.055f) / 1.055f, 2.4f); }
__host__ __device__ inline float linear_to_srgb(float l) { return l <= 0.0031308f ? l * 12.92f : 1.055f * powf(l, 1.0f / 2.4f) - 0.055f; }

__host__ __device__ inline C3 linear_to_xyz(C3 c)
{
    const float a11 = 10135552.0f / 24577794.0f, a12 = 8788810.0f / 24577794.0f, a13 = 4435075.0f / 24577794.0f;
    const float a21 = 2613072.0f / 12288897.0f, a22 = 8788810.0f / 12288897.0f, a23 = 887015.0f / 12288897.0f;
    const float a31 = 1425312.0f / 73733382.0f, a32 = 8788810.0f / 73733382.0f, a33 = 70074185.0f / 73733382.0f;
    return {a11 * c.x + a12 * c.y + a13 * c.z, a21 * c.x + a22 * c.y + a23 * c.z, a31 * c.x + a32 * c.y + a33 * c.z};
}

__host__ __device__ inline C3 xyz_to_linear(C3 c)
{
    return {3.241003275f * c.x + -1.537398934f * c.y + -0.498615861f * c.z, -0.969224334f * c.x + 1.875930071f * c.y + 0.041554224f * c.z,
            0.055639423f * c.x + -0.204011202f * c.y + 1.057148933f * c.z};
}

// D65 reference illuminant {0.950428545, 1, 1.088900371} and its inverse {1.052156925, 1, 0.918357670} as flip-cuda rounds them
// (color.cuh:65-66)

__host__ __device__ inline C3 xyz_to_ycxcz(C3 c)
{
    const float x = c.x * 1.052156925f, y = c.y * 1.0f, z = c.z * 0.918357670f;
    return {116.0f * y - 16.0f, 500.0f * (x - y), 200.0f * (y - z)};
}

__host__ __device__ inline C3 ycxcz_to_xyz(C3 c)
{
    const float y = (c.x + 16.0f) / 116.0f, cx = c.y / 500.0f, cz = c.z / 200.0f;
    return {(y + cx) * 0.950428545f, y * 1.0f, (y - cz) * 1.088900371f};
}

__host__ __device__ inline C3 xyz_to_lab(C3 c)
{
    const float delta = 6.0f / 29.0f, d2 = delta * delta, d3 = delta * d2, factor = 1.0f / (3.0f * d2), term = 4.0f / 29.0f;
    float x = c.x * 1.052156925f, y = c.y * 1.0f, z = c.z * 0.918357670f;
    x = x > d3 ? powf(x, 1.0f / 3.0f) : factor * x + term;
    y = y > d3 ? powf(y, 1.0f / 3.0f) : factor * y + term;
    z = z > d3 ? powf(z, 1.0f / 3.0f) : factor * z + term;
    return {116.0f * y - 16.0f, 500.0f * (x - y), 200.0f * (y - z)};
}

__host__ __device__ inline C3 hunt_lab(C3 lab) { return {lab.x, 0.01f * lab.x * lab.y, 0.01f * lab.x * lab.z}; }

__host__ __device__ inline float hyab(C3 a, C3 b)
{
    return fabsf(a.x - b.x) + sqrtf((a.y - b.y) * (a.y - b.y) + (a.z - b.z) * (a.z - b.z));
}

__host__ __device__ inline float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

__device__ inline C3 to_ycxcz(float r, float g, float b)
{
    const C3 lin = {srgb_to_linear(linear_to_srgb(r)), srgb_to_linear(linear_to_srgb(g)), srgb_to_linear(linear_to_srgb(b))};
    return xyz_to_ycxcz(linear_to_xyz(lin));
}

// ycc[i] = reference, ycc[n + i] = test (accum / test_divisor): one plane per image, so that a warp's loads are contiguous
__global__ void __launch_bounds__(256) k_flip_prepare(const float4 *reference, const float4 *test, float test_divisor, uint32_t n, float4 *ycc)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const float4 r = reference[i], t = test[i];
    const C3 a = to_ycxcz(r.x, r.y, r.z), b = to_ycxcz(t.x / test_divisor, t.y / test_divisor, t.z / test_divisor);
    ycc[i]             = make_float4(a.x, a.y, a.z, 0.f);
    ycc[(size_t)n + i] = make_float4(b.x, b.y, b.z, 0.f);
}

// rows[k * n + i], 14 planes of n = w * h floats (coalesced in both passes): per image (reference k = 0..6, test 7..13):
// Y, Cx, Cz partial 1, Cz partial 2, dx, ddx, Gaussian
__global__ void __launch_bounds__(256) k_flip_rows(const float4 *ycc, int w, int h, float *rows)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w)
        return;
    float s[14];
#pragma unroll
    for (int k = 0; k < 14; k++)
        s[k] = 0.f;
    const int rs = c_flip.r_spatial, rf = c_flip.r_feature;
    const size_t n = (size_t)w * h;
    for (int ix = -rs; ix <= rs; ix++) {
        const int xx     = min(max(0, x + ix), w - 1);
        const size_t src = (size_t)y * w + xx;
        const float wy = c_flip.ycx[ix + rs][0], wcx = c_flip.ycx[ix + rs][1], wz1 = c_flip.cz[ix + rs][0], wz2 = c_flip.cz[ix + rs][1];
#pragma unroll
        for (int im = 0; im < 2; im++) {
            const float4 c = ycc[im * n + src];
            s[7 * im + 0] = fmaf(wy, c.x, s[7 * im + 0]), s[7 * im + 1] = fmaf(wcx, c.y, s[7 * im + 1]);
            s[7 * im + 2] = fmaf(wz1, c.z, s[7 * im + 2]), s[7 * im + 3] = fmaf(wz2, c.z, s[7 * im + 3]);
        }
    }
    for (int ix = -rf; ix <= rf; ix++) {
        const int xx     = min(max(0, x + ix), w - 1);
        const size_t src = (size_t)y * w + xx;
        const float g = c_flip.feat[ix + rf][0], dg = c_flip.feat[ix + rf][1], ddg = c_flip.feat[ix + rf][2];
#pragma unroll
        for (int im = 0; im < 2; im++) {
            const float yn = fmaf(ycc[im * n + src].x, 1.0f / 116.0f, 16.0f / 116.0f);
            s[7 * im + 4] = fmaf(dg, yn, s[7 * im + 4]), s[7 * im + 5] = fmaf(ddg, yn, s[7 * im + 5]), s[7 * im + 6] = fmaf(g, yn, s[7 * im + 6]);
        }
    }
    float *dst = rows + ((size_t)y * w + x);
#pragma unroll
    for (int k = 0; k < 14; k++)
        dst[k * n] = s[k];
}

__global__ void __launch_bounds__(256) k_flip_cols(const float *rows, int w, int h, float *error_map, double *sum)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    float err = 0.f;
    if (x < w) {
        const int rs = c_flip.r_spatial, rf = c_flip.r_feature;
        const size_t n = (size_t)w * h;
        float col[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        for (int iy = -rs; iy <= rs; iy++) {
            const int yy   = min(max(0, y + iy), h - 1);
            const float *p = rows + ((size_t)yy * w + x);
            const float wy = c_flip.ycx[iy + rs][0], wcx = c_flip.ycx[iy + rs][1], wz1 = c_flip.cz[iy + rs][0], wz2 = c_flip.cz[iy + rs][1];
#pragma unroll
            for (int im = 0; im < 2; im++)
                col[im][0] = fmaf(wy, p[(7 * im) * n], col[im][0]), col[im][1] = fmaf(wcx, p[(7 * im + 1) * n], col[im][1]),
                col[im][2] = fmaf(wz1, p[(7 * im + 2) * n], col[im][2]), col[im][3] = fmaf(wz2, p[(7 * im + 3) * n], col[im][3]);
        }
        C3 lab[2];
#pragma unroll
        for (int im = 0; im < 2; im++) {
            C3 rgb = xyz_to_linear(ycxcz_to_xyz({col[im][0], col[im][1], col[im][2] + col[im][3]}));
            rgb    = {clamp01(rgb.x), clamp01(rgb.y), clamp01(rgb.z)};
            lab[im] = hunt_lab(xyz_to_lab(linear_to_xyz(rgb)));
        }
        float cd = powf(hyab(lab[0], lab[1]), kQc);
        if (cd < c_flip.pccmax)
            cd *= kPt / c_flip.pccmax;
        else
            cd = kPt + ((cd - c_flip.pccmax) / (c_flip.cmax - c_flip.pccmax)) * (1.0f - kPt);
        float f[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}; // dx, ddx, dy, ddy
        for (int iy = -rf; iy <= rf; iy++) {
            const int yy   = min(max(0, y + iy), h - 1);
            const float *p = rows + ((size_t)yy * w + x);
            const float g = c_flip.feat[iy + rf][0], dg = c_flip.feat[iy + rf][1], ddg = c_flip.feat[iy + rf][2];
#pragma unroll
            for (int im = 0; im < 2; im++)
                f[im][0] = fmaf(g, p[(7 * im + 4) * n], f[im][0]), f[im][1] = fmaf(g, p[(7 * im + 5) * n], f[im][1]),
                f[im][2] = fmaf(dg, p[(7 * im + 6) * n], f[im][2]), f[im][3] = fmaf(ddg, p[(7 * im + 6) * n], f[im][3]);
        }
        const float edge_r = sqrtf(f[0][0] * f[0][0] + f[0][2] * f[0][2]), edge_t = sqrtf(f[1][0] * f[1][0] + f[1][2] * f[1][2]);
        const float pt_r = sqrtf(f[0][1] * f[0][1] + f[0][3] * f[0][3]), pt_t = sqrtf(f[1][1] * f[1][1] + f[1][3] * f[1][3]);
        const float fd = powf((1.0f / sqrtf(2.0f)) * fmaxf(fabsf(edge_r - edge_t), fabsf(pt_r - pt_t)), kQf);
        err            = powf(cd, 1.0f - fd);
        if (error_map)
            error_map[(size_t)y * w + x] = err;
    }
    // block sum -> one double atomic per block (the reference pools on the host in a float; a double keeps the mean order-stable)
    __shared__ float warp_sums[8];
    float v = err;
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_down_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0)
        warp_sums[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++)
            s += (double)warp_sums[k];
        atomicAdd(sum, s);
    }
}

inline float gauss_ab(float x2, float a, float b) { return a * std::sqrt(kPi / b) * std::exp(-(kPi * kPi) * x2 / b); }
inline float gauss_sqrt_ab(float x2, float a, float b) { return std::sqrt(a * std::sqrt(kPi / b)) * std::exp(-(kPi * kPi) * x2 / b); }

} // namespace

size_t flip_workspace_bytes(uint32_t w, uint32_t h) { return (size_t)w * h * (2 * sizeof(float4) + 14 * sizeof(float)) + 256; }

// Filters for a given pixels-per-degree value (cudaImage.cuh:399-479, sharedflip.h:61-70)
cudaError_t flip_set_ppd(float ppd, cudaStream_t stream)
{
    static FlipFilters f; // staging copy (cudaMemcpyToSymbolAsync reads it at enqueue time for pageable memory)
    const float b_max = 0.04f;
    f.r_spatial = (int)std::ceil(3.0f * std::sqrt(b_max / (2.0f * kPi * kPi)) * ppd);
    const float sd = 0.5f * kW * ppd;
    f.r_feature = (int)std::ceil(3.0f * sd);
    if (f.r_spatial > kMaxFlipRadius || f.r_feature > kMaxFlipRadius || f.r_spatial < 1 || f.r_feature < 1)
        return cudaErrorInvalidValue;
    const float dx = 1.0f / ppd;
    float sy = 0.f, scx = 0.f, sz1 = 0.f, sz2 = 0.f;
    for (int i = 0; i <= 2 * f.r_spatial; i++) {
        const float ix = (float)(i - f.r_spatial) * dx, x2 = ix * ix;
        f.ycx[i][0] = gauss_ab(x2, 1.0f, 0.0047f), f.ycx[i][1] = gauss_ab(x2, 1.0f, 0.0053f);
        f.cz[i][0] = gauss_sqrt_ab(x2, 34.1f, 0.04f), f.cz[i][1] = gauss_sqrt_ab(x2, 13.5f, 0.025f);
        sy += f.ycx[i][0], scx += f.ycx[i][1], sz1 += f.cz[i][0], sz2 += f.cz[i][1];
    }
    const float ny = 1.0f / sy, ncx = 1.0f / scx, nz = 1.0f / std::sqrt(sz1 * sz1 + sz2 * sz2);
    for (int i = 0; i <= 2 * f.r_spatial; i++)
        f.ycx[i][0] *= ny, f.ycx[i][1] *= ncx, f.cz[i][0] *= nz, f.cz[i][1] *= nz;
    float gs = 0.f, dneg = 0.f, dpos = 0.f, ddneg = 0.f, ddpos = 0.f;
    for (int i = 0; i <= 2 * f.r_feature; i++) {
        const float xx = (float)(i - f.r_feature);
        const float g  = std::exp(-(xx * xx) / (2.0f * sd * sd));
        const float dg = -xx * g, ddg = (xx * xx / (sd * sd) - 1.0f) * g;
        gs += g;
        if (dg > 0.f)
            dpos += dg;
        else
            dneg -= dg;
        if (ddg > 0.f)
            ddpos += ddg;
        else
            ddneg -= ddg;
        f.feat[i][0] = g, f.feat[i][1] = dg, f.feat[i][2] = ddg;
    }
    for (int i = 0; i <= 2 * f.r_feature; i++) {
        f.feat[i][0] /= gs;
        f.feat[i][1] /= f.feat[i][1] > 0.f ? dpos : dneg;
        f.feat[i][2] /= f.feat[i][2] > 0.f ? ddpos : ddneg;
    }
    const C3 green = hunt_lab(xyz_to_lab(linear_to_xyz({0.f, 1.f, 0.f}))), blue = hunt_lab(xyz_to_lab(linear_to_xyz({0.f, 0.f, 1.f})));
    f.cmax   = powf(hyab(green, blue), kQc);
    f.pccmax = kPc * f.cmax;
    return cudaMemcpyToSymbolAsync(c_flip, &f, sizeof(f), 0, cudaMemcpyHostToDevice, stream);
}

cudaError_t launch_flip(const float4 *reference, const float4 *test, float test_divisor, uint32_t w, uint32_t h, void *workspace,
                        float *error_map, double *sum_dev, cudaStream_t stream)
{
    const uint32_t n = w * h;
    float4 *ycc      = reinterpret_cast<float4 *>(workspace);
    float *rows      = reinterpret_cast<float *>(ycc + 2 * (size_t)n);
    cudaError_t e;
    if ((e = cudaMemsetAsync(sum_dev, 0, sizeof(double), stream)) != cudaSuccess)
        return e;
    k_flip_prepare<<<(n + 255u) / 256u, 256, 0, stream>>>(reference, test, test_divisor, n, ycc);
    const dim3 grid((w + 255u) / 256u, h);
    k_flip_rows<<<grid, 256, 0, stream>>>(ycc, (int)w, (int)h, rows);
    k_flip_cols<<<grid, 256, 0, stream>>>(rows, (int)w, (int)h, error_map, sum_dev);
    return cudaGetLastError();
}

} // namespace nbvh

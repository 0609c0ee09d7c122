/*
 * nbvh_oracle.c -- TEST INFRASTRUCTURE ONLY (see nbvh_oracle.h for the rules and the parity status).
 *
 * Scalar CPU restatement of the reference's neural ray-query path.  Compile with
 *   gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC
 * so that the only fused multiply-adds are the explicit fmaf()/fma() calls below, each of which
 * mirrors an fma.rn the reference's own kernels contain when built by nvcc 12.9 for sm_100
 * (PTX of src/cuda/neural/path/path_inference_kernels.cu and src/cuda/neural/neural_bvh_module_kernels.cu).
 * All file:line citations are into /root/reference.
 */
#include "nbvh_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NEURAL_STACK_SIZE 24 /* include/cuda/neural/accel/neural_bvh.cuh:22 */
#define TLAS_STACK_SIZE 10   /* include/cuda/accel/tlas.cuh:16 */
#define BLAS_STACK_SIZE 32   /* include/cuda/accel/bvh2.cuh:24 */

/* glm::min / glm::max operand order (ext/glm/glm/detail/func_common.inl:17-30) */
static inline float gmin(float x, float y) { return (y < x) ? y : x; }
static inline float gmax(float x, float y) { return (x < y) ? y : x; }

/* ------------------------------------------------------------------ fp16 */
float onbvh_h2f(uint16_t h)
{
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp  = (h >> 10) & 0x1Fu;
    uint32_t man  = h & 0x3FFu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal */
            int e = -1;
            do {
                man <<= 1;
                e++;
            } while (!(man & 0x400u));
            man &= 0x3FFu;
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7F800000u | (man << 13);
    } else {
        bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

/* double -> half, round to nearest even (one rounding) */
static uint16_t d2h(double d)
{
    uint16_t sign = 0;
    if (d != d)
        return 0x7FFF;
    if (signbit(d)) {
        sign = 0x8000u;
        d    = -d;
    }
    if (d == 0.0)
        return sign;
    if (d >= 65520.0) /* rounds to inf */
        return sign | 0x7C00u;
    int e;
    double m = frexp(d, &e); /* d = m * 2^e, m in [0.5,1) */
    /* half normal range: 2^-14 .. ; unit in last place is 2^(E-10) with E = e-1 */
    int E = e - 1;
    if (E < -14)
        E = -14; /* subnormal: fixed ulp 2^-24 */
    double scaled = ldexp(d, 10 - E); /* value in units of ulp */
    double r      = nearbyint(scaled); /* default rounding mode = RNE */
    uint32_t q    = (uint32_t)r;
    (void)m;
    if (E == -14 && q < 1024u)
        return sign | (uint16_t)q; /* subnormal or zero */
    if (q == 2048u) {               /* mantissa overflow */
        q = 1024u;
        E++;
    }
    if (E > 15)
        return sign | 0x7C00u;
    return sign | (uint16_t)(((uint32_t)(E + 15) << 10) | (q - 1024u));
}

uint16_t onbvh_f2h(float f) { return d2h((double)f); }

/* __hfma: one rounding of a*b+c to half (tcnn vec.h:210-212,370) */
static inline uint16_t hfma(uint16_t a, uint16_t b, uint16_t c)
{
    return d2h((double)onbvh_h2f(a) * (double)onbvh_h2f(b) + (double)onbvh_h2f(c));
}

/* ------------------------------------------------------------------ RNG (include/cuda/utils/lcg_rng.cuh:12-72) */
static uint32_t murmur_mix(uint32_t hash, uint32_t k)
{
    k *= 0xcc9e2d51u;
    k = (k << 15) | (k >> 17);
    k *= 0x1b873593u;
    hash ^= k;
    hash = ((hash << 13) | (hash >> 19)) * 5u + 0xe6546b64u;
    return hash;
}

uint32_t onbvh_lcg_rng_init(uint32_t index, uint32_t frame, uint32_t linear)
{
    uint32_t h = murmur_mix(0u, linear);
    h          = murmur_mix(h, 512u * frame + index);
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}

float onbvh_lcg_randomf(uint32_t *state)
{
    *state = *state * 1664525u + 1013904223u;
    /* ldexp((float)u, -32): cvt.rn.f32.u32 then an exact scale by 2^-32 (can return 1.0f) */
    return (float)(*state) * 0x1p-32f;
}

/* ------------------------------------------------------------------ camera */
/* glm::inverse(mat4) cofactor formulation (ext/glm/glm/detail/func_matrix.inl compute_inverse<4,4>),
 * host code: no contraction.  m is column-major m[col][row]. */
static void mat4_inverse(const float a[16], float out[16])
{
#define M(c, r) a[(c)*4 + (r)]
    float c00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3);
    float c02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3);
    float c03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3);
    float c04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
    float c06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3);
    float c07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
    float c08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2);
    float c10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2);
    float c11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2);
    float c12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3);
    float c14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3);
    float c15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
    float c16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2);
    float c18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2);
    float c19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2);
    float c20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1);
    float c22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1);
    float c23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);
    float f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
    float f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
    float v0[4] = {M(1, 0), M(0, 0), M(0, 0), M(0, 0)}, v1[4] = {M(1, 1), M(0, 1), M(0, 1), M(0, 1)};
    float v2[4] = {M(1, 2), M(0, 2), M(0, 2), M(0, 2)}, v3[4] = {M(1, 3), M(0, 3), M(0, 3), M(0, 3)};
    float sa[4] = {1.f, -1.f, 1.f, -1.f}, sb[4] = {-1.f, 1.f, -1.f, 1.f};
    float inv[16];
    for (int i = 0; i < 4; i++) {
        float i0 = v1[i] * f0[i] - v2[i] * f1[i] + v3[i] * f2[i];
        float i1 = v0[i] * f0[i] - v2[i] * f3[i] + v3[i] * f4[i];
        float i2 = v0[i] * f1[i] - v1[i] * f3[i] + v3[i] * f5[i];
        float i3 = v0[i] * f2[i] - v1[i] * f4[i] + v2[i] * f5[i];
        inv[0 * 4 + i] = i0 * sa[i];
        inv[1 * 4 + i] = i1 * sb[i];
        inv[2 * 4 + i] = i2 * sa[i];
        inv[3 * 4 + i] = i3 * sb[i];
    }
    float d0 = M(0, 0) * inv[0], d1 = M(0, 1) * inv[4], d2 = M(0, 2) * inv[8], d3 = M(0, 3) * inv[12];
    float det = (d0 + d1) + (d2 + d3);
    float ood = 1.f / det;
    for (int i = 0; i < 16; i++)
        out[i] = inv[i] * ood;
#undef M
}

static void host_normalize(const float v[3], float out[3])
{
    /* glm::normalize = v * inversesqrt(dot(v,v)), inversesqrt = 1/sqrt (host: no fma) */
    float dot = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    float inv = 1.f / sqrtf(dot);
    out[0] = v[0] * inv, out[1] = v[1] * inv, out[2] = v[2] * inv;
}

static void host_cross(const float x[3], const float y[3], float out[3])
{
    out[0] = x[1] * y[2] - y[1] * x[2];
    out[1] = x[2] * y[0] - y[2] * x[0];
    out[2] = x[0] * y[1] - y[0] * x[1];
}

/* src/base/camera/camera.cpp:39-82 + src/cuda/cuda_backend.cpp:595-611 */
void onbvh_camera_params(const float matrix[16], float yfov_rad, uint32_t w, uint32_t h, float origin[3], float du[3],
                         float dv[3], float top_left[3])
{
    float inv[16];
    mat4_inverse(matrix, inv);
    /* inv * (0,0,0,1), inv*(0,0,-1,0), inv*(0,1,0,0): glm mat*vec = (c0*x + c1*y) + (c2*z + c3*w) */
    float dir_raw[3], up_raw[3], dir[3], up[3];
    for (int r = 0; r < 3; r++) {
        origin[r]  = (inv[0 + r] * 0.f + inv[4 + r] * 0.f) + (inv[8 + r] * 0.f + inv[12 + r] * 1.f);
        dir_raw[r] = (inv[0 + r] * 0.f + inv[4 + r] * 0.f) + (inv[8 + r] * -1.f + inv[12 + r] * 0.f);
        up_raw[r]  = (inv[0 + r] * 0.f + inv[4 + r] * 1.f) + (inv[8 + r] * 0.f + inv[12 + r] * 0.f);
    }
    host_normalize(dir_raw, dir);
    host_normalize(up_raw, up);
    /* set_fovy_from_radian stores degrees (glm::degrees = x * 57.29577951308232), update_camera goes back */
    float fovy_deg = yfov_rad * 57.295779513082320876798154814105f;
    float half_rad = (0.5f * fovy_deg) * 0.01745329251994329576923690768489f;
    float plane_y  = 2.f * tanf(half_rad);
    float aspect   = (float)w / (float)h;
    float plane_x  = plane_y * aspect;
    float c1[3], c1n[3], c2[3], c2n[3];
    host_cross(dir, up, c1);
    host_normalize(c1, c1n);
    for (int i = 0; i < 3; i++)
        du[i] = c1n[i] * plane_x;
    host_cross(du, dir, c2);
    host_normalize(c2, c2n);
    for (int i = 0; i < 3; i++)
        dv[i] = -c2n[i] * plane_y;
    for (int i = 0; i < 3; i++)
        top_left[i] = dir[i] - 0.5f * du[i] - 0.5f * dv[i];
}

/* a6: neural_bvh_module_common.cuh:440-446 + neural_bvh_renderer_kernels.cuh:74-79.
 * PTX: dir = fma(xf,du, yf*dv) + top_left;  len2 = fma(z,z, fma(x,x, y*y));  d = dir * rcp(sqrt(len2)) */
void onbvh_raygen(const oconst_t *c, uint32_t pixel, int sample_idx, int rnd_offset, float o[3], float d[3],
                  uint32_t *rng_state)
{
    uint32_t px = pixel % c->fb_w, py = pixel / c->fb_w;
    uint32_t rng = onbvh_lcg_rng_init((uint32_t)sample_idx, (uint32_t)rnd_offset, pixel);
    float r0     = onbvh_lcg_randomf(&rng);
    float xf     = (r0 + (float)px) / (float)c->fb_w;
    float r1     = onbvh_lcg_randomf(&rng);
    float yf     = (r1 + (float)py) / (float)c->fb_h;
    float v[3];
    for (int i = 0; i < 3; i++)
        v[i] = fmaf(xf, c->cam_du[i], yf * c->cam_dv[i]) + c->cam_top_left[i];
    float len2 = fmaf(v[2], v[2], fmaf(v[0], v[0], v[1] * v[1]));
    float inv  = 1.f / sqrtf(len2);
    for (int i = 0; i < 3; i++) {
        d[i] = v[i] * inv;
        o[i] = c->cam_origin[i];
    }
    if (rng_state)
        *rng_state = rng;
}

/* ------------------------------------------------------------------ slab test (a1) */
/* include/cuda/base/ray.cuh:116-132, include/base/accel/bvh.h:123-154 */
int onbvh_slab(const float mn[3], const float mx[3], const float o[3], const float d[3], float ray_mint,
               float ray_maxt, int check, float *mint, float *maxt)
{
    float sm[3], bg[3];
    for (int i = 0; i < 3; i++) {
        float t0 = (mn[i] - o[i]) / d[i];
        float t1 = (mx[i] - o[i]) / d[i];
        sm[i]    = gmin(t0, t1);
        bg[i]    = gmax(t0, t1);
    }
    float cmax = gmax(gmax(sm[0], sm[1]), sm[2]);
    float cmin = gmin(gmin(bg[0], bg[1]), bg[2]);
    *mint      = check ? gmax(ray_mint, cmax) : cmax;
    *maxt      = check ? gmin(ray_maxt, cmin) : cmin;
    return *mint <= *maxt;
}

/* dilated box about the centre (neural_bvh2.cuh:85-91,105-112; bvh.h:223-232) */
static void dilate(const onode_t *n, float dilation, float mn[3], float mx[3])
{
    const float scale = 1.f + dilation;
    for (int i = 0; i < 3; i++) {
        float ext    = n->mx[i] - n->mn[i];
        float center = n->mn[i] + ext * 0.5f;
        float e2     = ext * scale;
        mn[i]        = center - e2 * 0.5f;
        mx[i]        = center + e2 * 0.5f;
    }
}

static int node_primitive_min_max(const onode_t *n, const float o[3], const float d[3], float ray_mint,
                                  float ray_maxt, int check, float dilation, float *mint, float *maxt)
{
    if (dilation != 0.f) {
        float mn[3], mx[3];
        dilate(n, dilation, mn, mx);
        return onbvh_slab(mn, mx, o, d, ray_mint, ray_maxt, check, mint, maxt);
    }
    return onbvh_slab(n->mn, n->mx, o, d, ray_mint, ray_maxt, check, mint, maxt);
}

/* ------------------------------------------------------------------ N-BVH walk (a2, a3) */
/* include/cuda/neural/accel/neural_bvh2.cuh:170-292.  LoD predicates :119-149. */
int onbvh_neural_intersect(const onode_t *nodes, const oconst_t *c, uint32_t lod, const float o[3],
                           const float d[3], uint32_t *stack, int *stack_size, int inference, float current_maxt,
                           float dilation, float *out_mint, float *out_maxt, uint32_t *out_node,
                           uint64_t *interior_steps)
{
    const float ray_mint = -FLT_MAX, ray_maxt = current_maxt;
    const uint32_t leaf_lod = 2u + lod; /* LEAF_LOD(lod) = HIGHEST_BVH_LOD + lod */
    uint32_t node_idx       = stack[--(*stack_size)];
    for (;;) {
        const onode_t *node = &nodes[node_idx];
        const int is_leaf   = (node->a > 0u && node->a <= leaf_lod);
        if (is_leaf || node->b == 0xFFFFFFFFu) {
            float mint, maxt;
            if (node_primitive_min_max(node, o, d, ray_mint, ray_maxt, 0, dilation, &mint, &maxt)) {
                float maxt_eps = inference ? c->inference_maxt_epsilon : c->learning_maxt_epsilon;
                float mint_eps = inference ? c->inference_mint_epsilon : c->learning_mint_epsilon;
                int valid      = mint < ray_maxt;
                if (mint < 0.f)
                    valid &= maxt > maxt_eps;
                else
                    valid &= mint > mint_eps;
                if (valid) {
                    *out_mint = mint;
                    *out_maxt = maxt;
                    if (out_node)
                        *out_node = node_idx;
                    return 1;
                }
            }
            if (*stack_size == 0)
                return 0;
            node_idx = stack[--(*stack_size)];
            continue;
        }
        if (interior_steps)
            (*interior_steps)++;
        uint32_t c1 = node->b, c2 = node->b + 1u;
        float m1, m2, tmp;
        /* intersect_traversal_bbox: check_ray_bounds = true, returns mint or FLT_MAX */
        m1 = node_primitive_min_max(&nodes[c1], o, d, ray_mint, ray_maxt, 1, dilation, &m1, &tmp) ? m1 : FLT_MAX;
        m2 = node_primitive_min_max(&nodes[c2], o, d, ray_mint, ray_maxt, 1, dilation, &m2, &tmp) ? m2 : FLT_MAX;
        if (m1 > m2) {
            float t = m1;
            m1      = m2;
            m2      = t;
            uint32_t ti = c1;
            c1          = c2;
            c2          = ti;
        }
        if (m1 == FLT_MAX) {
            if (*stack_size == 0)
                return 0;
            node_idx = stack[--(*stack_size)];
        } else {
            node_idx = c1;
            if (m2 != FLT_MAX) {
                stack[(*stack_size)++] = c2;
                if (*stack_size >= NEURAL_STACK_SIZE)
                    return 0; /* "Out of bound stack!" -> ray dropped */
            }
        }
    }
}

void onbvh_trace_intervals(const onode_t *nodes, const oconst_t *c, uint32_t n_rays, const float *o, const float *d,
                           uint32_t max_intervals, uint32_t *counts, uint32_t *node_idx, float *mint, float *maxt,
                           uint64_t *interior_steps)
{
    uint64_t steps = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : steps)
    for (int64_t r = 0; r < (int64_t)n_rays; r++) {
        uint32_t stack[NEURAL_STACK_SIZE];
        int ss   = 1;
        stack[0] = 0;
        uint32_t n = 0;
        uint64_t local = 0;
        while (ss > 0) {
            float a, b;
            uint32_t ni = 0;
            if (!onbvh_neural_intersect(nodes, c, c->rendered_lod, o + 3 * r, d + 3 * r, stack, &ss, 1, FLT_MAX,
                                        c->fixed_node_dilation, &a, &b, &ni, &local))
                break;
            if (n < max_intervals) {
                node_idx[(size_t)r * max_intervals + n] = ni;
                mint[(size_t)r * max_intervals + n]     = a;
                maxt[(size_t)r * max_intervals + n]     = b;
            }
            n++;
        }
        counts[r] = n;
        steps += local;
    }
    if (interior_steps)
        *interior_steps = steps;
}

/* ------------------------------------------------------------------ BVH2 closest hit (a14) */
/* include/base/accel/bvh.h:123-137 (BVH2Node::intersect, honours ray bounds) */
static float node_intersect(const onode_t *n, const float o[3], const float d[3], float ray_mint, float ray_maxt)
{
    float mint, maxt;
    return onbvh_slab(n->mn, n->mx, o, d, ray_mint, ray_maxt, 1, &mint, &maxt) ? mint : FLT_MAX;
}

/* include/cuda/accel/bvh.cuh:98-136.  dot(a,b) = fma(a.z,b.z, fma(a.x,b.x, a.y*b.y)); inv_det = rcp.rn(det).
 * cross: the PTX shows mul,mul,sub WITHOUT .rn, and ptxas fuses the minuend product into the subtraction
 * (SASS of non_neural_bvh_intersection: FMUL r = x2*y2; FFMA out = x1*y1 - r).  Pinned by the reference's own
 * kernel run on a B200 (tests/golden/ref_chess_256x144.npz): the unfused form differs in 1-2 ulp of t on ~14 % of
 * the chessboard hits (needle triangles, det ~ 1e-6). */
static inline float dot_ptx(const float a[3], const float b[3]) { return fmaf(a[2], b[2], fmaf(a[0], b[0], a[1] * b[1])); }
static inline void cross_ptx(const float x[3], const float y[3], float out[3])
{
    out[0] = fmaf(x[1], y[2], -(y[1] * x[2]));
    out[1] = fmaf(x[2], y[0], -(y[2] * x[0]));
    out[2] = fmaf(x[0], y[1], -(y[0] * x[1]));
}

static float tri_intersect(const oblas_t *b, uint32_t prim, const float o[3], const float d[3], float ray_mint,
                           float ray_maxt, float *u_out, float *v_out)
{
    const uint32_t *id = b->indices + 3 * (size_t)prim;
    const float *v0 = b->vertices + 3 * (size_t)id[0], *v1 = b->vertices + 3 * (size_t)id[1],
                *v2 = b->vertices + 3 * (size_t)id[2];
    float e1[3], e2[3], pvec[3], tvec[3], qvec[3];
    for (int i = 0; i < 3; i++) {
        e1[i] = v1[i] - v0[i];
        e2[i] = v2[i] - v0[i];
    }
    cross_ptx(d, e2, pvec);
    float det = dot_ptx(e1, pvec);
    if (det > -1e-8f && det < 1e-8f)
        return FLT_MAX;
    float inv_det = 1.f / det;
    for (int i = 0; i < 3; i++)
        tvec[i] = o[i] - v0[i];
    float u = dot_ptx(tvec, pvec) * inv_det;
    if (u < 0.f || u > 1.f)
        return FLT_MAX;
    cross_ptx(tvec, e1, qvec);
    /* dot(ray.d, qvec) and dot(e2, qvec): PTX multiplies the .y pair first, then fuses .x and .z */
    float v = fmaf(qvec[2], d[2], fmaf(qvec[0], d[0], qvec[1] * d[1])) * inv_det;
    if (v < 0.f || (u + v) > 1.f)
        return FLT_MAX;
    float t = fmaf(qvec[2], e2[2], fmaf(e2[0], qvec[0], qvec[1] * e2[1])) * inv_det;
    if (t < ray_mint || t >= ray_maxt)
        return FLT_MAX;
    *u_out = u;
    *v_out = v;
    return t;
}

/* include/cuda/accel/bvh2.cuh:13-88 */
float onbvh_blas_intersect(const oblas_t *blas, const float o[3], const float d[3], float ray_mint, float ray_maxt,
                           float *u, float *v, uint32_t *prim_idx, uint64_t *interior_steps, uint64_t *tri_tests)
{
    uint32_t stack[BLAS_STACK_SIZE];
    int ss        = 0;
    uint32_t ni   = 0;
    float best_t  = FLT_MAX;
    for (;;) {
        const onode_t *node = &blas->nodes[ni];
        if (node->a > 0u) {
            for (uint32_t p = node->b; p < node->b + node->a; p++) {
                uint32_t prim = blas->prim_ids[p];
                float uu = 0.f, vv = 0.f;
                if (tri_tests)
                    (*tri_tests)++;
                float t = tri_intersect(blas, prim, o, d, ray_mint, ray_maxt, &uu, &vv);
                if (t < ray_maxt) {
                    best_t   = t;
                    *u       = uu;
                    *v       = vv;
                    *prim_idx = prim;
                    ray_maxt = t;
                }
            }
            if (ss == 0)
                return best_t;
            ni = stack[--ss];
            continue;
        }
        if (interior_steps)
            (*interior_steps)++;
        uint32_t c1 = node->b, c2 = node->b + 1u;
        float m1 = node_intersect(&blas->nodes[c1], o, d, ray_mint, ray_maxt);
        float m2 = node_intersect(&blas->nodes[c2], o, d, ray_mint, ray_maxt);
        if (m1 > m2) {
            float t = m1;
            m1      = m2;
            m2      = t;
            uint32_t ti = c1;
            c1          = c2;
            c2          = ti;
        }
        if (m1 == FLT_MAX) {
            if (ss == 0)
                return best_t;
            ni = stack[--ss];
        } else {
            ni = c1;
            if (m2 != FLT_MAX)
                stack[ss++] = c2;
        }
    }
}

void onbvh_blas_intersect_batch(const oblas_t *blas, uint32_t n, const float *o, const float *d, const float *maxt,
                                float *t, float *u, float *v, uint32_t *prim, uint64_t *counters)
{
    uint64_t steps = 0, tests = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : steps, tests)
    for (int64_t r = 0; r < (int64_t)n; r++) {
        uint64_t a = 0, b = 0;
        float uu = 0.f, vv = 0.f;
        uint32_t pp = 0xFFFFFFFFu;
        t[r]    = onbvh_blas_intersect(blas, o + 3 * r, d + 3 * r, 1e-4f, maxt ? maxt[r] : FLT_MAX, &uu, &vv, &pp, &a, &b);
        u[r]    = uu;
        v[r]    = vv;
        prim[r] = pp;
        steps += a;
        tests += b;
    }
    if (counters) {
        counters[0] = steps;
        counters[1] = tests;
    }
}

/* ------------------------------------------------------------------ TLAS (a4) */
/* include/cuda/accel/tlas.cuh:220-325 (local stack variant); the ray keeps [1e-4, FLT_MAX]. */
int onbvh_tlas_intersect_only(const oscene_t *s, const float o[3], const float d[3], uint32_t *stack,
                              int *stack_size, uint32_t *blas_idx, int *is_neural)
{
    *is_neural  = 0;
    uint32_t ni = stack[--(*stack_size)];
    for (;;) {
        const onode_t *node = &s->tlas_nodes[ni];
        if (node->b >= 1u) { /* leaf_type >= TLAS_LEAF */
            if (s->n_blas == 1 && node_intersect(node, o, d, 1e-4f, FLT_MAX) == FLT_MAX)
                return 0;
            *is_neural = node->b == 2u;
            *blas_idx  = s->blas_indices[node->a];
            return 1;
        }
        uint32_t c1 = node->a, c2 = node->a + 1u;
        float m1 = node_intersect(&s->tlas_nodes[c1], o, d, 1e-4f, FLT_MAX);
        float m2 = node_intersect(&s->tlas_nodes[c2], o, d, 1e-4f, FLT_MAX);
        if (m1 > m2) {
            float t = m1;
            m1      = m2;
            m2      = t;
            uint32_t ti = c1;
            c1          = c2;
            c2          = ti;
        }
        if (m1 == FLT_MAX) {
            if (*stack_size == 0)
                return 0;
            ni = stack[--(*stack_size)];
        } else {
            ni = c1;
            if (m2 != FLT_MAX) {
                stack[(*stack_size)++] = c2;
                if (*stack_size >= TLAS_STACK_SIZE)
                    return 0;
            }
        }
    }
}

/* ------------------------------------------------------------------ segment encoding (a8) */
/* src/cuda/neural/neural_bvh_module_kernels.cu:10-97,266-303.
 * PTX: world = fma(t, d, o); dilation = (extent_dil + 1) + fixed_dil; centre = (max+min)*0.5;
 *      enc = (world - centre) / (dilation*(max-min)) + 0.5; p = fma(enc_i, s, enc_o*(1-s)),
 *      s = rcp(float(n-1)) * (float(k) + offset). */
/* SegmentDistribution::Uniform (neural_bvh_module_kernels.cu:53-69): t = lcg_randomf(rng) per segment point, drawn from the RAY's
 * generator (encode_bvh_network_input_data :292-297 reads and writes back traversal_data's rng), same blend expression. */

void onbvh_encode_segment(const oconst_t *c, const float o[3], const float d[3], float mint, float maxt, float *out)
{
    onbvh_encode_segment_rng(c, o, d, mint, maxt, NULL, out);
}

void onbvh_encode_segment_rng(const oconst_t *c, const float o[3], const float d[3], float mint, float maxt, uint32_t *rng,
                              float *out)
{
    if (mint < 0.f) {
        float t = mint;
        mint    = maxt;
        maxt    = t;
    }
    const float dilation = (c->node_extent_dilation + 1.f) + c->fixed_node_dilation;
    float eo[3], ei[3];
    for (int i = 0; i < 3; i++) {
        float wo     = fmaf(mint, d[i], o[i]);
        float wi     = fmaf(maxt, d[i], o[i]);
        float extent = dilation * (c->bbox_max[i] - c->bbox_min[i]);
        float center = (c->bbox_max[i] + c->bbox_min[i]) * 0.5f;
        eo[i]        = (wo - center) / extent + 0.5f;
        ei[i]        = (wi - center) / extent + 0.5f;
    }
    const uint32_t n = c->n_segment_points;
    float delta, offset;
    if (c->segment_distribution == 2u && rng) { /* Uniform */
        for (uint32_t k = 0; k < n; k++) {
            float s = onbvh_lcg_randomf(rng);
            for (int i = 0; i < 3; i++)
                out[3 * k + i] = fmaf(ei[i], s, eo[i] * (1.0f - s));
        }
        return;
    }
    if (c->segment_distribution == 1u) { /* EquidistantExterior */
        delta  = (n == 1u) ? 1.f : 1.f / (float)(n - 1u);
        offset = 0.f;
    } else { /* EquidistantInterior */
        delta  = 1.f / (float)n;
        offset = 0.5f;
    }
    for (uint32_t k = 0; k < n; k++) {
        float s = delta * ((float)k + offset);
        for (int i = 0; i < 3; i++)
            out[3 * k + i] = fmaf(ei[i], s, eo[i] * (1.0f - s));
    }
}

/* ------------------------------------------------------------------ hash grid (a9) */
/* ext/tiny-cuda-nn/include/tiny-cuda-nn/common_device.h:709-718 */
static float grid_scale(uint32_t level, float log2_pls, uint32_t base_res)
{
    return exp2f((float)level * log2_pls) * (float)base_res - 1.0f;
}
static uint32_t grid_resolution(float scale) { return (uint32_t)ceilf(scale) + 1u; }

/* ext/tiny-cuda-nn/include/tiny-cuda-nn/encodings/grid.h:692-723 */
uint64_t onbvh_grid_n_params(const onet_t *net, uint32_t *offsets)
{
    uint32_t offset = 0;
    const float l2  = log2f(net->per_level_scale);
    for (uint32_t l = 0; l < net->n_levels; l++) {
        uint32_t res       = grid_resolution(grid_scale(l, l2, net->base_resolution));
        uint32_t max_p     = 0xFFFFFFFFu / 2u;
        uint32_t in_level  = powf((float)res, 3.f) > (float)max_p ? max_p : res * res * res;
        in_level           = (in_level + 7u) / 8u * 8u;
        uint32_t hash_size = 1u << net->log2_hashmap_size;
        if (in_level > hash_size)
            in_level = hash_size;
        if (offsets)
            offsets[l] = offset;
        offset += in_level;
    }
    if (offsets)
        offsets[net->n_levels] = offset;
    return (uint64_t)offset * net->n_features;
}

static uint64_t mlp_n_params(const onet_t *net)
{
    return (uint64_t)net->n_neurons * net->n_input + (uint64_t)(net->n_hidden_layers - 1) * net->n_neurons * net->n_neurons +
           (uint64_t)net->n_output_padded * net->n_neurons;
}

/* grid.h:49-169 (Linear interpolation, Hash grid type, CoherentPrime hash: common_device.h:652-655,691-707) */
void onbvh_grid_encode(const onet_t *net, const float *points, uint32_t n_points, uint16_t *features)
{
    uint32_t offsets[64];
    onbvh_grid_n_params(net, offsets);
    const uint16_t *grid_base = net->params + mlp_n_params(net);
    const float l2            = log2f(net->per_level_scale);
    const uint32_t F          = net->n_features;
    for (uint32_t p = 0; p < n_points; p++) {
        for (uint32_t l = 0; l < net->n_levels; l++) {
            const uint16_t *grid   = grid_base + (size_t)offsets[l] * F;
            const uint32_t hsize   = offsets[l + 1] - offsets[l];
            const float scale      = grid_scale(l, l2, net->base_resolution);
            const uint32_t res     = grid_resolution(scale);
            float pos[3];
            uint32_t pg[3];
            for (int dim = 0; dim < 3; dim++) { /* pos_fract, common_device.h:857-868 */
                float x   = fmaf(scale, points[3 * p + dim], 0.5f);
                float fl  = floorf(x);
                pg[dim]   = (uint32_t)(int)fl;
                pos[dim]  = x - fl;
            }
            uint16_t result[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (uint32_t idx = 0; idx < 8; idx++) {
                float w = 1.f;
                uint32_t q[3];
                for (int dim = 0; dim < 3; dim++) {
                    if ((idx & (1u << dim)) == 0) {
                        w *= 1.f - pos[dim];
                        q[dim] = pg[dim];
                    } else {
                        w *= pos[dim];
                        q[dim] = pg[dim] + 1u;
                    }
                }
                /* grid_index */
                uint32_t stride = 1, index = 0;
                for (int dim = 0; dim < 3 && stride <= hsize; dim++) {
                    index += q[dim] * stride;
                    stride *= res;
                }
                if (hsize < stride)
                    index = (q[0] * 1u) ^ (q[1] * 2654435761u) ^ (q[2] * 805459861u);
                index %= hsize;
                uint16_t wh = onbvh_f2h(w);
                for (uint32_t f = 0; f < F; f++)
                    result[f] = hfma(wh, grid[(size_t)index * F + f], result[f]);
            }
            for (uint32_t f = 0; f < F; f++)
                features[(size_t)p * net->n_levels * F + l * F + f] = result[f];
        }
    }
}

/* ------------------------------------------------------------------ MLP (a10) */
static inline float logistic(float x) { return 1.0f / (1.0f + expf(-x)); } /* common_device.h:129-133 */

/* Weight layout: row-major W[out][in], matrices back to back (fully_fused_mlp.cu:855-863). */
void onbvh_mlp_forward(const onet_t *net, const uint16_t *features, int mode, float *out, float *hidden)
{
    float act[256], next[256];
    uint32_t in_w = net->n_input;
    for (uint32_t i = 0; i < in_w; i++)
        act[i] = onbvh_h2f(features[i]);
    const uint16_t *W = net->params;
    const uint32_t n_mats = net->n_hidden_layers + 1;
    for (uint32_t m = 0; m < n_mats; m++) {
        const int last       = (m == n_mats - 1);
        const uint32_t out_w = last ? net->n_output_padded : net->n_neurons;
        for (uint32_t r = 0; r < out_w; r++) {
            const uint16_t *row = W + (size_t)r * in_w;
            float acc;
            if (mode == 1) { /* fp16 accumulator, rounded after every 16-wide K chunk (wmma m16n16k16) */
                uint16_t acc_h = 0;
                for (uint32_t k0 = 0; k0 < in_w; k0 += 16) {
                    float part = onbvh_h2f(acc_h);
                    for (uint32_t k = k0; k < k0 + 16 && k < in_w; k++)
                        part += onbvh_h2f(row[k]) * act[k];
                    acc_h = onbvh_f2h(part);
                }
                acc = onbvh_h2f(acc_h);
            } else { /* fp32 accumulate, sequential in K like one MMA lane would sum exact products */
                double a = 0.0;
                for (uint32_t k = 0; k < in_w; k++)
                    a += (double)onbvh_h2f(row[k]) * (double)act[k];
                acc = (float)a;
            }
            if (!last) {
                float v = acc > 0.f ? acc : 0.f;
                next[r] = (mode == 2) ? v : onbvh_h2f(onbvh_f2h(v)); /* activations travel as fp16 */
            } else {
                float s = logistic(acc);
                next[r] = (mode == 1) ? onbvh_h2f(onbvh_f2h(s)) : s;
            }
        }
        W += (size_t)out_w * in_w;
        if (!last && hidden)
            memcpy(hidden + (size_t)m * net->n_neurons, next, sizeof(float) * net->n_neurons);
        memcpy(act, next, sizeof(float) * out_w);
        in_w = out_w;
    }
    for (uint32_t i = 0; i < net->n_output; i++)
        out[i] = act[i];
}

static void query_one(const onet_t *net, const float *pts, uint32_t n_points, int mode, float *out)
{
    uint16_t feat[256];
    onbvh_grid_encode(net, pts, n_points, feat);
    if (mode == 2) { /* keep fp16-valued features; fp32 everywhere else */
    }
    onbvh_mlp_forward(net, feat, mode, out, NULL);
}

void onbvh_query(const onet_t *net, const float *in, uint32_t n, int mode, float *out)
{
    const uint32_t n_points = net->n_input / (net->n_levels * net->n_features);
    const uint32_t in_dims  = 3 * n_points;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++)
        query_one(net, in + (size_t)i * in_dims, n_points, mode, out + (size_t)i * net->n_output);
}

/* ------------------------------------------------------------------ decode (a11) */
/* neural_bvh_module_common.cuh:23-57,160-209 (LocalT).  PTX: the literal 1e-3 is a double:
 *   outside: t = (float)fma((double)len + 1e-3, (double)u, (double)mint)
 *   inside : e = (float)(((double)len + 1e-3) * (double)u);  accept iff e < (maxt - len*eps) - fixed_eps */
int onbvh_decode(const oconst_t *c, const float *y, float seg_mint, float seg_maxt, uint32_t depth, float *t_out)
{
    float thr = depth == 1u ? c->primary_visibility_threshold : c->secondary_visibility_threshold;
    if (!(y[0] > thr)) {
        *t_out = FLT_MAX;
        return 0;
    }
    float len  = seg_maxt - seg_mint;
    float unit = y[1];
    unit       = gmin(gmax(unit, 0.f), 1.f); /* glm::clamp = min(max(x,lo),hi) */
    float t    = FLT_MAX;
    if (seg_mint < 0.f) {
        if (c->depth_epsilon != 1.f) {
            float inward = (float)(((double)len + 1e-3) * (double)unit);
            float lim    = (seg_maxt - len * c->depth_epsilon) - c->fixed_depth_epsilon;
            if (inward < lim)
                t = seg_maxt - inward;
        }
    } else {
        t = (float)fma((double)len + 1e-3, (double)unit, (double)seg_mint);
    }
    *t_out = t;
    return t != FLT_MAX;
}


/* ------------------------------------------------------------------ training-sample generator (a17, K10) */
/* CUDA texture unit, cudaFilterModeLinear + cudaAddressModeWrap + normalized coordinates + cudaReadModeNormalizedFloat
 * (CUDA programming guide, "Texture fetching": x = frac(u)*N - 0.5, weights in 1.8 fixed point). */
void onbvh_tex2d(const otexture_t *t, float u, float v, float out[4])
{
    float x = (u - floorf(u)) * (float)t->w - 0.5f, y = (v - floorf(v)) * (float)t->h - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float a = floorf((x - fx) * 256.f + 0.5f) / 256.f, b = floorf((y - fy) * 256.f + 0.5f) / 256.f;
    int i0 = (int)fx, j0 = (int)fy;
    int i1 = i0 + 1, j1 = j0 + 1;
    i0 = ((i0 % (int)t->w) + (int)t->w) % (int)t->w, i1 = ((i1 % (int)t->w) + (int)t->w) % (int)t->w;
    j0 = ((j0 % (int)t->h) + (int)t->h) % (int)t->h, j1 = ((j1 % (int)t->h) + (int)t->h) % (int)t->h;
    for (int ch = 0; ch < 4; ch++) {
        float t00 = t->rgba[4 * ((size_t)j0 * t->w + i0) + ch] / 255.f, t10 = t->rgba[4 * ((size_t)j0 * t->w + i1) + ch] / 255.f;
        float t01 = t->rgba[4 * ((size_t)j1 * t->w + i0) + ch] / 255.f, t11 = t->rgba[4 * ((size_t)j1 * t->w + i1) + ch] / 255.f;
        out[ch]   = (1.f - a) * (1.f - b) * t00 + a * (1.f - b) * t10 + (1.f - a) * b * t01 + a * b * t11;
    }
}

static inline void s_normalize(const float v[3], float out[3])
{
    float inv = 1.f / sqrtf(dot_ptx(v, v)); /* glm::normalize on the device: v * rcp(sqrt(dot)) */
    out[0] = v[0] * inv, out[1] = v[1] * inv, out[2] = v[2] * inv;
}
/* w*a + u*b + v*c as the reference's SASS evaluates it (pinned bit-exactly by tests/golden/ref_chess_sampler.npz) */
static inline float s_bary(float w, float a, float u, float b, float v, float c) { return fmaf(v, c, fmaf(w, a, u * b)); }

/* include/cuda/utils/common.cuh:21-36 */
static void s_ortho_basis(float tangent[3], float bitangent[3], const float n[3])
{
    float b[3] = {0.f, 0.f, 0.f}, c[3];
    if (n[0] < 0.6f && n[0] > -0.6f)
        b[0] = 1.f;
    else if (n[1] < 0.6f && n[1] > -0.6f)
        b[1] = 1.f;
    else if (n[2] < 0.6f && n[2] > -0.6f)
        b[2] = 1.f;
    else
        b[0] = 1.f;
    cross_ptx(b, n, c);
    s_normalize(c, tangent);
    cross_ptx(n, tangent, c);
    s_normalize(c, bitangent);
}

/* include/cuda/utils/sampling.cuh:8-14 */
static void s_uniform_sphere(float sx, float sy, float out[3])
{
    float phi       = 6.2831854820251465f * sx;
    float cos_theta = fmaf(sy, 2.0f, -1.0f);
    float sin_theta = sqrtf(fmaf(-cos_theta, cos_theta, 1.0f));
    out[0] = sin_theta * cosf(phi), out[1] = sin_theta * sinf(phi), out[2] = cos_theta;
}

/* include/cuda/base/material.cuh:101-120,193-200 */
static void s_lambertian_dir(const float n[3], const float tangent[3], const float bitangent[3], float ux, float uy, float out[3])
{
    float sx = fmaf(2.f, ux, -1.f), sy = fmaf(2.f, uy, -1.f), dx = sx, dy = sy;
    if (!(sx == 0.f && sy == 0.f)) {
        float radius, theta;
        if (fabsf(sx) > fabsf(sy)) {
            radius = sx;
            theta  = 0.78539818525314331f * (sy / sx);
        } else {
            radius = sy;
            theta  = 1.5707963705062866f - 0.78539818525314331f * (sx / sy);
        }
        dx = radius * cosf(theta), dy = radius * sinf(theta);
    }
    float h[3] = {dx, dy, sqrtf(fmaxf(0.f, (1.f - dx * dx) - dy * dy))}, hn[3];
    s_normalize(h, hn);
    for (int i = 0; i < 3; i++)
        out[i] = s_bary(hn[0], tangent[i], hn[1], bitangent[i], hn[2], n[i]);
}

typedef struct {
    float position[3], sh_normal[3], uv[2];
    int32_t material_id;
} oits_t;

/* BaseCudaBLAS::compute_its_data (include/cuda/accel/bvh.cuh:53-96) */
static void s_compute_its(const osampler_t *sp, uint32_t prim, float u, float v, oits_t *its)
{
    const uint32_t *id = sp->blas.indices + 3 * (size_t)prim;
    uint32_t left = 0, right = sp->num_materials;
    while (left < right) { /* upper_bound(material_map, prim) - 1 */
        uint32_t mid = left + (right - left) / 2;
        if (sp->material_map[mid] <= prim)
            left = mid + 1;
        else
            right = mid;
    }
    its->material_id = (sp->material_ids && sp->num_materials) ? sp->material_ids[left - 1] : -1;
    const float *v0 = sp->blas.vertices + 3 * (size_t)id[0], *v1 = sp->blas.vertices + 3 * (size_t)id[1],
                *v2 = sp->blas.vertices + 3 * (size_t)id[2];
    float w = (1.f - u) - v;
    if (sp->normals) {
        const float *n0 = sp->normals + 3 * (size_t)id[0], *n1 = sp->normals + 3 * (size_t)id[1], *n2 = sp->normals + 3 * (size_t)id[2];
        float n[3];
        for (int i = 0; i < 3; i++)
            n[i] = s_bary(w, n0[i], u, n1[i], v, n2[i]);
        s_normalize(n, its->sh_normal);
    } else {
        float e1[3], e2[3], g[3];
        for (int i = 0; i < 3; i++)
            e1[i] = v1[i] - v0[i], e2[i] = v2[i] - v0[i];
        cross_ptx(e1, e2, g);
        s_normalize(g, its->sh_normal);
    }
    for (int i = 0; i < 3; i++)
        its->position[i] = s_bary(w, v0[i], u, v1[i], v, v2[i]);
    if (sp->texcoords) {
        const float *t0 = sp->texcoords + 2 * (size_t)id[0], *t1 = sp->texcoords + 2 * (size_t)id[1], *t2 = sp->texcoords + 2 * (size_t)id[2];
        its->uv[0] = s_bary(w, t0[0], u, t1[0], v, t2[0]);
        its->uv[1] = s_bary(w, t0[1], u, t1[1], v, t2[1]);
    } else {
        its->uv[0] = u, its->uv[1] = v;
    }
    if (its->material_id != -1 && sp->materials) { /* unpack_normal (material.cuh:66-75) */
        uint32_t mask;
        memcpy(&mask, &sp->materials[its->material_id].normal[0], 4);
        if (mask & 0x80000000u) {
            float tx[4];
            onbvh_tex2d(&sp->textures[mask & 0x1fffffffu], its->uv[0], its->uv[1], tx);
            if (!(tx[0] == -1.f && tx[1] == -1.f && tx[2] == -1.f)) {
                float c[3] = {tx[0] - 0.5f, tx[1] - 0.5f, tx[2] - 0.5f}, tsn[3], tw[3], bw[3], n[3];
                s_normalize(c, tsn);
                s_ortho_basis(tw, bw, its->sh_normal);
                memcpy(n, its->sh_normal, 12);
                for (int i = 0; i < 3; i++)
                    its->sh_normal[i] = s_bary(tw[i], tsn[0], bw[i], tsn[1], n[i], tsn[2]);
            }
        }
    }
}

/* test hook: when set, receives (prim, u, v) of every written slot (3 floats per slot, prim bit-cast) */
float *onbvh_sampler_debug = NULL;

/* neural_bvh_module_kernels.cu:119-264 with stack_based_sample_uniform_bvh_density and
 * generate_single_bvh_network_output_reference (neural_bvh_module_common.cuh:703-861); LocalT depth targets */
void onbvh_generate_batch(const onode_t *nbvh, const oconst_t *c, const osampler_t *sp, uint32_t training_step,
                          uint32_t batch_size, uint32_t log2_bvh_max_traversals, uint32_t *sampled_nodes,
                          float current_max_loss, const otreecut_t *tree_cut, float *inputs, float *targets)
{
    const uint32_t sub_batch = batch_size >> log2_bvh_max_traversals, n_sub = 1u << log2_bvh_max_traversals;
    const uint32_t n_in = 3u * c->n_segment_points, n_out = 8u;
    const float dilation_enc = (c->node_extent_dilation + 1.f) + c->fixed_node_dilation;
    const uint32_t n_lods    = 7u - sp->current_learned_lod; /* NUM_BVH_LODS = 7 */
#pragma omp parallel for schedule(dynamic, 64)
    for (uint32_t i = 0; i < sub_batch; i++) {
        uint32_t stack[NEURAL_STACK_SIZE];
        int ss = 0, cont = 0, swap_dir = 0, bounce = 0;
        float leaf_mint = 0.f, leaf_maxt = 0.f, o[3] = {0, 0, 0}, d[3] = {1, 1, 1};
        uint32_t slot = i * n_sub;
        for (uint32_t inner = 0; inner < n_sub; inner++) {
            uint32_t rng = onbvh_lcg_rng_init(training_step * batch_size, 0u, slot);
            float s0 = onbvh_lcg_randomf(&rng), s1 = onbvh_lcg_randomf(&rng), dir[3];
            s_uniform_sphere(s0, s1, dir);
            int valid = cont && ss != 0, restarted = 0;
            uint32_t lod = (uint32_t)(onbvh_lcg_randomf(&rng) * (float)n_lods);
            lod          = sp->current_learned_lod + (lod < n_lods - 1u ? lod : n_lods - 1u);
            float sampled_dilation = fmaf(onbvh_lcg_randomf(&rng), c->node_extent_dilation, c->fixed_node_dilation);
            uint32_t node_idx      = 0;
            do {
                if (!valid) {
                    stack[0] = 0, ss = 1;
                    float off = (sp->sampling_bbox_scaling - 1.f) * 0.5f;
                    for (int k = 0; k < 3; k++) {
                        float l = fmaf(sp->sampling_bbox_scaling, onbvh_lcg_randomf(&rng), -off);
                        o[k]    = fmaf(l, c->bbox_max[k] - c->bbox_min[k], c->bbox_min[k]);
                        d[k]    = dir[k];
                    }
                    restarted = 1;
                }
                valid = onbvh_neural_intersect(nbvh, c, lod, o, d, stack, &ss, 0, FLT_MAX, sampled_dilation, &leaf_mint, &leaf_maxt,
                                               &node_idx, NULL);
            } while (!valid);
            sampled_nodes[slot] = node_idx;
            if (restarted)
                bounce = 0, swap_dir = 0;
            if (swap_dir && leaf_mint >= 0.f)
                swap_dir = 0;
            if (sp->training_node_survival) {
                otreecut_t nd = tree_cut[node_idx];
                float est     = -FLT_MAX;
                if (nd.samples != 0u) {
                    float ia = (float)nd.iter_alive;
                    est      = nd.loss / (ia * sqrtf((float)nd.samples / ia));
                }
                float keep = fmaxf(sp->min_rr_survival_prob, est / current_max_loss);
                if (!(est <= current_max_loss))
                    keep = 1.f;
                if (sp->survival_probability_scale * keep < onbvh_lcg_randomf(&rng))
                    continue; /* slot not advanced, contents kept */
            }
            { /* encode_segment_from_world_po_pi, no negative-mint swap here */
                float t_o = swap_dir ? leaf_maxt : leaf_mint, t_i = swap_dir ? leaf_mint : leaf_maxt, eo[3], ei[3];
                for (int k = 0; k < 3; k++) {
                    float extent = dilation_enc * (c->bbox_max[k] - c->bbox_min[k]);
                    float center = (c->bbox_max[k] + c->bbox_min[k]) * 0.5f;
                    eo[k]        = (fmaf(t_o, d[k], o[k]) - center) / extent + 0.5f;
                    ei[k]        = (fmaf(t_i, d[k], o[k]) - center) / extent + 0.5f;
                }
                const uint32_t n = c->n_segment_points;
                float delta = c->segment_distribution == 1u ? (n == 1u ? 1.f : 1.f / (float)(n - 1u)) : 1.f / (float)n;
                float offs  = c->segment_distribution == 1u ? 0.f : 0.5f;
                for (uint32_t k = 0; k < n; k++) {
                    float s = delta * ((float)k + offs);
                    for (int a = 0; a < 3; a++)
                        inputs[(size_t)slot * n_in + 3 * k + a] = fmaf(ei[a], s, eo[a] * (1.0f - s));
                }
            }
            float td[3] = {d[0], d[1], d[2]}, seg_mint = leaf_mint, seg_maxt = leaf_maxt;
            if (swap_dir) {
                td[0] = -d[0], td[1] = -d[1], td[2] = -d[2];
                seg_mint = -leaf_maxt, seg_maxt = -leaf_mint;
            }
            float hu = 0.f, hv = 0.f;
            uint32_t prim = 0xFFFFFFFFu;
            float ht      = onbvh_blas_intersect(&sp->blas, o, td, seg_mint, seg_maxt, &hu, &hv, &prim, NULL, NULL);
            int is_hit    = ht != FLT_MAX;
            float *tg     = targets + (size_t)slot * n_out;
            oits_t its;
            memset(&its, 0, sizeof(its));
            if (is_hit)
                s_compute_its(sp, prim, hu, hv, &its);
            if (onbvh_sampler_debug) {
                memcpy(&onbvh_sampler_debug[3 * (size_t)slot], &prim, 4);
                onbvh_sampler_debug[3 * (size_t)slot + 1] = hu, onbvh_sampler_debug[3 * (size_t)slot + 2] = hv;
            }
            tg[0] = is_hit ? 1.f : 0.f;
            float depth = 0.f;
            if (is_hit) {
                depth = (float)((double)(ht - seg_mint) / ((double)(seg_maxt - seg_mint) + 1e-3));
                if (sp->depth_linear_activation)
                    depth = fmaf(depth, 0.1f, 0.45f);
            }
            tg[1] = depth;
            for (int k = 0; k < 3; k++)
                tg[2 + k] = is_hit ? (its.sh_normal[k] + 1.f) * 0.5f : 0.f;
            float alb[3] = {0.f, 0.f, 0.f};
            if (is_hit) {
                alb[0] = alb[1] = alb[2] = 0.9f;
                if (its.material_id != -1 && sp->materials) {
                    const omaterial_t *m = &sp->materials[its.material_id];
                    uint32_t mask;
                    memcpy(&mask, &m->base_color[0], 4);
                    if (mask & 0x80000000u) {
                        float tx[4];
                        onbvh_tex2d(&sp->textures[mask & 0x1fffffffu], its.uv[0], its.uv[1], tx);
                        alb[0] = tx[0], alb[1] = tx[1], alb[2] = tx[2];
                    } else {
                        alb[0] = m->base_color[0], alb[1] = m->base_color[1], alb[2] = m->base_color[2];
                    }
                }
            }
            tg[5] = alb[0], tg[6] = alb[1], tg[7] = alb[2];
            swap_dir = 0;
            cont     = !is_hit;
            if (is_hit && sp->max_training_bounces > 0) {
                bounce++;
                float wo[3] = {-d[0], -d[1], -d[2]}, shn[3] = {its.sh_normal[0], its.sh_normal[1], its.sh_normal[2]}, tangent[3], bitangent[3];
                int backface = 0;
                if (dot_ptx(wo, shn) < 0.f) {
                    shn[0] = -shn[0], shn[1] = -shn[1], shn[2] = -shn[2];
                    backface = 1;
                }
                s_ortho_basis(tangent, bitangent, shn);
                if ((backface && !sp->backface_bounce_training) || bounce >= sp->max_training_bounces) {
                    bounce = 0;
                    cont   = 0;
                } else {
                    float u0 = onbvh_lcg_randomf(&rng), u1 = onbvh_lcg_randomf(&rng), nd[3];
                    if (sp->sample_cosine_weighted) {
                        s_lambertian_dir(shn, tangent, bitangent, u0, u1, nd);
                    } else {
                        s_uniform_sphere(u0, u1, nd);
                        if (!(dot_ptx(wo, shn) * dot_ptx(nd, shn) > 0.f))
                            nd[0] = -nd[0], nd[1] = -nd[1], nd[2] = -nd[2];
                    }
                    for (int k = 0; k < 3; k++)
                        o[k] = its.position[k], d[k] = nd[k];
                    cont = 1, swap_dir = 1;
                    stack[0] = 0, ss = 1;
                }
            }
            slot++;
        }
    }
}

/* ------------------------------------------------------------------ primary-ray pipeline */
/* closest hit of one ray over the whole TLAS: classical BLAS closest hits + neural leaf queries, in TLAS order */
/* rng: the ray's generator (Uniform segment points draw from it and leave it advanced for the shading step); NULL = none */
static void trace_ray(const oscene_t *s, const onode_t *nbvh, const oconst_t *c, const onet_t *net, const float o[3],
                      const float d[3], uint32_t ray_depth, uint32_t lod, int mlp_mode, ohit_t *hit, uint64_t *cnt, uint32_t *rng)
{
    memset(hit, 0, sizeof(*hit));
    hit->t = FLT_MAX;
    uint32_t tstack[TLAS_STACK_SIZE];
    int tss   = 1;
    tstack[0] = 0;
    const uint32_t n_points = c->n_segment_points;
    while (tss > 0) {
        uint32_t blas_idx;
        int is_neural;
        if (!onbvh_tlas_intersect_only(s, o, d, tstack, &tss, &blas_idx, &is_neural))
            break;
        if (is_neural && !c->ignore_neural_blas) {
            /* tlas_traversal_and_fill_next_hybrid_traversal (common.cuh:323-380): first interval with FLT_MAX */
            uint32_t stack[NEURAL_STACK_SIZE];
            int ss      = 1;
            stack[0]    = 0;
            float maxt_ray = FLT_MAX;
            for (;;) {
                float mint, maxt;
                uint32_t node;
                if (!onbvh_neural_intersect(nbvh, c, lod, o, d, stack, &ss, 1, maxt_ray,
                                            c->fixed_node_dilation, &mint, &maxt, &node, &cnt[1]))
                    break;
                float pts[64], y[16];
                onbvh_encode_segment_rng(c, o, d, mint, maxt, rng, pts);
                query_one(net, pts, n_points, mlp_mode, y);
                cnt[0]++;
                float t;
                /* process_neural_bvh_traversal_step (path_inference_kernels.cu:61-131) */
                if (onbvh_decode(c, y, mint, maxt, ray_depth, &t) && t < hit->t) {
                    hit->valid_hit = 1;
                    hit->t         = t;
                    float n[3]     = {y[2] * 2.f - 1.f, y[3] * 2.f - 1.f, y[4] * 2.f - 1.f};
                    float len2     = fmaf(n[2], n[2], fmaf(n[0], n[0], n[1] * n[1]));
                    float inv      = 1.f / sqrtf(len2);
                    for (int i = 0; i < 3; i++) {
                        hit->normal_or_uv[i]  = n[i] * inv;
                        hit->albedo_or_ids[i] = y[5 + i];
                    }
                }
                /* resume with ray.maxt = current best t (common.cuh:529-545) */
                maxt_ray = hit->t;
                if (ss == 0)
                    break;
            }
        } else {
            /* non_neural_bvh_intersection (path_inference_kernels.cu:195-222): ray.maxt = last_tlas_t */
            float u = 0.f, v = 0.f;
            uint32_t prim = 0;
            float t = onbvh_blas_intersect(&s->blas[blas_idx], o, d, 1e-4f, hit->t, &u, &v, &prim, &cnt[2], &cnt[3]);
            if (t < hit->t) {
                hit->valid_hit       = 2;
                hit->t               = t;
                hit->normal_or_uv[0] = u;
                hit->normal_or_uv[1] = v;
                memcpy(&hit->albedo_or_ids[0], &blas_idx, 4);
                memcpy(&hit->albedo_or_ids[1], &prim, 4);
            }
        }
    }
}

static void render_one(const oscene_t *s, const onode_t *nbvh, const oconst_t *c, const onet_t *net, uint32_t pixel,
                       int sample_idx, int rnd_offset, int mlp_mode, ohit_t *hit, uint64_t *cnt)
{
    float o[3], d[3];
    uint32_t rng;
    onbvh_raygen(c, pixel, sample_idx, rnd_offset, o, d, &rng);
    trace_ray(s, nbvh, c, net, o, d, 1u, c->rendered_lod, mlp_mode, hit, cnt, &rng);
}

void onbvh_render_primary(const oscene_t *s, const onode_t *nbvh, const oconst_t *c, const onet_t *net,
                          uint32_t pixel_begin, uint32_t pixel_end, uint32_t pixel_stride, int sample_idx,
                          int rnd_offset, int mlp_mode, ohit_t *hits, uint64_t *counters, int n_threads)
{
    uint64_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    if (pixel_stride == 0)
        pixel_stride = 1;
    const int64_t n = ((int64_t)pixel_end - (int64_t)pixel_begin + pixel_stride - 1) / pixel_stride;
#ifdef _OPENMP
    if (n_threads > 0)
        omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : c0, c1, c2, c3)
    for (int64_t i = 0; i < n; i++) {
        uint64_t cnt[4] = {0, 0, 0, 0};
        render_one(s, nbvh, c, net, pixel_begin + (uint32_t)i * pixel_stride, sample_idx, rnd_offset, mlp_mode,
                   &hits[i], cnt);
        c0 += cnt[0], c1 += cnt[1], c2 += cnt[2], c3 += cnt[3];
    }
    if (counters) {
        counters[0] = c0, counters[1] = c1, counters[2] = c2, counters[3] = c3;
    }
}

/* ================================================================== hybrid neural path tracing (a15, f1) */
/* Restates, in plain fp32 without contraction:
 *   unpack_material / unpack_albedo           include/cuda/base/material.cuh:24-89
 *   Disney BSDF sample / eval / pdf           include/cuda/base/material.cuh:91-592
 *   Fr / Refract                              include/cuda/base/wip_disney.cuh:128-152
 *   shading_frame / rr_terminate              include/cuda/utils/rendering.cuh:11-40
 *   background_hit_eval                       src/cuda/neural/path/path_inference_kernels.cu:37-59
 *   hybrid_scatter_and_get_next_tlas_traversal_step + neural_path_tracing
 *                                             src/cuda/neural/path/path_inference_kernels.cu:224-429, path_module.cu:139-245 */
#define O_PI 3.14159265358979323846f
#define O_INV_PI 0.31830988618379067154f
typedef struct {
    float base_color[3], metallic, specular, roughness, specular_tint, anisotropy, sheen, sheen_tint, clearcoat,
        clearcoat_gloss, base_emission[3], emission_scale, ior, specular_transmission;
} obsdf_t;

static float p_scalar(float x, const float uv[2], const otexture_t *tex)
{
    uint32_t mask;
    memcpy(&mask, &x, 4);
    if (mask & 0x80000000u) {
        float t[4];
        onbvh_tex2d(&tex[mask & 0x1fffffffu], uv[0], uv[1], t);
        return t[(mask >> 29) & 3u];
    }
    return x;
}
static void p_color(const float p[3], const float uv[2], const otexture_t *tex, float out[3])
{
    uint32_t mask;
    memcpy(&mask, &p[0], 4);
    if (mask & 0x80000000u) {
        float t[4];
        onbvh_tex2d(&tex[mask & 0x1fffffffu], uv[0], uv[1], t);
        out[0] = t[0], out[1] = t[1], out[2] = t[2];
    } else {
        out[0] = p[0], out[1] = p[1], out[2] = p[2];
    }
}
static void p_default_bsdf(obsdf_t *b)
{
    memset(b, 0, sizeof(*b));
    b->base_color[0] = b->base_color[1] = b->base_color[2] = 0.9f;
    b->roughness = 1.f, b->emission_scale = 1.f, b->ior = 1.5f;
}
static void p_unpack_material(const omaterial_t *m, const float uv[2], const otexture_t *tex, obsdf_t *b)
{
    p_color(m->base_color, uv, tex, b->base_color);
    b->metallic              = p_scalar(m->metallic, uv, tex);
    b->specular              = p_scalar(m->specular, uv, tex);
    b->roughness             = p_scalar(m->roughness, uv, tex);
    b->specular_tint         = p_scalar(m->specular_tint, uv, tex);
    b->anisotropy            = p_scalar(m->anisotropy, uv, tex);
    b->sheen                 = p_scalar(m->sheen, uv, tex);
    b->sheen_tint            = p_scalar(m->sheen_tint, uv, tex);
    b->clearcoat             = p_scalar(m->clearcoat, uv, tex);
    b->clearcoat_gloss       = p_scalar(m->clearcoat_gloss, uv, tex);
    b->ior                   = p_scalar(m->ior, uv, tex);
    b->specular_transmission = p_scalar(m->specular_transmission, uv, tex);
    p_color(m->base_emission, uv, tex, b->base_emission);
    b->emission_scale = m->emission_scale;
}

static inline float p_sqr(float x) { return x * x; }
static inline float p_lerp(float a, float b, float t) { return a * (1.f - t) + b * t; } /* glm::mix */
static inline float p_sat(float x) { return fminf(fmaxf(x, 0.f), 1.f); }
static inline float p_schlick(float c) { return powf(p_sat(1.f - c), 5.f); }
static inline float p_lum(const float c[3]) { return 0.2126f * c[0] + 0.7152f * c[1] + 0.0722f * c[2]; }
static inline int p_same_hemi(const float wo[3], const float wi[3], const float n[3]) { return dot_ptx(wo, n) * dot_ptx(wi, n) > 0.f; }
static void p_half(const float wi[3], const float wo[3], float h[3])
{
    float t[3] = {wi[0] + wo[0], wi[1] + wo[1], wi[2] + wo[2]};
    s_normalize(t, h);
}
static void p_reflect(const float i[3], const float n[3], float out[3])
{
    float d = dot_ptx(i, n);
    for (int k = 0; k < 3; k++)
        out[k] = i[k] - 2.f * n[k] * d;
}
static void p_to_world(const float h[3], const float t[3], const float b[3], const float n[3], float out[3])
{
    for (int i = 0; i < 3; i++)
        out[i] = s_bary(h[0], t[i], h[1], b[i], h[2], n[i]);
}
static float p_gtr1(float c, float alpha)
{
    if (alpha >= 1.f)
        return O_INV_PI;
    float a2 = alpha * alpha;
    return O_INV_PI * (a2 - 1.f) / (logf(a2) * (1.f + (a2 - 1.f) * c * c));
}
static float p_gtr2(float c, float alpha)
{
    float a2 = alpha * alpha;
    return O_INV_PI * a2 / p_sqr(1.f + (a2 - 1.f) * c * c);
}
static float p_gtr2_aniso(float hn, float hx, float hy, float ax, float ay)
{
    return O_INV_PI / (ax * ay * p_sqr(p_sqr(hx / ax) + p_sqr(hy / ay) + hn * hn));
}
static float p_smith(float n_dot_o, float alpha_g)
{
    float a = alpha_g * alpha_g, b = n_dot_o * n_dot_o;
    return 1.f / (n_dot_o + sqrtf(a + b - a * b));
}
static float p_smith_aniso(float n_dot_o, float ox, float oy, float ax, float ay)
{
    return 1.f / (n_dot_o + sqrtf(p_sqr(ox * ax) + p_sqr(oy * ay) + p_sqr(n_dot_o)));
}
static void p_spec_colour(const obsdf_t *m, float out[3])
{
    float lum = p_lum(m->base_color);
    for (int i = 0; i < 3; i++) {
        float tint = lum > 0.f ? m->base_color[i] * (1.f / lum) : 1.f;
        out[i]     = p_lerp(p_lerp(1.f, tint, m->specular_tint) * (m->specular * 0.08f), m->base_color[i], m->metallic);
    }
}
static float p_fresnel_dielectric(float ci, float eta_i, float eta_t)
{
    float g = p_sqr(eta_t) / p_sqr(eta_i) - 1.f + p_sqr(ci);
    if (g < 0.f)
        return 1.f;
    return 0.5f * p_sqr(g - ci) / p_sqr(g + ci) * (1.f + p_sqr(ci * (g + ci) - 1.f) / p_sqr(ci * (g - ci) + 1.f));
}
static void p_transmission_half(const float wo[3], const float wi[3], float eta_i, float eta_o, float h[3])
{
    float t[3];
    for (int i = 0; i < 3; i++)
        t[i] = wo[i] + wi[i] * eta_i * (1.f / eta_o);
    s_normalize(t, h);
}

static void p_disney_eval(const obsdf_t *m, const float n[3], const float wo[3], const float wi[3], const float tg[3],
                          const float bt[3], float out[3])
{
    out[0] = out[1] = out[2] = 0.f;
    if (!p_same_hemi(wo, wi, n)) {
        if (m->specular_transmission > 0.f) {
            float on = dot_ptx(wo, n), in = dot_ptx(wi, n);
            if (on == 0.f || in == 0.f)
                return;
            int entering = on > 0.f;
            float eta_o = entering ? 1.f : m->ior, eta_i = entering ? m->ior : 1.f, h[3];
            p_transmission_half(wo, wi, eta_i, eta_o, h);
            float alpha = fmaxf(0.001f, m->roughness * m->roughness);
            float d = p_gtr2(fabsf(dot_ptx(n, h)), alpha), f = p_fresnel_dielectric(fabsf(in), eta_o, eta_i);
            float g = p_smith(fabsf(in), alpha) * p_smith(fabsf(on), alpha);
            float ih = dot_ptx(wi, h), oh = dot_ptx(wo, h);
            float c = fabsf(oh) / fabsf(on) * fabsf(ih) / fabsf(in) * p_sqr(eta_o) / p_sqr(eta_o * oh + eta_i * ih);
            for (int i = 0; i < 3; i++)
                out[i] = m->base_color[i] * (c * (1.f - f) * g * d) * ((1.f - m->metallic) * m->specular_transmission);
        }
        return;
    }
    float h[3], spec[3];
    p_half(wi, wo, h);
    p_spec_colour(m, spec);
    /* clear coat */
    float coat = 0.25f * m->clearcoat * p_gtr1(dot_ptx(n, h), p_lerp(0.1f, 0.001f, m->clearcoat_gloss)) *
                 p_lerp(0.04f, 1.f, p_schlick(dot_ptx(wi, n))) * (p_smith(dot_ptx(n, wi), 0.25f) * p_smith(dot_ptx(n, wo), 0.25f));
    /* sheen + diffuse */
    float lum = p_lum(m->base_color), fh = p_schlick(dot_ptx(wi, h));
    float n_dot_o = fabsf(dot_ptx(wo, n)), n_dot_i = fabsf(dot_ptx(wi, n)), ih = dot_ptx(wi, h);
    float fd90 = 0.5f + 2.f * m->roughness * ih * ih;
    float dif  = O_INV_PI * p_lerp(1.f, fd90, p_schlick(n_dot_i)) * p_lerp(1.f, fd90, p_schlick(n_dot_o));
    /* gloss */
    float dg;
    if (m->anisotropy == 0.f) {
        float alpha = fmaxf(0.001f, m->roughness * m->roughness);
        dg = p_gtr2(dot_ptx(n, h), alpha) * (p_smith(dot_ptx(n, wi), alpha) * p_smith(dot_ptx(n, wo), alpha));
    } else {
        float aspect = sqrtf(1.f - m->anisotropy * 0.9f), a = m->roughness * m->roughness;
        float ax = fmaxf(0.001f, a / aspect), ay = fmaxf(0.001f, a * aspect);
        dg = p_gtr2_aniso(dot_ptx(n, h), fabsf(dot_ptx(h, tg)), fabsf(dot_ptx(h, bt)), ax, ay) *
             (p_smith_aniso(dot_ptx(n, wi), fabsf(dot_ptx(wi, tg)), fabsf(dot_ptx(wi, bt)), ax, ay) *
              p_smith_aniso(dot_ptx(n, wo), fabsf(dot_ptx(wo, tg)), fabsf(dot_ptx(wo, bt)), ax, ay));
    }
    float k = (1.f - m->metallic) * (1.f - m->specular_transmission);
    for (int i = 0; i < 3; i++) {
        float tint  = lum > 0.f ? m->base_color[i] * (1.f / lum) : 1.f;
        float sheen = p_lerp(1.f, tint, m->sheen_tint) * (fh * m->sheen);
        float gloss = p_lerp(spec[i], 1.f, fh) * dg;
        out[i]      = (m->base_color[i] * dif + sheen) * k + gloss + coat;
    }
}

static float p_disney_pdf(const obsdf_t *m, const float n[3], const float wo[3], const float wi[3], const float tg[3],
                          const float bt[3])
{
    float alpha = fmaxf(0.001f, m->roughness * m->roughness), aspect = sqrtf(1.f - m->anisotropy * 0.9f);
    float ax = fmaxf(0.001f, alpha / aspect), ay = fmaxf(0.001f, alpha * aspect);
    float dn = dot_ptx(wi, n), diffuse = dn > 0.f ? dn * O_INV_PI : 0.f;
    float coat = 0.f, micro = 0.f, trans = 0.f, n_comp = 3.f;
    if (p_same_hemi(wo, wi, n)) {
        float h[3];
        p_half(wi, wo, h);
        float c = dot_ptx(n, h), oh = dot_ptx(wo, h);
        coat = p_gtr1(c, p_lerp(0.1f, 0.001f, m->clearcoat_gloss)) * c / (4.f * oh);
        if (m->anisotropy == 0.f)
            micro = p_gtr2(c, alpha) * c / (4.f * oh);
        else
            micro = p_gtr2_aniso(c, fabsf(dot_ptx(h, tg)), fabsf(dot_ptx(h, bt)), ax, ay) * c / (4.f * oh);
    }
    if (m->specular_transmission > 0.f) {
        n_comp = 4.f;
        if (!p_same_hemi(wo, wi, n)) {
            int entering = dot_ptx(wo, n) > 0.f;
            float eta_o = entering ? 1.f : m->ior, eta_i = entering ? m->ior : 1.f, h[3];
            p_transmission_half(wo, wi, eta_i, eta_o, h);
            float c = fabsf(dot_ptx(n, h)), ih = dot_ptx(wi, h), oh = dot_ptx(wo, h);
            trans = p_gtr2(c, alpha) * c * fabsf(oh * p_sqr(eta_o) / p_sqr(eta_o * oh + eta_i * ih));
        }
    }
    return (diffuse + micro + trans + coat) / n_comp;
}

static void p_spherical(float sin_t, float cos_t, float phi, float out[3])
{
    out[0] = sin_t * cosf(phi), out[1] = sin_t * sinf(phi), out[2] = cos_t;
}

/* disney_sample_eval (material.cuh:517-592) */
static void p_disney_sample_eval(const obsdf_t *m, const float n[3], const float wo[3], const float tg[3], const float bt[3],
                                 uint32_t *rng, float wi[3], float *pdf, float f[3])
{
    float sx = onbvh_lcg_randomf(rng), sy = onbvh_lcg_randomf(rng);
    f[0] = f[1] = f[2] = 0.f;
    if (!(m->specular_transmission < 1.f)) { /* dielectric_sample_eval (material.cuh:473-512) */
        float eta = 1.f / m->ior, ci = dot_ptx(n, wo), F, bs = 0.f;
        float s2t = p_sqr(eta) * (1.f - ci * ci);
        if (s2t > 1.f) {
            F = 1.f;
        } else {
            float ln = sqrtf(1.f - s2t), e = 1.f / eta;
            float r1 = (ci - e * ln) / (ci + e * ln), r2 = (ln - e * ci) / (ln + e * ci);
            F = 0.5f * (p_sqr(r1) + p_sqr(r2));
        }
        wi[0] = wi[1] = wi[2] = 0.f;
        if (sy < F) {
            float mwo[3] = {-wo[0], -wo[1], -wo[2]};
            p_reflect(mwo, n, wi);
            if (dot_ptx(wi, n) > 0.f)
                bs = F;
        } else {
            float s2i = fmaxf(0.f, 1.f - ci * ci), s2 = eta * eta * s2i;
            if (s2 < 1.f) {
                float ct = sqrtf(1.f - s2), k = eta * ci - ct;
                for (int i = 0; i < 3; i++)
                    wi[i] = eta * -wo[i] + k * n[i];
                bs = 1.f - F;
            }
        }
        *pdf = 1.f;
        float inv = 1.f / fabsf(dot_ptx(n, wi));
        f[0] = f[1] = f[2] = bs * inv;
        return;
    }
    int component = (int)(onbvh_lcg_randomf(rng) * 3.f);
    component     = component < 0 ? 0 : component > 2 ? 2 : component;
    if (component == 0) {
        s_lambertian_dir(n, tg, bt, sx, sy, wi);
    } else {
        float hl[3], hn[3], h[3], mwo[3] = {-wo[0], -wo[1], -wo[2]};
        if (component == 1) {
            float alpha = fmaxf(0.001f, m->roughness * m->roughness);
            if (m->anisotropy == 0.f) {
                float c2 = (1.f - sy) / (1.f + (alpha * alpha - 1.f) * sy);
                p_spherical(1.f - c2, sqrtf(c2), 2.f * O_PI * sx, hl);
                s_normalize(hl, hn);
                p_to_world(hn, tg, bt, n, h);
            } else {
                float aspect = sqrtf(1.f - m->anisotropy * 0.9f);
                float ax = fmaxf(0.001f, alpha / aspect), ay = fmaxf(0.001f, alpha * aspect);
                float x = 2.f * O_PI * sx, r = sqrtf(sy / (1.f - sy)), cx = ax * cosf(x), cy = ay * sinf(x);
                for (int i = 0; i < 3; i++)
                    hl[i] = r * (cx * tg[i] + cy * bt[i]) + n[i];
                s_normalize(hl, h);
            }
        } else {
            float alpha = p_lerp(0.1f, 0.001f, m->clearcoat_gloss), a2 = alpha * alpha;
            float c2    = (1.f - powf(a2, 1.f - sy)) / (1.f - a2);
            p_spherical(1.f - c2, sqrtf(c2), 2.f * O_PI * sx, hl);
            s_normalize(hl, hn);
            p_to_world(hn, tg, bt, n, h);
        }
        p_reflect(mwo, h, wi);
        if (!p_same_hemi(wo, wi, n)) {
            *pdf  = 0.f;
            wi[0] = wi[1] = wi[2] = 0.f;
            return;
        }
    }
    *pdf = p_disney_pdf(m, n, wo, wi, tg, bt);
    p_disney_eval(m, n, wo, wi, tg, bt, f);
}

/* float4 texture, cudaFilterModeLinear + cudaAddressModeClamp + normalized coordinates */
static void p_tex2d_f32_clamp(const float *px, uint32_t w, uint32_t h, float u, float v, float out[4])
{
    float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float a = floorf((x - fx) * 256.f + 0.5f) / 256.f, b = floorf((y - fy) * 256.f + 0.5f) / 256.f;
    int i0 = (int)fx, j0 = (int)fy, i1 = i0 + 1, j1 = j0 + 1;
#define CLAMPI(i, n) ((i) < 0 ? 0 : (i) >= (int)(n) ? (int)(n) - 1 : (i))
    i0 = CLAMPI(i0, w), i1 = CLAMPI(i1, w), j0 = CLAMPI(j0, h), j1 = CLAMPI(j1, h);
#undef CLAMPI
    for (int ch = 0; ch < 4; ch++) {
        float t00 = px[4 * ((size_t)j0 * w + i0) + ch], t10 = px[4 * ((size_t)j0 * w + i1) + ch];
        float t01 = px[4 * ((size_t)j1 * w + i0) + ch], t11 = px[4 * ((size_t)j1 * w + i1) + ch];
        out[ch]   = (1.f - a) * (1.f - b) * t00 + a * (1.f - b) * t10 + (1.f - a) * b * t01 + a * b * t11;
    }
}

static void p_background(const opath_t *pc, const float d[3], float out[3])
{
    out[0] = out[1] = out[2] = 0.f;
    if (!pc->envmap_eval)
        return;
    if (pc->skylike_background) {
        float a = 0.5f * (d[1] + 1.f);
        out[0] = (1.f - a) * 1.f + a * 0.5f, out[1] = (1.f - a) * 1.f + a * 0.7f, out[2] = (1.f - a) * 1.f + a * 1.f;
        return;
    }
    float ang = pc->envmap_rotation_deg * 0.01745329251994329576923690768489f, c = cosf(ang), s = sinf(ang);
    float rx = c * d[0] + s * d[2], rz = c * d[2] - s * d[0];
    float eu = atan2f(rx, -rz) * O_INV_PI * 0.5f + 0.5f, ev = acosf(fminf(fmaxf(d[1], -1.f), 1.f)) * O_INV_PI;
    float val[4] = {1.f, 1.f, 1.f, 1.f};
    if (pc->envmap)
        p_tex2d_f32_clamp(pc->envmap, pc->env_w, pc->env_h, eu, ev, val);
    for (int i = 0; i < 3; i++)
        out[i] = fminf(100.f, val[i] * pc->envmap_strength);
}

static void path_one(const oscene_t *s, const onode_t *nbvh, const oconst_t *c, const onet_t *net, const opath_t *pc,
                     uint32_t pixel, int sample_idx, int rnd_offset, int mlp_mode, float radiance[3], ohit_t *primary,
                     uint64_t *cnt)
{
    float o[3], d[3], thr[3] = {1.f, 1.f, 1.f}, ill[3] = {0.f, 0.f, 0.f};
    uint32_t rng;
    onbvh_raygen(c, pixel, sample_idx, rnd_offset, o, d, &rng);
    for (uint32_t depth = 1;; depth++) {
        ohit_t hit;
        uint32_t lod = (pc->path_lod_switch && pc->num_lods_learned > 0 && depth > 1) ? pc->lod_switch : c->rendered_lod;
        trace_ray(s, nbvh, c, net, o, d, depth, lod, mlp_mode, &hit, cnt, &rng);
        cnt[4]++;
        if (depth == 1 && primary)
            *primary = hit;
        if (hit.valid_hit == 0 || depth == pc->max_depth) {
            if (hit.valid_hit == 0) {
                float bg[3];
                p_background(pc, d, bg);
                for (int i = 0; i < 3; i++)
                    ill[i] += thr[i] * bg[i];
            }
            break;
        }
        obsdf_t bsdf;
        p_default_bsdf(&bsdf);
        float wo[3] = {-d[0], -d[1], -d[2]}, its_n[3], pos[3];
        if (hit.valid_hit == 1) {
            for (int i = 0; i < 3; i++) {
                bsdf.base_color[i] = hit.albedo_or_ids[i];
                its_n[i]           = hit.normal_or_uv[i];
                pos[i]             = o[i] + hit.t * d[i];
            }
            bsdf.metallic = pc->neural_metallic, bsdf.roughness = pc->neural_roughness, bsdf.specular = pc->neural_specular;
        } else {
            uint32_t mesh, prim;
            memcpy(&mesh, &hit.albedo_or_ids[0], 4);
            memcpy(&prim, &hit.albedo_or_ids[1], 4);
            osampler_t sp;
            memset(&sp, 0, sizeof(sp));
            sp.blas          = s->blas[mesh];
            sp.normals       = pc->attr[mesh].normals;
            sp.texcoords     = pc->attr[mesh].texcoords;
            sp.material_ids  = pc->attr[mesh].material_ids;
            sp.material_map  = pc->attr[mesh].material_map;
            sp.num_materials = pc->attr[mesh].material_map ? pc->attr[mesh].num_materials : 0;
            sp.materials     = pc->materials;
            sp.textures      = pc->textures;
            oits_t its;
            s_compute_its(&sp, prim, hit.normal_or_uv[0], hit.normal_or_uv[1], &its);
            memcpy(pos, its.position, 12);
            memcpy(its_n, its.sh_normal, 12);
            int has_mat = its.material_id != -1 && pc->materials;
            if (c->ignore_neural_blas && mesh == pc->neural_bvh_idx) {
                if (has_mat)
                    p_color(pc->materials[its.material_id].base_color, its.uv, pc->textures, bsdf.base_color);
                bsdf.metallic = pc->neural_metallic, bsdf.roughness = pc->neural_roughness, bsdf.specular = pc->neural_specular;
            } else if (pc->force_constant_material) {
                if (has_mat)
                    p_color(pc->materials[its.material_id].base_color, its.uv, pc->textures, bsdf.base_color);
                bsdf.metallic = pc->constant_metallic, bsdf.roughness = pc->constant_roughness, bsdf.specular = pc->constant_specular;
            } else if (has_mat) {
                p_unpack_material(&pc->materials[its.material_id], its.uv, pc->textures, &bsdf);
            }
        }
        float shn[3], tg[3], bt[3];
        int flip = dot_ptx(wo, its_n) < 0.f; /* shading_frame */
        for (int i = 0; i < 3; i++)
            shn[i] = flip ? -its_n[i] : its_n[i];
        s_ortho_basis(tg, bt, shn);
        if ((depth == 1 || pc->emission_eval) && !(bsdf.base_emission[0] == 0.f && bsdf.base_emission[1] == 0.f && bsdf.base_emission[2] == 0.f))
            for (int i = 0; i < 3; i++)
                ill[i] += thr[i] * (bsdf.base_emission[i] * bsdf.emission_scale);
        float wi[3], pdf, f[3];
        p_disney_sample_eval(&bsdf, shn, wo, tg, bt, &rng, wi, &pdf, f);
        float cti = fabsf(dot_ptx(wi, shn));
        for (int i = 0; i < 3; i++)
            thr[i] *= f[i] * cti / pdf;
        if (pdf == 0.f || (f[0] == 0.f && f[1] == 0.f && f[2] == 0.f))
            break;
        if (pc->apply_russian_roulette && depth >= pc->rr_start_depth) { /* rr_terminate */
            float survival = fminf(1.f, fmaxf(fmaxf(thr[0], thr[1]), thr[2]));
            if (onbvh_lcg_randomf(&rng) < survival) {
                for (int i = 0; i < 3; i++)
                    thr[i] /= survival;
            } else {
                break;
            }
        }
        for (int i = 0; i < 3; i++) {
            o[i] = pos[i] + shn[i] * pc->normal_epsilon_offset;
            d[i] = wi[i];
        }
    }
    radiance[0] = ill[0], radiance[1] = ill[1], radiance[2] = ill[2];
}

/* closest-hit query of a ray batch: one trace_ray per ray (what nbvh_trace_rays computes) */
void onbvh_trace_rays(const oscene_t *s, const onode_t *nbvh, const oconst_t *c, const onet_t *net, uint32_t n, const float *o,
                      const float *d, uint32_t ray_depth, uint32_t lod, int mlp_mode, ohit_t *hits, uint64_t *counters, int n_threads)
{
    uint64_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#ifdef _OPENMP
    if (n_threads > 0)
        omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : c0, c1, c2, c3)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        uint64_t cnt[4] = {0, 0, 0, 0};
        trace_ray(s, nbvh, c, net, o + 3 * i, d + 3 * i, ray_depth, lod, mlp_mode, &hits[i], cnt, NULL);
        c0 += cnt[0], c1 += cnt[1], c2 += cnt[2], c3 += cnt[3];
    }
    if (counters) {
        counters[0] = c0, counters[1] = c1, counters[2] = c2, counters[3] = c3;
    }
}

void onbvh_render_path(const oscene_t *s, const onode_t *nbvh, const oconst_t *c, const onet_t *net, const opath_t *pc,
                       uint32_t pixel_begin, uint32_t pixel_end, uint32_t pixel_stride, int sample_idx, int rnd_offset,
                       int mlp_mode, float *radiance, ohit_t *primary_hits, uint64_t *counters, int n_threads)
{
    uint64_t c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
    if (pixel_stride == 0)
        pixel_stride = 1;
    const int64_t n = ((int64_t)pixel_end - (int64_t)pixel_begin + pixel_stride - 1) / pixel_stride;
#ifdef _OPENMP
    if (n_threads > 0)
        omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : c0, c1, c2, c3, c4)
    for (int64_t i = 0; i < n; i++) {
        uint64_t cnt[5] = {0, 0, 0, 0, 0};
        path_one(s, nbvh, c, net, pc, pixel_begin + (uint32_t)i * pixel_stride, sample_idx, rnd_offset, mlp_mode,
                 radiance + 3 * i, primary_hits ? &primary_hits[i] : NULL, cnt);
        c0 += cnt[0], c1 += cnt[1], c2 += cnt[2], c3 += cnt[3], c4 += cnt[4];
    }
    if (counters) {
        counters[0] = c0, counters[1] = c1, counters[2] = c2, counters[3] = c3, counters[4] = c4;
    }
}

/* ================================================================== online optimisation (a18-a21) */
/* include/cuda/neural/utils/loss_gradient.cuh:24-95 */
static void loss_and_gradient(float target, float prediction, uint32_t type, float *loss, float *grad)
{
    float diff = prediction - target;
    switch (type) {
    case 0: { /* RelativeL2 */
        float denom = prediction * prediction + 1e-2f;
        *loss = diff * diff / denom, *grad = 2.0f * diff / denom;
        break;
    }
    case 3: /* L1 */
        *loss = fabsf(diff), *grad = copysignf(1.0f, diff);
        break;
    case 2: { /* RelativeL1 */
        float scale = 1.0f / (fabsf(prediction) + 1e-2f);
        *loss = fabsf(diff) * scale, *grad = copysignf(scale, diff);
        break;
    }
    case 5: { /* BinaryCrossEntropy */
        const float eps = 1e-5f;
        float pred      = fminf(fmaxf(prediction, eps), 1.f - eps);
        *loss = -(target * logf(pred) + (1.f - target) * logf(1.f - pred));
        *grad = -(target - pred) / (pred - (pred * pred));
        break;
    }
    case 6: { /* LogL2 */
        float d = logf(1 + prediction) - logf(1 + target);
        *loss = d * d, *grad = 2.0f * d / (prediction + 1);
        break;
    }
    case 7: { /* LHalf */
        float s = sqrtf(fabsf(diff));
        *loss = s, *grad = 1.f / (2.f * s + 1e-5f);
        break;
    }
    default: /* L2 (and RelativeL2Luminance, which the reference's switch sends to L2) */
        *loss = diff * diff, *grad = 2.0f * diff;
        break;
    }
}

/* pt_composite_loss.cuh:19-123 */
float onbvh_loss_element(const oloss_t *lc, const float *target8, float prediction, uint32_t elem, uint32_t n_total,
                         float loss_scale, float *grad_out)
{
    const uint32_t depth0 = lc->vis_dim, normal0 = depth0 + lc->depth_dim, albedo0 = normal0 + lc->normal_dim,
                   end = albedo0 + lc->albedo_dim;
    float loss = 0.f, grad = 0.f;
    if (elem < end) {
        const float target = target8[elem];
        const int ignore   = lc->vis_masking ? (target8[0] < 0.5f) : 0;
        if (lc->vis_dim > 0 && elem == 0) {
            loss_and_gradient(target, prediction, lc->vis_loss, &loss, &grad);
            loss *= lc->vis_scale, grad *= lc->vis_scale;
        } else if (!ignore && elem >= depth0 && elem < normal0) {
            loss_and_gradient(target, prediction, lc->depth_loss, &loss, &grad);
            if (lc->depth_linear)
                grad *= (lc->depth_loss == 1 || lc->depth_loss == 0) ? 100.f : 10.f;
            loss *= lc->depth_scale, grad *= lc->depth_scale;
        } else if (!ignore && elem >= normal0 && elem < albedo0) {
            if (lc->normal_linear) {
                float n_length = 0.1f;
                float n_axis   = (prediction * 2.f - 1.f) / n_length;
                float lin      = n_axis * 0.5f + 0.5f;
                loss_and_gradient(target, lin, lc->normal_loss, &loss, &grad);
                grad *= 1.f / n_length;
            } else {
                loss_and_gradient(target, prediction, lc->normal_loss, &loss, &grad);
            }
            loss *= lc->normal_scale, grad *= lc->normal_scale;
        } else if (!ignore && elem >= albedo0 && elem < end) {
            if (lc->albedo_linear) {
                float n_length = 0.1f;
                float lin      = prediction / n_length;
                loss_and_gradient(target, lin, lc->albedo_loss, &loss, &grad);
                grad *= 1.f / n_length;
            } else {
                loss_and_gradient(target, prediction, lc->albedo_loss, &loss, &grad);
            }
            loss *= lc->albedo_scale, grad *= lc->albedo_scale;
        }
    }
    *grad_out = loss_scale * grad / (float)n_total;
    return loss / (float)n_total;
}

static inline float round_h(float x) { return onbvh_h2f(onbvh_f2h(x)); }

/* Scatter of one point's feature gradient into the grid (kernel_grid_backward, grid.h:215-320): same positions,
 * weights and indices as the forward pass; gradient += weight * dL/dfeature per corner. */
static void grid_backward_point(const onet_t *net, const uint32_t *offsets, const float *pt, const float *dfeat /* levels*F */,
                                float *grid_grads)
{
    const float l2   = log2f(net->per_level_scale);
    const uint32_t F = net->n_features;
    for (uint32_t l = 0; l < net->n_levels; l++) {
        const uint32_t hsize = offsets[l + 1] - offsets[l];
        const float scale    = grid_scale(l, l2, net->base_resolution);
        const uint32_t res   = grid_resolution(scale);
        float pos[3];
        uint32_t pg[3];
        for (int dim = 0; dim < 3; dim++) {
            float x  = fmaf(scale, pt[dim], 0.5f);
            float fl = floorf(x);
            pg[dim]  = (uint32_t)(int)fl;
            pos[dim] = x - fl;
        }
        for (uint32_t idx = 0; idx < 8; idx++) {
            float w = 1.f;
            uint32_t q[3];
            for (int dim = 0; dim < 3; dim++) {
                if ((idx & (1u << dim)) == 0) {
                    w *= 1.f - pos[dim];
                    q[dim] = pg[dim];
                } else {
                    w *= pos[dim];
                    q[dim] = pg[dim] + 1u;
                }
            }
            uint32_t stride = 1, index = 0;
            for (int dim = 0; dim < 3 && stride <= hsize; dim++) {
                index += q[dim] * stride;
                stride *= res;
            }
            if (hsize < stride)
                index = (q[0] * 1u) ^ (q[1] * 2654435761u) ^ (q[2] * 805459861u);
            index %= hsize;
            float *g = grid_grads + ((size_t)offsets[l] + index) * F;
            for (uint32_t f = 0; f < F; f++)
                g[f] += w * dfeat[l * F + f];
        }
    }
}

void onbvh_train_forward_backward(const onet_t *net, const oloss_t *lc, const float *in, const float *target,
                                  uint32_t batch, uint32_t global_batch, float loss_scale, float *L, float *grads,
                                  double *loss_sum)
{
    uint32_t offsets[64];
    onbvh_grid_n_params(net, offsets);
    const uint64_t n_mlp    = mlp_n_params(net);
    const uint32_t n_points = net->n_input / (net->n_levels * net->n_features);
    const uint32_t in_dims  = 3 * n_points;
    const uint32_t n_mats   = net->n_hidden_layers + 1;
    const uint32_t W        = net->n_neurons, K0 = net->n_input, NO = net->n_output_padded;
    const uint32_t n_total  = global_batch * net->n_output;
    const uint32_t ppf      = net->n_levels * net->n_features; /* features per point */
    memset(grads, 0, sizeof(float) * net->n_params);
    double lsum = 0.0;
    /* fp32 copies of the fp16 weights */
    float *Wf = (float *)malloc(sizeof(float) * n_mlp);
    for (uint64_t i = 0; i < n_mlp; i++)
        Wf[i] = onbvh_h2f(net->params[i]);
    double *wg = (double *)calloc(n_mlp, sizeof(double)); /* weight gradients accumulate in high precision */
    float *gg  = grads + n_mlp;
    for (uint32_t s = 0; s < batch; s++) {
        uint16_t feat[256];
        float act[9][256]; /* act[0] = features, act[l] = hidden l (post ReLU, fp16 valued) */
        float y[64], dz[256], dprev[256];
        onbvh_grid_encode(net, in + (size_t)s * in_dims, n_points, feat);
        for (uint32_t i = 0; i < K0; i++)
            act[0][i] = onbvh_h2f(feat[i]);
        const float *Wl = Wf;
        uint32_t in_w   = K0;
        for (uint32_t m = 0; m < n_mats; m++) {
            const int last       = m == n_mats - 1;
            const uint32_t out_w = last ? NO : W;
            for (uint32_t r = 0; r < out_w; r++) {
                double a = 0.0;
                for (uint32_t k = 0; k < in_w; k++)
                    a += (double)Wl[(size_t)r * in_w + k] * (double)act[m][k];
                float acc = (float)a;
                if (!last)
                    act[m + 1][r] = round_h(acc > 0.f ? acc : 0.f);
                else
                    y[r] = round_h(logistic(acc)); /* the training forward hands fp16 predictions to the loss */
            }
            Wl += (size_t)out_w * in_w;
            in_w = out_w;
        }
        /* loss + output-activation backward: dz = dL/dy * y(1-y), operands fp16-rounded */
        for (uint32_t r = 0; r < NO; r++) {
            float g = 0.f, lv = 0.f;
            if (r < net->n_output)
                lv = onbvh_loss_element(lc, target + (size_t)s * net->n_output, y[r], r, n_total, loss_scale, &g);
            if (L)
                L[(size_t)s * NO + r] = lv;
            lsum += lv;
            g     = round_h(g);
            dz[r] = round_h(g * (y[r] * (1.f - y[r])));
        }
        /* backward through the matrices, last to first */
        const float *Wend = Wf + n_mlp;
        uint32_t out_w    = NO;
        for (int m = (int)n_mats - 1; m >= 0; m--) {
            const uint32_t k_w = m == 0 ? K0 : W;
            Wend -= (size_t)out_w * k_w;
            double *wgm = wg + (Wend - Wf);
            for (uint32_t r = 0; r < out_w; r++)
                if (dz[r] != 0.f)
                    for (uint32_t k = 0; k < k_w; k++)
                        wgm[(size_t)r * k_w + k] += (double)dz[r] * (double)act[m][k];
            for (uint32_t k = 0; k < k_w; k++) {
                double a = 0.0;
                for (uint32_t r = 0; r < out_w; r++)
                    a += (double)Wend[(size_t)r * k_w + k] * (double)dz[r];
                float d = (float)a;
                if (m > 0)
                    d = act[m][k] > 0.f ? d : 0.f; /* ReLU' from the stored activation */
                dprev[k] = round_h(d);
            }
            memcpy(dz, dprev, sizeof(float) * k_w);
            out_w = k_w;
        }
        /* dz now holds dL/dfeatures (K0 values): scatter into the grid, point by point */
        for (uint32_t p = 0; p < n_points; p++)
            grid_backward_point(net, offsets, in + (size_t)s * in_dims + 3 * p, dz + p * ppf, gg);
    }
    for (uint64_t i = 0; i < n_mlp; i++)
        grads[i] = (float)wg[i];
    free(Wf);
    free(wg);
    if (loss_sum)
        *loss_sum = lsum;
}

/* exponential_decay.h:60-71: the factor is multiplied BEFORE the nested step when step >= decay_start and
 * (step - decay_start) % decay_interval == 0 */
float onbvh_learning_rate(const oopt_t *oc, uint32_t step)
{
    float factor = 1.f;
    for (uint32_t s = 0; s <= step; s++)
        if ((int64_t)s >= oc->decay_start && oc->decay_interval > 0 && ((int64_t)s - oc->decay_start) % oc->decay_interval == 0)
            factor *= oc->decay_base;
    return oc->learning_rate * factor;
}

/* optimizers/adam.h:48-123 with relative/absolute decay 0, no weight clipping, AdaBound off */
void onbvh_adam_step(const oopt_t *oc, uint32_t step, uint64_t n_params, uint64_t n_matrix, float loss_scale,
                     const float *grads, float *master, uint16_t *params, float *m, float *v, uint32_t *param_steps)
{
    const float base_lr = onbvh_learning_rate(oc, step);
    for (uint64_t i = 0; i < n_params; i++) {
        float gradient = round_h(grads[i]) / loss_scale;
        if (i >= n_matrix && gradient == 0.f)
            continue;
        const float w = master[i];
        if (i < n_matrix)
            gradient += oc->l2_reg * w;
        const float g2 = gradient * gradient;
        const float fm = m[i] = oc->beta1 * m[i] + (1.f - oc->beta1) * gradient;
        const float sm = v[i] = oc->beta2 * v[i] + (1.f - oc->beta2) * g2;
        const uint32_t t = ++param_steps[i];
        float lr = base_lr * (sqrtf(1.f - powf(oc->beta2, (float)t)) / (1.f - powf(oc->beta1, (float)t)));
        const float eff = lr / (sqrtf(sm) + oc->epsilon);
        const float nw  = w - eff * fm;
        master[i] = nw;
        params[i] = onbvh_f2h(nw);
    }
}

// shading.cuh -- surface interaction, materials and the Disney BSDF for the bounce part of the wavefront path
// (SURVEY 8 rows a15 / f1): replaces, for the path kernels in path.cu and the sample generator in sampler.cu,
//   BaseCudaBLAS::compute_its_data                 /root/reference/include/cuda/accel/bvh.cuh:53-96
//   unpack_material / unpack_albedo / unpack_normal /root/reference/include/cuda/base/material.cuh:24-89
//   Disney BSDF sample / eval / pdf                 /root/reference/include/cuda/base/material.cuh:91-592
//   Fr / Refract (dielectric lobe)                  /root/reference/include/cuda/base/wip_disney.cuh:128-152
//   shading_frame / rr_terminate                    /root/reference/include/cuda/utils/rendering.cuh:11-40
//   ortho_basis / reflect / luminance               /root/reference/include/cuda/utils/common.cuh:21-95
//   background_hit_eval                             /root/reference/src/cuda/neural/path/path_inference_kernels.cu:37-59
//
// Arithmetic: the geometric part (compute_its_data, ortho_basis, sample_lambertian_dir) keeps the contraction pattern that
// reproduces the reference's training targets bit for bit (tests/test_sampler.py).  The BSDF itself is evaluated in plain
// fp32 without contraction (-fmad=false); the reference build contracts freely there, so shaded radiance agrees to
// rounding only -- image parity is tolerance-based (tests/test_path_goldens.py).
#pragma once
#include "kernels.cuh"

namespace nbvh {

// PackedMaterial == Material + vtable pointer (include/base/material.h:52-87): 96 bytes
struct MaterialPod {
    unsigned long long vptr;
    float base_color[3], metallic;
    float specular, roughness, specular_tint, anisotropy;
    float sheen, sheen_tint, clearcoat, clearcoat_gloss;
    float base_emission[3], emission_scale;
    float normal[3], ior;
    float specular_transmission, pad;
};
static_assert(sizeof(MaterialPod) == 96, "PackedMaterial layout");

struct BlasAttr { // BaseCudaBLAS shading attributes (include/cuda/accel/bvh.cuh:11-23)
    const float *normals, *texcoords;
    const int *material_ids;
    const uint32_t *material_map;
    uint32_t num_materials;
};

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 ld3(const float *p, uint32_t i) { return V3{__ldg(p + 3u * i), __ldg(p + 3u * i + 1u), __ldg(p + 3u * i + 2u)}; }
__device__ __forceinline__ V3 sub3(V3 a, V3 b) { return V3{__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)}; }
__device__ __forceinline__ V3 add3(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 mul3(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ V3 scale3(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 neg3(V3 a) { return V3{-a.x, -a.y, -a.z}; }
__device__ __forceinline__ V3 splat3(float s) { return V3{s, s, s}; }
__device__ __forceinline__ bool is_zero3(V3 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }
// glm::lerp == glm::mix: x * (1 - a) + y * a
__device__ __forceinline__ float lerp1(float a, float b, float t) { return a * (1.f - t) + b * t; }
__device__ __forceinline__ V3 lerp3(V3 a, V3 b, float t) { return V3{lerp1(a.x, b.x, t), lerp1(a.y, b.y, t), lerp1(a.z, b.z, t)}; }
__device__ __forceinline__ float dot3(V3 a, V3 b) { return __fmaf_rn(a.z, b.z, __fmaf_rn(a.x, b.x, __fmul_rn(a.y, b.y))); }
__device__ __forceinline__ V3 cross3(V3 a, V3 b)
{
    return V3{__fmaf_rn(a.y, b.z, -__fmul_rn(b.y, a.z)), __fmaf_rn(a.z, b.x, -__fmul_rn(b.z, a.x)), __fmaf_rn(a.x, b.y, -__fmul_rn(b.x, a.y))};
}
// glm::normalize: v * inversesqrt(dot(v, v)), inversesqrt = rcp(sqrt) on the device (same pattern as camera_ray)
__device__ __forceinline__ V3 normalize3(V3 a)
{
    const float inv = __frcp_rn(__fsqrt_rn(dot3(a, a)));
    return V3{__fmul_rn(a.x, inv), __fmul_rn(a.y, inv), __fmul_rn(a.z, inv)};
}
// w*a + u*b + v*c per component: the reference's SASS evaluates fma(v,c, fma(w,a, u*b)) (found by matching the reference
// kernel's outputs bit for bit, tests/golden/ref_chess_sampler.npz)
__device__ __forceinline__ float bary(float w, float a, float u, float b, float v, float c)
{
    return __fmaf_rn(v, c, __fmaf_rn(w, a, __fmul_rn(u, b)));
}

// include/cuda/utils/common.cuh:21-36
__device__ __forceinline__ void ortho_basis(V3 &tangent, V3 &bitangent, V3 n)
{
    bitangent = v3(0.f, 0.f, 0.f);
    if (n.x < 0.6f && n.x > -0.6f)
        bitangent.x = 1.f;
    else if (n.y < 0.6f && n.y > -0.6f)
        bitangent.y = 1.f;
    else if (n.z < 0.6f && n.z > -0.6f)
        bitangent.z = 1.f;
    else
        bitangent.x = 1.f;
    tangent   = normalize3(cross3(bitangent, n));
    bitangent = normalize3(cross3(n, tangent));
}

// local (x, y, z) -> x*tangent + y*bitangent + z*n
__device__ __forceinline__ V3 to_world(V3 h, V3 tangent, V3 bitangent, V3 n)
{
    return V3{bary(h.x, tangent.x, h.y, bitangent.x, h.z, n.x), bary(h.x, tangent.y, h.y, bitangent.y, h.z, n.y),
              bary(h.x, tangent.z, h.y, bitangent.z, h.z, n.z)};
}

// include/cuda/base/material.cuh:101-120,193-200
__device__ __forceinline__ V3 sample_lambertian_dir(V3 n, V3 tangent, V3 bitangent, float ux, float uy)
{
    const float sx = __fmaf_rn(2.f, ux, -1.f), sy = __fmaf_rn(2.f, uy, -1.f);
    float dx = sx, dy = sy;
    if (!(sx == 0.f && sy == 0.f)) {
        float radius, theta;
        if (fabsf(sx) > fabsf(sy)) {
            radius = sx;
            theta  = __fmul_rn(0.78539818525314331f, __fdiv_rn(sy, sx));
        } else {
            radius = sy;
            theta  = __fsub_rn(1.5707963705062866f, __fmul_rn(0.78539818525314331f, __fdiv_rn(sx, sy)));
        }
        dx = __fmul_rn(radius, cosf(theta)), dy = __fmul_rn(radius, sinf(theta));
    }
    const float dz = __fsqrt_rn(fmaxf(0.f, __fsub_rn(__fsub_rn(1.f, __fmul_rn(dx, dx)), __fmul_rn(dy, dy))));
    return to_world(normalize3(v3(dx, dy, dz)), tangent, bitangent, n);
}

struct ItsData {
    V3 position, sh_normal;
    float u, v; // texture coordinates
    int material_id;
};

__device__ __forceinline__ float4 tex_fetch(const cudaTextureObject_t *textures, uint32_t id, float u, float v)
{
    return tex2D<float4>(textures[id], u, v);
}

// BaseCudaBLAS::compute_its_data (bvh.cuh:53-96)
__device__ __forceinline__ ItsData compute_its_data(const BlasConst &blas, const BlasAttr &attr, const MaterialPod *materials,
                                                    const cudaTextureObject_t *textures, const TriHit &h)
{
    ItsData its;
    const uint32_t i0 = __ldg(blas.indices + 3u * h.prim), i1 = __ldg(blas.indices + 3u * h.prim + 1u),
                   i2 = __ldg(blas.indices + 3u * h.prim + 2u);
    // upper_bound(material_map, num_materials, prim) - 1
    uint32_t left = 0, right = attr.num_materials;
    while (left < right) {
        const uint32_t mid = left + (right - left) / 2u;
        if (__ldg(attr.material_map + mid) <= h.prim)
            left = mid + 1u;
        else
            right = mid;
    }
    its.material_id = (attr.material_ids && left > 0u) ? __ldg(attr.material_ids + (left - 1u)) : -1;
    const V3 v0 = ld3(blas.vertices, i0), v1 = ld3(blas.vertices, i1), v2 = ld3(blas.vertices, i2);
    const float w = __fsub_rn(__fsub_rn(1.f, h.u), h.v);
    if (attr.normals) {
        const V3 n0 = ld3(attr.normals, i0), n1 = ld3(attr.normals, i1), n2 = ld3(attr.normals, i2);
        its.sh_normal = normalize3(v3(bary(w, n0.x, h.u, n1.x, h.v, n2.x), bary(w, n0.y, h.u, n1.y, h.v, n2.y), bary(w, n0.z, h.u, n1.z, h.v, n2.z)));
    } else {
        its.sh_normal = normalize3(cross3(sub3(v1, v0), sub3(v2, v0)));
    }
    its.position = v3(bary(w, v0.x, h.u, v1.x, h.v, v2.x), bary(w, v0.y, h.u, v1.y, h.v, v2.y), bary(w, v0.z, h.u, v1.z, h.v, v2.z));
    if (attr.texcoords) {
        const float *t = attr.texcoords;
        its.u = bary(w, __ldg(t + 2u * i0), h.u, __ldg(t + 2u * i1), h.v, __ldg(t + 2u * i2));
        its.v = bary(w, __ldg(t + 2u * i0 + 1u), h.u, __ldg(t + 2u * i1 + 1u), h.v, __ldg(t + 2u * i2 + 1u));
    } else {
        its.u = h.u, its.v = h.v;
    }
    if (its.material_id != -1 && materials) {
        // unpack_normal (material.cuh:66-75): tangent-space normal map -> world
        const uint32_t mask = __float_as_uint(materials[its.material_id].normal[0]);
        if (mask & 0x80000000u) {
            const float4 tx = tex_fetch(textures, mask & 0x1fffffffu, its.u, its.v);
            if (!(tx.x == -1.f && tx.y == -1.f && tx.z == -1.f)) {
                const V3 tsn = normalize3(v3(__fsub_rn(tx.x, 0.5f), __fsub_rn(tx.y, 0.5f), __fsub_rn(tx.z, 0.5f)));
                V3 tw, bw;
                ortho_basis(tw, bw, its.sh_normal);
                const V3 n = its.sh_normal; // mat3(tw, bw, n) * tsn = tw*x + bw*y + n*z
                its.sh_normal = v3(bary(tw.x, tsn.x, bw.x, tsn.y, n.x, tsn.z), bary(tw.y, tsn.x, bw.y, tsn.y, n.y, tsn.z),
                                   bary(tw.z, tsn.x, bw.z, tsn.y, n.z, tsn.z));
            }
        }
    }
    return its;
}

// ------------------------------------------------------------------------------------------------ materials
struct Bsdf { // Material (include/base/material.h:52-78) unpacked: every parameter a plain value
    V3 base_color;
    float metallic, specular, roughness, specular_tint, anisotropy, sheen, sheen_tint, clearcoat, clearcoat_gloss;
    V3 base_emission;
    float emission_scale, ior, specular_transmission;
};

__device__ __forceinline__ Bsdf default_bsdf()
{
    Bsdf b;
    b.base_color = splat3(0.9f);
    b.metallic = 0.f, b.specular = 0.f, b.roughness = 1.f, b.specular_tint = 0.f, b.anisotropy = 0.f;
    b.sheen = 0.f, b.sheen_tint = 0.f, b.clearcoat = 0.f, b.clearcoat_gloss = 0.f;
    b.base_emission  = splat3(0.f);
    b.emission_scale = 1.f, b.ior = 1.5f, b.specular_transmission = 0.f;
    return b;
}

// textured_scalar_param (material.cuh:24-33): negative floats are texture handles [31] flag, [30:29] channel, [28:0] id
__device__ __forceinline__ float scalar_param(float x, float u, float v, const cudaTextureObject_t *textures)
{
    const uint32_t mask = __float_as_uint(x);
    if (mask & 0x80000000u) {
        const float4 t      = tex_fetch(textures, mask & 0x1fffffffu, u, v);
        const uint32_t chan = (mask >> 29) & 3u;
        return chan == 0u ? t.x : chan == 1u ? t.y : chan == 2u ? t.z : t.w;
    }
    return x;
}

__device__ __forceinline__ V3 color_param(const float p[3], float u, float v, const cudaTextureObject_t *textures)
{
    const uint32_t mask = __float_as_uint(p[0]);
    if (mask & 0x80000000u) {
        const float4 t = tex_fetch(textures, mask & 0x1fffffffu, u, v);
        return v3(t.x, t.y, t.z);
    }
    return v3(p[0], p[1], p[2]);
}

// unpack_material (material.cuh:35-64)
__device__ __forceinline__ Bsdf unpack_material(const MaterialPod &p, float u, float v, const cudaTextureObject_t *textures)
{
    Bsdf m;
    m.base_color            = color_param(p.base_color, u, v, textures);
    m.metallic              = scalar_param(p.metallic, u, v, textures);
    m.specular              = scalar_param(p.specular, u, v, textures);
    m.roughness             = scalar_param(p.roughness, u, v, textures);
    m.specular_tint         = scalar_param(p.specular_tint, u, v, textures);
    m.anisotropy            = scalar_param(p.anisotropy, u, v, textures);
    m.sheen                 = scalar_param(p.sheen, u, v, textures);
    m.sheen_tint            = scalar_param(p.sheen_tint, u, v, textures);
    m.clearcoat             = scalar_param(p.clearcoat, u, v, textures);
    m.clearcoat_gloss       = scalar_param(p.clearcoat_gloss, u, v, textures);
    m.ior                   = scalar_param(p.ior, u, v, textures);
    m.specular_transmission = scalar_param(p.specular_transmission, u, v, textures);
    m.base_emission         = color_param(p.base_emission, u, v, textures);
    m.emission_scale        = p.emission_scale;
    return m;
}

// ------------------------------------------------------------------------------------------------ Disney BSDF
constexpr float kPi    = 3.14159265358979323846f;
constexpr float kInvPi = 0.31830988618379067154f;

__device__ __forceinline__ float sqr1(float x) { return x * x; }
__device__ __forceinline__ float saturate1(float x) { return fminf(fmaxf(x, 0.f), 1.f); }
__device__ __forceinline__ float luminance3(V3 c) { return 0.2126f * c.x + 0.7152f * c.y + 0.0722f * c.z; }
__device__ __forceinline__ bool same_hemisphere(V3 wo, V3 wi, V3 n) { return dot3(wo, n) * dot3(wi, n) > 0.f; }
__device__ __forceinline__ V3 reflect3(V3 i, V3 n) // i - 2 n dot(i, n)
{
    const float d = dot3(i, n);
    return V3{i.x - 2.f * n.x * d, i.y - 2.f * n.y * d, i.z - 2.f * n.z * d};
}
__device__ __forceinline__ float schlick_weight(float cos_theta) { return powf(saturate1(1.f - cos_theta), 5.f); }

__device__ __forceinline__ float gtr_1(float cos_theta_h, float alpha)
{
    if (alpha >= 1.f)
        return kInvPi;
    const float a2 = alpha * alpha;
    return kInvPi * (a2 - 1.f) / (logf(a2) * (1.f + (a2 - 1.f) * cos_theta_h * cos_theta_h));
}
__device__ __forceinline__ float gtr_2(float cos_theta_h, float alpha)
{
    const float a2 = alpha * alpha;
    return kInvPi * a2 / sqr1(1.f + (a2 - 1.f) * cos_theta_h * cos_theta_h);
}
__device__ __forceinline__ float gtr_2_aniso(float h_dot_n, float h_dot_x, float h_dot_y, float ax, float ay)
{
    return kInvPi / (ax * ay * sqr1(sqr1(h_dot_x / ax) + sqr1(h_dot_y / ay) + h_dot_n * h_dot_n));
}
__device__ __forceinline__ float smith_ggx(float n_dot_o, float alpha_g)
{
    const float a = alpha_g * alpha_g, b = n_dot_o * n_dot_o;
    return 1.f / (n_dot_o + sqrtf(a + b - a * b));
}
__device__ __forceinline__ float smith_ggx_aniso(float n_dot_o, float o_dot_x, float o_dot_y, float ax, float ay)
{
    return 1.f / (n_dot_o + sqrtf(sqr1(o_dot_x * ax) + sqr1(o_dot_y * ay) + sqr1(n_dot_o)));
}

__device__ __forceinline__ V3 spherical_dir(float sin_theta, float cos_theta, float phi)
{
    return v3(sin_theta * cosf(phi), sin_theta * sinf(phi), cos_theta);
}
// material.cuh:203-224 (the reference passes 1 - cos^2 as "sin_theta" and normalises afterwards; kept)
__device__ __forceinline__ V3 sample_gtr_1_h(V3 n, V3 tangent, V3 bitangent, float alpha, float sx, float sy)
{
    const float phi_h = 2.f * kPi * sx, a2 = alpha * alpha;
    const float c2    = (1.f - powf(a2, 1.f - sy)) / (1.f - a2);
    return to_world(normalize3(spherical_dir(1.f - c2, sqrtf(c2), phi_h)), tangent, bitangent, n);
}
__device__ __forceinline__ V3 sample_gtr_2_h(V3 n, V3 tangent, V3 bitangent, float alpha, float sx, float sy)
{
    const float phi_h = 2.f * kPi * sx;
    const float c2    = (1.f - sy) / (1.f + (alpha * alpha - 1.f) * sy);
    return to_world(normalize3(spherical_dir(1.f - c2, sqrtf(c2), phi_h)), tangent, bitangent, n);
}
__device__ __forceinline__ V3 sample_gtr_2_aniso_h(V3 n, V3 tangent, V3 bitangent, float ax, float ay, float sx, float sy)
{
    const float x = 2.f * kPi * sx, r = sqrtf(sy / (1.f - sy));
    const float cx = ax * cosf(x), cy = ay * sinf(x);
    return normalize3(v3(r * (cx * tangent.x + cy * bitangent.x) + n.x, r * (cx * tangent.y + cy * bitangent.y) + n.y,
                         r * (cx * tangent.z + cy * bitangent.z) + n.z));
}

__device__ __forceinline__ float lambertian_pdf(V3 wi, V3 n)
{
    const float d = dot3(wi, n);
    return d > 0.f ? d * kInvPi : 0.f;
}
__device__ __forceinline__ float gtr_1_pdf(V3 wo, V3 wi, V3 n, float alpha)
{
    if (!same_hemisphere(wo, wi, n))
        return 0.f;
    const V3 w_h  = normalize3(add3(wi, wo));
    const float c = dot3(n, w_h);
    return gtr_1(c, alpha) * c / (4.f * dot3(wo, w_h));
}
__device__ __forceinline__ float gtr_2_pdf(V3 wo, V3 wi, V3 n, float alpha)
{
    if (!same_hemisphere(wo, wi, n))
        return 0.f;
    const V3 w_h  = normalize3(add3(wi, wo));
    const float c = dot3(n, w_h);
    return gtr_2(c, alpha) * c / (4.f * dot3(wo, w_h));
}
__device__ __forceinline__ float gtr_2_transmission_pdf(V3 wo, V3 wi, V3 n, float alpha, float ior)
{
    if (same_hemisphere(wo, wi, n))
        return 0.f;
    const bool entering = dot3(wo, n) > 0.f;
    const float eta_o = entering ? 1.f : ior, eta_i = entering ? ior : 1.f;
    const V3 w_h        = normalize3(add3(wo, scale3(scale3(wi, eta_i), 1.f / eta_o)));
    const float c       = fabsf(dot3(n, w_h));
    const float i_dot_h = dot3(wi, w_h), o_dot_h = dot3(wo, w_h);
    const float dwh_dwi = o_dot_h * sqr1(eta_o) / sqr1(eta_o * o_dot_h + eta_i * i_dot_h);
    return gtr_2(c, alpha) * c * fabsf(dwh_dwi);
}
__device__ __forceinline__ float gtr_2_aniso_pdf(V3 wo, V3 wi, V3 n, V3 tangent, V3 bitangent, float ax, float ay)
{
    if (!same_hemisphere(wo, wi, n))
        return 0.f;
    const V3 w_h  = normalize3(add3(wi, wo));
    const float c = dot3(n, w_h);
    return gtr_2_aniso(c, fabsf(dot3(w_h, tangent)), fabsf(dot3(w_h, bitangent)), ax, ay) * c / (4.f * dot3(wo, w_h));
}

// material.cuh:142-151
__device__ __forceinline__ float fresnel_dielectric(float cos_theta_i, float eta_i, float eta_t)
{
    const float g = sqr1(eta_t) / sqr1(eta_i) - 1.f + sqr1(cos_theta_i);
    if (g < 0.f)
        return 1.f;
    return 0.5f * sqr1(g - cos_theta_i) / sqr1(g + cos_theta_i) *
           (1.f + sqr1(cos_theta_i * (g + cos_theta_i) - 1.f) / sqr1(cos_theta_i * (g - cos_theta_i) + 1.f));
}

__device__ __forceinline__ V3 spec_colour(const Bsdf &m)
{
    const float lum = luminance3(m.base_color);
    const V3 tint   = lum > 0.f ? scale3(m.base_color, 1.f / lum) : splat3(1.f);
    return lerp3(scale3(lerp3(splat3(1.f), tint, m.specular_tint), m.specular * 0.08f), m.base_color, m.metallic);
}

__device__ __forceinline__ V3 disney_diffuse(const Bsdf &m, V3 n, V3 wo, V3 wi)
{
    const V3 w_h        = normalize3(add3(wi, wo));
    const float n_dot_o = fabsf(dot3(wo, n)), n_dot_i = fabsf(dot3(wi, n)), i_dot_h = dot3(wi, w_h);
    const float fd90    = 0.5f + 2.f * m.roughness * i_dot_h * i_dot_h;
    const float fi = schlick_weight(n_dot_i), fo = schlick_weight(n_dot_o);
    return scale3(m.base_color, kInvPi * lerp1(1.f, fd90, fi) * lerp1(1.f, fd90, fo));
}
__device__ __forceinline__ V3 disney_microfacet_isotropic(const Bsdf &m, V3 n, V3 wo, V3 wi)
{
    const V3 w_h      = normalize3(add3(wi, wo));
    const float alpha = fmaxf(0.001f, m.roughness * m.roughness);
    const float d     = gtr_2(dot3(n, w_h), alpha);
    const V3 f        = lerp3(spec_colour(m), splat3(1.f), schlick_weight(dot3(wi, w_h)));
    const float g     = smith_ggx(dot3(n, wi), alpha) * smith_ggx(dot3(n, wo), alpha);
    return scale3(f, d * g);
}
__device__ __forceinline__ V3 disney_microfacet_anisotropic(const Bsdf &m, V3 n, V3 wo, V3 wi, V3 tangent, V3 bitangent)
{
    const V3 w_h       = normalize3(add3(wi, wo));
    const float aspect = sqrtf(1.f - m.anisotropy * 0.9f), a = m.roughness * m.roughness;
    const float ax = fmaxf(0.001f, a / aspect), ay = fmaxf(0.001f, a * aspect);
    const float d = gtr_2_aniso(dot3(n, w_h), fabsf(dot3(w_h, tangent)), fabsf(dot3(w_h, bitangent)), ax, ay);
    const V3 f    = lerp3(spec_colour(m), splat3(1.f), schlick_weight(dot3(wi, w_h)));
    const float g = smith_ggx_aniso(dot3(n, wi), fabsf(dot3(wi, tangent)), fabsf(dot3(wi, bitangent)), ax, ay) *
                    smith_ggx_aniso(dot3(n, wo), fabsf(dot3(wo, tangent)), fabsf(dot3(wo, bitangent)), ax, ay);
    return scale3(f, d * g);
}
__device__ __forceinline__ V3 disney_microfacet_transmission_isotropic(const Bsdf &m, V3 n, V3 wo, V3 wi)
{
    const float o_dot_n = dot3(wo, n), i_dot_n = dot3(wi, n);
    if (o_dot_n == 0.f || i_dot_n == 0.f)
        return splat3(0.f);
    const bool entering = o_dot_n > 0.f;
    const float eta_o = entering ? 1.f : m.ior, eta_i = entering ? m.ior : 1.f;
    const V3 w_h      = normalize3(add3(wo, scale3(scale3(wi, eta_i), 1.f / eta_o)));
    const float alpha = fmaxf(0.001f, m.roughness * m.roughness);
    const float d     = gtr_2(fabsf(dot3(n, w_h)), alpha);
    const float f     = fresnel_dielectric(fabsf(i_dot_n), eta_o, eta_i);
    const float g     = smith_ggx(fabsf(i_dot_n), alpha) * smith_ggx(fabsf(o_dot_n), alpha);
    const float i_dot_h = dot3(wi, w_h), o_dot_h = dot3(wo, w_h);
    const float c = fabsf(o_dot_h) / fabsf(o_dot_n) * fabsf(i_dot_h) / fabsf(i_dot_n) * sqr1(eta_o) / sqr1(eta_o * o_dot_h + eta_i * i_dot_h);
    return scale3(m.base_color, c * (1.f - f) * g * d);
}
__device__ __forceinline__ float disney_clear_coat(const Bsdf &m, V3 n, V3 wo, V3 wi)
{
    const V3 w_h      = normalize3(add3(wi, wo));
    const float alpha = lerp1(0.1f, 0.001f, m.clearcoat_gloss);
    const float d     = gtr_1(dot3(n, w_h), alpha);
    const float f     = lerp1(0.04f, 1.f, schlick_weight(dot3(wi, n)));
    const float g     = smith_ggx(dot3(n, wi), 0.25f) * smith_ggx(dot3(n, wo), 0.25f);
    return 0.25f * m.clearcoat * d * f * g;
}
__device__ __forceinline__ V3 disney_sheen(const Bsdf &m, V3 wo, V3 wi)
{
    const V3 w_h    = normalize3(add3(wi, wo));
    const float lum = luminance3(m.base_color);
    const V3 tint   = lum > 0.f ? scale3(m.base_color, 1.f / lum) : splat3(1.f);
    return scale3(lerp3(splat3(1.f), tint, m.sheen_tint), schlick_weight(dot3(wi, w_h)) * m.sheen);
}

// disney_eval (material.cuh:418-441)
__device__ __forceinline__ V3 disney_eval(const Bsdf &m, V3 n, V3 wo, V3 wi, V3 tangent, V3 bitangent)
{
    if (!same_hemisphere(wo, wi, n)) {
        if (m.specular_transmission > 0.f)
            return scale3(disney_microfacet_transmission_isotropic(m, n, wo, wi), (1.f - m.metallic) * m.specular_transmission);
        return splat3(0.f);
    }
    const float coat = disney_clear_coat(m, n, wo, wi);
    const V3 sheen = disney_sheen(m, wo, wi), diffuse = disney_diffuse(m, n, wo, wi);
    const V3 gloss = m.anisotropy == 0.f ? disney_microfacet_isotropic(m, n, wo, wi)
                                         : disney_microfacet_anisotropic(m, n, wo, wi, tangent, bitangent);
    const float k = (1.f - m.metallic) * (1.f - m.specular_transmission);
    return V3{(diffuse.x + sheen.x) * k + gloss.x + coat, (diffuse.y + sheen.y) * k + gloss.y + coat, (diffuse.z + sheen.z) * k + gloss.z + coat};
}

// disney_pdf (material.cuh:443-471)
__device__ __forceinline__ float disney_pdf(const Bsdf &m, V3 n, V3 wo, V3 wi, V3 tangent, V3 bitangent)
{
    const float alpha  = fmaxf(0.001f, m.roughness * m.roughness);
    const float aspect = sqrtf(1.f - m.anisotropy * 0.9f);
    const float ax = fmaxf(0.001f, alpha / aspect), ay = fmaxf(0.001f, alpha * aspect);
    const float diffuse    = lambertian_pdf(wi, n);
    const float clear_coat = gtr_1_pdf(wo, wi, n, lerp1(0.1f, 0.001f, m.clearcoat_gloss));
    float n_comp = 3.f, microfacet_transmission = 0.f;
    const float microfacet = m.anisotropy == 0.f ? gtr_2_pdf(wo, wi, n, alpha) : gtr_2_aniso_pdf(wo, wi, n, tangent, bitangent, ax, ay);
    if (m.specular_transmission > 0.f) {
        n_comp                  = 4.f;
        microfacet_transmission = gtr_2_transmission_pdf(wo, wi, n, alpha, m.ior);
    }
    return (diffuse + microfacet + microfacet_transmission + clear_coat) / n_comp;
}

// Fr / Refract (wip_disney.cuh:128-152) and dielectric_sample_eval (material.cuh:473-512)
__device__ __forceinline__ float fresnel_fr(float v_dot_n, float eio)
{
    const float sin2_t = sqr1(eio) * (1.f - v_dot_n * v_dot_n);
    if (sin2_t > 1.f)
        return 1.f;
    const float l_dot_n = sqrtf(1.f - sin2_t), eta = 1.f / eio;
    const float r1 = (v_dot_n - eta * l_dot_n) / (v_dot_n + eta * l_dot_n);
    const float r2 = (l_dot_n - eta * v_dot_n) / (l_dot_n + eta * v_dot_n);
    return 0.5f * (sqr1(r1) + sqr1(r2));
}
__device__ __forceinline__ V3 dielectric_sample_eval(const Bsdf &m, V3 n, V3 wo, float sample_y, V3 &wi, float &pdf)
{
    const float eta = 1.f / m.ior;
    V3 bsdf         = splat3(0.f);
    const float cos_i = dot3(n, wo);
    const float F     = fresnel_fr(cos_i, eta);
    wi = splat3(0.f); // the reference leaves wi unset on total internal reflection; zero keeps the path deterministic
    if (sample_y < F) {
        wi = reflect3(neg3(wo), n);
        if (dot3(wi, n) > 0.f)
            bsdf = splat3(F);
    } else {
        const float sin2_i = fmaxf(0.f, 1.f - cos_i * cos_i), sin2_t = eta * eta * sin2_i;
        if (sin2_t < 1.f) {
            const float cos_t = sqrtf(1.f - sin2_t), k = eta * cos_i - cos_t;
            wi   = v3(eta * -wo.x + k * n.x, eta * -wo.y + k * n.y, eta * -wo.z + k * n.z);
            bsdf = splat3(1.f - F);
        }
    }
    pdf = 1.f;
    return scale3(bsdf, 1.f / fabsf(dot3(n, wi)));
}

// disney_sample_eval (material.cuh:517-592): draws 3 random numbers (2 when the material is fully transmissive)
__device__ __forceinline__ V3 disney_sample_eval(const Bsdf &m, V3 n, V3 wo, V3 tangent, V3 bitangent, uint32_t &rng, V3 &wi, float &pdf)
{
    const float sx = lcg_randomf(rng), sy = lcg_randomf(rng);
    if (!(m.specular_transmission < 1.f))
        return dielectric_sample_eval(m, n, wo, sy, wi, pdf);
    int component = (int)(lcg_randomf(rng) * 3.f);
    component     = min(max(component, 0), 2);
    if (component == 0) {
        wi = sample_lambertian_dir(n, tangent, bitangent, sx, sy);
    } else {
        V3 w_h;
        if (component == 1) {
            const float alpha = fmaxf(0.001f, m.roughness * m.roughness);
            if (m.anisotropy == 0.f) {
                w_h = sample_gtr_2_h(n, tangent, bitangent, alpha, sx, sy);
            } else {
                const float aspect = sqrtf(1.f - m.anisotropy * 0.9f);
                w_h = sample_gtr_2_aniso_h(n, tangent, bitangent, fmaxf(0.001f, alpha / aspect), fmaxf(0.001f, alpha * aspect), sx, sy);
            }
        } else {
            w_h = sample_gtr_1_h(n, tangent, bitangent, lerp1(0.1f, 0.001f, m.clearcoat_gloss), sx, sy);
        }
        wi = reflect3(neg3(wo), w_h);
        if (!same_hemisphere(wo, wi, n)) { // invalid reflection: the path ends
            pdf = 0.f;
            wi  = splat3(0.f);
            return splat3(0.f);
        }
    }
    pdf = disney_pdf(m, n, wo, wi, tangent, bitangent);
    return disney_eval(m, n, wo, wi, tangent, bitangent);
}

// shading_frame (rendering.cuh:11-23)
__device__ __forceinline__ void shading_frame(V3 wo, V3 its_normal, V3 &sh_normal, V3 &tangent, V3 &bitangent)
{
    sh_normal = dot3(wo, its_normal) < 0.f ? neg3(its_normal) : its_normal;
    ortho_basis(tangent, bitangent, sh_normal);
}

// rr_terminate (rendering.cuh:25-40); draws one random number when depth >= rr_start_depth
__device__ __forceinline__ bool rr_terminate(uint32_t depth, V3 &throughput, uint32_t rr_start_depth, uint32_t &rng)
{
    if (depth >= rr_start_depth) {
        const float survival = fminf(1.f, fmaxf(fmaxf(throughput.x, throughput.y), throughput.z));
        if (lcg_randomf(rng) < survival)
            throughput = v3(throughput.x / survival, throughput.y / survival, throughput.z / survival);
        else
            return true;
    }
    return false;
}

// background_hit_eval (path_inference_kernels.cu:37-59)
__device__ __forceinline__ V3 background_eval(const BackgroundConst &b, float dx, float dy, float dz)
{
    if (!b.envmap_eval)
        return splat3(0.f);
    if (b.skylike_background) {
        const float a = 0.5f * (dy + 1.f);
        return v3((1.f - a) * 1.f + a * 0.5f, (1.f - a) * 1.f + a * 0.7f, (1.f - a) * 1.f + a * 1.f);
    }
    const float rx = b.rot_cos * dx + b.rot_sin * dz, rz = b.rot_cos * dz - b.rot_sin * dx;
    const float eu = atan2f(rx, -rz) * kInvPi * 0.5f + 0.5f;
    const float ev = acosf(fminf(fmaxf(dy, -1.f), 1.f)) * kInvPi;
    float4 val     = make_float4(1.f, 1.f, 1.f, 1.f);
    if (b.envmap)
        val = tex2D<float4>(b.envmap, eu, ev);
    return v3(fminf(100.f, val.x * b.strength), fminf(100.f, val.y * b.strength), fminf(100.f, val.z * b.strength));
}

} // namespace nbvh

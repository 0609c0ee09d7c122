// path.cuh -- parameter block of the multi-bounce path kernels (path.cu), shared with context.cu
#pragma once
#include "shading.cuh"

namespace nbvh {

struct PathConst {
    BlasAttr attr[kMaxBlas];       // shading attributes of every BLAS (CudaSceneData.tlas.mesh_bvhs[i]) ...
    const BlasAttr *attr_ext;      // ... or a device table when the scene has more than kMaxBlas meshes
    const MaterialPod *materials;  // CudaSceneData.materials
    const cudaTextureObject_t *textures;
    uint32_t neural_bvh_idx;       // bvh_const.neural_bvh_idx
    // path_module_const (path_module.h:217-254)
    float normal_epsilon_offset;
    uint32_t apply_russian_roulette, rr_start_depth;
    float neural_roughness, neural_specular, neural_metallic;
    // scene_constants (cuda_backend.h:58-72)
    uint32_t force_constant_material, emission_eval;
    float constant_roughness, constant_specular, constant_metallic;
};

__device__ __forceinline__ BlasAttr path_attr(const PathConst &pc, uint32_t i) { return pc.attr_ext ? pc.attr_ext[i] : pc.attr[i]; }

// path lists of one shade pass: in/hits_in (length *count_in, or the pixel count when null) -> survivors appended to out
struct PathIo {
    const PathState *in;
    const PathHit *hits_in;
    PathState *out;
    const uint32_t *count_in;
    uint32_t *count_out;
};

constexpr uint32_t kPathCountersPerDepth = 4; // neural queue count, queue cursor, path-list length, pad

cudaError_t launch_render_path(const RenderParams &rp, const PathConst &pc, const NetConst &nc, const NbvhConst &nbvh_bounce,
                               PathState *const paths[2], PathHit *const hits[2], uint32_t *counters, int num_sms, cudaStream_t stream);

cudaError_t launch_trace_rays(const RenderParams &rp, const NetConst &nc, const float *o, const float *d, uint32_t n, uint32_t ray_depth,
                              PathState *paths, PathHit *hits, uint32_t *counters, nbvh_hit_record *out, int num_sms, cudaStream_t stream);

cudaError_t launch_image_error(const float4 *accum, const float4 *reference, uint32_t n, uint32_t spp, int mode, float *sum_dev, uchar4 *ldr,
                               cudaStream_t stream);

} // namespace nbvh

// TEST INFRASTRUCTURE (golden generator + reference GPU arm) -- runs the reference's OWN wavefront path headless on
// the GPU box.  Nothing here is product code and no reference source is copied: this file only *includes* the
// reference's headers where they lie under /root/reference and *calls* its device functions and __global__ kernels
// (compiled from /root/reference/src/cuda/neural/{neural_bvh_module_kernels,neural_bvh_renderer_kernels}.cu,
// path/path_inference_kernels.cu and the four *_constants.cu by oracle/Makefile.tcnn `render_harness`), plus its
// tiny-cuda-nn fork for network->inference.  The host loop below drives those kernels in the order
// NeuralPathModule::neural_path_tracing / neural_bvh_traversal do (/root/reference/src/cuda/neural/path/path_module.cu:54-245);
// the GLFW/GL display, CudaBackend and the NeuralBVHModule class (which cannot run without a GL context,
// render_app.cpp:307-312) are replaced by plain cudaMalloc buffers and cudaArray surfaces.
//
// usage: ref_render_harness <workdir> <W> <H> <max_depth> <log2_hashmap> <n_timed_frames>
//        ref_render_harness <workdir> <W> <H> <max_depth> <log2_hashmap> <n_timed_frames> timing
//          (bench.py's live reference-GPU arm: the reference's frame loop only, 3 warm-up + <n_timed_frames> timed frames,
//           no parity hooks and no dumps besides ref_frame.json; shading inputs are read when present, as in `path` mode)
//        ref_render_harness <workdir> <W> <H> <max_depth> <log2_hashmap> <n_timed_frames> path <spp>
//          (hybrid neural path tracing with shading: also reads materials.bin, tex<k>.rgba8 + tex<k>.dims, envmap.f32 + envmap.dims,
//           mesh<i>_{normals.f32,texcoords.f32,material_ids.i32,material_map.u32}; writes path_accum_1spp.f32 / path_accum.f32 =
//           the RGBA32F accumulation surface after 1 and after <spp> samples)
//        ref_render_harness <workdir> <W> <H> <max_depth> <log2_hashmap> 0 sampler <batch> <log2_max_traversals> <steps> [max_loss]
//          (K10: granular_sample_bvh_network_input_data; also reads mesh<i>_{normals.f32,texcoords.f32,material_ids.i32,
//           material_map.u32}, materials.bin, tex<k>.rgba8 + tex<k>.dims, tree_cut.bin; writes sampler_{in.f32,tgt.f32,nodes.u32})
//   reads  <workdir>/const_{renderer,scene_bbox,module,pt_output,path,camera,scene}.bin (POD blobs, sizes checked
//          against the reference's own sizeof), nbvh_nodes.bin, tlas_nodes.bin, tlas_indices.u32,
//          mesh<i>_{nodes.bin,prim_ids.u64,vertices.f32,indices.u32} (i = 0..n_meshes-1), params.u16
//   writes rays.f32 (N x 6), iv_count.u32, iv_node.u32, iv_mint.f32, iv_maxt.f32 (N x 32: the unconditional interval
//          list of NeuralBVH2::intersect), blas_hit.bin (N x {t,u,v,prim,mesh}: CudaTLAS::intersect<false> over the
//          non-neural meshes), step0_in.f32 / step0_out.f32 / step0_pixel.u32 (first wavefront step's network I/O),
//          hits.bin (N x {valid_hit,t,normal_or_uv[3],albedo_or_ids[3]}: final HybridIntersectionResult per pixel),
//          ref_frame.json (frame time, launches, queries)
#include "cuda/cuda_backend_constants.cuh"
#include "cuda/neural/neural_bvh_module_constants.cuh"
#include "cuda/neural/neural_bvh_module_kernels.cuh"
#include "cuda/neural/neural_bvh_renderer_constants.cuh"
#include "cuda/neural/neural_bvh_renderer_kernels.cuh"
#include "cuda/neural/path/path_constants.cuh"
#include "cuda/neural/path/path_inference_kernels.cuh"
#include "cuda/neural/neural_bvh_renderer.h"

#include <tiny-cuda-nn/common.h>
#include <tiny-cuda-nn/gpu_matrix.h>
#include <tiny-cuda-nn/network_with_input_encoding.h>

#include <algorithm>
#include <cstdio>
#include <fstream>
#include <string>
#include <vector>

using namespace ntwr;
using namespace ntwr::neural;
using namespace ntwr::neural::path;
using precision_t = tcnn::network_precision_t;

#define CK(x)                                                                                        \
    do {                                                                                             \
        cudaError_t e_ = (x);                                                                        \
        if (e_ != cudaSuccess) {                                                                     \
            fprintf(stderr, "%s:%d CUDA error %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));    \
            exit(3);                                                                                 \
        }                                                                                            \
    } while (0)

template <typename T>
static bool slurp(const std::string &p, std::vector<T> &v)
{
    std::ifstream f(p, std::ios::binary | std::ios::ate);
    if (!f)
        return false;
    const size_t bytes = (size_t)f.tellg();
    v.resize(bytes / sizeof(T));
    f.seekg(0);
    f.read(reinterpret_cast<char *>(v.data()), bytes);
    return true;
}
template <typename T>
static void dump(const std::string &p, const std::vector<T> &v)
{
    std::ofstream f(p, std::ios::binary);
    f.write(reinterpret_cast<const char *>(v.data()), sizeof(T) * v.size());
}
template <typename T>
static T *upload(const std::vector<T> &v)
{
    T *d = nullptr;
    CK(cudaMalloc(&d, std::max<size_t>(sizeof(T) * v.size(), 16)));
    CK(cudaMemcpy(d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
    return d;
}
template <typename T>
static T *dalloc(size_t n)
{
    T *d = nullptr;
    CK(cudaMalloc(&d, std::max<size_t>(sizeof(T) * n, 16)));
    CK(cudaMemset(d, 0, std::max<size_t>(sizeof(T) * n, 16)));
    return d;
}
template <typename T>
static std::vector<T> download(const T *d, size_t n)
{
    std::vector<T> v(n);
    CK(cudaMemcpy(v.data(), d, sizeof(T) * n, cudaMemcpyDeviceToHost));
    return v;
}
template <typename POD>
static POD read_pod(const std::string &path)
{
    std::vector<uint8_t> b;
    if (!slurp(path, b) || b.size() != sizeof(POD)) {
        fprintf(stderr, "%s: %zu bytes, the reference struct has %zu\n", path.c_str(), b.size(), sizeof(POD));
        exit(4);
    }
    POD p;
    memcpy((void *)&p, b.data(), sizeof(POD));
    return p;
}

constexpr int MAX_IV = 32;

// Same launch shape as tcnn::linear_kernel (128 threads, ceil(n/128) blocks; common_host.h:331-337), written as a direct
// launch because two of the reference kernels share their name with a device-side template.
#define LINEAR_KERNEL(kernel, stream, n, ...)                                                                  \
    do {                                                                                                       \
        if ((n) > 0)                                                                                           \
            kernel<<<tcnn::n_blocks_linear((uint32_t)(n)), tcnn::N_THREADS_LINEAR, 0, stream>>>(n, __VA_ARGS__); \
    } while (0)

// ---- parity hooks: the reference's device functions called one ray per thread -------------------------------
__global__ void hk_raygen(uint32_t n, int sample_idx, int offset, float *rays)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    glm::uvec2 px = glm::uvec2(i % fb_size.x, i / fb_size.x);
    LCGRand rng   = get_lcg_rng(sample_idx, offset, i);
    float xf      = (float(px.x) + lcg_randomf(rng)) / float(fb_size.x);
    float yf      = (float(px.y) + lcg_randomf(rng)) / float(fb_size.y);
    Ray3f ray     = camera_raygen(xf, yf);
    rays[6 * i + 0] = ray.o.x, rays[6 * i + 1] = ray.o.y, rays[6 * i + 2] = ray.o.z;
    rays[6 * i + 3] = ray.d.x, rays[6 * i + 4] = ray.d.y, rays[6 * i + 5] = ray.d.z;
}

__global__ void hk_intervals(uint32_t n, const float *rays, NeuralBVH nbvh, uint32_t *cnt, uint32_t *node, float *mint, float *maxt)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    Ray3f ray(glm::vec3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), glm::vec3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]));
    uint32_t stack[NEURAL_STACK_SIZE];
    stack[0]       = 0;
    int stack_size = 1;
    uint32_t c     = 0;
    while (stack_size > 0) {
        float lm[1], lM[1];
        uint32_t nid[1];
        bool got = nbvh.intersect(bvh_const.rendered_lod, ray, 0u, stack, stack_size, true, lm, lM, true, FLT_MAX, nid);
        if (!got)
            break;
        if (c < MAX_IV) {
            node[(size_t)i * MAX_IV + c] = nid[0];
            mint[(size_t)i * MAX_IV + c] = lm[0];
            maxt[(size_t)i * MAX_IV + c] = lM[0];
        }
        c++;
    }
    cnt[i] = c;
}

struct BlasHit {
    float t, u, v;
    uint32_t prim, mesh;
};
__global__ void hk_blas(uint32_t n, const float *rays, CudaSceneData scene, uint32_t skip_mesh, BlasHit *out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    Ray3f ray(glm::vec3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), glm::vec3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]));
    BlasHit best{FLT_MAX, 0.f, 0.f, 0xFFFFFFFFu, 0xFFFFFFFFu};
    for (uint32_t m = 0; m < scene.tlas.num_meshes; m++) {
        if (m == skip_mesh)
            continue;
        Ray3f r = ray;
        r.maxt  = best.t;
        PreliminaryItsData pi = scene.tlas.mesh_bvhs[m].intersect<false>(r);
        if (pi.is_valid() && pi.t < best.t)
            best = BlasHit{pi.t, pi.u, pi.v, pi.prim_idx, m};
    }
    out[i] = best;
}

struct HitRec {
    uint32_t valid_hit;
    float t;
    float a[3], b[3];
};
__global__ void hk_collect_final(uint32_t n, HybridIntersectionResult r, HitRec *out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    // records are cumulative per ray (best hit so far); launches are stream-ordered, so the last record written for a
    // pixel is its final primary hit.  Only camera rays (depth 1) are recorded.
    if (i >= n || r.ray_depth[i] != 1)
        return;
    HitRec h;
    h.valid_hit = r.valid_hit[i];
    h.t         = r.t[i];
    glm::vec3 a = r.normal_or_uv[i], b = r.albedo_or_mesh_prim_id[i];
    h.a[0] = a.x, h.a[1] = a.y, h.a[2] = a.z, h.b[0] = b.x, h.b[1] = b.y, h.b[2] = b.z;
    if (h.valid_hit == 0)
        h.a[0] = h.a[1] = h.a[2] = h.b[0] = h.b[1] = h.b[2] = 0.f;
    out[r.pixel_coord[i]] = h;
}
__global__ void hk_step0_pixels(uint32_t n, NeuralTraversalData td, uint32_t *px)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        px[i] = td.pixel_coord(i);
}

// ---- buffers the module would own (path_module_init.cu:83-198, neural_bvh_module.cu:22-86) ---------------------
struct Buffers {
    NeuralWavefrontData<NeuralTraversalResult> wf[2];
    NormalTraversalData normal;
    HybridIntersectionResult hybrid;
    float *net_in, *net_out;
    uint32_t *c_neural, *c_normal, *c_its;
};
static vec3_soa soa(size_t n)
{
    return vec3_soa(dalloc<float>(n), dalloc<float>(n), dalloc<float>(n));
}
static Buffers make_buffers(uint32_t n_pix, uint32_t in_dims, uint32_t out_dims)
{
    Buffers b;
    const size_t n  = tcnn::next_multiple(n_pix, 4u);
    const size_t ne = tcnn::next_multiple(n_pix, tcnn::batch_size_granularity);
    for (int k = 0; k < 2; k++) {
        vec3_soa o = soa(n), d = soa(n);
        b.wf[k].traversal_data = NeuralTraversalData(o, d, dalloc<uint8_t>(n), dalloc<uint8_t>(n), dalloc<float>(n), dalloc<float>(n),
                                                     dalloc<int>(n), dalloc<uint32_t>(n * NEURAL_STACK_SIZE), dalloc<uint32_t>(n),
                                                     dalloc<uint32_t>(n));
        b.wf[k].traversal_data.illumination    = dalloc<glm::vec3>(n);
        b.wf[k].traversal_data.throughput_rng  = dalloc<glm::vec4>(n);
        b.wf[k].traversal_data.tlas_stack_size = dalloc<int>(n);
        b.wf[k].traversal_data.tlas_stack      = dalloc<uint32_t>(n * TLAS_STACK_SIZE);
        b.wf[k].results = NeuralTraversalResult(dalloc<uint8_t>(n), dalloc<float>(n), dalloc<glm::vec3>(n), dalloc<glm::vec3>(n));
    }
    {
        vec3_soa o = soa(n), d = soa(n);
        b.normal = NormalTraversalData(o, d, dalloc<uint8_t>(n), dalloc<float>(n), dalloc<uint8_t>(n), dalloc<glm::vec3>(n),
                                       dalloc<glm::vec3>(n), dalloc<uint32_t>(n), dalloc<uint32_t>(n), dalloc<int>(n),
                                       dalloc<uint32_t>(n * TLAS_STACK_SIZE), dalloc<glm::vec3>(n), dalloc<glm::vec4>(n));
    }
    {
        vec3_soa o = soa(n), d = soa(n);
        b.hybrid = HybridIntersectionResult(dalloc<uint8_t>(n), dalloc<float>(n), dalloc<glm::vec3>(n), dalloc<glm::vec3>(n), o, d,
                                            dalloc<uint8_t>(n), dalloc<glm::vec3>(n), dalloc<glm::vec4>(n), dalloc<uint32_t>(n),
                                            dalloc<int>(n), dalloc<uint32_t>(n * TLAS_STACK_SIZE));
    }
    b.net_in   = dalloc<float>(ne * in_dims);
    b.net_out  = dalloc<float>(ne * out_dims);
    b.c_neural = dalloc<uint32_t>(1);
    b.c_normal = dalloc<uint32_t>(1);
    b.c_its    = dalloc<uint32_t>(1);
    return b;
}

struct FrameStats {
    uint64_t queries = 0, launches = 0, neural_steps = 0, tlas_steps = 0, host_syncs = 0;
};

// One frame == NeuralPathModule::neural_path_tracing (path_module.cu:139-245) with neural_bvh_traversal (:54-137) inlined.
static FrameStats reference_frame(Buffers &b, tcnn::NetworkWithInputEncoding<precision_t> &net, const NeuralBVH &nbvh, CudaSceneData scene,
                                  uint32_t n_pix, int spp, int offset, uint32_t max_depth, uint32_t in_dims, uint32_t out_dims,
                                  HitRec *final_hits, std::vector<float> *step0_in, std::vector<float> *step0_out,
                                  std::vector<uint32_t> *step0_px)
{
    const int max_traversal_steps = 1000, max_tlas_traversal = 1000;  // path_module.h defaults
    FrameStats st;
    cudaStream_t s0 = 0, s1;
    CK(cudaStreamCreate(&s1));
    uint32_t n_neural = 0, n_normal = 0;
    CK(cudaMemsetAsync(b.c_neural, 0, 4, s0));
    CK(cudaMemsetAsync(b.c_normal, 0, 4, s0));
    LINEAR_KERNEL(hybrid_camera_raygen_and_hybrid_first_hit_and_traversal, s0, n_pix, spp, offset, b.normal, nbvh,
                        b.wf[0].traversal_data, b.wf[0].results, scene, b.c_neural, b.c_normal);
    st.launches++;
    CK(cudaMemcpyAsync(&n_neural, b.c_neural, 4, cudaMemcpyDeviceToHost, s0));
    CK(cudaMemcpyAsync(&n_normal, b.c_normal, 4, cudaMemcpyDeviceToHost, s1));
    CK(cudaStreamSynchronize(s0));
    CK(cudaStreamSynchronize(s1));
    st.host_syncs++;
    int tlas_step = 0;
    bool first_neural = true;
    while (n_normal || n_neural) {
        CK(cudaMemsetAsync(b.c_its, 0, 4, s0));
        if (n_normal) {
            LINEAR_KERNEL(non_neural_bvh_intersection, s1, n_normal, spp, b.normal, b.hybrid, scene, b.c_its);
            st.launches++;
        }
        // neural_bvh_traversal
        {
            auto *cur = &b.wf[0];
            uint32_t n = n_neural;
            for (uint32_t step = 0; step <= (uint32_t)max_traversal_steps && n; step++) {
                const uint32_t n_pad = tcnn::next_multiple(n, tcnn::batch_size_granularity);
                tcnn::GPUMatrix<float> in(b.net_in, in_dims, n_pad), out(b.net_out, out_dims, n_pad);
                LINEAR_KERNEL(encode_bvh_network_input_data, s0, n, cur->traversal_data, in.data(), true);
                net.inference(s0, in, out);
                if (first_neural && step0_in) {
                    CK(cudaStreamSynchronize(s0));
                    *step0_in  = download(b.net_in, (size_t)n * in_dims);
                    *step0_out = download(b.net_out, (size_t)n * out_dims);
                    uint32_t *px = dalloc<uint32_t>(n);
                    hk_step0_pixels<<<(n + 255) / 256, 256, 0, s0>>>(n, cur->traversal_data, px);
                    CK(cudaStreamSynchronize(s0));
                    *step0_px = download(px, n);
                    cudaFree(px);
                }
                first_neural = false;
                LINEAR_KERNEL(process_bvh_network_output_and_intersect_neural_bvh, s0, n, step, out.data(), cur->traversal_data,
                                    cur->results, nbvh);
                auto *next = &b.wf[(step + 1) % 2];
                auto *prev = &b.wf[step % 2];
                CK(cudaMemsetAsync(b.c_neural, 0, 4, s0));
                LINEAR_KERNEL(compact_neural_data_and_accumulate, s0, n, spp, *prev, *next, b.hybrid, b.c_neural, b.c_its,
                                    step == (uint32_t)max_traversal_steps);
                st.queries += n;
                st.launches += 3 + 4;  // encode + process + compact + tcnn inference (3 grid kernels + fused MLP)
                st.neural_steps++;
                CK(cudaMemcpyAsync(&n, b.c_neural, 4, cudaMemcpyDeviceToHost, s0));
                CK(cudaStreamSynchronize(s0));
                st.host_syncs++;
                cur = next;
            }
        }
        CK(cudaDeviceSynchronize());
        uint32_t n_its = 0;
        CK(cudaMemcpyAsync(&n_its, b.c_its, 4, cudaMemcpyDeviceToHost));
        CK(cudaMemsetAsync(b.c_neural, 0, 4, s0));
        CK(cudaMemsetAsync(b.c_normal, 0, 4, s0));
        CK(cudaStreamSynchronize(s0));
        if (final_hits && n_its)
            hk_collect_final<<<(n_its + 255) / 256, 256, 0, s0>>>(n_its, b.hybrid, final_hits);
        const bool last = tlas_step == max_tlas_traversal - 1;
        LINEAR_KERNEL(hybrid_scatter_and_get_next_tlas_traversal_step, s0, n_its, spp, b.hybrid, b.wf[0].traversal_data,
                            b.wf[0].results, b.normal, nbvh, scene, b.c_neural, b.c_normal, last, max_depth);
        st.launches++;
        CK(cudaMemcpyAsync(&n_normal, b.c_normal, 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpyAsync(&n_neural, b.c_neural, 4, cudaMemcpyDeviceToHost));
        CK(cudaStreamSynchronize(s0));
        st.host_syncs += 2;
        st.tlas_steps++;
        if (last)
            break;
        tlas_step++;
    }
    CK(cudaStreamDestroy(s1));
    CK(cudaDeviceSynchronize());
    return st;
}

int main(int argc, char **argv)
{
    if (argc < 7) {
        fprintf(stderr, "usage: ref_render_harness <workdir> <W> <H> <max_depth> <log2_hashmap> <n_timed_frames>\n");
        return 2;
    }
    const std::string wd = argv[1];
    const uint32_t W = atoi(argv[2]), H = atoi(argv[3]), max_depth = atoi(argv[4]);
    const int log2_hashmap = atoi(argv[5]), n_timed = atoi(argv[6]);
    const uint32_t n_pix = W * H;

    // ---- constants (cudaMemcpyToSymbol like CudaBackend / NeuralBVHRenderer / the modules do every frame)
    auto c_renderer = read_pod<NeuralBVHRendererConstants>(wd + "/const_renderer.bin");
    auto c_bbox     = read_pod<Bbox3f>(wd + "/const_scene_bbox.bin");
    auto c_module   = read_pod<NeuralBVHModuleConstants>(wd + "/const_module.bin");
    auto c_pt       = read_pod<PathTracingOutputConstants>(wd + "/const_pt_output.bin");
    auto c_path     = read_pod<NeuralPathModuleConstants>(wd + "/const_path.bin");
    auto c_camera   = read_pod<CameraParams>(wd + "/const_camera.bin");
    auto c_scene    = read_pod<CudaSceneConstants>(wd + "/const_scene.bin");
    CK(cudaMemcpyToSymbol(bvh_const, &c_renderer, sizeof(c_renderer)));
    CK(cudaMemcpyToSymbol(neural_scene_bbox, &c_bbox, sizeof(c_bbox)));
    CK(cudaMemcpyToSymbol(bvh_module_const, &c_module, sizeof(c_module)));
    CK(cudaMemcpyToSymbol(pt_target_const, &c_pt, sizeof(c_pt)));
    CK(cudaMemcpyToSymbol(path_module_const, &c_path, sizeof(c_path)));
    CK(cudaMemcpyToSymbol(camera_params, &c_camera, sizeof(c_camera)));
    CK(cudaMemcpyToSymbol(scene_constants, &c_scene, sizeof(c_scene)));
    glm::uvec2 fb(W, H);
    CK(cudaMemcpyToSymbol(fb_size, &fb, sizeof(fb)));
    CK(cudaMemcpyToSymbol(fb_size_flat, &n_pix, sizeof(n_pix)));

    // ---- surfaces replacing the GL textures (cuda_backend.cpp:87-123)
    cudaArray_t arr_ldr, arr_acc;
    cudaChannelFormatDesc d8 = cudaCreateChannelDesc<uchar4>(), d32 = cudaCreateChannelDesc<float4>();
    CK(cudaMallocArray(&arr_ldr, &d8, W, H, cudaArraySurfaceLoadStore));
    CK(cudaMallocArray(&arr_acc, &d32, W, H, cudaArraySurfaceLoadStore));
    cudaResourceDesc rd = {};
    rd.resType          = cudaResourceTypeArray;
    cudaSurfaceObject_t s_ldr, s_acc;
    rd.res.array.array = arr_ldr;
    CK(cudaCreateSurfaceObject(&s_ldr, &rd));
    rd.res.array.array = arr_acc;
    CK(cudaCreateSurfaceObject(&s_acc, &rd));
    CK(cudaMemcpyToSymbol(framebuffer, &s_ldr, sizeof(s_ldr)));
    CK(cudaMemcpyToSymbol(accum_buffer, &s_acc, sizeof(s_acc)));

    // ---- scene + tree-cut
    std::vector<uint8_t> raw;
    std::vector<TLASNode> tlas_nodes;
    std::vector<uint32_t> tlas_idx;
    std::vector<NeuralBVH2Node> nb_nodes;
    if (!slurp(wd + "/tlas_nodes.bin", tlas_nodes) || !slurp(wd + "/tlas_indices.u32", tlas_idx) || !slurp(wd + "/nbvh_nodes.bin", nb_nodes)) {
        fprintf(stderr, "scene arrays missing in %s\n", wd.c_str());
        return 4;
    }
    std::vector<CudaBLAS> meshes;
    for (int m = 0;; m++) {
        std::vector<BVH2Node> nodes;
        std::vector<size_t> prim;
        std::vector<float> verts;
        std::vector<uint32_t> idx;
        const std::string p = wd + "/mesh" + std::to_string(m);
        if (!slurp(p + "_nodes.bin", nodes))
            break;
        slurp(p + "_prim_ids.u64", prim), slurp(p + "_vertices.f32", verts), slurp(p + "_indices.u32", idx);
        CudaBLAS bl;
        memset((void *)&bl, 0, sizeof(bl));
        std::vector<float> nrm, tc;
        std::vector<int> mids;
        std::vector<uint32_t> mmap;
        if (slurp(p + "_normals.f32", nrm) && !nrm.empty())
            bl.normals = (glm::vec3 *)upload(nrm);
        if (slurp(p + "_texcoords.f32", tc) && !tc.empty())
            bl.texcoords = (glm::vec2 *)upload(tc);
        if (slurp(p + "_material_ids.i32", mids) && slurp(p + "_material_map.u32", mmap) && !mmap.empty()) {
            bl.material_ids  = upload(mids);
            bl.material_map  = upload(mmap);
            bl.num_materials = (uint32_t)mmap.size();
        }
        bl.num_nodes        = (uint32_t)nodes.size();
        bl.prim_indices     = upload(prim);
        bl.num_prim_indices = (uint32_t)prim.size();
        bl.vertices         = (glm::vec3 *)upload(verts);
        bl.indices          = (glm::uvec3 *)upload(idx);
        bl.nodes            = upload(nodes);
        meshes.push_back(bl);
    }
    CudaSceneData scene;
    memset((void *)&scene, 0, sizeof(scene));
    scene.tlas.tlas_nodes     = upload(tlas_nodes);
    scene.tlas.num_tlas_nodes = (uint32_t)tlas_nodes.size();
    scene.tlas.blas_indices   = upload(tlas_idx);
    scene.tlas.mesh_bvhs      = upload(meshes);
    scene.tlas.num_meshes     = (uint32_t)meshes.size();
    NeuralBVH nbvh;
    nbvh.nodes     = upload(nb_nodes);
    nbvh.num_nodes = (uint32_t)nb_nodes.size();
    printf("[ref_render_harness] %ux%u, %zu tree-cut nodes, %zu meshes, neural_bvh_idx %u\n", W, H, nb_nodes.size(), meshes.size(),
           c_renderer.neural_bvh_idx);

    // materials (raw PackedMaterial bytes; the vtable slot is never read on the device) and RGBA8 textures
    std::vector<uint8_t> mats;
    if (slurp(wd + "/materials.bin", mats) && !mats.empty()) {
        if (mats.size() % sizeof(PackedMaterial)) {
            fprintf(stderr, "materials.bin: %zu bytes is not a multiple of sizeof(PackedMaterial)=%zu\n", mats.size(), sizeof(PackedMaterial));
            return 4;
        }
        scene.materials = (PackedMaterial *)upload(mats);
    }
    std::vector<cudaTextureObject_t> texs;
    for (int k = 0;; k++) {
        std::vector<uint8_t> px;
        std::vector<uint32_t> dims;
        if (!slurp(wd + "/tex" + std::to_string(k) + ".rgba8", px) || !slurp(wd + "/tex" + std::to_string(k) + ".dims", dims))
            break;
        const uint32_t tw = dims[0], th = dims[1];
        cudaChannelFormatDesc cd = cudaCreateChannelDesc<uchar4>();
        cudaArray_t arr;
        CK(cudaMallocArray(&arr, &cd, tw, th));
        CK(cudaMemcpy2DToArray(arr, 0, 0, px.data(), tw * 4, tw * 4, th, cudaMemcpyHostToDevice));
        cudaResourceDesc trd = {};
        trd.resType          = cudaResourceTypeArray;
        trd.res.array.array  = arr;
        cudaTextureDesc td   = {}; // CudaBackend::create_cuda_textures_buffer (cuda_backend.cpp:421-432)
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeWrap;
        td.filterMode = cudaFilterModeLinear, td.readMode = cudaReadModeNormalizedFloat, td.normalizedCoords = 1;
        td.maxAnisotropy = 1, td.maxMipmapLevelClamp = 99, td.minMipmapLevelClamp = 0, td.mipmapFilterMode = cudaFilterModePoint;
        td.borderColor[0] = 1.0f, td.sRGB = 0;
        cudaTextureObject_t t;
        CK(cudaCreateTextureObject(&t, &trd, &td, nullptr));
        texs.push_back(t);
    }
    if (!texs.empty())
        scene.textures = upload(texs);

    // ---- environment map == CudaBackend::load_envmap_texture (cuda_backend.cpp:444-500); absent file -> the reference's white image
    {
        std::vector<float> env;
        std::vector<uint32_t> dims;
        int ew = 1, eh = 1;
        if (slurp(wd + "/envmap.f32", env) && slurp(wd + "/envmap.dims", dims) && dims.size() == 2 && env.size() == (size_t)dims[0] * dims[1] * 4) {
            ew = (int)dims[0], eh = (int)dims[1];
        } else {
            env.assign(4, 1.f);
        }
        cudaChannelFormatDesc cd = cudaCreateChannelDesc(32, 32, 32, 32, cudaChannelFormatKindFloat);
        cudaArray_t arr;
        CK(cudaMallocArray(&arr, &cd, ew, eh));
        CK(cudaMemcpy2DToArray(arr, 0, 0, env.data(), (size_t)ew * 16, (size_t)ew * 16, eh, cudaMemcpyHostToDevice));
        cudaResourceDesc erd = {};
        erd.resType          = cudaResourceTypeArray;
        erd.res.array.array  = arr;
        cudaTextureDesc etd  = {};
        etd.filterMode = cudaFilterModeLinear, etd.normalizedCoords = true;
        etd.addressMode[0] = etd.addressMode[1] = cudaAddressModeClamp;
        cudaTextureObject_t et;
        CK(cudaCreateTextureObject(&et, &erd, &etd, nullptr));
        CK(cudaMemcpyToSymbol(envmap_texture, &et, sizeof(et)));
        printf("[ref_render_harness] envmap %dx%d\n", ew, eh);
    }

    // ---- treecut mode: the reference's tree-cut learning kernels (K17) for one optimisation step
    //      usage: ... 0 treecut <batch> <num_splits>; reads tree_cut.bin, tc_L.f32 (batch x 16), tc_sampled.u32 (batch);
    //      runs update_neural_tree_cut_loss + (k + 1) x update_neural_tree_cut_decisions exactly as neural_bvh_renderer.cu:199-236;
    //      writes tc_cut_out.bin (NeuralTreeCutData per node), tc_max_losses.f32 (k), tc_split.u32 (k)
    if (argc >= 10 && std::string(argv[7]) == "treecut") {
        const uint32_t batch = atoi(argv[8]), k = atoi(argv[9]);
        std::vector<NeuralTreeCutData> cut;
        std::vector<float> L;
        std::vector<uint32_t> sampled;
        if (!slurp(wd + "/tree_cut.bin", cut) || cut.size() != nb_nodes.size() || !slurp(wd + "/tc_L.f32", L) || L.size() != (size_t)batch * 16 ||
            !slurp(wd + "/tc_sampled.u32", sampled) || sampled.size() != batch) {
            fprintf(stderr, "treecut mode: bad inputs in %s\n", wd.c_str());
            return 4;
        }
        NeuralTreeCutData *d_cut = upload(cut);
        float *d_L               = upload(L);
        uint32_t *d_sampled      = upload(sampled);
        float *d_max             = dalloc<float>(std::max(1u, k));
        uint32_t *d_split        = dalloc<uint32_t>(std::max(1u, k));
        LINEAR_KERNEL(update_neural_tree_cut_loss, 0, batch * 16u, 16u, c_pt.output_network_dims, d_L, d_sampled, d_cut, batch);
        CK(cudaDeviceSynchronize());
        for (uint32_t i = 0; i <= k; i++)
            LINEAR_KERNEL(update_neural_tree_cut_decisions, 0, nbvh.num_nodes, nbvh, d_cut, d_max, d_split, i, i == k);
        CK(cudaDeviceSynchronize());
        dump(wd + "/tc_cut_out.bin", download(d_cut, cut.size()));
        dump(wd + "/tc_max_losses.f32", download(d_max, k));
        dump(wd + "/tc_split.u32", download(d_split, k));
        printf("[ref_render_harness] treecut: %zu nodes, batch %u, %u splits requested\n", cut.size(), batch, k);
        return 0;
    }

    // ---- sampler mode: the reference's training-sample generator (K10) for a few steps
    if (argc >= 11 && std::string(argv[7]) == "sampler") {
        const uint32_t batch = atoi(argv[8]), log2_trav = atoi(argv[9]), steps = atoi(argv[10]);
        const float max_loss = argc >= 12 ? (float)atof(argv[11]) : 1.f;
        std::vector<NeuralTreeCutData> cut;
        if (!slurp(wd + "/tree_cut.bin", cut) || cut.size() != nb_nodes.size()) {
            fprintf(stderr, "tree_cut.bin: %zu entries, tree-cut has %zu nodes\n", cut.size(), nb_nodes.size());
            return 4;
        }
        NeuralTreeCutData *d_cut = upload(cut);
        const uint32_t in_d = c_module.input_network_dims, out_d = c_pt.output_network_dims;
        uint32_t *d_nodes = dalloc<uint32_t>(batch);
        float *d_in = dalloc<float>((size_t)batch * in_d), *d_tg = dalloc<float>((size_t)batch * out_d);
        const uint32_t sub = batch >> log2_trav, nsub = 1u << log2_trav;
        std::vector<float> all_in, all_tg;
        std::vector<uint32_t> all_nodes;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0), cudaEventCreate(&e1);
        float ms_sum = 0.f;
        for (uint32_t st = 0; st < steps; st++) {
            cudaEventRecord(e0, 0);
            LINEAR_KERNEL(granular_sample_bvh_network_input_data, 0, (int)sub, (int)nsub, (int)st, nbvh, scene, d_nodes, d_in, d_tg, max_loss, d_cut);
            cudaEventRecord(e1, 0);
            CK(cudaDeviceSynchronize());
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (st)
                ms_sum += ms;
            auto a = download(d_in, (size_t)batch * in_d);
            auto b2 = download(d_tg, (size_t)batch * out_d);
            auto c2 = download(d_nodes, batch);
            all_in.insert(all_in.end(), a.begin(), a.end());
            all_tg.insert(all_tg.end(), b2.begin(), b2.end());
            all_nodes.insert(all_nodes.end(), c2.begin(), c2.end());
        }
        dump(wd + "/sampler_in.f32", all_in);
        dump(wd + "/sampler_tgt.f32", all_tg);
        dump(wd + "/sampler_nodes.u32", all_nodes);
        FILE *fs = fopen((wd + "/sampler.json").c_str(), "w");
        fprintf(fs, "{\"batch\": %u, \"log2_bvh_max_traversals\": %u, \"steps\": %u, \"ms_per_step\": %.5f, \"n_textures\": %zu}\n", batch,
                log2_trav, steps, steps > 1 ? ms_sum / (steps - 1) : 0.f, texs.size());
        fclose(fs);
        printf("[ref_render_harness] sampler: batch %u, 2^%u traversals/thread, %u steps, %.4f ms/step\n", batch, log2_trav, steps,
               steps > 1 ? ms_sum / (steps - 1) : 0.f);
        return 0;
    }

    // ---- network exactly as NeuralBVHModule::reinitialise_neural_model builds it (neural_bvh_module.cu:118-263) for the shipped config
    using json = nlohmann::json;
    json point_encoding = {{"n_dims_to_encode", 3}, {"otype", "HashGrid"}, {"hash", "CoherentPrime"}, {"n_levels", 8},
                           {"n_features_per_level", 4}, {"log2_hashmap_size", log2_hashmap}, {"base_resolution", 8},
                           {"per_level_scale", 2.0f}, {"stochastic_interpolation", false}, {"interpolation", "Linear"}};
    json segment = {{"otype", "RepeatedComposite"}, {"reduction", "Concatenation"}, {"n_repetitions", 3},
                    {"n_dims_to_encode", 9}, {"nested", {point_encoding}}};
    json input_encoding = json::array();
    input_encoding.push_back(segment);
    json encoding_opts = {{"otype", "Composite"}, {"n_dims_to_encode", 9}, {"nested", input_encoding}};
    json network_opts  = {{"otype", "FullyFusedMLP"}, {"n_neurons", 64}, {"n_hidden_layers", 4}, {"activation", "ReLU"},
                          {"output_activation", "Sigmoid"}};
    tcnn::NetworkWithInputEncoding<precision_t> net(9u, 8u, encoding_opts, network_opts);
    std::vector<uint16_t> params;
    if (!slurp(wd + "/params.u16", params) || params.size() != net.n_params()) {
        fprintf(stderr, "params.u16: %zu entries, network needs %zu\n", params.size(), net.n_params());
        return 4;
    }
    precision_t *d_params = (precision_t *)upload(params);
    net.set_params(d_params, d_params, nullptr);

    const uint32_t in_dims = c_module.input_network_dims, out_dims = c_pt.output_network_dims;
    Buffers b = make_buffers(n_pix, in_dims, out_dims);

    // ---- parity hooks (skipped in path mode: image goldens and timings only)
    // `timing` (bench.py's live reference-GPU arm): the frame loop only -- no parity hooks, no dumps besides ref_frame.json
    const bool timing_only = argc >= 8 && std::string(argv[7]) == "timing";
    const bool path_mode  = timing_only || (argc >= 9 && std::string(argv[7]) == "path");
    const uint32_t blocks = (n_pix + 255) / 256;
    float *d_rays = dalloc<float>((size_t)n_pix * 6);
    if (!path_mode) {
    hk_raygen<<<blocks, 256>>>(n_pix, 0, 0, d_rays);
    CK(cudaDeviceSynchronize());
    dump(wd + "/rays.f32", download(d_rays, (size_t)n_pix * 6));
    {
        uint32_t *cnt = dalloc<uint32_t>(n_pix), *node = dalloc<uint32_t>((size_t)n_pix * MAX_IV);
        float *mint = dalloc<float>((size_t)n_pix * MAX_IV), *maxt = dalloc<float>((size_t)n_pix * MAX_IV);
        hk_intervals<<<blocks, 256>>>(n_pix, d_rays, nbvh, cnt, node, mint, maxt);
        CK(cudaDeviceSynchronize());
        dump(wd + "/iv_count.u32", download(cnt, n_pix));
        dump(wd + "/iv_node.u32", download(node, (size_t)n_pix * MAX_IV));
        dump(wd + "/iv_mint.f32", download(mint, (size_t)n_pix * MAX_IV));
        dump(wd + "/iv_maxt.f32", download(maxt, (size_t)n_pix * MAX_IV));
        cudaFree(cnt), cudaFree(node), cudaFree(mint), cudaFree(maxt);
    }
    {
        BlasHit *bh = dalloc<BlasHit>(n_pix);
        hk_blas<<<blocks, 256>>>(n_pix, d_rays, scene, c_renderer.ignore_neural_blas ? 0xFFFFFFFFu : c_renderer.neural_bvh_idx, bh);
        CK(cudaDeviceSynchronize());
        dump(wd + "/blas_hit.bin", download(bh, n_pix));
        cudaFree(bh);
    }
    }

    // ---- one recorded frame (sample 0), then timed frames
    HitRec *d_hits = dalloc<HitRec>(n_pix);
    std::vector<float> s0in, s0out;
    std::vector<uint32_t> s0px;
    FrameStats st = path_mode ? reference_frame(b, net, nbvh, scene, n_pix, 0, 0, max_depth, in_dims, out_dims, nullptr, nullptr, nullptr, nullptr)
                              : reference_frame(b, net, nbvh, scene, n_pix, 0, 0, max_depth, in_dims, out_dims, d_hits, &s0in, &s0out, &s0px);
    if (!path_mode) {
        dump(wd + "/hits.bin", download(d_hits, n_pix));
        dump(wd + "/step0_in.f32", s0in);
        dump(wd + "/step0_out.f32", s0out);
        dump(wd + "/step0_pixel.u32", s0px);
    }
    if (!timing_only) {
        std::vector<uchar4> ldr((size_t)n_pix);
        CK(cudaMemcpy2DFromArray(ldr.data(), W * 4, arr_ldr, 0, 0, W * 4, H, cudaMemcpyDeviceToHost));
        dump(wd + "/frame_ldr.u8", ldr);
    }
    printf("[ref_render_harness] frame: %llu queries, %llu neural steps, %llu tlas steps, %llu launches, %llu host syncs\n",
           (unsigned long long)st.queries, (unsigned long long)st.neural_steps, (unsigned long long)st.tlas_steps,
           (unsigned long long)st.launches, (unsigned long long)st.host_syncs);

    // ---- path mode: accumulate `spp` samples through the reference's frame loop and dump the RGBA32F accumulation buffer
    if (path_mode && !timing_only) {
        const int spp = atoi(argv[8]);
        std::vector<float4> acc((size_t)n_pix);
        CK(cudaMemcpy2DFromArray(acc.data(), W * 16, arr_acc, 0, 0, W * 16, H, cudaMemcpyDeviceToHost));
        dump(wd + "/path_accum_1spp.f32", acc);
        unsigned long long q = st.queries;
        for (int sidx = 1; sidx < spp; sidx++)
            q += reference_frame(b, net, nbvh, scene, n_pix, sidx, 0, max_depth, in_dims, out_dims, nullptr, nullptr, nullptr, nullptr).queries;
        CK(cudaMemcpy2DFromArray(acc.data(), W * 16, arr_acc, 0, 0, W * 16, H, cudaMemcpyDeviceToHost));
        dump(wd + "/path_accum.f32", acc);
        printf("[ref_render_harness] path: %d spp accumulated, %llu queries\n", spp, q);
    }

    float best_ms = 1e30f, sum_ms = 0.f;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    for (int f = 0; f < 3 + n_timed; f++) {
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0, 0);
        reference_frame(b, net, nbvh, scene, n_pix, 0, 0, max_depth, in_dims, out_dims, nullptr, nullptr, nullptr, nullptr);
        cudaEventRecord(e1, 0);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (f >= 3) {
            best_ms = std::min(best_ms, ms);
            sum_ms += ms;
        }
    }
    FILE *fj = fopen((wd + "/ref_frame.json").c_str(), "w");
    fprintf(fj,
            "{\"W\": %u, \"H\": %u, \"max_depth\": %u, \"rays\": %u, \"queries\": %llu, \"neural_steps\": %llu, \"tlas_steps\": %llu, "
            "\"launches\": %llu, \"host_syncs\": %llu, \"frames_timed\": %d, \"ms_best\": %.5f, \"ms_mean\": %.5f}\n",
            W, H, max_depth, n_pix, (unsigned long long)st.queries, (unsigned long long)st.neural_steps,
            (unsigned long long)st.tlas_steps, (unsigned long long)st.launches, (unsigned long long)st.host_syncs, n_timed, best_ms,
            n_timed ? sum_ms / n_timed : 0.f);
    fclose(fj);
    printf("[ref_render_harness] reference frame %ux%u max_depth %u: best %.3f ms, mean %.3f ms over %d frames\n", W, H, max_depth,
           best_ms, n_timed ? sum_ms / n_timed : 0.f, n_timed);
    return 0;
}

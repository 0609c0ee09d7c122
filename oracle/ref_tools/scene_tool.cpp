// TEST INFRASTRUCTURE (fixture generator) -- not part of the product path.
//
// Builds the device-side scene inputs of the hot path exactly the way the reference's host
// setup does, by compiling against the reference's OWN vendored headers where they lie under
// /root/reference (glm for the node transforms, ext/bvh for BLAS/TLAS construction).  Only the
// produced binaries (baseline/_ref/...) travel to the GPU box; no reference source is copied.
//
// Follows (behaviour, not text):
//   * glTF node transform composition          /root/reference/src/base/gltf/gltf_flatten.cpp:28-47
//   * flattening into one neural mesh + others /root/reference/src/base/scene.cpp:38-167
//   * BLAS build + 1e-4 bbox regularisation     /root/reference/src/cuda/cuda_backend.cpp:186-310
//   * TLAS build                                /root/reference/src/cuda/cuda_backend.cpp:312-385
//   * 32-byte node packing                      /root/reference/src/base/accel/bvh.cpp:133-151
//
// Input: a text manifest written by tools/make_fixtures.py (it parses the glTF JSON).
// Output: raw little-endian arrays + manifest.json in the output directory.
#include <bvh/v2/bvh.h>
#include <bvh/v2/default_builder.h>
#include <bvh/v2/executor.h>
#include <bvh/v2/node.h>
#include <bvh/v2/thread_pool.h>
#include <bvh/v2/tri.h>
#include <bvh/v2/vec.h>

#define GLM_ENABLE_EXPERIMENTAL
#include <glm/glm.hpp>
#include <glm/gtc/matrix_inverse.hpp>
#include <glm/gtc/quaternion.hpp>
#include <glm/gtc/type_ptr.hpp>
#include <glm/gtx/transform.hpp>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

using LibNode = bvh::v2::Node<float, 3>;
using LibBVH  = bvh::v2::Bvh<LibNode>;
using Vec3    = bvh::v2::Vec<float, 3>;
using BBox    = bvh::v2::BBox<float, 3>;
using Tri     = bvh::v2::Tri<float, 3>;

struct PackedNode {  // == BVH2Node / TLASNode byte layout (32 B)
    float mn[3];
    uint32_t a;
    float mx[3];
    uint32_t b;
};
static_assert(sizeof(PackedNode) == 32, "node must be 32 bytes");

struct OutMesh {
    std::string name;
    std::vector<glm::vec3> vertices, normals;
    std::vector<glm::vec2> texcoords;
    std::vector<glm::uvec3> indices;
    std::vector<int32_t> material_ids;
    std::vector<uint32_t> material_map;
    uint32_t index_offset = 0;
};

struct NodeTRS {
    bool has_t = false, has_r = false, has_s = false;
    double t[3], r[4], s[3];
};

static glm::mat4 node_matrix(const NodeTRS &n)
{
    glm::mat4 m(1.f);
    if (n.has_s)
        m = glm::scale(glm::vec3(n.s[0], n.s[1], n.s[2]));
    if (n.has_r) {
        const glm::quat q = glm::quat(n.r[3], n.r[0], n.r[1], n.r[2]);
        m                 = glm::mat4_cast(q) * m;
    }
    if (n.has_t) {
        m = glm::translate(glm::vec3(n.t[0], n.t[1], n.t[2])) * m;
    }
    return m;
}

template <typename T>
static void dump(const std::string &path, const std::vector<T> &v)
{
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char *>(v.data()), sizeof(T) * v.size());
}

static LibBVH build_blas(const OutMesh &mesh, bvh::v2::ThreadPool &pool)
{
    bvh::v2::ParallelExecutor executor(pool);
    const size_t n = mesh.indices.size();
    std::vector<BBox> bboxes(n);
    std::vector<Vec3> centers(n);
    executor.for_each(0, n, [&](size_t begin, size_t end) {
        for (size_t i = begin; i < end; ++i) {
            const glm::uvec3 id = mesh.indices[i];
            const glm::vec3 a = mesh.vertices[id.x], b = mesh.vertices[id.y], c = mesh.vertices[id.z];
            Tri tri(Vec3(a.x, a.y, a.z), Vec3(b.x, b.y, b.z), Vec3(c.x, c.y, c.z));
            bboxes[i]  = tri.get_bbox();
            centers[i] = tri.get_center();
        }
    });
    typename bvh::v2::DefaultBuilder<LibNode>::Config cfg;
    cfg.quality       = bvh::v2::DefaultBuilder<LibNode>::Quality::High;
    cfg.min_leaf_size = 1;
    cfg.max_leaf_size = 2;
    LibBVH bvh        = bvh::v2::DefaultBuilder<LibNode>::build(pool, bboxes, centers, cfg);
    const float eps   = 1e-4f;
    for (auto &nd : bvh.nodes) {
        auto diag = nd.get_bbox().get_diagonal();
        for (size_t i = 0; i < 3; i++) {
            if (diag[i] <= eps) {
                nd.bounds[2 * i] -= eps * 0.5f;
                nd.bounds[2 * i + 1] += eps * 0.5f;
            }
        }
    }
    return bvh;
}

static std::vector<PackedNode> pack(const LibBVH &bvh)
{
    std::vector<PackedNode> out(bvh.nodes.size());
    for (size_t i = 0; i < bvh.nodes.size(); i++) {
        auto bb = bvh.nodes[i].get_bbox();
        for (int k = 0; k < 3; k++) {
            out[i].mn[k] = bb.min[k];
            out[i].mx[k] = bb.max[k];
        }
        out[i].a = (uint32_t)bvh.nodes[i].index.prim_count;
        out[i].b = (uint32_t)bvh.nodes[i].index.first_id;
    }
    return out;
}

int main(int argc, char **argv)
{
    if (argc < 2) {
        fprintf(stderr, "usage: scene_tool manifest.txt\n");
        return 2;
    }
    std::ifstream mf(argv[1]);
    std::string line, outdir;
    std::vector<char> bin;
    std::vector<OutMesh> meshes;
    std::vector<NodeTRS> chain;
    int neural_mesh = 0;
    while (std::getline(mf, line)) {
        std::istringstream ss(line);
        std::string tok;
        ss >> tok;
        if (tok == "bin") {
            std::string p;
            ss >> p;
            std::ifstream bf(p, std::ios::binary);
            bin.assign(std::istreambuf_iterator<char>(bf), {});
        } else if (tok == "out") {
            ss >> outdir;
        } else if (tok == "mesh") {
            meshes.emplace_back();
            ss >> meshes.back().name;
        } else if (tok == "inst") {
            chain.clear();
        } else if (tok == "node") {
            NodeTRS n;
            std::string k;
            while (ss >> k) {
                if (k == "T") { n.has_t = true; ss >> n.t[0] >> n.t[1] >> n.t[2]; }
                if (k == "R") { n.has_r = true; ss >> n.r[0] >> n.r[1] >> n.r[2] >> n.r[3]; }
                if (k == "S") { n.has_s = true; ss >> n.s[0] >> n.s[1] >> n.s[2]; }
            }
            chain.push_back(n);
        } else if (tok == "prim") {
            long pos_off, pos_cnt, nrm_off, uv_off, idx_off, idx_cnt, idx_type, material;
            ss >> pos_off >> pos_cnt >> nrm_off >> uv_off >> idx_off >> idx_cnt >> idx_type >> material;
            OutMesh &m = meshes.back();
            glm::mat4 xf(1.f);
            for (auto &n : chain)
                xf = xf * node_matrix(n);
            const glm::mat3 xf_n = glm::inverseTranspose(glm::mat3(xf));
            m.material_map.push_back((uint32_t)m.indices.size());
            const float *pos = reinterpret_cast<const float *>(bin.data() + pos_off);
            for (long i = 0; i < pos_cnt; i++) {
                glm::vec3 v(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
                if (xf != glm::mat4(1.f))
                    v = xf * glm::vec4(v, 1.f);
                m.vertices.push_back(v);
            }
            if (uv_off >= 0) {
                const float *uv = reinterpret_cast<const float *>(bin.data() + uv_off);
                for (long i = 0; i < pos_cnt; i++)
                    m.texcoords.push_back(glm::vec2(uv[2 * i], uv[2 * i + 1]));
            } else if (m.index_offset != 0) {
                m.texcoords.resize(m.index_offset + pos_cnt);
            }
            if (nrm_off >= 0) {
                const float *nr = reinterpret_cast<const float *>(bin.data() + nrm_off);
                for (long i = 0; i < pos_cnt; i++) {
                    glm::vec3 v(nr[3 * i], nr[3 * i + 1], nr[3 * i + 2]);
                    if (xf != glm::mat4(1.f))
                        v = xf_n * v;
                    m.normals.push_back(v);
                }
            } else if (m.index_offset != 0) {
                m.normals.resize(m.index_offset + pos_cnt);
            }
            for (long i = 0; i < idx_cnt / 3; i++) {
                uint32_t id[3];
                for (int k = 0; k < 3; k++) {
                    if (idx_type == 2)
                        id[k] = reinterpret_cast<const uint16_t *>(bin.data() + idx_off)[3 * i + k];
                    else
                        id[k] = reinterpret_cast<const uint32_t *>(bin.data() + idx_off)[3 * i + k];
                }
                m.indices.push_back(glm::uvec3(m.index_offset + id[0], m.index_offset + id[1], m.index_offset + id[2]));
            }
            m.index_offset += (uint32_t)pos_cnt;
            m.material_ids.push_back((int32_t)material);
        } else if (tok == "neural") {
            ss >> neural_mesh;
        } else if (tok == "camera") {
            // camera <w> <h> <yfov_rad> <16 matrix floats>: golden for the camera-parameter derivation
            // (src/base/camera/camera.cpp:39-82, src/cuda/cuda_backend.cpp:595-611), computed by glm itself.
            unsigned w, h;
            float yfov, mtx[16];
            ss >> w >> h >> yfov;
            for (int i = 0; i < 16; i++)
                ss >> mtx[i];
            const glm::mat4 to_world = glm::make_mat4(mtx);
            const glm::mat4 inv      = glm::inverse(to_world);
            const float fovy_deg     = glm::degrees(yfov);
            const glm::vec3 origin   = glm::vec3{inv * glm::vec4{0, 0, 0, 1}};
            const glm::vec3 cam_dir  = glm::normalize(glm::vec3{inv * glm::vec4{0, 0, -1, 0}});
            const glm::vec3 cam_up   = glm::normalize(glm::vec3{inv * glm::vec4{0, 1, 0, 0}});
            glm::vec2 img;
            img.y                = 2.f * std::tan(glm::radians(0.5f * fovy_deg));
            const float aspect   = static_cast<float>(w) / h;
            img.x                = img.y * aspect;
            const glm::vec3 du   = glm::normalize(glm::cross(cam_dir, cam_up)) * img.x;
            const glm::vec3 dv   = -glm::normalize(glm::cross(du, cam_dir)) * img.y;
            const glm::vec3 tl   = cam_dir - 0.5f * du - 0.5f * dv;
            const float vals[12] = {origin.x, origin.y, origin.z, du.x, du.y, du.z, dv.x, dv.y, dv.z, tl.x, tl.y, tl.z};
            std::vector<float> v(vals, vals + 12);
            dump(outdir + "/camera_params.f32", v);
        }
    }

    bvh::v2::ThreadPool pool;
    std::vector<BBox> roots;
    std::ostringstream js;
    js << "{\n  \"neural_mesh\": " << neural_mesh << ",\n  \"meshes\": [\n";
    for (size_t mi = 0; mi < meshes.size(); mi++) {
        OutMesh &m = meshes[mi];
        fprintf(stderr, "[scene_tool] mesh %zu '%s': %zu verts, %zu tris -> building BLAS...\n", mi, m.name.c_str(),
                m.vertices.size(), m.indices.size());
        LibBVH bvh = build_blas(m, pool);
        roots.push_back(bvh.get_root().get_bbox());
        auto nodes = pack(bvh);
        std::vector<uint32_t> prim_ids(bvh.prim_ids.begin(), bvh.prim_ids.end());
        const std::string p = outdir + "/mesh" + std::to_string(mi);
        dump(p + "_vertices.f32", m.vertices);
        dump(p + "_indices.u32", m.indices);
        dump(p + "_normals.f32", m.normals);
        dump(p + "_texcoords.f32", m.texcoords);
        dump(p + "_material_ids.i32", m.material_ids);
        dump(p + "_material_map.u32", m.material_map);
        dump(p + "_bvh_nodes.bin", nodes);
        dump(p + "_prim_ids.u32", prim_ids);
        js << "    {\"name\": \"" << m.name << "\", \"n_vertices\": " << m.vertices.size()
           << ", \"n_triangles\": " << m.indices.size() << ", \"n_normals\": " << m.normals.size()
           << ", \"n_texcoords\": " << m.texcoords.size() << ", \"n_nodes\": " << nodes.size()
           << ", \"n_prim_ids\": " << prim_ids.size() << ", \"n_materials\": " << m.material_ids.size() << "}"
           << (mi + 1 < meshes.size() ? "," : "") << "\n";
    }
    js << "  ],\n";

    // TLAS over the BLAS root boxes
    {
        std::vector<BBox> bboxes(roots);
        std::vector<Vec3> centers(roots.size());
        for (size_t i = 0; i < roots.size(); i++)
            centers[i] = bboxes[i].get_center();
        typename bvh::v2::DefaultBuilder<LibNode>::Config cfg;
        cfg.quality       = bvh::v2::DefaultBuilder<LibNode>::Quality::Medium;
        cfg.min_leaf_size = 1;
        cfg.max_leaf_size = 1;
        LibBVH tlas       = bvh::v2::DefaultBuilder<LibNode>::build(pool, bboxes, centers, cfg);
        std::vector<PackedNode> nodes(tlas.nodes.size());
        for (size_t i = 0; i < tlas.nodes.size(); i++) {
            auto bb = tlas.nodes[i].get_bbox();
            for (int k = 0; k < 3; k++) {
                nodes[i].mn[k] = bb.min[k];
                nodes[i].mx[k] = bb.max[k];
            }
            nodes[i].a = (uint32_t)tlas.nodes[i].index.first_id;  // left_or_blas_offset
            uint32_t type = 0;
            if (tlas.nodes[i].is_leaf())
                type = ((int)tlas.prim_ids[tlas.nodes[i].index.first_id] == neural_mesh) ? 2u : 1u;
            nodes[i].b = type;  // leaf_type
        }
        std::vector<uint32_t> idx(tlas.prim_ids.begin(), tlas.prim_ids.end());
        dump(outdir + "/tlas_nodes.bin", nodes);
        dump(outdir + "/tlas_indices.u32", idx);
        js << "  \"n_tlas_nodes\": " << nodes.size() << ",\n  \"n_tlas_indices\": " << idx.size() << "\n}\n";
    }
    std::ofstream(outdir + "/manifest.json") << js.str();
    fprintf(stderr, "[scene_tool] done -> %s\n", outdir.c_str());
    return 0;
}

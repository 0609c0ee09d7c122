// io.cu -- on-disk formats of the reference handled by libnbvh_b200.so: .nbvh tree-cuts, network_models JSON, EXR frames,
// timing CSV and result paths.
#include "host_context.hpp"

extern "C" {

// .nbvh = bvh::v2::Bvh<Node<float,3,64,32>>::serialize: u64 n_nodes, u64 n_prim_ids, n_nodes x {6 f32 bounds
// (minx,maxx,miny,maxy,minz,maxz), u64 first_id, u64 prim_count}, n_prim_ids x u64
// (/root/reference/ext/bvh/src/bvh/v2/bvh.h:133-153, node.h:119-133)
int nbvh_load_neural_bvh_file(nbvh_ctx *ctx, const char *path)
{
    DeviceGuard device_guard(ctx);
    if (!ctx || !path)
        return NBVH_ERR_INVALID_ARG;
    std::string data;
    if (!read_file(path, data))
        return fail(ctx, NBVH_ERR_IO, std::string("cannot read ") + path);
    if (data.size() < 16)
        return fail(ctx, NBVH_ERR_PARSE, ".nbvh too short");
    uint64_t n_nodes, n_prims;
    memcpy(&n_nodes, data.data(), 8);
    memcpy(&n_prims, data.data() + 8, 8);
    if (n_nodes == 0 || n_nodes > 0x7FFFFFFFull || n_prims > data.size() / 8 || data.size() != 16 + n_nodes * 40 + n_prims * 8)
        return fail(ctx, NBVH_ERR_PARSE, ".nbvh size does not match its header");
    std::vector<nbvh_node32> nodes(n_nodes);
    const char *p = data.data() + 16;
    for (uint64_t i = 0; i < n_nodes; i++, p += 40) {
        float b[6];
        uint64_t first, count;
        memcpy(b, p, 24);
        memcpy(&first, p + 24, 8);
        memcpy(&count, p + 32, 8);
        nodes[i].aabb_min[0] = b[0], nodes[i].aabb_max[0] = b[1];
        nodes[i].aabb_min[1] = b[2], nodes[i].aabb_max[1] = b[3];
        nodes[i].aabb_min[2] = b[4], nodes[i].aabb_max[2] = b[5];
        nodes[i].prim_count_or_left  = (uint32_t)count;
        nodes[i].left_or_prim_offset = (uint32_t)first;
    }
    return adopt_nbvh(ctx, nodes.data(), (uint32_t)n_nodes, cudaMemcpyHostToDevice);
}

int nbvh_save_neural_bvh_file(nbvh_ctx *ctx, const char *path)
{
    DeviceGuard device_guard(ctx);
    if (!ctx || !path)
        return NBVH_ERR_INVALID_ARG;
    if (!ctx->d_nbvh)
        return fail(ctx, NBVH_ERR_STATE, "no neural BVH set");
    std::vector<nbvh_node32> nodes(ctx->nbvh_nodes);
    CU_TRY(ctx, cudaMemcpy(nodes.data(), ctx->d_nbvh, (size_t)ctx->nbvh_nodes * 32, cudaMemcpyDeviceToHost));
    std::ofstream f(path, std::ios::binary);
    if (!f)
        return fail(ctx, NBVH_ERR_IO, std::string("cannot write ") + path);
    const uint64_t n = nodes.size(), one = 1, zero = 0;
    f.write((const char *)&n, 8);
    f.write((const char *)&one, 8); // one dummy prim id (neural_bvh_renderer_accel.cpp:186-187)
    for (auto &nd : nodes) {
        const float b[6] = {nd.aabb_min[0], nd.aabb_max[0], nd.aabb_min[1], nd.aabb_max[1], nd.aabb_min[2], nd.aabb_max[2]};
        const uint64_t first = nd.left_or_prim_offset, count = nd.prim_count_or_left;
        f.write((const char *)b, 24);
        f.write((const char *)&first, 8);
        f.write((const char *)&count, 8);
    }
    f.write((const char *)&zero, 8);
    return f ? NBVH_OK : fail(ctx, NBVH_ERR_IO, "write failed");
}

// network_models/*_network.json: {"n_params":N,"params_binary":{"bytes":[...],"subtype":null},"params_type":"__half"}
// (ext/tiny-cuda-nn trainer.h:275-320, gpu_memory_json.h:36-60); "float" payloads are accepted too (trainer.h:292-295)
int nbvh_load_network_file(nbvh_ctx *ctx, const char *path)
{
    DeviceGuard device_guard(ctx);
    if (!ctx || !path)
        return NBVH_ERR_INVALID_ARG;
    std::string text;
    if (!read_file(path, text))
        return fail(ctx, NBVH_ERR_IO, std::string("cannot read ") + path);
    Json root;
    try {
        root = json_parse(text);
    } catch (const std::exception &e) {
        return fail(ctx, NBVH_ERR_PARSE, e.what());
    }
    const Json *bin = root.find("params_binary");
    const Json *bytes = bin ? bin->find("bytes") : nullptr;
    if (!bytes || (bytes->type != Json::NumArray && bytes->type != Json::Array))
        return fail(ctx, NBVH_ERR_PARSE, "params_binary.bytes missing");
    const uint64_t n         = (uint64_t)root.value("n_params", 0.0);
    const std::string type   = root.value("params_type", std::string("__half"));
    const size_t elem        = type == "float" ? 4 : 2;
    if (type != "float" && type != "__half")
        return fail(ctx, NBVH_ERR_PARSE, "unknown params_type " + type);
    if (bytes->size() != n * elem)
        return fail(ctx, NBVH_ERR_PARSE, "params_binary.bytes length does not match n_params");
    std::vector<uint8_t> raw(bytes->size());
    for (size_t i = 0; i < raw.size(); i++)
        raw[i] = (uint8_t)bytes->at_num(i);
    std::vector<uint16_t> half(n);
    if (elem == 2) {
        memcpy(half.data(), raw.data(), raw.size());
    } else {
        for (uint64_t i = 0; i < n; i++) {
            float f;
            memcpy(&f, raw.data() + 4 * i, 4);
            half[i] = f32_to_f16_host(f);
        }
    }
    return nbvh_load_params_host(ctx, half.data(), n);
}

int nbvh_save_network_file(nbvh_ctx *ctx, const char *path)
{
    DeviceGuard device_guard(ctx);
    if (!ctx || !path)
        return NBVH_ERR_INVALID_ARG;
    std::vector<uint16_t> half(ctx->n_params);
    CU_TRY(ctx, cudaMemcpy(half.data(), ctx->d_params, half.size() * 2, cudaMemcpyDeviceToHost));
    FILE *f = fopen(path, "wb");
    if (!f)
        return fail(ctx, NBVH_ERR_IO, std::string("cannot write ") + path);
    fprintf(f, "{\"n_params\":%llu,\"params_binary\":{\"bytes\":[", (unsigned long long)ctx->n_params);
    const uint8_t *b = reinterpret_cast<const uint8_t *>(half.data());
    std::string buf;
    buf.reserve(1 << 20);
    for (size_t i = 0; i < half.size() * 2; i++) {
        if (i)
            buf += ',';
        buf += std::to_string((unsigned)b[i]);
        if (buf.size() > (1 << 20) - 8) {
            fwrite(buf.data(), 1, buf.size(), f);
            buf.clear();
        }
    }
    fwrite(buf.data(), 1, buf.size(), f);
    fprintf(f, "],\"subtype\":null},\"params_type\":\"__half\"}");
    fclose(f);
    return NBVH_OK;
}

} // extern "C"

// == write_exr (include/base/bitmap.h:97-182) as the reference calls it from RenderApp::save_framebuffer (render_app.cpp:132-165):
// scanline OpenEXR 2.0, channels A, B, G, R as 32-bit float, no compression (one scanline per chunk), every value -- alpha too --
// divided by rgb_renormalisation (the accumulated sample count).  Byte-identical to the file the reference's tinyexr writes
// (tests/golden/ref_exr_7x5.exr, generated by oracle/ref_tools/exr_golden.cpp).
namespace {
void exr_attr(std::string &h, const char *name, const char *type, const void *data, int32_t size)
{
    h.append(name, strlen(name) + 1);
    h.append(type, strlen(type) + 1);
    h.append(reinterpret_cast<const char *>(&size), 4);
    h.append(reinterpret_cast<const char *>(data), (size_t)size);
}
} // namespace

int nbvh_write_exr_host(const char *path, uint32_t width, uint32_t height, const float *rgba_host, float rgb_renormalisation)
{
    if (!path || !rgba_host || width == 0 || height == 0)
        return NBVH_ERR_INVALID_ARG;
    std::string h;
    const unsigned char magic[8] = {0x76, 0x2f, 0x31, 0x01, 2, 0, 0, 0};
    h.append(reinterpret_cast<const char *>(magic), 8);
    std::string ch;
    for (const char name : {'A', 'B', 'G', 'R'}) {
        ch.push_back(name), ch.push_back('\0');
        const int32_t v[4] = {2 /* FLOAT */, 0 /* pLinear + 3 reserved bytes */, 1, 1};
        ch.append(reinterpret_cast<const char *>(v), 16);
    }
    ch.push_back('\0');
    exr_attr(h, "channels", "chlist", ch.data(), (int32_t)ch.size());
    const unsigned char zero = 0;
    exr_attr(h, "compression", "compression", &zero, 1);
    const int32_t win[4] = {0, 0, (int32_t)width - 1, (int32_t)height - 1};
    exr_attr(h, "dataWindow", "box2i", win, 16);
    exr_attr(h, "displayWindow", "box2i", win, 16);
    exr_attr(h, "lineOrder", "lineOrder", &zero, 1);
    const float one = 1.f, centre[2] = {0.f, 0.f};
    exr_attr(h, "pixelAspectRatio", "float", &one, 4);
    exr_attr(h, "screenWindowCenter", "v2f", centre, 8);
    exr_attr(h, "screenWindowWidth", "float", &one, 4);
    h.push_back('\0');
    FILE *f = fopen(path, "wb");
    if (!f)
        return NBVH_ERR_IO;
    const size_t row_bytes = (size_t)width * 16, chunk = 8 + row_bytes;
    std::vector<uint64_t> offsets(height);
    for (uint32_t y = 0; y < height; y++)
        offsets[y] = h.size() + (size_t)height * 8 + (size_t)y * chunk;
    bool ok = fwrite(h.data(), 1, h.size(), f) == h.size() && fwrite(offsets.data(), 8, height, f) == height;
    std::vector<float> row((size_t)width * 4);
    for (uint32_t y = 0; y < height && ok; y++) {
        const float *src = rgba_host + (size_t)y * width * 4;
        for (int c = 0; c < 4; c++) // file channel c = A, B, G, R = source component 3 - c
            for (uint32_t x = 0; x < width; x++)
                row[(size_t)c * width + x] = src[4 * x + (3 - c)] / rgb_renormalisation;
        const int32_t hdr[2] = {(int32_t)y, (int32_t)row_bytes};
        ok = fwrite(hdr, 4, 2, f) == 2 && fwrite(row.data(), 4, row.size(), f) == row.size();
    }
    ok = (fclose(f) == 0) && ok;
    return ok ? NBVH_OK : NBVH_ERR_IO;
}

// == NeuralBVHRenderer::save_image (neural_bvh_renderer_init.cu:361-383): reads the accumulation buffer back and writes it divided by
// accumulated_spp; refuses while ema_accumulate is on, like the reference ("would just save incorrect images").
int nbvh_save_image_exr(nbvh_ctx *ctx, const char *path, uint32_t accumulated_spp, void *stream)
{
    DeviceGuard device_guard(ctx);
    if (!ctx || !path || accumulated_spp == 0)
        return NBVH_ERR_INVALID_ARG;
    if (!ctx->accum)
        return fail(ctx, NBVH_ERR_STATE, "framebuffer not set");
    if (ctx->scene_const.ema_accumulate)
        return fail(ctx, NBVH_ERR_STATE, "saving images while ema_accumulate is on is not supported");
    std::vector<float> host((size_t)ctx->fb_w * ctx->fb_h * 4);
    cudaStream_t st = (cudaStream_t)stream;
    CU_TRY(ctx, cudaMemcpyAsync(host.data(), ctx->accum, host.size() * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    const int rc = nbvh_write_exr_host(path, ctx->fb_w, ctx->fb_h, host.data(), (float)accumulated_spp);
    return rc == NBVH_OK ? rc : fail(ctx, rc, std::string("cannot write ") + path);
}

// == FrameTimer::write_to_file (include/utils/frame_timer.h:33-58): "frame index,frame time (ms)[, loss]" and one line per frame with
// the default ostream formatting (6 significant digits).
int nbvh_write_frame_times_csv_host(const char *path, const double *frame_ms, const float *losses_or_null, uint32_t n)
{
    if (!path || (!frame_ms && n))
        return NBVH_ERR_INVALID_ARG;
    FILE *f = fopen(path, "w");
    if (!f)
        return NBVH_ERR_IO;
    fputs(losses_or_null ? "frame index,frame time (ms), loss\n" : "frame index,frame time (ms)\n", f);
    for (uint32_t i = 0; i < n; i++) {
        if (losses_or_null)
            fprintf(f, "%u,%g,%g\n", i, frame_ms[i], (double)losses_or_null[i]);
        else
            fprintf(f, "%u,%g\n", i, frame_ms[i]);
    }
    return fclose(f) == 0 ? NBVH_OK : NBVH_ERR_IO;
}

// Result file names of the reference's validation mode (neural_bvh_renderer_init.cu:328-372):
// kind 0: <dir>/results/<base>_<variant>_spp_<n>.exr, 1: ..._learning_spp_<n>.csv, 2: ..._inference_spp_<n>.csv
int nbvh_result_path(const char *scene_dir, const char *base_config_name, const char *variant_config_name, int kind, uint32_t accumulated_spp,
                     char *out, size_t out_capacity)
{
    if (!scene_dir || !base_config_name || !variant_config_name || !out || kind < 0 || kind > 2)
        return NBVH_ERR_INVALID_ARG;
    const char *tag[] = {"", "_learning", "_inference"};
    const std::string p = std::string(scene_dir) + "/results/" + base_config_name + "_" + variant_config_name + tag[kind] + "_spp_" +
                          std::to_string(accumulated_spp) + (kind == 0 ? ".exr" : ".csv");
    if (p.size() + 1 > out_capacity)
        return NBVH_ERR_INVALID_ARG;
    memcpy(out, p.c_str(), p.size() + 1);
    return NBVH_OK;
}


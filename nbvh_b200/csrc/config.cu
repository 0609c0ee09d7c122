// config.cu -- the reference's configuration JSON: defaults of its constant structs, nbvh_load_config_json, nbvh_save_config_json
// (/root/reference/src/cuda/neural/neural_bvh_renderer_init.cu:566-749, include/cuda/neural/neural_bvh_renderer_constants.cuh:13-47,
//  include/cuda/neural/neural_bvh_module.h:79-94,243-274).
#include "host_context.hpp"

namespace nbvh_host {

// ---- defaults of the reference structs ---------------------------------------------------------
nbvh_renderer_constants default_renderer_constants()
{
    nbvh_renderer_constants c;
    memset(&c, 0, sizeof(c));
    c.neural_bvh_idx = 0, c.num_lods_learned = 1, c.current_learned_lod = 0, c.rendered_lod = 0;
    c.ignore_neural_blas = false, c.path_lod_switch = false, c.lod_switch = 0;
    c.leaf_boundary_sampling_ratio = 0.f;
    c.inference_mint_epsilon = 1e-4f, c.inference_maxt_epsilon = 1e-3f;
    c.learning_mint_epsilon = 1e-3f, c.learning_maxt_epsilon = 1e-3f;
    c.sampling_bbox_scaling = 1.5f, c.fixed_node_dilation = 1e-5f, c.node_extent_dilation = 1e-4f;
    c.loss_exposure = 2e-7f, c.show_loss_heuristic = false, c.ellipsoidal_bvh = false, c.rescale_ellipsoid = false;
    c.log2_bvh_max_traversals = 0, c.sample_cosine_weighted = true, c.backface_bounce_training = false;
    c.max_training_bounces = 1, c.training_node_survival = true, c.survival_probability_scale = 1.f;
    c.min_rr_survival_prob = 0.005f, c.view_dependent_sampling = false, c.view_dependent_ratio = 0.f;
    return c;
}

nbvh_module_constants default_module_constants()
{
    nbvh_module_constants c;
    memset(&c, 0, sizeof(c));
    c.input_segment_point_dim = 3, c.input_segment_points = 3, c.input_segment_points_dim = 9;
    c.input_segment_direction_dim = 0, c.input_network_dims = 9;
    c.segment_distribution = 0; // EquidistantInterior
    return c;
}

nbvh_pt_output_constants default_pt_constants()
{
    nbvh_pt_output_constants c;
    memset(&c, 0, sizeof(c));
    c.output_visibility_dim = 1, c.output_depth_dim = 1, c.output_normal_dim = 0, c.output_albedo_dim = 0;
    c.output_network_dims = 2;
    c.visibility_loss = 5, c.visibility_loss_scale = 1.f, c.depth_loss = 3, c.depth_loss_scale = 1.f;
    c.normal_loss = 3, c.normal_loss_scale = 1.f, c.albedo_loss = 0, c.albedo_loss_scale = 1.f;
    c.depth_learning_mode = 0;
    c.output_normal_linear_activation = true, c.output_depth_linear_activation = false;
    c.output_albedo_linear_activation = false, c.visibility_masking_loss = true;
    c.primary_visibility_threshold = 0.5f, c.secondary_visibility_threshold = 0.5f;
    c.depth_epsilon = 0.2f, c.fixed_depth_epsilon = 1e-5f, c.interpolate_depth = false, c.depth_interpolation_eps = 1e-3f;
    return c;
}

nbvh_path_constants default_path_constants()
{
    nbvh_path_constants c;
    memset(&c, 0, sizeof(c));
    c.output_mode = 0, c.normal_epsilon_offset = 1e-4f, c.apply_russian_roulette = true, c.rr_start_depth = 2;
    c.neural_roughness = 1.f, c.neural_specular = 0.f, c.neural_metallic = 0.f;
    return c;
}

nbvh_scene_constants default_scene_constants()
{
    nbvh_scene_constants c;
    memset(&c, 0, sizeof(c));
    c.constant_roughness = 1.f, c.constant_specular = 0.f, c.constant_metallic = 0.f;
    c.envmap_strength = 1.f, c.envmap_rotation = 0.f, c.envmap_eval = true, c.emission_eval = true;
    c.skylike_background = false, c.sample_light = false, c.ema_weight = 0.05f, c.ema_accumulate = true;
    c.force_constant_material = false;
    return c;
}

// ---- camera (host float arithmetic, no contraction: compiled with -fmad=false, and x86 baseline has no FMA)
// glm::inverse cofactor expansion + Camera::origin/dir/up + CudaBackend::update_camera
// (/root/reference/src/base/camera/camera.cpp:39-82, /root/reference/src/cuda/cuda_backend.cpp:595-611)
void invert4(const float *a, float *out)
{
    auto M = [&](int c, int r) { return a[c * 4 + r]; };
    const float c00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3), c02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3);
    const float c03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3), c04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
    const float c06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3), c07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
    const float c08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2), c10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2);
    const float c11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2), c12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3);
    const float c14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3), c15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
    const float c16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2), c18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2);
    const float c19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2), c20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1);
    const float c22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1), c23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);
    const float f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
    const float f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
    const float v0[4] = {M(1, 0), M(0, 0), M(0, 0), M(0, 0)}, v1[4] = {M(1, 1), M(0, 1), M(0, 1), M(0, 1)};
    const float v2[4] = {M(1, 2), M(0, 2), M(0, 2), M(0, 2)}, v3[4] = {M(1, 3), M(0, 3), M(0, 3), M(0, 3)};
    float inv[16];
    for (int i = 0; i < 4; i++) {
        const float sa = (i & 1) ? -1.f : 1.f, sb = -sa;
        inv[0 + i]  = (v1[i] * f0[i] - v2[i] * f1[i] + v3[i] * f2[i]) * sa;
        inv[4 + i]  = (v0[i] * f0[i] - v2[i] * f3[i] + v3[i] * f4[i]) * sb;
        inv[8 + i]  = (v0[i] * f1[i] - v1[i] * f3[i] + v3[i] * f5[i]) * sa;
        inv[12 + i] = (v0[i] * f2[i] - v1[i] * f4[i] + v2[i] * f5[i]) * sb;
    }
    const float det = (M(0, 0) * inv[0] + M(0, 1) * inv[4]) + (M(0, 2) * inv[8] + M(0, 3) * inv[12]);
    const float ood = 1.f / det;
    for (int i = 0; i < 16; i++)
        out[i] = inv[i] * ood;
}

void normalize3(const float *v, float *o)
{
    const float dot = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const float inv = 1.f / sqrtf(dot);
    o[0] = v[0] * inv, o[1] = v[1] * inv, o[2] = v[2] * inv;
}

void cross3(const float *x, const float *y, float *o)
{
    o[0] = x[1] * y[2] - y[1] * x[2];
    o[1] = x[2] * y[0] - y[2] * x[0];
    o[2] = x[0] * y[1] - y[0] * x[1];
}

nbvh_camera_params camera_from_matrix(const float *m, float yfov_rad, uint32_t w, uint32_t h)
{
    float inv[16];
    invert4(m, inv);
    nbvh_camera_params cp;
    float dir_raw[3], up_raw[3], dir[3], up[3];
    for (int r = 0; r < 3; r++) {
        cp.origin[r] = (inv[0 + r] * 0.f + inv[4 + r] * 0.f) + (inv[8 + r] * 0.f + inv[12 + r] * 1.f);
        dir_raw[r]   = (inv[0 + r] * 0.f + inv[4 + r] * 0.f) + (inv[8 + r] * -1.f + inv[12 + r] * 0.f);
        up_raw[r]    = (inv[0 + r] * 0.f + inv[4 + r] * 1.f) + (inv[8 + r] * 0.f + inv[12 + r] * 0.f);
    }
    normalize3(dir_raw, dir);
    normalize3(up_raw, up);
    const float fovy_deg = yfov_rad * 57.295779513082320876798154814105f;        // glm::degrees
    const float half_rad = (0.5f * fovy_deg) * 0.01745329251994329576923690768489f; // glm::radians
    const float plane_y  = 2.f * tanf(half_rad);
    const float aspect   = (float)w / (float)h;
    const float plane_x  = plane_y * aspect;
    float c1[3], c1n[3], c2[3], c2n[3];
    cross3(dir, up, c1);
    normalize3(c1, c1n);
    for (int i = 0; i < 3; i++)
        cp.du[i] = c1n[i] * plane_x;
    cross3(cp.du, dir, c2);
    normalize3(c2, c2n);
    for (int i = 0; i < 3; i++)
        cp.dv[i] = -c2n[i] * plane_y;
    for (int i = 0; i < 3; i++)
        cp.dir_top_left[i] = dir[i] - 0.5f * cp.du[i] - 0.5f * cp.dv[i];
    return cp;
}

} // namespace nbvh_host

extern "C" {

// ---- configuration ---------------------------------------------------------------------------------
} // extern "C"

// The constant blocks and descriptions a config file decides: a file that is rejected half-way (unknown loss, unsupported network
// shape, malformed camera) must leave all of them -- and with them the allocated parameter layout -- as they were.
namespace {
struct ConfigSnapshot {
    nbvh_scene_constants scene_const;
    nbvh_renderer_constants bvh_const;
    nbvh_module_constants module_const;
    nbvh_pt_output_constants pt_const;
    nbvh_path_constants path_const;
    nbvh_network_desc net_desc;
    nbvh_optimizer_desc opt_desc;
    std::string dir_enc_type;
    int max_depth;
    uint32_t training_step;
    explicit ConfigSnapshot(const nbvh_ctx *c)
        : scene_const(c->scene_const), bvh_const(c->bvh_const), module_const(c->module_const), pt_const(c->pt_const),
          path_const(c->path_const), net_desc(c->net_desc), opt_desc(c->opt_desc), dir_enc_type(c->dir_enc_type), max_depth(c->max_depth),
          training_step(c->training_step)
    {
    }
    void restore(nbvh_ctx *c) const
    {
        c->scene_const = scene_const, c->bvh_const = bvh_const, c->module_const = module_const, c->pt_const = pt_const;
        c->path_const = path_const, c->net_desc = net_desc, c->opt_desc = opt_desc, c->dir_enc_type = dir_enc_type;
        c->max_depth = max_depth, c->training_step = training_step;
    }
};
int load_config_body(nbvh_ctx *ctx, const Json &root);
} // namespace

extern "C" int nbvh_load_config_json(nbvh_ctx *ctx, const char *json_text)
{
    DeviceGuard device_guard(ctx);
    if (!ctx || !json_text)
        return NBVH_ERR_INVALID_ARG;
    Json root;
    try {
        root = json_parse(json_text);
    } catch (const std::exception &e) {
        return fail(ctx, NBVH_ERR_PARSE, e.what());
    }
    const ConfigSnapshot snapshot(ctx);
    int rc;
    try {
        rc = load_config_body(ctx, root);
    } catch (const std::exception &e) {
        rc = fail(ctx, NBVH_ERR_PARSE, e.what());
    }
    if (rc != NBVH_OK)
        snapshot.restore(ctx);
    return rc;
}

namespace {
int load_config_body(nbvh_ctx *ctx, const Json &root)
{
    const Json &scene = root.object("cuda_scene");
    const Json &rend  = root.object("neural_bvh_renderer");
    const Json &mod   = rend.object("neural_module");

    // scene constants (cuda_backend.h:93-107)
    {
        const Json &j           = scene.object("scene_constants");
        nbvh_scene_constants d  = default_scene_constants(), &c = ctx->scene_const;
        c.force_constant_material = j.value("force_constant_material", d.force_constant_material);
        c.constant_roughness    = (float)j.value("constant_roughness", (double)d.constant_roughness);
        c.constant_specular     = (float)j.value("constant_specular", (double)d.constant_specular);
        c.constant_metallic     = (float)j.value("constant_metallic", (double)d.constant_metallic);
        c.envmap_rotation       = (float)j.value("envmap_rotation", (double)d.envmap_rotation);
        c.envmap_strength       = (float)j.value("envmap_strength", (double)d.envmap_strength);
        c.ema_accumulate        = j.value("ema_accumulate", d.ema_accumulate);
        c.ema_weight            = (float)j.value("ema_weight", (double)d.ema_weight);
        c.envmap_eval           = j.value("envmap_eval", d.envmap_eval);
        c.emission_eval         = j.value("emission_eval", d.emission_eval);
        c.skylike_background    = j.value("skylike_background", d.skylike_background);
    }
    // bvh constants (neural_bvh_renderer_constants.cuh:77-106)
    {
        const Json &j = rend.object("bvh_constants");
        nbvh_renderer_constants d = default_renderer_constants(), &c = ctx->bvh_const;
#define GETI(name) c.name = (int32_t)j.value(#name, (double)d.name)
#define GETF(name) c.name = (float)j.value(#name, (double)d.name)
#define GETB(name) c.name = j.value(#name, (bool)d.name)
        GETI(num_lods_learned);
        GETI(current_learned_lod);
        GETI(rendered_lod);
        GETB(ignore_neural_blas);
        GETF(leaf_boundary_sampling_ratio);
        GETF(inference_mint_epsilon);
        GETF(inference_maxt_epsilon);
        GETF(learning_mint_epsilon);
        GETF(learning_maxt_epsilon);
        GETF(sampling_bbox_scaling);
        GETF(fixed_node_dilation);
        GETF(node_extent_dilation);
        GETF(loss_exposure);
        GETB(show_loss_heuristic);
        GETB(ellipsoidal_bvh);
        GETB(rescale_ellipsoid);
        GETI(log2_bvh_max_traversals);
        GETB(sample_cosine_weighted);
        GETB(backface_bounce_training);
        GETI(max_training_bounces);
        GETB(training_node_survival);
        GETF(survival_probability_scale);
        GETF(min_rr_survival_prob);
#undef GETI
#undef GETF
#undef GETB
    }
    // module constants (neural_bvh_module.h:112-130)
    {
        const Json &j = mod.object("bvh_module_constants");
        nbvh_module_constants d = default_module_constants(), &c = ctx->module_const;
        c.input_segment_point_dim     = (uint32_t)j.value("input_segment_point_dim", (double)d.input_segment_point_dim);
        c.input_segment_points        = (uint32_t)j.value("input_segment_points", (double)d.input_segment_points);
        c.input_segment_points_dim    = (uint32_t)j.value("input_segment_points_dim", (double)d.input_segment_points_dim);
        c.input_segment_direction_dim = (uint32_t)j.value("input_segment_direction_dim", (double)d.input_segment_direction_dim);
        c.input_network_dims          = (uint32_t)j.value("input_network_dims", (double)d.input_network_dims);
        const int sd = find_index(j.value("segment_distribution", std::string(kSegmentDistributions[d.segment_distribution])),
                                  kSegmentDistributions);
        if (sd < 0)
            return fail(ctx, NBVH_ERR_PARSE, "unknown segment_distribution");
        c.segment_distribution = sd;
        c.antialias_direction  = j.value("antialias_direction", d.antialias_direction);
        c.merge_operator       = j.value("merge_operator", d.merge_operator);
        c.apply_encoding_trafo = j.value("apply_encoding_trafo", d.apply_encoding_trafo);
        // derived, whatever the file says: NeuralBVHModule::reinitialise_neural_model (neural_bvh_module.cu:126-131) -- the shipped
        // partial configs state only input_segment_points / input_segment_direction_dim
        c.input_segment_points_dim = c.input_segment_points * c.input_segment_point_dim;
        c.input_network_dims       = c.input_segment_points_dim + c.input_segment_direction_dim;
        if (c.apply_encoding_trafo)
            return fail(ctx, NBVH_ERR_UNSUPPORTED, "apply_encoding_trafo (asserted off in the reference) is not implemented");
    }
    // path-tracing output constants (neural_bvh_module.h:305-351)
    {
        const Json &j = mod.object("pt_output_constants");
        nbvh_pt_output_constants d = default_pt_constants(), &c = ctx->pt_const;
        c.output_visibility_dim = (uint32_t)j.value("output_visibility_dim", (double)d.output_visibility_dim);
        c.output_depth_dim      = (uint32_t)j.value("output_depth_dim", (double)d.output_depth_dim);
        c.output_normal_dim     = (uint32_t)j.value("output_normal_dim", (double)d.output_normal_dim);
        c.output_albedo_dim     = (uint32_t)j.value("output_albedo_dim", (double)d.output_albedo_dim);
        // derived: NeuralPathModule::update_neural_model_module (path_module_init.cu:259-265) overrides the file's value with the sum
        c.output_network_dims   = c.output_visibility_dim + c.output_depth_dim + c.output_normal_dim + c.output_albedo_dim;
        struct {
            const char *key;
            uint8_t *dst;
            uint8_t def;
        } losses[] = {{"visibility_loss", &c.visibility_loss, d.visibility_loss},
                      {"depth_loss", &c.depth_loss, d.depth_loss},
                      {"normal_loss", &c.normal_loss, d.normal_loss},
                      {"albedo_loss", &c.albedo_loss, d.albedo_loss}};
        for (auto &l : losses) {
            const int idx = find_index(j.value(l.key, std::string(kLossTypes[l.def])), kLossTypes);
            if (idx < 0)
                return fail(ctx, NBVH_ERR_PARSE, std::string("unknown loss type for ") + l.key);
            *l.dst = (uint8_t)idx;
        }
        c.visibility_loss_scale = (float)j.value("visibility_loss_scale", (double)d.visibility_loss_scale);
        c.depth_loss_scale      = (float)j.value("depth_loss_scale", (double)d.depth_loss_scale);
        c.normal_loss_scale     = (float)j.value("normal_loss_scale", (double)d.normal_loss_scale);
        c.albedo_loss_scale     = (float)j.value("albedo_loss_scale", (double)d.albedo_loss_scale);
        const int dm = find_index(j.value("depth_learning_mode", std::string(kDepthModes[d.depth_learning_mode])), kDepthModes);
        if (dm < 0)
            return fail(ctx, NBVH_ERR_PARSE, "unknown depth_learning_mode");
        c.depth_learning_mode             = (uint8_t)dm;
        c.output_normal_linear_activation = j.value("output_normal_linear_activation", d.output_normal_linear_activation);
        c.output_depth_linear_activation  = j.value("output_depth_linear_activation", d.output_depth_linear_activation);
        c.output_albedo_linear_activation = j.value("output_albedo_linear_activation", d.output_albedo_linear_activation);
        c.visibility_masking_loss         = j.value("visibility_masking_loss", d.visibility_masking_loss);
        c.primary_visibility_threshold    = (float)j.value("primary_visibility_threshold", (double)d.primary_visibility_threshold);
        c.secondary_visibility_threshold  = (float)j.value("secondary_visibility_threshold", (double)d.secondary_visibility_threshold);
        c.depth_epsilon                   = (float)j.value("depth_epsilon", (double)d.depth_epsilon);
        c.fixed_depth_epsilon             = (float)j.value("fixed_depth_epsilon", (double)d.fixed_depth_epsilon);
    }
    // path module constants (path_module.h:242-253)
    {
        const Json &j = mod.object("path_module_constants");
        nbvh_path_constants d = default_path_constants(), &c = ctx->path_const;
        const int om = find_index(j.value("output_mode", std::string(kOutputModes[d.output_mode])), kOutputModes);
        if (om < 0)
            return fail(ctx, NBVH_ERR_PARSE, "unknown output_mode");
        c.output_mode            = (uint8_t)om;
        c.normal_epsilon_offset  = (float)j.value("normal_epsilon_offset", (double)d.normal_epsilon_offset);
        c.apply_russian_roulette = j.value("apply_russian_roulette", d.apply_russian_roulette);
        c.rr_start_depth         = (int32_t)j.value("rr_start_depth", (double)d.rr_start_depth);
        c.neural_roughness       = (float)j.value("neural_roughness", (double)d.neural_roughness);
        c.neural_specular        = (float)j.value("neural_specular", (double)d.neural_specular);
        c.neural_metallic        = (float)j.value("neural_metallic", (double)d.neural_metallic);
    }
    ctx->max_depth = (int)mod.value("max_depth", (double)ctx->max_depth);
    // network + encoding + optimizer (tcnn_config.h)
    {
        const Json &nw = mod.object("network");
        const Json &hg = mod.object("segment_encoding").object("hashgrid_options");
        nbvh_network_desc &d = ctx->net_desc;
        const std::string net_type = nw.value("network", std::string("FullyFusedMLP"));
        if (net_type != "FullyFusedMLP" && net_type != "CutlassMLP")
            return fail(ctx, NBVH_ERR_PARSE, "unknown network type");
        if (nw.value("activation", std::string("ReLU")) != "ReLU" ||
            nw.value("output_activation", std::string("Sigmoid")) != "Sigmoid")
            return fail(ctx, NBVH_ERR_UNSUPPORTED, "only ReLU hidden / Sigmoid output activations are implemented");
        const std::string enc_type = mod.object("segment_encoding").value("type", std::string("HashGrid"));
        if (enc_type != "HashGrid")
            return fail(ctx, NBVH_ERR_UNSUPPORTED, "only the HashGrid segment encoding is implemented");
        if (mod.object("segment_encoding").value("reduction", std::string("Concatenation")) != "Concatenation")
            return fail(ctx, NBVH_ERR_UNSUPPORTED, "only the Concatenation reduction is implemented");
        if (hg.value("hash", std::string("CoherentPrime")) != "CoherentPrime" ||
            hg.value("interpolation", std::string("Linear")) != "Linear" || hg.value("stochastic_interpolation", false))
            return fail(ctx, NBVH_ERR_UNSUPPORTED, "only CoherentPrime hash with Linear interpolation is implemented");
        d.n_neurons            = (int32_t)nw.value("n_neurons", 64.0);
        d.n_hidden_layers      = (int32_t)nw.value("n_hidden_layers", 4.0);
        d.n_levels             = (int32_t)hg.value("n_levels", 8.0);
        d.n_features_per_level = (int32_t)hg.value("n_features_per_level", 4.0);
        d.log2_hashmap_size    = (int32_t)hg.value("log2_hashmap_size", 18.0);
        d.base_resolution      = (int32_t)hg.value("base_resolution", 8.0);
        d.per_level_scale      = (float)hg.value("per_level_scale", 2.0);
        const Json &op         = mod.object("optimizer");
        nbvh_optimizer_desc &o = ctx->opt_desc;
        if (op.value("type", std::string("Adam")) != "Adam")
            return fail(ctx, NBVH_ERR_UNSUPPORTED, "only the Adam optimizer is implemented");
        o.learning_rate  = (float)op.value("learning_rate", 1e-2);
        o.decay_start    = (int32_t)op.value("decay_start", 10000.0);
        o.decay_interval = (int32_t)op.value("decay_interval", 2000.0);
        o.decay_base     = (float)op.value("decay_base", (double)0.33f);
        o.beta1          = (float)op.value("beta1", (double)0.9f);
        o.beta2          = (float)op.value("beta2", (double)0.999f);
        o.epsilon        = (float)op.value("epsilon", (double)1e-8f);
        o.l2_reg         = (float)op.value("l2_reg", (double)1e-8f);
    }
    ctx->dir_enc_type  = mod.object("direction_encoding").value("type", std::string("Identity"));
    ctx->training_step = (uint32_t)rend.value("training_step", 0.0);
    // everything that can be refused is checked before the first side effect (framebuffer, camera, parameter layout)
    {
        const std::string err = layout_error(ctx);
        if (!err.empty())
            return fail(ctx, NBVH_ERR_UNSUPPORTED, err);
        if (const Json *m = scene.object("camera").find("matrix"))
            if (m->size() != 16)
                return fail(ctx, NBVH_ERR_PARSE, "camera.matrix must have 16 entries");
    }
    // framebuffer size + camera
    if (const Json *fb = scene.find("fb_size")) {
        if (fb->size() == 2) {
            int rc = nbvh_set_framebuffer(ctx, nullptr, nullptr, (uint32_t)fb->at_num(0), (uint32_t)fb->at_num(1));
            if (rc != NBVH_OK)
                return rc;
        }
    }
    const Json &cam = scene.object("camera");
    if (const Json *m = cam.find("matrix")) {
        if (m->size() != 16)
            return fail(ctx, NBVH_ERR_PARSE, "camera.matrix must have 16 entries");
        float mtx[16];
        for (int i = 0; i < 16; i++)
            mtx[i] = (float)m->at_num(i);
        const float yfov = (float)cam.value("yfov", (double)(45.0f * 0.01745329251994329576923690768489f));
        int rc           = nbvh_set_camera(ctx, mtx, yfov);
        if (rc != NBVH_OK)
            return rc;
    }
    ctx->dilated_with = -1.f, ctx->walk_version++;
    return rebuild_layout(ctx);
}
} // namespace

extern "C" {

// == the module-owned part of NeuralBVHRenderer::save_config (neural_bvh_renderer_init.cu:406-509): cuda_scene {fb_size, camera,
// scene_constants} and neural_bvh_renderer {neural_module {max_depth, pt_output_constants, path_module_constants, bvh_module_constants,
// network, optimizer, segment_encoding, direction_encoding}, training_step, bvh_constants} with the reference's key names, floats
// printed with 9 significant digits so that nbvh_load_config_json(save(x)) restores every constant block bit for bit.  The renderer
// adds the keys it owns (batch_size, tree_cut_learning, bvh_split_scheduler, ...).  Writes at most `capacity` bytes (NUL-terminated)
// and the needed size (including the NUL) into *needed.
int nbvh_save_config_json(const nbvh_ctx *ctx, char *out, size_t capacity, size_t *needed)
{
    DeviceGuard device_guard(ctx);
    if (!ctx || (!out && capacity))
        return NBVH_ERR_INVALID_ARG;
    std::string s;
    char buf[64];
    bool first = true;
    auto key = [&](const char *k) {
        if (!first)
            s += ", ";
        first = false;
        s += "\"", s += k, s += "\": ";
    };
    auto open  = [&](const char *k) { key(k), s += "{", first = true; };
    auto close = [&]() { s += "}", first = false; };
    auto F = [&](const char *k, float v) { key(k), snprintf(buf, sizeof(buf), "%.9g", (double)v), s += buf; };
    auto I = [&](const char *k, long long v) { key(k), s += std::to_string(v); };
    auto B = [&](const char *k, bool v) { key(k), s += v ? "true" : "false"; };
    auto S = [&](const char *k, const char *v) { key(k), s += "\"", s += v, s += "\""; };
    s += "{", first = true;
    open("cuda_scene");
    key("fb_size"), s += "[" + std::to_string(ctx->fb_w) + ", " + std::to_string(ctx->fb_h) + "]";
    if (ctx->cam_set) {
        open("camera");
        S("type", "perspective");
        F("yfov", ctx->cam_yfov);
        key("matrix"), s += "[";
        for (int i = 0; i < 16; i++)
            snprintf(buf, sizeof(buf), i ? ", %.9g" : "%.9g", (double)ctx->cam_matrix[i]), s += buf;
        s += "]";
        close();
    }
    {
        const nbvh_scene_constants &c = ctx->scene_const;
        open("scene_constants");
        B("force_constant_material", c.force_constant_material), F("constant_roughness", c.constant_roughness);
        F("constant_specular", c.constant_specular), F("constant_metallic", c.constant_metallic);
        F("envmap_rotation", c.envmap_rotation), F("envmap_strength", c.envmap_strength);
        B("ema_accumulate", c.ema_accumulate), F("ema_weight", c.ema_weight), B("envmap_eval", c.envmap_eval);
        B("emission_eval", c.emission_eval), B("skylike_background", c.skylike_background);
        close();
    }
    close();
    open("neural_bvh_renderer");
    S("bvh_module_type", "Path");
    open("neural_module");
    I("max_depth", ctx->max_depth);
    {
        const nbvh_pt_output_constants &c = ctx->pt_const;
        open("pt_output_constants");
        I("output_visibility_dim", c.output_visibility_dim), I("output_depth_dim", c.output_depth_dim);
        I("output_normal_dim", c.output_normal_dim), I("output_albedo_dim", c.output_albedo_dim);
        I("output_network_dims", c.output_network_dims);
        S("visibility_loss", kLossTypes[c.visibility_loss]), S("depth_loss", kLossTypes[c.depth_loss]);
        S("normal_loss", kLossTypes[c.normal_loss]), S("albedo_loss", kLossTypes[c.albedo_loss]);
        F("visibility_loss_scale", c.visibility_loss_scale), F("depth_loss_scale", c.depth_loss_scale);
        F("normal_loss_scale", c.normal_loss_scale), F("albedo_loss_scale", c.albedo_loss_scale);
        S("depth_learning_mode", kDepthModes[c.depth_learning_mode]);
        B("output_normal_linear_activation", c.output_normal_linear_activation);
        B("output_depth_linear_activation", c.output_depth_linear_activation);
        B("output_albedo_linear_activation", c.output_albedo_linear_activation);
        B("visibility_masking_loss", c.visibility_masking_loss);
        F("primary_visibility_threshold", c.primary_visibility_threshold);
        F("secondary_visibility_threshold", c.secondary_visibility_threshold);
        F("depth_epsilon", c.depth_epsilon), F("fixed_depth_epsilon", c.fixed_depth_epsilon);
        close();
    }
    {
        const nbvh_path_constants &c = ctx->path_const;
        open("path_module_constants");
        S("output_mode", kOutputModes[c.output_mode]), F("normal_epsilon_offset", c.normal_epsilon_offset);
        B("apply_russian_roulette", c.apply_russian_roulette), I("rr_start_depth", c.rr_start_depth);
        F("neural_roughness", c.neural_roughness), F("neural_specular", c.neural_specular), F("neural_metallic", c.neural_metallic);
        close();
    }
    {
        const nbvh_module_constants &c = ctx->module_const;
        open("bvh_module_constants");
        I("input_segment_point_dim", c.input_segment_point_dim), I("input_segment_points", c.input_segment_points);
        I("input_segment_points_dim", c.input_segment_points_dim), I("input_segment_direction_dim", c.input_segment_direction_dim);
        I("input_network_dims", c.input_network_dims), S("segment_distribution", kSegmentDistributions[c.segment_distribution]);
        B("antialias_direction", c.antialias_direction), B("merge_operator", c.merge_operator);
        B("apply_encoding_trafo", c.apply_encoding_trafo);
        close();
    }
    {
        const nbvh_network_desc &d = ctx->net_desc;
        open("network");
        S("network", "FullyFusedMLP"), I("n_neurons", d.n_neurons), I("n_hidden_layers", d.n_hidden_layers);
        S("activation", "ReLU"), S("output_activation", "Sigmoid");
        close();
        const nbvh_optimizer_desc &o = ctx->opt_desc;
        open("optimizer");
        S("type", "Adam"), F("learning_rate", o.learning_rate), I("decay_start", o.decay_start), I("decay_interval", o.decay_interval);
        F("decay_base", o.decay_base), F("beta1", o.beta1), F("beta2", o.beta2), F("epsilon", o.epsilon), F("l2_reg", o.l2_reg);
        F("relative_decay", 0.f), F("absolute_decay", 0.f), B("adabound", false);
        close();
        open("segment_encoding");
        S("type", "HashGrid");
        open("hashgrid_options");
        S("otype", "HashGrid"), I("n_levels", d.n_levels), I("n_features_per_level", d.n_features_per_level);
        I("log2_hashmap_size", d.log2_hashmap_size), I("base_resolution", d.base_resolution), F("per_level_scale", d.per_level_scale);
        S("hash", "CoherentPrime"), S("interpolation", "Linear"), B("stochastic_interpolation", false);
        close();
        S("reduction", "Concatenation");
        close();
        open("direction_encoding");
        S("type", ctx->dir_enc_type.c_str());
        close();
    }
    close(); // neural_module
    I("training_step", ctx->training_step);
    {
        const nbvh_renderer_constants &c = ctx->bvh_const;
        open("bvh_constants");
        I("num_lods_learned", c.num_lods_learned), I("current_learned_lod", c.current_learned_lod), I("rendered_lod", c.rendered_lod);
        B("ignore_neural_blas", c.ignore_neural_blas), F("leaf_boundary_sampling_ratio", c.leaf_boundary_sampling_ratio);
        F("inference_mint_epsilon", c.inference_mint_epsilon), F("inference_maxt_epsilon", c.inference_maxt_epsilon);
        F("learning_mint_epsilon", c.learning_mint_epsilon), F("learning_maxt_epsilon", c.learning_maxt_epsilon);
        F("sampling_bbox_scaling", c.sampling_bbox_scaling), F("fixed_node_dilation", c.fixed_node_dilation);
        F("node_extent_dilation", c.node_extent_dilation), F("loss_exposure", c.loss_exposure);
        B("show_loss_heuristic", c.show_loss_heuristic), B("ellipsoidal_bvh", c.ellipsoidal_bvh), B("rescale_ellipsoid", c.rescale_ellipsoid);
        I("log2_bvh_max_traversals", c.log2_bvh_max_traversals), B("sample_cosine_weighted", c.sample_cosine_weighted);
        B("backface_bounce_training", c.backface_bounce_training), I("max_training_bounces", c.max_training_bounces);
        B("training_node_survival", c.training_node_survival), F("survival_probability_scale", c.survival_probability_scale);
        F("min_rr_survival_prob", c.min_rr_survival_prob);
        close();
    }
    close(); // neural_bvh_renderer
    s += "}";
    if (needed)
        *needed = s.size() + 1;
    if (capacity) {
        const size_t n = s.size() < capacity - 1 ? s.size() : capacity - 1;
        memcpy(out, s.data(), n);
        out[n] = '\0';
    }
    return s.size() + 1 <= capacity || !capacity ? NBVH_OK : NBVH_ERR_INVALID_ARG;
}

} // extern "C"

// TEST INFRASTRUCTURE (golden generator) -- runs the reference's OWN network path on the GPU box:
// its tiny-cuda-nn fork (RepeatedComposite hash grid + FullyFusedMLP + Trainer + Adam/ExponentialDecay) and its
// PathTracingComposite loss, compiled from the sources where they lie under /root/reference (oracle/Makefile.tcnn).
// Nothing here is product code and no reference source is copied: this file only *calls* the reference's API the
// way /root/reference/src/cuda/neural/neural_bvh_module.cu:118-263 and neural_bvh_renderer.cu:172-175,241 do.
//
// usage: tcnn_harness <workdir> <log2_hashmap> <N_query> <batch> <steps>
//   reads  <workdir>/params_init.u16 (fp16 bits, optional), query_in.f32 (N x 9), train_in.f32 (steps x B x 9),
//          train_tgt.f32 (steps x B x 8)
//   writes params_init.u16 (when absent: tcnn's own initialisation), query_out.f32 (N x 8), train_loss.f32 (steps),
//          train_L0.f32 (B x 16, per-output loss of step 0), grads0.u16 (fp16 gradients of step 0, loss scale 128),
//          params_after.u16 / params_after.f32 (after `steps` optimiser steps), timing.json
#define NEURAL_BVH_MODULE_CONSTANT __constant__
#include "cuda/neural/neural_bvh_module_constants.cuh"
#include "cuda/neural/utils/pt_composite_loss.cuh"

#include <tiny-cuda-nn/common.h>
#include <tiny-cuda-nn/gpu_matrix.h>
#include <tiny-cuda-nn/loss.h>
#include <tiny-cuda-nn/network_with_input_encoding.h>
#include <tiny-cuda-nn/optimizer.h>
#include <tiny-cuda-nn/trainer.h>

#include <cstdio>
#include <fstream>
#include <string>
#include <vector>

using namespace tcnn;
using precision_t = network_precision_t;

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

int main(int argc, char **argv)
{
    if (argc < 6) {
        fprintf(stderr, "usage: tcnn_harness <workdir> <log2_hashmap> <N_query> <batch> <steps>\n");
        return 2;
    }
    const std::string wd = argv[1];
    const int log2_hashmap = atoi(argv[2]);
    const uint32_t n_query = (uint32_t)atoll(argv[3]), batch = (uint32_t)atoll(argv[4]), steps = (uint32_t)atoll(argv[5]);

    // constants of the shipped Chess config (PathTracingOutputConstants with the shipped dims)
    ntwr::neural::PathTracingOutputConstants pt;
    pt.output_visibility_dim = 1, pt.output_depth_dim = 1, pt.output_normal_dim = 3, pt.output_albedo_dim = 3;
    pt.output_network_dims = 8;
    CUDA_CHECK_THROW(cudaMemcpyToSymbol(ntwr::neural::pt_target_const, &pt, sizeof(pt)));

    register_loss<precision_t>("PathTracingComposite", [](const json &loss) {
        return new ntwr::neural::PathTracingCompositeLoss<precision_t>{
            loss.value("output_visibility_dim", 1u), loss.value("visibility_loss", "BinaryCrossEntropy"),
            loss.value("output_depth_dim", 1u),      loss.value("depth_loss", "L1"),
            loss.value("output_normal_dim", 3u),     loss.value("normal_loss", "L1"),
            loss.value("output_albedo_dim", 3u),     loss.value("albedo_loss", "RelativeL2")};
    });
    json loss_opts = {{"otype", "PathTracingComposite"},
                      {"output_visibility_dim", 1}, {"visibility_loss", ntwr::loss_types[pt.visibility_loss]},
                      {"output_depth_dim", 1},      {"depth_loss", ntwr::loss_types[pt.depth_loss]},
                      {"output_normal_dim", 3},     {"normal_loss", ntwr::loss_types[pt.normal_loss]},
                      {"output_albedo_dim", 3},     {"albedo_loss", ntwr::loss_types[pt.albedo_loss]}};
    json optimizer_opts = {{"otype", "ExponentialDecay"}, {"decay_start", 10000}, {"decay_interval", 2000}, {"decay_base", 0.33f},
                           {"nested", {{"otype", "Adam"}, {"learning_rate", 1e-2f}, {"beta1", 0.9f}, {"beta2", 0.999f},
                                       {"epsilon", 1e-8f}, {"l2_reg", 1e-8f}, {"relative_decay", 0.0f}, {"absolute_decay", 0.0f},
                                       {"adabound", false}}}};
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

    auto loss      = std::shared_ptr<Loss<precision_t>>{create_loss<precision_t>(loss_opts)};
    auto optimizer = std::shared_ptr<Optimizer<precision_t>>{create_optimizer<precision_t>(optimizer_opts)};
    auto network   = std::make_shared<NetworkWithInputEncoding<precision_t>>(9u, 8u, encoding_opts, network_opts);
    auto trainer   = std::make_shared<Trainer<float, precision_t, precision_t>>(network, optimizer, loss);
    const size_t n_params = network->n_params();
    printf("n_params %zu padded_output %u\n", n_params, network->padded_output_width());

    std::vector<uint16_t> params;
    if (slurp(wd + "/params_init.u16", params)) {
        if (params.size() != n_params) {
            fprintf(stderr, "params_init.u16 has %zu entries, network needs %zu\n", params.size(), n_params);
            return 3;
        }
        trainer->set_params(reinterpret_cast<const precision_t *>(params.data()), n_params, false);
    } else {
        params.resize(n_params);
        CUDA_CHECK_THROW(cudaMemcpy(params.data(), trainer->params(), n_params * 2, cudaMemcpyDeviceToHost));
        dump(wd + "/params_init.u16", params);
    }
    cudaStream_t stream = nullptr;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    float infer_ms = 0.f, train_ms = 0.f;

    // ---- inference: network->inference(stream, in 9xN, out 8xN)  (path_module.cu:92)
    std::vector<float> qin;
    if (n_query && slurp(wd + "/query_in.f32", qin) && qin.size() == (size_t)n_query * 9) {
        GPUMatrix<float> in(9, n_query), out(8, n_query);
        CUDA_CHECK_THROW(cudaMemcpy(in.data(), qin.data(), qin.size() * 4, cudaMemcpyHostToDevice));
        network->inference(stream, in, out);
        CUDA_CHECK_THROW(cudaDeviceSynchronize());
        cudaEventRecord(e0, stream);
        for (int i = 0; i < 10; i++)
            network->inference(stream, in, out);
        cudaEventRecord(e1, stream);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&infer_ms, e0, e1);
        infer_ms /= 10.f;
        std::vector<float> qout((size_t)n_query * 8);
        CUDA_CHECK_THROW(cudaMemcpy(qout.data(), out.data(), qout.size() * 4, cudaMemcpyDeviceToHost));
        dump(wd + "/query_out.f32", qout);
    }

    // ---- training: trainer->training_step(stream, in, target); trainer->loss(stream, *ctx)
    std::vector<float> tin, ttgt;
    if (steps && slurp(wd + "/train_in.f32", tin) && slurp(wd + "/train_tgt.f32", ttgt)) {
        const size_t avail = tin.size() / ((size_t)batch * 9);
        GPUMatrix<float> in(9, batch), tgt(8, batch);
        std::vector<float> losses(steps);
        for (uint32_t s = 0; s < steps; s++) {
            const size_t k = s % avail;
            CUDA_CHECK_THROW(cudaMemcpy(in.data(), tin.data() + k * batch * 9, (size_t)batch * 9 * 4, cudaMemcpyHostToDevice));
            CUDA_CHECK_THROW(cudaMemcpy(tgt.data(), ttgt.data() + k * batch * 8, (size_t)batch * 8 * 4, cudaMemcpyHostToDevice));
            cudaEventRecord(e0, stream);
            auto ctx = trainer->training_step(stream, in, tgt);
            cudaEventRecord(e1, stream);
            losses[s] = trainer->loss(stream, *ctx);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (s >= steps / 2)
                train_ms += ms;
            if (s == 0) {
                std::vector<float> L((size_t)batch * 16);
                CUDA_CHECK_THROW(cudaMemcpy(L.data(), ctx->L.data(), L.size() * 4, cudaMemcpyDeviceToHost));
                dump(wd + "/train_L0.f32", L);
                std::vector<uint16_t> g(n_params);
                CUDA_CHECK_THROW(cudaMemcpy(g.data(), trainer->param_gradients(), n_params * 2, cudaMemcpyDeviceToHost));
                dump(wd + "/grads0.u16", g);
                std::vector<uint16_t> o((size_t)batch * 16);
                CUDA_CHECK_THROW(cudaMemcpy(o.data(), ctx->output.data(), o.size() * 2, cudaMemcpyDeviceToHost));
                dump(wd + "/train_out0.u16", o);
                std::vector<uint16_t> p1(n_params);
                CUDA_CHECK_THROW(cudaMemcpy(p1.data(), trainer->params(), n_params * 2, cudaMemcpyDeviceToHost));
                dump(wd + "/params_after1.u16", p1);
            }
        }
        train_ms /= (float)(steps - steps / 2);
        dump(wd + "/train_loss.f32", losses);
        CUDA_CHECK_THROW(cudaMemcpy(params.data(), trainer->params(), n_params * 2, cudaMemcpyDeviceToHost));
        dump(wd + "/params_after.u16", params);
    }
    FILE *f = fopen((wd + "/timing.json").c_str(), "w");
    fprintf(f, "{\"n_params\": %zu, \"inference_ms\": %.5f, \"n_query\": %u, \"training_step_ms\": %.5f, \"batch\": %u}\n", n_params,
            infer_ms, n_query, train_ms, batch);
    fclose(f);
    printf("inference %.4f ms for %u queries; training step %.4f ms for batch %u\n", infer_ms, n_query, train_ms, batch);
    return 0;
}

"""Headless equivalent of the reference's validation mode (ValidationMode::Full): optimise, then render, and leave the same files
behind -- what NeuralBVHRenderer::render + process_validation_config_end do between load_next_config calls
(/root/reference/src/cuda/neural/neural_bvh_renderer.cu:113-312, neural_bvh_renderer_init.cu:260-383,406-509):

    <dir>/results/<base>_<variant>_learning_spp_0.csv      per-step time + loss   (FrameTimer::write_to_file with the loss graph)
    <dir>/nbvhs/<base>.nbvh                                 learned tree-cut       (save_neural_bvh)
    <dir>/network_models/<base>_network.json                weights                (save_network_model)
    <dir>/results/<base>_<variant>_spp_<n>.exr              accumulation / n       (save_image)
    <dir>/results/<base>_<variant>_inference_spp_<n>.csv    per-frame time
    <dir>/configs/<base>_<variant>.json                     configuration          (save_config)

All compute goes through the C ABI (module.NeuralPathModule); this file is host glue only."""
import json
import os

from . import results
from .treecut import BvhSplitScheduler, TreeCutLearner

# renderer-owned keys of save_config (neural_bvh_renderer_init.cu:479-499) that the module does not hold
_RENDERER_KEYS = ("validation_mode", "batch_size", "max_training_steps", "max_accumulated_spp", "base_config_name", "variant_config_name",
                  "run_optimisation", "run_inference", "continuous_reset", "neural_bvh_min_depth", "show_bvh", "show_neural_node_losses",
                  "show_tree_cut_depth", "tree_cut_learning", "bvh_split_scheduler")


def merged_config(mod, source_cfg, training_step=None, envmap_filename=None):
    """The full save_config document: the module-owned blocks from nbvh_save_config_json + the renderer-owned keys of source_cfg."""
    doc = json.loads(mod.save_config())
    rend = doc["neural_bvh_renderer"]
    for k in _RENDERER_KEYS:
        if k in source_cfg.get("neural_bvh_renderer", {}):
            rend[k] = source_cfg["neural_bvh_renderer"][k]
    if training_step is not None:
        rend["training_step"] = int(training_step)
    name = envmap_filename if envmap_filename is not None else source_cfg.get("cuda_scene", {}).get("envmap_filename")
    if name is not None:
        doc["cuda_scene"]["envmap_filename"] = name
    return doc


def run_full_validation(mod, cfg, out_dir, max_training_steps=None, max_accumulated_spp=None, base=None, variant=None, log=None):
    """mod: a NeuralPathModule with the scene set (neural mesh BLAS + materials included) and a framebuffer; cfg: the config dict it was
    created from.  Returns a dict with the written paths, the loss curve and the timings."""
    import torch
    rend = cfg["neural_bvh_renderer"]
    base = base or rend.get("base_config_name", "config")
    variant = variant or rend.get("variant_config_name", "neural")
    steps = int(max_training_steps if max_training_steps is not None else rend["max_training_steps"])
    spp = int(max_accumulated_spp if max_accumulated_spp is not None else rend["max_accumulated_spp"])
    for d in ("results", "nbvhs", "network_models", "configs"):
        os.makedirs(os.path.join(out_dir, d), exist_ok=True)
    paths = {}

    # ---- optimisation (m_run_optimisation): reinitialise_neural_bvh_learning_states + one step per frame
    mod.init_params(1337)
    mod.tree_cut_init(rend.get("neural_bvh_min_depth", 0))
    sched = BvhSplitScheduler.from_config(cfg)
    # the schedule's statistics start from the node count of the initial cut (neural_bvh_renderer_init.cu: update_stats_and_reset
    # (m_neural_nodes_used, ...)), and the sampler runs with the config's log2_bvh_max_traversals
    sched.update_stats_and_reset(len(mod.neural_bvh_nodes()), rend["max_training_steps"])
    learner = TreeCutLearner(mod, sched, batch_size=rend["batch_size"], num_lods_learned=rend["bvh_constants"]["num_lods_learned"],
                             log2_bvh_max_traversals=rend["bvh_constants"].get("log2_bvh_max_traversals", 0))
    learner.tree_cut_learning = bool(rend.get("tree_cut_learning", True))
    timer, losses = results.FrameTimer(), []
    torch.cuda.synchronize()
    for s in range(steps):
        timer.start_frame(s)
        loss, _ = learner.step()          # reads the loss back, like the reference's per-step loss graph
        timer.end_frame()
        losses.append(loss)
        if log and ((s + 1) % 100 == 0 or s + 1 == steps):
            log("step %d loss %.5g tree-cut nodes %d" % (s + 1, loss, len(mod.neural_bvh_nodes())))
    paths["learning_csv"] = results.result_path(out_dir, base, variant, "learning", 0)
    timer.write_to_file(paths["learning_csv"], losses)
    learning_ms = list(timer.frame_times)
    paths["nbvh"] = os.path.join(out_dir, "nbvhs", base + ".nbvh")
    mod.save_neural_bvh(paths["nbvh"])
    paths["network"] = os.path.join(out_dir, "network_models", base + "_network.json")
    mod.save_network_model(paths["network"])

    # ---- inference (m_run_inference): one neural_render per frame until max_accumulated_spp
    timer = results.FrameTimer()
    for s in range(spp):
        timer.start_frame(s)
        mod.neural_render(s, 0)
        torch.cuda.synchronize()          # the reference's frame ends with the display sync
        timer.end_frame()
    paths["image"] = results.result_path(out_dir, base, variant, "image", spp)
    try:
        mod.save_image(paths["image"], spp)
    except Exception as ex:  # e.g. ema_accumulate: the reference logs the failed save and goes on with the CSV and the config
        paths["image"] = None
        if log:
            log("image not saved: %s" % ex)
    paths["inference_csv"] = results.result_path(out_dir, base, variant, "inference", spp)
    timer.write_to_file(paths["inference_csv"])
    paths["config"] = os.path.join(out_dir, "configs", "%s_%s.json" % (base, variant))
    with open(paths["config"], "w") as f:
        json.dump(merged_config(mod, cfg, training_step=steps), f, indent=4)
        f.write("\n")
    return {"paths": paths, "losses": losses, "learning_ms": learning_ms, "inference_ms": list(timer.frame_times),
            "tree_cut_nodes": len(mod.neural_bvh_nodes()), "steps": steps, "spp": spp}

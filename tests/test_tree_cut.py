"""Tree-cut learning: per-node loss statistics, top-k split decisions, device-side split, LoD assignment,
split scheduler.

  * scheduler: the Python mirror against tests/golden/split_scheduler.json, printed by the reference's OWN BvhSplitScheduler
    header (oracle/ref_tools/split_scheduler_golden.cpp; the shipped Chess schedule ends at 25,249 nodes == the shipped .nbvh)
  * initial cut / attach / step: the CUDA path (through the C ABI) against the sequential numpy restatement in oracle/oracle.py
    -- node arrays, node map, depths, sample counts and split decisions must be IDENTICAL (integer work); loss sums within
    1e-5 relative (fp32 atomics are unordered in the reference too)
  * a short from-scratch optimisation with tree-cut learning on the synthetic scene: the cut grows as scheduled, the loss falls."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_split_scheduler_matches_reference_header():
    from nbvh_b200.treecut import BvhSplitScheduler
    g = json.load(open(os.path.join(HERE, "golden", "split_scheduler.json")))
    for case in g["cases"]:
        interval, last, start, scaling, iscaling, iters = case["args"]
        s = BvhSplitScheduler(interval, last, start, scaling, iscaling)
        s.update_stats_and_reset(1, iters)
        st = s.stats
        assert [st["total_splits"], st["num_cuts"], st["expected_bvh_num_nodes"], st["expected_bvh_depth"]] == case["out"]["stats"]
        splits = [[i, n] for i in range(iters) for n in [s.get_bvh_splits(i)] if n]
        assert splits == case["out"]["splits"]
        assert [int(s.should_assign_new_lod(i, 4, 3, 7)) for i in range(0, iters, 97)] == case["out"]["lod"]


def test_shipped_schedule_reproduces_shipped_tree_cut_size(chess_root):
    from nbvh_b200.scene_io import CHESS_BASE
    from nbvh_b200.treecut import BvhSplitScheduler
    cfg = json.load(open(os.path.join(chess_root, CHESS_BASE + "_neural.json")))
    s = BvhSplitScheduler.from_config(cfg)
    s.update_stats_and_reset(1, cfg["neural_bvh_renderer"]["max_training_steps"])
    n_file = int(np.frombuffer(open(os.path.join(chess_root, CHESS_BASE + ".nbvh"), "rb").read(8), dtype="<u8")[0])
    assert s.stats["expected_bvh_num_nodes"] == n_file == 25249


def test_oracle_tree_cut_step_semantics(oracle, synthetic):
    cfg, nbvh, tlas, tlas_idx, blas = synthetic
    bn = blas[0]["nodes"]
    nodes, cut, nmap = oracle.tree_cut_initial(bn, 3, 6)
    assert len(nodes) == 15 and (nodes["a"][nodes["b"] == 0xFFFFFFFF] == 8).all() and cut["depth"].max() == 3
    rng = np.random.default_rng(0)
    B = 4096
    leaves = np.flatnonzero(nodes["b"] == 0xFFFFFFFF)
    sampled = rng.choice(leaves, B).astype(np.uint32)
    L = np.zeros((B, 16), np.float32)
    L[:, :8] = rng.uniform(0, 1, (B, 8)) * (1 + sampled[:, None] % 5)
    n2, c2, m2, mx, ns = oracle.tree_cut_step(nodes, cut, nmap, bn, L, sampled, 3, 6)
    assert ns == 3 and len(n2) == 21 and mx > 0
    assert c2["samples"][: len(nodes)].sum() == 8 * B and (c2["iter_alive"][leaves] == 2).all()
    split = np.flatnonzero((n2["b"][: len(nodes)] != 0xFFFFFFFF) & (nodes["b"] == 0xFFFFFFFF))
    assert len(split) == 3 and (n2["a"][split] == 0).all()
    for k, idx in enumerate(np.sort(split)):  # new nodes numbered in ascending order of the split node
        assert n2["b"][idx] == len(nodes) + 2 * k
        assert (m2[n2["b"][idx]] == bn["b"][nmap[idx]]) and (c2["depth"][n2["b"][idx]] == cut["depth"][idx] + 1)


@pytest.fixture(scope="module")
def tc_env(oracle, synthetic):
    import torch
    import nbvh_b200
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg, nbvh, tlas, tlas_idx, blas = synthetic
    mod = nbvh_b200.NeuralPathModule(cfg)
    mod.set_neural_bvh(nbvh)
    mod.set_scene(tlas, tlas_idx, blas)
    mod.init_params(1337)
    yield dict(mod=mod, cfg=cfg, blas=blas, torch=torch)
    mod.close()


def same_nodes(a, b):
    return len(a) == len(b) and a.tobytes() == np.ascontiguousarray(b).tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("min_depth", [0, 4, 9])
def test_gpu_initial_cut_and_attach_match_oracle(tc_env, oracle, min_depth):
    mod, bn = tc_env["mod"], tc_env["blas"][0]["nodes"]
    mod.tree_cut_init(min_depth)
    nodes, cut, nmap = mod.tree_cut_state()
    rn, rc, rm = oracle.tree_cut_initial(bn, min_depth, 6)
    assert same_nodes(nodes, rn) and (nmap == rm).all() and cut.tobytes() == rc.tobytes()
    # attach: rebuild the learning state from the node array alone
    mod.set_neural_bvh(rn)
    mod.tree_cut_attach()
    nodes2, cut2, nmap2 = mod.tree_cut_state()
    assert same_nodes(nodes2, rn) and (nmap2 == rm).all() and (cut2["depth"] == rc["depth"]).all()


@pytest.mark.gpu
def test_gpu_tree_cut_init_resets_and_attach_rederives_the_learned_lod(oracle, synthetic):
    """build_and_upload_neural_bvh starts learning at NUM_BVH_LODS - 1 whatever the config says (accel.cpp:520; 34 of the 35 shipped
    Chess configs omit current_learned_lod), and reinitialise_neural_bvh_learning_states re-derives it from the coarsest level a
    childless node is reached at (accel.cpp:496-501)."""
    import copy
    import torch
    import nbvh_b200
    from nbvh_b200 import _capi as A
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg, nbvh, tlas, tlas_idx, blas = synthetic
    cfg = copy.deepcopy(cfg)
    cfg["neural_bvh_renderer"]["bvh_constants"].pop("current_learned_lod", None)  # a partial config: the key defaults to 0
    mod = nbvh_b200.NeuralPathModule(cfg)
    mod.set_scene(tlas, tlas_idx, blas)
    assert mod.get_constants(A.CONST_RENDERER).current_learned_lod == 0
    mod.tree_cut_init(3)
    assert mod.get_constants(A.CONST_RENDERER).current_learned_lod == 6
    nodes, _, _ = mod.tree_cut_state()
    leaves = nodes["b"] == 0xFFFFFFFF
    assert set(np.unique(nodes["a"][leaves]).tolist()) <= {1, 8}  # LEAF_FINAL or LEAF_LOD(6), like the shipped .nbvh files
    # attach on a cut whose childless nodes are all reached at level 6: learning resumes at 6
    mod2 = nbvh_b200.NeuralPathModule(cfg)
    mod2.set_scene(tlas, tlas_idx, blas)
    mod2.set_neural_bvh(nodes)
    mod2.tree_cut_attach()
    assert mod2.get_constants(A.CONST_RENDERER).current_learned_lod == 6
    # an interior node tagged LEAF_LOD(6) (= a leaf of level 6 that was split while level 5 was learned): its subtree is level 5
    rn, _, _ = oracle.tree_cut_initial(blas[0]["nodes"], 4, 5)
    deep = np.ascontiguousarray(rn).copy()
    interior = np.flatnonzero(deep["b"] != 0xFFFFFFFF)
    deep["a"][interior[len(interior) // 2:]] = 8
    mod2.set_neural_bvh(deep)
    mod2.tree_cut_attach()
    assert mod2.get_constants(A.CONST_RENDERER).current_learned_lod == 5
    mod.close()
    mod2.close()


@pytest.mark.gpu
def test_gpu_tree_cut_steps_match_oracle(tc_env, oracle):
    mod, bn, torch = tc_env["mod"], tc_env["blas"][0]["nodes"], tc_env["torch"]
    mod.tree_cut_init(2)
    nodes, cut, nmap = oracle.tree_cut_initial(bn, 2, 6)
    rng = np.random.default_rng(1)
    B = 1 << 14
    for step, k in enumerate([1, 2, 0, 5, 40, 400, 3, 0, 1000]):
        leaves = np.flatnonzero(nodes["b"] == 0xFFFFFFFF)
        sampled = rng.choice(leaves, B).astype(np.uint32)
        L = np.zeros((B, 16), np.float32)
        L[:, :8] = (rng.uniform(0, 1, (B, 8)) * (0.1 + (sampled[:, None] * 2654435761 % 97) / 97.0)).astype(np.float32)
        mx, ns = mod.tree_cut_step(torch.from_numpy(L).cuda(), torch.from_numpy(sampled.astype(np.int32)).cuda(), k, dims=8)
        nodes, cut, nmap, rmx, rns = oracle.tree_cut_step(nodes, cut, nmap, bn, L, sampled, k, 6)
        gn, gc, gm = mod.tree_cut_state()
        assert ns == rns, "step %d: %d splits, oracle %d" % (step, ns, rns)
        assert same_nodes(gn, nodes), "step %d: node arrays differ" % step
        assert (gm == nmap).all() and (gc["depth"] == cut["depth"]).all()
        assert (gc["samples"] == cut["samples"]).all() and (gc["iter_alive"] == cut["iter_alive"]).all()
        assert np.allclose(gc["loss"], cut["loss"], rtol=1e-5, atol=1e-6)
        assert abs(mx - rmx) <= 1e-5 * max(1.0, abs(rmx))
    assert len(nodes) > 150  # the cut grew well past the initial 7 nodes (each step splits at most half of the leaves)
    # inference still works on the grown cut, and agrees with the oracle walking the same nodes
    const = oracle.config_constants(tc_env["cfg"], nodes)
    o, d = mod.raygen(0, 4096)
    cnt, node, mint, maxt = mod.trace_intervals(o, d, 32)
    oc, dc = oracle.raygen(const, np.arange(4096))
    cnt_c, node_c, mint_c, maxt_c, _ = oracle.trace_intervals(nodes, const, oc, dc, 32)
    m = np.arange(32)[None, :] < np.minimum(cnt_c, 32)[:, None]
    assert (cnt.cpu().numpy().astype(np.uint32) == cnt_c).all() and (node.cpu().numpy().astype(np.uint32)[m] == node_c[m]).all()
    assert (mint.cpu().numpy().view(np.uint32)[m] == mint_c.view(np.uint32)[m]).all()


@pytest.mark.gpu
def test_gpu_from_scratch_optimisation_with_tree_cut_learning(tc_env):
    """Sampler -> training step -> tree-cut statistics / split, 120 steps from the root node: the cut follows the schedule and
    the training loss falls (the reference's loop, neural_bvh_renderer.cu:140-290, with every kernel on the device)."""
    import nbvh_b200
    from nbvh_b200.treecut import BvhSplitScheduler, TreeCutLearner
    mod, torch = tc_env["mod"], tc_env["torch"]
    tl, ti, bl = tc_env["blas"], None, None
    mod.init_params(1337)
    mod.tree_cut_init(0)
    sched = BvhSplitScheduler(splitting_interval=10, last_split_iteration=100, start_split_num=1, split_scaling=2.0, split_interval_scaling=2.0)
    sched.update_stats_and_reset(1, 120)
    learner = TreeCutLearner(mod, sched, batch_size=1 << 14)
    losses, grown = [], 0
    for _ in range(120):
        loss, ns = learner.step()
        losses.append(loss)
        grown += ns
    nodes, cut, nmap = mod.tree_cut_state()
    assert len(nodes) == 1 + 2 * grown and 0 < grown <= sched.stats["total_splits"]
    assert np.isfinite(losses).all() and np.mean(losses[-10:]) < np.mean(losses[:10])
    leaves = nodes["b"] == 0xFFFFFFFF
    assert leaves.sum() == grown + 1 and cut["depth"].max() == mod._lib.nbvh_tree_cut_max_depth(mod._h)


# ------------------------------------------------------------------------------------------------ the reference's own K17 kernels
G_TC = os.path.join(HERE, "golden", "ref_chess_treecut.npz")


def _treecut_case():
    import sys
    from nbvh_b200.scene_io import CHESS_BASE
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
    from make_ref_treecut_goldens import treecut_inputs
    if not os.path.exists(G_TC):
        pytest.skip("ref_chess_treecut.npz not generated yet (tools/make_ref_treecut_goldens.py on a GPU box)")
    g = np.load(G_TC)
    return g, treecut_inputs, CHESS_BASE


def check_against_reference_kernels(g, nodes_before, nodes_after, cut_after, max_loss, blas_nodes, nmap):
    n = len(nodes_before)
    # statistics: sample counts and leaf ages exactly, loss sums to the order of the fp32 atomics
    assert (cut_after["samples"][:n] == g["samples"]).all() and (cut_after["iter_alive"][:n] == g["iter_alive"]).all()
    assert np.allclose(cut_after["loss"][:n], g["loss"], rtol=2e-5, atol=1e-6)
    assert abs(max_loss - g["max_losses"][0]) <= 2e-5 * g["max_losses"][0]
    # decisions: the reference's k split indices (distinct maxima, best first) are exactly the leaves that were split here,
    # except those whose BLAS node is a leaf (the reference's host split skips them: "Leaf node reached, no split")
    ref_split = set(int(i) for i in g["split"])
    mine = set(np.flatnonzero((nodes_before["b"] == 0xFFFFFFFF) & (nodes_after["b"][:n] != 0xFFFFFFFF)).tolist())
    assert mine <= ref_split
    for i in ref_split - mine:
        assert nodes_before["b"][i] != 0xFFFFFFFF or blas_nodes["a"][nmap[i]] > 0 or i == 0
    assert len(mine) > 0.9 * int(g["k"])


def test_oracle_tree_cut_step_matches_reference_kernels(oracle, chess_root):
    g, treecut_inputs, base = _treecut_case()
    blas_nodes = np.fromfile(os.path.join(chess_root, "mesh0_bvh_nodes.bin"), dtype=oracle.NODE_DTYPE)
    nodes, cut0, nmap = oracle.tree_cut_initial(blas_nodes, int(g["cut_depth"]), 6)
    cut, L, sampled = treecut_inputs(nodes, int(g["batch"]), int(g["seed"]))
    n2, c2, m2, mx, ns = oracle.tree_cut_step(nodes, cut.astype(oracle.TREE_CUT_DTYPE), nmap, blas_nodes, L, sampled, int(g["k"]), 6)
    check_against_reference_kernels(g, nodes, n2, c2, mx, blas_nodes, nmap)


@pytest.mark.gpu
def test_gpu_tree_cut_step_matches_reference_kernels(oracle, chess_root):
    import torch
    import nbvh_b200
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    g, treecut_inputs, base = _treecut_case()
    mod = nbvh_b200.load_chess_module(with_neural_mesh=True)
    with pytest.raises(nbvh_b200.NbvhError):  # the shipped .nbvh is a cut of the authors' BLAS build, not of the one built here
        mod.tree_cut_attach()
    mod.tree_cut_init(int(g["cut_depth"]))
    nodes, cut0, nmap = mod.tree_cut_state()
    rn, _, rm = oracle.tree_cut_initial(np.fromfile(os.path.join(chess_root, "mesh0_bvh_nodes.bin"), dtype=oracle.NODE_DTYPE), int(g["cut_depth"]), 6)
    assert same_nodes(nodes, rn) and (nmap == rm).all()
    cut, L, sampled = treecut_inputs(nodes, int(g["batch"]), int(g["seed"]))
    cut["depth"] = cut0["depth"]
    mod.set_tree_cut_data(cut)
    mx, ns = mod.tree_cut_step(torch.from_numpy(L).cuda(), torch.from_numpy(sampled.astype(np.int32)).cuda(), int(g["k"]), dims=8)
    n2, c2, m2 = mod.tree_cut_state()
    blas_nodes = np.fromfile(os.path.join(chess_root, "mesh0_bvh_nodes.bin"), dtype=oracle.NODE_DTYPE)
    check_against_reference_kernels(g, nodes, n2, c2, mx, blas_nodes, nmap)
    assert len(n2) == len(nodes) + 2 * ns
    mod.close()

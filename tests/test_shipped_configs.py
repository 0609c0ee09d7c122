"""Every configuration file the reference ships for the Chess scene (7 node budgets x 5 hash-table sizes, 2^14 ... 2^22 entries per
level; most of them partial files that rely on the reference's defaults) loads unchanged through nbvh_create:

  * the values a file states arrive in the constant blocks; the parameter count equals the tcnn layout's (grid.h:1010-1041 offsets,
    fully_fused_mlp.cu:636-678 matrices) as computed by the oracle
  * at every shipped table size -- up to the largest, 2^22 entries x 8 levels x 4 features = 110 M parameters, far beyond L2 --
    network->inference agrees with the oracle within the fp16 tolerance of the other parity tests (2e-3 max-abs on the 8 outputs), and
    optimisation steps on a fixed batch run and reduce the loss."""
import glob
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _configs(chess_root):
    files = sorted(glob.glob(os.path.join(chess_root, "configs", "*_neural.json")))
    if len(files) < 35:
        pytest.skip("shipped config set not in the fixture (tools/make_fixtures.py copies it in the authoring container)")
    return files


def test_all_shipped_configs_load(chess_root, oracle):
    import torch
    import nbvh_b200
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    O = oracle
    seen_sizes = set()
    for path in _configs(chess_root):
        text = open(path).read()
        cfg = json.loads(text)
        mod = nbvh_b200.NeuralPathModule(text)
        rend = cfg["neural_bvh_renderer"]
        hg = rend["neural_module"]["segment_encoding"]["hashgrid_options"]
        net = mod.get_constants(nbvh_b200.CONST_NETWORK)
        assert net.log2_hashmap_size == hg["log2_hashmap_size"] and net.n_levels == hg["n_levels"], path
        seen_sizes.add(net.log2_hashmap_size)
        want, _ = O.expected_n_params(O.make_net(cfg, np.zeros(1, np.uint16)))
        assert mod.n_params() == want, path
        # derived like reinitialise_neural_model / update_neural_model_module: the partial files state only the four output dims
        assert mod.network_output_dims() == 8 and mod.get_constants(nbvh_b200.CONST_MODULE).input_network_dims == 9, path
        bc = mod.get_constants(nbvh_b200.CONST_RENDERER)
        for k, v in rend.get("bvh_constants", {}).items():
            if hasattr(bc, k):
                got = getattr(bc, k)
                assert (np.float32(got) == np.float32(v)) if isinstance(v, float) else (got == v), (path, k)
        pc = mod.get_constants(nbvh_b200.CONST_PATH)
        for k, v in rend["neural_module"].get("path_module_constants", {}).items():
            if hasattr(pc, k) and not isinstance(v, str):
                assert np.float32(getattr(pc, k)) == np.float32(v), (path, k)
        assert mod.max_depth == rend["neural_module"].get("max_depth", mod.max_depth)
        assert json.loads(mod.save_config())["cuda_scene"]["fb_size"] == cfg["cuda_scene"]["fb_size"]
        mod.close()
    assert seen_sizes == {14, 16, 18, 20, 22}


@pytest.mark.parametrize("log2", [14, 16, 18, 20, 22])
def test_query_and_training_at_every_shipped_table_size(chess_root, oracle, log2):
    import torch
    import nbvh_b200
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    O = oracle
    path = os.path.join(chess_root, "configs", "nodes_7_hash_%d_lvl_8_tsteps_10000_neural.json" % log2)
    if not os.path.exists(path):
        pytest.skip("shipped config set not in the fixture")
    text = open(path).read()
    cfg = json.loads(text)
    mod = nbvh_b200.NeuralPathModule(text)
    mod.init_params(1337)
    # grid values of the tcnn initialisation are +-1e-4: scale them up so that the table lookups matter in the comparison
    bits = mod.get_params()
    rng = np.random.default_rng(log2)
    n_mlp = 64 * 96 + 3 * 64 * 64 + 16 * 64
    vals = bits.view(np.float16).copy()
    vals[n_mlp:] = rng.uniform(-1, 1, vals.size - n_mlp).astype(np.float16)
    mod.load_params(vals.view(np.uint16))
    net = O.make_net(cfg, vals.view(np.uint16))
    x = rng.uniform(0.0, 1.0, (8192, 9)).astype(np.float32)
    y = mod.inference(torch.from_numpy(x).cuda()).cpu().numpy()
    err = float(np.abs(y - O.query(net, x, mode=0)).max())
    assert err < 2e-3, "2^%d table: query max|err| %g" % (log2, err)
    # optimisation on a fixed batch: runs at this table size and the loss falls
    mod.init_params(1337)
    B = 1 << 14
    xin = torch.from_numpy(rng.uniform(0.05, 0.95, (B, 9)).astype(np.float32)).cuda()
    tgt = np.zeros((B, 8), np.float32)
    tgt[:, 0] = rng.uniform(size=B) < 0.5
    tgt[:, 1] = rng.uniform(0, 1, B)
    n = rng.normal(size=(B, 3))
    tgt[:, 2:5] = n / np.linalg.norm(n, axis=1, keepdims=True)
    tgt[:, 5:8] = rng.uniform(0, 1, (B, 3))
    tgt_t = torch.from_numpy(tgt).cuda()
    losses = [mod.training_step(xin, tgt_t) for _ in range(30)]
    assert np.isfinite(losses).all() and np.mean(losses[-5:]) < losses[0], losses
    mod.close()

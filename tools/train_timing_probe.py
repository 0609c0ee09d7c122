import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch, nbvh_b200
root = nbvh_b200.chess_fixture_dir()
mod = nbvh_b200.load_chess_module(max_depth=1, envmap_eval=False)
B = 1 << 18
rng = np.random.default_rng(0)
x = torch.from_numpy(rng.uniform(0.05, 0.95, size=(B, 9)).astype(np.float32)).cuda()
t = torch.from_numpy(rng.uniform(0, 1, size=(B, 8)).astype(np.float32)).cuda()
for i in range(3):
    mod.training_forward_backward(x, t)
torch.cuda.synchronize()
mod.close()

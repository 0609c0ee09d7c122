import sys, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, nbvh_b200
from tools.gpu_perf import timeit
out = {}
for mode in ("split", "mega"):
    os.environ["NBVH_RENDER_MODE"] = mode
    mod = nbvh_b200.load_chess_module()
    mod.reset_counters()
    mod.neural_render(0, 0)
    torch.cuda.synchronize()
    out[mode + "_counters"] = mod.counters()
    out[mode + "_render_ms"] = timeit(lambda: mod.neural_render(0, 0), iters=30)
    mod.close()
print(json.dumps(out, indent=1))

import sys, time
sys.path.insert(0, '/root/repo')
import torch, nbvh_b200
cfg, nbvh, tlas, ti, blas, info = nbvh_b200.synthetic_instanced_scene(device="cuda", return_mesh=True)
mod = nbvh_b200.NeuralPathModule(cfg)
mv, mi = info["mesh_vertices"], info["mesh_indices"]
for _ in range(2):
    torch.cuda.synchronize(); t=time.time()
    n,p,d = mod.build_blas(mv, mi)
    torch.cuda.synchronize(); print('build ms', (time.time()-t)*1e3, 'depth', d)

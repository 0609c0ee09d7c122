"""nbvh_b200 -- B200-native (sm_100a) neural ray-query hot path of N-BVH behind the reference's module interface.

The compute lives in libnbvh_b200.so (hand-written CUDA, C ABI in include/nbvh_b200.h); this package is the
host-side mirror of the reference's NeuralPathModule plus loaders for the harness.  No CPU fallback exists:
importing works anywhere, creating a module requires the built library and an sm_100 GPU.
"""
from ._capi import (CONST_CAMERA, CONST_MODULE, CONST_NETWORK, CONST_OPTIMIZER, CONST_PATH, CONST_PT_OUTPUT,
                    CONST_RENDERER, CONST_SCENE, CONST_SCENE_BBOX, load_library)
from .module import HIT_DTYPE, MATERIAL_DTYPE, NODE_DTYPE, TREE_CUT_DTYPE, NbvhError, NeuralPathModule
from .scene_io import (chess_fixture_dir, has_shading_fixture, load_chess_module, load_scene_arrays, load_shading_arrays,
                       procedural_materials, procedural_tree_cut, synthetic_instanced_scene, synthetic_ray_batch, synthetic_scene)

__all__ = ["NeuralPathModule", "NbvhError", "NODE_DTYPE", "HIT_DTYPE", "load_library", "chess_fixture_dir",
           "load_chess_module", "load_scene_arrays", "synthetic_scene", "procedural_materials", "procedural_tree_cut",
           "MATERIAL_DTYPE", "TREE_CUT_DTYPE", "load_shading_arrays", "has_shading_fixture", "synthetic_instanced_scene", "synthetic_ray_batch"]

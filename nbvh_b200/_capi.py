"""ctypes declarations for libnbvh_b200.so (include/nbvh_b200.h).  No compute happens in Python."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NBVH_B200_LIB") or os.path.join(HERE, "libnbvh_b200.so")  # override: A/B builds (build.build_variant)

NBVH_OK = 0
ERR_NAMES = {0: "OK", -1: "INVALID_ARG", -2: "CUDA", -3: "IO", -4: "PARSE", -5: "UNSUPPORTED", -6: "STATE"}

CONST_RENDERER, CONST_MODULE, CONST_PT_OUTPUT, CONST_PATH, CONST_SCENE, CONST_CAMERA, CONST_SCENE_BBOX, \
    CONST_NETWORK, CONST_OPTIMIZER = range(9)

COUNTER_NAMES = ["rays", "queries", "nbvh_steps", "blas_steps", "tri_tests", "tile_iterations", "queued_rays",
                 "nbvh_steps_neural", "t_start_inv", "t_exhaust_inv", "t_end", "t_cta_sum",
                 "ph_walk", "ph_sync", "ph_encode", "ph_mlp", "ph_walk_d", "ph_sync_d", "ph_encode_d", "ph_mlp_d", "iters_drain"]


class RendererConstants(C.Structure):
    _fields_ = [("neural_bvh_idx", C.c_uint32), ("num_lods_learned", C.c_int32), ("current_learned_lod", C.c_int32),
                ("rendered_lod", C.c_int32), ("ignore_neural_blas", C.c_bool), ("path_lod_switch", C.c_bool),
                ("lod_switch", C.c_int32), ("leaf_boundary_sampling_ratio", C.c_float),
                ("inference_mint_epsilon", C.c_float), ("inference_maxt_epsilon", C.c_float),
                ("learning_mint_epsilon", C.c_float), ("learning_maxt_epsilon", C.c_float),
                ("sampling_bbox_scaling", C.c_float), ("fixed_node_dilation", C.c_float),
                ("node_extent_dilation", C.c_float), ("loss_exposure", C.c_float), ("show_loss_heuristic", C.c_bool),
                ("ellipsoidal_bvh", C.c_bool), ("rescale_ellipsoid", C.c_bool), ("log2_bvh_max_traversals", C.c_int32),
                ("sample_cosine_weighted", C.c_bool), ("backface_bounce_training", C.c_bool),
                ("max_training_bounces", C.c_int32), ("training_node_survival", C.c_bool),
                ("survival_probability_scale", C.c_float), ("min_rr_survival_prob", C.c_float),
                ("view_dependent_sampling", C.c_bool), ("view_dependent_ratio", C.c_float)]


class ModuleConstants(C.Structure):
    _fields_ = [("input_segment_point_dim", C.c_uint32), ("input_segment_points", C.c_uint32),
                ("input_segment_points_dim", C.c_uint32), ("input_segment_direction_dim", C.c_uint32),
                ("input_network_dims", C.c_uint32), ("segment_distribution", C.c_int32),
                ("antialias_direction", C.c_bool), ("merge_operator", C.c_bool), ("apply_encoding_trafo", C.c_bool),
                ("_pad", C.c_uint8 * 5)]


class PtOutputConstants(C.Structure):
    _fields_ = [("output_visibility_dim", C.c_uint32), ("output_depth_dim", C.c_uint32),
                ("output_normal_dim", C.c_uint32), ("output_albedo_dim", C.c_uint32),
                ("output_network_dims", C.c_uint32), ("visibility_loss", C.c_uint8),
                ("visibility_loss_scale", C.c_float), ("depth_loss", C.c_uint8), ("depth_loss_scale", C.c_float),
                ("normal_loss", C.c_uint8), ("normal_loss_scale", C.c_float), ("albedo_loss", C.c_uint8),
                ("albedo_loss_scale", C.c_float), ("depth_learning_mode", C.c_uint8),
                ("output_normal_linear_activation", C.c_bool), ("output_depth_linear_activation", C.c_bool),
                ("output_albedo_linear_activation", C.c_bool), ("visibility_masking_loss", C.c_bool),
                ("primary_visibility_threshold", C.c_float), ("secondary_visibility_threshold", C.c_float),
                ("depth_epsilon", C.c_float), ("fixed_depth_epsilon", C.c_float), ("interpolate_depth", C.c_bool),
                ("depth_interpolation_eps", C.c_float), ("_pad", C.c_uint8 * 4)]


class PathConstants(C.Structure):
    _fields_ = [("output_mode", C.c_uint8), ("normal_epsilon_offset", C.c_float), ("apply_russian_roulette", C.c_bool),
                ("rr_start_depth", C.c_int32), ("neural_roughness", C.c_float), ("neural_specular", C.c_float),
                ("neural_metallic", C.c_float), ("_pad", C.c_uint8 * 4)]


class SceneConstants(C.Structure):
    _fields_ = [("constant_roughness", C.c_float), ("constant_specular", C.c_float), ("constant_metallic", C.c_float),
                ("envmap_strength", C.c_float), ("envmap_rotation", C.c_float), ("envmap_eval", C.c_bool),
                ("emission_eval", C.c_bool), ("skylike_background", C.c_bool), ("sample_light", C.c_bool),
                ("ema_weight", C.c_float), ("ema_accumulate", C.c_bool), ("force_constant_material", C.c_bool),
                ("_pad", C.c_uint8 * 2)]


class CameraParams(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("du", C.c_float * 3), ("dv", C.c_float * 3), ("dir_top_left", C.c_float * 3)]


class Bbox3f(C.Structure):
    _fields_ = [("min", C.c_float * 3), ("max", C.c_float * 3)]


class NetworkDesc(C.Structure):
    _fields_ = [("n_levels", C.c_int32), ("n_features_per_level", C.c_int32), ("log2_hashmap_size", C.c_int32),
                ("base_resolution", C.c_int32), ("per_level_scale", C.c_float), ("n_neurons", C.c_int32),
                ("n_hidden_layers", C.c_int32)]


class OptimizerDesc(C.Structure):
    _fields_ = [("learning_rate", C.c_float), ("decay_start", C.c_int32), ("decay_interval", C.c_int32),
                ("decay_base", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("epsilon", C.c_float),
                ("l2_reg", C.c_float)]


class CudaBlas(C.Structure):
    _fields_ = [("num_nodes", C.c_uint32), ("prim_indices", C.c_void_p), ("num_prim_indices", C.c_uint32),
                ("vertices", C.c_void_p), ("indices", C.c_void_p), ("normals", C.c_void_p), ("texcoords", C.c_void_p),
                ("material_ids", C.c_void_p), ("material_map", C.c_void_p), ("num_materials", C.c_uint32),
                ("nodes", C.c_void_p)]


class SceneData(C.Structure):
    _fields_ = [("tlas_nodes", C.c_void_p), ("num_tlas_nodes", C.c_uint32), ("blas_indices", C.c_void_p),
                ("mesh_bvhs", C.c_void_p), ("num_meshes", C.c_uint32), ("materials", C.c_void_p),
                ("textures", C.c_void_p)]


CONST_TYPES = {CONST_RENDERER: RendererConstants, CONST_MODULE: ModuleConstants, CONST_PT_OUTPUT: PtOutputConstants,
               CONST_PATH: PathConstants, CONST_SCENE: SceneConstants, CONST_CAMERA: CameraParams,
               CONST_SCENE_BBOX: Bbox3f, CONST_NETWORK: NetworkDesc, CONST_OPTIMIZER: OptimizerDesc}

# every symbol include/nbvh_b200.h declares: name -> (restype, argtypes)
_vp, _u32, _u64, _i, _f = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_float
SYMBOLS = {
    "nbvh_abi_version": (_u32, []),
    "nbvh_create": (_i, [_i, C.c_char_p, C.POINTER(_vp)]),
    "nbvh_destroy": (None, [_vp]),
    "nbvh_last_error": (C.c_char_p, [_vp]),
    "nbvh_load_config_json": (_i, [_vp, C.c_char_p]),
    "nbvh_set_constants": (_i, [_vp, _i, _vp, C.c_size_t]),
    "nbvh_get_constants": (_i, [_vp, _i, _vp, C.c_size_t]),
    "nbvh_set_camera": (_i, [_vp, C.POINTER(_f), _f]),
    "nbvh_get_max_depth": (_i, [_vp]),
    "nbvh_set_max_depth": (_i, [_vp, _i]),
    "nbvh_set_scene": (_i, [_vp, C.POINTER(SceneData)]),
    "nbvh_set_neural_bvh": (_i, [_vp, _vp, _u32]),
    "nbvh_load_neural_bvh_file": (_i, [_vp, C.c_char_p]),
    "nbvh_save_neural_bvh_file": (_i, [_vp, C.c_char_p]),
    "nbvh_neural_bvh_num_nodes": (_u32, [_vp]),
    "nbvh_get_neural_bvh_host": (_i, [_vp, _vp, _u32]),
    "nbvh_load_network_file": (_i, [_vp, C.c_char_p]),
    "nbvh_save_network_file": (_i, [_vp, C.c_char_p]),
    "nbvh_n_params": (_u64, [_vp]),
    "nbvh_load_params_host": (_i, [_vp, _vp, _u64]),
    "nbvh_get_params_host": (_i, [_vp, _vp, _u64]),
    "nbvh_init_params": (_i, [_vp, _u64]),
    "nbvh_set_framebuffer": (_i, [_vp, _vp, _vp, _u32, _u32]),
    "nbvh_set_hit_buffer": (_i, [_vp, _vp]),
    "nbvh_query": (_i, [_vp, _vp, _vp, _u32, _vp]),
    "nbvh_trace_intervals": (_i, [_vp, _vp, _vp, _u32, _u32, _vp, _vp, _vp, _vp, _vp]),
    "nbvh_blas_intersect": (_i, [_vp, _u32, _vp, _vp, _vp, _u32, _vp, _vp, _vp, _vp, _vp]),
    "nbvh_raygen": (_i, [_vp, _u32, _u32, _i, _i, _vp, _vp, _vp]),
    "nbvh_encode_segments": (_i, [_vp, _vp, _vp, _vp, _vp, _u32, _vp, _vp]),
    "nbvh_render": (_i, [_vp, _u32, _u32, _u32, _u32, _vp]),
    "nbvh_trace_rays": (_i, [_vp, _vp, _vp, _u32, _u32, _vp, _vp]),
    "nbvh_render_samples": (_i, [_vp, _u32, _u32, _u32, _u32, _u32, _vp]),
    "nbvh_render_interleaved": (_i, [_vp, _u32, _u32, _u32, _u32, _vp]),
    "nbvh_build_blas": (_i, [_vp, _vp, _vp, _u32, _vp, _vp, C.POINTER(_u32), _vp]),
    "nbvh_render_host_submit": (_i, [_vp, _u32, _u32, _vp, _vp, _u32]),
    "nbvh_render_host_wait": (_i, [_vp, _u32]),
    "nbvh_build_tlas": (_i, [_vp, _vp, _u32, C.POINTER(_u32), _u32, _vp, _vp, _vp]),
    "nbvh_flip_error": (_i, [_vp, _vp, _u32, _f, _vp, C.POINTER(_f), _vp]),
    "nbvh_flip_images": (_i, [_vp, _vp, _vp, _f, _u32, _u32, _f, _vp, C.POINTER(_f), _vp]),
    "nbvh_write_exr_host": (_i, [C.c_char_p, _u32, _u32, _vp, _f]),
    "nbvh_save_image_exr": (_i, [_vp, C.c_char_p, _u32, _vp]),
    "nbvh_write_frame_times_csv_host": (_i, [C.c_char_p, _vp, _vp, _u32]),
    "nbvh_result_path": (_i, [C.c_char_p, C.c_char_p, C.c_char_p, _i, _u32, C.c_char_p, C.c_size_t]),
    "nbvh_save_config_json": (_i, [_vp, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "nbvh_image_error": (_i, [_vp, _vp, _u32, _i, _i, C.POINTER(_f), _vp]),
    "nbvh_render_host": (_i, [_vp, _u32, _u32, _vp, _vp, _vp]),
    "nbvh_selftest_division": (_i, [_vp, _u64, _u64, C.POINTER(_u64)]),
    "nbvh_get_counters": (_i, [_vp, _vp, _u32]),
    "nbvh_reset_counters": (_i, [_vp]),
    "nbvh_get_kernel_time": (_i, [_vp, C.POINTER(_f), C.POINTER(_u32)]),
    "nbvh_enable_kernel_timing": (_i, [_vp, _i]),
    "nbvh_enable_stage_timing": (_i, [_vp, _u32]),
    "nbvh_set_queue_order": (_i, [_vp, _u32, _u32, _u32, _u32, _u32]),
    "nbvh_get_stage_times": (_i, [_vp, C.POINTER(_f), C.POINTER(_f), C.POINTER(_u32)]),
    "nbvh_training_step": (_i, [_vp, _vp, _vp, _u32, _vp, C.POINTER(_f), _vp]),
    "nbvh_training_forward_backward": (_i, [_vp, _vp, _vp, _u32, _u32, _vp, _vp, _vp]),
    "nbvh_optimisation_data_generation": (_i, [_vp, _u32, _u32, _u32, _vp, _f, _vp, _vp, _vp, _vp]),
    "nbvh_sampler_walk_ahead": (_i, [_vp, _u32, _u32]),
    "nbvh_create_texture_rgba8": (_i, [_vp, _vp, _u32, _u32, C.POINTER(_u64)]),
    "nbvh_destroy_textures": (_i, [_vp]),
    "nbvh_create_envmap_rgba32f": (_i, [_vp, _vp, _u32, _u32, C.POINTER(_u64)]),
    "nbvh_set_envmap": (_i, [_vp, _u64]),
    "nbvh_gradients": (_vp, [_vp]),
    "nbvh_set_gradient_buffer": (_i, [_vp, _vp]),
    "nbvh_optimizer_step": (_i, [_vp, _vp]),
    "nbvh_training_step_count": (_u32, [_vp]),
    "nbvh_tree_cut_init": (_i, [_vp, _u32]),
    "nbvh_tree_cut_attach": (_i, [_vp]),
    "nbvh_tree_cut_step": (_i, [_vp, _vp, _u32, _u32, _vp, _u32, _u32, C.POINTER(_f), C.POINTER(_u32), _vp]),
    "nbvh_tree_cut_accumulate": (_i, [_vp, _vp, _u32, _u32, _vp, _u32, _vp, _vp, _vp]),
    "nbvh_tree_cut_decide": (_i, [_vp, _vp, _vp, _u32, C.POINTER(_f), C.POINTER(_u32), _vp]),
    "nbvh_tree_cut_assign_lod": (_i, [_vp, C.POINTER(_i)]),
    "nbvh_tree_cut_data_device": (_vp, [_vp]),
    "nbvh_neural_bvh_device": (_vp, [_vp]),
    "nbvh_set_tree_cut_host": (_i, [_vp, _vp, _u32]),
    "nbvh_get_tree_cut_host": (_i, [_vp, _vp, _vp, _u32]),
    "nbvh_tree_cut_max_depth": (_u32, [_vp]),
    "nbvh_enable_training_stage_timing": (_i, [_vp, _u32]),
    "nbvh_get_training_stage_times": (_i, [_vp, C.POINTER(_f), C.POINTER(_u32)]),
    "nbvh_nccl_unique_id": (_i, [_vp]),
    "nbvh_allreduce_setup": (_i, [_vp, _vp]),
    "nbvh_allreduce_setup_unique_id": (_i, [_vp, _vp, _i, _i]),
    "nbvh_allreduce_teardown": (_i, [_vp]),
    "nbvh_allreduce_buffer": (_i, [_vp, _vp, _u64, _i, _vp]),
    "nbvh_data_parallel_world": (_i, [_vp, C.POINTER(_i)]),
    "nbvh_nccl_comm": (_vp, [_vp]),
}
NCCL_UNIQUE_ID_BYTES = 128
DTYPE_F32, DTYPE_U32, DTYPE_I32, DTYPE_U64, DTYPE_F64 = range(5)

_lib = None


def load_library():
    """Loads the CUDA extension.  Fails loudly when it has not been built: there is no CPU fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libnbvh_b200.so is missing (%s): run `python -m nbvh_b200.build` / "
                               "__graft_entry__.build() -- nbvh_b200 has no CPU fallback" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the header and the library ever diverge
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib

"""Host-side schedule of the tree-cut learning: mirror of BvhSplitScheduler
(/root/reference/include/cuda/neural/utils/bvh_split_scheduler.h:14-150) and of the renderer's per-step control flow
(/root/reference/src/cuda/neural/neural_bvh_renderer.cu:140-290).  Pure integer logic: the device work is nbvh_tree_cut_step."""

NUM_BVH_LODS = 7  # include/cuda/neural/accel/neural_bvh2.cuh:14


class BvhSplitScheduler:
    def __init__(self, splitting_interval=50, last_split_iteration=2000, start_split_num=1, split_scaling=2.0,
                 split_interval_scaling=2.0, lod_decision_type="EqualIterations"):
        self.splitting_interval = int(splitting_interval)
        self.last_split_iteration = int(last_split_iteration)
        self.start_split_num = int(start_split_num)
        self.split_scaling = float(split_scaling)
        self.split_interval_scaling = float(split_interval_scaling)
        self.lod_decision_type = lod_decision_type
        self.stats = dict(total_splits=0, num_cuts=0, expected_bvh_num_nodes=0, expected_bvh_depth=0)
        self._reset()

    @classmethod
    def from_config(cls, cfg):
        """bvh_split_scheduler block of the config JSON (neural_bvh_renderer_init.cu:712, from_json :139-148)."""
        j = cfg.get("neural_bvh_renderer", {}).get("bvh_split_scheduler", {})
        return cls(j.get("splitting_interval", 50), j.get("last_split_iteration", 2000), j.get("start_split_num", 1),
                   j.get("split_scaling", 2.0), j.get("split_interval_scaling", 2.0))

    def _reset(self):
        self.current_num_splits = self.start_split_num
        self.current_splitting_interval = self.splitting_interval
        self.current_total_splits = 0

    def get_bvh_splits(self, it):
        """get_bvh_splits (:73-95): C++ `int *= float` truncates toward zero."""
        if it == 0:
            self._reset()
        if it > self.last_split_iteration:
            return 0
        n = 0
        if it % self.splitting_interval == 0:
            n = self.current_num_splits
            self.current_total_splits += n
        if it % self.current_splitting_interval == 0:
            self.current_num_splits = int(self.current_num_splits * self.split_scaling)
            self.current_splitting_interval = int(self.current_splitting_interval * self.split_interval_scaling)
        return n

    def update_stats_and_reset(self, bvh_start_num_nodes, num_iterations):
        """update_stats_and_reset (:38-71).  Returns the cumulative-splits graph."""
        import math
        self._reset()
        st = dict(total_splits=0, num_cuts=0, expected_bvh_num_nodes=int(bvh_start_num_nodes), expected_bvh_depth=0)
        graph = []
        for i in range(num_iterations):
            leaves = (st["expected_bvh_num_nodes"] + 1) // 2
            s = self.get_bvh_splits(i)
            if s > 0:
                s = min(s, leaves)
                st["total_splits"] += s
                st["expected_bvh_num_nodes"] += 2 * s
                st["num_cuts"] += 1
            graph.append(float(st["total_splits"]))
        st["expected_bvh_depth"] = int(math.log2(float(st["expected_bvh_num_nodes"]))) if st["expected_bvh_num_nodes"] > 0 else 0
        self.stats = st
        self._reset()
        return graph

    def should_assign_new_lod(self, it, current_learned_lod, num_lods_learned, max_bvh_lods=NUM_BVH_LODS):
        """should_assign_new_lod (:97-125); integer arithmetic as in C++ (left-to-right, truncating division)."""
        if self.lod_decision_type == "EqualIterations":
            nxt = (max_bvh_lods - current_learned_lod) * self.last_split_iteration // num_lods_learned
            assign = nxt <= it
        else:
            nxt = (max_bvh_lods - current_learned_lod) * self.stats["total_splits"] // num_lods_learned
            assign = nxt <= self.current_total_splits
        return assign or it >= self.last_split_iteration


class TreeCutLearner:
    """Per-step control flow of NeuralBVHRenderer's optimisation loop around one NeuralPathModule
    (neural_bvh_renderer.cu:140-290): sample generation -> training step -> tree-cut statistics / split -> LoD assignment."""

    def __init__(self, module, scheduler, batch_size, num_lods_learned=1, log2_bvh_max_traversals=0, rank=0, world=1):
        """batch_size = slots of THIS rank; with world > 1 the ranks train data parallel: gradients and tree-cut statistics are
        allreduced (torch.distributed must be initialised), weights and tree-cut stay identical on all ranks."""
        import torch
        self.rank, self.world = int(rank), int(world)
        self.mod, self.sched = module, scheduler
        self.batch = int(batch_size)
        self.num_lods_learned = int(num_lods_learned)
        self.log2_trav = int(log2_bvh_max_traversals)
        dev = torch.device("cuda", module.device)
        self.sampled = torch.zeros(self.batch, dtype=torch.int32, device=dev)
        # the sampler and the trainer stride these by the config's dims (3 * input_segment_points, output_network_dims)
        from ._capi import CONST_MODULE
        in_dims = int(module.get_constants(CONST_MODULE).input_network_dims)
        out_dims = int(module.network_output_dims())
        self.inputs = torch.zeros((self.batch, in_dims), dtype=torch.float32, device=dev)
        self.targets = torch.zeros((self.batch, out_dims), dtype=torch.float32, device=dev)
        self.L = torch.zeros((self.batch, 16), dtype=torch.float32, device=dev)
        self.max_node_loss = 0.0
        self.step_no = 0
        self.tree_cut_learning = True
        self.all_lods_assigned = False
        self.walk_ahead = 16

    def step(self, stream=None):
        from ._capi import CONST_RENDERER
        m = self.mod
        # walk-ahead of the sample generator: up to 16 steps per walk launch, but not past the next scheduled split (a split changes the
        # tree-cut, and with it the walks of every later step)
        ahead = self.walk_ahead
        if self.tree_cut_learning and self.step_no <= self.sched.last_split_iteration:
            iv = max(1, self.sched.splitting_interval)
            ahead = min(ahead, ((self.step_no + iv - 1) // iv) * iv - self.step_no + 1)
        m.sampler_walk_ahead(max(1, ahead), self.world)
        current_max_loss = 0.0 if self.step_no == 0 else self.max_node_loss
        # the sampler's RNG is keyed by training_step * batch + slot: step * world + rank gives every rank its own global slots
        m.optimisation_data_generation(self.step_no * self.world + self.rank, self.batch, self.log2_trav, self.sampled, current_max_loss,
                                       m.tree_cut_data_ptr(), self.inputs, self.targets, stream=stream)
        n_splits = self.sched.get_bvh_splits(self.step_no) if self.tree_cut_learning else 0
        if self.world == 1:
            loss = m.training_step(self.inputs, self.targets, per_sample_loss=self.L, want_loss=True, stream=stream)
            self.max_node_loss, n_split = m.tree_cut_step(self.L, self.sampled, n_splits, stream=stream)
        else:
            import torch
            import torch.distributed as dist
            if not hasattr(self, "grads"):
                self.grads = m.gradient_buffer()
                self.loss_sum = torch.zeros(1, dtype=torch.float32, device=self.L.device)
            self.loss_sum.zero_()
            m.training_forward_backward(self.inputs, self.targets, global_batch=self.batch * self.world, per_sample_loss=self.L,
                                        loss_sum=self.loss_sum, stream=stream)
            if m.data_parallel_world()[1] <= 1:  # no communicator inside the library (setup_data_parallel): the host's process group sums
                dist.all_reduce(self.grads)
                dist.all_reduce(self.loss_sum)
            m.optimizer_step(stream=stream)
            loss = float(self.loss_sum.item())
            self.max_node_loss, n_split = m.tree_cut_step_data_parallel(self.L, self.sampled, n_splits, stream=stream)
        if self.tree_cut_learning and self.num_lods_learned > 1 and not self.all_lods_assigned:
            lod = m.get_constants(CONST_RENDERER).current_learned_lod
            if self.sched.should_assign_new_lod(self.step_no, lod, self.num_lods_learned):
                self.all_lods_assigned = m.tree_cut_assign_lod()
                if self.all_lods_assigned:
                    self.tree_cut_learning = False
        self.step_no += 1
        return loss, n_split

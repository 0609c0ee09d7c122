// TEST INFRASTRUCTURE: prints the schedule of the reference's OWN BvhSplitScheduler (included from /root/reference where it lies) for
// tests/golden/split_scheduler.json.  Build + run: tools/make_scheduler_goldens.sh (authoring container only).
#include <cstdio>
#include <cassert>
#include <cmath>
#include <json/json.hpp>
#define UNREACHABLE() assert(false)
#include "cuda/neural/utils/bvh_split_scheduler.h"
int main(int argc, char** argv) {
    ntwr::neural::BvhSplitScheduler s;
    s.m_splitting_interval = atoi(argv[1]); s.m_last_split_iteration = atoi(argv[2]); s.m_start_split_num = atoi(argv[3]);
    s.m_split_scaling = atof(argv[4]); s.m_split_interval_scaling = atof(argv[5]);
    int iters = atoi(argv[6]);
    std::vector<float> g;
    s.update_stats_and_reset(1, g, iters);
    printf("{\"stats\": [%d, %d, %d, %d], \"splits\": [", s.m_stats.total_splits, s.m_stats.num_cuts, s.m_stats.expected_bvh_num_nodes, s.m_stats.expected_bvh_depth);
    bool first = true;
    for (int i = 0; i < iters; i++) { int n = s.get_bvh_splits(i); if (n) { printf("%s[%d, %d]", first ? "" : ", ", i, n); first = false; } }
    printf("], \"lod\": [");
    for (int i = 0; i < iters; i += 97) printf("%s%d", i ? ", " : "", (int)s.should_assign_new_lod(i, 4, 3, 7));
    printf("]}\n");
}

// TEST INFRASTRUCTURE -- see cuda_runtime.h in this directory
#pragma once
void kry_emul_grid_sync();
namespace cooperative_groups {
struct grid_group {
    void sync() { kry_emul_grid_sync(); }
};
static inline grid_group this_grid() { return grid_group(); }
}   // namespace cooperative_groups

// plan.hpp -- internal (C++) view of the host-side planners in plan.cpp
#pragma once
#include "bicgstab_b200.h"
#include <vector>

namespace bicg {

struct HaloRun { int first; int len; int owner; };   // global columns [first, first+len) owned by `owner`

int  plan_tiles(const unsigned *ptr, int rows, int rows_per_tile, int cap_nnz, std::vector<int> &tile_row);
void plan_halo_runs(const CSR_Matrix *offd, const INFO_Matrix *info, int world, int gap, int self,
                    std::vector<HaloRun> &runs);

} // namespace bicg

// plan.hpp -- internal (C++) view of the host-side planners in plan.cpp
#pragma once
#include "bicgstab_b200.h"
#include <vector>

namespace bicg {

struct HaloRun { int first; int len; int owner; };
struct PushRunHost { int src; int len; int dst_off; };   // local rows [src, src+len) -> ghost slots dst_off.. on the peer   // global columns [first, first+len) owned by `owner`

int  plan_tiles(const unsigned *ptr, int rows, int rows_per_tile, int cap_nnz, std::vector<int> &tile_row);
unsigned plan_cta_tiles(const unsigned *ptr, int rows, int ctas, int rows_per_tile, const unsigned char *row_extra,
                        int extra_weight, std::vector<int> &tile_row, std::vector<int> &cta_tile, int cap_limit,
                        std::vector<unsigned> *tile_nz, std::vector<int> *tile_flag, int row_weight);
void plan_halo_runs(const CSR_Matrix *offd, const INFO_Matrix *info, int world, int gap, int self,
                    std::vector<HaloRun> &runs);

void merge_blocks(const CSR_Matrix *diag, const CSR_Matrix *offd, const INFO_Matrix *info, int self, int world, int gap,
                  int ghost_off, std::vector<unsigned> &mptr, std::vector<unsigned> &mcol, std::vector<double> &mval,
                  std::vector<int> &recv, int &n_ghost);
void plan_merged_layout(const CSR_Matrix *diag, const CSR_Matrix *offd, const INFO_Matrix *info, int self, int world, int gap,
                        std::vector<unsigned> &mptr, std::vector<int> &recv, int &n_ghost);
void plan_push_runs(const int *all_recv, const int *cnts, int stride, int self, int dest, int my_first,
                    std::vector<PushRunHost> &out);

} // namespace bicg

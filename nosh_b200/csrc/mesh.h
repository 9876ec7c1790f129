// mesh.h -- host entry points of the device mesh pipeline (mesh.cu)
#pragma once
#include "common.cuh"
namespace nosh {
void partition_range(int64_t n_global, int nranks, int rank, int64_t group, int64_t *begin, int64_t *end,
                     int64_t *group_used, int64_t *ngroups_out, int64_t *gbegin, int64_t *gcount);
void setup_partition(Ctx *ctx, int64_t n_global);
void mesh_from_host(Ctx *ctx, int dim, int64_t nv, const double *coords, int64_t ncells,
                    const int32_t *cells);
void mesh_from_host_local(Ctx *ctx, int dim, int64_t n_global, int64_t nv_local, const int64_t *gids,
                          const double *coords, int64_t ncells, const int32_t *cells);
void mesh_tetgrid(Ctx *ctx, int nx, int ny, int nz, const double lo[3], const double hi[3],
                  double jitter, uint64_t seed);
}  // namespace nosh

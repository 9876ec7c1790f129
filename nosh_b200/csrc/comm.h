// comm.h -- NCCL plumbing (comm.cu)
#pragma once
#include "common.cuh"
namespace nosh {
void comm_unique_id(void *id128);
void comm_init(Ctx *ctx, const void *id128, int rank, int nranks);
void comm_destroy(Ctx *ctx);
void comm_allreduce_sum(Ctx *ctx, const double *send, double *recv, int64_t n);
void halo_setup(Ctx *ctx);
void halo_exchange(Ctx *ctx, double2 *vec, cudaStream_t stream = nullptr);
// peer-memory path (CUDA IPC over NVLink): set-up after the work vectors exist
void p2p_setup(Ctx *ctx);
void p2p_teardown(Ctx *ctx);
void p2p_halo_push(Ctx *ctx, int which_r, const double2 *vec);
int p2p_check_error(Ctx *ctx);
}  // namespace nosh

// comm.h -- multi-GPU plumbing: peer memory + set-up exchange (comm.cu)
#pragma once
#include "common.cuh"
namespace nosh {
void comm_unique_id(void *id128);
void comm_init(Ctx *ctx, const void *id128, int rank, int nranks);
// set-up through the caller's communicator (host all-gather callback), data path over peer memory only
void comm_init_host(Ctx *ctx, int rank, int nranks, HostAllgather fn, void *user);
void comm_destroy(Ctx *ctx);
// set-up only: all-gather of one fixed-size HOST record per rank (caller's communicator or NCCL)
void exchange_allgather(Ctx *ctx, const void *send, void *recv, size_t bytes);
void comm_allreduce_sum(Ctx *ctx, const double *send, double *recv, int64_t n);
void halo_setup(Ctx *ctx);
void halo_exchange(Ctx *ctx, double2 *vec, cudaStream_t stream = nullptr);
// peer-memory path (CUDA IPC over NVLink): set-up after the work vectors exist
void p2p_setup(Ctx *ctx);
void p2p_teardown(Ctx *ctx);
void p2p_halo_push(Ctx *ctx, int which_r, const double2 *vec);
// stand-alone exchange over peer memory (p2p.ok): begin = push + signal, end = wait; the ghosts of this
// exchange are at the returned pointer (Ng entries) and, with ghost_out != NULL, also copied there
void halo_begin(Ctx *ctx, const double2 *vec, cudaStream_t stream = nullptr);
const double2 *halo_end(Ctx *ctx, double2 *ghost_out, cudaStream_t stream = nullptr);
const double2 *halo_slot(Ctx *ctx);
unsigned long long halo_next_epoch(Ctx *ctx);
int p2p_check_error(Ctx *ctx);
}  // namespace nosh

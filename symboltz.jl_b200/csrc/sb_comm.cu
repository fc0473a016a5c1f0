// symboltz.jl_b200 -- multi-GPU exchange steps of the hot path behind the C ABI (libsbc.so): the library owns the NCCL communicator,
// the host language only passes an opaque handle (SURVEY §8b "library owns device buffers, streams and NCCL communicator behind the handle").
//
// Replaces, for a host without torch.distributed (Julia over ccall): the fan-out of src/solve.jl:566 (`Threads.@spawn` per mode) across the
// GPUs of one box -- one process (or one host thread) per GPU:
//   rank r integrates the modes r, r + world, ... (cost grows with k: strided ownership balances it)        sbm_solvept_src
//   sbc_allreduce_sum over the zero-initialised full S[nk][nS][nt] (disjoint supports: the sum is an all-gather)
//   rank r runs the line of sight for its contiguous slice of fine wavenumbers                               sbl_los(k0, nk)
//   partial C_l over that slice                                                                               sbl_cl(k0, k1)
//   sbc_allreduce_sum over C_l[nmodes][nl]  (north star item 4: "NCCL all-reduce of the partial C_l sums")
// and for parameter sweeps: sbc_allreduce_sum over P[ncosmo][nk] with every rank filling the rows of its own cosmologies ("gather of P(k)").
// The messages are 3 KB ... 10 MB: latency-bound, so plain NCCL ring/tree (NVLS on NVSwitch when available) is the right tool; there is no
// compute step to fuse a transfer into -- the solve that produces S runs for tens of milliseconds before the 9.7 MB exchange.
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdio.h>
#include <string.h>

#define SBC_CHECK_NCCL(x)                                                                                             \
    do {                                                                                                              \
        ncclResult_t r_ = (x);                                                                                        \
        if (r_ != ncclSuccess) { fprintf(stderr, "symboltz_b200(comm): NCCL error %s at %s:%d\n", ncclGetErrorString(r_), __FILE__, __LINE__); return -2000 - (int)r_; } \
    } while (0)

extern "C" {

int sbc_unique_id_bytes(void) { return (int)sizeof(ncclUniqueId); }

// Rank 0 creates the id (out: sbc_unique_id_bytes() = 128 bytes) and hands it to the other ranks by any host-side means (file, socket, MPI, ...).
int sbc_unique_id(char* out) {
    ncclUniqueId id;
    SBC_CHECK_NCCL(ncclGetUniqueId(&id));
    memcpy(out, &id, sizeof(id));
    return 0;
}

// Collective over all ranks: create the communicator for the CURRENT CUDA device of the calling thread.  *comm receives the opaque handle.
int sbc_comm_init(const char* id_bytes, int rank, int world, void** comm) {
    if (!id_bytes || !comm || world < 1 || rank < 0 || rank >= world) return -1;
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    ncclComm_t c;
    SBC_CHECK_NCCL(ncclCommInitRank(&c, world, id, rank));
    *comm = (void*)c;
    return 0;
}

int sbc_comm_destroy(void* comm) {
    if (!comm) return 0;
    SBC_CHECK_NCCL(ncclCommDestroy((ncclComm_t)comm));
    return 0;
}

// In-place sum over ranks of dbuf[n] (device), asynchronous on `stream`.
int sbc_allreduce_sum(void* comm, double* dbuf, long long n, void* stream) {
    if (!comm || n < 0) return -1;
    if (n == 0) return 0;
    SBC_CHECK_NCCL(ncclAllReduce(dbuf, dbuf, (size_t)n, ncclDouble, ncclSum, (ncclComm_t)comm, (cudaStream_t)stream));
    return 0;
}

// Ownership rules of the sharded paths, so that every host language uses the same ones:
// modes (and cosmologies of a sweep) are strided, fine wavenumbers of the line of sight are contiguous slices.
int sbc_owned_count(int n, int rank, int world) { return n <= rank ? 0 : (n - rank + world - 1) / world; }
int sbc_owned_index(int j, int rank, int world) { return rank + j * world; }
int sbc_slice_begin(int n, int rank, int world) { return (int)(((long long)n * rank) / world); }
int sbc_slice_end(int n, int rank, int world) { return (int)(((long long)n * (rank + 1)) / world); }

} // extern "C"

// Row-sharded chains: the exchange window every rank exposes to its peers, and the peer-memory Gram sum.
//
// The reference has no multi-device path; this is the B200-native design of SURVEY.md 8(e): ranks
// hold disjoint rows of X / eps, the chain is replicated, and the only per-block exchange -- the B partial dots -- is
// written by the reducer warps of the persistent sweep kernel straight into every peer's window over NVLink.
#pragma once
#include "sweep.cuh"

namespace brr {

// What a rank tells the others about itself (all-gathered through brr_comm::allgather at creation).
struct PeerBlob {
    int32_t pid, device;
    uint64_t nonce;                // random per-process token: equal nonce and pid <=> the peer is a thread of this process
    uint64_t base;                 // device address of the window in the owner's address space
    cudaIpcMemHandle_t handle;     // for peers in other processes
    int64_t n_local;               // rows of this rank
    int32_t block, kind;
    int64_t M;
};

// One cudaMalloc per rank.  Everything up to `off_eps` has the same offset on every rank.
struct Window {
    int rank = 0, R = 1;
    uint8_t *base = nullptr; size_t bytes = 0;
    size_t off_xred = 0, off_xfin = 0, off_ready = 0, off_gram = 0, off_eps = 0;
    uint8_t *peer[MAXR] = {};      // base of every rank's window as seen from this rank (peer[rank] == base)
    bool ipc_opened[MAXR] = {};
    int64_t n_rows[MAXR] = {}, row0[MAXR] = {};
    int64_t n_total = 0;
    bool keep_on_release = false;    // a failed chain leaks its window instead of freeing it (peers may still be writing)
    int colocated = 1;               // ranks of this chain that run on this rank's device (threads of one process: a test configuration)

    void layout(int PS, int nb, int B, int64_t Npad, int gram_elem_bytes = 4);
    void allocate();
    void connect(const brr_comm &comm, int device, int64_t n_local, int B, int kind, int64_t M);   // collective
    void release();
    uint64_t *xred(int r) const { return reinterpret_cast<uint64_t *>(peer[r] + off_xred); }
    uint64_t *xfin(int r) const { return reinterpret_cast<uint64_t *>(peer[r] + off_xfin); }
    uint32_t *ready(int r) const { return reinterpret_cast<uint32_t *>(peer[r] + off_ready); }
    size_t gram_bytes = 0;           // one partial-Gram buffer (two alternate: iteration parity)
    int32_t *gram(int r, int buf) const { return reinterpret_cast<int32_t *>(peer[r] + off_gram + (size_t)buf * gram_bytes); }
    double *eps(int r) const { return reinterpret_cast<double *>(peer[r] + off_eps); }
};

// G_sum = sum over ranks of their partial block Grams (exact int32), every rank reading its peers' partials over NVLink.
// Signals "my partial of iteration `epoch - 1` is complete" to every peer first, then waits for theirs.
// f64: the partials are fp64 tiles (stores with dense columns), summed in rank order
void launch_gram_allsum(const Window &w, int buf, uint32_t epoch, void *d_sum, size_t n_elems, bool f64, int *abort_flag, cudaStream_t stream, int max_ctas);

void comm_check(int rc, const char *what);
void comm_allreduce(const brr_comm &comm, double *buf, int64_t n);

}  // namespace brr

// Host-side handle definitions shared by genotypes.cu and net.cu.
#pragma once

#include <string.h>

#include "comm.cuh"
#include "common.cuh"

struct bann_ctx {
    int device = 0;
    int rank = 0;
    int world = 1;
    int num_sms = 148;
    int cc_major = 10;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    // cross-rank sums over peer memory (comm.cuh / comm.cu); connected by bann_ctx_comm_connect
    uint2* xr_inbox[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool xr_ipc[8] = {false, false, false, false, false, false, false, false};   // opened with cudaIpcOpenMemHandle
    bool xr_connected = false;
    uint32_t xr_epoch = 0;
};

namespace bann {
// the descriptor of the NEXT exchange (advances the epoch); single-rank contexts get world = 1
XrComm xr_next(bann_ctx* ctx, int* error_flag);
void xr_release(bann_ctx* ctx);
// peer-memory plumbing shared by the context's inbox and a net's bulk exchange region (comm.cu)
int comm_export(void* dptr, int device, uint8_t* out /* BANN_COMM_HANDLE_BYTES */);
int comm_import(bann_ctx* ctx, const uint8_t* blob, void** out, bool* opened_ipc);
}  // namespace bann

struct bann_genotypes {
    bann_ctx* ctx = nullptr;
    uint64_t n = 0;        // local rows
    uint64_t n_total = 0;  // rows over all ranks
    uint64_t m = 0;
    uint64_t num_branches = 0;
    uint32_t ntiles = 0;
    uint64_t total_cols = 0;
    uint64_t store_bytes = 0;
    uint64_t packed_bytes = 0;  // sum_b m_b * ceil(n/4): the algorithmic genotype bytes of one pass
    std::vector<uint32_t> m_b, m_pad4;
    std::vector<uint64_t> tile_off, col_off;
    uint8_t* d_store = nullptr;
    // tensor-core store (only when every branch has <= 2048 markers): per branch, per 256-row super-tile,
    // [ceil(m/8) chunks][128 row pairs] 32-bit words, see k_build_tc
    uint32_t* d_store_tc = nullptr;
    uint32_t nst = 0;            // super-tiles of 256 rows
    uint64_t tc_bytes = 0;
    std::vector<uint64_t> tc_off;
    float* d_means = nullptr;
    float* d_stds = nullptr;
    float* d_mu = nullptr;  // per-branch gathered means (sum m_b)
    float* d_sd = nullptr;
    uint64_t* d_col_ids = nullptr;
    unsigned long long* d_counts = nullptr;
};

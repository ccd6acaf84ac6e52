// Device-resident 2-bit genotype store.
//
// Replaces BedVM + MarkerGrouping + GroupedGenotypes::x_group_af of the reference
// (io/bed.rs:123-133,193-245,325-355; group/grouping.rs:7-15; data/genotypes.rs:44-48).
// The reference LUT-decodes a branch's columns to f32 on the host and uploads N x m_b floats at
// every branch visit.  Here the packed payload is uploaded once and re-tiled into a
// branch-major, row-tile-major layout so that the bytes one CTA needs for one (branch, 128-row
// tile) are one contiguous block:
//
//   tile(b, t) = [32 row-quads][m_pad4 bytes] (m_pad4 = 4 * odd);  byte(q, j) holds the four individuals
//   t*128 + 4q .. +3 (LSB first, as in PLINK) of the branch's j-th marker.
//
// Codes are re-encoded from PLINK's {00->2, 01->missing(0), 10->1, 11->0}
// (io/bed_lookup_tables.rs:4) to the genotype value itself {00->0, 01->1, 10->2}; rows beyond
// n and marker padding hold 0.  Standardisation (bed.rs:354) is folded into the first layer by
// the kernels; bann_genotypes_decode_branch reproduces (g - mean)/std bit-exactly for tests.
#include "store.cuh"

namespace bann {

thread_local std::string g_last_error;
unsigned long long g_launch_count = 0;
void set_error(const std::string& s) { g_last_error = s; }

// ------------------------------------------------------------------ kernels
__device__ __forceinline__ uint8_t plink_to_value_codes(uint8_t b) {
    // per 2-bit field (hi,lo): 00->10, 01->00, 10->01, 11->00
    uint8_t H = (b >> 1) & 0x55, L = b & 0x55;
    uint8_t ghi = (uint8_t)(~H & ~L) & 0x55;
    uint8_t glo = (uint8_t)(H & ~L) & 0x55;
    return (uint8_t)((ghi << 1) | glo);
}

// Sequential f32 column statistics, exactly as io/bed.rs:231-238: mean = sum(v)/n,
// std = sqrt(sum((v-mean)^2)/n), both sums accumulated left to right in f32.
__global__ void k_col_stats(const uint8_t* __restrict__ payload, uint64_t n, uint64_t m, uint64_t bpc,
                            float* __restrict__ means, float* __restrict__ stds) {
    uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (j >= m) return;
    const uint8_t* col = payload + j * bpc;
    const float lut[4] = {2.f, 0.f, 1.f, 0.f};
    float s = 0.f;
    for (uint64_t i = 0; i < n; ++i) {
        uint8_t b = col[i >> 2];
        s = __fadd_rn(s, lut[(b >> (2 * (i & 3))) & 3]);
    }
    float mean = __fdiv_rn(s, (float)n);
    float ss = 0.f;
    for (uint64_t i = 0; i < n; ++i) {
        uint8_t b = col[i >> 2];
        float d = __fsub_rn(lut[(b >> (2 * (i & 3))) & 3], mean);
        ss = __fadd_rn(ss, __fmul_rn(d, d));
    }
    means[j] = mean;
    stds[j] = __fsqrt_rn(__fdiv_rn(ss, (float)n));
}

__global__ void k_col_counts(const uint8_t* __restrict__ payload, uint64_t n, uint64_t m, uint64_t bpc,
                             unsigned long long* __restrict__ out) {
    uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (j >= m) return;
    const uint8_t* col = payload + j * bpc;
    unsigned long long c[3] = {0, 0, 0};
    const int val[4] = {2, 0, 1, 0};
    for (uint64_t i = 0; i < n; ++i) c[val[(col[i >> 2] >> (2 * (i & 3))) & 3]]++;
    out[3 * j] = c[0]; out[3 * j + 1] = c[1]; out[3 * j + 2] = c[2];
}

// One block per (branch, tile): gather the branch's columns, re-encode, transpose to [q][j].
__global__ void k_build_tiles(const uint8_t* __restrict__ payload, uint64_t n, uint64_t bpc,
                              const uint64_t* __restrict__ tile_branch_start,  // prefix of tiles per branch
                              const BranchDesc* __restrict__ descs, const uint64_t* __restrict__ col_ids,
                              uint32_t num_branches, uint32_t ntiles, uint8_t* __restrict__ store) {
    extern __shared__ uint8_t sm[];
    uint64_t gt = blockIdx.x;  // global tile index = b * ntiles + t
    uint32_t b = (uint32_t)(gt / ntiles), t = (uint32_t)(gt % ntiles);
    const BranchDesc& d = descs[b];
    const uint64_t* cols = col_ids + d.col_off;
    uint32_t mp = d.m_pad4;
    uint32_t total = mp * kTileQuads;
    for (uint32_t idx = threadIdx.x; idx < total; idx += blockDim.x) {
        uint32_t j = idx / kTileQuads, q = idx % kTileQuads;
        uint8_t v = 0;
        uint64_t src = (uint64_t)t * kTileQuads + q;  // byte within the column
        if (j < d.m && src < bpc) {
            v = plink_to_value_codes(payload[cols[j] * bpc + src]);
            uint64_t row0 = src * 4;
            if (row0 + 4 > n) {  // mask the individuals beyond n (PLINK pads with 00 == value 2)
                uint32_t valid = (uint32_t)(n - row0);
                v &= (uint8_t)((1u << (2 * valid)) - 1u);
            }
        }
        sm[q * mp + j] = v;
    }
    __syncthreads();
    uint8_t* dst = store + d.tile_off + (uint64_t)t * total;
    for (uint32_t idx = threadIdx.x; idx < total; idx += blockDim.x) dst[idx] = sm[idx];
}

// Tensor-core store (k1_tc.cuh).  One block per (branch, 256-row super-tile), thread t = row pair
// (256 s + t, 256 s + 128 + t).  Word i of the thread holds the pair's i-th 8-marker chunk: genotype value
// codes (00->0, 01->1, 10->2; PLINK missing and padding -> 0) of marker 8i + 2p     at bits [2p+1 : 2p]      (row A)
//                                                              marker 8i + 2p + 1 at bits [16+2p+1 : 16+2p]  (row A)
// and the same two markers of row B 8 bits higher (p = 0..3).  `x & (0x00030003 << 2p)` is then the pair of
// bf16 subnormals g * 4^p * 2^-133 of row A, `(x >> 8) & ...` that of row B.
__device__ __forceinline__ uint32_t plink_value_code(const uint8_t* __restrict__ payload, uint64_t col, uint64_t bpc,
                                                     uint64_t row) {
    const uint32_t c = (payload[col * bpc + (row >> 2)] >> (2 * (row & 3))) & 3u;
    return c == 0 ? 2u : (c == 2 ? 1u : 0u);   // io/bed_lookup_tables.rs:4
}
__global__ void __launch_bounds__(128) k_build_tc(const uint8_t* __restrict__ payload, uint64_t n, uint64_t bpc,
                                                  const BranchDesc* __restrict__ descs,
                                                  const uint64_t* __restrict__ col_ids, uint32_t nst,
                                                  uint32_t* __restrict__ store_tc) {
    const uint32_t b = blockIdx.x / nst, s = blockIdx.x % nst, t = threadIdx.x;
    const BranchDesc& d = descs[b];
    const uint64_t* cols = col_ids + d.col_off;
    const uint64_t rowA = (uint64_t)s * 256 + t, rowB = rowA + 128;
    uint32_t* dst = store_tc + (d.tc_off >> 2) + (size_t)s * d.nc * 128 + t;
    for (uint32_t i = 0; i < d.nc; ++i) {
        uint32_t w = 0;
        for (uint32_t p = 0; p < 4; ++p)
            for (uint32_t hi = 0; hi < 2; ++hi) {
                const uint32_t j = 8 * i + 2 * p + hi;
                if (j >= d.m) continue;
                const uint32_t sh = 16 * hi + 2 * p;
                if (rowA < n) w |= plink_value_code(payload, cols[j], bpc, rowA) << sh;
                if (rowB < n) w |= plink_value_code(payload, cols[j], bpc, rowB) << (sh + 8);
            }
        dst[(size_t)i * 128] = w;
    }
}

// test hook: decode branch b from the tensor-core store exactly as k1_tc expands it
__global__ void k_decode_branch_tc(const uint32_t* __restrict__ store_tc, BranchDesc d, uint64_t n,
                                   const float* __restrict__ mu, const float* __restrict__ sd, int standardized,
                                   float* __restrict__ out) {
    uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (idx >= n * d.m) return;
    const uint64_t i = idx % n, j = idx / n;
    const uint64_t s = i / 256;
    const uint32_t r = (uint32_t)(i % 256), t = r & 127u, rowB = r >> 7;
    uint32_t x = store_tc[(d.tc_off >> 2) + (size_t)s * d.nc * 128 + (size_t)(j >> 3) * 128 + t];
    if (rowB) x >>= 8;
    const uint32_t p = (uint32_t)((j & 7u) >> 1);
    const uint32_t pair = x & (0x00030003u << (2 * p));               // two bf16 bit patterns
    const uint32_t bits = (j & 1u) ? (pair >> 16) : (pair & 0xFFFFu);
    float g = (float)(bits >> (2 * p));                               // the subnormal's integer significand / 4^p
    if (standardized) g = __fdiv_rn(__fsub_rn(g, mu[d.col_off + j]), sd[d.col_off + j]);  // bed.rs:354
    out[idx] = g;
}

// test hook: decode branch b -> f32 [n x m_b] column-major
__global__ void k_decode_branch(const uint8_t* __restrict__ store, BranchDesc d, uint64_t n,
                                const float* __restrict__ mu, const float* __restrict__ sd, int standardized,
                                float* __restrict__ out) {
    uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (idx >= n * d.m) return;
    uint64_t i = idx % n, j = idx / n;
    uint64_t t = i / kTileRows;
    uint32_t r = (uint32_t)(i % kTileRows);
    uint8_t byte = store[d.tile_off + t * (uint64_t)(kTileQuads * d.m_pad4) + (r >> 2) * d.m_pad4 + j];
    float g = (float)((byte >> (2 * (r & 3))) & 3);
    if (standardized) g = __fdiv_rn(__fsub_rn(g, mu[d.col_off + j]), sd[d.col_off + j]);  // bed.rs:354
    out[idx] = g;
}

// Synthetic genotypes in the spirit of BedVM::random (io/bed.rs:136-188): maf_j ~ U(lo, hi),
// g_ij ~ Binomial(2, maf_j), written as PLINK codes.  Counter-based (Philox keyed by column and
// GLOBAL row), so the data do not depend on how rows are sharded over GPUs.
__global__ void k_random_payload(uint8_t* __restrict__ payload, uint64_t n, uint64_t row_offset, uint64_t m,
                                 uint64_t bpc, uint64_t seed, float maf_lo, float maf_hi) {
    uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (idx >= m * bpc) return;
    const uint64_t j = idx / bpc, q = idx % bpc;
    Philox pm(seed ^ 0x6a09e667f3bcc908ull, j, 0);
    uint32_t r[4];
    pm.next(r);
    const float maf = maf_lo + (maf_hi - maf_lo) * u01_half_open(r[0]);
    const float p2 = maf * maf, p1 = p2 + 2.f * maf * (1.f - maf);
    Philox pg(seed, j, (row_offset >> 2) + q);
    pg.next(r);
    uint8_t byte = 0;
    for (int k = 0; k < 4; ++k) {
        if (4 * q + k >= n) break;
        const float u = u01_half_open(r[k]);
        const uint8_t code = (u < p2) ? 0x0 : (u < p1 ? 0x2 : 0x3);   // value 2 -> 00, 1 -> 10, 0 -> 11 (bed.rs:16)
        byte |= (uint8_t)(code << (2 * k));
    }
    payload[idx] = byte;
}

__global__ void k_gather_stats(const float* __restrict__ means, const float* __restrict__ stds,
                               const uint64_t* __restrict__ col_ids, uint64_t total, float* __restrict__ mu,
                               float* __restrict__ sd) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= total) return;
    mu[i] = means[col_ids[i]];
    sd[i] = stds[col_ids[i]];
}

}  // namespace bann

using namespace bann;

extern "C" {

const char* bann_last_error(void) { return g_last_error.c_str(); }

int bann_cuda_available(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return n > 0 ? 1 : 0;
}

uint64_t bann_launch_count(int reset) {
    uint64_t v = g_launch_count;
    if (reset) g_launch_count = 0;
    return v;
}

int bann_ctx_create(int device, void* stream, int rank, int world, bann_ctx** out) {
    if (!out) BANN_FAIL("out is NULL");
    if (!bann_cuda_available()) BANN_FAIL("no CUDA device: libbann_b200 has no CPU fallback");
    if (world < 1 || rank < 0 || rank >= world) BANN_FAIL("bad rank/world");
    BANN_CUDA(cudaSetDevice(device));
    bann_ctx* c = new bann_ctx();
    c->device = device;
    c->rank = rank;
    c->world = world;
    if (stream) {
        c->stream = (cudaStream_t)stream;
        c->owns_stream = false;
    } else {
        BANN_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->owns_stream = true;
    }
    cudaDeviceProp prop;
    BANN_CUDA(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    c->cc_major = prop.major;
    *out = c;
    return 0;
}

void bann_ctx_destroy(bann_ctx* c) {
    if (!c) return;
    bann::xr_release(c);
    if (c->owns_stream) cudaStreamDestroy(c->stream);
    delete c;
}

int bann_ctx_sync(bann_ctx* c) {
    if (!c) BANN_FAIL("ctx is NULL");
    BANN_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

static int build_from_device_payload(bann_ctx* ctx, uint8_t* d_payload /* consumed */, uint64_t n, uint64_t n_total,
                                     uint64_t m, const float* col_means, const float* col_stds, int device_stats,
                                     uint64_t num_branches, const uint64_t* branch_offsets, const uint64_t* col_ids,
                                     bann_genotypes** out) {
    uint64_t total_cols = branch_offsets[num_branches];
    bann_genotypes* g = new bann_genotypes();
    g->ctx = ctx;
    g->n = n;
    g->n_total = n_total ? n_total : n;
    g->m = m;
    g->num_branches = num_branches;
    g->ntiles = (uint32_t)((n + kTileRows - 1) / kTileRows);
    g->total_cols = total_cols;
    uint64_t bpc = (n + 3) / 4;
    cudaStream_t st = ctx->stream;

    // branch geometry
    g->m_b.resize(num_branches);
    g->m_pad4.resize(num_branches);
    g->tile_off.resize(num_branches);
    g->col_off.resize(num_branches);
    uint64_t off = 0;
    uint32_t max_mp = 0;
    for (uint64_t b = 0; b < num_branches; ++b) {
        uint32_t mb = (uint32_t)(branch_offsets[b + 1] - branch_offsets[b]);
        uint32_t mp = 4u * (((mb + 3) / 4) | 1u);   // bytes per row-quad: whole words, ODD word count (bank-conflict free)
        g->m_b[b] = mb;
        g->m_pad4[b] = mp;
        g->tile_off[b] = off;
        g->col_off[b] = branch_offsets[b];
        off += (uint64_t)g->ntiles * kTileQuads * mp;
        off = (off + 15) & ~15ull;  // keep every branch 16-byte aligned
        if (mp > max_mp) max_mp = mp;
    }
    g->store_bytes = off;
    // tensor-core store geometry (all-or-nothing: every branch within the 512 markers of the K-blocked tensor-core kernel)
    g->nst = (uint32_t)((n + 255) / 256);
    g->tc_off.assign(num_branches, 0);
    bool tc_ok = true;
    for (uint64_t b = 0; b < num_branches; ++b) tc_ok = tc_ok && g->m_b[b] <= 2048;   // kTcxMaxMarkers (k1_tcx.cuh)
    g->tc_bytes = 0;
    if (tc_ok) {
        for (uint64_t b = 0; b < num_branches; ++b) {
            g->tc_off[b] = g->tc_bytes;
            g->tc_bytes += (uint64_t)g->nst * ((g->m_b[b] + 7) / 8) * 128 * 4;
        }
    }
    g->packed_bytes = 0;
    for (uint64_t b = 0; b < num_branches; ++b) g->packed_bytes += (uint64_t)g->m_b[b] * bpc;

    uint64_t* d_cols = nullptr;
    BranchDesc* d_descs = nullptr;
    BANN_CUDA(cudaMalloc(&d_cols, total_cols * sizeof(uint64_t)));
    BANN_CUDA(cudaMemcpyAsync(d_cols, col_ids, total_cols * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    g->d_col_ids = d_cols;
    BANN_CUDA(cudaMalloc(&g->d_means, m * sizeof(float)));
    BANN_CUDA(cudaMalloc(&g->d_stds, m * sizeof(float)));
    BANN_CUDA(cudaMalloc(&g->d_mu, total_cols * sizeof(float)));
    BANN_CUDA(cudaMalloc(&g->d_sd, total_cols * sizeof(float)));
    BANN_CUDA(cudaMalloc(&g->d_store, g->store_bytes));
    BANN_CUDA(cudaMemsetAsync(g->d_store, 0, g->store_bytes, st));

    if (col_means) {
        BANN_CUDA(cudaMemcpyAsync(g->d_means, col_means, m * sizeof(float), cudaMemcpyHostToDevice, st));
        BANN_CUDA(cudaMemcpyAsync(g->d_stds, col_stds, m * sizeof(float), cudaMemcpyHostToDevice, st));
    } else if (device_stats) {
        k_col_stats<<<(unsigned)((m + 127) / 128), 128, 0, st>>>(d_payload, n, m, bpc, g->d_means, g->d_stds);
        BANN_LAUNCHED();
        BANN_CUDA(cudaGetLastError());
    } else {   // caller will provide global statistics through bann_genotypes_set_col_stats
        std::vector<float> zeros(m, 0.f), ones(m, 1.f);
        BANN_CUDA(cudaMemcpyAsync(g->d_means, zeros.data(), m * sizeof(float), cudaMemcpyHostToDevice, st));
        BANN_CUDA(cudaMemcpyAsync(g->d_stds, ones.data(), m * sizeof(float), cudaMemcpyHostToDevice, st));
        BANN_CUDA(cudaStreamSynchronize(st));
    }
    BANN_CUDA(cudaMalloc(&g->d_counts, 3 * m * sizeof(unsigned long long)));
    k_col_counts<<<(unsigned)((m + 127) / 128), 128, 0, st>>>(d_payload, n, m, bpc, g->d_counts);
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());

    // minimal descs for the tile builder
    std::vector<BranchDesc> descs(num_branches);
    for (uint64_t b = 0; b < num_branches; ++b) {
        memset(&descs[b], 0, sizeof(BranchDesc));
        descs[b].m = g->m_b[b];
        descs[b].m_pad4 = g->m_pad4[b];
        descs[b].tile_off = g->tile_off[b];
        descs[b].col_off = g->col_off[b];
        descs[b].tc_off = g->tc_off[b];
        descs[b].nc = (g->m_b[b] + 7) / 8;
    }
    BANN_CUDA(cudaMalloc(&d_descs, num_branches * sizeof(BranchDesc)));
    BANN_CUDA(cudaMemcpyAsync(d_descs, descs.data(), num_branches * sizeof(BranchDesc), cudaMemcpyHostToDevice, st));
    size_t smem = (size_t)max_mp * kTileQuads;
    if (smem > 200 * 1024) BANN_FAIL("branch with more than 6400 markers is not supported by the tile builder");
    if (smem > 48 * 1024) BANN_CUDA(cudaFuncSetAttribute(k_build_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    uint64_t nblocks = num_branches * (uint64_t)g->ntiles;
    if (nblocks > 0x7fffffffull) BANN_FAIL("too many tiles for one launch");
    k_build_tiles<<<(unsigned)nblocks, 256, smem, st>>>(d_payload, n, bpc, nullptr, d_descs, d_cols,
                                                      (uint32_t)num_branches, g->ntiles, g->d_store);
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    if (g->tc_bytes) {
        uint64_t nb_tc = num_branches * (uint64_t)g->nst;
        if (nb_tc > 0x7fffffffull) BANN_FAIL("too many super-tiles for one launch");
        BANN_CUDA(cudaMalloc(&g->d_store_tc, g->tc_bytes));
        k_build_tc<<<(unsigned)nb_tc, 128, 0, st>>>(d_payload, n, bpc, d_descs, d_cols, g->nst, g->d_store_tc);
        BANN_LAUNCHED();
        BANN_CUDA(cudaGetLastError());
    }
    k_gather_stats<<<(unsigned)((total_cols + 255) / 256), 256, 0, st>>>(g->d_means, g->d_stds, d_cols, total_cols,
                                                                         g->d_mu, g->d_sd);
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    BANN_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_payload);
    cudaFree(d_descs);
    *out = g;
    return 0;
}

static int check_csr(uint64_t m, uint64_t num_branches, const uint64_t* branch_offsets, const uint64_t* col_ids) {
    for (uint64_t b = 0; b < num_branches; ++b)
        if (branch_offsets[b + 1] <= branch_offsets[b]) BANN_FAIL("branch with no markers / offsets not increasing");
    for (uint64_t k = 0; k < branch_offsets[num_branches]; ++k)
        if (col_ids[k] >= m) BANN_FAIL("column id out of range");
    return 0;
}

int bann_genotypes_create(bann_ctx* ctx, const uint8_t* bed_payload, uint64_t n, uint64_t n_total, uint64_t m,
                          const float* col_means, const float* col_stds, uint64_t num_branches,
                          const uint64_t* branch_offsets, const uint64_t* col_ids, bann_genotypes** out) {
    if (!ctx || !bed_payload || !branch_offsets || !col_ids || !out) BANN_FAIL("NULL argument");
    if (n == 0 || m == 0 || num_branches == 0) BANN_FAIL("empty genotype store");
    if ((col_means == nullptr) != (col_stds == nullptr)) BANN_FAIL("col_means and col_stds must both be given or both NULL");
    if (!col_means && ctx->world > 1 && n_total != 0 && n_total != n)   // a replicated store (n == n_total, e.g. test data) is fine
        BANN_FAIL("column statistics must be global: pass col_means/col_stds when rows are sharded");
    BANN_CUDA(cudaSetDevice(ctx->device));
    BANN_CHECK(check_csr(m, num_branches, branch_offsets, col_ids));
    uint64_t bpc = (n + 3) / 4;
    uint8_t* d_payload = nullptr;
    BANN_CUDA(cudaMalloc(&d_payload, m * bpc));
    BANN_CUDA(cudaMemcpyAsync(d_payload, bed_payload, m * bpc, cudaMemcpyHostToDevice, ctx->stream));
    return build_from_device_payload(ctx, d_payload, n, n_total, m, col_means, col_stds, 1, num_branches, branch_offsets,
                                     col_ids, out);
}

int bann_genotypes_random(bann_ctx* ctx, uint64_t n, uint64_t row_offset, uint64_t n_total, uint64_t m, uint64_t seed,
                          float maf_lo, float maf_hi, uint64_t num_branches, const uint64_t* branch_offsets,
                          const uint64_t* col_ids, bann_genotypes** out) {
    if (!ctx || !branch_offsets || !col_ids || !out) BANN_FAIL("NULL argument");
    if (n == 0 || m == 0 || num_branches == 0) BANN_FAIL("empty genotype store");
    if (row_offset % 4 != 0) BANN_FAIL("row_offset must be a multiple of 4");
    BANN_CUDA(cudaSetDevice(ctx->device));
    BANN_CHECK(check_csr(m, num_branches, branch_offsets, col_ids));
    uint64_t bpc = (n + 3) / 4;
    uint8_t* d_payload = nullptr;
    BANN_CUDA(cudaMalloc(&d_payload, m * bpc));
    uint64_t total = m * bpc;
    k_random_payload<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(d_payload, n, row_offset, m, bpc, seed,
                                                                              maf_lo, maf_hi);
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    // statistics: exact sequential ones when this rank holds every row, else the caller combines
    // bann_genotypes_col_counts over ranks and calls bann_genotypes_set_col_stats
    return build_from_device_payload(ctx, d_payload, n, n_total, m, nullptr, nullptr, ctx->world == 1 ? 1 : 0, num_branches,
                                     branch_offsets, col_ids, out);
}

void bann_genotypes_destroy(bann_genotypes* g) {
    if (!g) return;
    cudaFree(g->d_store);
    cudaFree(g->d_store_tc);
    cudaFree(g->d_means);
    cudaFree(g->d_stds);
    cudaFree(g->d_mu);
    cudaFree(g->d_sd);
    cudaFree(g->d_col_ids);
    cudaFree(g->d_counts);
    delete g;
}

int bann_genotypes_col_stats(bann_genotypes* g, float* col_means, float* col_stds) {
    if (!g || !col_means || !col_stds) BANN_FAIL("NULL argument");
    BANN_CUDA(cudaMemcpyAsync(col_means, g->d_means, g->m * sizeof(float), cudaMemcpyDeviceToHost, g->ctx->stream));
    BANN_CUDA(cudaMemcpyAsync(col_stds, g->d_stds, g->m * sizeof(float), cudaMemcpyDeviceToHost, g->ctx->stream));
    BANN_CUDA(cudaStreamSynchronize(g->ctx->stream));
    return 0;
}

int bann_genotypes_col_counts(bann_genotypes* g, uint64_t* out) {
    if (!g || !out) BANN_FAIL("NULL argument");
    BANN_CUDA(cudaMemcpyAsync(out, g->d_counts, 3 * g->m * sizeof(uint64_t), cudaMemcpyDeviceToHost, g->ctx->stream));
    BANN_CUDA(cudaStreamSynchronize(g->ctx->stream));
    return 0;
}

int bann_genotypes_set_col_stats(bann_genotypes* g, const float* col_means, const float* col_stds) {
    if (!g || !col_means || !col_stds) BANN_FAIL("NULL argument");
    cudaStream_t st = g->ctx->stream;
    BANN_CUDA(cudaMemcpyAsync(g->d_means, col_means, g->m * sizeof(float), cudaMemcpyHostToDevice, st));
    BANN_CUDA(cudaMemcpyAsync(g->d_stds, col_stds, g->m * sizeof(float), cudaMemcpyHostToDevice, st));
    k_gather_stats<<<(unsigned)((g->total_cols + 255) / 256), 256, 0, st>>>(g->d_means, g->d_stds, g->d_col_ids,
                                                                            g->total_cols, g->d_mu, g->d_sd);
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    BANN_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int bann_genotypes_decode_branch(bann_genotypes* g, uint64_t b, int standardized, float* out) {
    if (!g || !out) BANN_FAIL("NULL argument");
    if (b >= g->num_branches) BANN_FAIL("branch index out of range");
    BranchDesc d;
    memset(&d, 0, sizeof(d));
    d.m = g->m_b[b];
    d.m_pad4 = g->m_pad4[b];
    d.tile_off = g->tile_off[b];
    d.col_off = g->col_off[b];
    if (!g->d_store) BANN_FAIL("the byte-tile store was released (bann_genotypes_release_byte_store): use bann_genotypes_decode_branch_tc");
    uint64_t total = g->n * d.m;
    float* dout = nullptr;
    BANN_CUDA(cudaMalloc(&dout, total * sizeof(float)));
    k_decode_branch<<<(unsigned)((total + 255) / 256), 256, 0, g->ctx->stream>>>(g->d_store, d, g->n, g->d_mu, g->d_sd,
                                                                                 standardized, dout);
    BANN_LAUNCHED();
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, total * sizeof(float), cudaMemcpyDeviceToHost, g->ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g->ctx->stream);
    cudaFree(dout);
    BANN_CUDA(e);
    return 0;
}

int bann_genotypes_decode_branch_tc(bann_genotypes* g, uint64_t b, int standardized, float* out) {
    if (!g || !out) BANN_FAIL("NULL argument");
    if (b >= g->num_branches) BANN_FAIL("branch index out of range");
    if (!g->d_store_tc) BANN_FAIL("no tensor-core store (a branch has more than 512 markers)");
    BranchDesc d;
    memset(&d, 0, sizeof(d));
    d.m = g->m_b[b];
    d.col_off = g->col_off[b];
    d.tc_off = g->tc_off[b];
    d.nc = (d.m + 7) / 8;
    uint64_t total = g->n * d.m;
    float* dout = nullptr;
    BANN_CUDA(cudaMalloc(&dout, total * sizeof(float)));
    k_decode_branch_tc<<<(unsigned)((total + 255) / 256), 256, 0, g->ctx->stream>>>(g->d_store_tc, d, g->n, g->d_mu, g->d_sd,
                                                                                    standardized, dout);
    BANN_LAUNCHED();
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, total * sizeof(float), cudaMemcpyDeviceToHost, g->ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g->ctx->stream);
    cudaFree(dout);
    BANN_CUDA(e);
    return 0;
}

int bann_genotypes_has_tc_store(bann_genotypes* g) { return g && g->d_store_tc ? 1 : 0; }
int bann_genotypes_has_byte_store(bann_genotypes* g) { return g && g->d_store ? 1 : 0; }

// The byte-tile store is read by the FFMA / shape-agnostic kernels, the probe kernels (activations, effect sizes) and
// bann_genotypes_decode_branch only; where the tensor-core store exists a caller that runs the tensor-core kernels can give its
// memory back (13 GB of 27 at BASELINE configs[2]).  Anything that still needs it afterwards fails with a message.
int bann_genotypes_release_byte_store(bann_genotypes* g) {
    if (!g) BANN_FAIL("NULL argument");
    if (!g->d_store_tc) BANN_FAIL("no tensor-core store (a branch has more than 2048 markers): the byte-tile store is the only copy of the genotypes");
    BANN_CUDA(cudaStreamSynchronize(g->ctx->stream));
    BANN_CUDA(cudaFree(g->d_store));
    g->d_store = nullptr;
    return 0;
}

}  // extern "C"

// Hardware probe for the tensor-core building blocks of K1 (run on a B200 through gpurun):
//   * SWIZZLE_NONE operand layouts / descriptor fields (K-major and MN-major views of one image),
//   * accumulator lane mapping for M = 128 and M = 64,
//   * f16 SUBNORMAL A operands (a 2-bit genotype code masked in place is g * 4^p * 2^-24) against
//     bf16 B operands in one kind::f16 instruction,
// each checked against a host computation in double precision.  Not part of the library.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../rs-bann_b200/csrc/umma.cuh"

using namespace bann::umma;

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e = (x);                                                                   \
        if (e != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);     \
            exit(2);                                                                           \
        }                                                                                      \
    } while (0)

struct ProbeArgs {
    const uint8_t* a_img; uint32_t a_bytes;
    const uint8_t* b_img; uint32_t b_bytes;
    uint32_t nsteps;
    uint32_t a_off, a_step, a_lbo, a_sbo;
    uint32_t b_off, b_step, b_lbo, b_sbo;
    uint32_t idesc;
    uint32_t ncol;      // columns to dump (multiple of 16)
    float* out;         // [128 lanes][ncol]
};

__global__ void __launch_bounds__(128) k_probe(ProbeArgs p) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base;
    uint8_t* sa = sm;
    uint8_t* sb = sm + ((p.a_bytes + 127) & ~127u);
    for (uint32_t i = threadIdx.x; i < p.a_bytes / 4; i += 128) ((uint32_t*)sa)[i] = ((const uint32_t*)p.a_img)[i];
    for (uint32_t i = threadIdx.x; i < p.b_bytes / 4; i += 128) ((uint32_t*)sb)[i] = ((const uint32_t*)p.b_img)[i];
    if (threadIdx.x == 0) { mbar_init(&mbar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tmem_base, 64);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t td = tmem_base;
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < p.nsteps; ++s) {
            const uint64_t ad = make_desc(smem_u32(sa) + p.a_off + s * p.a_step, p.a_lbo, p.a_sbo);
            const uint64_t bd = make_desc(smem_u32(sb) + p.b_off + s * p.b_step, p.b_lbo, p.b_sbo);
            mma_f16(td, ad, bd, p.idesc, s > 0);
        }
        commit(&mbar);
    }
    mbar_wait(&mbar, 0);
    fence_after_sync();
    const uint32_t warp = threadIdx.x >> 5;
    for (uint32_t c = 0; c < p.ncol; c += 16) {
        float v[16];
        tmem_ld16(td + ((warp * 32u) << 16) + c, v);
        for (int i = 0; i < 16; ++i) p.out[threadIdx.x * p.ncol + c + i] = v[i];
    }
    fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(td, 64);
}

// ---------------------------------------------------------------- host helpers
static float f16_to_float(uint16_t h) {
    const int s = h >> 15, e = (h >> 10) & 31, m = h & 1023;
    double v;
    if (e == 0) v = ldexp((double)m, -24);
    else if (e == 31) v = m ? NAN : INFINITY;
    else v = ldexp(1.0 + m / 1024.0, e - 15);
    return (float)(s ? -v : v);
}
static float bf16_to_float(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static uint16_t float_to_bf16(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static uint16_t float_to_f16(float f) {   // exactly representable values only (test data)
    if (f == 0.f) return 0;
    int e;
    const double m = frexp(fabs((double)f), &e);   // |f| = m * 2^e, m in [0.5, 1)
    const int E = e - 1 + 15;
    double mant = E >= 1 ? (m * 2.0 - 1.0) * 1024.0 : ldexp(fabs((double)f), 24);
    if (mant != floor(mant) || E > 30 || mant > 1023.0) { printf("value %g not exact in f16\n", f); exit(3); }
    const uint16_t h = (uint16_t)(((E >= 1 ? E : 0) << 10) | (int)mant);
    return (uint16_t)(h | (f < 0 ? 0x8000 : 0));
}
static uint32_t lcg = 12345u;
static uint32_t rnd() { lcg = lcg * 1664525u + 1013904223u; return lcg >> 8; }

// image with chunk (r, c) at c * cstride + r * 16; element e of the chunk at + 2 * e
static void put(std::vector<uint8_t>& img, uint32_t cstride, uint32_t r, uint32_t col, uint16_t bits) {
    const uint32_t c = col / 8, e = col % 8;
    const size_t off = (size_t)c * cstride + (size_t)r * 16 + 2 * e;
    if (off + 2 > img.size()) img.resize(off + 2, 0);
    memcpy(&img[off], &bits, 2);
}

struct Result { std::vector<float> out; };

static Result run(const std::vector<uint8_t>& a, const std::vector<uint8_t>& b, ProbeArgs p, size_t smem_min) {
    uint8_t *da, *db;
    float* dout;
    p.a_bytes = (uint32_t)((a.size() + 15) & ~15u);
    p.b_bytes = (uint32_t)((b.size() + 15) & ~15u);
    std::vector<uint8_t> ap(a), bp(b);
    ap.resize(p.a_bytes, 0);
    bp.resize(p.b_bytes, 0);
    CK(cudaMalloc(&da, p.a_bytes)); CK(cudaMalloc(&db, p.b_bytes));
    CK(cudaMemcpy(da, ap.data(), p.a_bytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, bp.data(), p.b_bytes, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dout, 128 * p.ncol * sizeof(float)));
    CK(cudaMemset(dout, 0xFF, 128 * p.ncol * sizeof(float)));
    p.a_img = da; p.b_img = db; p.out = dout;
    size_t smem = ((p.a_bytes + 127) & ~127u) + p.b_bytes + 256;
    if (smem < smem_min) smem = smem_min;
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_probe<<<1, 128, smem>>>(p);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    Result r;
    r.out.resize(128 * p.ncol);
    CK(cudaMemcpy(r.out.data(), dout, r.out.size() * 4, cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db); cudaFree(dout);
    return r;
}

// compare D[row][n] against expected; lane(row) mapping given
static int check(const char* name, const Result& r, uint32_t ncol, int M, int N, const std::vector<double>& expect,
                 int lane_of_row(int)) {
    double maxerr = 0, maxref = 0;
    int bad = 0;
    for (int i = 0; i < M; ++i)
        for (int n = 0; n < N; ++n) {
            const double e = expect[(size_t)i * N + n];
            const double g = r.out[(size_t)lane_of_row(i) * ncol + n];
            const double err = fabs(e - g);
            if (!(err <= 1e-6 * fabs(e) + 1e-30)) ++bad;
            if (err > maxerr || err != err) maxerr = err;
            if (fabs(e) > maxref) maxref = fabs(e);
        }
    printf("%-46s M=%3d N=%2d  max|err|=%.3e  max|ref|=%.3e  mismatches=%d  -> %s\n", name, M, N, maxerr, maxref, bad,
           bad == 0 ? "OK" : "FAIL");
    if (bad) {
        printf("   first rows (lane: got | expect):\n");
        for (int i = 0; i < 4; ++i) {
            printf("   row %d lane %d:", i, lane_of_row(i));
            for (int n = 0; n < 4; ++n) printf(" %.6e|%.6e", r.out[(size_t)lane_of_row(i) * ncol + n], expect[(size_t)i * N + n]);
            printf("\n");
        }
    }
    return bad;
}
static int lane128(int i) { return i; }
static int lane64(int i) { return (i % 16) + 32 * (i / 16); }

int main(int argc, char** argv) {
    const int mixed = argc > 1 && atoi(argv[1]) == 1;   // 1: also try f16 A x bf16 B (illegal instruction on sm_100a)
    int dev = 0;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    printf("device: %s cc %d.%d\n", prop.name, prop.major, prop.minor);
    int fails = 0;
    const int R = 128, K = 64, N = 16;
    // ---- operands: modes 0 = normal f16 A (small integers), 1 = subnormal f16 A (masked 2-bit codes), 2 = subnormal bf16 A
    for (int mode = 0; mode < 3; ++mode) {
        for (int bfmt = 0; bfmt < 2; ++bfmt) {   // B: 0 = f16, 1 = bf16
            if (mode == 2 && bfmt == 0) continue;
            if (mode != 2 && bfmt == 1 && !mixed) continue;
            std::vector<uint8_t> A, B, Dl;
            std::vector<double> av((size_t)R * K), bv((size_t)N * K), dv((size_t)R * N);
            const uint32_t a_cs = R * 16;     // chunk stride of the A image (bytes)
            for (int r = 0; r < R; ++r)
                for (int k = 0; k < K; ++k) {
                    const uint32_t g = rnd() % 3;
                    uint16_t bits;
                    if (mode == 0) bits = float_to_f16((float)g);
                    else if (mode == 1) bits = (uint16_t)(g << (2 * (k % 6)));        // f16 subnormal / first binades: g * 4^p * 2^-24
                    else bits = (uint16_t)(g << (2 * (k % 4)));                        // bf16: g * 4^p * 2^-133
                    put(A, a_cs, r, k, bits);
                    av[(size_t)r * K + k] = mode == 2 ? (double)g * ldexp(1.0, 2 * (k % 4) - 133) : (double)f16_to_float(bits);
                }
            // forward B: K-major [n][k]: chunk (n, kc) at kc * (N*16) + n * 16
            const double bscale = mode == 2 ? ldexp(1.0, 100) : (mode == 1 ? (bfmt ? ldexp(1.0, 24) : 256.0) : 1.0);
            for (int n = 0; n < N; ++n)
                for (int k = 0; k < K; ++k) {
                    const float w = (float)(((int)(rnd() % 255) - 127) / 32.0 * bscale);
                    const uint16_t bits = bfmt ? float_to_bf16(w) : float_to_f16(w);
                    put(B, N * 16, n, k, bits);
                    bv[(size_t)n * K + k] = bfmt ? bf16_to_float(bits) : f16_to_float(bits);
                }
            // backward B (delta): MN-major [row k'][n]: chunk (row, nc) at nc * (R*16) + row * 16
            for (int r = 0; r < R; ++r)
                for (int n = 0; n < N; ++n) {
                    const float w = (float)(((int)(rnd() % 255) - 127) / 64.0 * bscale);
                    const uint16_t bits = bfmt ? float_to_bf16(w) : float_to_f16(w);
                    put(Dl, R * 16, r, n, bits);
                    dv[(size_t)r * N + n] = bfmt ? bf16_to_float(bits) : f16_to_float(bits);
                }
            const uint32_t afmt = mode == 2 ? FMT_BF16 : FMT_F16, bf = bfmt ? FMT_BF16 : FMT_F16;
            char name[128];
            // ---- forward: D[r][n] = sum_k A[r][k] B[n][k]; M = 128, A K-major, B K-major
            {
                std::vector<double> ex((size_t)R * N, 0.0);
                for (int r = 0; r < R; ++r)
                    for (int n = 0; n < N; ++n) {
                        double s = 0;
                        for (int k = 0; k < K; ++k) s += av[(size_t)r * K + k] * bv[(size_t)n * K + k];
                        ex[(size_t)r * N + n] = s;
                    }
                ProbeArgs p{};
                p.nsteps = K / 16;
                p.a_off = 0; p.a_step = 2 * a_cs; p.a_lbo = a_cs; p.a_sbo = 128;
                p.b_off = 0; p.b_step = 2 * N * 16; p.b_lbo = N * 16; p.b_sbo = 128;
                p.idesc = make_idesc(afmt, bf, 0, 0, 128, N);
                p.ncol = 16;
                Result r = run(A, B, p, 0);
                snprintf(name, sizeof name, "fwd  K-major  Amode=%d Bfmt=%s", mode, bfmt ? "bf16" : "f16");
                fails += check(name, r, 16, R, N, ex, lane128);
            }
            // ---- backward: D[j][n] = sum_r A[r][j] delta[r][n]; A^T MN-major (M = 64 markers), B MN-major, K = rows
            for (int M = 64; M <= 128; M += 64) {
                std::vector<double> ex((size_t)M * N, 0.0);
                for (int j = 0; j < M && j < K; ++j)
                    for (int n = 0; n < N; ++n) {
                        double s = 0;
                        for (int r = 0; r < R; ++r) s += av[(size_t)r * K + j] * dv[(size_t)r * N + n];
                        ex[(size_t)j * N + n] = s;
                    }
                ProbeArgs p{};
                p.nsteps = R / 16;
                p.a_off = 0; p.a_step = 16 * 16; p.a_lbo = 128; p.a_sbo = a_cs;
                p.b_off = 0; p.b_step = 16 * 16; p.b_lbo = 128; p.b_sbo = R * 16;
                p.idesc = make_idesc(afmt, bf, 1, 1, M, N);
                p.ncol = 16;
                // M = 128 reads 16 marker chunks: rows 64.. come from whatever follows the image (zero padded here)
                std::vector<uint8_t> A2(A);
                if (M == 128) A2.resize((size_t)16 * a_cs, 0);
                Result r = run(A2, Dl, p, 0);
                snprintf(name, sizeof name, "bwd  MN-major Amode=%d Bfmt=%s", mode, bfmt ? "bf16" : "f16");
                fails += check(name, r, 16, M == 128 ? 64 : M, N, ex, M == 64 ? lane64 : lane128);
            }
        }
    }
    printf(fails ? "PROBE: %d mismatches\n" : "PROBE: all OK\n", fails);
    return 0;
}

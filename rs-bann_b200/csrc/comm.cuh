// Cross-rank sums over NVLink peer memory for the sequential-exact schedule ("in
// sequential-exact mode the all-reduce is latency-critical").
//
// Rows are sharded over ranks (one process per GPU); parameters, precisions and RNG keys are replicated.
// Every cross-row sum of a branch visit -- the raw gradient sums [gW | gb | rss] of each leapfrog step, the
// residual statistics after the transition -- has to be summed over ranks before the (replicated) update
// that consumes it.  These vectors are tiny (P_b + 1 floats, 1.2 KB at config 3), so the exchange is pure
// latency: it is FUSED into the kernel that produces the value (the fixed-order chunk reduction KR, the
// residual reductions) instead of being a separate collective launch.
//
// Protocol (one-shot, push, "low latency" words): every rank owns an inbox [2 parities][world sources][cap]
// of 8-byte words {payload, epoch}.  The thread that holds element `slot` stores {value, epoch} into its own
// column of EVERY peer's inbox (one 8-byte store per peer: delivered whole, so the flag travels with the
// data and no fence is needed), then polls its own inbox until the word of every source carries the epoch,
// and adds the values in RANK ORDER.  All ranks therefore compute bit-identical sums and take identical
// accept / reject decisions.  Parities alternate with the epoch: a peer can only write epoch e + 2 into a
// slot after it has completed epoch e + 1, which needed this rank's e + 1 contribution, which this rank
// sent after it had consumed epoch e (stream order) -- two parities suffice.  That argument needs EVERY epoch
// the host hands out (xr_next) to be exchanged by every rank: a kernel that received an XrComm must run its
// exchange unconditionally (k_reduce_partials does so for early-rejected branches too, with values nobody uses).
// A rank that never shows up trips a wall-clock limit: the error flag is raised and the kernel returns.
#pragma once

#include "common.cuh"

namespace bann {

constexpr int kXrMaxWorld = 8;
constexpr uint32_t kXrCap = 32768;                 // 8-byte words per (parity, source)
constexpr unsigned long long kXrTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

struct XrComm {
    uint2* inbox[kXrMaxWorld];   // inbox[r]: rank r's inbox (peer mapped for r != rank); NULL table = single rank
    uint32_t rank, world;
    uint32_t epoch;              // flag value of this exchange (>= 1, same on all ranks)
    int* error_flag;
};

// {payload, epoch} travel as ONE 64-bit word (a single st.b64 / ld.b64: 8-byte aligned accesses are single-copy atomic, so the
// flag can never be seen without its data; a .v2.u32 access is modelled as two scalar accesses and gives no such guarantee)
__device__ __forceinline__ void xr_store(uint2* p, uint32_t payload, uint32_t epoch) {
    const unsigned long long w = (unsigned long long)payload | ((unsigned long long)epoch << 32);
    asm volatile("st.relaxed.sys.global.b64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ uint2 xr_load(const uint2* p) {
    unsigned long long w;
    asm volatile("ld.relaxed.sys.global.b64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return make_uint2((uint32_t)w, (uint32_t)(w >> 32));
}
__device__ __forceinline__ unsigned long long xr_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// push this rank's 32-bit payload of element `slot` to every peer
__device__ __forceinline__ void xr_push(const XrComm& c, uint32_t slot, uint32_t payload) {
    const size_t mine = ((size_t)(c.epoch & 1u) * c.world + c.rank) * kXrCap + slot;
    for (uint32_t r = 0; r < c.world; ++r)
        if (r != c.rank) xr_store(c.inbox[r] + mine, payload, c.epoch);
}
// wait for source `src`'s payload of element `slot`
__device__ __forceinline__ uint32_t xr_pull(const XrComm& c, uint32_t slot, uint32_t src) {
    const uint2* p = c.inbox[c.rank] + ((size_t)(c.epoch & 1u) * c.world + src) * kXrCap + slot;
    uint2 v = xr_load(p);
    if (v.y == c.epoch) return v.x;
    if (*reinterpret_cast<volatile int*>(c.error_flag) == 2) return 0u;   // an earlier exchange already timed out: do not wait again
    const unsigned long long t0 = xr_now();
    while (true) {
        v = xr_load(p);
        if (v.y == c.epoch) return v.x;
        if (xr_now() - t0 > kXrTimeoutNs) {
            atomicExch(c.error_flag, 2);
            return 0u;
        }
    }
}

// sum over ranks of a float held by exactly one thread per rank (same `slot` on every rank), rank order
__device__ __forceinline__ float xr_sum(const XrComm& c, uint32_t slot, float v) {
    if (c.world <= 1) return v;
    xr_push(c, slot, __float_as_uint(v));
    float s = 0.f;
    for (uint32_t src = 0; src < c.world; ++src) {
        const float x = (src == c.rank) ? v : __uint_as_float(xr_pull(c, slot, src));
        s = (src == 0) ? x : s + x;
    }
    return s;
}
// the same for a double (two words: slots 2 * slot, 2 * slot + 1)
__device__ __forceinline__ double xr_sum(const XrComm& c, uint32_t slot, double v) {
    if (c.world <= 1) return v;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    xr_push(c, 2 * slot, (uint32_t)bits);
    xr_push(c, 2 * slot + 1, (uint32_t)(bits >> 32));
    double s = 0.0;
    for (uint32_t src = 0; src < c.world; ++src) {
        double x = v;
        if (src != c.rank) {
            const unsigned long long lo = xr_pull(c, 2 * slot, src), hi = xr_pull(c, 2 * slot + 1, src);
            x = __longlong_as_double((long long)(lo | (hi << 32)));
        }
        s = (src == 0) ? x : s + x;
    }
    return s;
}


// ------------------------------------------------------------------ bulk all-reduce / all-gather over peer memory (grouped schedule)
// The grouped schedule sums [gW | gb | rss] of EVERY listed branch over ranks once per leapfrog step: B x (P_b + 1) floats
// (11.7 MB at config 3) -- bandwidth, not latency, so the push protocol above (8 B per value and peer) is the wrong tool.
// Each rank owns an exchange region in its HBM, mapped by every peer (bann_net_comm_handle / _connect):
//     flagsA[8] flagsB[8] | in[2][cap] | out[2][cap]
// One all-reduce (epoch e, parity e & 1), two kernels on the caller's stream, no host involvement:
//   k_xg_publish: copies the local sums into in[parity]; the last block to finish fences at system scope and writes e into
//                 flagsA[rank] of every peer.
//   k_xg_reduce_scatter: waits for flagsA[src] >= e of every source, sums ITS slice (1 / world of the values) over the ranks'
//                 in[parity] in RANK ORDER (bit-identical everywhere), writes it to out[parity] and to the local result, then
//                 (last block) writes e into flagsB[rank] of every peer.
//   k_xg_all_gather: waits for flagsB, pulls the other slices from the owners' out[parity].
// Every rank moves 2 x (world - 1) / world x bytes over NVLink, like a ring / NVLS all-reduce, in three tiny launches.
// Buffer reuse: in[parity] / out[parity] of epoch e are overwritten at e + 2, after this rank completed epoch e + 1, which
// needed every peer's e + 1 flags, which the peers raised after finishing their reads of epoch e (stream order).
// A rank that never shows up trips the same 20 s wall-clock limit as above (error flag 2).
struct XgComm {
    uint8_t* region[kXrMaxWorld];   // region[r]: rank r's exchange region (peer mapped for r != rank)
    uint32_t rank, world;
    uint32_t epoch;
    uint64_t cap;                   // floats per in / out buffer
    int* error_flag;
    unsigned int* counter;          // local: blocks finished (self-resetting)
};
__device__ __forceinline__ uint32_t* xg_flags(uint8_t* region, int which) { return reinterpret_cast<uint32_t*>(region) + 8 * which; }
__device__ __forceinline__ float* xg_in(const XgComm& c, uint32_t r) {
    return reinterpret_cast<float*>(c.region[r] + 256) + (size_t)(c.epoch & 1u) * c.cap;
}
__device__ __forceinline__ float* xg_out(const XgComm& c, uint32_t r) {
    return reinterpret_cast<float*>(c.region[r] + 256) + (size_t)(2u + (c.epoch & 1u)) * c.cap;
}
__device__ __forceinline__ void xg_signal(const XgComm& c, int which) {     // one thread, after the data is written
    __threadfence_system();
    for (uint32_t r = 0; r < c.world; ++r) {
        uint32_t* f = xg_flags(c.region[r], which) + c.rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(c.epoch) : "memory");
    }
}
__device__ __forceinline__ bool xg_wait(const XgComm& c, int which) {       // one thread per block, then __syncthreads
    const uint32_t* f = xg_flags(c.region[c.rank], which);
    const unsigned long long t0 = xr_now();
    for (uint32_t src = 0; src < c.world; ++src) {
        while (true) {
            uint32_t v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f + src) : "memory");
            if ((int32_t)(v - c.epoch) >= 0) break;
            if (*reinterpret_cast<volatile int*>(c.error_flag) == 2) return false;
            if (xr_now() - t0 > kXrTimeoutNs) { atomicExch(c.error_flag, 2); return false; }
        }
    }
    return true;
}
__device__ __forceinline__ float4 xg_ld4(const float* p) {                   // peer data: never from a stale L1 line
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
// the last block of a grid to get here returns true (and resets the counter for the next launch)
__device__ __forceinline__ bool xg_last_block(unsigned int* counter) {
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(counter, 1u);
        last = done == gridDim.x - 1;
        if (last) *counter = 0;
    }
    __syncthreads();
    return last;
}
// count: floats, a multiple of 4; src / dst 16-byte aligned
static __global__ void __launch_bounds__(256) k_xg_publish(XgComm c, const float* __restrict__ src, uint64_t count) {
    float4* dst = reinterpret_cast<float4*>(xg_in(c, c.rank));
    const float4* s4 = reinterpret_cast<const float4*>(src);
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < count / 4; i += (uint64_t)gridDim.x * 256) dst[i] = s4[i];
    if (xg_last_block(c.counter) && threadIdx.x == 0) xg_signal(c, 0);
}
// slice of rank r: float4 indices [r * per4, min((r + 1) * per4, count / 4))
static __global__ void __launch_bounds__(256) k_xg_reduce_scatter(XgComm c, float* __restrict__ result, uint64_t count, uint64_t per4) {
    __shared__ bool ok;
    if (threadIdx.x == 0) ok = xg_wait(c, 0);
    __syncthreads();
    if (ok) {
        const uint64_t lo = c.rank * per4, hi = min((c.rank + 1) * per4, count / 4);
        float4* out = reinterpret_cast<float4*>(xg_out(c, c.rank));
        float4* res = reinterpret_cast<float4*>(result);
        for (uint64_t i = lo + (uint64_t)blockIdx.x * 256 + threadIdx.x; i < hi; i += (uint64_t)gridDim.x * 256) {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            for (uint32_t r = 0; r < c.world; ++r) {          // rank order: identical bits on every rank
                const float4 v = (r == c.rank) ? res[i] : xg_ld4(xg_in(c, r) + 4 * i);
                if (r == 0) s = v;
                else { s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
            }
            out[i] = s;
            res[i] = s;
        }
    }
    if (xg_last_block(c.counter) && threadIdx.x == 0) xg_signal(c, 1);
}
static __global__ void __launch_bounds__(256) k_xg_all_gather(XgComm c, float* __restrict__ result, uint64_t count, uint64_t per4) {
    __shared__ bool ok;
    if (threadIdx.x == 0) ok = xg_wait(c, 1);
    __syncthreads();
    if (!ok) return;
    float4* res = reinterpret_cast<float4*>(result);
    const uint64_t n4 = count / 4;
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * 256) {
        const uint32_t owner = (uint32_t)(i / per4);
        if (owner != c.rank) res[i] = xg_ld4(xg_out(c, owner) + 4 * i);
    }
}
// all-gather of slices the ranks hold locally (parameter upload: every rank copied only its own slice from the host):
// publish writes the slice into out[parity] and raises flagsB; k_xg_all_gather then fills in the rest
static __global__ void __launch_bounds__(256) k_xg_publish_slice(XgComm c, const float* __restrict__ src, uint64_t count, uint64_t per4) {
    const uint64_t lo = c.rank * per4, hi = min((c.rank + 1) * per4, count / 4);
    float4* out = reinterpret_cast<float4*>(xg_out(c, c.rank));
    const float4* s4 = reinterpret_cast<const float4*>(src);
    for (uint64_t i = lo + (uint64_t)blockIdx.x * 256 + threadIdx.x; i < hi; i += (uint64_t)gridDim.x * 256) out[i] = s4[i];
    if (xg_last_block(c.counter) && threadIdx.x == 0) xg_signal(c, 1);
}

}  // namespace bann

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

}  // namespace bann

// K1 on the 5th-generation tensor cores (tcgen05 + tensor memory), for branches with <= 64 markers.
//
// The two big contractions of a branch -- z0 = X W0 (forward) and S = X^T delta_0 (backward) -- run as
// tcgen05.mma kind::f16 with bf16 operands and f32 accumulators in tensor memory; everything in
// between (remaining layers, tanh, error, deltas, cross-row sums of the small layers) stays in FP32
// registers, one thread per individual, exactly as in k1_small.
//
// Genotypes never get converted arithmetically.  The tensor-core store (genotypes.cu, k_build_tc)
// keeps, per 256-row super-tile and 8-marker chunk, one 32-bit word per ROW PAIR (t, 128 + t) whose
// 2-bit fields sit where a single AND turns them into bf16 numbers: a genotype code g at bits
// [2p+1:2p] (p = 0..3) of a 16-bit half is the bf16 SUBNORMAL g * 4^p * 2^-133 (exact; checked on
// hardware by tests/probe/probe_umma.cu).  One LOP3 yields two operand elements; the per-marker power
// of two 4^p(j) is folded into the staged weights (forward) and undone on the accumulator (backward).
// The expanded tile [chunk][row] x 16 B is written once per super-tile and read by BOTH contractions:
// as the K-major A operand (rows x markers) and as the MN-major A operand (markers x rows).
//
// f32 weights and deltas enter as three bf16 pieces each (hi + mid + lo = the f32 value, 24 significand
// bits), side by side in the N dimension, so products are exact and only the f32 accumulation rounds.
// Pieces are pre-scaled by 2^100 (exact) so that subnormal * piece stays a normal f32.
//
// Per CTA (128 threads, one branch, a range of super-tiles), software pipelined over two operand buffers:
//   wait fwd(i) -> tcgen05.ld z0 -> tail part 1 (layers, tanh, error; both rows of a thread as packed f32x2)
//   -> wait bwd(i-1) -> expand(i+1) -> issue MMA fwd(i+1) (2 x M128 N16 K16*ks)  [runs under part 2]
//   -> tail part 2 (deltas, cross-row sums) -> delta pieces to shared memory
//   -> issue MMA bwd(i) (M64 N16 K16 x 16, accumulating in tensor memory over ALL super-tiles of the CTA)
//   [runs under part 1 of the next super-tile] -> one read of the first-layer gradient at the end.
#pragma once
#include <cuda_bf16.h>

#include "k1_small.cuh"
#include "umma.cuh"

// Round-1 history of this kernel (cfg3s K1 ms): separate word loads 1.99 -> batched word loads in expand 1.86 -> deferred MMA issue
// 1.68 -> no proxy fence after expand 1.645.  Tried, no gain: forward issue after the delta pieces, backward issue later in part 1.

namespace bann {

constexpr int kTcMaxMarkers = 64;     // one M = 64 backward accumulator, four K = 16 forward steps
constexpr int kTcRows = 256;          // rows per super-tile: thread t owns rows t and 128 + t
constexpr uint32_t kTcChunkStride = kTcRows * 16;   // bytes between 8-marker chunks of the expanded tile

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ float pow2f(int e) { return __int_as_float((127 + e) << 23); }

// ---- packed FP32 pairs (Blackwell FFMA2 / FMUL2 / FADD2): lane .x = row t, lane .y = row 128 + t
struct f2 { unsigned long long v; };
__device__ __forceinline__ f2 mk2(float x, float y) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ f2 dup2(float x) { return mk2(x, x); }
__device__ __forceinline__ float lo2(f2 a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); (void)y; return x; }
__device__ __forceinline__ float hi2(f2 a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); (void)x; return y; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 ld2(const float2* p) { f2 r; r.v = *reinterpret_cast<const unsigned long long*>(p); return r; }

// tanh of a pair: 1 - 2 / (2^(2x log2 e) + 1); saturates correctly (ex2 -> inf / 0), absolute error ~1.5e-7
__device__ __forceinline__ f2 tanh2(f2 x) {
    const f2 t = mul2(x, dup2(2.8853900817779268f));
    float e0, e1, r0, r1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(lo2(t)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(hi2(t)));
    const f2 s = add2(mk2(e0, e1), dup2(1.f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(lo2(s)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(hi2(s)));
    return fma2(mk2(r0, r1), dup2(-2.f), dup2(1.f));
}
// 1 - a^2 (activation_functions.rs:33-45 for tanh, evaluated from the activation)
__device__ __forceinline__ f2 dtanh2(f2 a) { return fma2(a, mul2(a, dup2(-1.f)), dup2(1.f)); }

// ---- activations of the tensor-core tails (activation_functions.rs:23-45), on packed pairs.
// The pre-activation arrives PRE-SCALED by act_prescale<ACT>() (folded into the staged weights and biases, so the
// scaling costs no instruction): tanh wants 2 x log2(e), everything else the plain value.
template <int ACT> __host__ __device__ constexpr float act_prescale() { return ACT == BANN_TANH ? 2.8853900817779268f : 1.f; }
// activation of a pre-scaled pair; `aux` = what the derivative needs besides the activation (SiLU: the sigmoid)
template <int ACT>
__device__ __forceinline__ f2 act2(f2 t, f2& aux) {
    if constexpr (ACT == BANN_TANH) {          // 1 - 2 / (2^t + 1), t = 2 x log2(e); saturates correctly (ex2 -> inf / 0)
        float e0, e1, r0, r1;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(lo2(t)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(hi2(t)));
        const f2 s = add2(mk2(e0, e1), dup2(1.f));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(lo2(s)));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(hi2(s)));
        aux = t;
        return fma2(mk2(r0, r1), dup2(-2.f), dup2(1.f));
    } else if constexpr (ACT == BANN_RELU) {
        aux = t;
        return mk2(fmaxf(lo2(t), 0.f), fmaxf(hi2(t), 0.f));
    } else if constexpr (ACT == BANN_LEAKY_RELU) {   // max(x, 0.01 x): x for x > 0, 0.01 x for x < 0, 0 at 0
        const f2 s = mul2(t, dup2(0.01f));
        aux = t;
        return mk2(fmaxf(lo2(t), lo2(s)), fmaxf(hi2(t), hi2(s)));
    } else if constexpr (ACT == BANN_SILU) {         // x * sigma(x), sigma = 1 / (1 + 2^(-x log2 e))
        const f2 u = mul2(t, dup2(-1.4426950408889634f));
        float e0, e1, r0, r1;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(lo2(u)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(hi2(u)));
        const f2 s = add2(mk2(e0, e1), dup2(1.f));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(lo2(s)));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(hi2(s)));
        aux = mk2(r0, r1);
        // rcp(inf) = 0 and x -> -inf: 0 * (-inf) would be NaN; the reference's x * (1 / (1 + exp(-x))) has the same hazard only at -inf itself
        return mul2(t, aux);
    } else {
        aux = t;
        return t;
    }
}
// MINUS the derivative of the activation at the pre-activation, from the activation (and aux): the backward pass carries
// alternating signs so that tanh's a^2 - 1 needs one instruction instead of two
template <int ACT>
__device__ __forceinline__ f2 neg_dact2(f2 a, f2 aux) {
    if constexpr (ACT == BANN_TANH) return fma2(a, a, dup2(-1.f));
    else if constexpr (ACT == BANN_RELU) return mk2(lo2(a) > 0.f ? -1.f : 0.f, hi2(a) > 0.f ? -1.f : 0.f);     // a > 0 <=> x > 0
    else if constexpr (ACT == BANN_LEAKY_RELU)
        return mk2(lo2(a) > 0.f ? -1.f : (lo2(a) < 0.f ? -0.01f : 0.f), hi2(a) > 0.f ? -1.f : (hi2(a) < 0.f ? -0.01f : 0.f));
    else if constexpr (ACT == BANN_SILU) return fma2(aux, a, mul2(add2(a, aux), dup2(-1.f)));   // -(f + s (1 - f)), f = x s = a
    else return dup2(-1.f);
}

template <int H, int S, int D>
struct TcShape {
    using T = TailShape<H, S, D>;
    static constexpr int W0 = T::W0;
    static constexpr int NN = 16;                       // accumulator columns: 3 pieces x W0 units, padded
    static_assert(3 * W0 <= NN, "first-layer width too large for one N = 16 accumulator");
    static constexpr int TMEM_COLS = 128;               // 2 x NN forward + 2 x NN backward accumulators, 2 x 32 columns: forward A operand
    static constexpr size_t SD = (NN / 8) * kTcChunkStride;   // delta pieces  [n-chunk][row] x 16 B
    static constexpr size_t SW = 8 * NN * 16;           // weight pieces [k-chunk][n] x 16 B
    static constexpr int NRED = 4 * (T::NTACC > 64 ? T::NTACC : 64);
    // Layout: [expanded genotypes: 2 buffers x ncb chunks][weight pieces][packed words of the next tile][delta pieces][misc].
    // The operand reads overrun a buffer on purpose: the last forward K-step of an odd ncb reads chunk ncb (times
    // zero weights; what lies there -- the other buffer, weight pieces, packed words -- is always finite) and the
    // M = 64 backward operand always reads 8 chunks (result rows >= m are never used), so ncb + 8 chunks must exist.
    static constexpr size_t MISC = (size_t)(2 * ((T::n_tail() + 3) & ~3) + 2 * T::W0P + NRED + 16) * 4 + 64;
    static size_t smem(uint32_t ncb) {
        const size_t used = (size_t)2 * ncb * kTcChunkStride + SW + (size_t)ncb * 512 + SD + MISC;
        const size_t need = (size_t)(ncb + 8) * kTcChunkStride;
        return (used > need ? used : need) + 128;
    }
};

// ---- the FP32 tail shared by the tensor-core kernels (k1_tc, k1_tcw): thread = row pair (t, 128 + t) as packed f32x2.
// part1: remaining layers, error, and every part of the backward pass that needs no other data than the thread's own rows:
//   all deltas are linear in the error, delta_l = e * rho_l with rho_{NLA-1} = phi'(z) (.) W_out and
//   rho_{l-1} = phi'(z_{l-1}) (.) (rho_l W_l^T)  (branch_sampler.rs:813-875 with the error factored out), so rho_l is computed right
//   behind the forward pass, while the weights of layer l are still in registers: one shared-memory read per weight and
//   super-tile instead of two (the kernels are shared-memory-wavefront bound, profiles/r2_k1_tc_ncu_summary.md).
//   Signs alternate -- sg_l = (-1)^(NLA-l) cA^(NLA-1-l) rho_l -- so that tanh's a^2 - 1 is one instruction and the pre-scaled
//   weights are used as they are; the factor is undone on the error, one multiplication per layer.
// part2: delta_0 = sg_0 * (signed, unscaled error) -> three bf16 pieces per unit by truncation (exact: 3 x 8 significand bits).
// recursive-halving warp reduction of CNT values per lane (TcTail::reduce_and_store<true>): after the steps 16, 8, 4, 2, 1 every lane
// holds halved_count(CNT, 5) fully reduced values
__host__ __device__ constexpr int halved_count(int cnt, int steps) { return steps == 0 ? cnt : halved_count((cnt + 1) / 2, steps - 1); }
template <int CNT, int O, int N>
__device__ __forceinline__ void halve_step(float (&v)[N], uint32_t lane) {
    if constexpr (O >= 1) {
        constexpr int HALF = (CNT + 1) / 2;
        const bool up = (lane & (uint32_t)O) != 0;
#pragma unroll
        for (int i = 0; i < HALF; ++i) {
            const float hi = 2 * i + 1 < CNT ? v[2 * i + 1] : 0.f;
            const float send = up ? v[2 * i] : hi;
            const float keep = up ? hi : v[2 * i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, O);
        }
        halve_step<HALF, O / 2>(v, lane);
    }
}

// SACC: the cross-row sums as ONE float per value (row t and row 128 + t summed as they arrive: two FFMA instead of one FFMA2,
// the same FMA-pipe cycles, one more issue slot) instead of a packed pair -- half the accumulator registers (41 instead of 82 for
// [5,5,1]), which is what lets k1_tc5's fifth (issuing) warp fit three CTAs per SM (k1_tc5.cuh).
template <bool SACC> struct TcAccElem { using type = f2; };
template <> struct TcAccElem<true> { using type = float; };
template <bool SACC> __device__ __forceinline__ typename TcAccElem<SACC>::type acc_zero() {
    if constexpr (SACC) return 0.f; else return dup2(0.f);
}
__device__ __forceinline__ void acc_fma(f2& d, f2 a, f2 b) { d = fma2(a, b, d); }
__device__ __forceinline__ void acc_fma(float& d, f2 a, f2 b) { d = fmaf(hi2(a), hi2(b), fmaf(lo2(a), lo2(b), d)); }
__device__ __forceinline__ void acc_add(f2& d, f2 a) { d = add2(d, a); }
__device__ __forceinline__ void acc_add(float& d, f2 a) { d = (d + lo2(a)) + hi2(a); }
__device__ __forceinline__ float acc_total(f2 v) { return lo2(v) + hi2(v); }
__device__ __forceinline__ float acc_total(float v) { return v; }

template <int H, int S, int D, int ACT, bool SACC = false>
struct TcTail {
    using T = TailShape<H, S, D>;
    static constexpr int NLA = T::NLA, W0 = T::W0, MW = T::MW, NN = 16;
    static constexpr int NL1 = NLA > 1 ? NLA - 1 : 1;
    static constexpr float cA = act_prescale<ACT>();
    using AE = typename TcAccElem<SACC>::type;
    struct Acc {                          // cross-row sums, persistent over the super-tiles of a CTA
        AE gb0[W0], gWo[S], rss, gWt[NL1][MW][MW], gbt[NL1][MW];
        __device__ __forceinline__ void clear() {
            const AE z = acc_zero<SACC>();
            rss = z;
#pragma unroll
            for (int c = 0; c < W0; ++c) gb0[c] = z;
#pragma unroll
            for (int c = 0; c < S; ++c) gWo[c] = z;
#pragma unroll
            for (int l = 0; l < NL1; ++l)
#pragma unroll
                for (int i = 0; i < MW; ++i) {
                    gbt[l][i] = z;
#pragma unroll
                    for (int c = 0; c < MW; ++c) gWt[l][i][c] = z;
                }
        }
    };
    // wp: tail parameters (weights / biases feeding an activated layer >= 1 pre-scaled by cA), b0p: cA * mean-folded first-layer bias
    __device__ __forceinline__ static void part1(const float* accA, const float* accB, const float* wp, const float* b0p, f2& tg,
                                                 bool add_pred_to_target, f2 valid, bool bwd, Acc& A, f2& yh, f2 (&sg0)[W0], f2& ef0) {
        const f2 zero2 = dup2(0.f);
        f2 act[NLA][MW], aux[NLA][MW];
#pragma unroll
        for (int c = 0; c < W0; ++c) {
            const f2 z = mk2(accA[c] + (accA[W0 + c] + accA[2 * W0 + c]), accB[c] + (accB[W0 + c] + accB[2 * W0 + c]));
            act[0][c] = act2<ACT>(fma2(z, dup2(8589934592.f /* 2^33 */ * cA), dup2(b0p[c])), aux[0][c]);
        }
#pragma unroll
        for (int l = 1; l < NLA; ++l) {
#pragma unroll
            for (int c = 0; c < MW; ++c) {
                if (c < T::width(l)) {
                    f2 zz = dup2(wp[T::b_off(l) + c]);
#pragma unroll
                    for (int i = 0; i < MW; ++i)
                        if (i < T::in_w(l)) zz = fma2(act[l - 1][i], dup2(wp[T::w_off(l) + c * T::in_w(l) + i]), zz);
                    act[l][c] = act2<ACT>(zz, aux[l][c]);
                }
            }
        }
        yh = zero2;
#pragma unroll
        for (int i = 0; i < S; ++i) yh = fma2(act[NLA - 1][i], dup2(wp[T::w_off(NLA) + i]), yh);
        if (add_pred_to_target) tg = add2(tg, yh);                         // net.rs:280
        const f2 e = mul2(fma2(tg, dup2(-1.f), yh), valid);                // branch_sampler.rs:821
        ef0 = zero2;
#pragma unroll
        for (int c = 0; c < W0; ++c) sg0[c] = zero2;
        if (!bwd) return;
        f2 sg[MW];
#pragma unroll
        for (int i = 0; i < S; ++i) sg[i] = mul2(neg_dact2<ACT>(act[NLA - 1][i], aux[NLA - 1][i]), dup2(wp[T::w_off(NLA) + i]));
        acc_fma(A.rss, e, e);
#pragma unroll
        for (int i = 0; i < S; ++i) acc_fma(A.gWo[i], act[NLA - 1][i], e);
        float fac = -1.f;                  // (-1)^(NLA-l) cA^-(NLA-1-l) at l = NLA - 1
#pragma unroll
        for (int l = NLA - 1; l >= 1; --l) {
            f2 nd[MW];                     // sg of the layer below, from the weights the forward pass of layer l just used
#pragma unroll
            for (int i = 0; i < MW; ++i) nd[i] = zero2;
#pragma unroll
            for (int c = 0; c < MW; ++c)
                if (c < T::width(l)) {
#pragma unroll
                    for (int i = 0; i < MW; ++i)
                        if (i < T::in_w(l)) nd[i] = fma2(sg[c], dup2(wp[T::w_off(l) + c * T::in_w(l) + i]), nd[i]);
                }
            const f2 ef = mul2(e, dup2(fac));        // cross-row sums of layer l
#pragma unroll
            for (int c = 0; c < MW; ++c)
                if (c < T::width(l)) {
                    const f2 dl = mul2(sg[c], ef);
                    acc_add(A.gbt[l - 1][c], dl);
#pragma unroll
                    for (int i = 0; i < MW; ++i)
                        if (i < T::in_w(l)) acc_fma(A.gWt[l - 1][i][c], act[l - 1][i], dl);
                }
#pragma unroll
            for (int i = 0; i < MW; ++i)
                if (i < T::in_w(l)) sg[i] = mul2(neg_dact2<ACT>(act[l - 1][i], aux[l - 1][i]), nd[i]);
            fac = -fac / cA;
        }
#pragma unroll
        for (int c = 0; c < W0; ++c) sg0[c] = sg[c];
        ef0 = mul2(e, dup2(fac));
    }
    // ---- part1 without the cross-row sums (k1_tc5): those need nothing but registers and nobody waits for them, so the caller
    // takes them later, in slices (`accumulate_slice`) placed by hand where a warp would otherwise only wait for one pipe --
    // ptxas keeps blocks of different kinds of work apart even inside one basic block, hand placement is what interleaves them.
    // Mid: what the sums need.  part1x calls between(c) behind unit c of the first layer (ex2 / rcp on the MUFU pipe) and
    // after() behind the whole layer; both run BEFORE any field of M other than act[0] is written, so the caller may keep the
    // previous super-tile's terms in the same M (and a copy of its act[0]).
    struct Mid { f2 act[NLA][MW], sgl[NL1][MW], e; };
    template <class Between, class After>
    __device__ __forceinline__ static void part1x(const float* accA, const float* accB, const float* wp, const float* b0p, f2& tg,
                                                  bool add_pred_to_target, f2 valid, bool bwd, Mid& M, f2& yh, f2 (&sg0)[W0], f2& ef0,
                                                  Between between, After after) {
        const f2 zero2 = dup2(0.f);
        f2 aux[NLA][MW];
#pragma unroll
        for (int c = 0; c < W0; ++c) {
            const f2 z = mk2(accA[c] + (accA[W0 + c] + accA[2 * W0 + c]), accB[c] + (accB[W0 + c] + accB[2 * W0 + c]));
            M.act[0][c] = act2<ACT>(fma2(z, dup2(8589934592.f /* 2^33 */ * cA), dup2(b0p[c])), aux[0][c]);
            between(c);
        }
        after();
#pragma unroll
        for (int l = 1; l < NLA; ++l) {
#pragma unroll
            for (int c = 0; c < MW; ++c) {
                if (c < T::width(l)) {
                    f2 zz = dup2(wp[T::b_off(l) + c]);
#pragma unroll
                    for (int i = 0; i < MW; ++i)
                        if (i < T::in_w(l)) zz = fma2(M.act[l - 1][i], dup2(wp[T::w_off(l) + c * T::in_w(l) + i]), zz);
                    M.act[l][c] = act2<ACT>(zz, aux[l][c]);
                }
            }
        }
        yh = zero2;
#pragma unroll
        for (int i = 0; i < S; ++i) yh = fma2(M.act[NLA - 1][i], dup2(wp[T::w_off(NLA) + i]), yh);
        if (add_pred_to_target) tg = add2(tg, yh);                         // net.rs:280
        M.e = mul2(fma2(tg, dup2(-1.f), yh), valid);                       // branch_sampler.rs:821
        ef0 = zero2;
#pragma unroll
        for (int c = 0; c < W0; ++c) sg0[c] = zero2;
        if (!bwd) return;
        f2 sg[MW];
#pragma unroll
        for (int i = 0; i < S; ++i) sg[i] = mul2(neg_dact2<ACT>(M.act[NLA - 1][i], aux[NLA - 1][i]), dup2(wp[T::w_off(NLA) + i]));
        float fac = -1.f;
#pragma unroll
        for (int l = NLA - 1; l >= 1; --l) {
            f2 nd[MW];
#pragma unroll
            for (int i = 0; i < MW; ++i) nd[i] = zero2;
#pragma unroll
            for (int c = 0; c < MW; ++c)
                if (c < T::width(l)) {
                    M.sgl[l - 1][c] = sg[c];
#pragma unroll
                    for (int i = 0; i < MW; ++i)
                        if (i < T::in_w(l)) nd[i] = fma2(sg[c], dup2(wp[T::w_off(l) + c * T::in_w(l) + i]), nd[i]);
                }
#pragma unroll
            for (int i = 0; i < MW; ++i)
                if (i < T::in_w(l)) sg[i] = mul2(neg_dact2<ACT>(M.act[l - 1][i], aux[l - 1][i]), nd[i]);
            fac = -fac / cA;
        }
#pragma unroll
        for (int c = 0; c < W0; ++c) sg0[c] = sg[c];
        ef0 = mul2(M.e, dup2(fac));
    }
    // the cross-row sums of the layers >= 1, the output layer and rss (same terms per accumulator as part1) in NSLICE slices:
    // slice 0 = rss + output layer, slice 1 + (NLA-1-l) MW + c = unit c of layer l.  a0: the first layer's activations (M.act[0],
    // or the caller's copy of them)
    static constexpr int NSLICE = 1 + (NLA - 1) * MW;
    __device__ __forceinline__ static void accumulate_slice(const Mid& M, const f2 (&a0)[MW], Acc& A, int k) {
        if (k == 0) {
            acc_fma(A.rss, M.e, M.e);
#pragma unroll
            for (int i = 0; i < S; ++i) acc_fma(A.gWo[i], NLA == 1 ? a0[i] : M.act[NLA - 1][i], M.e);
        }
        float fac = -1.f;
#pragma unroll
        for (int l = NLA - 1; l >= 1; --l) {
#pragma unroll
            for (int c = 0; c < MW; ++c)
                if (c < T::width(l) && k == 1 + (NLA - 1 - l) * MW + c) {
                    const f2 dl = mul2(M.sgl[l - 1][c], mul2(M.e, dup2(fac)));
                    acc_add(A.gbt[l - 1][c], dl);
#pragma unroll
                    for (int i = 0; i < MW; ++i)
                        if (i < T::in_w(l)) acc_fma(A.gWt[l - 1][i][c], l == 1 ? a0[i] : M.act[l - 1][i], dl);
                }
            fac = -fac / cA;
        }
    }
    // delta_0 of the pair, accumulated into gb0 and lifted by 2^100 (exact) for the piece split
    __device__ __forceinline__ static void delta0(const f2 (&sg0)[W0], f2 ef0, Acc& A, f2 (&v)[W0]) {
#pragma unroll
        for (int c = 0; c < W0; ++c) {
            const f2 d0 = mul2(sg0[c], ef0);
            acc_add(A.gb0[c], d0);
            v[c] = mul2(d0, dup2(1.2676506002282294e30f));
        }
    }
    // three bf16 pieces per unit -> the MN-major B operand rows of this thread: n = piece * W0 + unit, one 16-byte chunk per 8 n
    __device__ __forceinline__ static void store_pieces(f2 (&v)[W0], uint8_t* dst /* sD + tid * 16 */) {
        uint32_t pa[NN], pb[NN];      // f32 bit patterns whose upper halves are the bf16 pieces (row A / row B)
#pragma unroll
        for (int n = 0; n < NN; ++n) { pa[n] = 0u; pb[n] = 0u; }
#pragma unroll
        for (int c = 0; c < W0; ++c) {
#pragma unroll
            for (int piece = 0; piece < 3; ++piece) {
                if (piece < 2) {
                    const uint32_t ua = __float_as_uint(lo2(v[c])) & 0xFFFF0000u, ub = __float_as_uint(hi2(v[c])) & 0xFFFF0000u;
                    pa[piece * W0 + c] = ua; pb[piece * W0 + c] = ub;
                    v[c] = add2(v[c], mk2(-__uint_as_float(ua), -__uint_as_float(ub)));
                } else {      // the last piece needs no mask: the byte permutation below takes the upper halves only
                    pa[piece * W0 + c] = __float_as_uint(lo2(v[c])); pb[piece * W0 + c] = __float_as_uint(hi2(v[c]));
                }
            }
        }
#pragma unroll
        for (int q = 0; q < NN / 8; ++q) {
            uint4 wa, wb;
            wa.x = __byte_perm(pa[8 * q], pa[8 * q + 1], 0x7632); wa.y = __byte_perm(pa[8 * q + 2], pa[8 * q + 3], 0x7632);
            wa.z = __byte_perm(pa[8 * q + 4], pa[8 * q + 5], 0x7632); wa.w = __byte_perm(pa[8 * q + 6], pa[8 * q + 7], 0x7632);
            wb.x = __byte_perm(pb[8 * q], pb[8 * q + 1], 0x7632); wb.y = __byte_perm(pb[8 * q + 2], pb[8 * q + 3], 0x7632);
            wb.z = __byte_perm(pb[8 * q + 4], pb[8 * q + 5], 0x7632); wb.w = __byte_perm(pb[8 * q + 6], pb[8 * q + 7], 0x7632);
            *reinterpret_cast<uint4*>(dst + q * kTcChunkStride) = wa;
            *reinterpret_cast<uint4*>(dst + q * kTcChunkStride + 128 * 16) = wb;
        }
    }
    // stage the tail parameters of a branch: th_tail = theta + m * W0, n_tail floats
    __device__ __forceinline__ static void stage_tail(const float* th_tail, float* wp, uint32_t tid, uint32_t nthreads) {
        for (uint32_t k = tid; k < (uint32_t)T::n_tail(); k += nthreads) {
            const float w = th_tail[k];
            const bool scaled = NLA > 1 && ((int)k < T::w_off(NLA) || (int)k >= T::b_off(NLA > 1 ? 1 : 0));
            wp[k] = scaled ? w * cA : w;
        }
    }
    // CTA epilogue, first half: fixed-order reduction of the cross-row sums over row pairs, lanes and the NW compute warps,
    // written to the partial slot `pp` in param_vec order (rss at [P]); gb0 also to s_gb0 for the standardisation unfold
    // HALVING: the warp stage as a recursive-halving exchange -- at every step a lane hands one element of each pair to its
    // partner and keeps (and sums) the other, so NTACC values cost ~NTACC shuffles instead of 5 x NTACC (the stage is bound by the
    // SM's shuffle throughput); a fixed tree as well, in another order.  Used where the epilogue runs once per leapfrog step.
    template <bool HALVING = false, int NW = 4>
    __device__ __forceinline__ static void reduce_and_store(const Acc& A, float* red, uint32_t warp, uint32_t lane, uint32_t tid, float* pp,
                                                            uint32_t m, uint32_t P, float* s_gb0) {
        constexpr int NTACC = T::NTACC;
        if constexpr (HALVING) {
            if (warp < (uint32_t)NW) {
                float v[NTACC];
                int idx = 0;
                auto put = [&](AE v2) { v[idx++] = acc_total(v2); };
                put(A.rss);
#pragma unroll
                for (int c = 0; c < S; ++c) put(A.gWo[c]);
#pragma unroll
                for (int c = 0; c < W0; ++c) put(A.gb0[c]);
#pragma unroll
                for (int l = 1; l < NLA; ++l) {
#pragma unroll
                    for (int c = 0; c < MW; ++c) put(A.gbt[l - 1][c]);
#pragma unroll
                    for (int i = 0; i < MW; ++i)
#pragma unroll
                        for (int c = 0; c < MW; ++c) put(A.gWt[l - 1][i][c]);
                }
                halve_step<NTACC, 16>(v, lane);
                constexpr int cnt = halved_count(NTACC, 5);
                // element j of this lane is the sum of original value ((((2 j + b1) 2 + b2) 2 + b4) 2 + b8) 2 + b16
                float* rw = red + warp * NTACC;
#pragma unroll
                for (int j = 0; j < cnt; ++j) {
                    uint32_t k = (uint32_t)j;
#pragma unroll
                    for (int o = 1; o <= 16; o <<= 1) k = 2 * k + ((lane & (uint32_t)o) ? 1u : 0u);
                    if (k < (uint32_t)NTACC) rw[k] = v[j];
                }
            }
        } else
        if (warp < (uint32_t)NW) {
            float* rw = red + warp * NTACC;
            int idx = 0;
            auto put = [&](AE v2) {
                float v = acc_total(v2);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) rw[idx] = v;
                ++idx;
            };
            put(A.rss);
#pragma unroll
            for (int c = 0; c < S; ++c) put(A.gWo[c]);
#pragma unroll
            for (int c = 0; c < W0; ++c) put(A.gb0[c]);
#pragma unroll
            for (int l = 1; l < NLA; ++l) {
#pragma unroll
                for (int c = 0; c < MW; ++c) put(A.gbt[l - 1][c]);
#pragma unroll
                for (int i = 0; i < MW; ++i)
#pragma unroll
                    for (int c = 0; c < MW; ++c) put(A.gWt[l - 1][i][c]);
            }
        }
        __syncthreads();
        if (tid < NTACC) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) s += red[w * NTACC + tid];
            int idx = tid;
            if (idx == 0) pp[P] = s;
            else if (idx < 1 + S) pp[m * W0 + T::w_off(NLA) + (idx - 1)] = s;                       // output weights
            else if (idx < 1 + S + W0) { pp[m * W0 + T::b_off(0) + (idx - 1 - S)] = s; s_gb0[idx - 1 - S] = s; }
            else {
                int k = idx - (1 + S + W0);
                const int per = MW + MW * MW;
                const int l = 1 + k / per;
                k %= per;
                if (k < MW) {
                    if (k < T::width(l)) pp[m * W0 + T::b_off(l) + k] = s;
                } else {
                    k -= MW;
                    const int i = k / MW, c = k % MW;
                    if (i < T::in_w(l) && c < T::width(l)) pp[m * W0 + T::w_off(l) + c * T::in_w(l) + i] = s;
                }
            }
        }
        __syncthreads();
    }
};

// LEAN: gradient / leapfrog launches (targets given, no per-row outputs, backward always) -- the hot configuration
template <int H, int S, int D, int ACT, bool LEAN, int NCT>
__global__ void __launch_bounds__(128, 3) k1_tc(K1Args a) {
    using T = TailShape<H, S, D>;
    using C = TcShape<H, S, D>;
    constexpr int W0 = T::W0, W0P = T::W0P, NN = C::NN;
    extern __shared__ __align__(16) uint8_t smraw[];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform
    const uint32_t li = blockIdx.y, chunk = blockIdx.x;
    const uint32_t b = a.list ? a.list[li] : li;
    const BranchDesc& d = a.descs[b];        // descriptors and branch lists are written once, at net creation
    const uint32_t m = d.m, NC = NCT ? (uint32_t)NCT : d.nc, NKS = (NC + 1) >> 1, NCB = a.ncb;
    // Issuing a tcgen05.mma costs the issuing warp ~40 clk (measured: 8 forward MMAs + bulk copy 600 clk, 16 backward MMAs
    // 630 clk); on one warp that made it the straggler every other warp waited for (~1250 clk per super-tile).  The work is
    // therefore split: warp 0 forward row half 0, warp 1 forward row half 1 + the bulk copy, warp 2 / 3 backward row half 0 / 1.
    // The two backward halves accumulate in SEPARATE tensor-memory columns (each in its own fixed order) and are added at the end.
    // ---- shared memory carve-up
    uint8_t* sA = smraw + ((128u - (umma::smem_u32(smraw) & 127u)) & 127u);   // whole core matrices (keeps the shared address space)
    const uint32_t sa_bytes = NCB * kTcChunkStride;                        // one expanded-genotype buffer
    uint8_t* sW = sA + (size_t)2 * sa_bytes;
    uint32_t* sG = reinterpret_cast<uint32_t*>(sW + C::SW);               // [NC][128] packed words of the next super-tile (bulk copy target)
    uint8_t* sD = reinterpret_cast<uint8_t*>(sG) + (size_t)NCB * 512;
    // tail parameters as plain floats: FFMA2 takes a scalar register as a broadcast operand (.F32), so no {w, w} pairs are
    // needed and one LDS.128 brings four weights (the duplicated layout cost 2 shared-memory wavefronts per pair)
    float* wp = reinterpret_cast<float*>(__builtin_assume_aligned(sD + C::SD, 16));   // [n_tail]
    float* b0p = wp + ((T::n_tail() + 3) & ~3);                            // [W0P] first-layer bias with the means folded in
    float* red = b0p + 2 * ((T::n_tail() + 3) & ~3) - ((T::n_tail() + 3) & ~3) + 2 * W0P;   // [NRED] reduction scratch (layout size as before)
    // [0] forward MMAs done, [1] backward MMAs done (tcgen05.commit); [2] next tile expanded, [3] delta pieces written
    // (128 arrivals each); [4] packed words landed (bulk copy transaction bytes)
    uint64_t* mbar = reinterpret_cast<uint64_t*>(red + C::NRED);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 6);

    const float* th = a.theta + d.param_off;
    const float* mu = a.mu + d.col_off;
    const float* sd = a.sd + d.col_off;

    // ---- one-time setup: zero operand buffers, barriers, tensor memory
    {
        const uint32_t nz = (uint32_t)((sD + C::SD - sA) / 16);
        for (uint32_t k = tid; k < nz; k += 128) reinterpret_cast<uint4*>(sA)[k] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) {
        umma::mbar_init(&mbar[0], 2);        // two commits each: the MMA issue work is split over the four warps (see issue_*)
        umma::mbar_init(&mbar[1], 2);
        umma::mbar_init(&mbar[2], 128);
        umma::mbar_init(&mbar[3], 128);
        umma::mbar_init(&mbar[4], 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, C::TMEM_COLS);
    // Programmatic dependent launch (common.cuh): everything above touched only shared / tensor memory and ran under the
    // previous kernel (K2 of the last leapfrog step); parameters, targets and the branch status are read from here on.
    pdl_launch_dependents();
    pdl_wait();
    if (a.states && a.states[b].status != ST_RUNNING) {      // early-rejected / finished branch: release the tensor memory and leave
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();
        if (warp == 0) umma::tmem_dealloc(*tmem_slot, C::TMEM_COLS);
        return;
    }
    // the packed words of the CTA's first super-tile: requested NOW, so that the HBM round trip runs under the parameter
    // staging below instead of after it (at 8 GPUs a CTA covers only 49 super-tiles and its fixed cost is ~6 % of its time)
    const uint32_t t_begin = chunk * a.st_per_chunk;
    const uint32_t t_end = min(a.nst, t_begin + a.st_per_chunk);
    const uint32_t nit = t_end > t_begin ? t_end - t_begin : 0;
    const uint32_t* gwords = a.store_tc + (d.tc_off >> 2);
    umma::fence_async_smem();              // the zero fill above (generic proxy) precedes the bulk copy (async proxy) into the same buffer
    __syncthreads();                       // ... in every thread; and the barrier initialisation is visible to warp 1
    if (warp == 1 && nit > 0) {
        if (umma::elect_one()) umma::bulk_load(sG, gwords + (size_t)t_begin * NC * 128, NC * 512u, &mbar[4]);
        __syncwarp();
    }
    // tail parameters; what feeds an activated layer >= 1 (its weights and biases) carries the activation's pre-scale
    using TT = TcTail<H, S, D, ACT>;
    constexpr float cA = TT::cA;
    TT::stage_tail(th + m * W0, wp, tid, 128);
    __syncthreads();
    // ---- stage W' = W0 / sd (f32, in the delta buffer for the bias fold) and its three bf16 pieces
    float* wtmp = reinterpret_cast<float*>(sD);                // [m][W0], transient
    for (uint32_t k = tid; k < m * W0; k += 128) {
        const uint32_t j = k / W0, c = k % W0;
        const float w = __fdiv_rn(th[c * m + j], sd[j]);       // bed.rs:354 folded into the first layer
        wtmp[k] = w;
        const float v = w * pow2f(100 - 2 * (int)((j & 7u) >> 1));   // exact: undoes 4^p of the operand, lifts out of the subnormals
        const float p0 = bf16_round(v), r1 = v - p0, p1 = bf16_round(r1), p2 = bf16_round(r1 - p1);
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(sW + (j >> 3) * (NN * 16) + (j & 7u) * 2);
        dst[(0 * W0 + c) * 8] = __float2bfloat16_rn(p0);
        dst[(1 * W0 + c) * 8] = __float2bfloat16_rn(p1);
        dst[(2 * W0 + c) * 8] = __float2bfloat16_rn(p2);
    }
    __syncthreads();
    if (tid < W0P) {
        float acc = 0.f;
        if (tid < W0) {
            for (uint32_t j = 0; j < m; ++j) acc = fmaf(mu[j], wtmp[j * W0 + tid], acc);
            acc = (th[m * W0 + T::b_off(0) + tid] - acc) * cA;
        }
        b0p[tid] = acc;
    }
    __syncthreads();
    for (uint32_t k = tid; k < m * W0; k += 128) wtmp[k] = 0.f;   // the pad columns of the delta operand must stay zero
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tlane = tmem + ((warp * 32u) << 16);
    // forward A operand in tensor memory: row half h at columns [64 + 32 h, 96 + 32 h), column 4 i + o = markers 8 i + 2 o, + 1
    // of this thread's row.  The forward MMA then reads no shared memory for A (the kernel is shared-memory-wavefront bound).
    const uint32_t tA = tlane + 4 * NN;
    for (uint32_t c = 0; c < 64; c += 4) umma::tmem_st4(tA + c, 0u, 0u, 0u, 0u);
    umma::tmem_st_wait();
    const uint32_t sA_u = umma::smem_u32(sA), sD_u = umma::smem_u32(sD), sW_u = umma::smem_u32(sW);
    constexpr uint32_t idesc_f = umma::make_idesc(umma::FMT_BF16, umma::FMT_BF16, 0, 0, 128, NN);
    constexpr uint32_t idesc_b = umma::make_idesc(umma::FMT_BF16, umma::FMT_BF16, 1, 1, 64, NN);

    // ---- persistent per-thread accumulators (cross-row sums of the layers >= 1), one lane per row of the pair
    const f2 zero2 = dup2(0.f);
    typename TT::Acc A;
    A.clear();

    const size_t eoff = a.out_per_entry ? (size_t)li * a.n : 0;
    const size_t toff = (a.target_mode == TGT_PER_ENTRY) ? (size_t)li * a.n : 0;
    const float* tsrc = (!LEAN && a.target_mode == TGT_RESID_PLUS_PRED) ? a.resid : (a.tgt ? a.tgt + toff : nullptr);
    const bool bwd = LEAN || !a.fwd_only;

    // packed words of super-tile `st` -> shared memory, one bulk copy (the branch's super-tiles are contiguous: [st][NC][128] u32)
    auto issue_load = [&](uint32_t st) {        // whole issuer warp enters
        if (umma::elect_one()) umma::bulk_load(sG, gwords + (size_t)st * NC * 128, NC * 512u, &mbar[4]);
        __syncwarp();
    };
    // expand: one AND per two operand elements, 16-byte conflict-free stores (row t and row 128 + t)
    auto expand = [&](uint32_t buf) {
        uint8_t* rowA = sA + buf * sa_bytes + tid * 16;
        uint32_t xw[8];
        // all word loads first: the stores below are volatile asm, the compiler will not hoist loads over them
#pragma unroll
        for (int i = 0; i < 8; ++i) xw[i] = ((uint32_t)i < NC) ? sG[i * 128 + tid] : 0u;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if ((uint32_t)i < NC) {
                const uint32_t x = xw[i], y = x >> 8;
                const uint4 oa = make_uint4(x & 0x00030003u, x & 0x000C000Cu, x & 0x00300030u, x & 0x00C000C0u);
                const uint4 ob = make_uint4(y & 0x00030003u, y & 0x000C000Cu, y & 0x00300030u, y & 0x00C000C0u);
                *reinterpret_cast<uint4*>(rowA + i * kTcChunkStride) = oa;                   // backward operand (MN-major view)
                *reinterpret_cast<uint4*>(rowA + i * kTcChunkStride + 128 * 16) = ob;
                umma::tmem_st4(tA + 4 * i, oa.x, oa.y, oa.z, oa.w);                          // forward operand
                umma::tmem_st4(tA + 32 + 4 * i, ob.x, ob.y, ob.z, ob.w);
            }
        umma::tmem_st_wait();
        umma::fence_before_sync();
    };
    // forward contraction of the super-tile in buffer `buf`: z0 pieces -> tensor memory columns [0, 2 NN)
    // (whole issuer warp enters; one elected lane issues)
    const uint64_t dW_f = umma::make_desc(sW_u, NN * 16, 128);
    const uint64_t dA_b = umma::make_desc(sA_u, 128, kTcChunkStride), dD_b = umma::make_desc(sD_u, 128, kTcChunkStride);
    auto issue_fwd = [&](uint32_t h) {       // whole warp enters; row half h of the super-tile just expanded
        umma::fence_after_sync();
        if (umma::elect_one()) {
#pragma unroll
            for (uint32_t ks = 0; ks < 4; ++ks)
                if (ks < NKS)
                    umma::mma_f16_ts(tmem + h * NN, tmem + 4 * NN + h * 32 + ks * 8, dW_f + ((ks * 2u * (NN * 16)) >> 4), idesc_f, ks > 0);
            umma::commit(&mbar[0]);
        }
        __syncwarp();
    };
    // backward contraction, row half h (K steps 8h .. 8h + 7) into accumulator columns [(2 + h) NN, (3 + h) NN)
    auto issue_bwd = [&](uint32_t buf, uint32_t h, uint32_t it) {
        umma::fence_after_sync();
        const uint64_t base = dA_b + ((buf * sa_bytes) >> 4) + h * 128u, dbase = dD_b + h * 128u;
        if (umma::elect_one()) {
#pragma unroll
            for (uint32_t ks = 0; ks < kTcRows / 32; ++ks)
                umma::mma_f16(tmem + (2 + h) * NN, base + ks * 16u, dbase + ks * 16u, idesc_b, (it | ks) != 0);
            umma::commit(&mbar[1]);
        }
        __syncwarp();
    };

    auto load_targets = [&](uint32_t st) -> f2 {
        const uint32_t rA = st * kTcRows + tid, rB = rA + 128;
        if (!tsrc || st >= t_end) return zero2;
        return mk2(rA < a.n ? __ldg(tsrc + rA) : 0.f, rB < a.n ? __ldg(tsrc + rB) : 0.f);
    };
    f2 tg_next = load_targets(t_begin);
    if (nit > 0) {
        umma::mbar_wait(&mbar[4], 0);          // requested at the top of the kernel
        expand(0);
        umma::fence_async_smem();
        __syncthreads();
        if (warp < 2) issue_fwd(warp);
        if (warp == 1 && nit > 1) issue_load(t_begin + 1);
    }
    for (uint32_t it = 0; it < nit; ++it) {
        const uint32_t st = t_begin + it, buf = it & 1u;
        const uint32_t rowA_g = st * kTcRows + tid, rowB_g = rowA_g + 128;
        const bool vA = rowA_g < a.n, vB = rowB_g < a.n;
        f2 tg = tg_next;                 // loaded one super-tile ahead: nothing in this iteration waits on HBM
        tg_next = load_targets(st + 1);
        // ---- z0 of this super-tile (its forward contraction was issued one stage ago)
        umma::mbar_wait(&mbar[0], it & 1u);
        umma::fence_after_sync();
        float accA[16], accB[16];
        umma::tmem_ld16x2(tlane, tlane + NN, accA, accB);
        umma::fence_before_sync();      // these reads precede the next forward MMA (ordered by the barrier below)
        if (bwd && it > 0 && warp >= 2) {   // deferred: by now every thread has long written its delta pieces
            umma::mbar_wait(&mbar[3], (it - 1) & 1u);
            issue_bwd(buf ^ 1u, warp - 2, it - 1);
        }

        // ---- tail, part 1 (TcTail): remaining layers, error, rho_l and the cross-row sums of the layers >= 1
        f2 yh, sg0[W0], ef0;
        TT::part1(accA, accB, wp, b0p, tg, !LEAN && a.target_mode == TGT_RESID_PLUS_PRED, mk2(vA ? 1.f : 0.f, vB ? 1.f : 0.f), bwd, A,
                  yh, sg0, ef0);
        // ---- per-row outputs
        if (!LEAN) {
            auto put = [&](float* dst, uint32_t row, float v, int accumulate) {
                if (!dst || row >= a.n) return;
                float* p = dst + eoff + row;
                if (accumulate > 0) *p += v;
                else if (accumulate < 0) *p -= v;
                else *p = v;
            };
            if (a.target_mode == TGT_RESID_PLUS_PRED) {
                put(a.tgt_out, rowA_g, lo2(tg), 0); put(a.tgt_out, rowB_g, hi2(tg), 0);
                put(a.prev_out, rowA_g, lo2(yh), 0); put(a.prev_out, rowB_g, hi2(yh), 0);   // net.rs:279
            }
            put(a.yhat_out, rowA_g, lo2(yh), a.yhat_accumulate);
            put(a.yhat_out, rowB_g, hi2(yh), a.yhat_accumulate);
        }
        // ---- stage the next super-tile and start its forward contraction; it runs under part 2 of this tail
        if (bwd && it > 0) umma::mbar_wait(&mbar[1], (it - 1) & 1u);   // backward of the previous super-tile read buffer buf^1 and sD
        if (it + 1 < nit) {
            umma::mbar_wait(&mbar[4], (it + 1) & 1u);    // words of super-tile it + 1 (requested one stage ago)
            expand(buf ^ 1u);
        }
        // no CTA-wide barrier: every thread arrives and moves on, only the issuer warp waits for all 128 arrivals
        if (!bwd) umma::fence_async_smem();   // with a backward pass the image written above is first read by a backward MMA gated by the NEXT delta barrier, whose fence covers it
        umma::mbar_arrive(&mbar[2]);
        auto fwd_issue_point = [&]() {
            if (warp < 2) {
                umma::mbar_wait(&mbar[2], it & 1u);
                if (it + 1 < nit) issue_fwd(warp);
                if (warp == 1 && it + 2 < nit) issue_load(st + 2);        // every thread has read the staged words by now
            }
        };
        if (!bwd) { fwd_issue_point(); continue; }

        // ---- tail, part 2: delta_0 -> three bf16 pieces per unit in the backward B operand
        {
            f2 v[W0];
            TT::delta0(sg0, ef0, A, v);
            fwd_issue_point();            // deferred: the other warps have expanded their rows while this one did the above
            TT::store_pieces(v, sD + tid * 16);
        }
        umma::fence_async_smem();
        umma::mbar_arrive(&mbar[3]);
    }
    if (bwd && nit > 0 && warp >= 2) {       // the last super-tile's backward contraction
        umma::mbar_wait(&mbar[3], (nit - 1) & 1u);
        issue_bwd((nit - 1) & 1u, warp - 2, nit - 1);
    }
    const bool has_bwd = bwd && a.part && nit > 0;
    float sacc[16];
    if (has_bwd) {
        umma::mbar_wait(&mbar[1], (nit - 1) & 1u);
        umma::fence_after_sync();
        float sacc1[16];
        umma::tmem_ld16x2(tlane + 2 * NN, tlane + 3 * NN, sacc, sacc1);
#pragma unroll
        for (int k = 0; k < 16; ++k) sacc[k] += sacc1[k];       // row half 0 + row half 1, fixed order
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, C::TMEM_COLS);
    if (!bwd || !a.part) return;

    // ---- CTA epilogue: fixed-order reduction over row pairs, lanes and warps; unfold the standardisation
    float* pp = a.part + ((size_t)li * a.nchunk + chunk) * a.pstride;
    __shared__ float s_gb0[W0];
    TT::reduce_and_store(A, red, warp, lane, tid, pp, m, d.P, s_gb0);
    // first-layer weight gradient: accumulator row j lives in lane (j % 16) of warp j / 16 (M = 64 layout);
    // S_jc = sum of the three pieces * 2^33 / 4^p(j), then (S_jc - mu_j * gb0_c) / sd_j
    if (lane < 16) {
        const uint32_t j = warp * 16 + lane;
        if (j < m) {
            const float unscale = pow2f(33 - 2 * (int)((j & 7u) >> 1));
#pragma unroll
            for (int c = 0; c < W0; ++c) {
                const float s = (nit > 0 ? (sacc[c] + (sacc[W0 + c] + sacc[2 * W0 + c])) : 0.f) * unscale;
                pp[c * m + j] = __fdiv_rn(s - mu[j] * s_gb0[c], sd[j]);
            }
        }
    }
}

// k1_tc5_inst.cu: launches k1_tc5<H,S,D,ACT,true,*> when that tuple is instantiated (*launched), else leaves the launch to k1_tc
int launch_one_tc5(int H, int S, int D, int act, K1Args& a, uint32_t nlist, size_t smem, cudaStream_t st, bool* launched);

template <int H, int S, int D, int ACT>
int launch_one_tc_act(K1Args& a, uint32_t nlist, cudaStream_t st) {
    using C = TcShape<H, S, D>;
    constexpr bool kNct7 = ACT == BANN_TANH;      // the 49..56-marker specialisation exists for the benchmarked activation only
    static bool configured = false;
    if (!configured) {
        BANN_CUDA(cudaFuncSetAttribute(k1_tc<H, S, D, ACT, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem(8)));
        if (kNct7) BANN_CUDA(cudaFuncSetAttribute(k1_tc<H, S, D, ACT, true, kNct7 ? 7 : 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem(8)));
        BANN_CUDA(cudaFuncSetAttribute(k1_tc<H, S, D, ACT, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem(8)));
        configured = true;
    }
    const size_t smem = C::smem(a.ncb);
    dim3 grid(a.nchunk, nlist);
    const bool lean = !a.fwd_only && !a.yhat_out && a.target_mode != TGT_RESID_PLUS_PRED && a.tgt;
    if (lean && a.tc_variant >= 1) {           // the five-warp kernel (k1_tc5.cuh) where it is instantiated
        bool done = false;
        int rc = launch_one_tc5(H, S, D, ACT, a, nlist, smem, st, &done);
        if (rc) return rc;
        if (done) { a.tc_variant = 101; return 0; }   // tells launch_k1 which family ran (bann_net_last_k1_kernel)
    }
    if (lean && kNct7 && a.nc_uniform == 7) BANN_CUDA(launch_pdl(k1_tc<H, S, D, ACT, true, kNct7 ? 7 : 0>, grid, dim3(128), smem, st, a));   // 49..56 markers in every listed branch
    else if (lean) BANN_CUDA(launch_pdl(k1_tc<H, S, D, ACT, true, 0>, grid, dim3(128), smem, st, a));
    else BANN_CUDA(launch_pdl(k1_tc<H, S, D, ACT, false, 0>, grid, dim3(128), smem, st, a));
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    return 0;
}
template <int H, int S, int D>
int launch_one_tc(K1Args& a, uint32_t nlist, cudaStream_t st) {
    switch (a.act) {
        case BANN_TANH: return launch_one_tc_act<H, S, D, BANN_TANH>(a, nlist, st);
        case BANN_RELU: return launch_one_tc_act<H, S, D, BANN_RELU>(a, nlist, st);
        case BANN_LEAKY_RELU: return launch_one_tc_act<H, S, D, BANN_LEAKY_RELU>(a, nlist, st);
        case BANN_SILU: return launch_one_tc_act<H, S, D, BANN_SILU>(a, nlist, st);
        default: return launch_one_tc_act<H, S, D, BANN_IDENTITY>(a, nlist, st);
    }
}

// Tensor-core K1 for a homogeneous launch: any of the five activations, every listed branch of the same architecture with at
// most 64 markers and 3 * W0 <= 16, and a tensor-core store present.  Otherwise *launched stays false.
#ifdef BANN_K1_TC_IMPL
int launch_k1_tc(const std::vector<BranchDesc>& descs, int single_branch, K1Args& a, uint32_t nlist, int num_sms,
                        cudaStream_t st, bool* launched, uint32_t* nchunk_io, float** part_io, bann_net* net) {
    *launched = false;
    if (!a.store_tc) return 0;
    const BranchDesc& d0 = descs[single_branch >= 0 ? single_branch : 0];
    uint32_t max_m = d0.m;
    if (single_branch < 0) {
        for (const BranchDesc& d : descs) {
            if (d.nl != d0.nl) return 0;
            for (uint32_t l = 0; l < d.nl; ++l)
                if (d.widths[l] != d0.widths[l]) return 0;
            max_m = std::max(max_m, d.m);
        }
    }
    if (max_m > (uint32_t)kTcMaxMarkers) return 0;
    a.ncb = (max_m + 7) / 8;
    a.nc_uniform = a.ncb;
    if (single_branch < 0)
        for (const BranchDesc& d : descs)
            if (d.nc != a.ncb) a.nc_uniform = 0;
    const int D = (int)d0.nl - 2;
    const int S = (int)d0.widths[d0.nl - 2];
    const int H = D > 0 ? (int)d0.widths[0] : S;
    for (int l = 0; l < D; ++l)
        if ((int)d0.widths[l] != H) return 0;
    const uint32_t nst = a.nst;
    // single-branch launches (sequential schedule): 2 CTAs per SM instead of 4 -- fewer chunks for the reduction kernel on the
    // critical path of every leapfrog step (measured 27.7 -> 25.7 us per leapfrog at N = 100k); grouped launches as before
    const uint32_t want_mult = nlist == 1 ? 2 : 4;
    uint32_t want = (uint32_t)std::max<uint64_t>(1, ((uint64_t)num_sms * want_mult + nlist - 1) / nlist);
    uint32_t nchunk = std::min<uint32_t>(want, std::max<uint32_t>(1, nst));
    uint32_t spc = (nst + nchunk - 1) / nchunk;
    nchunk = (nst + spc - 1) / spc;
#define BANN_TRY_TC(HH, SS, DD)                                                                           \
    if (!*launched && H == HH && S == SS && D == DD) {                                                    \
        a.nchunk = nchunk;                                                                                \
        a.st_per_chunk = spc;                                                                             \
        if (part_io) {                                                                                    \
            if (nchunk == 1) *part_io = bann_net_gsum(net);                                               \
            else {                                                                                        \
                float* p = bann_net_partials(net, (size_t)nlist * nchunk * bann_net_pstride(net));        \
                if (!p) return -2;                                                                        \
                *part_io = p;                                                                             \
            }                                                                                             \
            a.part = *part_io;                                                                            \
        }                                                                                                 \
        *nchunk_io = nchunk;                                                                              \
        int rc = launch_one_tc<HH, SS, DD>(a, nlist, st);                                                 \
        if (rc) return rc;                                                                                \
        *launched = true;                                                                                 \
    }
    BANN_TRY_TC(5, 5, 1)
    BANN_TRY_TC(2, 2, 1)
    BANN_TRY_TC(4, 3, 1)
    BANN_TRY_TC(4, 3, 2)
    BANN_TRY_TC(5, 3, 2)
    BANN_TRY_TC(2, 2, 0)
    BANN_TRY_TC(3, 3, 1)
    BANN_TRY_TC(4, 4, 1)
    BANN_TRY_TC(5, 5, 2)
    BANN_TRY_TC(5, 5, 0)
#undef BANN_TRY_TC
    return 0;
}
#else
int launch_k1_tc(const std::vector<BranchDesc>& descs, int single_branch, K1Args& a, uint32_t nlist, int num_sms,
                        cudaStream_t st, bool* launched, uint32_t* nchunk_io, float** part_io, bann_net* net);
#endif

}  // namespace bann

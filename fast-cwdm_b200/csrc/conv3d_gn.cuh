// GroupNorm helpers shared by the conv3d kernels' epilogues / operand transforms (conv3d.cu, conv3d_chain.cu).
#pragma once
#include "common.cuh"

namespace fcwdm {

// ---- fused GroupNorm statistics (epilogue) ----------------------------------------------------------
// Reduce V per-lane values across the 32 lanes of a warp with a halving butterfly (lane pairs exchange HALF of their
// values per step: V/2 + V/4 + ... + 1 shuffles, then log2(32/V) full steps) and add value i into dst[i].
// 9 shuffles for V = 8 instead of 40 with one full butterfly per value.
template <int N, int OFF>
__device__ __forceinline__ void halve_step(float* a, int lane) {
    const bool upper = (lane & OFF) != 0;
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        const float send = upper ? a[i] : a[i + N / 2];
        const float keep = upper ? a[i + N / 2] : a[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
    }
}
template <int V>
__device__ __forceinline__ void warp_reduce_scatter(float* a, int lane, float* dst) {
    static_assert(V == 2 || V == 4 || V == 8 || V == 16, "unsupported value count");
    if constexpr (V == 16) { halve_step<16, 16>(a, lane); halve_step<8, 8>(a, lane); halve_step<4, 4>(a, lane); halve_step<2, 2>(a, lane); }
    if constexpr (V == 8) { halve_step<8, 16>(a, lane); halve_step<4, 8>(a, lane); halve_step<2, 4>(a, lane); }
    if constexpr (V == 4) { halve_step<4, 16>(a, lane); halve_step<2, 8>(a, lane); }
    if constexpr (V == 2) { halve_step<2, 16>(a, lane); }
#pragma unroll
    for (int off = 16 / V; off > 0; off >>= 1) a[0] += __shfl_xor_sync(0xffffffffu, a[0], off);
    if ((lane & (32 / V - 1)) == 0) dst[lane / (32 / V)] += a[0];
}
// vr: the 8 consecutive output channels [co, co+8) of this lane's voxel; CPG channels per group (CPG <= 4 here)
template <int CPG>
__device__ __forceinline__ void gn_accumulate(const float* vr, float* my_stat, int co, int lane) {
    constexpr int V = 2 * (8 / CPG);
    float a[V];
#pragma unroll
    for (int g = 0; g < 8 / CPG; ++g) {
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int e = 0; e < CPG; ++e) {
            s += vr[g * CPG + e];
            q = fmaf(vr[g * CPG + e], vr[g * CPG + e], q);
        }
        a[2 * g] = s;
        a[2 * g + 1] = q;
    }
    warp_reduce_scatter<V>(a, lane, my_stat + 2 * (co / CPG));
}
// CPG >= 8: the whole 8-channel run belongs to one group
__device__ __forceinline__ void gn_accumulate_wide(const float* vr, float* my_stat, int grp, int lane) {
    float a[2] = {0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        a[0] += vr[e];
        a[1] = fmaf(vr[e], vr[e], a[1]);
    }
    warp_reduce_scatter<2>(a, lane, my_stat + 2 * grp);
}
// ---- statistics of a 32-column chunk with ONE warp reduction --------------------------------------------
// One reduce per 8 columns (above) is a dependent chain of ~6 shuffles + a shared-memory update, 16 times per 128-column
// row: ~2.7 us per tile on an epilogue warp that has its scheduler to itself.  Here the lane first adds up (sum, sum of
// squares) of every 4-column run of a 32-column chunk in registers -- ca[16], run r at ca[2r], ca[2r+1] -- and the warp
// reduces once per chunk.  Supported channels per group: 4, 8, 16 and multiples of 32 (gn_chunk_ok).
__device__ __forceinline__ bool gn_chunk_ok(int cpg) { return cpg == 4 || cpg == 8 || cpg == 16 || (cpg > 0 && (cpg & 31) == 0); }
// vr: the 8 stored values of columns [8 g4, 8 g4 + 8) of the chunk (g4 = 0..3, compile-time after unrolling)
__device__ __forceinline__ void gn_chunk_add(float* ca, int g4, const float* vr) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        ca[4 * g4] += vr[e];
        ca[4 * g4 + 1] = fmaf(vr[e], vr[e], ca[4 * g4 + 1]);
        ca[4 * g4 + 2] += vr[4 + e];
        ca[4 * g4 + 3] = fmaf(vr[4 + e], vr[4 + e], ca[4 * g4 + 3]);
    }
}
// ch0: absolute channel of the chunk's first column (a multiple of 32); my_stat: this warp's [group][2] row; ca is consumed
__device__ __forceinline__ void gn_chunk_flush(float* ca, float* my_stat, int ch0, int cpg, int lane) {
    if (cpg == 4) {
        warp_reduce_scatter<16>(ca, lane, my_stat + 2 * (ch0 >> 2));
        return;
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {                                   // 8-column runs
        ca[2 * g] = ca[4 * g] + ca[4 * g + 2];
        ca[2 * g + 1] = ca[4 * g + 1] + ca[4 * g + 3];
    }
    if (cpg == 8) {
        warp_reduce_scatter<8>(ca, lane, my_stat + 2 * (ch0 >> 3));
        return;
    }
    ca[0] += ca[2]; ca[1] += ca[3];                                 // 16-column runs
    ca[2] = ca[4] + ca[6]; ca[3] = ca[5] + ca[7];
    if (cpg == 16) {
        warp_reduce_scatter<4>(ca, lane, my_stat + 2 * (ch0 >> 4));
        return;
    }
    ca[0] += ca[2]; ca[1] += ca[3];                                 // 32 or more channels per group: one group per chunk
    warp_reduce_scatter<2>(ca, lane, my_stat + 2 * (ch0 / cpg));
}

__device__ __forceinline__ float c_silu(float x) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    return x * fmaf(0.5f, t, 0.5f);
}

}  // namespace fcwdm

// K5w: weight gradient of the 3-D convolution (3x3x3 "same" or 1x1x1, stride 1) as a tcgen05 GEMM for sm_100a.
//
// Replaces the cuDNN wgrad that autograd runs for nn.Conv3d under scripts/train.py (guided_diffusion/nn.py:22-32,
// train_util.py:396-460).
//
//   dW[tap][co][ci] = sum over voxels v of  dY[v][co] * X[v + off(tap)][ci]          (X zero padded)
//
// GEMM view per tap: M = C_out, N = C_in, K = voxels.  Both operands are channels-last bf16, i.e. the GEMM's
// K dimension (voxels) is the OUTER dimension in memory: both are "MN-major" UMMA operands.  A TMA box of voxel rows
// x 64 channels with SWIZZLE_128B is exactly the canonical MN-major SWIZZLE_128B layout (8 voxel rows x 128 B per
// swizzle atom), so, as in the forward kernel, TMA output is consumed by tcgen05.mma without repacking:
//   A = dY tile, 16 (H) x 8 (W) voxels of one depth plane, M_TILE channels (64-channel chunks LBO apart);
//   B = X halo plane (rows x (8 + 2) voxels, N_TILE channels); tap (kh, kw) is the SAME plane read through a
//       descriptor whose start address is shifted by (kh*10 + kw) voxel rows and whose 8-row-group stride (SBO) is
//       the halo pitch -- zero padding comes from the TMA unit's out-of-bounds fill.
// One MMA consumes K = 16 voxels (two H rows of the tile); a tile is 8 such steps per tap.
//
// A CTA owns a group of taps (one kd, and one kh or all three) and an (M_TILE x N_TILE) channel block; it walks a
// strided subset ("split") of the voxel tiles, accumulating every tap in its own TMEM accumulator (C_out = 64 uses
// M = 64 MMAs; two taps then share a column range on interleaved datapath halves), and drains once at the end into a
// per-split fp32 workspace.  A second kernel sums the splits (fixed order: deterministic) and scatters into the
// reference's (C_out, C_in, k, k, k) fp32 gradient layout, optionally accumulating (weight-tied ResBlocks).
#include <stdlib.h>

#include "tc_ptx.cuh"

namespace fcwdm {

struct WgradArgs {
    int N, D, H, W;
    int n_wt, n_ht;
    int num_tiles;      // N * D * n_ht * n_wt
    int n_nb;           // blocks of N_TILE input channels (blockIdx.z = mb * n_nb + nb)
    int co_p, ci_p;     // workspace channel extents
    float* ws;          // [n_split][taps][co_p][ci_p]
};

template <int M_TILE, int N_TILE, int KS, int NKH>
struct WgCfg {
    static constexpr int PAD = KS / 2;
    static constexpr int NKW = KS;
    static constexpr int NT = NKH * NKW;                    // taps per CTA
    static constexpr int TAPS = KS * KS * KS;
    static constexpr int ROWP = 8 + 2 * PAD;                // halo pitch in voxel rows
    static constexpr int XROWS = 16 + (NKH == 3 ? 2 : 0);
    static constexpr int A_CHUNK = 128 * 128;               // 128 voxels x 64 channels bf16
    static constexpr int X_CHUNK_RAW = XROWS * ROWP * 128;
    static constexpr int X_CHUNK = (X_CHUNK_RAW + 1023) / 1024 * 1024;
    static constexpr int MC = M_TILE / 64, NC = N_TILE / 64;
    static constexpr int A_BYTES = MC * A_CHUNK;
    static constexpr int STAGE_BYTES = A_BYTES + NC * X_CHUNK;
    static constexpr int TX_BYTES = A_BYTES + NC * X_CHUNK_RAW;
    static constexpr int STAGES_RAW = (227 * 1024 - 2048) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
    static constexpr int COL_SLOTS = (M_TILE == 64) ? (NT + 1) / 2 : NT;
    static constexpr int ACC_COLS = COL_SLOTS * N_TILE;
    static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + 1024;
    static_assert(M_TILE == 64 || M_TILE == 128, "UMMA M");
    static_assert(N_TILE == 64 || N_TILE == 128 || N_TILE == 256, "N tile");
    static_assert(ACC_COLS <= 512, "accumulators exceed tensor memory");
    static_assert(STAGES >= 2, "not enough shared memory for a double buffer");
    static_assert(KS == 1 || KS == 3, "kernel size");
    static_assert(NKH == 1 || (NKH == 3 && KS == 3), "kh grouping");
};

// MN-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): canonical layout in 16-byte
// units ((8,n),(8,k)) : ((1,LBO),(8,SBO)) -- 64 contiguous channels, channel chunks LBO apart; 8 voxel rows of 128 B,
// 8-row groups SBO apart.
__device__ __forceinline__ uint64_t make_sw128_mn_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

constexpr int kWgWarpProd = 4, kWgWarpAlloc = 6, kWgWarpMma = 7;

template <int M_TILE, int N_TILE, int KS, int NKH>
__global__ void __launch_bounds__(256, 1) conv3d_wgrad_kernel(const __grid_constant__ CUtensorMap map_dy,
                                                              const __grid_constant__ CUtensorMap map_x,
                                                              const WgradArgs args) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    using Cfg = WgCfg<M_TILE, N_TILE, KS, NKH>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
    const uint32_t full = bars;
    const uint32_t empty = full + 8 * Cfg::STAGES;
    const uint32_t done = empty + 8 * Cfg::STAGES;
    const uint32_t tmem_slot = done + 8;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < Cfg::STAGES; ++i) {
            mbar_init(full + 8 * i, 1);
            mbar_init(empty + 8 * i, 1);
        }
        mbar_init(done, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == kWgWarpProd && lane == 0) {
        tma_prefetch_desc(&map_dy);
        tma_prefetch_desc(&map_x);
    }
    if (warp == kWgWarpAlloc) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // this CTA's share of the problem
    const int split = blockIdx.x, n_split = gridDim.x;
    const int grp = blockIdx.y;
    const int kd = (KS == 1) ? 0 : (NKH == 3 ? grp : grp / 3);
    const int kh0 = (KS == 1 || NKH == 3) ? 0 : grp % 3;
    const int mb = blockIdx.z / args.n_nb, nb = blockIdx.z % args.n_nb;
    const int co0 = mb * M_TILE, ci0 = nb * N_TILE;

    if (warp == kWgWarpProd) {
        // ================================ producer: dY tile + X halo plane per voxel tile ======================
        if (lane == 0) {
            uint32_t q = 0;
            for (int tile = split; tile < args.num_tiles; tile += n_split, ++q) {
                int r = tile;
                const int wt = r % args.n_wt; r /= args.n_wt;
                const int ht = r % args.n_ht; r /= args.n_ht;
                const int d = r % args.D;
                const int n = r / args.D;
                const uint32_t st = q % Cfg::STAGES, ph = (q / Cfg::STAGES) & 1;
                mbar_wait(empty + 8 * st, ph ^ 1);
                mbar_arrive_expect_tx(full + 8 * st, Cfg::TX_BYTES);
                const uint32_t sa = smem_base + st * Cfg::STAGE_BYTES;
#pragma unroll
                for (int c = 0; c < Cfg::MC; ++c)
                    tma_load_5d(sa + c * Cfg::A_CHUNK, &map_dy, full + 8 * st, co0 + 64 * c, wt * 8, ht * 16, d, n);
#pragma unroll
                for (int c = 0; c < Cfg::NC; ++c)
                    tma_load_5d(sa + Cfg::A_BYTES + c * Cfg::X_CHUNK, &map_x, full + 8 * st, ci0 + 64 * c,
                                wt * 8 - Cfg::PAD, ht * 16 - Cfg::PAD + kh0, d + kd - Cfg::PAD, n);
            }
        }
    } else if (warp == kWgWarpMma) {
        // ================================ MMA issuer ============================================================
        // instruction descriptor: D = f32, A = B = bf16, A and B MN-major (bits 15, 16), N >> 3 at [17,23), M >> 4 at [24,29)
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(N_TILE >> 3) << 17) | ((uint32_t)(M_TILE >> 4) << 24);
        const uint64_t a_base = make_sw128_mn_desc(smem_base, Cfg::A_CHUNK, 1024);
        const uint64_t b_base = make_sw128_mn_desc(smem_base + Cfg::A_BYTES, Cfg::X_CHUNK, Cfg::ROWP * 128);
        uint32_t q = 0;
        for (int tile = split; tile < args.num_tiles; tile += n_split, ++q) {
            const uint32_t st = q % Cfg::STAGES;
            mbar_wait(full + 8 * st, (q / Cfg::STAGES) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t so = (uint64_t)((st * Cfg::STAGE_BYTES) >> 4);
                const uint32_t acc = (q == 0) ? 0u : 1u;
#pragma unroll
                for (int j = 0; j < Cfg::NT; ++j) {
                    const int khl = j / Cfg::NKW, kw = j % Cfg::NKW;
                    const uint32_t dcol = (M_TILE == 64) ? (uint32_t)((j >> 1) * N_TILE) + ((uint32_t)((j & 1) * 16) << 16)
                                                         : (uint32_t)(j * N_TILE);
                    const uint64_t bo = so + (uint64_t)(((khl * Cfg::ROWP + kw) * 128) >> 4);
#pragma unroll
                    for (int s = 0; s < 8; ++s)
                        umma_bf16(tmem_base + dcol, a_base + so + (uint64_t)((s * 2048) >> 4),
                                  b_base + bo + (uint64_t)((s * 2 * Cfg::ROWP * 128) >> 4), idesc, (s == 0) ? acc : 1u);
                }
                umma_commit(empty + 8 * st);
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(done);
        __syncwarp();
    } else if (warp < 4) {
        // ================================ drain: TMEM -> fp32 workspace =========================================
        mbar_wait(done, 0);
        tc_fence_after();
        const int ew = warp;
        float* wsb = args.ws + (size_t)split * Cfg::TAPS * args.co_p * args.ci_p;
#pragma unroll 1
        for (int slot = 0; slot < Cfg::COL_SLOTS; ++slot) {
            int j, row;
            if (M_TILE == 64) {
                j = 2 * slot + (lane >> 4);
                row = ew * 16 + (lane & 15);
            } else {
                j = slot;
                row = ew * 32 + lane;
            }
            const bool live = j < Cfg::NT;
            const int khl = j / Cfg::NKW, kw = j % Cfg::NKW;
            const int tap = (KS == 1) ? 0 : ((kd * 3 + kh0 + khl) * 3 + kw);
            float* dst = wsb + ((size_t)tap * args.co_p + co0 + row) * args.ci_p + ci0;
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + slot * N_TILE;
#pragma unroll 1
            for (int c = 0; c < N_TILE; c += 32) {
                uint32_t v[32];
                tmem_ld_x16(taddr + c, v);
                tmem_ld_x16(taddr + c + 16, v + 16);
                tmem_ld_wait();
                if (live) {
#pragma unroll
                    for (int e = 0; e < 32; e += 4)
                        *reinterpret_cast<float4*>(dst + c + e) =
                            make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]),
                                        __uint_as_float(v[e + 3]));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kWgWarpAlloc) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------------
// C_out <= 64, C_in block of 64 (the full-resolution layers, 60 % of the wgrad FLOPs): tap-packed MMAs.
//
// One tcgen05.mma costs >= ~60 cycles whatever its shape (profiles/r01_mma_microbench.md), so a 64 x 64 x 16 MMA per
// tap runs the tensor pipe at ~40 %.  Here the NINE taps of one kd are produced by TWO MMAs per K-step:
//   dW[oh, ow] = sum_p dY[p + a] * X[p + b]   with tap (oh, ow) = b - a,
//   A operand (M) = dY loaded with a halo in H only (18 x 8 voxels, row pitch 8 = 1024 B, every shift 1024-B aligned):
//                   MMA 1 stacks the shifts a_h = -1 and a_h = 0 as two 64-channel chunks LBO = 1024 B apart (M = 128),
//                   MMA 2 takes a_h = +1 alone (M = 64);
//   B operand (N) = X loaded with a halo in W only (16 x 10 voxels): the shifts b_w = -1, 0, +1 are the three 64-channel
//                   "chunks" of ONE N = 192 operand whose chunk stride (LBO) is a single voxel row, 128 B.
//   MMA 1 -> taps (kh = 2, 1) x (kw = 0, 1, 2) in a 128-lane x 192-column accumulator, MMA 2 -> (kh = 0) x (kw = 0, 1, 2).
// 16 instructions per tile instead of 72; accumulators: 192 + 192 tensor-memory columns.
// ---------------------------------------------------------------------------------------------------
struct Wg64Cfg {
    static constexpr int DY_BYTES = 18 * 8 * 128;          // 18432, H-halo plane of dY
    static constexpr int X_BYTES = 16 * 10 * 128;          // 20480, W-halo plane of X
    static constexpr int STAGE_BYTES = DY_BYTES + X_BYTES; // 38912 (both 1024-B multiples)
    static constexpr int STAGES = 5;
    static constexpr int TMEM_COLS = 512;                  // 192 (M = 128) + 192 (M = 64) used
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + 1024;
};

__global__ void __launch_bounds__(256, 1) conv3d_wgrad64_kernel(const __grid_constant__ CUtensorMap map_dy,
                                                                const __grid_constant__ CUtensorMap map_x,
                                                                const WgradArgs args) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    using Cfg = Wg64Cfg;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
    const uint32_t full = bars;
    const uint32_t empty = full + 8 * Cfg::STAGES;
    const uint32_t done = empty + 8 * Cfg::STAGES;
    const uint32_t tmem_slot = done + 8;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < Cfg::STAGES; ++i) {
            mbar_init(full + 8 * i, 1);
            mbar_init(empty + 8 * i, 1);
        }
        mbar_init(done, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == kWgWarpProd && lane == 0) {
        tma_prefetch_desc(&map_dy);
        tma_prefetch_desc(&map_x);
    }
    if (warp == kWgWarpAlloc) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    asm volatile("griddepcontrol.wait;" ::: "memory");

    const int split = blockIdx.x, n_split = gridDim.x;
    const int kd = blockIdx.y;
    const int ci0 = (blockIdx.z % args.n_nb) * 64;            // C_out <= 64: a single output-channel block

    if (warp == kWgWarpProd) {
        if (lane == 0) {
            uint32_t q = 0;
            for (int tile = split; tile < args.num_tiles; tile += n_split, ++q) {
                int r = tile;
                const int wt = r % args.n_wt; r /= args.n_wt;
                const int ht = r % args.n_ht; r /= args.n_ht;
                const int d = r % args.D;
                const int n = r / args.D;
                const uint32_t st = q % Cfg::STAGES, ph = (q / Cfg::STAGES) & 1;
                mbar_wait(empty + 8 * st, ph ^ 1);
                mbar_arrive_expect_tx(full + 8 * st, Cfg::STAGE_BYTES);
                const uint32_t sa = smem_base + st * Cfg::STAGE_BYTES;
                tma_load_5d(sa, &map_dy, full + 8 * st, 0, wt * 8, ht * 16 - 1, d, n);                         // H halo
                tma_load_5d(sa + Cfg::DY_BYTES, &map_x, full + 8 * st, ci0, wt * 8 - 1, ht * 16, d + kd - 1, n);   // W halo
            }
        }
    } else if (warp == kWgWarpMma) {
        // D = f32, A = B = bf16, both MN-major; N = 192; M = 128 / 64
        constexpr uint32_t idesc_common = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(192 >> 3) << 17);
        constexpr uint32_t idesc128 = idesc_common | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t idesc64 = idesc_common | ((uint32_t)(64 >> 4) << 24);
        // A: 64-channel chunks 1024 B apart (the next H row of the halo plane), 8-row groups 1024 B apart (pitch 8)
        const uint64_t a_base = make_sw128_mn_desc(smem_base, 1024, 1024);
        // B: 64-channel chunks 128 B apart (the next voxel along W), 8-row groups 1280 B apart (pitch 10)
        const uint64_t b_base = make_sw128_mn_desc(smem_base + Cfg::DY_BYTES, 128, 1280);
        if (elect_one()) {
            uint32_t q = 0;
            for (int tile = split; tile < args.num_tiles; tile += n_split, ++q) {
                const uint32_t st = q % Cfg::STAGES;
                mbar_wait(full + 8 * st, (q / Cfg::STAGES) & 1);
                tc_fence_after();
                const uint64_t so = (uint64_t)((st * Cfg::STAGE_BYTES) >> 4);
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const uint32_t acc = (q == 0 && s == 0) ? 0u : 1u;
                    // output rows 2s, 2s+1: dY halo rows (2s + 1 + a_h); X rows 2s, 2s+1 (pitch 10)
                    const uint64_t b = b_base + so + (uint64_t)((s * 2 * 1280) >> 4);
                    umma_bf16(tmem_base, a_base + so + (uint64_t)(((2 * s + 0) * 1024) >> 4), b, idesc128, acc);          // a_h = -1, 0
                    umma_bf16(tmem_base + 192, a_base + so + (uint64_t)(((2 * s + 2) * 1024) >> 4), b, idesc64, acc);     // a_h = +1
                }
                umma_commit(empty + 8 * st);
            }
            umma_commit(done);
        }
        __syncwarp();
    } else if (warp < 4) {
        mbar_wait(done, 0);
        tc_fence_after();
        const int ew = warp;
        float* wsb = args.ws + (size_t)split * 27 * args.co_p * args.ci_p;
        // region 1 (M = 128): lane r < 64 -> a_h = -1 (kh = 2), output channel r; r >= 64 -> a_h = 0 (kh = 1), channel r - 64
        {
            const int r = ew * 32 + lane;
            const int kh = r < 64 ? 2 : 1;
            const int co = r & 63;
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16);
#pragma unroll 1
            for (int kw = 0; kw < 3; ++kw) {
                float* dst = wsb + ((size_t)((kd * 3 + kh) * 3 + kw) * args.co_p + co) * args.ci_p + ci0;
#pragma unroll 1
                for (int c = 0; c < 64; c += 32) {
                    uint32_t v[32];
                    tmem_ld_x16(taddr + kw * 64 + c, v);
                    tmem_ld_x16(taddr + kw * 64 + c + 16, v + 16);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 32; e += 4)
                        *reinterpret_cast<float4*>(dst + c + e) = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                                                              __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
                }
            }
        }
        // region 2 (M = 64, columns 192..383): accumulator row r lives in lane (r % 16) + 32 * (r / 16): a_h = +1 (kh = 0)
        {
            const bool live = lane < 16;
            const int co = ew * 16 + (lane & 15);
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + 192;
#pragma unroll 1
            for (int kw = 0; kw < 3; ++kw) {
                float* dst = wsb + ((size_t)((kd * 3 + 0) * 3 + kw) * args.co_p + co) * args.ci_p + ci0;
#pragma unroll 1
                for (int c = 0; c < 64; c += 32) {
                    uint32_t v[32];
                    tmem_ld_x16(taddr + kw * 64 + c, v);
                    tmem_ld_x16(taddr + kw * 64 + c + 16, v + 16);
                    tmem_ld_wait();
                    if (live) {
#pragma unroll
                        for (int e = 0; e < 32; e += 4)
                            *reinterpret_cast<float4*>(dst + c + e) = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                                                                  __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kWgWarpAlloc) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}


// dW[co][ci][tap] (+)= sum over splits of ws[split][tap][co][ci].  One block = one (tap, output channel, 32 input
// channels): 8 thread groups stride over the splits with coalesced 128-byte reads, a shared-memory tree adds the 8
// partial sums in a fixed order (deterministic), 32 threads scatter the result into the reference's layout.
__global__ void __launch_bounds__(256) wgrad_finalize_kernel(const float* __restrict__ ws, float* __restrict__ dw,
                                                             int n_split, int taps, int Cout, int Cin, int co_p,
                                                             int ci_p, int accumulate) {
    pdl_prologue();
    __shared__ float part[8][32];
    const int tap = blockIdx.z;
    const int co = blockIdx.y;
    const int c = threadIdx.x & 31, lane = threadIdx.x >> 5;
    const int ci = blockIdx.x * 32 + c;
    const size_t split_stride = (size_t)taps * co_p * ci_p;
    float a = 0.f;
    if (ci < Cin) {
        const float* p = ws + ((size_t)tap * co_p + co) * ci_p + ci;
        float a0 = 0.f, a1 = 0.f;
        int s = lane;
        for (; s + 8 < n_split; s += 16) {
            a0 += p[(size_t)s * split_stride];
            a1 += p[(size_t)(s + 8) * split_stride];
        }
        if (s < n_split) a0 += p[(size_t)s * split_stride];
        a = a0 + a1;
    }
    part[lane][c] = a;
    __syncthreads();
    if (lane == 0 && ci < Cin) {
        float v = ((part[0][c] + part[1][c]) + (part[2][c] + part[3][c])) + ((part[4][c] + part[5][c]) + (part[6][c] + part[7][c]));
        float* out = dw + ((size_t)co * Cin + ci) * taps + tap;
        *out = accumulate ? *out + v : v;
    }
}

// Few splits (the low-resolution layers, where the output-channel blocks already fill the GPU): one block = one output
// channel x 32 input channels x all taps; coalesced reads along ci, shared-memory transpose, coalesced writes of the
// 32 x taps contiguous floats.
__global__ void __launch_bounds__(256) wgrad_finalize_small_kernel(const float* __restrict__ ws, float* __restrict__ dw,
                                                                   int n_split, int taps, int Cout, int Cin, int co_p,
                                                                   int ci_p, int accumulate) {
    pdl_prologue();
    __shared__ float tile[27][33];
    const int co = blockIdx.y;
    const int cib = blockIdx.x * 32;
    const size_t split_stride = (size_t)taps * co_p * ci_p;
    for (int i = threadIdx.x; i < taps * 32; i += blockDim.x) {
        const int tap = i >> 5, c = i & 31;
        float a = 0.f;
        if (cib + c < Cin) {
            const float* p = ws + ((size_t)tap * co_p + co) * ci_p + cib + c;
            float v[8];                                   // n_split <= 8 here: all splits in flight at once, summed in order
#pragma unroll
            for (int sp = 0; sp < 8; ++sp) v[sp] = sp < n_split ? p[sp * split_stride] : 0.f;
#pragma unroll
            for (int sp = 0; sp < 8; ++sp) a += v[sp];
        }
        tile[tap][c] = a;
    }
    __syncthreads();
    const int n_ci = (Cin - cib) < 32 ? (Cin - cib) : 32;
    float* out = dw + ((size_t)co * Cin + cib) * taps;
    for (int i = threadIdx.x; i < n_ci * taps; i += blockDim.x) {
        const int c = i / taps, tap = i % taps;
        const float v = tile[tap][c];
        out[i] = accumulate ? out[i] + v : v;
    }
}

// (Cout, Cin, k^3) f32 -> (Cin, Cout, k^3) f32 with the taps reversed: the data-gradient of a stride-1 "same"
// convolution is the forward convolution of dY with these weights.
__global__ void __launch_bounds__(256) transpose_flip_weights_kernel(const float* __restrict__ w, float* __restrict__ wt,
                                                                     int Cout, int Cin, int taps) {
    pdl_prologue();
    const long long total = (long long)Cout * Cin * taps;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int tap = (int)(idx % taps);
    long long r = idx / taps;
    const int co = (int)(r % Cout);
    const int ci = (int)(r / Cout);
    wt[idx] = w[((long long)co * Cin + ci) * taps + (taps - 1 - tap)];
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_wg_encode = nullptr;

template <int M_TILE, int N_TILE, int KS, int NKH>
static cudaError_t wg_set_attr() {
    return cudaFuncSetAttribute(conv3d_wgrad_kernel<M_TILE, N_TILE, KS, NKH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                WgCfg<M_TILE, N_TILE, KS, NKH>::SMEM_BYTES);
}

int conv3d_wgrad_init_device() {
    if (g_wg_encode == nullptr) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        FCWDM_REQUIRE(e == cudaSuccess && fn != nullptr && qres == cudaDriverEntryPointSuccess, FCWDM_ERR_CUDA,
                      "fcwdm_init: cuTensorMapEncodeTiled entry point not available (%s)", cudaGetErrorString(e));
        g_wg_encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    cudaError_t e = cudaSuccess;
#define FCWDM_WG_SET(M, N, K, H) \
    if (e == cudaSuccess) e = wg_set_attr<M, N, K, H>();
    FCWDM_WG_SET(64, 64, 3, 3) FCWDM_WG_SET(64, 128, 3, 1) FCWDM_WG_SET(128, 64, 3, 1) FCWDM_WG_SET(128, 128, 3, 1)
    FCWDM_WG_SET(64, 64, 1, 1) FCWDM_WG_SET(128, 64, 1, 1) FCWDM_WG_SET(128, 128, 1, 1)
#undef FCWDM_WG_SET
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(conv3d_wgrad64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Wg64Cfg::SMEM_BYTES);
    FCWDM_REQUIRE(e == cudaSuccess, FCWDM_ERR_CUDA, "fcwdm_init: cudaFuncSetAttribute (wgrad) failed: %s",
                  cudaGetErrorString(e));
    return FCWDM_OK;
}

struct WgPlan {
    int m_tile, n_tile, nkh;
    int groups, n_mb, n_nb, co_p, ci_p, n_split;
    long long num_tiles;
};

static WgPlan wg_plan(int64_t N, int64_t D, int64_t H, int64_t W, int64_t Cin, int64_t Cout, int ksize) {
    WgPlan p;
    p.m_tile = Cout <= 64 ? 64 : 128;
    const int64_t ci64 = (Cin + 63) / 64 * 64;
    if (ksize == 1) {
        p.nkh = 1;
        p.n_tile = (p.m_tile == 128 && ci64 % 128 == 0) ? 128 : 64;
        p.groups = 1;
    } else if (p.m_tile == 64) {
        // C_in blocks of 64: all nine (kh, kw) taps of one kd per CTA (two taps share a column range on interleaved
        // datapath halves, 5 x 64 columns); C_in a multiple of 128: blocks of 128, three kw taps per CTA
        const bool wide = (ci64 % 128 == 0);
        p.nkh = wide ? 1 : 3;
        p.n_tile = wide ? 128 : 64;
        p.groups = wide ? 9 : 3;
    } else {
        p.nkh = 1;
        p.n_tile = (ci64 % 128 == 0) ? 128 : 64;
        p.groups = 9;
    }
    p.co_p = (int)((Cout + p.m_tile - 1) / p.m_tile * p.m_tile);
    p.ci_p = (int)((ci64 + p.n_tile - 1) / p.n_tile * p.n_tile);
    p.n_mb = p.co_p / p.m_tile;
    p.n_nb = p.ci_p / p.n_tile;
    p.num_tiles = N * D * ((H + 15) / 16) * ((W + 7) / 8);
    const long long per_split = (long long)p.groups * p.n_mb * p.n_nb;
    long long s = num_sms() / per_split;                             // one full wave of CTAs (one CTA per SM)
    if (s > p.num_tiles) s = p.num_tiles;
    if (s > 1024) s = 1024;
    if (s < 1) s = 1;
    p.n_split = (int)s;
    return p;
}

template <int M_TILE, int N_TILE, int KS, int NKH>
static int launch_wgrad(const CUtensorMap& mdy, const CUtensorMap& mx, const WgradArgs& a, const WgPlan& p, cudaStream_t st) {
    using Cfg = WgCfg<M_TILE, N_TILE, KS, NKH>;
    launch_k(conv3d_wgrad_kernel<M_TILE, N_TILE, KS, NKH>, dim3(p.n_split, p.groups, p.n_mb * p.n_nb), dim3(256),
             Cfg::SMEM_BYTES, st, mdy, mx, a);
    FCWDM_CHECK_LAUNCH("fcwdm_conv3d_wgrad");
    return FCWDM_OK;
}

}  // namespace fcwdm

using namespace fcwdm;

extern "C" int64_t fcwdm_conv3d_wgrad_workspace_bytes(int64_t N, int64_t D, int64_t H, int64_t W, int64_t Cin,
                                                      int64_t Cout, int ksize) {
    if (N < 0 || D < 0 || H < 0 || W < 0 || Cin <= 0 || Cout <= 0 || (ksize != 1 && ksize != 3)) return -1;
    if (N * D * H * W == 0) return 0;
    const WgPlan p = wg_plan(N, D, H, W, Cin, Cout, ksize);
    return (int64_t)p.n_split * ksize * ksize * ksize * p.co_p * p.ci_p * (int64_t)sizeof(float);
}

extern "C" int fcwdm_conv3d_wgrad(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, float* dw, void* workspace,
                                  int64_t workspace_bytes, int accumulate, int64_t N, int64_t D, int64_t H, int64_t W,
                                  int64_t Cin, int64_t Cout, int ksize, void* stream) {
    FCWDM_REQUIRE(x && dy && dw, FCWDM_ERR_INVALID, "fcwdm_conv3d_wgrad: null pointer");
    FCWDM_REQUIRE(N >= 0 && D >= 0 && H >= 0 && W >= 0 && Cin > 0 && Cout > 0, FCWDM_ERR_INVALID,
                  "fcwdm_conv3d_wgrad: bad dimension");
    FCWDM_REQUIRE(ksize == 1 || ksize == 3, FCWDM_ERR_UNSUPPORTED, "fcwdm_conv3d_wgrad: kernel size %d (only 1, 3)", ksize);
    FCWDM_REQUIRE(D < 32768 && H < 32768 && W < 32768 && N < 32768, FCWDM_ERR_UNSUPPORTED, "fcwdm_conv3d_wgrad: dim too large");
    const int taps = ksize * ksize * ksize;
    cudaStream_t st = (cudaStream_t)stream;
    if (N * D * H * W == 0) {
        if (!accumulate) cudaMemsetAsync(dw, 0, sizeof(float) * Cout * Cin * taps, st);
        return FCWDM_OK;
    }
    const WgPlan p = wg_plan(N, D, H, W, Cin, Cout, ksize);
    const int64_t ci64 = (Cin + 63) / 64 * 64, co64 = (Cout + 63) / 64 * 64;
    FCWDM_REQUIRE(x_ld >= ci64 && x_ld % 8 == 0 && dy_ld >= co64 && dy_ld % 8 == 0, FCWDM_ERR_INVALID,
                  "fcwdm_conv3d_wgrad: x_ld (%lld) / dy_ld (%lld) must cover the channel counts rounded up to 64 "
                  "(%lld / %lld); pad channels must hold finite values",
                  (long long)x_ld, (long long)dy_ld, (long long)ci64, (long long)co64);
    FCWDM_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)dy % 16 == 0) && ((uintptr_t)dw % 16 == 0) &&
                      ((uintptr_t)workspace % 16 == 0),
                  FCWDM_ERR_INVALID, "fcwdm_conv3d_wgrad: pointers must be 16-byte aligned");
    const int64_t need = (int64_t)p.n_split * taps * p.co_p * p.ci_p * (int64_t)sizeof(float);
    FCWDM_REQUIRE(workspace != nullptr && workspace_bytes >= need, FCWDM_ERR_INVALID,
                  "fcwdm_conv3d_wgrad: workspace of %lld bytes needed (fcwdm_conv3d_wgrad_workspace_bytes), got %lld",
                  (long long)need, (long long)workspace_bytes);
    if (g_wg_encode == nullptr) {
        int dev = 0;
        cudaGetDevice(&dev);
        int rc = fcwdm_init(dev);
        if (rc) return rc;
    }
    const int pad = ksize / 2;
    CUtensorMap mdy, mx;
    {
        cuuint64_t dims[5] = {(cuuint64_t)dy_ld, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
        cuuint64_t strides[4] = {(cuuint64_t)dy_ld * 2, (cuuint64_t)W * dy_ld * 2, (cuuint64_t)H * W * dy_ld * 2,
                                 (cuuint64_t)D * H * W * dy_ld * 2};
        cuuint32_t box[5] = {64, 8, 16, 1, 1};
        cuuint32_t es[5] = {1, 1, 1, 1, 1};
        CUresult r = g_wg_encode(&mdy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(dy), dims, strides, box, es,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        FCWDM_REQUIRE(r == CUDA_SUCCESS, FCWDM_ERR_CUDA, "fcwdm_conv3d_wgrad: dY tensor map encode failed (%d)", (int)r);
    }
    {
        cuuint64_t dims[5] = {(cuuint64_t)x_ld, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
        cuuint64_t strides[4] = {(cuuint64_t)x_ld * 2, (cuuint64_t)W * x_ld * 2, (cuuint64_t)H * W * x_ld * 2,
                                 (cuuint64_t)D * H * W * x_ld * 2};
        cuuint32_t box[5] = {64, (cuuint32_t)(8 + 2 * pad), (cuuint32_t)(16 + (p.nkh == 3 ? 2 : 0)), 1, 1};
        cuuint32_t es[5] = {1, 1, 1, 1, 1};
        CUresult r = g_wg_encode(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides, box, es,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        FCWDM_REQUIRE(r == CUDA_SUCCESS, FCWDM_ERR_CUDA, "fcwdm_conv3d_wgrad: X tensor map encode failed (%d)", (int)r);
    }
    WgradArgs a;
    a.N = (int)N; a.D = (int)D; a.H = (int)H; a.W = (int)W;
    a.n_wt = (int)((W + 7) / 8);
    a.n_ht = (int)((H + 15) / 16);
    a.num_tiles = (int)p.num_tiles;
    a.n_nb = p.n_nb;
    a.co_p = p.co_p; a.ci_p = p.ci_p;
    a.ws = (float*)workspace;
    int rc = FCWDM_ERR_UNSUPPORTED;
    if (ksize == 3) {
        if (p.m_tile == 64 && p.n_tile == 64 && p.nkh == 3) {
            static const bool packed_off = getenv("FCWDM_WGRAD_PACKED") != nullptr && atoi(getenv("FCWDM_WGRAD_PACKED")) == 0;
            if (packed_off) {
                rc = launch_wgrad<64, 64, 3, 3>(mdy, mx, a, p, st);
            } else {
                // tap-packed kernel: dY with a halo in H only, X with a halo in W only
                CUtensorMap mdy2, mx2;
                cuuint32_t es[5] = {1, 1, 1, 1, 1};
                cuuint64_t ddims[5] = {(cuuint64_t)dy_ld, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
                cuuint64_t dstr[4] = {(cuuint64_t)dy_ld * 2, (cuuint64_t)W * dy_ld * 2, (cuuint64_t)H * W * dy_ld * 2,
                                      (cuuint64_t)D * H * W * dy_ld * 2};
                cuuint32_t dbox[5] = {64, 8, 18, 1, 1};
                CUresult r1 = g_wg_encode(&mdy2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(dy), ddims, dstr, dbox, es,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                cuuint64_t xdims[5] = {(cuuint64_t)x_ld, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
                cuuint64_t xstr[4] = {(cuuint64_t)x_ld * 2, (cuuint64_t)W * x_ld * 2, (cuuint64_t)H * W * x_ld * 2,
                                      (cuuint64_t)D * H * W * x_ld * 2};
                cuuint32_t xbox[5] = {64, 10, 16, 1, 1};
                CUresult r2 = g_wg_encode(&mx2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), xdims, xstr, xbox, es,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                FCWDM_REQUIRE(r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS, FCWDM_ERR_CUDA,
                              "fcwdm_conv3d_wgrad: tensor map encode failed (%d, %d)", (int)r1, (int)r2);
                launch_k(conv3d_wgrad64_kernel, dim3(p.n_split, 3, p.n_nb), dim3(256), Wg64Cfg::SMEM_BYTES, st, mdy2, mx2, a);
                FCWDM_CHECK_LAUNCH("fcwdm_conv3d_wgrad");
                rc = FCWDM_OK;
            }
        }
        else if (p.m_tile == 64 && p.n_tile == 128) rc = launch_wgrad<64, 128, 3, 1>(mdy, mx, a, p, st);
        else if (p.m_tile == 128 && p.n_tile == 64) rc = launch_wgrad<128, 64, 3, 1>(mdy, mx, a, p, st);
        else if (p.m_tile == 128 && p.n_tile == 128) rc = launch_wgrad<128, 128, 3, 1>(mdy, mx, a, p, st);
    } else {
        if (p.m_tile == 64) rc = launch_wgrad<64, 64, 1, 1>(mdy, mx, a, p, st);
        else if (p.n_tile == 64) rc = launch_wgrad<128, 64, 1, 1>(mdy, mx, a, p, st);
        else rc = launch_wgrad<128, 128, 1, 1>(mdy, mx, a, p, st);
    }
    FCWDM_REQUIRE(rc != FCWDM_ERR_UNSUPPORTED, FCWDM_ERR_UNSUPPORTED, "fcwdm_conv3d_wgrad: no kernel variant for this shape");
    if (rc) return rc;
    if (p.n_split <= 8)
        launch_k(wgrad_finalize_small_kernel, dim3((unsigned)((Cin + 31) / 32), (unsigned)Cout), dim3(256), 0, st,
                 (const float*)workspace, dw, p.n_split, taps, (int)Cout, (int)Cin, p.co_p, p.ci_p, accumulate);
    else
        launch_k(wgrad_finalize_kernel, dim3((unsigned)((Cin + 31) / 32), (unsigned)Cout, (unsigned)taps), dim3(256), 0, st,
                 (const float*)workspace, dw, p.n_split, taps, (int)Cout, (int)Cin, p.co_p, p.ci_p, accumulate);
    FCWDM_CHECK_LAUNCH("fcwdm_conv3d_wgrad (finalize)");
    return FCWDM_OK;
}

extern "C" int fcwdm_conv3d_transpose_flip_weights(const float* w, float* wt, int64_t Cout, int64_t Cin, int ksize,
                                                   void* stream) {
    FCWDM_REQUIRE(w && wt, FCWDM_ERR_INVALID, "fcwdm_conv3d_transpose_flip_weights: null pointer");
    FCWDM_REQUIRE(Cout > 0 && Cin > 0 && (ksize == 1 || ksize == 3), FCWDM_ERR_INVALID,
                  "fcwdm_conv3d_transpose_flip_weights: bad argument");
    const int taps = ksize * ksize * ksize;
    const long long total = (long long)Cout * Cin * taps;
    launch_k(transpose_flip_weights_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, w, wt,
             (int)Cout, (int)Cin, taps);
    FCWDM_CHECK_LAUNCH("fcwdm_conv3d_transpose_flip_weights");
    return FCWDM_OK;
}

// ---------------------------------------------------------------------------------------------------
// One-launch re-packing of every conv weight of a model (after each optimizer step the fp32 master weights change and
// all bf16 operand copies -- forward and data-gradient forms, single-CTA and CTA-pair layouts -- are rebuilt).
// ---------------------------------------------------------------------------------------------------
namespace fcwdm {

struct PackJob {            // mirrored by fcwdm/engine.py (8 x int64)
    const float* src;       // (Cout_w, Cin_w, taps) f32 master weight
    __nv_bfloat16* dst;
    long long O, I;         // output / input channels of the conv being packed (swapped w.r.t. the master for dgrad)
    long long taps;
    long long pair;         // 1: CTA-pair layout [kh*3+kw][kd][O_p][64]; 0: [tap][O_p][I_p]
    long long transposed;   // 1: data-gradient form w'[o][i][tap] = w[i][o][taps-1-tap]
    long long total;        // packed elements
};

// One block = one tile of 512 (o, i) pairs x all taps, staged through shared memory so that BOTH sides are coalesced:
// the master weight is read in runs that are contiguous in (inner channel, tap) -- 27 floats per (o, i) pair lie together --
// and the packed tensor is written in runs that are contiguous in i.  Forward form: tile 8 o x 64 i (source rows = o);
// data-gradient form: tile 16 o x 32 i (source rows = i, contiguous along o).  (The first version gathered one element
// per thread with a 108-byte stride: 0.5 ms per training step for 54 M parameters; this one is bound by the 0.33 GB it moves.)
constexpr int kPackPairs = 512;
constexpr int kPackTapsMax = 27;

// TAPS and the tile shape are compile-time so that every index split below divides by a constant (the generic form spent
// most of its time in 32-bit integer division: 0.55 ms for the 0.65 GB it moves)
template <int TAPS, bool TRANSPOSED>
__device__ __forceinline__ void pack_job_tiles(const PackJob& j, __nv_bfloat16* tile) {
    const int O = (int)j.O, I = (int)j.I;
    const int O_p = j.pair ? (O <= 16 ? 16 : 64) : (O + 15) / 16 * 16;
    const int I_p = j.pair ? 64 : (I + 63) / 64 * 64;
    constexpr int TO = TRANSPOSED ? 16 : 8, TI = kPackPairs / TO;
    constexpr int RUN = (TRANSPOSED ? TO : TI) * TAPS;            // source elements that are contiguous per source row
    const int tiles_i = (I_p + TI - 1) / TI, tiles_o = (O_p + TO - 1) / TO;
    const bool pair = j.pair != 0;
    for (int t = blockIdx.x; t < tiles_o * tiles_i; t += gridDim.x) {
        const int o0 = (t / tiles_i) * TO, i0 = (t % tiles_i) * TI;
        __syncthreads();                                          // the previous tile has been written out
#pragma unroll 4
        for (int e = threadIdx.x; e < kPackPairs * TAPS; e += 256) {
            const int row = e / RUN, rem = e - row * RUN;         // source row inside the tile, offset inside its run
            const int col = rem / TAPS, tap = rem - col * TAPS;
            float v = 0.f;
            int o_l, i_l, tap_d;
            if (TRANSPOSED) {                                     // w'[o][i][tap] = w[i][o][taps - 1 - tap]
                i_l = row; o_l = col; tap_d = TAPS - 1 - tap;
                if (i0 + i_l < I && o0 + o_l < O) v = __ldg(j.src + ((long long)(i0 + i_l) * O + o0) * TAPS + rem);
            } else {
                o_l = row; i_l = col; tap_d = tap;
                if (o0 + o_l < O && i0 + i_l < I) v = __ldg(j.src + ((long long)(o0 + o_l) * I + i0) * TAPS + rem);
            }
            tile[(o_l * TI + i_l) * TAPS + tap_d] = __float2bfloat16_rn(v);
        }
        __syncthreads();
#pragma unroll 4
        for (int f = threadIdx.x; f < kPackPairs * TAPS; f += 256) {
            const int tap = f / kPackPairs, p = f - tap * kPackPairs;
            const int o = o0 + p / TI, i = i0 + p % TI;
            if (o < O_p && i < I_p) {
                long long idx;
                if (pair) {
                    const int kd = tap / 9, t2 = tap - kd * 9;    // pair layout: [kh*3+kw][kd][O_p][64]
                    idx = (((long long)t2 * 3 + kd) * O_p + o) * 64 + i;
                } else {
                    idx = ((long long)tap * O_p + o) * I_p + i;
                }
                j.dst[idx] = tile[p * TAPS + tap];
            }
        }
    }
}

__global__ void __launch_bounds__(256) pack_all_kernel(const PackJob* __restrict__ jobs) {
    pdl_prologue();
    __shared__ __nv_bfloat16 tile[kPackPairs * kPackTapsMax];
    const PackJob j = jobs[blockIdx.y];
    if (j.taps == 27) {
        if (j.transposed) pack_job_tiles<27, true>(j, tile); else pack_job_tiles<27, false>(j, tile);
    } else {
        if (j.transposed) pack_job_tiles<1, true>(j, tile); else pack_job_tiles<1, false>(j, tile);
    }
}

}  // namespace fcwdm

extern "C" int fcwdm_conv3d_pack_all(const void* jobs, int64_t n_jobs, int64_t max_total, void* stream) {
    FCWDM_REQUIRE(jobs != nullptr || n_jobs == 0, FCWDM_ERR_INVALID, "fcwdm_conv3d_pack_all: null job table");
    FCWDM_REQUIRE(n_jobs >= 0 && n_jobs <= 65535 && max_total >= 0, FCWDM_ERR_INVALID, "fcwdm_conv3d_pack_all: bad argument");
    if (n_jobs == 0 || max_total == 0) return FCWDM_OK;
    long long bx = (max_total + fcwdm::kPackPairs * 27 - 1) / (fcwdm::kPackPairs * 27);   // tiles of the largest 3x3x3 job
    if (bx > 2048) bx = 2048;
    if (bx < 1) bx = 1;
    launch_k(fcwdm::pack_all_kernel, dim3((unsigned)bx, (unsigned)n_jobs), dim3(256), 0, (cudaStream_t)stream,
             (const fcwdm::PackJob*)jobs);
    FCWDM_CHECK_LAUNCH("fcwdm_conv3d_pack_all");
    return FCWDM_OK;
}

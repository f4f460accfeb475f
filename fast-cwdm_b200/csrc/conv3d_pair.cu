// K5b: 3x3x3 "same" conv3d for C_in <= 64 and C_out <= 64 as a kd-fused, 2-CTA (cta_group::2) implicit GEMM.
//
// This is the kernel for the full-resolution 64-channel layers of the wavelet U-Net (ten 64->64 convs + stem +
// output conv per denoiser call = 64 % of its FLOPs).  The single-CTA kernel (conv3d.cu) is bound by the 128 B/clk
// shared-memory port there: in SS mode every MMA re-reads A (4 KB) and B (N*32 B) and with N = 64 that is
// 192 B/clk.  Two changes remove the bound:
//
//  * kd fusion.  The three depth taps (kd = 0,1,2) of a filter column (kh,kw) are concatenated along N:
//    B' = [kd][C_out][64] = 192 rows.  One MMA of input plane p against B'(kh,kw) produces the contributions to
//    the THREE output planes p+1, p, p-1 in adjacent accumulator column blocks, so every activation row is read
//    from shared memory 9 times instead of 27.
//  * CTA pairs.  Two CTAs of a cluster run one M = 256 MMA (128 voxels each, side by side in H or W); each CTA
//    keeps HALF of B' (96 rows per filter column, all 9 columns = 108 KB) resident in its shared memory for the
//    whole kernel, so weights are fetched from L2 once per CTA and B smem reads per SM are halved.
//    Per SM and MMA: A 4 KB + B 3 KB per 96 cycles = 73 B/clk.
//
// Work decomposition: a pair walks a depth segment of one (32x8 or 16x16 voxel) column, plane by plane.  Input
// planes stream through a small TMA ring (halo planes of 18x10 voxels x 64 ch, SWIZZLE_128B, zero-filled
// out of bounds); every input plane gets the same 36 MMAs (9 filter columns x 4 K-steps, N = 192, always
// accumulating).  Output-plane accumulators live in a 6-deep logical ring over 8 physical TMEM slots:
// the triple (p+1, p, p-1) must be column-contiguous, so two of every six planes are split over two physical
// slots (6/7 mirror 0/1) and the epilogue adds the halves.  The epilogue (4 warps per CTA, each CTA drains its own
// 128 TMEM lanes) adds bias / timestep embedding / residual, stores bf16, optionally accumulates GroupNorm
// statistics, and ZEROES the drained slots (so the MMAs never need a non-accumulating first touch).  Segment edges
// are handled by two dummy output planes per side that are accumulated, drained and discarded.
#include "tc_ptx.cuh"

namespace fcwdm {

struct PairArgs {
    int N, D, H, W;
    int Cout;
    int pair_w;              // 1: the two CTAs of a pair sit side by side along W, 0: along H
    int n_hu, n_wu;          // pair-units along H and W
    int seg_len, n_seg;      // output planes per depth segment, segments per column
    int num_items;           // N * n_hu * n_wu * n_seg
    const float* bias;
    const float* chan_bias;
    long long cb_ld;
    const __nv_bfloat16* residual;
    long long res_ld;
    __nv_bfloat16* y;
    long long y_ld;
    double* gn_stats;
    int gn_cpg, gn_groups;
    // GN_IN variant: the conv input is SiLU(GroupNorm(x)) applied on the fly by the operand producers
    const __nv_bfloat16* x;      // raw (un-normalised) input, channels-last
    long long x_ld;
    int Cin;
    const double* gi_stats;      // [N][FCWDM_GN_STAT_REPLICAS][gi_groups][2] statistics of x
    const float* gi_gamma;
    const float* gi_beta;
    int gi_groups;
    float gi_eps;
};

template <int N_TILE>
struct PairCfg {
    static constexpr int NF = 3 * N_TILE;                     // fused N (kd-major)
    static constexpr int B_TAP_BYTES = (NF / 2) * 128;        // this CTA's half of one filter column
    static constexpr int B_BYTES = 9 * B_TAP_BYTES;
    static constexpr int ROWP = 10, HROWS = 18;
    static constexpr int PLANE_BYTES = HROWS * ROWP * 128;    // 23040
    static constexpr int SLOT_BYTES = 23552;                  // 1024-aligned
    static constexpr int A_SLOTS_RAW = (227 * 1024 - 3584 - B_BYTES) / SLOT_BYTES;
    static constexpr int A_SLOTS = A_SLOTS_RAW > 6 ? 6 : A_SLOTS_RAW;
    static constexpr int RING = 6;
    static constexpr int TMEM_COLS = 8 * N_TILE < 32 ? 32 : 8 * N_TILE;
    static constexpr int SMEM_BYTES = 1024 + B_BYTES + A_SLOTS * SLOT_BYTES + 2048 + 512;
    static_assert(B_TAP_BYTES % 1024 == 0, "weight tiles must stay 1024-B aligned");
    static_assert(A_SLOTS >= 3, "not enough plane slots");
    static_assert(TMEM_COLS <= 512, "accumulator ring exceeds tensor memory");
};

struct PairItem {
    int n, h0, w0, d_begin, L;
};
__device__ __forceinline__ PairItem decode_item(int item, const PairArgs& a, int rank) {
    PairItem it;
    int r = item;
    const int seg = r % a.n_seg; r /= a.n_seg;
    const int wu = r % a.n_wu; r /= a.n_wu;
    const int hu = r % a.n_hu; r /= a.n_hu;
    it.n = r;
    it.h0 = a.pair_w ? hu * 16 : hu * 32 + rank * 16;
    it.w0 = a.pair_w ? wu * 16 + rank * 8 : wu * 8;
    it.d_begin = seg * a.seg_len;
    const int d_end = it.d_begin + a.seg_len < a.D ? it.d_begin + a.seg_len : a.D;
    it.L = d_end - it.d_begin;
    return it;
}

// ---- GroupNorm statistics helpers (same scheme as conv3d.cu) -----------------------------------------
template <int N, int OFF>
__device__ __forceinline__ void p_halve_step(float* a, int lane) {
    const bool upper = (lane & OFF) != 0;
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        const float send = upper ? a[i] : a[i + N / 2];
        const float keep = upper ? a[i + N / 2] : a[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
    }
}
template <int V>
__device__ __forceinline__ void p_reduce_scatter(float* a, int lane, float* dst) {
    if constexpr (V == 32) { p_halve_step<32, 16>(a, lane); p_halve_step<16, 8>(a, lane); p_halve_step<8, 4>(a, lane); p_halve_step<4, 2>(a, lane); p_halve_step<2, 1>(a, lane); }
    if constexpr (V == 16) { p_halve_step<16, 16>(a, lane); p_halve_step<8, 8>(a, lane); p_halve_step<4, 4>(a, lane); p_halve_step<2, 2>(a, lane); }
    if constexpr (V == 8) { p_halve_step<8, 16>(a, lane); p_halve_step<4, 8>(a, lane); p_halve_step<2, 4>(a, lane); }
    if constexpr (V == 4) { p_halve_step<4, 16>(a, lane); p_halve_step<2, 8>(a, lane); }
    if constexpr (V == 2) { p_halve_step<2, 16>(a, lane); }
#pragma unroll
    for (int off = 16 / V; off > 0; off >>= 1) a[0] += __shfl_xor_sync(0xffffffffu, a[0], off);
    if ((lane & (32 / V - 1)) == 0) dst[lane / (32 / V)] += a[0];
}
template <int V>
__device__ __forceinline__ void p_zero_then_reduce(float* a, int lane, float* dst) {
    dst[lane] = 0.f;
    __syncwarp();
    float tmp[V];
#pragma unroll
    for (int i = 0; i < V; ++i) tmp[i] = a[i];
    p_reduce_scatter<V>(tmp, lane, dst);
    __syncwarp();
}
// Per-thread channel-pair sums (ps: sum, pq: sum of squares; N_TILE/2 each) -> warp totals (halving butterfly) ->
// shared memory [warp][2][N_TILE/2] -> one fp64 atomic per (group, component), groups = runs of cpg/2 pairs.
template <int N_TILE>
__device__ __forceinline__ void p_flush_pair_stats(float* ps, float* pq, float* wstat, const PairArgs& args, int n,
                                                   int ew, int lane) {
    constexpr int P = N_TILE / 2;                 // 32 (N_TILE = 64) or 8 (N_TILE = 16)
    float* mine = wstat + ew * 64;
    p_zero_then_reduce<P>(ps, lane, mine);
    p_zero_then_reduce<P>(pq, lane, mine + 32);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int e = ew * 32 + lane;
    if (e < 2 * args.gn_groups) {
        const int g = e >> 1, comp = e & 1, ppg = args.gn_cpg >> 1;
        double v = 0.0;
        for (int wq = 0; wq < 4; ++wq)
            for (int j = 0; j < ppg; ++j) v += (double)wstat[wq * 64 + comp * 32 + g * ppg + j];
        double* dst = args.gn_stats +
                      (((long long)n * FCWDM_GN_STAT_REPLICAS + (blockIdx.x % FCWDM_GN_STAT_REPLICAS)) * args.gn_groups) * 2;
        atomicAdd(dst + e, v);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
    for (int i = 0; i < P; ++i) ps[i] = pq[i] = 0.f;
}

__device__ __forceinline__ float p_silu(float x) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    return x * fmaf(0.5f, t, 0.5f);
}

constexpr int kPWarpProdA = 4, kPWarpProdB = 5, kPWarpAlloc = 6, kPWarpMma = 7;

// Transform (fused GroupNorm + SiLU) warps of the GN_IN variant.  A plane costs them ~1.6 us with 4 warps: hidden behind the
// 36 N = 192 MMAs of a 64-channel layer (1.8 us), but twice the ~0.9 us the 36 N = 48 MMAs of the output conv (C_out = 8,
// N_TILE = 16) take -- that layer was transform-bound (104 us for 28 GFLOP); its small epilogue leaves the registers for 8.
template <int N_TILE, bool GN_IN>
struct PairXf {
    static constexpr int WARPS = GN_IN ? (N_TILE == 16 ? 8 : 4) : 0;
    static constexpr int THREADS = WARPS * 32;
};

template <int N_TILE, bool GN_IN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256 + PairXf<N_TILE, GN_IN>::THREADS, 1)
    conv3d_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                       const PairArgs args) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    using Cfg = PairCfg<N_TILE>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_b = smem_base;
    const uint32_t smem_a = smem_b + Cfg::B_BYTES;
    const uint32_t bars = smem_a + Cfg::A_SLOTS * Cfg::SLOT_BYTES;
    const uint32_t full_a = bars;                                   // [A_SLOTS]  (used in the leader)
    const uint32_t empty_a = full_a + 8 * Cfg::A_SLOTS;             // [A_SLOTS]  (each CTA, multicast commit)
    const uint32_t acc_full = empty_a + 8 * Cfg::A_SLOTS;           // [RING]     (each CTA, multicast commit)
    const uint32_t acc_empty = acc_full + 8 * Cfg::RING;            // [RING]     (leader; 8 remote warp arrivals)
    const uint32_t b_full = acc_empty + 8 * Cfg::RING;              // [1]        (leader)
    const uint32_t landed_a = b_full + 8;                           // [A_SLOTS]  (GN_IN: this CTA's raw plane landed)
    const uint32_t tmem_slot = landed_a + 8 * Cfg::A_SLOTS;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    float* wstat = reinterpret_cast<float*>(smem_raw + (bars + 1024 - smem_u32(smem_raw)));
    float* sbias = reinterpret_cast<float*>(smem_raw + (bars + 512 - smem_u32(smem_raw)));   // [N_TILE]
    float* sgn = reinterpret_cast<float*>(smem_raw + (bars + 2048 - smem_u32(smem_raw)));    // [2][64] GN_IN scale / shift

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < Cfg::A_SLOTS; ++i) {
            mbar_init(full_a + 8 * i, GN_IN ? 2 * PairXf<N_TILE, GN_IN>::WARPS : 1);   // GN_IN: transform warps x 2 CTAs arrive remotely
            mbar_init(empty_a + 8 * i, 1);
            mbar_init(landed_a + 8 * i, 1);
        }
        for (int i = 0; i < Cfg::RING; ++i) {
            mbar_init(acc_full + 8 * i, 1);
            mbar_init(acc_empty + 8 * i, 8);
        }
        mbar_init(b_full, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == kPWarpProdA && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
    }
    if (warp == kPWarpAlloc) tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // both CTAs' barriers are initialised before anybody signals across the pair
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // Register rebalancing (GN_IN, 64 channels: 384 threads are capped at 168 registers each): the four single-lane control
    // warps hand most of theirs to the epilogue warps, which then hold a whole residual row one plane ahead.
    // Compile-time experiments, OFF by default (measured on B200, same-box A/B, round 2): giving the epilogue warps the
    // control warps' registers with setmaxnreg (232 / 96 / 168 per thread) and holding a whole residual row one plane ahead
    // made BOTH in-step forms ~18 us slower (141 -> 160 us, 163 -> 183 us): the MMA-issuing warp spills below 168 registers.
#ifdef FCWDM_PAIR_REBALANCE_INC
    constexpr bool REBALANCE = GN_IN && N_TILE == 64;
#else
    constexpr bool REBALANCE = false;
#define FCWDM_PAIR_REBALANCE_INC 232
#define FCWDM_PAIR_REBALANCE_DEC 96
#endif
#ifdef FCWDM_PAIR_RES_AHEAD
    constexpr bool RES_AHEAD = N_TILE == 64;
#else
    constexpr bool RES_AHEAD = false;
#endif

    if (warp >= 4 && warp < 8) {
      // the control warpgroup (one ptxas region, so the smaller register budget applies to exactly this code)
      if constexpr (REBALANCE) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FCWDM_PAIR_REBALANCE_DEC));
      if (warp == kPWarpProdB) {
        // ============ resident weights: this CTA's half (NF/2 rows) of all 9 filter columns, loaded once ============
        if (lane == 0) {
            if (leader) mbar_arrive_expect_tx(b_full, 2 * Cfg::B_BYTES);
            for (int tap = 0; tap < 9; ++tap)
                tma_load_3d_2sm(smem_b + tap * Cfg::B_TAP_BYTES, &map_b, b_full, 0, (int)rank * (Cfg::NF / 2), tap);
        }
    } else if (warp == kPWarpProdA) {
        // ============ A producer: this CTA's halo planes, in the same order and slots as the peer's ============
        // plain variant: completion bytes go straight to the leader's full_a barrier (MMA may start);
        // GN_IN variant: they go to this CTA's own landed_a barrier; the transform warps below take it from there.
        if (lane == 0) {
            uint32_t J = 0;
            for (int item = cluster_id; item < args.num_items; item += num_clusters) {
                const PairItem it = decode_item(item, args, (int)rank);
                for (int k = 0; k < it.L + 2; ++k, ++J) {
                    const uint32_t slot = J % Cfg::A_SLOTS, use = J / Cfg::A_SLOTS;
                    mbar_wait(empty_a + 8 * slot, (use & 1) ^ 1);
                    if (GN_IN) {
                        mbar_arrive_expect_tx(landed_a + 8 * slot, Cfg::PLANE_BYTES);
                        tma_load_5d(smem_a + slot * Cfg::SLOT_BYTES, &map_a, landed_a + 8 * slot, 0, it.w0 - 1, it.h0 - 1,
                                    it.d_begin - 1 + k, it.n);
                    } else {
                        if (leader) mbar_arrive_expect_tx(full_a + 8 * slot, 2 * Cfg::PLANE_BYTES);
                        tma_load_5d_2sm(smem_a + slot * Cfg::SLOT_BYTES, &map_a, full_a + 8 * slot, 0, it.w0 - 1,
                                        it.h0 - 1, it.d_begin - 1 + k, it.n);
                    }
                }
            }
        }
      } else if (warp == kPWarpMma) {
        // ============ MMA issuer (leader CTA only) ============
        if (leader) {
            // M = 256 (m_dim = 16), N = NF, bf16 x bf16 -> f32, K-major A and B
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Cfg::NF >> 3) << 17) |
                                       ((uint32_t)(256 >> 4) << 24);
            const uint64_t a_desc_base = make_sw128_desc(smem_a, Cfg::ROWP * 128);
            const uint64_t b_desc_base = make_sw128_desc(smem_b, 1024);
            mbar_wait(b_full, 0);
            tc_fence_after();
            uint32_t J = 0, G = 0;           // running input-plane / output-plane counters
            for (int item = cluster_id; item < args.num_items; item += num_clusters) {
                const PairItem it = decode_item(item, args, 0);
                for (int k = 0; k < it.L + 2; ++k, ++J) {
                    // outputs touched: item-local i = k+2 (kd 0), k+1 (kd 1), k (kd 2); first touch of i = k+2
                    // (and of i = 0, 1 at k = 0): their TMEM slots must have been zeroed by the epilogues
                    if (k == 0) {
                        mbar_wait(acc_empty + 8 * (G % Cfg::RING), (G / Cfg::RING) & 1);
                        mbar_wait(acc_empty + 8 * ((G + 1) % Cfg::RING), ((G + 1) / Cfg::RING) & 1);
                    }
                    const uint32_t Gt = G + k + 2;
                    mbar_wait(acc_empty + 8 * (Gt % Cfg::RING), (Gt / Cfg::RING) & 1);
                    const uint32_t slot = J % Cfg::A_SLOTS;
                    mbar_wait(full_a + 8 * slot, (J / Cfg::A_SLOTS) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t v = 5 - (Gt % 6);                    // lowest physical slot of the triple
                        const uint32_t d_addr = tmem_base + v * N_TILE;
                        const uint64_t a_plane = a_desc_base + (uint64_t)((slot * Cfg::SLOT_BYTES) >> 4);
                        if (args.Cin > 32) {
#pragma unroll
                            for (int tap = 0; tap < 9; ++tap) {
                                const uint64_t ad = a_plane + (uint64_t)((((tap / 3) * Cfg::ROWP + (tap % 3)) * 128) >> 4);
                                const uint64_t bd = b_desc_base + (uint64_t)((tap * Cfg::B_TAP_BYTES) >> 4);
                                umma_bf16_2sm(d_addr, ad, bd, idesc, 1u);
                                umma_bf16_2sm(d_addr, ad + 2, bd + 2, idesc, 1u);
                                umma_bf16_2sm(d_addr, ad + 4, bd + 4, idesc, 1u);
                                umma_bf16_2sm(d_addr, ad + 6, bd + 6, idesc, 1u);
                            }
                        } else {              // C_in <= 32 (the stem conv): channels 32..63 are padding, skip their K steps
#pragma unroll
                            for (int tap = 0; tap < 9; ++tap) {
                                const uint64_t ad = a_plane + (uint64_t)((((tap / 3) * Cfg::ROWP + (tap % 3)) * 128) >> 4);
                                const uint64_t bd = b_desc_base + (uint64_t)((tap * Cfg::B_TAP_BYTES) >> 4);
                                umma_bf16_2sm(d_addr, ad, bd, idesc, 1u);
                                umma_bf16_2sm(d_addr, ad + 2, bd + 2, idesc, 1u);
                            }
                        }
                        umma_commit_2sm(empty_a + 8 * slot);                        // plane slot free (both CTAs)
                        const uint32_t Gc = G + k;                                  // output i = k is complete
                        umma_commit_2sm(acc_full + 8 * (Gc % Cfg::RING));
                        if (k == it.L + 1) {                                        // segment end: flush the two trailing dummies
                            umma_commit_2sm(acc_full + 8 * ((Gc + 1) % Cfg::RING));
                            umma_commit_2sm(acc_full + 8 * ((Gc + 2) % Cfg::RING));
                        }
                    }
                    __syncwarp();
                }
                G += it.L + 4;
            }
        }
      }
    } else if (GN_IN && warp >= 8) {
        // ============ A producers with fused GroupNorm + SiLU (4 warps): global -> registers -> normalise, activate ->
        // swizzled shared memory (the layout TMA SWIZZLE_128B would have produced).  Out-of-range halo voxels are
        // written as ZERO: the convolution pads the ACTIVATED tensor.  ============
        constexpr int XFT = GN_IN ? PairXf<N_TILE, GN_IN>::THREADS : 128;   // 128, or 256 for the output conv (branch dead without GN_IN)
        const int pt = threadIdx.x - 256;                       // 0..XFT-1
        const uint32_t full_a_leader = mapa_u32(full_a, 0);
        constexpr int CHUNKS = Cfg::HROWS * Cfg::ROWP * 8;       // 16-byte chunks per plane (1440)
        constexpr int PER_THREAD = (CHUNKS + XFT - 1) / XFT;     // 12 (6)
        const int jmine = (pt & 7) ^ ((pt >> 3) & 7);
        float sc[8], sh[8];
        int cur_n = -1;
        uint32_t J = 0;
        for (int item = cluster_id; item < args.num_items; item += num_clusters) {
            const PairItem it = decode_item(item, args, (int)rank);
            if (it.n != cur_n) {
                cur_n = it.n;
                asm volatile("bar.sync 2, %0;" ::"n"(XFT) : "memory");
                if (pt < 64) {
                    float sc0 = 0.f, sh0 = 0.f;
                    if (pt < args.Cin) {
                        const int cpg = args.Cin / args.gi_groups;
                        const int g = pt / cpg;
                        double sum = 0.0, sq = 0.0;
                        for (int r = 0; r < FCWDM_GN_STAT_REPLICAS; ++r) {
                            const double* sp = args.gi_stats + (((long long)it.n * FCWDM_GN_STAT_REPLICAS + r) * args.gi_groups + g) * 2;
                            sum += sp[0];
                            sq += sp[1];
                        }
                        const double cnt = (double)args.D * args.H * args.W * cpg;
                        const double mean = sum / cnt;
                        double var = sq / cnt - mean * mean;
                        var = var < 0.0 ? 0.0 : var;
                        const float rstd = (float)(1.0 / sqrt(var + (double)args.gi_eps));
                        sc0 = rstd * __ldg(args.gi_gamma + pt);
                        sh0 = __ldg(args.gi_beta + pt) - (float)mean * sc0;
                    }
                    sgn[pt] = sc0;
                    sgn[64 + pt] = sh0;
                }
                asm volatile("bar.sync 2, %0;" ::"n"(XFT) : "memory");
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    sc[e] = sgn[jmine * 8 + e];
                    sh[e] = sgn[64 + jmine * 8 + e];
                }
            }
            for (int k = 0; k < it.L + 2; ++k, ++J) {
                const uint32_t slot = J % Cfg::A_SLOTS, use = J / Cfg::A_SLOTS;
                const int d = it.d_begin - 1 + k;
                const bool d_ok = (d >= 0) && (d < args.D);
                mbar_wait(landed_a + 8 * slot, use & 1);          // TMA has written the raw plane (zeros out of range)
                const uint32_t base = smem_a + slot * Cfg::SLOT_BYTES;
                if (d_ok) {
                    // all shared-memory loads first (independent, pipelined), then the arithmetic and the stores:
                    // a load -> compute -> store chain per chunk is latency-bound (~450 cycles per chunk)
                    uint4* plane = reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)));
                    uint4 raw[PER_THREAD];
                    uint32_t valid = 0;
#pragma unroll
                    for (int q = 0; q < PER_THREAD; ++q) {
                        const int c = pt + q * XFT;
                        const int r = c >> 3;
                        const int hr = r / Cfg::ROWP, wc = r - hr * Cfg::ROWP;
                        const int h = it.h0 - 1 + hr, w = it.w0 - 1 + wc;
                        // out-of-range halo voxels stay ZERO: the convolution pads the ACTIVATED tensor
                        const bool in = (c < CHUNKS) && (h >= 0) && (h < args.H) && (w >= 0) && (w < args.W);
                        raw[q] = make_uint4(0u, 0u, 0u, 0u);
                        if (in) {
                            raw[q] = plane[c];
                            valid |= 1u << q;
                        }
                    }
                    // this thread's logical 16-byte chunk index is the same for all of its chunks
                    // (c = pt + XFT q, XFT a multiple of 64  =>  (c & 7) ^ ((c >> 3) & 7) does not depend on q): scale / shift live in registers
#pragma unroll
                    for (int q = 0; q < PER_THREAD; ++q) {
                        float f[8];
                        unpack8(raw[q], f);
#pragma unroll
                        for (int e = 0; e < 8; ++e) f[e] = p_silu(fmaf(f[e], sc[e], sh[e]));
                        if (valid & (1u << q)) plane[pt + q * XFT] = pack8(f);
                    }
                }
                fence_proxy_async();                 // generic-proxy smem writes -> visible to the tensor-core (async) proxy
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(full_a_leader + 8 * slot);
            }
        }
    } else if (warp < 4) {
        // ============ epilogue: this CTA's 128 accumulator lanes ============
        if constexpr (REBALANCE) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(FCWDM_PAIR_REBALANCE_INC));
        const int ew = warp;
        const int row = ew * 32 + lane;
        const int hh = row >> 3, ww = row & 7;
        const uint32_t lane_base = tmem_base + ((uint32_t)(ew * 32) << 16);
        const uint32_t acc_empty_leader = mapa_u32(acc_empty, 0);
        // Fused GroupNorm statistics of the output: per-thread fp32 sums over CHANNEL PAIRS (exact for any even number
        // of channels per group) kept in registers across all planes of the sample, reduced across lanes only when
        // the sample index changes / at exit -- the per-plane cost is 2 FADD/FFMA per channel.
        const bool want_stats = args.gn_stats != nullptr;
        float ps[N_TILE / 2], pq[N_TILE / 2];
#pragma unroll
        for (int e = 0; e < N_TILE / 2; ++e) ps[e] = pq[e] = 0.f;
        // initial state: all accumulator slots zero, every ring entry "empty" (phase 0 of acc_empty)
        for (int c = 0; c < 8 * N_TILE; c += 16) tmem_st_zero_x16(lane_base + c);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0)
            for (int r = 0; r < Cfg::RING; ++r) mbar_arrive_remote(acc_empty_leader + 8 * r);
        int cur_n = -1;
        uint32_t G = 0;
        constexpr int HALF = N_TILE < 32 ? N_TILE : 32;          // accumulator columns drained per batch of TMEM loads
        uint4 res_next[N_TILE / 8];                              // RES_AHEAD: the next plane's residual row
#pragma unroll
        for (int g = 0; g < N_TILE / 8; ++g) res_next[g] = make_uint4(0u, 0u, 0u, 0u);
        for (int item = cluster_id; item < args.num_items; item += num_clusters) {
            const PairItem it = decode_item(item, args, (int)rank);
            if (it.n != cur_n) {
                if (want_stats && cur_n >= 0) p_flush_pair_stats<N_TILE>(ps, pq, wstat, args, cur_n, ew, lane);
                cur_n = it.n;
                // per-sample additive term bias[c] + chan_bias[n][c], staged once in shared memory (broadcast reads)
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (row < N_TILE) {
                    float bv = 0.f;
                    if (row < args.Cout) {
                        if (args.bias != nullptr) bv += __ldg(args.bias + row);
                        if (args.chan_bias != nullptr) bv += __ldg(args.chan_bias + (long long)it.n * args.cb_ld + row);
                    }
                    sbias[row] = bv;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            const int h = it.h0 + hh, w = it.w0 + ww;
            const bool hw_ok = (h < args.H) && (w < args.W);
            for (int i = 0; i < it.L + 4; ++i, ++G) {
                const uint32_t ring = G % Cfg::RING;
                const uint32_t v = 5 - (G % 6);
                const uint32_t main_col = v * N_TILE;
                const bool split = v < 2;                                 // slots 6 / 7 hold the other half
                const uint32_t extra_col = (6 + v) * N_TILE;
                const int d = it.d_begin - 2 + i;
                const bool real = (i >= 2) && (i < it.L + 2);             // dummy planes are drained and dropped
                const bool ok = real && hw_ok;
                const bool use_res = ok && args.residual != nullptr;
                const long long vox = (((long long)it.n * args.D + d) * args.H + h) * args.W + w;
                // the residual row (one 128-byte line per thread) of a LATER plane is pulled into L2 now, so that its loads
                // are L2 hits instead of HBM round trips (the epilogue has ~1.7 us per plane)
                constexpr int PF = RES_AHEAD ? 2 : 1;
                if (args.residual != nullptr && hw_ok && d + PF >= 0 && d + PF < args.D)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(args.residual + (vox + (long long)PF * args.H * args.W) * args.res_ld));
                uint4 res[RES_AHEAD ? N_TILE / 8 : HALF / 8];
                if constexpr (RES_AHEAD) {
                    // whole rows, one plane ahead: this plane's row was requested a plane ago (or right here for the first
                    // plane of an item), the next plane's row goes out now and lands while this plane is drained and stored
                    if (use_res && i == 2) {
#pragma unroll
                        for (int g = 0; g < N_TILE / 8; ++g)
                            if (g * 8 < args.Cout) res_next[g] = *reinterpret_cast<const uint4*>(args.residual + vox * args.res_ld + g * 8);
                    }
#pragma unroll
                    for (int g = 0; g < N_TILE / 8; ++g) res[g] = res_next[g];
                    if (args.residual != nullptr && hw_ok && i + 1 >= 2 && i + 1 < it.L + 2) {
                        const __nv_bfloat16* rn = args.residual + (vox + (long long)args.H * args.W) * args.res_ld;
#pragma unroll
                        for (int g = 0; g < N_TILE / 8; ++g)
                            if (g * 8 < args.Cout) res_next[g] = *reinterpret_cast<const uint4*>(rn + g * 8);
                    }
                } else if (use_res) {
                    // first half of the residual row: in flight while we wait for the accumulator
#pragma unroll
                    for (int g = 0; g < HALF / 8; ++g)
                        if (g * 8 < args.Cout) res[g] = *reinterpret_cast<const uint4*>(args.residual + vox * args.res_ld + g * 8);
                }
                mbar_wait(acc_full + 8 * ring, (G / Cfg::RING) & 1);
                tc_fence_after();
#pragma unroll
                for (int c0 = 0; c0 < N_TILE; c0 += HALF) {
                    // drain HALF fp32 columns (both physical halves of a split plane) with all TMEM loads in flight,
                    // zero them for the next use of the slot
                    uint32_t acc[HALF];
#pragma unroll
                    for (int q = 0; q < HALF; q += 16) tmem_ld_x16(lane_base + main_col + c0 + q, acc + q);
                    if (split) {
                        uint32_t acc2[HALF];
#pragma unroll
                        for (int q = 0; q < HALF; q += 16) tmem_ld_x16(lane_base + extra_col + c0 + q, acc2 + q);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < HALF; ++e)
                            acc[e] = __float_as_uint(__uint_as_float(acc[e]) + __uint_as_float(acc2[e]));
#pragma unroll
                        for (int q = 0; q < HALF; q += 16) tmem_st_zero_x16(lane_base + extra_col + c0 + q);
                    } else {
                        tmem_ld_wait();
                    }
#pragma unroll
                    for (int q = 0; q < HALF; q += 16) tmem_st_zero_x16(lane_base + main_col + c0 + q);
                    if (c0 + HALF >= N_TILE) {                            // whole row drained: slots free for the MMA again
                        tmem_st_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_remote(acc_empty_leader + 8 * ring);
                    }
                    if (real) {                                           // warp-uniform
#pragma unroll
                        for (int g = 0; g < HALF / 8; ++g) {
                            const int co = c0 + g * 8;
                            if (co < args.Cout) {
                                float vv[8];
                                const float4 b0 = *reinterpret_cast<const float4*>(sbias + co);
                                const float4 b1 = *reinterpret_cast<const float4*>(sbias + co + 4);
                                vv[0] = __uint_as_float(acc[g * 8 + 0]) + b0.x; vv[1] = __uint_as_float(acc[g * 8 + 1]) + b0.y;
                                vv[2] = __uint_as_float(acc[g * 8 + 2]) + b0.z; vv[3] = __uint_as_float(acc[g * 8 + 3]) + b0.w;
                                vv[4] = __uint_as_float(acc[g * 8 + 4]) + b1.x; vv[5] = __uint_as_float(acc[g * 8 + 5]) + b1.y;
                                vv[6] = __uint_as_float(acc[g * 8 + 6]) + b1.z; vv[7] = __uint_as_float(acc[g * 8 + 7]) + b1.w;
                                if (use_res) {
                                    float rr[8];
                                    unpack8(res[RES_AHEAD ? (c0 / 8 + g) : g], rr);
#pragma unroll
                                    for (int e = 0; e < 8; ++e) vv[e] += rr[e];
                                }
                                const uint4 packed = pack8(vv);
                                if (ok) *reinterpret_cast<uint4*>(args.y + vox * args.y_ld + co) = packed;
                                if (want_stats && ok) {                   // statistics of the STORED (bf16) values
                                    float vr[8];
                                    unpack8(packed, vr);
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        ps[(co >> 1) + e] += vr[2 * e] + vr[2 * e + 1];
                                        pq[(co >> 1) + e] = fmaf(vr[2 * e], vr[2 * e], fmaf(vr[2 * e + 1], vr[2 * e + 1], pq[(co >> 1) + e]));
                                    }
                                }
                            }
                        }
                        if (!RES_AHEAD && use_res && c0 + HALF < N_TILE) {   // next half of the residual row
#pragma unroll
                            for (int g = 0; g < HALF / 8; ++g)
                                if (c0 + HALF + g * 8 < args.Cout)
                                    res[g] = *reinterpret_cast<const uint4*>(args.residual + vox * args.res_ld + c0 + HALF + g * 8);
                        }
                    }
                }
            }
        }
        if (want_stats && cur_n >= 0) p_flush_pair_stats<N_TILE>(ps, pq, wstat, args, cur_n, ew, lane);
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // the peer may still be reading TMEM / signalling our barriers
    if (warp == kPWarpAlloc) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
    }
}

// weights (Cout, Cin, 3,3,3) f32 -> [kh*3+kw][kd][Cout_p][64] bf16 (Cin <= 64 zero padded, Cout_p = 16 or 64)
__global__ void __launch_bounds__(256) pack_weights_pair_kernel(const float* __restrict__ w,
                                                                __nv_bfloat16* __restrict__ wp, int Cout, int Cin,
                                                                int Cout_p) {
    pdl_prologue();
    const long long total = 27ll * Cout_p * 64;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int ci = (int)(idx % 64);
    long long r = idx / 64;
    const int co = (int)(r % Cout_p); r /= Cout_p;
    const int kd = (int)(r % 3);
    const int tap2 = (int)(r / 3);                   // kh*3 + kw
    float v = 0.f;
    if (co < Cout && ci < Cin) v = w[((long long)co * Cin + ci) * 27 + kd * 9 + tap2];
    wp[idx] = __float2bfloat16_rn(v);
}

typedef CUresult (*EncodeTiledFnP)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnP g_encode_p = nullptr;

int conv3d_pair_init_device() {
    if (g_encode_p == nullptr) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        FCWDM_REQUIRE(e == cudaSuccess && fn != nullptr && qres == cudaDriverEntryPointSuccess, FCWDM_ERR_CUDA,
                      "fcwdm_init: cuTensorMapEncodeTiled entry point not available (%s)", cudaGetErrorString(e));
        g_encode_p = reinterpret_cast<EncodeTiledFnP>(fn);
    }
    cudaError_t e = cudaFuncSetAttribute(conv3d_pair_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         PairCfg<64>::SMEM_BYTES);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(conv3d_pair_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 PairCfg<16>::SMEM_BYTES);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(conv3d_pair_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 PairCfg<64>::SMEM_BYTES);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(conv3d_pair_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 PairCfg<16>::SMEM_BYTES);
    FCWDM_REQUIRE(e == cudaSuccess, FCWDM_ERR_CUDA, "fcwdm_init: cudaFuncSetAttribute(pair) failed: %s",
                  cudaGetErrorString(e));
    return FCWDM_OK;
}

template <int N_TILE, bool GN_IN>
static int launch_pair(const CUtensorMap& ma, const CUtensorMap& mb, const PairArgs& a, cudaStream_t st) {
    const int clusters = num_sms() / 2;
    const int grid = 2 * (a.num_items < clusters ? a.num_items : clusters);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(256 + PairXf<N_TILE, GN_IN>::THREADS);
    cfg.dynamicSmemBytes = PairCfg<N_TILE>::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;            // the cluster shape is static (__cluster_dims__)
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv3d_pair_kernel<N_TILE, GN_IN>, ma, mb, a);
    FCWDM_REQUIRE(e == cudaSuccess, FCWDM_ERR_CUDA, "fcwdm_conv3d_pair_fwd: launch failed: %s", cudaGetErrorString(e));
    FCWDM_CHECK_LAUNCH("fcwdm_conv3d_pair_fwd");
    return FCWDM_OK;
}

}  // namespace fcwdm

using namespace fcwdm;

extern "C" int fcwdm_conv3d_pair_supported(int64_t Cin, int64_t Cout, int ksize) {
    return (ksize == 3 && Cin > 0 && Cin <= 64 && Cout > 0 && Cout <= 64 && Cout % 8 == 0) ? 1 : 0;
}

extern "C" int64_t fcwdm_conv3d_pair_packed_elems(int64_t Cout, int64_t Cin) {
    if (!fcwdm_conv3d_pair_supported(Cin, Cout, 3)) return -1;
    return 27 * (Cout <= 16 ? 16 : 64) * 64;
}

extern "C" int fcwdm_conv3d_pair_pack_weights(const float* w, void* wp, int64_t Cout, int64_t Cin, void* stream) {
    FCWDM_REQUIRE(w && wp, FCWDM_ERR_INVALID, "fcwdm_conv3d_pair_pack_weights: null pointer");
    FCWDM_REQUIRE(fcwdm_conv3d_pair_supported(Cin, Cout, 3), FCWDM_ERR_UNSUPPORTED,
                  "fcwdm_conv3d_pair_pack_weights: needs C_in <= 64, C_out <= 64, C_out %% 8 == 0");
    const int cout_p = Cout <= 16 ? 16 : 64;
    const long long total = 27ll * cout_p * 64;
    launch_k(pack_weights_pair_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, w,
             (__nv_bfloat16*)wp, (int)Cout, (int)Cin, cout_p);
    FCWDM_CHECK_LAUNCH("fcwdm_conv3d_pair_pack_weights");
    return FCWDM_OK;
}

extern "C" int fcwdm_conv3d_pair_fwd(const void* x, int64_t x_ld, const void* wp, const float* bias,
                                     const float* chan_bias, int64_t cb_ld, const void* residual, int64_t res_ld, void* y,
                                     int64_t y_ld, double* gn_stats, int64_t gn_groups, const double* gn_in_stats,
                                     const float* gn_in_gamma, const float* gn_in_beta, int64_t gn_in_groups,
                                     float gn_in_eps, int64_t N, int64_t D, int64_t H, int64_t W, int64_t Cin,
                                     int64_t Cout, void* stream) {
    FCWDM_REQUIRE(x && wp && y, FCWDM_ERR_INVALID, "fcwdm_conv3d_pair_fwd: null pointer");
    FCWDM_REQUIRE(fcwdm_conv3d_pair_supported(Cin, Cout, 3), FCWDM_ERR_UNSUPPORTED,
                  "fcwdm_conv3d_pair_fwd: needs C_in <= 64, C_out <= 64, C_out %% 8 == 0 (3x3x3)");
    FCWDM_REQUIRE(N >= 0 && D >= 0 && H >= 0 && W >= 0 && D < 32768 && H < 32768 && W < 32768 && N < 32768,
                  FCWDM_ERR_INVALID, "fcwdm_conv3d_pair_fwd: bad dimension");
    FCWDM_REQUIRE(x_ld >= 64 && x_ld % 8 == 0 && y_ld >= Cout && y_ld % 8 == 0 &&
                      (residual == nullptr || (res_ld >= Cout && res_ld % 8 == 0)) && cb_ld % 4 == 0,
                  FCWDM_ERR_INVALID, "fcwdm_conv3d_pair_fwd: bad leading dimension");
    FCWDM_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)wp % 16 == 0) && ((uintptr_t)y % 16 == 0) &&
                      ((uintptr_t)residual % 16 == 0) && ((uintptr_t)bias % 16 == 0) && ((uintptr_t)chan_bias % 16 == 0),
                  FCWDM_ERR_INVALID, "fcwdm_conv3d_pair_fwd: pointers must be 16-byte aligned");
    if (gn_stats != nullptr) {
        FCWDM_REQUIRE(gn_groups > 0 && gn_groups <= 32 && Cout % gn_groups == 0, FCWDM_ERR_UNSUPPORTED,
                      "fcwdm_conv3d_pair_fwd: fused GroupNorm statistics need 1 <= groups <= 32 dividing C_out");
        const int64_t cpg = Cout / gn_groups;
        FCWDM_REQUIRE(cpg % 2 == 0, FCWDM_ERR_UNSUPPORTED,
                      "fcwdm_conv3d_pair_fwd: fused statistics need an even number of channels per group");
    }
    if (N * D * H * W == 0) return FCWDM_OK;
    if (g_encode_p == nullptr) {
        int dev = 0;
        cudaGetDevice(&dev);
        int rc = fcwdm_init(dev);
        if (rc) return rc;
    }
    const int n_tile = Cout <= 16 ? 16 : 64;
    CUtensorMap ma, mb;
    {
        cuuint64_t dims[5] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
        cuuint64_t strides[4] = {(cuuint64_t)x_ld * 2, (cuuint64_t)W * x_ld * 2, (cuuint64_t)H * W * x_ld * 2,
                                 (cuuint64_t)D * H * W * x_ld * 2};
        cuuint32_t box[5] = {64, 10, 18, 1, 1};
        cuuint32_t es[5] = {1, 1, 1, 1, 1};
        CUresult r = g_encode_p(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides, box, es,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        FCWDM_REQUIRE(r == CUDA_SUCCESS, FCWDM_ERR_CUDA, "fcwdm_conv3d_pair_fwd: activation tensor map failed (%d)", (int)r);
    }
    {
        cuuint64_t dims[3] = {64, (cuuint64_t)(3 * n_tile), 9};
        cuuint64_t strides[2] = {128, (cuuint64_t)(3 * n_tile) * 128};
        cuuint32_t box[3] = {64, (cuuint32_t)(3 * n_tile / 2), 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = g_encode_p(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(wp), dims, strides, box, es,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        FCWDM_REQUIRE(r == CUDA_SUCCESS, FCWDM_ERR_CUDA, "fcwdm_conv3d_pair_fwd: weight tensor map failed (%d)", (int)r);
    }
    PairArgs a;
    a.N = (int)N; a.D = (int)D; a.H = (int)H; a.W = (int)W;
    a.Cout = (int)Cout;
    // pair orientation: fewer wasted (fully out-of-range) CTA tiles
    const long long th = (H + 15) / 16, tw = (W + 7) / 8;
    const long long cost_w = th * (2 * ((tw + 1) / 2)), cost_h = (2 * ((th + 1) / 2)) * tw;
    a.pair_w = cost_w <= cost_h ? 1 : 0;
    a.n_hu = a.pair_w ? (int)th : (int)((th + 1) / 2);
    a.n_wu = a.pair_w ? (int)((tw + 1) / 2) : (int)tw;
    // depth segmentation: minimise rounds x (planes per segment + 2 halo planes)
    const long long columns = N * a.n_hu * a.n_wu;
    const long long clusters = num_sms() / 2;
    long long best_cost = -1;
    int best_seg = 1;
    for (int ns = 1; ns <= D && ns <= 64; ++ns) {
        const long long len = (D + ns - 1) / ns;
        if ((long long)(ns - 1) * len >= D) continue;      // empty trailing segment
        const long long rounds = (columns * ns + clusters - 1) / clusters;
        const long long cost = rounds * (len + 2);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_seg = ns; }
    }
    a.n_seg = best_seg;
    a.seg_len = (int)((D + best_seg - 1) / best_seg);
    const long long items = columns * a.n_seg;
    FCWDM_REQUIRE(items < (1ll << 31), FCWDM_ERR_UNSUPPORTED, "fcwdm_conv3d_pair_fwd: too many work items");
    a.num_items = (int)items;
    a.bias = bias; a.chan_bias = chan_bias; a.cb_ld = cb_ld;
    a.residual = (const __nv_bfloat16*)residual; a.res_ld = res_ld;
    a.y = (__nv_bfloat16*)y; a.y_ld = y_ld;
    a.gn_stats = gn_stats;
    a.gn_groups = gn_stats ? (int)gn_groups : 0;
    a.gn_cpg = gn_stats ? (int)(Cout / gn_groups) : 0;
    a.x = (const __nv_bfloat16*)x; a.x_ld = x_ld; a.Cin = (int)Cin;
    a.gi_stats = gn_in_stats; a.gi_gamma = gn_in_gamma; a.gi_beta = gn_in_beta;
    a.gi_groups = (int)gn_in_groups; a.gi_eps = gn_in_eps;
    cudaStream_t st = (cudaStream_t)stream;
    if (gn_in_stats != nullptr) {
        FCWDM_REQUIRE(gn_in_gamma && gn_in_beta && gn_in_groups > 0 && Cin % gn_in_groups == 0, FCWDM_ERR_INVALID,
                      "fcwdm_conv3d_pair_fwd: fused input GroupNorm needs gamma, beta and groups dividing C_in");
        return n_tile == 64 ? launch_pair<64, true>(ma, mb, a, st) : launch_pair<16, true>(ma, mb, a, st);
    }
    return n_tile == 64 ? launch_pair<64, false>(ma, mb, a, st) : launch_pair<16, false>(ma, mb, a, st);
}

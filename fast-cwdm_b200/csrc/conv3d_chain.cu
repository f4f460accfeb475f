// K5c: a RUN of consecutive low-resolution 3x3x3 conv3d layers in ONE persistent launch.
//
// Where: levels 2-3 and the bottleneck of the wavelet U-Net (guided_diffusion/wunet.py:533-609,615-675): 28x28x20,
// 14x14x10 and 7x7x5 voxels with 128-1024 input and 128/256 output channels.  One such layer has 10-160 output tiles
// for 148 SMs and its activations are L2 resident (<= 4 MB); launched one by one (conv3d.cu) each costs 17-43 us of
// which ~8 us are dependency wait, TMEM allocation, barrier set-up and pipeline fill/drain, and a tile is a serial
// chain of up to 432 (x C_in/128) MMAs on 57 % of the SMs: 29 % of a denoising step for 7 % of its FLOPs.
//
// Here a cluster-of-4 persistent grid (one CTA per SM, all co-resident) walks a LIST of layers:
//   * layer boundaries are grid barriers (one release-add on a global counter per CTA, an acquire spin in the consumers)
//     instead of kernel boundaries; TMEM, mbarriers, ring state and the weight prefetch survive across them (the weight
//     producer runs ahead into the next layer while the grid drains the current one);
//   * every layer is scheduled as full waves of one tile per CTA plus a REMAINDER wave in which each tile is shared by
//     s = 2 or 4 CTAs of a cluster (split-K over 64-channel blocks of C_in), so a layer with fewer tiles than SMs still
//     uses up to 4x as many of them and its serial MMA chain is s times shorter;
//   * the partial sums of a shared tile are REDUCE-SCATTERED through distributed shared memory: CTA r of the group
//     sends the 128/s accumulator columns owned by each peer into that peer's receive slots and finishes (bias, timestep
//     embedding, residual, GroupNorm statistics, bf16 store) only its own columns -- (s-1)/s of 64 KB per CTA instead of
//     a leader holding (s-1) x 64 KB, and the epilogue is spread over the group;
//   * the GroupNorm statistics a layer needs are accumulated by its producer's epilogue (fp64 atomics), published by the
//     grid barrier, and applied (normalise + SiLU) to the landed halo planes in shared memory by four transform warps.
// Tile = 16 x 8 voxels of one depth plane x 128 output channels (M = 128, N = 128); operand staging, UMMA descriptors
// and the row-shifted halo addressing are those of conv3d.cu.
#include <stdlib.h>
#include <string.h>

#include "conv3d_gn.cuh"
#include "tc_ptx.cuh"

namespace fcwdm {

constexpr int kChainMaxLayers = 32;
constexpr int kChainMaxCluster = 4;      // cluster size is chosen per launch (2 or 4): 4 shortens the serial MMA chain of the
                                         // smallest layers 4x, 2 keeps every SM usable (33 clusters of 4 = 132 of 148 SMs on B200)

struct ChainLayer {
    CUtensorMap map_a;       // activations (C_in_p, W, H, D, N) bf16, box 64 x 10 x 18 x 1 x 1, SWIZZLE_128B
    CUtensorMap map_b;       // weights (C_in_p, C_out_p, 27) bf16, box 64 x 128 x 3
    int N, D, H, W;
    int Cout, n_cb;
    int n_nt, n_wt, n_ht;
    int num_tiles;
    int waves_a;             // full waves of one tile per CTA
    int split_b;             // CTAs per tile in the remainder wave (1, 2 or 4; divides n_cb)
    const float* bias;
    const float* chan_bias;
    long long cb_ld;
    const __nv_bfloat16* residual;
    long long res_ld;
    __nv_bfloat16* y;
    long long y_ld;
    double* gn_stats;        // statistics of the OUTPUT (pre-zeroed) or null
    int gn_cpg, gn_groups;
    int Cin;
    const double* gi_stats;  // statistics of the INPUT: the layer convolves SiLU(GroupNorm(x)); null = plain input
    const float* gi_gamma;
    const float* gi_beta;
    int gi_groups;
    float gi_eps;
    // auxiliary (non-conv) operations of the same launch, executed by the epilogue + transform warps of every CTA:
    //   kind 1 = Haar DWT  x (N,D,H,W,Cin) -> y = LLL * lll_scale + chan_bias, aux = 7 high bands * hi_scale (or null)
    //   kind 2 = Haar IDWT x = LLL (N,D/2,H/2,W/2,Cin) * lll_scale, aux = 7 high bands -> y (N,D,H,W,Cin) + chan_bias
    // gn_stats (optional) receives the GroupNorm statistics of y in both cases
    int kind;
    __nv_bfloat16* aux;
    long long aux_ld, aux_sb;
    float lll_scale, hi_scale;
};

struct ChainParams {
    ChainLayer layers[kChainMaxLayers];
    int n_layers;
    int cluster;             // CTAs per cluster of this launch (2 or 4)
    long long* trace;        // development (-DFCWDM_CONV_TRACE): [grid][kChainMaxLayers][16] %globaltimer stamps, else null
    unsigned int* sync;      // grid-barrier counter, zero at launch; layer l is complete when it reaches (l+1) * gridDim.x
};
static_assert(sizeof(ChainParams) <= 32000, "kernel parameters exceed the 32 KB limit of CUDA >= 12.1");

struct ChainCfg {
    static constexpr int N_TILE = 128;
    static constexpr int ROWP = 10, HROWS = 18;
    static constexpr int PLANE_BYTES = HROWS * ROWP * 128;        // 23040
    static constexpr int SLOT_BYTES = 23552;                      // 1024-aligned
    static constexpr int A_SLOTS = 3;
    static constexpr int B_TAP_BYTES = N_TILE * 128;
    static constexpr int B_BYTES = 3 * B_TAP_BYTES;               // the 3 kw taps of one (kd, kh)
    static constexpr int B_STAGES = 2;
    // receive slots of the reduce-scatter: [sender][128 rows][CW + 4] fp32 (the 4-float pad keeps 128-bit row accesses
    // conflict-free); s = 4: 3 x 128 x 36, s = 2: 1 x 128 x 68
    static constexpr int RECV_BYTES = 3 * 128 * 36 * 4;
    static constexpr int XF_WARPS = 8, XF_THREADS = XF_WARPS * 32;   // operand hand-over / GroupNorm transform warps
    static constexpr int THREADS = 256 + XF_THREADS;
    static constexpr int TAIL_BYTES = 4608;                       // barriers | bias | statistics | GN scale/shift
    static constexpr int ACC_STAGES = 2;
    static constexpr int TMEM_COLS = 256;
    static constexpr int SMEM_BYTES = 1024 + A_SLOTS * SLOT_BYTES + B_STAGES * B_BYTES + RECV_BYTES + TAIL_BYTES;
    static_assert(SMEM_BYTES <= 227 * 1024, "chain kernel exceeds shared memory");
};

struct ChainWork {
    int tile, split, kr;
};
// Work item `it` of this CTA in layer L: its = 0 .. waves_a-1 are whole tiles, it = waves_a is the shared remainder wave.
__device__ __forceinline__ bool chain_work(const ChainLayer& L, int it, int crank, int csize, ChainWork& w) {
    const int G = (int)gridDim.x;
    if (it < L.waves_a) {
        w.tile = it * G + (int)blockIdx.x;
        w.split = 1;
        w.kr = 0;
        return true;
    }
    const int R = L.num_tiles - L.waves_a * G;
    const int s = L.split_b;
    const int gi = ((int)blockIdx.x / csize) * (csize / s) + crank / s;
    if (gi >= R) return false;
    w.tile = L.waves_a * G + gi;
    w.split = s;
    w.kr = crank % s;
    return true;
}

struct ChainTile {
    int n, d0, h0, w0, n0;
};
__device__ __forceinline__ ChainTile chain_decode(int tile, const ChainLayer& L) {
    ChainTile t;
    int r = tile;
    const int nt = r % L.n_nt; r /= L.n_nt;
    const int wt = r % L.n_wt; r /= L.n_wt;
    const int ht = r % L.n_ht; r /= L.n_ht;
    const int dt = r % L.D; r /= L.D;
    t.n = r;
    t.d0 = dt;
    t.h0 = ht * 16;
    t.w0 = wt * 8;
    t.n0 = nt * ChainCfg::N_TILE;
    return t;
}

// Bounded acquire spin on the grid-barrier counter: a protocol bug must surface as a trap, never as a hung GPU.
__device__ __forceinline__ void grid_wait(const unsigned int* ctr, unsigned int target) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    if (v >= target) return;
    const long long t0 = clock64();
    do {
        __nanosleep(64);
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        if (clock64() - t0 > 4000000000LL) {
            printf("fcwdm conv3d chain: grid barrier timed out (block %d thread %d have %u want %u)\n", blockIdx.x,
                   threadIdx.x, v, target);
            __trap();
        }
    } while (v < target);
}
__device__ __forceinline__ uint4 ld_cg_u4(const void* p) {       // L2-only load: data written by other SMs in THIS launch
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ double ld_cg_f64(const double* p) {
    double r;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(r) : "l"(p) : "memory");
    return r;
}

// %globaltimer stamps per (CTA, layer), compiled in only with -DFCWDM_CONV_TRACE (tools/chain_trace.py):
//  0 A producer reaches the layer | 1 grid barrier passed | 2 sgn table ready (transform warps) | 3 first plane landed
//  4 first plane handed to the MMA | 5 first MMA issued | 6 last MMA issued | 7 accumulator complete (epilogue)
//  8 partial sums exchanged | 9 tile stored | 10 statistics flushed | 11 arrived at the next grid barrier
#ifdef FCWDM_CONV_TRACE
__device__ __forceinline__ long long chain_now() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define CHAIN_TRACE(li, slot) do { if (P.trace != nullptr) P.trace[((size_t)blockIdx.x * kChainMaxLayers + (li)) * 16 + (slot)] = chain_now(); } while (0)
#else
#define CHAIN_TRACE(li, slot) do { } while (0)
#endif


// ---------------------------------------------------------------------------------------------------
// auxiliary operations (wavelet down / up-sampling between the convs of a run) -- HBM/L2-bound elementwise passes
// ---------------------------------------------------------------------------------------------------
constexpr int kAuxWorkers = 128 + ChainCfg::XF_THREADS;      // epilogue warps + transform warps of one CTA

__device__ __forceinline__ float chain_scale(float v, float scale) {
    return (scale == (1.0f / 3.0f)) ? (v / 3.0f) : (v * scale);      // LLL / 3. is a true division in the reference
}

// One Haar analysis (kind 1) or synthesis (kind 2) pass over sample n by the workers of the whole grid; wid = this
// thread's worker index in its CTA.  st8[0..7] / st8[8..15] accumulate per-channel sum / sum of squares of the STORED
// output values of this worker's fixed 8-channel chunk (the grid stride is a multiple of the chunks per voxel).
__device__ __forceinline__ void chain_aux_pass(const ChainLayer& L, int n, int wid, float* st8) {
    const int C8 = L.Cin >> 3;
    const int D2 = L.D >> 1, H2 = L.H >> 1, W2 = L.W >> 1;
    const long long items = (long long)D2 * H2 * W2 * C8;
    const long long stride = (long long)gridDim.x * kAuxWorkers;
    const bool want = L.gn_stats != nullptr;
    for (long long idx = (long long)blockIdx.x * kAuxWorkers + wid; idx < items; idx += stride) {
        const int cq = (int)(idx % C8);
        long long t = idx / C8;
        const int ww = (int)(t % W2); t /= W2;
        const int hh = (int)(t % H2);
        const int dd = (int)(t / H2);
        const long long vox2 = (((long long)n * D2 + dd) * H2 + hh) * W2 + ww;
        const long long vox0 = (((long long)n * L.D + 2 * dd) * L.H + 2 * hh) * L.W + 2 * ww;
        float bs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (L.chan_bias != nullptr) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(L.chan_bias + (long long)n * L.cb_ld + cq * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(L.chan_bias + (long long)n * L.cb_ld + cq * 8 + 4));
            bs[0] = b0.x; bs[1] = b0.y; bs[2] = b0.z; bs[3] = b0.w;
            bs[4] = b1.x; bs[5] = b1.y; bs[6] = b1.z; bs[7] = b1.w;
        }
        if (L.kind == 1) {
            const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(L.residual);    // aux ops carry their input in `residual`
            const __nv_bfloat16* src = x + vox0 * L.res_ld + cq * 8;
            float in[8][8];      // [brick position i*4+j*2+k][channel]
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        unpack8(ld_cg_u4(src + ((long long)(i * L.H + j) * L.W + k) * L.res_ld), in[i * 4 + j * 2 + k]);
            float ob[8][8];      // [band][channel]
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float xin[8], b[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) xin[q] = in[q][c];
                haar_analysis(xin, b);
#pragma unroll
                for (int q = 0; q < 8; ++q) ob[q][c] = b[q];
            }
            {
                float o[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) o[c] = chain_scale(ob[0][c], L.lll_scale) + bs[c];
                const uint4 packed = pack8(o);
                *reinterpret_cast<uint4*>(L.y + vox2 * L.y_ld + cq * 8) = packed;
                if (want) {
                    float vr[8];
                    unpack8(packed, vr);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        st8[c] += vr[c];
                        st8[8 + c] = fmaf(vr[c], vr[c], st8[8 + c]);
                    }
                }
            }
            if (L.aux != nullptr) {
#pragma unroll
                for (int b = 1; b < 8; ++b) {
                    float o[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) o[c] = chain_scale(ob[b][c], L.hi_scale);
                    *reinterpret_cast<uint4*>(L.aux + (long long)(b - 1) * L.aux_sb + vox2 * L.aux_ld + cq * 8) = pack8(o);
                }
            }
        } else {
            const __nv_bfloat16* lll = reinterpret_cast<const __nv_bfloat16*>(L.residual);
            float bnd[8][8];     // [band][channel]
            unpack8(ld_cg_u4(lll + vox2 * L.res_ld + cq * 8), bnd[0]);
#pragma unroll
            for (int c = 0; c < 8; ++c) bnd[0][c] *= L.lll_scale;
#pragma unroll
            for (int b = 1; b < 8; ++b) unpack8(ld_cg_u4(L.aux + (long long)(b - 1) * L.aux_sb + vox2 * L.aux_ld + cq * 8), bnd[b]);
            float out[8][8];     // [brick position][channel]
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float b[8], xo[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) b[q] = bnd[q][c];
                haar_synthesis(b, xo);
#pragma unroll
                for (int q = 0; q < 8; ++q) out[q][c] = xo[q] + bs[c];
            }
            __nv_bfloat16* dst = L.y + vox0 * L.y_ld + cq * 8;
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const uint4 packed = pack8(out[i * 4 + j * 2 + k]);
                        *reinterpret_cast<uint4*>(dst + ((long long)(i * L.H + j) * L.W + k) * L.y_ld) = packed;
                        if (want) {
                            float vr[8];
                            unpack8(packed, vr);
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                st8[c] += vr[c];
                                st8[8 + c] = fmaf(vr[c], vr[c], st8[8 + c]);
                            }
                        }
                    }
        }
    }
}

// All kAuxWorkers threads of the CTA: run the op sample by sample, reduce the statistics (lanes sharing a channel chunk
// -> per-warp rows in shared memory -> one fp64 atomic per (group, component) and CTA), fence the stores, meet on barrier 3.
__device__ __forceinline__ void chain_aux_op(const ChainLayer& L, int wid, float* scratch) {
    const int lane = wid & 31, wrp = wid >> 5;                    // 12 worker warps
    const int C8 = L.Cin >> 3;                                    // 8, 16 or 32 (checked on the host)
    const bool want = L.gn_stats != nullptr;
    for (int n = 0; n < L.N; ++n) {
        float st8[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) st8[i] = 0.f;
        chain_aux_pass(L, n, wid, st8);
        if (want) {
            for (int off = 16; off >= C8; off >>= 1)              // lanes l, l + C8, ... own the same channel chunk
#pragma unroll
                for (int i = 0; i < 16; ++i) st8[i] += __shfl_xor_sync(0xffffffffu, st8[i], off);
            if (lane < C8) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    scratch[(wrp * 2 + 0) * 256 + lane * 8 + c] = st8[c];          // [warp][component][channel]
                    scratch[(wrp * 2 + 1) * 256 + lane * 8 + c] = st8[8 + c];
                }
            }
            asm volatile("bar.sync 3, %0;" ::"n"(kAuxWorkers) : "memory");
            if (wid < 2 * L.gn_groups) {
                const int g = wid >> 1, comp = wid & 1;
                double v = 0.0;
                for (int w = 0; w < kAuxWorkers / 32; ++w)
                    for (int c = g * L.gn_cpg; c < (g + 1) * L.gn_cpg; ++c) v += (double)scratch[(w * 2 + comp) * 256 + c];
                atomicAdd(L.gn_stats + (((long long)n * FCWDM_GN_STAT_REPLICAS + (blockIdx.x % FCWDM_GN_STAT_REPLICAS)) * L.gn_groups) * 2 + wid, v);
            }
            asm volatile("bar.sync 3, %0;" ::"n"(kAuxWorkers) : "memory");
        }
    }
    __threadfence();                                              // this worker's stores / atomics are visible device-wide
    asm volatile("bar.sync 3, %0;" ::"n"(kAuxWorkers) : "memory");
}

constexpr int kCWarpProdA = 4, kCWarpProdB = 5, kCWarpAlloc = 6, kCWarpMma = 7;

__global__ void __launch_bounds__(ChainCfg::THREADS, 1) conv3d_chain_kernel(const __grid_constant__ ChainParams P) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    using Cfg = ChainCfg;
    constexpr int N_TILE = Cfg::N_TILE;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_a = smem_base;
    const uint32_t smem_b = smem_a + Cfg::A_SLOTS * Cfg::SLOT_BYTES;
    const uint32_t recv_u32 = smem_b + Cfg::B_STAGES * Cfg::B_BYTES;
    const uint32_t bars = recv_u32 + Cfg::RECV_BYTES;
    const uint32_t full_a = bars;
    const uint32_t empty_a = full_a + 8 * Cfg::A_SLOTS;
    const uint32_t landed_a = empty_a + 8 * Cfg::A_SLOTS;
    const uint32_t full_b = landed_a + 8 * Cfg::A_SLOTS;
    const uint32_t empty_b = full_b + 8 * Cfg::B_STAGES;
    const uint32_t tmem_full = empty_b + 8 * Cfg::B_STAGES;
    const uint32_t tmem_empty = tmem_full + 8 * Cfg::ACC_STAGES;
    const uint32_t ready = tmem_empty + 8 * Cfg::ACC_STAGES;      // [3]: sender slot j has delivered its partial columns
    const uint32_t tmem_slot = ready + 8 * 3;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    float* recv = reinterpret_cast<float*>(smem_raw + (recv_u32 - smem_u32(smem_raw)));
    float* sbias = reinterpret_cast<float*>(smem_raw + (bars + 1024 - smem_u32(smem_raw)));   // [N_TILE]
    float* wstat = reinterpret_cast<float*>(smem_raw + (bars + 1536 - smem_u32(smem_raw)));   // [4 warps][32 groups][2]
    float* sgn = reinterpret_cast<float*>(smem_raw + (bars + 2560 - smem_u32(smem_raw)));     // [2][256] GN scale / shift
    float* gstat = reinterpret_cast<float*>(smem_raw + (bars + 512 - smem_u32(smem_raw)));    // [2][64] group mean / rstd

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int crank = (int)cluster_ctarank();
    const unsigned int G = gridDim.x;

    if (threadIdx.x == 0) {
        for (int i = 0; i < Cfg::A_SLOTS; ++i) {
            mbar_init(full_a + 8 * i, Cfg::XF_WARPS);   // the transform warps hand every plane over
            mbar_init(empty_a + 8 * i, 1);
            mbar_init(landed_a + 8 * i, 1);
        }
        for (int i = 0; i < Cfg::B_STAGES; ++i) {
            mbar_init(full_b + 8 * i, 1);
            mbar_init(empty_b + 8 * i, 1);
        }
        for (int i = 0; i < Cfg::ACC_STAGES; ++i) {
            mbar_init(tmem_full + 8 * i, 1);
            mbar_init(tmem_empty + 8 * i, 4);
        }
        for (int i = 0; i < 3; ++i) mbar_init(ready + 8 * i, 1);   // the receiver's expect_tx; the senders' st.async complete the bytes
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == kCWarpAlloc) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();       // every CTA's barriers exist before a peer signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == kCWarpProdA) {
        // ================================ A producer: halo planes ================================
        if (lane == 0) {
            uint32_t q = 0;
            for (int li = 0; li < P.n_layers; ++li) {
                const ChainLayer& L = P.layers[li];
                if (L.kind != 0) continue;                            // auxiliary op: no operands to stage
                CHAIN_TRACE(li, 0);
                if (li > 0) {
                    grid_wait(P.sync, (unsigned int)li * G);          // the previous layer's output is complete everywhere
                    asm volatile("fence.proxy.async.global;" ::: "memory");   // ... and visible to the TMA (async proxy) reads
                }
                CHAIN_TRACE(li, 1);
                for (int it = 0; it <= L.waves_a; ++it) {
                    ChainWork wk;
                    if (!chain_work(L, it, crank, P.cluster, wk)) continue;
                    const ChainTile tc = chain_decode(wk.tile, L);
                    const int cbn = L.n_cb / wk.split, cb0 = wk.kr * cbn;
                    for (int cb = cb0; cb < cb0 + cbn; ++cb) {
                        for (int p = 0; p < 3; ++p, ++q) {
                            const uint32_t slot = q % Cfg::A_SLOTS, ph = (q / Cfg::A_SLOTS) & 1;
                            mbar_wait(empty_a + 8 * slot, ph ^ 1);
                            mbar_arrive_expect_tx(landed_a + 8 * slot, Cfg::PLANE_BYTES);
                            tma_load_5d(smem_a + slot * Cfg::SLOT_BYTES, &L.map_a, landed_a + 8 * slot, cb * 64, tc.w0 - 1,
                                        tc.h0 - 1, tc.d0 + p - 1, tc.n);
                        }
                    }
                }
            }
        }
    } else if (warp == kCWarpProdB) {
        // ================================ B producer: weight tiles (never waits for a grid barrier) ==================
        if (lane == 0) {
            uint32_t r = 0;
            for (int li = 0; li < P.n_layers; ++li) {
                const ChainLayer& L = P.layers[li];
                if (L.kind != 0) continue;
                for (int it = 0; it <= L.waves_a; ++it) {
                    ChainWork wk;
                    if (!chain_work(L, it, crank, P.cluster, wk)) continue;
                    const ChainTile tc = chain_decode(wk.tile, L);
                    const int cbn = L.n_cb / wk.split, cb0 = wk.kr * cbn;
                    for (int cb = cb0; cb < cb0 + cbn; ++cb) {
                        for (int tap = 0; tap < 27; tap += 3, ++r) {
                            const uint32_t st = r % Cfg::B_STAGES, ph = (r / Cfg::B_STAGES) & 1;
                            mbar_wait(empty_b + 8 * st, ph ^ 1);
                            mbar_arrive_expect_tx(full_b + 8 * st, Cfg::B_BYTES);
                            tma_load_3d(smem_b + st * Cfg::B_BYTES, &L.map_b, full_b + 8 * st, cb * 64, tc.n0, tap);
                        }
                    }
                }
            }
        }
    } else if (warp >= 8) {
        // ================================ operand hand-over (4 warps): fused GroupNorm + SiLU where the layer has one ===
        const int pt = threadIdx.x - 256;                        // 0 .. XF_THREADS-1
        constexpr int XT = Cfg::XF_THREADS;
        constexpr int CHUNKS = Cfg::HROWS * Cfg::ROWP * 8;        // 16-byte chunks per plane
        constexpr int PER_THREAD = (CHUNKS + XT - 1) / XT;
        // physical chunk c = pt + XT q sits in row c >> 3 at position c & 7; its logical (channel) chunk (c & 7) ^ ((c >> 3) & 7)
        // does not depend on q (XT is a multiple of 64): per-channel scale / shift live in registers per channel block
        const int jmine = (pt & 7) ^ ((pt >> 3) & 7);
        uint32_t q = 0;
        for (int li = 0; li < P.n_layers; ++li) {
            const ChainLayer& L = P.layers[li];
            if (L.kind != 0) {                                    // auxiliary op: these 8 warps are workers 128..383
                if (li > 0) {
                    if (pt == 0) grid_wait(P.sync, (unsigned int)li * G);
                    asm volatile("bar.sync 2, %0;" ::"n"(XT) : "memory");
                }
                chain_aux_op(L, 128 + pt, recv);
                continue;
            }
            const bool gn = L.gi_stats != nullptr;
            int cur_n = -1;
            float gam = 0.f, bet = 0.f;                           // this thread's channel (C_in <= 256 = XT), fetched before the wait
            if (gn && pt < L.Cin) {
                gam = __ldg(L.gi_gamma + pt);
                bet = __ldg(L.gi_beta + pt);
            }
            if (gn && li > 0) {                                   // the statistics were accumulated by the previous layers
                if (pt == 0) grid_wait(P.sync, (unsigned int)li * G);
                asm volatile("bar.sync 2, %0;" ::"n"(XT) : "memory");
            }
            for (int it = 0; it <= L.waves_a; ++it) {
                ChainWork wk;
                if (!chain_work(L, it, crank, P.cluster, wk)) continue;
                const ChainTile tc = chain_decode(wk.tile, L);
                if (gn && tc.n != cur_n) {
                    cur_n = tc.n;
                    asm volatile("bar.sync 2, %0;" ::"n"(XT) : "memory");
                    // group statistics: 4 lanes per group, each sums a quarter of the replicas (8 independent L2 loads in
                    // flight per lane instead of a serial 32-load chain per channel), butterfly, one fp64 finish per group
                    const int cpg = L.Cin / L.gi_groups;
                    const double inv_cnt = 1.0 / ((double)L.D * L.H * L.W * cpg);
                    for (int idx = pt; idx < L.gi_groups * 4; idx += XT) {
                        const int g = idx >> 2, rq = idx & 3;
                        constexpr int RPQ = FCWDM_GN_STAT_REPLICAS / 4;
                        double v[2 * RPQ];
#pragma unroll
                        for (int r = 0; r < RPQ; ++r) {
                            const double* sp = L.gi_stats + (((long long)tc.n * FCWDM_GN_STAT_REPLICAS + rq * RPQ + r) * L.gi_groups + g) * 2;
                            v[2 * r] = ld_cg_f64(sp);
                            v[2 * r + 1] = ld_cg_f64(sp + 1);
                        }
                        double sum = 0.0, sq = 0.0;
#pragma unroll
                        for (int r = 0; r < RPQ; ++r) {
                            sum += v[2 * r];
                            sq += v[2 * r + 1];
                        }
                        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                        sq += __shfl_xor_sync(0xffffffffu, sq, 1);
                        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                        sq += __shfl_xor_sync(0xffffffffu, sq, 2);
                        if (rq == 0) {
                            const double mean = sum * inv_cnt;
                            double var = sq * inv_cnt - mean * mean;           // fp64: the subtraction cancels
                            var = var < 0.0 ? 0.0 : var;
                            gstat[g] = (float)mean;
                            gstat[64 + g] = rsqrtf((float)var + L.gi_eps);
                        }
                    }
                    asm volatile("bar.sync 2, %0;" ::"n"(XT) : "memory");
                    if (pt < L.Cin) {
                        const int g = pt / cpg;
                        const float sc0 = gstat[64 + g] * gam;
                        sgn[pt] = sc0;
                        sgn[256 + pt] = bet - gstat[g] * sc0;
                    }
                    asm volatile("bar.sync 2, %0;" ::"n"(XT) : "memory");
                    if (pt == 0) CHAIN_TRACE(li, 2);
                }
                const int cbn = L.n_cb / wk.split, cb0 = wk.kr * cbn;
                for (int cb = cb0; cb < cb0 + cbn; ++cb) {
                    float sc[8], sh[8];
                    if (gn) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            sc[e] = sgn[cb * 64 + jmine * 8 + e];
                            sh[e] = sgn[256 + cb * 64 + jmine * 8 + e];
                        }
                    }
                    for (int p = 0; p < 3; ++p, ++q) {
                        const uint32_t slot = q % Cfg::A_SLOTS;
                        const int d = tc.d0 + p - 1;
                        mbar_wait(landed_a + 8 * slot, (q / Cfg::A_SLOTS) & 1);
                        if (pt == 0 && cb == cb0 && p == 0) CHAIN_TRACE(li, 3);
                        if (gn && d >= 0 && d < L.D) {
                            uint4* plane = reinterpret_cast<uint4*>(smem_raw + (smem_a + slot * Cfg::SLOT_BYTES - smem_u32(smem_raw)));
                            uint4 raw[PER_THREAD];
                            uint32_t valid = 0;
#pragma unroll
                            for (int i = 0; i < PER_THREAD; ++i) {
                                const int c = pt + i * XT;
                                const int r = c >> 3;
                                const int hr = r / Cfg::ROWP, wc = r - hr * Cfg::ROWP;
                                const int h = tc.h0 - 1 + hr, w = tc.w0 - 1 + wc;
                                // out-of-range halo voxels stay ZERO: the convolution pads the ACTIVATED tensor
                                const bool in = (c < CHUNKS) && (h >= 0) && (h < L.H) && (w >= 0) && (w < L.W);
                                if (in) {
                                    raw[i] = plane[c];
                                    valid |= 1u << i;
                                }
                            }
#pragma unroll
                            for (int i = 0; i < PER_THREAD; ++i) {
                                if (valid & (1u << i)) {             // ragged tiles (7x7, 14x14 planes): most of the halo is padding
                                    float f[8];
                                    unpack8(raw[i], f);
#pragma unroll
                                    for (int e = 0; e < 8; ++e) f[e] = c_silu(fmaf(f[e], sc[e], sh[e]));
                                    plane[pt + i * XT] = pack8(f);
                                }
                            }
                            fence_proxy_async();         // generic-proxy smem writes -> visible to the tensor-core (async) proxy
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(full_a + 8 * slot);
                        if (pt == 0 && cb == cb0 && p == 0) CHAIN_TRACE(li, 4);
                    }
                }
            }
        }
    } else if (warp == kCWarpMma) {
        // ================================ MMA issuer (one elected lane for the whole kernel) ==========================
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N_TILE >> 3) << 17) |
                                   ((uint32_t)(128 >> 4) << 24);
        const uint64_t a_desc_base = make_sw128_desc(smem_a, Cfg::ROWP * 128);
        const uint64_t b_desc_base = make_sw128_desc(smem_b, 1024);
        if (elect_one()) {
            uint32_t q = 0, r = 0, acc_it = 0;
            for (int li = 0; li < P.n_layers; ++li) {
                const ChainLayer& L = P.layers[li];
                if (L.kind != 0) continue;
                for (int it = 0; it <= L.waves_a; ++it) {
                    ChainWork wk;
                    if (!chain_work(L, it, crank, P.cluster, wk)) continue;
                    const uint32_t as = acc_it % Cfg::ACC_STAGES, aph = (acc_it / Cfg::ACC_STAGES) & 1;
                    mbar_wait(tmem_empty + 8 * as, aph ^ 1);
                    tc_fence_after();
                    const uint32_t acc0 = tmem_base + as * N_TILE;
                    const int cbn = L.n_cb / wk.split;
                    for (int cb = 0; cb < cbn; ++cb) {
                        for (int kd = 0; kd < 3; ++kd, ++q) {
                            const uint32_t slot = q % Cfg::A_SLOTS;
                            mbar_wait(full_a + 8 * slot, (q / Cfg::A_SLOTS) & 1);
                            if (cb == 0 && kd == 0) CHAIN_TRACE(li, 5);
                            const uint64_t a_desc = a_desc_base + (uint64_t)((slot * Cfg::SLOT_BYTES) >> 4);
                            for (int kh = 0; kh < 3; ++kh, ++r) {
                                const uint32_t st = r % Cfg::B_STAGES;
                                mbar_wait(full_b + 8 * st, (r / Cfg::B_STAGES) & 1);
                                tc_fence_after();
#pragma unroll
                                for (int kw = 0; kw < 3; ++kw) {
                                    const uint64_t bd = b_desc_base + (uint64_t)((st * Cfg::B_BYTES + kw * Cfg::B_TAP_BYTES) >> 4);
                                    const uint64_t ad = a_desc + (uint64_t)(((kh * Cfg::ROWP + kw) * 128) >> 4);
                                    const uint32_t first = ((cb == 0) && (kd == 0) && (kh == 0) && (kw == 0)) ? 0u : 1u;
                                    umma_bf16(acc0, ad, bd, idesc, first);
                                    umma_bf16(acc0, ad + 2, bd + 2, idesc, 1u);
                                    umma_bf16(acc0, ad + 4, bd + 4, idesc, 1u);
                                    umma_bf16(acc0, ad + 6, bd + 6, idesc, 1u);
                                }
                                umma_commit(empty_b + 8 * st);
                            }
                            umma_commit(empty_a + 8 * slot);      // plane kd is not needed by later taps of this channel block
                        }
                    }
                    umma_commit(tmem_full + 8 * as);
                    CHAIN_TRACE(li, 6);
                    ++acc_it;
                }
            }
        }
        __syncwarp();
    } else if (warp < 4) {
        // ================================ epilogue (4 warps = the TMEM lane quarters) ================================
        const int ew = warp;
        const int row = ew * 32 + lane;           // accumulator row = voxel within the 16 x 8 tile
        const int hh = row >> 3, ww = row & 7;
        float* my_stat = wstat + ew * 64;
        uint32_t acc_it = 0;
        uint32_t ready_uses[3] = {0u, 0u, 0u};
        for (int li = 0; li < P.n_layers; ++li) {
            const ChainLayer& L = P.layers[li];
            const bool want_stats = L.gn_stats != nullptr;
            if (li > 0) {
                // lockstep: a CTA without work in this layer must not arrive at THIS layer's barrier before the previous one
                // is complete (the counter is cumulative: early arrivals would be counted towards the previous layer)
                if (ew == 0 && lane == 0) grid_wait(P.sync, (unsigned int)li * G);
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            if (L.kind != 0) {                                    // auxiliary op: these 4 warps are workers 0..127
                chain_aux_op(L, row, recv);
                if (li + 1 < P.n_layers) {                        // every worker has fenced its stores (barrier 3)
                    asm volatile("fence.proxy.async.global;" ::: "memory");
                    if (ew == 0 && lane == 0)
                        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(P.sync), "r"(1u) : "memory");
                }
                continue;
            }
            if (want_stats) {
                my_stat[lane] = 0.f;
                my_stat[lane + 32] = 0.f;
                __syncwarp();
            }
            int cur_n = -1, cur_n0 = -1;
            for (int it = 0; it <= L.waves_a; ++it) {
                ChainWork wk;
                if (!chain_work(L, it, crank, P.cluster, wk)) continue;
                const ChainTile tc = chain_decode(wk.tile, L);
                if (tc.n != cur_n || tc.n0 != cur_n0) {
                    if (want_stats && tc.n != cur_n && cur_n >= 0) {
                        // flush the finished sample's statistics: 4 warps -> one fp64 atomic per (group, component)
                        asm volatile("bar.sync 1, 128;" ::: "memory");
                        const int e = ew * 32 + lane;
                        if (e < 2 * L.gn_groups) {
                            const double v = (double)wstat[e] + (double)wstat[64 + e] + (double)wstat[128 + e] + (double)wstat[192 + e];
                            atomicAdd(L.gn_stats + (((long long)cur_n * FCWDM_GN_STAT_REPLICAS + (blockIdx.x % FCWDM_GN_STAT_REPLICAS)) * L.gn_groups) * 2 + e, v);
                        }
                        asm volatile("bar.sync 1, 128;" ::: "memory");
                        my_stat[lane] = 0.f;
                        my_stat[lane + 32] = 0.f;
                        __syncwarp();
                    }
                    cur_n = tc.n;
                    cur_n0 = tc.n0;
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    {
                        const int co = tc.n0 + row;                   // row < N_TILE = 128 always
                        float bv = 0.f;
                        if (co < L.Cout) {
                            if (L.bias != nullptr) bv += __ldg(L.bias + co);
                            if (L.chan_bias != nullptr) bv += __ldg(L.chan_bias + (long long)tc.n * L.cb_ld + co);
                        }
                        sbias[row] = bv;
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
                const uint32_t as = acc_it % Cfg::ACC_STAGES, aph = (acc_it / Cfg::ACC_STAGES) & 1;
                ++acc_it;
                const int s = wk.split;
                const int CW = N_TILE / s;                            // accumulator columns this CTA finishes
                const int col0 = wk.kr * CW;
                const int h = tc.h0 + hh, w = tc.w0 + ww;
                const bool ok = (h < L.H) && (w < L.W);
                const long long vox = (((long long)tc.n * L.D + tc.d0) * L.H + h) * L.W + w;
                const bool use_res = ok && L.residual != nullptr;
                // the residual row of my first 32 columns: in flight while the MMAs finish and the partial sums travel
                uint4 res[4];
                if (use_res) {
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        if (tc.n0 + col0 + g * 8 < L.Cout) res[g] = ld_cg_u4(L.residual + vox * L.res_ld + tc.n0 + col0 + g * 8);
                }
                mbar_wait(tmem_full + 8 * as, aph);
                tc_fence_after();
                if (row == 0) CHAIN_TRACE(li, 7);
                const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + as * N_TILE;
                if (s > 1) {
                    // ---- reduce-scatter: send every peer the columns it owns, receive mine from every peer.
                    // A shared tile is always this CTA's LAST work of the layer, and the next layer's halo planes cannot
                    // land before the grid barrier this CTA has yet to arrive at: the plane ring is idle, so it serves as
                    // the staging buffer.  TMEM -> registers -> staging (padded rows, conflict-free), then ONE bulk
                    // distributed-shared-memory copy per peer that completes its bytes on the RECEIVER's mbarrier.
                    // (Measured alternatives: per-thread st.shared::cluster + cluster-scope release/acquire 5-6 us per tile --
                    // a MEMBAR.ALL.GPU per arrive; st.async 4 us -- one mbarrier transaction per 16 bytes.)
                    const uint32_t gb = (uint32_t)(crank - wk.kr);    // cluster rank of the group's first CTA
                    const int pitch = CW + 4;                         // floats per row of a staging / receive slot
                    const uint32_t slot_bytes = (uint32_t)(128 * pitch * 4);
                    if (ew == 0 && lane == 0)
                        for (int j = 0; j < s - 1; ++j) mbar_arrive_expect_tx(ready + 8 * j, slot_bytes);
                    for (int qd = 0; qd < s; ++qd) {
                        if (qd == wk.kr) continue;
                        const int j = (wk.kr - qd + s) % s - 1;       // my slot in peer qd's receive buffer (and in my staging)
                        uint4* stg = reinterpret_cast<uint4*>(smem_raw + (smem_a + (uint32_t)j * slot_bytes + (uint32_t)(row * pitch * 4) -
                                                                          smem_u32(smem_raw)));
#pragma unroll 1
                        for (int c = 0; c < CW; c += 32) {
                            uint32_t acc[32];
                            tmem_ld_x16(taddr + qd * CW + c, acc);
                            tmem_ld_x16(taddr + qd * CW + c + 16, acc + 16);
                            tmem_ld_wait();
#pragma unroll
                            for (int e = 0; e < 8; ++e)
                                stg[(c >> 2) + e] = make_uint4(acc[4 * e], acc[4 * e + 1], acc[4 * e + 2], acc[4 * e + 3]);
                        }
                    }
                    fence_proxy_async();                              // staging writes -> visible to the bulk copy (async proxy)
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (ew == 0 && lane == 0) {
                        for (int qd = 0; qd < s; ++qd) {
                            if (qd == wk.kr) continue;
                            const int j = (wk.kr - qd + s) % s - 1;
                            const uint32_t dst = mapa_u32(recv_u32 + (uint32_t)j * slot_bytes, gb + (uint32_t)qd);
                            const uint32_t rbar = mapa_u32(ready + 8 * j, gb + (uint32_t)qd);
                            asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                         ::"r"(dst), "r"(smem_a + (uint32_t)j * slot_bytes), "r"(slot_bytes), "r"(rbar) : "memory");
                        }
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    for (int j = 0; j < s - 1; ++j) {
                        mbar_wait(ready + 8 * j, ready_uses[j] & 1);
                        ++ready_uses[j];
                    }
                }
                if (row == 0) CHAIN_TRACE(li, 8);
#pragma unroll 1
                for (int c0 = col0; c0 < col0 + CW; c0 += 32) {
                    uint32_t acc[32];
                    tmem_ld_x16(taddr + c0, acc);
                    tmem_ld_x16(taddr + c0 + 16, acc + 16);
                    tmem_ld_wait();
                    if (row == 0 && c0 == col0) CHAIN_TRACE(li, 12);
                    for (int j = 0; j < s - 1; ++j) {                 // + the peers' partial sums for my columns
                        const float4* pr = reinterpret_cast<const float4*>(recv + (j * 128 + row) * (CW + 4) + (c0 - col0));
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float4 pv = pr[e];
                            acc[4 * e] = __float_as_uint(__uint_as_float(acc[4 * e]) + pv.x);
                            acc[4 * e + 1] = __float_as_uint(__uint_as_float(acc[4 * e + 1]) + pv.y);
                            acc[4 * e + 2] = __float_as_uint(__uint_as_float(acc[4 * e + 2]) + pv.z);
                            acc[4 * e + 3] = __float_as_uint(__uint_as_float(acc[4 * e + 3]) + pv.w);
                        }
                    }
                    if (row == 0 && c0 == col0) CHAIN_TRACE(li, 13);
                    // statistics: one warp reduction per 32-column chunk (conv3d_gn.cuh)
                    const bool chunk_stats = want_stats && gn_chunk_ok(L.gn_cpg);
                    float ca[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) ca[e] = 0.f;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int col = c0 + g * 8;
                        const int co = tc.n0 + col;
                        if (co < L.Cout) {                              // warp-uniform
                            float v[8];
                            const float4 b0 = *reinterpret_cast<const float4*>(sbias + col);
                            const float4 b1 = *reinterpret_cast<const float4*>(sbias + col + 4);
                            v[0] = __uint_as_float(acc[g * 8 + 0]) + b0.x; v[1] = __uint_as_float(acc[g * 8 + 1]) + b0.y;
                            v[2] = __uint_as_float(acc[g * 8 + 2]) + b0.z; v[3] = __uint_as_float(acc[g * 8 + 3]) + b0.w;
                            v[4] = __uint_as_float(acc[g * 8 + 4]) + b1.x; v[5] = __uint_as_float(acc[g * 8 + 5]) + b1.y;
                            v[6] = __uint_as_float(acc[g * 8 + 6]) + b1.z; v[7] = __uint_as_float(acc[g * 8 + 7]) + b1.w;
                            if (use_res) {
                                float rr[8];
                                unpack8(res[g], rr);
#pragma unroll
                                for (int e = 0; e < 8; ++e) v[e] += rr[e];
                            }
                            const uint4 packed = pack8(v);
#if !(defined(FCWDM_CHAIN_EXP) && (FCWDM_CHAIN_EXP & 2))
                            if (ok) *reinterpret_cast<uint4*>(L.y + vox * L.y_ld + co) = packed;
#endif
                            if (use_res && c0 + 32 < col0 + CW && co + 32 < L.Cout)   // next 32 columns' residual, one chunk ahead
                                res[g] = ld_cg_u4(L.residual + vox * L.res_ld + co + 32);
#if defined(FCWDM_CHAIN_EXP) && (FCWDM_CHAIN_EXP & 1)
                            if (want_stats && packed.x == 0x12345678u) {
#else
                            if (want_stats) {                          // statistics of the STORED (bf16) values
#endif
                                float vr[8];
                                unpack8(packed, vr);
                                if (!ok) {
#pragma unroll
                                    for (int e = 0; e < 8; ++e) vr[e] = 0.f;
                                }
                                if (chunk_stats) {
                                    gn_chunk_add(ca, g, vr);
                                } else switch (L.gn_cpg) {
                                    case 1: gn_accumulate<1>(vr, my_stat, co, lane); break;
                                    case 2: gn_accumulate<2>(vr, my_stat, co, lane); break;
                                    case 4: gn_accumulate<4>(vr, my_stat, co, lane); break;
                                    default: gn_accumulate_wide(vr, my_stat, co / L.gn_cpg, lane); break;
                                }
                            }
                        }
                    }
                    if (chunk_stats && tc.n0 + c0 < L.Cout) gn_chunk_flush(ca, my_stat, tc.n0 + c0, L.gn_cpg, lane);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tmem_empty + 8 * as);
                if (row == 0) CHAIN_TRACE(li, 9);
            }
            // ---- end of layer: publish this CTA's share (stores + statistics), then arrive at the grid barrier
            if (want_stats && cur_n >= 0) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int e = ew * 32 + lane;
                if (e < 2 * L.gn_groups) {
                    const double v = (double)wstat[e] + (double)wstat[64 + e] + (double)wstat[128 + e] + (double)wstat[192 + e];
                    atomicAdd(L.gn_stats + (((long long)cur_n * FCWDM_GN_STAT_REPLICAS + (blockIdx.x % FCWDM_GN_STAT_REPLICAS)) * L.gn_groups) * 2 + e, v);
                }
            }
            if (row == 0) CHAIN_TRACE(li, 10);
            if (ew == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging (plane ring) is free again
            if (li + 1 < P.n_layers) {
                __threadfence();                                          // my stores / atomics are visible device-wide ...
                asm volatile("fence.proxy.async.global;" ::: "memory");   // ... also to other SMs' TMA reads
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (ew == 0 && lane == 0)
                    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(P.sync), "r"(1u) : "memory");
                if (row == 0) CHAIN_TRACE(li, 11);
            } else {
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();       // nobody exits while a peer may still signal its barriers / store into its receive slots
    if (warp == kCWarpAlloc) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFnC)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnC g_encode_c = nullptr;
static long long* g_chain_trace = nullptr;      // development only (fcwdm_debug_set_chain_trace)
static int g_chain_clusters[kChainMaxCluster + 1] = {0, 0, 0, 0, 0};      // co-resident clusters (one CTA per SM) by cluster size

int conv3d_chain_init_device() {
    if (g_encode_c == nullptr) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        FCWDM_REQUIRE(e == cudaSuccess && fn != nullptr && qres == cudaDriverEntryPointSuccess, FCWDM_ERR_CUDA,
                      "fcwdm_init: cuTensorMapEncodeTiled entry point not available (%s)", cudaGetErrorString(e));
        g_encode_c = reinterpret_cast<EncodeTiledFnC>(fn);
    }
    cudaError_t e = cudaFuncSetAttribute(conv3d_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ChainCfg::SMEM_BYTES);
    FCWDM_REQUIRE(e == cudaSuccess, FCWDM_ERR_CUDA, "fcwdm_init: cudaFuncSetAttribute(chain) failed: %s", cudaGetErrorString(e));
    for (int cs = 2; cs <= kChainMaxCluster; cs *= 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs * 64);
        cfg.blockDim = dim3(ChainCfg::THREADS);
        cfg.dynamicSmemBytes = ChainCfg::SMEM_BYTES;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cs;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int n = 0;
        e = cudaOccupancyMaxActiveClusters(&n, conv3d_chain_kernel, &cfg);
        FCWDM_REQUIRE(e == cudaSuccess && n > 0, FCWDM_ERR_CUDA,
                      "fcwdm_init: cudaOccupancyMaxActiveClusters(chain, %d) failed: %s (%d)", cs, cudaGetErrorString(e), n);
        g_chain_clusters[cs] = n;
    }
    return FCWDM_OK;
}

}  // namespace fcwdm

using namespace fcwdm;

extern "C" int fcwdm_conv3d_chain_supported(int64_t Cin, int64_t Cout, int ksize) {
    return (ksize == 3 && Cin > 0 && Cin % 64 == 0 && Cout > 0 && Cout % 128 == 0) ? 1 : 0;
}

extern "C" int fcwdm_conv3d_chain_max_layers(void) { return kChainMaxLayers; }

extern "C" int fcwdm_conv3d_chain(const fcwdm_chain_layer* layers, int64_t n_layers, void* sync_counter, void* stream) {
    FCWDM_REQUIRE(layers != nullptr && sync_counter != nullptr, FCWDM_ERR_INVALID, "fcwdm_conv3d_chain: null pointer");
    FCWDM_REQUIRE(n_layers >= 1 && n_layers <= kChainMaxLayers, FCWDM_ERR_INVALID,
                  "fcwdm_conv3d_chain: 1 <= n_layers <= %d", kChainMaxLayers);
    FCWDM_REQUIRE((uintptr_t)sync_counter % 4 == 0, FCWDM_ERR_INVALID, "fcwdm_conv3d_chain: sync_counter must be 4-byte aligned");
    if (g_encode_c == nullptr || g_chain_clusters[2] == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        int rc = fcwdm_init(dev);
        if (rc) return rc;
    }
    // cluster size: 4 when every layer of the run fits one wave of the 4-CTA-cluster grid (then the remainder wave IS the
    // layer and a 4-way split pays), else 2 (all SMs, fewer idle CTAs behind the full waves of the larger layers)
    int csize = kChainMaxCluster;
    {
        const long long g4 = (long long)g_chain_clusters[4] * 4;
        for (int64_t i = 0; i < n_layers; ++i) {
            const fcwdm_chain_layer& l = layers[i];
            if (l.kind != FCWDM_CHAIN_CONV) continue;
            const long long tiles = l.N * l.D * ((l.H + 15) / 16) * ((l.W + 7) / 8) * (l.Cout / 128);
            if (tiles > g4) csize = 2;
        }
        static const int env_cs = getenv("FCWDM_CHAIN_CLUSTER") ? atoi(getenv("FCWDM_CHAIN_CLUSTER")) : 0;
        if (env_cs == 2 || env_cs == 4) csize = env_cs;
    }
    const int G = g_chain_clusters[csize] * csize;
    ChainParams params;            // 7 KB of kernel parameters (tensor maps live in parameter space)
    ChainParams* p = &params;
    p->n_layers = (int)n_layers;
    p->cluster = csize;
    p->trace = g_chain_trace;
    p->sync = (unsigned int*)sync_counter;
    for (int64_t i = 0; i < n_layers; ++i) {
        const fcwdm_chain_layer& l = layers[i];
        ChainLayer& L = p->layers[i];
        L.kind = (int)l.kind;
        if (l.kind != FCWDM_CHAIN_CONV) {
            // ---- auxiliary op: Haar DWT / IDWT on channels-last bf16, same operand meaning as fcwdm_dwt3d_cl / fcwdm_idwt3d_cl
            FCWDM_REQUIRE(l.kind == FCWDM_CHAIN_DWT || l.kind == FCWDM_CHAIN_IDWT, FCWDM_ERR_INVALID,
                          "fcwdm_conv3d_chain: op %d: unknown kind %d", (int)i, (int)l.kind);
            FCWDM_REQUIRE(l.x && l.y && (l.kind == FCWDM_CHAIN_DWT || l.aux), FCWDM_ERR_INVALID,
                          "fcwdm_conv3d_chain: op %d: null pointer", (int)i);
            FCWDM_REQUIRE(l.N > 0 && l.D > 0 && l.H > 0 && l.W > 0 && l.D % 2 == 0 && l.H % 2 == 0 && l.W % 2 == 0,
                          FCWDM_ERR_INVALID, "fcwdm_conv3d_chain: op %d: D, H, W must be positive and even", (int)i);
            FCWDM_REQUIRE(l.Cin == 64 || l.Cin == 128 || l.Cin == 256, FCWDM_ERR_UNSUPPORTED,
                          "fcwdm_conv3d_chain: op %d: in-chain DWT / IDWT need 64, 128 or 256 channels", (int)i);
            FCWDM_REQUIRE(l.x_ld >= l.Cin && l.x_ld % 8 == 0 && l.y_ld >= l.Cin && l.y_ld % 8 == 0 &&
                              (l.aux == nullptr || (l.aux_ld >= l.Cin && l.aux_ld % 8 == 0 && l.aux_sb % 8 == 0)) && l.cb_ld % 4 == 0,
                          FCWDM_ERR_INVALID, "fcwdm_conv3d_chain: op %d: bad leading dimension", (int)i);
            FCWDM_REQUIRE(((uintptr_t)l.x % 16 == 0) && ((uintptr_t)l.y % 16 == 0) && ((uintptr_t)l.aux % 16 == 0) &&
                              ((uintptr_t)l.chan_bias % 16 == 0),
                          FCWDM_ERR_INVALID, "fcwdm_conv3d_chain: op %d: pointers must be 16-byte aligned", (int)i);
            if (l.gn_stats != nullptr)
                FCWDM_REQUIRE(l.gn_groups > 0 && l.gn_groups <= 32 && l.Cin % l.gn_groups == 0, FCWDM_ERR_UNSUPPORTED,
                              "fcwdm_conv3d_chain: op %d: fused statistics need 1 <= groups <= 32 dividing C", (int)i);
            memset(&L.map_a, 0, sizeof(L.map_a));
            memset(&L.map_b, 0, sizeof(L.map_b));
            L.N = (int)l.N; L.D = (int)l.D; L.H = (int)l.H; L.W = (int)l.W;
            L.Cout = (int)l.Cin; L.Cin = (int)l.Cin;
            L.n_cb = 0; L.n_nt = L.n_wt = L.n_ht = 1; L.num_tiles = 0; L.waves_a = 0; L.split_b = 1;
            L.bias = nullptr; L.chan_bias = l.chan_bias; L.cb_ld = l.cb_ld;
            L.residual = (const __nv_bfloat16*)l.x; L.res_ld = l.x_ld;           // aux ops carry their input here
            L.y = (__nv_bfloat16*)l.y; L.y_ld = l.y_ld;
            L.gn_stats = l.gn_stats;
            L.gn_groups = l.gn_stats ? (int)l.gn_groups : 0;
            L.gn_cpg = l.gn_stats ? (int)(l.Cin / l.gn_groups) : 0;
            L.gi_stats = nullptr; L.gi_gamma = L.gi_beta = nullptr; L.gi_groups = 0; L.gi_eps = 0.f;
            L.aux = (__nv_bfloat16*)l.aux; L.aux_ld = l.aux_ld; L.aux_sb = l.aux_sb;
            L.lll_scale = l.lll_scale; L.hi_scale = l.hi_scale;
            continue;
        }
        L.aux = nullptr; L.aux_ld = L.aux_sb = 0; L.lll_scale = L.hi_scale = 1.f;
        FCWDM_REQUIRE(l.x && l.wp && l.y, FCWDM_ERR_INVALID, "fcwdm_conv3d_chain: layer %d: null pointer", (int)i);
        FCWDM_REQUIRE(l.N > 0 && l.D > 0 && l.H > 0 && l.W > 0 && l.N < 32768 && l.D < 32768 && l.H < 32768 && l.W < 32768,
                      FCWDM_ERR_INVALID, "fcwdm_conv3d_chain: layer %d: bad dimension", (int)i);
        FCWDM_REQUIRE(fcwdm_conv3d_chain_supported(l.Cin, l.Cout, 3), FCWDM_ERR_UNSUPPORTED,
                      "fcwdm_conv3d_chain: layer %d: needs C_in %% 64 == 0 and C_out %% 128 == 0 (3x3x3)", (int)i);
        FCWDM_REQUIRE(l.x_ld >= l.Cin && l.x_ld % 8 == 0 && l.y_ld >= l.Cout && l.y_ld % 8 == 0 &&
                          (l.residual == nullptr || (l.res_ld >= l.Cout && l.res_ld % 8 == 0)) && l.cb_ld % 4 == 0,
                      FCWDM_ERR_INVALID, "fcwdm_conv3d_chain: layer %d: bad leading dimension", (int)i);
        FCWDM_REQUIRE(((uintptr_t)l.x % 16 == 0) && ((uintptr_t)l.wp % 16 == 0) && ((uintptr_t)l.y % 16 == 0) &&
                          ((uintptr_t)l.residual % 16 == 0) && ((uintptr_t)l.bias % 16 == 0) && ((uintptr_t)l.chan_bias % 16 == 0),
                      FCWDM_ERR_INVALID, "fcwdm_conv3d_chain: layer %d: pointers must be 16-byte aligned", (int)i);
        if (l.gn_stats != nullptr) {
            FCWDM_REQUIRE(l.gn_groups > 0 && l.gn_groups <= 32 && l.Cout % l.gn_groups == 0, FCWDM_ERR_UNSUPPORTED,
                          "fcwdm_conv3d_chain: layer %d: fused statistics need 1 <= groups <= 32 dividing C_out", (int)i);
            const int64_t cpg = l.Cout / l.gn_groups;
            FCWDM_REQUIRE(cpg == 1 || cpg == 2 || cpg == 4 || cpg % 8 == 0, FCWDM_ERR_UNSUPPORTED,
                          "fcwdm_conv3d_chain: layer %d: fused statistics need channels/group in {1,2,4,8k}", (int)i);
        }
        if (l.gn_in_stats != nullptr) {
            FCWDM_REQUIRE(l.gn_in_gamma && l.gn_in_beta && l.Cin <= 256 && l.gn_in_groups > 0 && l.Cin % l.gn_in_groups == 0,
                          FCWDM_ERR_UNSUPPORTED,
                          "fcwdm_conv3d_chain: layer %d: fused input GroupNorm needs gamma, beta, C_in <= 256 divisible by the groups", (int)i);
        }
        const int64_t cin_p = l.Cin, cout_p = l.Cout;
        {
            cuuint64_t dims[5] = {(cuuint64_t)cin_p, (cuuint64_t)l.W, (cuuint64_t)l.H, (cuuint64_t)l.D, (cuuint64_t)l.N};
            cuuint64_t strides[4] = {(cuuint64_t)l.x_ld * 2, (cuuint64_t)l.W * l.x_ld * 2, (cuuint64_t)l.H * l.W * l.x_ld * 2,
                                     (cuuint64_t)l.D * l.H * l.W * l.x_ld * 2};
            cuuint32_t box[5] = {64, 10, 18, 1, 1};
            cuuint32_t es[5] = {1, 1, 1, 1, 1};
            CUresult r = g_encode_c(&L.map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(l.x), dims, strides, box, es,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            FCWDM_REQUIRE(r == CUDA_SUCCESS, FCWDM_ERR_CUDA, "fcwdm_conv3d_chain: layer %d: activation tensor map failed (%d)", (int)i, (int)r);
        }
        {
            cuuint64_t dims[3] = {(cuuint64_t)cin_p, (cuuint64_t)cout_p, 27};
            cuuint64_t strides[2] = {(cuuint64_t)cin_p * 2, (cuuint64_t)cout_p * cin_p * 2};
            cuuint32_t box[3] = {64, 128, 3};
            cuuint32_t es[3] = {1, 1, 1};
            CUresult r = g_encode_c(&L.map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(l.wp), dims, strides, box, es,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            FCWDM_REQUIRE(r == CUDA_SUCCESS, FCWDM_ERR_CUDA, "fcwdm_conv3d_chain: layer %d: weight tensor map failed (%d)", (int)i, (int)r);
        }
        L.N = (int)l.N; L.D = (int)l.D; L.H = (int)l.H; L.W = (int)l.W;
        L.Cout = (int)l.Cout;
        L.n_cb = (int)(cin_p / 64);
        L.n_nt = (int)(cout_p / 128);
        L.n_wt = (int)((l.W + 7) / 8);
        L.n_ht = (int)((l.H + 15) / 16);
        const long long tiles = (long long)l.N * l.D * L.n_ht * L.n_wt * L.n_nt;
        FCWDM_REQUIRE(tiles < (1ll << 30), FCWDM_ERR_UNSUPPORTED, "fcwdm_conv3d_chain: layer %d: too many tiles", (int)i);
        L.num_tiles = (int)tiles;
        L.waves_a = (int)(tiles / G);
        const long long rem = tiles - (long long)L.waves_a * G;
        int s = 1;
        for (int cand = csize; cand > 1; cand >>= 1)
            if (L.n_cb % cand == 0 && rem * cand <= G) { s = cand; break; }
        {
            static const int env_s = getenv("FCWDM_CHAIN_SPLIT") ? atoi(getenv("FCWDM_CHAIN_SPLIT")) : 0;   // development: cap the split
            if (env_s > 0 && s > env_s) s = env_s;
        }
        L.split_b = s;
        L.bias = l.bias; L.chan_bias = l.chan_bias; L.cb_ld = l.cb_ld;
        L.residual = (const __nv_bfloat16*)l.residual; L.res_ld = l.res_ld;
        L.y = (__nv_bfloat16*)l.y; L.y_ld = l.y_ld;
        L.gn_stats = l.gn_stats;
        L.gn_groups = l.gn_stats ? (int)l.gn_groups : 0;
        L.gn_cpg = l.gn_stats ? (int)(l.Cout / l.gn_groups) : 0;
        L.Cin = (int)l.Cin;
        L.gi_stats = l.gn_in_stats; L.gi_gamma = l.gn_in_gamma; L.gi_beta = l.gn_in_beta;
        L.gi_groups = (int)l.gn_in_groups; L.gi_eps = l.gn_in_eps;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)G);
    cfg.blockDim = dim3(ChainCfg::THREADS);
    cfg.dynamicSmemBytes = ChainCfg::SMEM_BYTES;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv3d_chain_kernel, *p);
    FCWDM_REQUIRE(e == cudaSuccess, FCWDM_ERR_CUDA, "fcwdm_conv3d_chain: launch failed: %s", cudaGetErrorString(e));
    FCWDM_CHECK_LAUNCH("fcwdm_conv3d_chain");
    return FCWDM_OK;
}

/* Development aid (tools/chain_trace.py): when set, every fcwdm_conv3d_chain launch writes %globaltimer stamps
 * [grid][32 ops][16] (slot meaning: see CHAIN_TRACE in conv3d_chain.cu) to this device buffer.  NULL switches it off. */
extern "C" int fcwdm_debug_set_chain_trace(void* device_buffer) {
#ifdef FCWDM_CONV_TRACE
    g_chain_trace = (long long*)device_buffer;
    return FCWDM_OK;
#else
    g_chain_trace = nullptr;
    FCWDM_REQUIRE(device_buffer == nullptr, FCWDM_ERR_UNSUPPORTED,
                  "fcwdm_debug_set_chain_trace: this build has no trace points (rebuild with FCWDM_CONV_TRACE=1)");
    return FCWDM_OK;
#endif
}

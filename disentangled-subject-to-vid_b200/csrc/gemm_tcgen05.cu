// Persistent warp-specialised tcgen05 GEMM for sm_100a:  Y[M,N] = X[M,K] W[N,K]^T (+ T[M,r] B[N,r]^T) with fused epilogues.
//
//   warp 0      : TMA producer  (cp.async.bulk.tensor 2D, 128B swizzle, BLOCK_K = 64 bf16 = one swizzle row)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (M = 128, N = BN, K = 16 per instruction)
//   warps 2..5  : epilogue (tcgen05.ld 32x32b -> bias / GELU / gate-residual -> 16-byte global stores)
//
// Accumulators are double buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the main loop of tile
// i+1.  Tiles are rasterised in groups of GROUP_M row-blocks so that one wave of 148 CTAs re-uses both the activation
// panel and the weight panel out of the 126 MB L2.  The LoRA up-projection is folded in as extra K blocks read through
// a second pair of tensor maps (no concatenated copies are materialised).
#include <cstdlib>
#include "common.cuh"
#include "host_util.h"
#include "s2v_b200.h"
#include <string.h>

namespace s2v {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_GROUP_M_DEFAULT = 16;
constexpr int GEMM_THREADS = 192;
// The fused q/k LayerNorm + RoPE epilogue is ~5x the plain one per element: with 4 epilogue warps it took longer than the tile's main
// loop (tensor pipe 76 % in ncu, profiles/r02_qkv_fused_ncu_full.txt).  That instantiation runs 8 epilogue warps — two per TMEM
// lane quarter, each taking half of the tile's head vectors.
constexpr int GEMM_THREADS_WIDE_EPI = 320;
__host__ __device__ constexpr int gemm_threads(int epi) { return epi == 4 /* S2V_EPI_QKV_NORM_ROPE */ ? GEMM_THREADS_WIDE_EPI : GEMM_THREADS; }

struct GemmKParams {
    int M, N, K, K2;
    int group_m;        // row-blocks per rasterisation group (see tile_coords)
    int lora_group_n;
    const bf16* bias;
    bf16* out;
    long long ldo;
    float alpha;
    const float* mod;
    int mod_stride, gate_off_text, gate_off_other, rows_per_batch, text_len;
    // implicit-GEMM convolution (S2V_EPI_CONV): K block kb reads channels [(kb % cin_blocks)*64, +64) of A rows shifted by
    // tap_off[kb / cin_blocks]; output row m is stored at row out_row0 + m, zeroed when it is a spatial border position
    int taps, cin_blocks, a_row0, out_row0, Hp, Wp;
    int tap_off[27];
    const bf16* res;
    long long ldres;
    // GroupNorm statistics of the convolution's OUTPUT, fused into the epilogue (S2V_EPI_CONV / S2V_EPI_CONV_T): per row-block partial
    // sums [num_row_blocks, N, 2] (sum, sum of squares of the bf16-rounded values that are stored; the zero border ring adds nothing)
    float* stats;
    // fused q/k LayerNorm + RoPE (S2V_EPI_QKV_NORM_ROPE): columns [0, qk_cols) are 64-wide q then k head vectors
    const bf16 *nq_w, *nq_b, *nk_w, *nk_b;
    const float *rope_cos, *rope_sin;
    int qk_cols;    // 2*H*64
    float qk_eps;
};

constexpr int S2V_EPI_CONV = 3;
constexpr int S2V_EPI_QKV_NORM_ROPE = 4;
// Convolution with the operands swapped (Cout = 128 layers): the weights [128, taps*Cin] are the M-side operand and 256
// consecutive output positions the N-side one, D^T[cout, position].  A 128x128 tile reads 8 KB of shared memory per 64 tensor
// cycles (the full 128 B/clk: the BN = 128 conv ran at 47 % tensor-pipe activity), a 128x256 tile 12 KB per 128.
constexpr int S2V_EPI_CONV_T = 5;

// Per-head LayerNorm(64) + interleaved RoPE on one head vector held by ONE thread, with exactly the arithmetic (operation
// order, explicit FMAs, bf16 rounding points) of qk_norm_rope_kernel in elementwise.cu, so that the fused epilogue and the
// stand-alone kernel produce identical bits.  f[] enters as bf16(acc + bias) values.
__device__ __forceinline__ void head_norm_rope(float (&f)[64], const bf16* __restrict__ w, const bf16* __restrict__ b,
                                               const float* __restrict__ cs, const float* __restrict__ sn, float eps) {
    float g[8];
#pragma unroll
    for (int l = 0; l < 8; ++l) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) a += f[l * 8 + j];
        g[l] = a;
    }
    // the stand-alone kernel combines its 8 lanes with an xor-shuffle tree (1, 2, 4)
    const float mean = (((g[0] + g[1]) + (g[2] + g[3])) + ((g[4] + g[5]) + (g[6] + g[7]))) * (1.0f / 64.0f);
#pragma unroll
    for (int l = 0; l < 8; ++l) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float d = f[l * 8 + j] - mean;
            a = fmaf(d, d, a);
        }
        g[l] = a;
    }
    const float rstd = rsqrtf((((g[0] + g[1]) + (g[2] + g[3])) + ((g[4] + g[5]) + (g[6] + g[7]))) * (1.0f / 64.0f) + eps);
#pragma unroll
    for (int v8 = 0; v8 < 8; ++v8) {
        const uint4 wu = __ldg(reinterpret_cast<const uint4*>(w) + v8), bu = __ldg(reinterpret_cast<const uint4*>(b) + v8);
        const uint32_t ww[4] = {wu.x, wu.y, wu.z, wu.w}, bw[4] = {bu.x, bu.y, bu.z, bu.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            f[v8 * 8 + 2 * j] = fmaf((f[v8 * 8 + 2 * j] - mean) * rstd, bf16_lo(ww[j]), bf16_lo(bw[j]));
            f[v8 * 8 + 2 * j + 1] = fmaf((f[v8 * 8 + 2 * j + 1] - mean) * rstd, bf16_hi(ww[j]), bf16_hi(bw[j]));
        }
    }
    if (cs) {
#pragma unroll
        for (int v4 = 0; v4 < 16; ++v4) {
            const float4 c4 = __ldg(reinterpret_cast<const float4*>(cs) + v4), s4 = __ldg(reinterpret_cast<const float4*>(sn) + v4);
            const float c[4] = {c4.x, c4.y, c4.z, c4.w}, sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
            for (int j = 0; j < 4; j += 2) {
                const float x0 = __bfloat162float(__float2bfloat16(f[v4 * 4 + j])), x1 = __bfloat162float(__float2bfloat16(f[v4 * 4 + j + 1]));
                f[v4 * 4 + j] = fmaf(x0, c[j], -(x1 * sv[j]));
                f[v4 * 4 + j + 1] = fmaf(x1, c[j + 1], x0 * sv[j + 1]);
            }
        }
    }
}

template <int BN, bool TWO = false>
struct GemmCfg {
    // TWO (cta_group::2): a CTA pair owns a 256 x BN tile; each CTA stages its own 128 A rows and HALF of the B rows
    static constexpr int STAGES = TWO ? 6 : (BN == 256) ? 4 : 6;
    static constexpr uint32_t A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr uint32_t B_BYTES = (TWO ? BN / 2 : BN) * GEMM_BK * 2;
    static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr uint32_t TMEM_COLS = 2 * BN;
    static constexpr uint32_t STATS_BYTES = 2 * 4 * BN * 2 * 4;   // conv epilogue: 2 buffers x 4 warps x BN columns x (sum, sumsq) fp32
    static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + STATS_BYTES;
};

// Tiles are walked group by group: a group is `group_m` consecutive row-blocks x ALL column-blocks, row-block fastest.  While a
// group is in flight its activation panel (group_m x 128 rows x K) is re-read by every wave and should stay in L2, and the weight
// matrix streams through once per group: DRAM reads ~ |X| + |W| * num_m / group_m (host side: pick_group_m).
__device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int group_m, int& m_blk, int& n_blk) {
    const int per_group = group_m * num_n;
    const int g = tile / per_group;
    const int r = tile - g * per_group;
    const int m_first = g * group_m;
    const int gm = min(group_m, num_m - m_first);
    m_blk = m_first + r % gm;
    n_blk = r / gm;
}

template <int BN, int EPI, bool TWO>
__global__ void __launch_bounds__(gemm_threads(EPI), 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                    const GemmKParams p) {
    using Cfg = GemmCfg<BN, TWO>;
    static_assert(!TWO || (EPI != S2V_EPI_CONV_T), "the swapped-operand convolution runs one CTA per tile");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + Cfg::STAGES;
    uint64_t* tmem_full = bars + 2 * Cfg::STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* stat_red = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256);   // [2][4][BN][2]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    constexpr bool TR = (EPI == S2V_EPI_CONV_T);
    constexpr int TILE_M = TR ? BN : TWO ? 2 * GEMM_BM : GEMM_BM;   // problem rows per tile (TR: output positions; TWO: the pair's 256)
    // TWO: both CTAs of a pair walk the same tile sequence; this CTA owns rows [m_blk * 256 + rank * 128, +128) of every tile
    const uint32_t cta_rank = TWO ? cluster_ctarank() : 0u;
    const int first_tile = TWO ? int(blockIdx.x >> 1) : int(blockIdx.x);
    const int tile_step = TWO ? int(gridDim.x >> 1) : int(gridDim.x);
    const int num_m = (p.M + TILE_M - 1) / TILE_M;
    const int num_n = TR ? 1 : (p.N + BN - 1) / BN;
    const int num_tiles = num_m * num_n;
    const int kb1 = (EPI == S2V_EPI_CONV || TR) ? p.taps * p.cin_blocks : (p.K + GEMM_BK - 1) / GEMM_BK;
    const int kb2 = (p.K2 + GEMM_BK - 1) / GEMM_BK;
    const int kb_total = kb1 + kb2;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (kb2) {
            tma_prefetch_desc(&tmA2);
            tma_prefetch_desc(&tmB2);
        }
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            // one arrive per epilogue warp (of both CTAs: the leader's MMA thread waits)
            mbar_init(&tmem_empty[a], (TWO ? 2 : 1) * (gemm_threads(EPI) / 32 - 2));
        }
        fence_barrier_init();
    }
    if (TWO) {
        __syncwarp();
        cluster_sync_all();     // the pair's barriers exist before either CTA signals them
    }
    if (warp == 1) {
        if (TWO) tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS); else tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
                int m_blk, n_blk;
                tile_coords(tile, num_m, num_n, p.group_m, m_blk, n_blk);
                const int m0 = m_blk * TILE_M + int(cta_rank) * GEMM_BM, n0 = n_blk * BN;
                const int nb0 = n0 + int(cta_rank) * (BN / 2);    // TWO: this CTA's half of the tile's B rows
                const int lora_col0 = kb2 ? (n0 / p.lora_group_n) * p.K2 : 0;
                for (int kb = 0; kb < kb_total; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + Cfg::A_BYTES;
                    if (TWO) {
                        // the leader's barrier counts the bytes of BOTH CTAs' loads (tma_load_2d_pair signals it from the peer too)
                        if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                        if (kb < kb1) {
                            if (EPI == S2V_EPI_CONV) {
                                const int tap = kb / p.cin_blocks;
                                tma_load_2d_pair(&tmA, &full_bar[stage], sa, (kb - tap * p.cin_blocks) * GEMM_BK, p.a_row0 + m0 + p.tap_off[tap]);
                            } else {
                                tma_load_2d_pair(&tmA, &full_bar[stage], sa, kb * GEMM_BK, m0);
                            }
                            tma_load_2d_pair(&tmB, &full_bar[stage], sb, kb * GEMM_BK, nb0);
                        } else {
                            const int kk = (kb - kb1) * GEMM_BK;
                            tma_load_2d_pair(&tmA2, &full_bar[stage], sa, lora_col0 + kk, m0);
                            tma_load_2d_pair(&tmB2, &full_bar[stage], sb, kk, nb0);
                        }
                        if (++stage == Cfg::STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                        continue;
                    }
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    if (TR) {   // weights -> M-side slot (128 rows), 256 tap-shifted activation rows -> N-side slot
                        const int tap = kb / p.cin_blocks;
                        tma_load_2d(&tmB, &full_bar[stage], sa, kb * GEMM_BK, 0);
                        tma_load_2d(&tmA, &full_bar[stage], sb, (kb - tap * p.cin_blocks) * GEMM_BK, p.a_row0 + m0 + p.tap_off[tap]);
                    } else if (kb < kb1) {
                        if (EPI == S2V_EPI_CONV) {
                            const int tap = kb / p.cin_blocks;
                            tma_load_2d(&tmA, &full_bar[stage], sa, (kb - tap * p.cin_blocks) * GEMM_BK,
                                        p.a_row0 + m0 + p.tap_off[tap]);
                        } else {
                            tma_load_2d(&tmA, &full_bar[stage], sa, kb * GEMM_BK, m0);
                        }
                        tma_load_2d(&tmB, &full_bar[stage], sb, kb * GEMM_BK, n0);
                    } else {
                        const int kk = (kb - kb1) * GEMM_BK;
                        tma_load_2d(&tmA2, &full_bar[stage], sa, lora_col0 + kk, m0);
                        tma_load_2d(&tmB2, &full_bar[stage], sb, kk, n0);
                    }
                    if (++stage == Cfg::STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        // single elected thread for the whole loop (see attn_tcgen05.cu: `if (lane == 0)` around each tcgen05.mma makes the
        // compiler wrap it in a warp-uniformisation loop)
        constexpr uint32_t idesc = make_idesc_bf16(TWO ? 2 * GEMM_BM : GEMM_BM, BN, 0, 0);
        if (cta_rank == 0 && elect_one()) {      // TWO: the leader CTA's thread issues for the pair
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < kb_total; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint64_t adesc = make_smem_desc_sw128(sa, 16, 1024);
                    const uint64_t bdesc = make_smem_desc_sw128(sa + Cfg::A_BYTES, 16, 1024);
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        // advance 16 bf16 = 32 bytes inside the 128B swizzle row: +2 in the (addr >> 4) field
                        if (TWO) umma_ss_pair(d_tmem, adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc, (kb | k) != 0);
                        else umma_ss(d_tmem, adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc, (kb | k) != 0);
                    }
                    if (TWO) {   // both CTAs' producers (stage free) and both CTAs' epilogue warps (accumulator ready) are told
                        umma_commit_pair(&empty_bar[stage], 3);
                        if (kb == kb_total - 1) umma_commit_pair(&tmem_full[acc], 3);
                    } else {
                        umma_commit(&empty_bar[stage]);
                        if (kb == kb_total - 1) umma_commit(&tmem_full[acc]);
                    }
                    if (++stage == Cfg::STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5; 2..9 for the fused q/k epilogue)
        const int q = warp & 3;  // TMEM lane quarter this warp may touch
        const int ehalf = (warp - 2) >> 2;   // 8 epilogue warps: which half of the tile's head vectors this warp takes
        // fused q/k epilogue: per-warp staging of the parameters every lane reads at the same address — the tile's bias slice
        // (BN bf16) and the four 64-wide LayerNorm vectors — in the (otherwise unused) statistics area: as global loads each use sat
        // behind a full L1/L2 round trip inside loops that are not unrolled (44 % long-scoreboard stalls in ncu)
        uint8_t* ep_s = reinterpret_cast<uint8_t*>(stat_red) + (warp - 2) * 1024;     // [0, 512) bias slice, [512, 1024) nq_w | nq_b | nk_w | nk_b
        if (EPI == S2V_EPI_QKV_NORM_ROPE) {
            const bf16* src = lane < 8 ? p.nq_w : lane < 16 ? p.nq_b : lane < 24 ? p.nk_w : p.nk_b;
            sts128(ep_s + 512 + lane * 16, __ldg(reinterpret_cast<const uint4*>(src) + (lane & 7)));
            __syncwarp();
        }
        int it = 0;
        for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
            int m_blk, n_blk;
            tile_coords(tile, num_m, num_n, p.group_m, m_blk, n_blk);
            if (TWO) m_blk = m_blk * 2 + int(cta_rank);     // this CTA's 128-row block of the pair's tile
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            if (TR) {
                // lane = output channel, column = output position: 2-byte stores, 64 contiguous bytes per warp and position
                const int ch = q * 32 + lane;
                const bool ch_ok = ch < p.N;
                const float bs = (p.bias && ch_ok) ? __bfloat162float(p.bias[ch]) : 0.f;
                const int pos0 = m_blk * TILE_M;
                const int rem = pos0 % (p.Hp * p.Wp);
                int hp = rem / p.Wp, wp = rem - hp * p.Wp;
                float st_s = 0.f, st_q = 0.f;     // this channel's sum / sum of squares over the tile's positions (p.stats)
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + acc * BN + c * 32, v);
                    tmem_ld_wait();
                    // the 32 residual values of the chunk are fetched first (read-only path, all in flight together): fetched
                    // one by one between the stores they serialise — the compiler must assume out and res alias — and the
                    // epilogue, not the MMAs, set the pace (250 instead of 1160 TFLOP/s)
                    float rv[32];
                    const int posc = pos0 + c * 32;
                    if (p.res) {   // branch-free inside: clamped (always valid) addresses, so the 32 loads issue back to back
                        const bf16* rp = p.res + (long long)p.out_row0 * p.ldres + (ch_ok ? ch : 0);
                        unsigned short rraw[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            rraw[j] = __ldg(reinterpret_cast<const unsigned short*>(rp + (long long)min(posc + j, p.M - 1) * p.ldres));
#pragma unroll
                        for (int j = 0; j < 32; ++j) rv[j] = __uint_as_float(uint32_t(rraw[j]) << 16);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) rv[j] = 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int pos = posc + j;
                        const bool brd = (hp == 0) | (hp == p.Hp - 1) | (wp == 0) | (wp == p.Wp - 1);
                        if (pos < p.M && ch_ok) {
                            float f = __uint_as_float(v[j]) * p.alpha;
                            f += bs;
                            f += rv[j];
                            if (brd) f = 0.f;
                            const bf16 fo = __float2bfloat16(f);
                            p.out[(long long)(pos + p.out_row0) * p.ldo + ch] = fo;
                            const float fr = __bfloat162float(fo);
                            st_s += fr;
                            st_q = fmaf(fr, fr, st_q);
                        }
                        if (++wp == p.Wp) {
                            wp = 0;
                            if (++hp == p.Hp) hp = 0;
                        }
                    }
                }
                if (p.stats && ch_ok) {
                    p.stats[((long long)m_blk * p.N + ch) * 2] = st_s;
                    p.stats[((long long)m_blk * p.N + ch) * 2 + 1] = st_q;
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_relaxed(&tmem_empty[acc]);
                continue;
            }
            const int row = m_blk * GEMM_BM + q * 32 + lane;
            const bool row_ok = row < p.M;
            const int n0 = n_blk * BN;
            bf16* orow = p.out + (long long)(row + (EPI == S2V_EPI_CONV ? p.out_row0 : 0)) * p.ldo + n0;
            bool border = false;
            const bf16* rrow = nullptr;
            if (EPI == S2V_EPI_CONV && row_ok) {
                const int rem = row % (p.Hp * p.Wp);
                const int hp = rem / p.Wp, wp = rem - hp * p.Wp;
                border = (hp == 0) | (hp == p.Hp - 1) | (wp == 0) | (wp == p.Wp - 1);
                if (p.res) rrow = p.res + (long long)(row + p.out_row0) * p.ldres + n0;
            }
            const float* gate = nullptr;
            if (EPI == S2V_EPI_GATE_RESIDUAL && row_ok) {
                const int b = row / p.rows_per_batch;
                const int s = row - b * p.rows_per_batch;
                gate = p.mod + (long long)b * p.mod_stride + (s < p.text_len ? p.gate_off_text : p.gate_off_other) + n0;
            }
            if (EPI == S2V_EPI_QKV_NORM_ROPE) {
                // Head vectors are 64 columns; a tile may hold q, k and v heads (small models).  Pass 1 computes each q/k head
                // vector's LayerNorm statistics (v head vectors are finished there); pass 2 walks the tile in 16-column chunks so
                // that the row's cos / sin (uncoalesced: one table row per thread) are read ONCE per tile instead of once per head
                // vector.  Same arithmetic as head_norm_rope / qk_norm_rope_kernel.
                constexpr int NH = BN / 64;
                if (lane * 8 < BN) {
                    uint4 bz = make_uint4(0u, 0u, 0u, 0u);
                    if (p.bias && n0 + lane * 8 < p.N) bz = __ldg(reinterpret_cast<const uint4*>(p.bias + n0) + lane);
                    sts128(ep_s + lane * 16, bz);
                }
                __syncwarp();
                const int sidx = row_ok ? row % p.rows_per_batch : 0;
                const bool rope = p.rope_cos != nullptr && sidx >= p.text_len;
                const float* cs = p.rope_cos + (long long)(rope ? sidx - p.text_len : 0) * 64;
                const float* sn = p.rope_sin + (long long)(rope ? sidx - p.text_len : 0) * 64;
                const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;
                const bool qk_tile = n0 + (ehalf * (BN / 128)) * 64 < p.qk_cols;   // at least this warp's first head vector is a q or k head
                float mean[NH], rstd[NH];
                constexpr int NHW = NH / 2;                 // head vectors per epilogue warp
                const int hv_lo = ehalf * NHW, hv_hi = hv_lo + NHW;
#pragma unroll 1
                for (int hv = hv_lo; hv < hv_hi; ++hv) {
                    uint32_t v0[32], v1[32];
                    tmem_ld32(trow + hv * 64, v0);
                    tmem_ld32(trow + hv * 64 + 32, v1);
                    tmem_ld_wait();
                    const int col0 = n0 + hv * 64;
                    float f[64];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        f[j] = __uint_as_float(v0[j]) * p.alpha;
                        f[32 + j] = __uint_as_float(v1[j]) * p.alpha;
                    }
                    if (p.bias && col0 < p.N) {
#pragma unroll
                        for (int v8 = 0; v8 < 8; ++v8) {
                            const uint4 bb = lds128(ep_s + (hv * 64 + v8 * 8) * 2);
                            const uint32_t bw[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                f[v8 * 8 + 2 * j] += bf16_lo(bw[j]);
                                f[v8 * 8 + 2 * j + 1] += bf16_hi(bw[j]);
                            }
                        }
                    }
                    if (col0 < p.qk_cols) {
#pragma unroll
                        for (int j = 0; j < 64; ++j) f[j] = __bfloat162float(__float2bfloat16(f[j]));   // the projection output is bf16
                        float g[8];
#pragma unroll
                        for (int l = 0; l < 8; ++l) {
                            float a = 0.f;
#pragma unroll
                            for (int j = 0; j < 8; ++j) a += f[l * 8 + j];
                            g[l] = a;
                        }
                        const float m = (((g[0] + g[1]) + (g[2] + g[3])) + ((g[4] + g[5]) + (g[6] + g[7]))) * (1.0f / 64.0f);
#pragma unroll
                        for (int l = 0; l < 8; ++l) {
                            float a = 0.f;
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float d = f[l * 8 + j] - m;
                                a = fmaf(d, d, a);
                            }
                            g[l] = a;
                        }
                        mean[hv] = m;
                        rstd[hv] = rsqrtf((((g[0] + g[1]) + (g[2] + g[3])) + ((g[4] + g[5]) + (g[6] + g[7]))) * (1.0f / 64.0f) + p.qk_eps);
                    } else if (row_ok && col0 < p.N) {   // v columns: bias only
#pragma unroll
                        for (int v8 = 0; v8 < 8; ++v8) {
                            uint4 o;
                            o.x = pack_bf16x2(f[v8 * 8 + 0], f[v8 * 8 + 1]);
                            o.y = pack_bf16x2(f[v8 * 8 + 2], f[v8 * 8 + 3]);
                            o.z = pack_bf16x2(f[v8 * 8 + 4], f[v8 * 8 + 5]);
                            o.w = pack_bf16x2(f[v8 * 8 + 6], f[v8 * 8 + 7]);
                            *reinterpret_cast<uint4*>(orow + hv * 64 + v8 * 8) = o;
                        }
                    }
                }
                if (qk_tile) {
#pragma unroll 1
                    for (int c16 = 0; c16 < 4; ++c16) {
                        float cc[16], ss[16];
                        if (rope) {
#pragma unroll
                            for (int v4 = 0; v4 < 4; ++v4) {
                                const float4 c4 = __ldg(reinterpret_cast<const float4*>(cs) + c16 * 4 + v4);
                                const float4 s4 = __ldg(reinterpret_cast<const float4*>(sn) + c16 * 4 + v4);
                                cc[v4 * 4] = c4.x; cc[v4 * 4 + 1] = c4.y; cc[v4 * 4 + 2] = c4.z; cc[v4 * 4 + 3] = c4.w;
                                ss[v4 * 4] = s4.x; ss[v4 * 4 + 1] = s4.y; ss[v4 * 4 + 2] = s4.z; ss[v4 * 4 + 3] = s4.w;
                            }
                        }
#pragma unroll 1
                        for (int hv = hv_lo; hv < hv_hi; ++hv) {
                            uint32_t v[16];
                            tmem_ld16(trow + hv * 64 + c16 * 16, v);
                            tmem_ld_wait();
                            const int col0 = n0 + hv * 64 + c16 * 16;
                            if (!(row_ok && col0 < p.qk_cols)) continue;   // v head vectors were stored by pass 1
                            const bool is_q = col0 < (p.qk_cols >> 1);
                            const uint8_t* nw = ep_s + (is_q ? 512 : 768);        // staged nq_w | nq_b | nk_w | nk_b
                            const uint8_t* nb = nw + 128;
                            float wf[16], bfv[16];
#pragma unroll
                            for (int v8 = 0; v8 < 2; ++v8) {   // same address in every lane: one shared-memory wavefront each
                                const uint4 wu = lds128(nw + (c16 * 2 + v8) * 16);
                                const uint4 bu = lds128(nb + (c16 * 2 + v8) * 16);
                                const uint32_t ww[4] = {wu.x, wu.y, wu.z, wu.w}, bw[4] = {bu.x, bu.y, bu.z, bu.w};
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    wf[v8 * 8 + 2 * j] = bf16_lo(ww[j]); wf[v8 * 8 + 2 * j + 1] = bf16_hi(ww[j]);
                                    bfv[v8 * 8 + 2 * j] = bf16_lo(bw[j]); bfv[v8 * 8 + 2 * j + 1] = bf16_hi(bw[j]);
                                }
                            }
                            float f[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) * p.alpha;
                            if (p.bias) {
#pragma unroll
                                for (int v8 = 0; v8 < 2; ++v8) {
                                    const uint4 bb = lds128(ep_s + (hv * 64 + c16 * 16 + v8 * 8) * 2);
                                    const uint32_t bw[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        f[v8 * 8 + 2 * j] += bf16_lo(bw[j]);
                                        f[v8 * 8 + 2 * j + 1] += bf16_hi(bw[j]);
                                    }
                                }
                            }
                            const float m = mean[hv], r = rstd[hv];
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const float xb = __bfloat162float(__float2bfloat16(f[j]));
                                f[j] = fmaf((xb - m) * r, wf[j], bfv[j]);
                            }
                            if (rope) {
#pragma unroll
                                for (int j = 0; j < 16; j += 2) {
                                    const float x0 = __bfloat162float(__float2bfloat16(f[j])), x1 = __bfloat162float(__float2bfloat16(f[j + 1]));
                                    f[j] = fmaf(x0, cc[j], -(x1 * ss[j]));
                                    f[j + 1] = fmaf(x1, cc[j + 1], x0 * ss[j + 1]);
                                }
                            }
#pragma unroll
                            for (int v8 = 0; v8 < 2; ++v8) {
                                uint4 o;
                                o.x = pack_bf16x2(f[v8 * 8 + 0], f[v8 * 8 + 1]);
                                o.y = pack_bf16x2(f[v8 * 8 + 2], f[v8 * 8 + 3]);
                                o.z = pack_bf16x2(f[v8 * 8 + 4], f[v8 * 8 + 5]);
                                o.w = pack_bf16x2(f[v8 * 8 + 6], f[v8 * 8 + 7]);
                                *reinterpret_cast<uint4*>(orow + hv * 64 + c16 * 16 + v8 * 8) = o;
                            }
                        }
                    }
                }
            } else
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t v[32];
                tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + acc * BN + c * 32, v);
                tmem_ld_wait();
                const int col0 = n0 + c * 32;
                float sv[32];      // S2V_EPI_CONV with p.stats: the bf16-rounded values this lane stores (0 where it stores nothing)
                if (EPI == S2V_EPI_CONV) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) sv[j] = 0.f;
                }
                if (row_ok && col0 < p.N) {
#pragma unroll
                    for (int g8 = 0; g8 < 4; ++g8) {
                        const int col = col0 + g8 * 8;
                        if (col < p.N) {
                            float f[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[g8 * 8 + j]) * p.alpha;
                            if (p.bias) {
                                const uint4 bb = __ldg(reinterpret_cast<const uint4*>(p.bias + col));
                                const uint32_t bw[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    f[2 * j] += bf16_lo(bw[j]);
                                    f[2 * j + 1] += bf16_hi(bw[j]);
                                }
                            }
                            if (EPI == S2V_EPI_BIAS_GELU) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) f[j] = gelu_tanh(f[j]);
                            }
                            if (EPI == S2V_EPI_GATE_RESIDUAL) {
                                const float4 g0 = __ldg(reinterpret_cast<const float4*>(gate + c * 32 + g8 * 8));
                                const float4 g1 = __ldg(reinterpret_cast<const float4*>(gate + c * 32 + g8 * 8 + 4));
                                const uint4 rr = *reinterpret_cast<const uint4*>(orow + c * 32 + g8 * 8);
                                const uint32_t rw[4] = {rr.x, rr.y, rr.z, rr.w};
                                const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    f[2 * j] = bf16_lo(rw[j]) + gg[2 * j] * f[2 * j];
                                    f[2 * j + 1] = bf16_hi(rw[j]) + gg[2 * j + 1] * f[2 * j + 1];
                                }
                            }
                            if (EPI == S2V_EPI_CONV) {
                                if (rrow) {
                                    const uint4 rr = *reinterpret_cast<const uint4*>(rrow + c * 32 + g8 * 8);
                                    const uint32_t rw[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        f[2 * j] += bf16_lo(rw[j]);
                                        f[2 * j + 1] += bf16_hi(rw[j]);
                                    }
                                }
                                if (border) {
#pragma unroll
                                    for (int j = 0; j < 8; ++j) f[j] = 0.f;
                                }
                            }
                            uint4 o;
                            o.x = pack_bf16x2(f[0], f[1]);
                            o.y = pack_bf16x2(f[2], f[3]);
                            o.z = pack_bf16x2(f[4], f[5]);
                            o.w = pack_bf16x2(f[6], f[7]);
                            *reinterpret_cast<uint4*>(orow + c * 32 + g8 * 8) = o;
                            if (EPI == S2V_EPI_CONV) {
                                const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    sv[g8 * 8 + 2 * j] = bf16_lo(ow[j]);
                                    sv[g8 * 8 + 2 * j + 1] = bf16_hi(ow[j]);
                                }
                            }
                        }
                    }
                }
                if (EPI == S2V_EPI_CONV && p.stats) {
                    // column sums over this warp's 32 rows by a butterfly transpose-reduce: after the step with distance s a lane
                    // keeps s of its columns (the half its bit selects) plus the partner's partial sums for them — 31 shuffles per
                    // quantity instead of 160 — and lane l ends up with column c*32 + l
                    float sq[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) sq[j] = sv[j] * sv[j];
#pragma unroll
                    for (int sdist = 16; sdist >= 1; sdist >>= 1) {
                        const bool upper = (lane & sdist) != 0;
#pragma unroll
                        for (int j = 0; j < sdist; ++j) {
                            const float keep_s = upper ? sv[j + sdist] : sv[j], send_s = upper ? sv[j] : sv[j + sdist];
                            const float keep_q = upper ? sq[j + sdist] : sq[j], send_q = upper ? sq[j] : sq[j + sdist];
                            sv[j] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, sdist);
                            sq[j] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, sdist);
                        }
                    }
                    float* dst = stat_red + ((size_t(acc) * 4 + q) * BN + c * 32 + lane) * 2;
                    dst[0] = sv[0];
                    dst[1] = sq[0];
                }
            }
            if (EPI == S2V_EPI_CONV && p.stats) {
                // the four epilogue warps' partial sums -> one row-block entry, in a fixed order (deterministic); the buffer is
                // double-buffered by accumulator parity, so one named barrier per tile is enough
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int col = (warp - 2) * 32 + lane; col < BN; col += 128) {
                    if (n0 + col < p.N) {
                        float a0 = 0.f, a1 = 0.f;
#pragma unroll
                        for (int w4 = 0; w4 < 4; ++w4) {
                            const float* srcp = stat_red + ((size_t(acc) * 4 + w4) * BN + col) * 2;
                            a0 += srcp[0];
                            a1 += srcp[1];
                        }
                        p.stats[((long long)m_blk * p.N + n0 + col) * 2] = a0;
                        p.stats[((long long)m_blk * p.N + n0 + col) * 2 + 1] = a1;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                // relaxed: the accumulator has been read (tcgen05.wait::ld); the row stores need not be visible to the issuing thread
                if (TWO) mbar_arrive_cluster_relaxed(&tmem_empty[acc], 0); else mbar_arrive_relaxed(&tmem_empty[acc]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (TWO) cluster_sync_all();    // neither CTA frees its TMEM or retires while the pair's MMAs / barrier arrivals can still touch it
    if (warp == 1) {
        tc_fence_after();
        if (TWO) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS); else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------------ host side
struct ConvExtra {
    int taps, cin, a_row0, out_row0, Hp, Wp;
    int64_t a_rows;          // rows of the A volume (TMA bound)
    const int* tap_off;
    const void* res;
    long long ldres;
    float* stats;
};

// Rasterisation group height (in 128-row blocks).  DRAM reads of a launch ~ |X| + |W| * num_m / group_m as long as the group's
// activation panel (group_m x 128 rows x K bf16) survives in L2 while the weights stream past it; measured on B200 with ncu
// (profiles/r02_summary.md, "GEMM rasterisation"): K = 3072 + LoRA: 16 -> 1.72 GB, 32 -> 1.04, 40 -> 0.89, 48 -> 1.01, 64 -> 1.74 (panel
// no longer fits), 96 -> 5.4 GB for the FFN-up projection (0.34 GB of operands) — the optimum is a ~31 MB panel.  When the whole
// weight matrix is small enough to stay resident anyway (out-projection: 19 MB) short groups are better (12: 0.51 GB, 48: 0.72 GB).
// S2V_GEMM_GROUP_M overrides for measurements.
static int pick_group_m(int M, int N, int K) {
    static const int forced = [] { const char* e = getenv("S2V_GEMM_GROUP_M"); return e ? atoi(e) : 0; }();
    if (forced > 0) return forced;
    (void)M;
    if ((long long)N * K * 2 <= (32ll << 20)) return 12;
    const long long panel_row = 128ll * K * 2;
    long long g = ((31ll << 20) + panel_row / 2) / panel_row;
    return int(g < 8 ? 8 : g > 64 ? 64 : g);
}

// The big projections run as CTA pairs (cta_group::2, 256 x 256 tiles): per MMA each SM stages and reads half of the B rows, so
// shared-memory operand traffic and L2 -> SM traffic per flop drop by a third against the 128 x 256 single-CTA tile — under the
// 1 kW cap that is what the sustained GEMM rate responds to (profiles/r02_summary.md).  S2V_GEMM_2CTA=0 keeps single CTAs (A/B).
static bool use_cta_pairs(int M, int N) {
    static const int forced = [] { const char* e = getenv("S2V_GEMM_2CTA"); return e ? atoi(e) : -1; }();
    if (forced >= 0) return forced != 0;
    return M >= 256 && N >= 256;     // (also the T5 encoder's 452-row projections: 8.58 -> 8.06 ms per encode)
}

// Persistent CTAs (SMs) a projection GEMM occupies.  S2V_GEMM_SMS overrides for the power-density experiments of
// profiles/r02_summary.md ("the board's power governor"): fewer SMs = lower instantaneous power at a given clock.
static int gemm_sms() {
    static const int forced = [] { const char* e = getenv("S2V_GEMM_SMS"); return e ? atoi(e) : 0; }();
    const int n = sm_count();
    return (forced > 0 && forced < n) ? forced : n;
}

template <int BN, int EPI, bool TWO = false>
static int launch_gemm(const s2v_linear_args* a, cudaStream_t stream, const ConvExtra* conv = nullptr, const s2v_qk_norm_args* qk = nullptr) {
    using Cfg = GemmCfg<BN, TWO>;
    CUtensorMap tmA, tmB, tmA2, tmB2;
    int rc;
    constexpr bool TR = (EPI == S2V_EPI_CONV_T);   // operands swapped: BN activation rows per box, 128 weight rows
    constexpr int B_BOX = TR ? GEMM_BM : TWO ? BN / 2 : BN;
    if ((rc = make_tmap_2d_bf16(&tmA, a->x, conv ? conv->cin : a->K, conv ? conv->a_rows : a->M, a->ldx, GEMM_BK, TR ? BN : GEMM_BM))) return rc;
    if ((rc = make_tmap_2d_bf16(&tmB, a->w, a->K, a->N, a->ldw, GEMM_BK, B_BOX))) return rc;
    const int K2 = a->lora_t ? a->lora_r : 0;
    if (K2) {
        const int groups = (a->N + a->lora_group_n - 1) / a->lora_group_n;
        if ((rc = make_tmap_2d_bf16(&tmA2, a->lora_t, (int64_t)groups * K2, a->M, a->ldt, GEMM_BK, GEMM_BM))) return rc;
        if ((rc = make_tmap_2d_bf16(&tmB2, a->lora_b, K2, a->N, a->ldb, GEMM_BK, B_BOX))) return rc;
    } else {
        tmA2 = tmA;
        tmB2 = tmB;
    }
    GemmKParams p;
    p.M = a->M; p.N = a->N; p.K = a->K; p.K2 = K2;
    p.group_m = pick_group_m(a->M, a->N, a->K + K2);
    p.lora_group_n = a->lora_group_n > 0 ? a->lora_group_n : a->N;
    p.bias = static_cast<const bf16*>(a->bias);
    p.out = static_cast<bf16*>(a->out);
    p.ldo = a->ldo;
    p.alpha = a->alpha;
    p.mod = a->mod;
    p.mod_stride = a->mod_stride; p.gate_off_text = a->gate_off_text; p.gate_off_other = a->gate_off_other;
    p.rows_per_batch = a->rows_per_batch > 0 ? a->rows_per_batch : a->M; p.text_len = a->text_len;
    p.taps = 1; p.cin_blocks = (a->K + GEMM_BK - 1) / GEMM_BK; p.a_row0 = 0; p.out_row0 = 0; p.Hp = 1; p.Wp = 1;
    p.res = nullptr; p.ldres = 0; p.stats = nullptr;
    p.nq_w = p.nq_b = p.nk_w = p.nk_b = nullptr; p.rope_cos = p.rope_sin = nullptr; p.qk_cols = 0; p.qk_eps = 0.f;
    if (qk) {
        p.nq_w = static_cast<const bf16*>(qk->nq_w); p.nq_b = static_cast<const bf16*>(qk->nq_b);
        p.nk_w = static_cast<const bf16*>(qk->nk_w); p.nk_b = static_cast<const bf16*>(qk->nk_b);
        p.rope_cos = qk->cos; p.rope_sin = qk->sin; p.qk_cols = 2 * qk->H * 64; p.qk_eps = qk->eps;
        p.rows_per_batch = qk->S; p.text_len = qk->text_len;
    }
    if (conv) {
        p.taps = conv->taps; p.cin_blocks = (conv->cin + GEMM_BK - 1) / GEMM_BK; p.a_row0 = conv->a_row0; p.out_row0 = conv->out_row0;
        p.Hp = conv->Hp; p.Wp = conv->Wp; p.res = static_cast<const bf16*>(conv->res); p.ldres = conv->ldres; p.stats = conv->stats;
        for (int i = 0; i < 27; ++i) p.tap_off[i] = i < conv->taps ? conv->tap_off[i] : 0;
    }

    auto kern = gemm_tcgen05_kernel<BN, EPI, TWO>;
    if ((rc = ensure_smem_optin(reinterpret_cast<const void*>(kern), Cfg::SMEM_BYTES, "cudaFuncSetAttribute(gemm)"))) return rc;
    constexpr int TILE_M = TWO ? 2 * GEMM_BM : GEMM_BM;
    const int num_tiles = TR ? (a->M + BN - 1) / BN : ((a->M + TILE_M - 1) / TILE_M) * ((a->N + BN - 1) / BN);
    if (TWO) {
        if (p.group_m > 1) p.group_m = (p.group_m + 1) / 2;      // group height is counted in tiles (256-row pair-blocks here)
        const int sms = conv ? sm_count() : gemm_sms();
        const int clusters = num_tiles < sms / 2 ? num_tiles : sms / 2;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * clusters);
        cfg.blockDim = dim3(gemm_threads(EPI));
        cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmA2, tmB2, p);
        if (e != cudaSuccess) return set_cuda_error(e, "gemm_tcgen05_kernel(cta pair)");
        return check_launch("gemm_tcgen05_kernel(cta pair)");
    }
    const int sms1 = conv ? sm_count() : gemm_sms();
    const int grid = num_tiles < sms1 ? num_tiles : sms1;
    kern<<<grid, gemm_threads(EPI), Cfg::SMEM_BYTES, stream>>>(tmA, tmB, tmA2, tmB2, p);
    return check_launch("gemm_tcgen05_kernel");
}

}  // namespace s2v

using namespace s2v;

extern "C" int s2v_linear(const s2v_linear_args* a, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!a || !a->x || !a->w || !a->out) return set_error(S2V_E_BADARG, "s2v_linear: null pointer");
    if (a->M <= 0 || a->N <= 0 || a->K <= 0) return set_error(S2V_E_BADARG, "s2v_linear: empty problem");
    if ((a->K % 8) || (a->N % 8) || (a->ldx % 8) || (a->ldw % 8) || (a->ldo % 8))
        return set_error(S2V_E_UNSUPPORTED, "s2v_linear: K, N and leading dims must be multiples of 8 (16-byte rows)");
    if (a->lora_t) {
        if (!a->lora_b || a->lora_r <= 0 || (a->lora_r % 8) || (a->ldt % 8) || (a->ldb % 8))
            return set_error(S2V_E_BADARG, "s2v_linear: bad LoRA operands");
        const int gn = a->lora_group_n > 0 ? a->lora_group_n : a->N;
        if (gn % 128 && gn != a->N) return set_error(S2V_E_UNSUPPORTED, "s2v_linear: lora_group_n must be a multiple of 128");
    }
    if (a->epilogue == S2V_EPI_GATE_RESIDUAL && !a->mod) return set_error(S2V_E_BADARG, "s2v_linear: gate table missing");
    int rc = ensure_device();
    if (rc) return rc;
    // 256-wide tiles for the big projections, 128-wide when N is small or a LoRA group boundary is not 256-aligned
    const int gn = (a->lora_t && a->lora_group_n > 0) ? a->lora_group_n : a->N;
    bool wide = (a->N >= 256) && (gn % 256 == 0 || !a->lora_t || gn == a->N);
    // few rows (the T5 prompt encoder: 452): 256-wide tiles would leave SMs without a tile — the launch is weight-bandwidth
    // bound, so more, narrower tiles stream the weights through more SMs
    if (wide && (long long)((a->M + GEMM_BM - 1) / GEMM_BM) * ((a->N + 255) / 256) < sm_count()) wide = false;
    const bool two = wide && use_cta_pairs(a->M, a->N);
    switch (a->epilogue) {
        case S2V_EPI_BIAS:
            if (two) return launch_gemm<256, S2V_EPI_BIAS, true>(a, stream);
            return wide ? launch_gemm<256, S2V_EPI_BIAS>(a, stream) : launch_gemm<128, S2V_EPI_BIAS>(a, stream);
        case S2V_EPI_BIAS_GELU:
            if (two) return launch_gemm<256, S2V_EPI_BIAS_GELU, true>(a, stream);
            return wide ? launch_gemm<256, S2V_EPI_BIAS_GELU>(a, stream) : launch_gemm<128, S2V_EPI_BIAS_GELU>(a, stream);
        case S2V_EPI_GATE_RESIDUAL:
            if (two) return launch_gemm<256, S2V_EPI_GATE_RESIDUAL, true>(a, stream);
            return wide ? launch_gemm<256, S2V_EPI_GATE_RESIDUAL>(a, stream)
                        : launch_gemm<128, S2V_EPI_GATE_RESIDUAL>(a, stream);
        default:
            return set_error(S2V_E_BADARG, "s2v_linear: unknown epilogue");
    }
}

extern "C" int s2v_qkv_lora(const s2v_linear_args* a, void* stream) { return s2v_linear(a, stream); }

extern "C" int s2v_qkv_lora_norm_rope(const s2v_linear_args* a, const s2v_qk_norm_args* qk, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!a || !qk || !a->x || !a->w || !a->out || !qk->nq_w || !qk->nq_b || !qk->nk_w || !qk->nk_b)
        return set_error(S2V_E_BADARG, "s2v_qkv_lora_norm_rope: null pointer");
    if (a->M <= 0 || a->K <= 0 || qk->H <= 0 || qk->S <= 0 || a->N != 3 * qk->H * 64 || (a->M % qk->S))
        return set_error(S2V_E_BADARG, "s2v_qkv_lora_norm_rope: expected N = 3*H*64 and M = B*S");
    if ((a->K % 8) || (a->ldx % 8) || (a->ldw % 8) || (a->ldo % 8)) return set_error(S2V_E_UNSUPPORTED, "s2v_qkv_lora_norm_rope: K and leading dims must be multiples of 8");
    if ((qk->cos == nullptr) != (qk->sin == nullptr)) return set_error(S2V_E_BADARG, "s2v_qkv_lora_norm_rope: cos and sin go together");
    if (a->lora_t) {
        if (!a->lora_b || a->lora_r <= 0 || (a->lora_r % 8) || (a->ldt % 8) || (a->ldb % 8) || (a->lora_group_n != qk->H * 64))
            return set_error(S2V_E_BADARG, "s2v_qkv_lora_norm_rope: bad LoRA operands (one group per q, k, v)");
    }
    int rc = ensure_device();
    if (rc) return rc;
    const int gn = qk->H * 64;
    const bool wide = (gn % 256 == 0) || !a->lora_t;
    if (wide && use_cta_pairs(a->M, a->N)) return launch_gemm<256, S2V_EPI_QKV_NORM_ROPE, true>(a, stream, nullptr, qk);
    return wide ? launch_gemm<256, S2V_EPI_QKV_NORM_ROPE>(a, stream, nullptr, qk) : launch_gemm<128, S2V_EPI_QKV_NORM_ROPE>(a, stream, nullptr, qk);
}
extern "C" int s2v_outproj_lora_gate_residual(const s2v_linear_args* a, void* stream) {
    if (a && a->epilogue != S2V_EPI_GATE_RESIDUAL) return set_error(S2V_E_BADARG, "outproj: epilogue must be GATE_RESIDUAL");
    return s2v_linear(a, stream);
}
extern "C" int s2v_ffn_up_gelu_lora(const s2v_linear_args* a, void* stream) {
    if (a && a->epilogue != S2V_EPI_BIAS_GELU) return set_error(S2V_E_BADARG, "ffn_up: epilogue must be BIAS_GELU");
    return s2v_linear(a, stream);
}
extern "C" int s2v_ffn_down_lora_gate_residual(const s2v_linear_args* a, void* stream) {
    if (a && a->epilogue != S2V_EPI_GATE_RESIDUAL) return set_error(S2V_E_BADARG, "ffn_down: epilogue must be GATE_RESIDUAL");
    return s2v_linear(a, stream);
}

// ------------------------------------------------------------------------------------------------ implicit-GEMM convolution
extern "C" int s2v_conv_gemm(const s2v_conv_args* c, void* stream) { return s2v_conv_gemm_stats(c, nullptr, nullptr, stream); }

extern "C" int s2v_conv_gemm_stats(const s2v_conv_args* c, float* stats_partial, int32_t* row_blocks, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!c || !c->x || !c->w || !c->out) return set_error(S2V_E_BADARG, "s2v_conv_gemm: null pointer");
    if (stats_partial && !row_blocks) return set_error(S2V_E_BADARG, "s2v_conv_gemm_stats: row_blocks must be given with stats_partial");
    if (c->taps != 1 && c->taps != 9 && c->taps != 27) return set_error(S2V_E_BADARG, "s2v_conv_gemm: taps must be 1, 9 or 27");
    if (c->T <= 0 || c->Hp < 3 || c->Wp < 3 || c->cin <= 0 || c->cout <= 0) return set_error(S2V_E_BADARG, "s2v_conv_gemm: empty problem");
    if ((c->cin % 8) || (c->cout % 8) || (c->ldx % 8) || (c->ldo % 8) || (c->ldw % 8) || (c->res && (c->ldres % 8)))
        return set_error(S2V_E_UNSUPPORTED, "s2v_conv_gemm: channel counts and leading dims must be multiples of 8");
    if (c->taps > 1 && (c->cin % 64)) return set_error(S2V_E_UNSUPPORTED, "s2v_conv_gemm: 3x3(x3) taps need Cin % 64 == 0");
    int rc = ensure_device();
    if (rc) return rc;
    const long long plane = (long long)c->Hp * c->Wp;
    const long long M = (long long)c->T * plane;
    if (M > 0x7fffffffLL || (long long)(c->T + c->t_pad) * plane > 0x7fffffffLL)
        return set_error(S2V_E_UNSUPPORTED, "s2v_conv_gemm: more than 2^31 rows");
    int off[27];
    if (c->taps == 27) {
        for (int dt = 0; dt < 3; ++dt)
            for (int dh = 0; dh < 3; ++dh)
                for (int dw = 0; dw < 3; ++dw) off[(dt * 3 + dh) * 3 + dw] = (int)((dt - 2) * plane + (dh - 1) * c->Wp + (dw - 1));
    } else if (c->taps == 9) {
        for (int dh = 0; dh < 3; ++dh)
            for (int dw = 0; dw < 3; ++dw) off[dh * 3 + dw] = (dh - 1) * c->Wp + (dw - 1);
    } else {
        off[0] = 0;
    }
    s2v_linear_args a;
    memset(&a, 0, sizeof(a));
    a.x = c->x; a.ldx = c->ldx; a.w = c->w; a.ldw = c->ldw; a.bias = c->bias; a.out = c->out; a.ldo = c->ldo;
    a.M = (int)M; a.N = c->cout; a.K = c->taps * c->cin; a.epilogue = S2V_EPI_CONV; a.alpha = 1.0f;
    ConvExtra ex;
    ex.taps = c->taps; ex.cin = c->cin; ex.a_row0 = (int)(c->t_pad * plane); ex.out_row0 = (int)(c->t_pad * plane);
    ex.Hp = c->Hp; ex.Wp = c->Wp; ex.a_rows = (long long)(c->T + c->t_pad) * plane; ex.tap_off = off; ex.res = c->res; ex.ldres = c->ldres;
    ex.stats = stats_partial;
    if (row_blocks) *row_blocks = (int32_t)((M + 127) / 128);      // the swapped-operand form below overrides (256 positions per block)
    if (c->cout >= 256) {
        // CTA pairs (256 positions x 256 channels per tile) like the projections; S2V_GEMM_2CTA=0 keeps single CTAs
        if (use_cta_pairs((int)M, c->cout)) return launch_gemm<256, S2V_EPI_CONV, true>(&a, stream, &ex);
        return launch_gemm<256, S2V_EPI_CONV>(&a, stream, &ex);
    }
    // Cout <= 128: one N block.  With a 3x3(x3) kernel the swapped form (weights on the M side,
    // 256 positions on the N side) halves the shared-memory operand traffic per flop; S2V_CONV_T=0 keeps the plain form (A/B).
    static const bool conv_t = [] { const char* e = getenv("S2V_CONV_T"); return !(e && e[0] == '0'); }();
    if (conv_t && c->cout <= 128 && c->taps > 1) {
        if (row_blocks) *row_blocks = (int32_t)((M + 255) / 256);
        return launch_gemm<256, S2V_EPI_CONV_T>(&a, stream, &ex);
    }
    return launch_gemm<128, S2V_EPI_CONV>(&a, stream, &ex);
}

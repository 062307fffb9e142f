// Joint (text | reference image | video) self-attention forward for sm_100a, head_dim 64, no mask.
//
// One CTA owns one (batch, head) and TWO 128-row query tiles (256 query rows) and streams all S keys in 128-key tiles:
//
//   warp 0        : TMA producer   — Q0,Q1 once, then K(j),V(j) tiles through a KV_STAGES-deep mbarrier ring
//   warp 1        : tcgen05.mma issuer (single thread)
//                     S_q(j)  = Q_q K(j)^T        SS-MMA  M128 N128 K64   -> TMEM columns [q*128, q*128+128)  (fp32)
//                     O_q    += P_q(j) V(j)       TS-MMA  M128 N64  K128  -> TMEM columns [256+q*64, +64)     (fp32)
//                   P_q(j) (bf16) aliases the first 64 columns of S_q; V is consumed straight from its [key][d] TMA
//                   layout as an MN-major B operand, so no transpose is ever materialised.
//   warps 4..7    : softmax warpgroup for query tile 0   (one thread = one query row = one TMEM lane)
//   warps 8..11   : softmax warpgroup for query tile 1
//
// The two query tiles ping-pong: while the tensor core runs PV_0(j) + S_0(j+1), warpgroup 1 does softmax on S_1(j),
// and vice versa.  Softmax is the FA-style online form in the exp2 domain with LAZY rescaling: the running reference
// max is only moved (and O, l rescaled in TMEM) when the row max grows by more than 2^8, so the O read-modify-write is
// off the critical path for almost every tile.  exp2 arguments are formed with one FFMA (s*c - m*c).
//
// Global layout: qkv [B, S, 3*H*64] (q|k|v, heads contiguous) read through ONE 4-D tensor map
// {64, 3H, S, B}; out [B, S, H*64].  Rows >= S are zero-filled by TMA; keys >= S are masked to -inf.
#include "common.cuh"
#include "host_util.h"
#include "s2v_b200.h"

namespace s2v {

constexpr int ATT_D = 64;
constexpr int ATT_BQ = 128;         // rows per query tile (UMMA M)
constexpr int ATT_QTILES = 2;       // query tiles per CTA
constexpr int ATT_BK = 128;         // keys per tile
constexpr int ATT_STAGES = 4;       // K/V ring depth
constexpr int ATT_THREADS = 384;    // 12 warps
constexpr uint32_t ATT_TILE_BYTES = ATT_BK * ATT_D * 2;  // 16 KB
constexpr uint32_t ATT_SMEM_BYTES = (ATT_QTILES + 2 * ATT_STAGES) * ATT_TILE_BYTES + 1024 + 256;
constexpr float ATT_RESCALE_THRESHOLD = 8.0f;  // log2 units

constexpr uint32_t TM_S0 = 0, TM_O0 = 256;  // TMEM column map: S_q at q*128, O_q at 256 + q*64

__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, bf16* __restrict__ out, int S, int H, float scale_log2) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                                     // 2 x 16 KB
    uint8_t* sK = smem + ATT_QTILES * ATT_TILE_BYTES;       // STAGES x 16 KB
    uint8_t* sV = sK + ATT_STAGES * ATT_TILE_BYTES;         // STAGES x 16 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT_STAGES * ATT_TILE_BYTES);
    uint64_t* q_full = bars;                  // 1
    uint64_t* kv_full = bars + 1;             // STAGES
    uint64_t* kv_empty = kv_full + ATT_STAGES;  // STAGES
    uint64_t* s_full = kv_empty + ATT_STAGES;   // 2
    uint64_t* p_ready = s_full + 2;           // 2
    uint64_t* o_final = p_ready + 2;          // 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_final + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int qblk = blockIdx.x, head = blockIdx.y, batch = blockIdx.z;
    const int q_row0 = qblk * (ATT_BQ * ATT_QTILES);
    const int n_kv = (S + ATT_BK - 1) / ATT_BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        mbar_init(q_full, 1);
        for (int s = 0; s < ATT_STAGES; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        for (int q = 0; q < 2; ++q) {
            mbar_init(&s_full[q], 1);
            mbar_init(&p_ready[q], 4);  // one arrive per softmax warp
        }
        mbar_init(o_final, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        setmaxnreg_dec<56>();
        if (warp == 0) {
            // ---------------------------------------------------------------- TMA producer
            if (lane == 0) {
                mbar_arrive_expect_tx(q_full, ATT_QTILES * ATT_TILE_BYTES);
                for (int q = 0; q < ATT_QTILES; ++q)
                    tma_load_4d(&tmQKV, q_full, sQ + q * ATT_TILE_BYTES, 0, head, q_row0 + q * ATT_BQ, batch);
                int stage = 0;
                uint32_t phase = 0;
                for (int j = 0; j < n_kv; ++j) {
                    mbar_wait(&kv_empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&kv_full[stage], 2 * ATT_TILE_BYTES);
                    tma_load_4d(&tmQKV, &kv_full[stage], sK + stage * ATT_TILE_BYTES, 0, H + head, j * ATT_BK, batch);
                    tma_load_4d(&tmQKV, &kv_full[stage], sV + stage * ATT_TILE_BYTES, 0, 2 * H + head, j * ATT_BK, batch);
                    if (++stage == ATT_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        } else if (warp == 1) {
            // ---------------------------------------------------------------- MMA issuer
            constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BQ, ATT_BK, 0, 0);   // S = Q K^T   (both K-major)
            constexpr uint32_t idesc_o = make_idesc_bf16(ATT_BQ, ATT_D, 0, 1);    // O += P V    (A in TMEM, B MN-major)
            auto issue_s = [&](int q, int stage) {
                const uint64_t adesc = make_smem_desc_sw128(smem_u32(sQ + q * ATT_TILE_BYTES), 16, 1024);
                const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sK + stage * ATT_TILE_BYTES), 16, 1024);
#pragma unroll
                for (int k = 0; k < ATT_D / 16; ++k)
                    umma_ss(tmem_base + TM_S0 + q * 128, adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc_s, k != 0);
            };
            auto issue_pv = [&](int q, int stage, bool accumulate) {
                // V tile [128 keys][64 d]: 16 keys per MMA = two 8-row groups = 2048 bytes
                const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sV + stage * ATT_TILE_BYTES), 1024, 1024);
#pragma unroll
                for (int k = 0; k < ATT_BK / 16; ++k)
                    umma_ts(tmem_base + TM_O0 + q * 64, tmem_base + TM_S0 + q * 128 + k * 8, bdesc + uint64_t(k * 128),
                            idesc_o, (accumulate || k != 0) ? 1u : 0u);
            };
            mbar_wait(q_full, 0);
            mbar_wait(&kv_full[0], 0);
            tc_fence_after();
            if (lane == 0) {
                issue_s(0, 0);
                umma_commit(&s_full[0]);
                issue_s(1, 0);
                umma_commit(&s_full[1]);
            }
            __syncwarp();
            int stage = 0;
            uint32_t phase = 0;
            for (int j = 0; j < n_kv; ++j) {
                int nstage = stage + 1;
                uint32_t nphase = phase;
                if (nstage == ATT_STAGES) {
                    nstage = 0;
                    nphase ^= 1;
                }
                const bool has_next = (j + 1 < n_kv);
                // ---- query tile 0
                mbar_wait(&p_ready[0], j & 1);
                tc_fence_after();
                if (lane == 0) issue_pv(0, stage, j != 0);
                __syncwarp();
                if (has_next) {
                    mbar_wait(&kv_full[nstage], nphase);
                    tc_fence_after();
                    if (lane == 0) {
                        issue_s(0, nstage);
                        umma_commit(&s_full[0]);
                    }
                    __syncwarp();
                }
                // ---- query tile 1
                mbar_wait(&p_ready[1], j & 1);
                tc_fence_after();
                if (lane == 0) {
                    issue_pv(1, stage, j != 0);
                    umma_commit(&kv_empty[stage]);  // every MMA reading K(j)/V(j) has been issued before this point
                    if (has_next) {
                        issue_s(1, nstage);
                        umma_commit(&s_full[1]);
                    } else {
                        umma_commit(o_final);
                    }
                }
                __syncwarp();
                stage = nstage;
                phase = nphase;
            }
        }
    } else {
        // -------------------------------------------------------------------- softmax warpgroups
        setmaxnreg_inc<224>();
        const int q = (warp - 4) >> 2;          // query tile of this warpgroup
        const int lq = warp & 3;                // TMEM lane quarter
        const uint32_t lane_off = uint32_t(lq * 32) << 16;
        const uint32_t tS = tmem_base + lane_off + TM_S0 + q * 128;
        const uint32_t tO = tmem_base + lane_off + TM_O0 + q * 64;
        const int row = q_row0 + q * ATT_BQ + lq * 32 + lane;

        float m_ref = -INFINITY;   // reference max (raw score units) used in the exponent
        float l_sum = 0.f;

        for (int j = 0; j < n_kv; ++j) {
            mbar_wait(&s_full[q], j & 1);
            tc_fence_after();
            uint32_t s[128];
            {
                uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
                uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
                uint32_t(&s2)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[64]);
                uint32_t(&s3)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[96]);
                tmem_ld32(tS, s0);
                tmem_ld32(tS + 32, s1);
                tmem_ld32(tS + 64, s2);
                tmem_ld32(tS + 96, s3);
                tmem_ld_wait();
            }
            const int valid = S - j * ATT_BK;  // keys valid in this tile (>= 128 except for the last tile)
            if (valid < ATT_BK) {
#pragma unroll
                for (int i = 0; i < 128; ++i)
                    if (i >= valid) s[i] = __float_as_uint(-INFINITY);
            }
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int i = 0; i < 128; i += 4) {
                mx0 = fmaxf(mx0, __uint_as_float(s[i]));
                mx1 = fmaxf(mx1, __uint_as_float(s[i + 1]));
                mx2 = fmaxf(mx2, __uint_as_float(s[i + 2]));
                mx3 = fmaxf(mx3, __uint_as_float(s[i + 3]));
            }
            const float m_tile = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
            if (j == 0) {
                m_ref = m_tile;
            } else {
                const bool need = (m_tile - m_ref) * scale_log2 > ATT_RESCALE_THRESHOLD;
                if (__any_sync(0xffffffffu, need)) {
                    // s_full[q] of tile j was committed after PV_q(j-1): O_q is quiescent here.
                    const float factor = need ? ex2_approx((m_ref - m_tile) * scale_log2) : 1.0f;
                    if (need) m_ref = m_tile;
                    l_sum *= factor;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        uint32_t o[32];
                        tmem_ld32(tO + c * 32, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
                        tmem_st32(tO + c * 32, o);
                    }
                    tmem_st_wait();
                }
            }
            const float mneg = -m_ref * scale_log2;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            uint32_t pk[64];
#pragma unroll
            for (int i = 0; i < 128; i += 4) {
                const float p0 = ex2_approx(fmaf(__uint_as_float(s[i]), scale_log2, mneg));
                const float p1 = ex2_approx(fmaf(__uint_as_float(s[i + 1]), scale_log2, mneg));
                const float p2 = ex2_approx(fmaf(__uint_as_float(s[i + 2]), scale_log2, mneg));
                const float p3 = ex2_approx(fmaf(__uint_as_float(s[i + 3]), scale_log2, mneg));
                a0 += p0; a1 += p1; a2 += p2; a3 += p3;
                pk[i / 2] = pack_bf16x2(p0, p1);
                pk[i / 2 + 1] = pack_bf16x2(p2, p3);
            }
            l_sum += (a0 + a1) + (a2 + a3);
            {
                const uint32_t(&p0)[32] = *reinterpret_cast<const uint32_t(*)[32]>(&pk[0]);
                const uint32_t(&p1)[32] = *reinterpret_cast<const uint32_t(*)[32]>(&pk[32]);
                tmem_st32(tS, p0);
                tmem_st32(tS + 32, p1);
                tmem_st_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_ready[q]);
        }
        // ---- epilogue: O / l -> bf16 -> global
        mbar_wait(o_final, 0);
        tc_fence_after();
        const float inv = 1.0f / l_sum;
        bf16* orow = out + ((long long)batch * S + row) * (long long)(H * ATT_D) + head * ATT_D;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld32(tO + c * 32, o);
            tmem_ld_wait();
            if (row < S) {
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {
                    uint4 w;
                    w.x = pack_bf16x2(__uint_as_float(o[g8 * 8 + 0]) * inv, __uint_as_float(o[g8 * 8 + 1]) * inv);
                    w.y = pack_bf16x2(__uint_as_float(o[g8 * 8 + 2]) * inv, __uint_as_float(o[g8 * 8 + 3]) * inv);
                    w.z = pack_bf16x2(__uint_as_float(o[g8 * 8 + 4]) * inv, __uint_as_float(o[g8 * 8 + 5]) * inv);
                    w.w = pack_bf16x2(__uint_as_float(o[g8 * 8 + 6]) * inv, __uint_as_float(o[g8 * 8 + 7]) * inv);
                    *reinterpret_cast<uint4*>(orow + c * 32 + g8 * 8) = w;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace s2v

using namespace s2v;

extern "C" int s2v_attn_fwd(const void* qkv, void* o, int32_t B, int32_t S, int32_t H, float softmax_scale, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!qkv || !o) return set_error(S2V_E_BADARG, "s2v_attn_fwd: null pointer");
    if (B <= 0 || S <= 0 || H <= 0) return set_error(S2V_E_BADARG, "s2v_attn_fwd: empty problem");
    if (B > 65535 || H > 65535) return set_error(S2V_E_UNSUPPORTED, "s2v_attn_fwd: B and H must fit a grid dimension");
    int rc = ensure_device();
    if (rc) return rc;
    CUtensorMap tm;
    const uint64_t row_bytes = (uint64_t)3 * H * ATT_D * 2;
    const uint64_t dims[4] = {(uint64_t)ATT_D, (uint64_t)3 * H, (uint64_t)S, (uint64_t)B};
    const uint64_t strides[4] = {2, (uint64_t)ATT_D * 2, row_bytes, row_bytes * (uint64_t)S};
    const uint32_t box[4] = {ATT_D, 1, ATT_BK, 1};
    if ((rc = make_tmap_nd_bf16(&tm, qkv, 4, dims, strides, box))) return rc;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES);
        if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(attn)");
        attr_done = true;
    }
    dim3 grid((S + ATT_BQ * ATT_QTILES - 1) / (ATT_BQ * ATT_QTILES), H, B);
    const float scale_log2 = softmax_scale * 1.4426950408889634f;
    attn_fwd_kernel<<<grid, ATT_THREADS, ATT_SMEM_BYTES, stream>>>(tm, static_cast<bf16*>(o), S, H, scale_log2);
    return check_launch("attn_fwd_kernel");
}

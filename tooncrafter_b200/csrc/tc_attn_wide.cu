// tooncrafter_b200 — fused single-head attention with a WIDE head (D = 64..512 channels, D % 64 == 0): the mid-block
// AttnBlock of the VAE encoder / decoder (lvdm/models/autoencoder_dualref.py:172-206, lvdm/modules/networks/
// ae_modules.py AttnBlock): softmax(Q K^T / sqrt(D)) V over the H*W tokens of one frame, D = 512 at full width.
//
// Replaces the unfused path (per frame: score GEMM -> fp16 [L][L] score matrix in HBM -> softmax_rows -> PV GEMM; 48 GEMM
// launches and ~0.4 GB of score traffic per decode) by ONE kernel; scores never leave the SM.
//
//   * O for 128 queries x 512 channels would fill all 512 TMEM columns, so a CTA owns 128 queries x ONE HALF of the value
//     channels (<= 256 columns of O) and the two CTAs of a query tile each compute S = Q K^T themselves (S is 2/3 of a
//     CTA's flops; the pair executes 1.5x the minimal flops and needs no cross-CTA exchange).  D <= 256: one CTA.
//   * TMEM: S/P buffer 0 [0,128), S/P buffer 1 [128,256), O [256, 256 + D_cta).  P (fp16) overwrites the first 64 columns of
//     the S buffer it was computed from and is the TMEM A operand of the PV MMA (as in tc_attn3_kernel); S(g+2) is issued
//     into that buffer once PV(g) has retired.
//   * Q (128 x D, 16 KiB per 64-channel k-block) stays in shared memory; K k-blocks and V sub-tiles ([128 keys][64 channels])
//     stream through TWO rings of 16 KiB slots, each with one producer warp and one consumer (issuer) warp walking it in
//     order.  (One shared ring walked "by position" by two consumers is wrong: mbarrier waits are by phase PARITY, and a
//     consumer that skips the other's slots can reach a slot two fills early, where the parity matches again — stale tiles
//     under load.)
//   * warps 0-3: softmax, one query row per thread (fp32, online, lazy rescale: O is only touched when a row maximum grew
//     by more than 2^8; otherwise PV(g-1) is awaited AFTER block g's exponentials); warp 4: issuer of the S MMAs; warp 5:
//     TMA producer of Q and K; warp 6: issuer of the PV MMAs; warp 7: TMA producer of V.  Two issuers because issuing is
//     the bottleneck of small MMAs: ~100 cycles of issue work per tcgen05.mma against 32-64 cycles of execution (64 MMAs
//     per key block at D = 512; one issuer ran 7200 cycles per block against 3072 tensor cycles).
//   The kernel is tensor-bound by construction (per 128-key block 3072 tensor cycles vs 1024 MUFU cycles at D = 512).
//
// Algorithmic flops = 4 * N * Lq * Lk * D (the S recompute is not counted).
#include "tc_common.cuh"
#include "tc_host.h"

namespace {

constexpr int kQTile = 128;
constexpr int kKVTile = 128;
constexpr int kTileBytes = 128 * 64 * 2;   // 16 KiB: [128 rows][64 halfs], 128B-swizzled
constexpr int kThreads = 256;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemO = 256;
constexpr float kRescaleThreshold = 8.0f;

struct alignas(64) AttnWideParams {
    CUtensorMap tmQ, tmK, tmV;
    int Lq, Lk;
    int kq;          // D / 64: k-blocks of the score contraction
    int vs;          // value sub-tiles (64 channels each) per CTA
    int rk, rv;      // slots of the K ring / of the V ring
    __half* out;
    long long ldo;
    float scale_log2;
};

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

__global__ void __launch_bounds__(kThreads, 1) tc_attn_wide_kernel(const __grid_constant__ AttnWideParams p) {
    tc::pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int KQ = p.kq, VS = p.vs, RK = p.rk, RV = p.rv;
    uint8_t* sQ = smem;                                   // KQ tiles
    uint8_t* sKr = smem + (size_t)KQ * kTileBytes;        // RK slots: K k-blocks
    uint8_t* sVr = sKr + (size_t)RK * kTileBytes;         // RV slots: V sub-tiles
    uint64_t* bars = reinterpret_cast<uint64_t*>(sVr + (size_t)RV * kTileBytes);
    uint64_t* bar_q = bars + 0;
    uint64_t* k_full = bars + 1;             // [RK]
    uint64_t* k_free = k_full + RK;          // [RK]
    uint64_t* v_full = k_free + RK;          // [RV]
    uint64_t* v_free = v_full + RV;          // [RV]
    uint64_t* s_full = v_free + RV;          // [2] S(g) written (buffer g & 1)
    uint64_t* p_ready = s_full + 2;          // [2] 128 arrivals: P(g) is in TMEM, O rescaled if it had to be
    uint64_t* o_done = p_ready + 2;          // [2] PV(g) retired (buffer g & 1: one phase per TWO key blocks, so a waiter
                                             //     can never fall two phases behind — parity waits would alias)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_done + 2);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int q0 = blockIdx.x * kQTile;
    const int half_idx = blockIdx.y;          // which D_cta-wide slice of the value channels
    const int nb = blockIdx.z;                // frame
    const int G = (p.Lk + kKVTile - 1) / kKVTile;

    if (tid == 0) {
        tc::mbar_init(bar_q, 1);
        for (int i = 0; i < RK; ++i) {
            tc::mbar_init(&k_full[i], 1);
            tc::mbar_init(&k_free[i], 1);
        }
        for (int i = 0; i < RV; ++i) {
            tc::mbar_init(&v_full[i], 1);
            tc::mbar_init(&v_free[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&s_full[i], 1);
            tc::mbar_init(&p_ready[i], 128);
        }
        tc::mbar_init(&o_done[0], 1);
        tc::mbar_init(&o_done[1], 1);
        tc::fence_mbar_init();
    }
    if (warp == 4) {
        tc::tmem_alloc(tmem_ptr_smem, kTmemCols);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    tc::pdl_wait();

    auto block_nk = [&](int g) {
        const int left = p.Lk - g * kKVTile;
        return left < kKVTile ? ((left + 15) & ~15) : kKVTile;
    };

    if (warp == 5) {
        // ------------------------------------------------------------------------------ TMA producer: Q, then the K ring
        if (tc::elect_one()) {
            tc::tma_prefetch_desc(&p.tmQ);
            tc::tma_prefetch_desc(&p.tmK);
            tc::mbar_arrive_expect_tx(bar_q, (uint32_t)(KQ * kTileBytes));
            for (int kb = 0; kb < KQ; ++kb) tc::tma_load_3d(sQ + kb * kTileBytes, &p.tmQ, bar_q, kb * 64, q0, nb);
            int slot = 0;
            uint32_t ph = 0;
            for (int g = 0; g < G; ++g)
                for (int kb = 0; kb < KQ; ++kb) {
                    tc::mbar_wait(&k_free[slot], ph ^ 1u);
                    tc::mbar_arrive_expect_tx(&k_full[slot], kTileBytes);
                    tc::tma_load_3d(sKr + slot * kTileBytes, &p.tmK, &k_full[slot], kb * 64, g * kKVTile, nb);
                    if (++slot == RK) { slot = 0; ph ^= 1u; }
                }
        }
        __syncwarp();
    } else if (warp == 7) {
        // ------------------------------------------------------------------------------ TMA producer: the V ring
        if (tc::elect_one()) {
            tc::tma_prefetch_desc(&p.tmV);
            int slot = 0;
            uint32_t ph = 0;
            for (int g = 0; g < G; ++g)
                for (int j = 0; j < VS; ++j) {
                    tc::mbar_wait(&v_free[slot], ph ^ 1u);
                    tc::mbar_arrive_expect_tx(&v_full[slot], kTileBytes);
                    tc::tma_load_3d(sVr + slot * kTileBytes, &p.tmV, &v_full[slot], (half_idx * VS + j) * 64, g * kKVTile, nb);
                    if (++slot == RV) { slot = 0; ph ^= 1u; }
                }
        }
        __syncwarp();
    } else if (warp == 4) {
        // ------------------------------------------------------------------------------ issuer of the S MMAs (one elected lane)
        if (tc::elect_one()) {
            const uint32_t sQ_a = tc::smem_u32(sQ), sK_a = tc::smem_u32(sKr);
            const uint64_t qd0 = tc::umma_desc_sw128(sQ_a), kd0 = tc::umma_desc_sw128(sK_a);
            int slot = 0;
            uint32_t ph = 0;
            tc::mbar_wait(bar_q, 0);
            for (int g = 0; g < G; ++g) {
                if (g >= 2) {
                    tc::mbar_wait(&o_done[g & 1], (uint32_t)(((g - 2) >> 1) & 1));   // PV(g-2) retired: its P buffer may be overwritten
                    tc::tc_fence_after();
                }
                // S(g) = Q K(g)^T into buffer g & 1: KQ k-blocks x 4 MMAs (K = 16 each)
                const uint32_t idesc = tc::umma_idesc_f16(128, (uint32_t)block_nk(g), 0, 0);
                const uint32_t d_tmem = tmem_base + (uint32_t)(g & 1) * 128u;
                for (int kb = 0; kb < KQ; ++kb) {
                    tc::mbar_wait(&k_full[slot], ph);
                    tc::tc_fence_after();
                    const uint64_t qd = qd0 + (uint64_t)(kb * (kTileBytes >> 4));
                    const uint64_t kd = kd0 + (uint64_t)(slot * (kTileBytes >> 4));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc::umma_f16(d_tmem, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                    tc::umma_commit(&k_free[slot]);
                    if (++slot == RK) { slot = 0; ph ^= 1u; }
                }
                tc::umma_commit(&s_full[g & 1]);
            }
        }
        __syncwarp();
    } else if (warp == 6) {
        // ------------------------------------------------------------------------------ issuer of the PV MMAs (one elected lane)
        // O[:, 64 j : 64 j + 64] += P(g) V(g)[:, sub-tile j]   (P from tensor memory, V sub-tile MN-major)
        if (tc::elect_one()) {
            const uint64_t vd0 = tc::umma_desc_sw128(tc::smem_u32(sVr));
            const uint32_t idesc_o = tc::umma_idesc_f16(128, 64, 0, 1);
            int slot = 0;
            uint32_t ph = 0;
            for (int g = 0; g < G; ++g) {
                const int nk = block_nk(g);
                const uint32_t p_tmem = tmem_base + (uint32_t)(g & 1) * 128u;
                tc::mbar_wait(&p_ready[g & 1], (uint32_t)((g >> 1) & 1));
                tc::tc_fence_after();
                for (int j = 0; j < VS; ++j) {
                    tc::mbar_wait(&v_full[slot], ph);
                    tc::tc_fence_after();
                    const uint64_t vd = vd0 + (uint64_t)(slot * (kTileBytes >> 4));
                    const uint32_t d_tmem = tmem_base + kTmemO + (uint32_t)(j * 64);
                    if (nk == kKVTile) {
#pragma unroll
                        for (int t = 0; t < kKVTile / 16; ++t)
                            tc::umma_f16_ts(d_tmem, p_tmem + (uint32_t)(t * 8), vd + (uint64_t)(t * 128), idesc_o, (g != 0 || t != 0) ? 1u : 0u);
                    } else {
                        for (int t = 0; t < nk / 16; ++t)
                            tc::umma_f16_ts(d_tmem, p_tmem + (uint32_t)(t * 8), vd + (uint64_t)(t * 128), idesc_o, (g != 0 || t != 0) ? 1u : 0u);
                    }
                    tc::umma_commit(&v_free[slot]);
                    if (++slot == RV) { slot = 0; ph ^= 1u; }
                }
                tc::umma_commit(&o_done[g & 1]);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------------------ softmax (one row per thread)
        const int row = tid;                                  // 0..127
        const uint32_t lane_off = ((uint32_t)(warp * 32)) << 16;
        const uint32_t tmem_o = tmem_base + kTmemO + lane_off;
        const int n_o16 = VS * 4;                             // 16-column chunks of this CTA's O
        const float c = p.scale_log2;
        const u64 c2 = pack2(c, c);
        float m_run = -INFINITY, l_run = 0.f;
        for (int g = 0; g < G; ++g) {
            const uint32_t tmem_s = tmem_base + (uint32_t)(g & 1) * 128u + lane_off;
            const int kv_left = p.Lk - g * kKVTile;
            const int nvalid = kv_left < kKVTile ? kv_left : kKVTile;
            tc::mbar_wait(&s_full[g & 1], (uint32_t)((g >> 1) & 1));
            tc::tc_fence_after();
            uint32_t s[128];
            tc::tmem_ld32(tmem_s, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
            tc::tmem_ld32(tmem_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
            tc::tmem_ld32(tmem_s + 64, *reinterpret_cast<uint32_t(*)[32]>(&s[64]));
            tc::tmem_ld32(tmem_s + 96, *reinterpret_cast<uint32_t(*)[32]>(&s[96]));
            tc::tmem_ld_wait();
            if (nvalid < kKVTile) {
#pragma unroll
                for (int i = 0; i < 128; ++i)
                    if (i >= nvalid) s[i] = 0xff800000u;   // -inf
            }
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int i = 0; i < 128; i += 8) {
                mx0 = max3(mx0, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
                mx1 = max3(mx1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
                mx2 = max3(mx2, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
                mx3 = max3(mx3, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
            }
            const float m_new = fmaxf(m_run, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)));
            const bool grow = (m_new - m_run) * c > kRescaleThreshold;    // first block: +inf -> true
            const float m_use = grow ? m_new : m_run;
            // o_done[] is awaited by PARITY: every block's phase is consumed exactly once and in order (a wait that lags two
            // phases behind would return at once) — before the exponentials when O has to be rescaled, after them
            // otherwise (the PV MMA of the previous block then retires underneath).
            const bool rescale = g > 0 && __any_sync(0xffffffffu, grow);
            if (rescale) {
                tc::mbar_wait(&o_done[(g - 1) & 1], (uint32_t)(((g - 1) >> 1) & 1));
                tc::tc_fence_after();
                const float alpha = ex2((m_run - m_use) * c);
                l_run *= alpha;
                for (int cc = 0; cc < n_o16; ++cc) {
                    uint32_t r[16];
                    tc::tmem_ld16(tmem_o + (uint32_t)(cc * 16), r);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                    tc::tmem_st16(tmem_o + (uint32_t)(cc * 16), r);
                }
                tc::tmem_st_wait();
            }
            m_run = m_use;
            const float nm = -m_use * c;
            const u64 nm2 = pack2(nm, nm);
            u64 sum_a = 0ull, sum_b = 0ull;
            uint32_t pk[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                const u64 x = fma2(pack2(__uint_as_float(s[2 * j]), __uint_as_float(s[2 * j + 1])), c2, nm2);
                float x0, x1;
                unpack2(x, x0, x1);
                const float e0 = ex2(x0), e1 = ex2(x1);
                if (j & 1) sum_b = add2(sum_b, pack2(e0, e1));
                else sum_a = add2(sum_a, pack2(e0, e1));
                pk[j] = pack_h2(e0, e1);
            }
            if (g > 0 && !rescale) tc::mbar_wait(&o_done[(g - 1) & 1], (uint32_t)(((g - 1) >> 1) & 1));
            // P overwrites the first 64 columns of this block's S buffer (every thread has its whole S row in registers)
            tc::tmem_st32(tmem_s, &pk[0]);
            tc::tmem_st32(tmem_s + 32, &pk[32]);
            tc::tmem_st_wait();
            tc::tc_fence_before();
            tc::mbar_arrive(&p_ready[g & 1]);
            float a0, a1, b0, b1;
            unpack2(sum_a, a0, a1);
            unpack2(sum_b, b0, b1);
            l_run += (a0 + a1) + (b0 + b1);
        }
        // ---- epilogue: O / l -> fp16 -> global (this CTA's channel slice)
        tc::mbar_wait(&o_done[(G - 1) & 1], (uint32_t)(((G - 1) >> 1) & 1));
        tc::tc_fence_after();
        const float inv_l = 1.0f / l_run;
        const int qrow = q0 + row;
        __half* dst = p.out + ((long long)nb * p.Lq + qrow) * p.ldo + (long long)half_idx * VS * 64;
        for (int cc = 0; cc < VS * 2; ++cc) {
            uint32_t r[32];
            tc::tmem_ld32(tmem_o + (uint32_t)(cc * 32), r);
            tc::tmem_ld_wait();
            if (qrow < p.Lq) {
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    uint4 u;
                    u.x = pack_h2(__uint_as_float(r[8 * q4 + 0]) * inv_l, __uint_as_float(r[8 * q4 + 1]) * inv_l);
                    u.y = pack_h2(__uint_as_float(r[8 * q4 + 2]) * inv_l, __uint_as_float(r[8 * q4 + 3]) * inv_l);
                    u.z = pack_h2(__uint_as_float(r[8 * q4 + 4]) * inv_l, __uint_as_float(r[8 * q4 + 5]) * inv_l);
                    u.w = pack_h2(__uint_as_float(r[8 * q4 + 6]) * inv_l, __uint_as_float(r[8 * q4 + 7]) * inv_l);
                    reinterpret_cast<uint4*>(dst)[cc * 4 + q4] = u;
                }
            }
        }
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace

using namespace tc_host;

extern "C" int tc_attention_wide(const void* q, const void* k, const void* v, long long ldq, long long ldk, long long ldv,
                                 void* out, long long ldo, int batches, int Lq, int Lk, int D, float scale, void* stream_v) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    TC_CHECK_ARG(q && k && v && out, "tc_attention_wide: null pointer");
    TC_CHECK_ARG(batches > 0 && Lq > 0 && Lk > 0, "tc_attention_wide: empty problem");
    TC_CHECK_ARG(D >= 64 && D <= 512 && D % 64 == 0 && (D <= 256 || D % 128 == 0),
                 "tc_attention_wide: head dim must be a multiple of 64 (of 128 above 256), at most 512");
    TC_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "tc_attention_wide: strides must be multiples of 8");
    TC_CHECK_ARG(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                   reinterpret_cast<uintptr_t>(out)) & 15) == 0, "tc_attention_wide: operands must be 16-byte aligned");
    AttnWideParams p;
    memset(&p, 0, sizeof(p));
    const uint32_t box[3] = {64, 128, 1};
    {
        uint64_t dims[3] = {(uint64_t)D, (uint64_t)Lq, (uint64_t)batches};
        uint64_t str[2] = {(uint64_t)ldq * 2, (uint64_t)Lq * (uint64_t)ldq * 2};
        const CUtensorMap* m = get_tensor_map(q, 3, dims, str, box);
        if (!m) return TC_ERR_CUDA;
        p.tmQ = *m;
    }
    {
        uint64_t dims[3] = {(uint64_t)D, (uint64_t)Lk, (uint64_t)batches};
        uint64_t strk[2] = {(uint64_t)ldk * 2, (uint64_t)Lk * (uint64_t)ldk * 2};
        uint64_t strv[2] = {(uint64_t)ldv * 2, (uint64_t)Lk * (uint64_t)ldv * 2};
        const CUtensorMap* mk = get_tensor_map(k, 3, dims, strk, box);
        const CUtensorMap* mv = get_tensor_map(v, 3, dims, strv, box);
        if (!mk || !mv) return TC_ERR_CUDA;
        p.tmK = *mk;
        p.tmV = *mv;
    }
    const int halves = D > 256 ? 2 : 1;
    p.Lq = Lq;
    p.Lk = Lk;
    p.kq = D / 64;
    p.vs = D / 64 / halves;
    p.out = reinterpret_cast<__half*>(out);
    p.ldo = ldo;
    p.scale_log2 = scale * 1.4426950408889634f;
    const int budget = 227 * 1024 - 1024 - 1024;          // alignment slack, barriers
    int slots = (budget - p.kq * kTileBytes) / kTileBytes;
    if (slots > 12) slots = 12;
    TC_CHECK_ARG(slots >= 4, "tc_attention_wide: not enough shared memory for the K and V rings");
    p.rv = slots / 3 < 2 ? 2 : slots / 3;                 // a third of the slots (>= 2) stream V, the rest K
    if (p.rv > p.vs) p.rv = p.vs < 2 ? 2 : p.vs;
    p.rk = slots - p.rv;
    const size_t smem_bytes = (size_t)(p.kq + slots) * kTileBytes + 2048;
    static bool attr_set = false;
    if (!attr_set) {
        int rc = check_cuda(cudaFuncSetAttribute(tc_attn_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024),
                            "cudaFuncSetAttribute(tc_attn_wide_kernel)");
        if (rc) return rc;
        attr_set = true;
    }
    dim3 grid((Lq + kQTile - 1) / kQTile, halves, batches);
    tc_host::launch(tc_attn_wide_kernel, grid, dim3(kThreads), smem_bytes, stream, 1, p);
    count_launch();
    TC_CHECK_LAUNCH("tc_attn_wide_kernel");
    return TC_OK;
}

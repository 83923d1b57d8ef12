// tooncrafter_b200 — attention kernels.
//
//  tc_attention           fused softmax(QK^T)V, head dim 64, tcgen05 MMAs (S = QK^T and O = PV) with TMEM
//                         accumulators, TMA-staged Q/K/V tiles, fp32 online softmax, one query row per thread;
//                         two query tiles per CTA ping-pong so one tile's softmax overlaps the other's MMAs.
//  tc_temporal_attention  16-frame temporal self-attention, one (pixel, head) per warp, mma.sync 16x8x16 (HBM-bound).
//  tc_softmax_rows        row softmax for the unfused d=512 VAE mid-block attention.
//
// Reference sites: lvdm/modules/attention.py:81-209,365-412; lvdm/models/autoencoder_dualref.py:172-200,270-341.
#include <stdlib.h>
#include "tc_common.cuh"
#include "tc_host.h"

namespace {

// ===================================================================================== fused attention (d = 64)
constexpr int kQTile = 128;
constexpr int kKVTile = 128;
constexpr int kTileBytes = 128 * 64 * 2;  // 16 KiB: [128 rows][64 halfs], 128B-swizzled

struct alignas(64) AttnKParams {
    CUtensorMap tmQ;
    CUtensorMap tmK[2];
    CUtensorMap tmV[2];
    int Lq, heads, n_seg;
    int Lk[2];
    int kv_div[2];
    __half* out;
    long long ldo;
    float scale_log2;
};

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_h2(float x, float y) {
    __half2 h = __floats2half2_rn(x, y);
    return *reinterpret_cast<uint32_t*>(&h);
}
// (x, y) = hi + lo with hi = fp16(x, y) and lo = fp16 of the rounding residual
__device__ __forceinline__ void split_h2(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x, y);
    const float2 f = __half22float2(h);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = pack_h2(x - f.x, y - f.y);
}

// Ping-pong over TWO query tiles per CTA with specialised warps:
//   warps 0-3 / 4-7 : softmax warpgroups for query tile 0 / 1 (one query row per thread)
//   warp 8          : MMA issuer   (S_w = Q_w K^T, O_w = P_w V; tcgen05, accumulators in TMEM)
//   warp 9          : TMA producer (Q tiles once, K / V tiles double-buffered)
// While one warpgroup runs its softmax (MUFU/ALU bound) the tensor core works on the other tile's MMAs.
constexpr int kAttnThreads = 320;
constexpr int kAttnTmemCols = 512;  // S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384)

template <bool kTwoSeg>
__global__ void __launch_bounds__(kAttnThreads, 1) tc_attn_kernel(const __grid_constant__ AttnKParams p) {
    tc::pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                     // 2 tiles
    uint8_t* sK = smem + 2 * kTileBytes;    // 2 stages
    uint8_t* sV = smem + 4 * kTileBytes;    // 2 stages
    uint8_t* sP = smem + 6 * kTileBytes;    // 2 tiles x 2 sub-tiles
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 10 * kTileBytes);
    uint64_t* bar_q = bars + 0;
    uint64_t* k_full = bars + 1;    // [2]
    uint64_t* k_free = bars + 3;    // [2]
    uint64_t* v_full = bars + 5;    // [2]
    uint64_t* v_free = bars + 7;    // [2]
    uint64_t* s_full = bars + 9;    // [2] per query tile
    uint64_t* o_full = bars + 11;   // [2]
    uint64_t* p_ready = bars + 13;  // [2], 128 arrivals
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 15);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int q0 = blockIdx.x * 2 * kQTile;
    const int head = blockIdx.y;
    const int qb = blockIdx.z;
    const int ntiles = (q0 + kQTile < p.Lq) ? 2 : 1;

    if (tid == 0) {
        tc::mbar_init(bar_q, 1);
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&k_full[i], 1);
            tc::mbar_init(&k_free[i], 1);
            tc::mbar_init(&v_full[i], 1);
            tc::mbar_init(&v_free[i], 1);
            tc::mbar_init(&s_full[i], 1);
            tc::mbar_init(&o_full[i], 1);
            tc::mbar_init(&p_ready[i], 128);
        }
        tc::fence_mbar_init();
    }
    if (warp == 8) {
        tc::tmem_alloc(tmem_ptr_smem, kAttnTmemCols);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    tc::pdl_wait();   // prologue (barriers, TMEM) overlapped the predecessor; Q/K/V are its results

    if (warp == 9) {
        // ------------------------------------------------------------------------------ TMA producer
        // whole warp converged, one elected lane issues (see tc_gemm.cu for why `if (lane == 0)` is slow)
        if (tc::elect_one()) {
            tc::tma_prefetch_desc(&p.tmQ);
            tc::mbar_arrive_expect_tx(bar_q, (uint32_t)(ntiles * kTileBytes));
            for (int w = 0; w < ntiles; ++w) tc::tma_load_3d(sQ + w * kTileBytes, &p.tmQ, bar_q, head * 64, q0 + w * kQTile, qb);
        }
        __syncwarp();
        int g = 0;
        for (int seg = 0; seg < p.n_seg; ++seg) {
            const int nblk = (p.Lk[seg] + kKVTile - 1) / kKVTile;
            const int kvb = qb / p.kv_div[seg];
            for (int j = 0; j < nblk; ++j, ++g) {
                const int st = g & 1;
                const uint32_t ph = (uint32_t)((g >> 1) & 1);
                tc::mbar_wait(&k_free[st], ph ^ 1u);
                if (tc::elect_one()) {
                    tc::mbar_arrive_expect_tx(&k_full[st], kTileBytes);
                    tc::tma_load_3d(sK + st * kTileBytes, &p.tmK[seg], &k_full[st], head * 64, j * kKVTile, kvb);
                }
                __syncwarp();
                tc::mbar_wait(&v_free[st], ph ^ 1u);
                if (tc::elect_one()) {
                    tc::mbar_arrive_expect_tx(&v_full[st], kTileBytes);
                    tc::tma_load_3d(sV + st * kTileBytes, &p.tmV[seg], &v_full[st], head * 64, j * kKVTile, kvb);
                }
                __syncwarp();
            }
        }
    } else if (warp == 8) {
        // ------------------------------------------------------------------------------ MMA issuer
        {
            const uint32_t sQ_a = tc::smem_u32(sQ), sK_a = tc::smem_u32(sK), sV_a = tc::smem_u32(sV), sP_a = tc::smem_u32(sP);
            int G = 0;
            for (int seg = 0; seg < p.n_seg; ++seg) G += (p.Lk[seg] + kKVTile - 1) / kKVTile;
            // per-block key counts, walked in lock-step with the other roles
            auto block_nk = [&](int gi) {
                int seg = 0, j = gi;
                while (true) {
                    const int nb = (p.Lk[seg] + kKVTile - 1) / kKVTile;
                    if (j < nb) break;
                    j -= nb;
                    ++seg;
                }
                const int left = p.Lk[seg] - j * kKVTile;
                return left < kKVTile ? ((left + 15) & ~15) : kKVTile;
            };
            // S_w = Q_w K^T for key block in stage `st` (+ the commits that follow it); called by ONE elected lane
            auto issue_s = [&](int w, int st, int nk, bool last_tile) {
                const uint32_t idesc = tc::umma_idesc_f16(128, (uint32_t)nk, 0, 0);
                const uint64_t qd = tc::umma_desc_sw128(sQ_a + (uint32_t)w * kTileBytes);
                const uint64_t kd = tc::umma_desc_sw128(sK_a + (uint32_t)st * kTileBytes);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc::umma_f16(tmem_base + (uint32_t)w * 128, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc, k != 0);
                tc::umma_commit(&s_full[w]);
                if (last_tile) tc::umma_commit(&k_free[st]);
            };
            tc::mbar_wait(bar_q, 0);
            tc::mbar_wait(&k_full[0], 0);
            tc::tc_fence_after();
            {
                const int nk = block_nk(0);
                if (tc::elect_one()) {
                    for (int w = 0; w < ntiles; ++w) issue_s(w, 0, nk, w == ntiles - 1);
                }
                __syncwarp();
            }
            for (int g = 0; g < G; ++g) {
                const int st = g & 1;
                const int nk = block_nk(g);
                const int nk_next = (g + 1 < G) ? block_nk(g + 1) : 0;
                for (int w = 0; w < ntiles; ++w) {
                    tc::mbar_wait(&p_ready[w], (uint32_t)(g & 1));
                    if (w == 0) tc::mbar_wait(&v_full[st], (uint32_t)((g >> 1) & 1));
                    if (g + 1 < G && w == 0) tc::mbar_wait(&k_full[(g + 1) & 1], (uint32_t)(((g + 1) >> 1) & 1));
                    tc::tc_fence_after();
                    if (tc::elect_one()) {
                        const uint32_t idesc_o = tc::umma_idesc_f16(128, 64, 0, 1);  // B (= V tile) MN-major
                        const uint64_t vd = tc::umma_desc_sw128(sV_a + (uint32_t)st * kTileBytes);
                        for (int t = 0; t < nk / 16; ++t) {
                            const uint64_t pd = tc::umma_desc_sw128(sP_a + (uint32_t)(2 * w + (t >> 2)) * kTileBytes) + (uint64_t)((t & 3) * 2);
                            tc::umma_f16(tmem_base + 256 + (uint32_t)w * 64, pd, vd + (uint64_t)(t * 128), idesc_o, t != 0);
                        }
                        tc::umma_commit(&o_full[w]);
                        if (w == ntiles - 1) tc::umma_commit(&v_free[st]);
                        if (g + 1 < G) issue_s(w, (g + 1) & 1, nk_next, w == ntiles - 1);
                    }
                    __syncwarp();
                }
            }
        }
    } else if ((warp >> 2) < ntiles) {
        // ------------------------------------------------------------------------------ softmax warpgroups
        const int w = warp >> 2;
        const int row = tid & 127;
        const uint32_t lane_off = ((uint32_t)((warp & 3) * 32)) << 16;
        const uint32_t tmem_s = tmem_base + (uint32_t)w * 128 + lane_off;
        const uint32_t tmem_o = tmem_base + 256 + (uint32_t)w * 64 + lane_off;
        uint8_t* sPw = sP + (size_t)(2 * w) * kTileBytes;
        float o_total[kTwoSeg ? 64 : 1];   // sum over segments (text + image cross attention) only when needed
        if constexpr (kTwoSeg) {
#pragma unroll
            for (int i = 0; i < 64; ++i) o_total[i] = 0.f;
        }
        float o_acc[64];
        int g = 0;
        for (int seg = 0; seg < p.n_seg; ++seg) {
            const int Lk = p.Lk[seg];
            const int nblk = (Lk + kKVTile - 1) / kKVTile;
            float m_run = -INFINITY, l_run = 0.f;
#pragma unroll
            for (int i = 0; i < 64; ++i) o_acc[i] = 0.f;
            for (int j = 0; j < nblk; ++j, ++g) {
                const int kv_left = Lk - j * kKVTile;
                const int nvalid = kv_left < kKVTile ? kv_left : kKVTile;
                const int nk = (nvalid + 15) & ~15;
                tc::mbar_wait(&s_full[w], (uint32_t)(g & 1));
                tc::tc_fence_after();
                // ---- pass 1: row max.  Full blocks take the unpredicated path with 4 independent max chains
                // (the serial FMNMX / FADD chains were the softmax warps' critical path: 2 warps per scheduler).
                float m_blk;
                const bool full = (nvalid == kKVTile);
                if (full) {
                    float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t r[32];
                        tc::tmem_ld32(tmem_s + (uint32_t)(c * 32), r);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            mx[0] = fmaxf(mx[0], __uint_as_float(r[i]));
                            mx[1] = fmaxf(mx[1], __uint_as_float(r[i + 1]));
                            mx[2] = fmaxf(mx[2], __uint_as_float(r[i + 2]));
                            mx[3] = fmaxf(mx[3], __uint_as_float(r[i + 3]));
                        }
                    }
                    m_blk = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
                } else {
                    m_blk = -INFINITY;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        if (c * 16 < nk) {
                            uint32_t r[16];
                            tc::tmem_ld16(tmem_s + (uint32_t)(c * 16), r);
                            tc::tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (c * 16 + i < nvalid) m_blk = fmaxf(m_blk, __uint_as_float(r[i]));
                        }
                    }
                }
                const float m_new = fmaxf(m_run, m_blk);
                const float m_scaled = m_new * p.scale_log2;
                const float alpha = fast_exp2(m_run * p.scale_log2 - m_scaled);
                // ---- pass 2: p = exp2(s * scale - m * scale), row sum, P tile (fp16, K-major, 128B swizzle) to smem
                float l_blk;
                if (full) {
                    float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t r[32];
                        tc::tmem_ld32(tmem_s + (uint32_t)(c * 32), r);
                        tc::tmem_ld_wait();
                        uint32_t pk[16];
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float e0 = fast_exp2(fmaf(__uint_as_float(r[i]), p.scale_log2, -m_scaled));
                            const float e1 = fast_exp2(fmaf(__uint_as_float(r[i + 1]), p.scale_log2, -m_scaled));
                            const float e2 = fast_exp2(fmaf(__uint_as_float(r[i + 2]), p.scale_log2, -m_scaled));
                            const float e3 = fast_exp2(fmaf(__uint_as_float(r[i + 3]), p.scale_log2, -m_scaled));
                            ls[0] += e0;
                            ls[1] += e1;
                            ls[2] += e2;
                            ls[3] += e3;
                            pk[i / 2] = pack_h2(e0, e1);
                            pk[i / 2 + 1] = pack_h2(e2, e3);
                        }
                        uint8_t* sub = sPw + (c >> 1) * kTileBytes + row * 128;
                        const int ch0 = (c & 1) * 4;
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4)
                            *reinterpret_cast<uint4*>(sub + (((ch0 + q4) ^ (row & 7)) << 4)) =
                                make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
                    }
                    l_blk = (ls[0] + ls[1]) + (ls[2] + ls[3]);
                } else {
                    l_blk = 0.f;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        if (c * 16 < nk) {
                            uint32_t r[16];
                            tc::tmem_ld16(tmem_s + (uint32_t)(c * 16), r);
                            tc::tmem_ld_wait();
                            uint32_t pk[8];
#pragma unroll
                            for (int i = 0; i < 16; i += 2) {
                                float e0 = fast_exp2(fmaf(__uint_as_float(r[i]), p.scale_log2, -m_scaled));
                                float e1 = fast_exp2(fmaf(__uint_as_float(r[i + 1]), p.scale_log2, -m_scaled));
                                e0 = (c * 16 + i < nvalid) ? e0 : 0.f;
                                e1 = (c * 16 + i + 1 < nvalid) ? e1 : 0.f;
                                l_blk += e0 + e1;
                                pk[i / 2] = pack_h2(e0, e1);
                            }
                            uint8_t* sub = sPw + (c >> 2) * kTileBytes + row * 128;
                            const int ch0 = (c & 3) * 2;
                            *reinterpret_cast<uint4*>(sub + (((ch0) ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                            *reinterpret_cast<uint4*>(sub + (((ch0 + 1) ^ (row & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                        }
                    }
                }
                l_run = l_run * alpha + l_blk;
                m_run = m_new;
                tc::fence_proxy_async_smem();
                tc::tc_fence_before();
                tc::mbar_arrive(&p_ready[w]);
                // O_blk of this block: o_acc = o_acc * alpha + P V
                tc::mbar_wait(&o_full[w], (uint32_t)(g & 1));
                tc::tc_fence_after();
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t r[32];
                    tc::tmem_ld32(tmem_o + (uint32_t)(c * 32), r);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = fmaf(o_acc[c * 32 + i], alpha, __uint_as_float(r[i]));
                }
            }
            const float inv_l = 1.0f / l_run;
            if constexpr (kTwoSeg) {
#pragma unroll
                for (int i = 0; i < 64; ++i) o_total[i] += o_acc[i] * inv_l;
            } else {
#pragma unroll
                for (int i = 0; i < 64; ++i) o_acc[i] *= inv_l;
            }
        }
        const float* o_fin = kTwoSeg ? o_total : o_acc;
        const int qrow = q0 + w * kQTile + row;
        if (qrow < p.Lq) {
            __half* dst = p.out + ((long long)qb * p.Lq + qrow) * p.ldo + head * 64;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                __half2 h[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(o_fin[c * 8 + 2 * i], o_fin[c * 8 + 2 * i + 1]);
                uint4 u;
                u.x = *reinterpret_cast<uint32_t*>(&h[0]);
                u.y = *reinterpret_cast<uint32_t*>(&h[1]);
                u.z = *reinterpret_cast<uint32_t*>(&h[2]);
                u.w = *reinterpret_cast<uint32_t*>(&h[3]);
                reinterpret_cast<uint4*>(dst)[c] = u;
            }
        }
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, kAttnTmemCols);
    }
}

// ===================================================================================== temporal attention
// x[b][t][p][heads*64]; one (b, p, head) item per group of kGroup lanes (lane i of the group = query frame i).
template <int kGroup>
__global__ void temporal_attn_kernel(const __half* __restrict__ q, const __half* __restrict__ k,
                                     const __half* __restrict__ v, long long ld, __half* __restrict__ out,
                                     long long ldo, int B, int T, int P, int heads, float scale_log2) {
    tc::pdl_wait();   // no early launch_dependents: waiting successor CTAs would squat on this kernel's SM slots
    constexpr int kItemsPerWarp = 32 / kGroup;
    constexpr int kWarps = 4;
    __shared__ __align__(16) __half sK[kWarps * kItemsPerWarp][kGroup][64];
    __shared__ __align__(16) __half sV[kWarps * kItemsPerWarp][kGroup][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane / kGroup, t = lane % kGroup;
    const long long total = (long long)B * P * heads;
    const long long item = ((long long)blockIdx.x * kWarps + warp) * kItemsPerWarp + sub;
    const bool item_ok = item < total;
    const int slot = warp * kItemsPerWarp + sub;

    long long b = 0, pix = 0;
    int h = 0;
    if (item_ok) {
        h = (int)(item % heads);
        const long long r = item / heads;
        pix = r % P;
        b = r / P;
    }
    const bool row_ok = item_ok && t < T;
    const long long tok = ((b * T + t) * P + pix);
    float qf[64];
    if (row_ok) {
        const uint4* qp = reinterpret_cast<const uint4*>(q + tok * ld + h * 64);
        const uint4* kp = reinterpret_cast<const uint4*>(k + tok * ld + h * 64);
        const uint4* vp = reinterpret_cast<const uint4*>(v + tok * ld + h * 64);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint4 u = qp[c];
            const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(hh[i]);
                qf[c * 8 + 2 * i] = f.x;
                qf[c * 8 + 2 * i + 1] = f.y;
            }
            reinterpret_cast<uint4*>(&sK[slot][t][0])[c] = kp[c];
            reinterpret_cast<uint4*>(&sV[slot][t][0])[c] = vp[c];
        }
    }
    __syncwarp();
    if (!row_ok) return;

    float s[kGroup];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < kGroup; ++j) {
        if (j < T) {
            float acc = 0.f;
            const __half2* kr = reinterpret_cast<const __half2*>(&sK[slot][j][0]);
#pragma unroll
            for (int d = 0; d < 32; ++d) {
                const float2 f = __half22float2(kr[d]);
                acc += qf[2 * d] * f.x + qf[2 * d + 1] * f.y;
            }
            s[j] = acc * scale_log2;
            m = fmaxf(m, s[j]);
        }
    }
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < kGroup; ++j) {
        if (j < T) {
            s[j] = exp2f(s[j] - m);
            l += s[j];
        }
    }
    const float inv = 1.0f / l;
    float o[64];
#pragma unroll
    for (int d = 0; d < 64; ++d) o[d] = 0.f;
#pragma unroll
    for (int j = 0; j < kGroup; ++j) {
        if (j < T) {
            const float pj = s[j] * inv;
            const __half2* vr = reinterpret_cast<const __half2*>(&sV[slot][j][0]);
#pragma unroll
            for (int d = 0; d < 32; ++d) {
                const float2 f = __half22float2(vr[d]);
                o[2 * d] += pj * f.x;
                o[2 * d + 1] += pj * f.y;
            }
        }
    }
    __half* dst = out + tok * ldo + h * 64;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        __half2 hh[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) hh[i] = __floats2half2_rn(o[c * 8 + 2 * i], o[c * 8 + 2 * i + 1]);
        uint4 u;
        u.x = *reinterpret_cast<uint32_t*>(&hh[0]);
        u.y = *reinterpret_cast<uint32_t*>(&hh[1]);
        u.z = *reinterpret_cast<uint32_t*>(&hh[2]);
        u.w = *reinterpret_cast<uint32_t*>(&hh[3]);
        reinterpret_cast<uint4*>(dst)[c] = u;
    }
}

// ===================================================================================== temporal attention, T <= 16
// One warp per (batch, pixel, head): the whole problem is a 16 x 16 x 64 attention, far below the 128-row tcgen05 tile,
// and the kernel is HBM-bound (reads q, k, v once, writes o once).  S = Q K^T and O = P V use warp-level
// mma.sync.m16n8k16 with fragments loaded straight from global memory (Q, K) or via ldmatrix.trans from a
// swizzled 2 KiB smem tile (V); softmax in fp32 on the accumulator fragments.
__device__ __forceinline__ void mma_m16n8k16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__global__ void __launch_bounds__(128) temporal_attn_mma_kernel(const __half* __restrict__ q, const __half* __restrict__ k,
                                                                const __half* __restrict__ v, long long ld,
                                                                __half* __restrict__ out, long long ldo, int B, int T,
                                                                int P, int heads, float scale_log2) {
    tc::pdl_wait();   // no early launch_dependents: waiting successor CTAs would squat on this kernel's SM slots
    __shared__ __align__(128) uint8_t sV[4][16 * 128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const long long total = (long long)B * P * heads;
    const long long frame_stride = (long long)P * ld;       // elements between consecutive frames of one pixel
    uint8_t* sv = sV[warp];
    for (long long item = (long long)blockIdx.x * 4 + warp; item < total; item += (long long)gridDim.x * 4) {
        const int h = (int)(item % heads);
        const long long r = item / heads;
        const long long pix = r % P, b = r / P;
        const long long base = ((b * T) * P + pix) * ld + h * 64;   // frame 0 of this (b, pix, head)
        // ---- V tile -> smem (16-byte chunks, XOR-swizzled so ldmatrix rows hit distinct banks)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int chunk = lane + 32 * i;
            const int row = chunk >> 3, c16 = chunk & 7;
            uint4 u = make_uint4(0u, 0u, 0u, 0u);
            if (row < T) u = *reinterpret_cast<const uint4*>(v + base + row * frame_stride + c16 * 8);
            *reinterpret_cast<uint4*>(sv + row * 128 + ((c16 ^ (row & 7)) << 4)) = u;
        }
        // ---- Q (A fragments) and K (B fragments) straight from global
        uint32_t qa[4][4], kb[2][4][2];
        const bool r0 = g < T, r1 = g + 8 < T;
        const __half* q0p = q + base + (long long)g * frame_stride;
        const __half* q1p = q + base + (long long)(g + 8) * frame_stride;
        const __half* k0p = k + base + (long long)g * frame_stride;
        const __half* k1p = k + base + (long long)(g + 8) * frame_stride;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const int c = 16 * ks + 2 * t4;
            qa[ks][0] = r0 ? *reinterpret_cast<const uint32_t*>(q0p + c) : 0u;
            qa[ks][1] = r1 ? *reinterpret_cast<const uint32_t*>(q1p + c) : 0u;
            qa[ks][2] = r0 ? *reinterpret_cast<const uint32_t*>(q0p + c + 8) : 0u;
            qa[ks][3] = r1 ? *reinterpret_cast<const uint32_t*>(q1p + c + 8) : 0u;
            kb[0][ks][0] = r0 ? *reinterpret_cast<const uint32_t*>(k0p + c) : 0u;
            kb[0][ks][1] = r0 ? *reinterpret_cast<const uint32_t*>(k0p + c + 8) : 0u;
            kb[1][ks][0] = r1 ? *reinterpret_cast<const uint32_t*>(k1p + c) : 0u;
            kb[1][ks][1] = r1 ? *reinterpret_cast<const uint32_t*>(k1p + c + 8) : 0u;
        }
        float sacc[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
            for (int i = 0; i < 4; ++i) sacc[nt][i] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) mma_m16n8k16(sacc[nt], qa[ks], kb[nt][ks][0], kb[nt][ks][1]);
        }
        // ---- softmax over the 16 keys of rows g and g + 8 (a row lives in the 4 lanes of a quad)
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const bool key_ok = (8 * nt + 2 * t4 + i) < T;
                sacc[nt][i] = key_ok ? sacc[nt][i] * scale_log2 : -INFINITY;
                sacc[nt][2 + i] = key_ok ? sacc[nt][2 + i] * scale_log2 : -INFINITY;
                m0 = fmaxf(m0, sacc[nt][i]);
                m1 = fmaxf(m1, sacc[nt][2 + i]);
            }
        }
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                sacc[nt][i] = exp2f(sacc[nt][i] - m0);
                sacc[nt][2 + i] = exp2f(sacc[nt][2 + i] - m1);
                l0 += sacc[nt][i];
                l1 += sacc[nt][2 + i];
            }
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
        // P = hi + lo with both parts fp16: with only T <= 16 keys per query the fp16 rounding of P is not averaged
        // away, and the kernel is HBM-bound, so the second (residual) MMA is free
        uint32_t pa[4], pl[4];
        split_h2(sacc[0][0] * inv0, sacc[0][1] * inv0, pa[0], pl[0]);
        split_h2(sacc[0][2] * inv1, sacc[0][3] * inv1, pa[1], pl[1]);
        split_h2(sacc[1][0] * inv0, sacc[1][1] * inv0, pa[2], pl[2]);
        split_h2(sacc[1][2] * inv1, sacc[1][3] * inv1, pa[3], pl[3]);
        __syncwarp();
        // ---- O = P V, V fragments by ldmatrix.trans (two 8-wide d tiles per instruction)
        __half* o0p = out + ((b * T + g) * P + pix) * ldo + h * 64;
        __half* o1p = out + ((b * T + g + 8) * P + pix) * ldo + h * 64;
#pragma unroll
        for (int nd = 0; nd < 8; nd += 2) {
            // lane i supplies the row address of matrix i/8: matrices (k 0-7, nd), (k 8-15, nd), (k 0-7, nd+1), (k 8-15, nd+1)
            const int mrow = (lane & 7) + ((lane >> 3) & 1) * 8;
            const int mcol = nd + (lane >> 4);
            const uint32_t addr = tc::smem_u32(sv + mrow * 128 + ((mcol ^ (mrow & 7)) << 4));
            uint32_t b0, b1, b2, b3;
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                         : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                         : "r"(addr));
            float oa[4] = {0.f, 0.f, 0.f, 0.f}, ob[4] = {0.f, 0.f, 0.f, 0.f};
            mma_m16n8k16(oa, pl, b0, b1);
            mma_m16n8k16(ob, pl, b2, b3);
            mma_m16n8k16(oa, pa, b0, b1);
            mma_m16n8k16(ob, pa, b2, b3);
            if (r0) {
                *reinterpret_cast<uint32_t*>(o0p + 8 * nd + 2 * t4) = pack_h2(oa[0], oa[1]);
                *reinterpret_cast<uint32_t*>(o0p + 8 * (nd + 1) + 2 * t4) = pack_h2(ob[0], ob[1]);
            }
            if (r1) {
                *reinterpret_cast<uint32_t*>(o1p + 8 * nd + 2 * t4) = pack_h2(oa[2], oa[3]);
                *reinterpret_cast<uint32_t*>(o1p + 8 * (nd + 1) + 2 * t4) = pack_h2(ob[2], ob[3]);
            }
        }
        __syncwarp();   // smem V tile is reused by the next item
    }
}

// ===================================================================================== row softmax (in place)
__global__ void softmax_rows_kernel(__half* __restrict__ s, long long lds, int rows, int cols, float scale_log2) {
    tc::pdl_wait();   // no early launch_dependents: waiting successor CTAs would squat on this kernel's SM slots
    const int row = blockIdx.x;
    if (row >= rows) return;
    __half* r = s + (long long)row * lds;
    __shared__ float red[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float m = -INFINITY;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) m = fmaxf(m, __half2float(r[c]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
    for (int w = 1; w < nw; ++w) m = fmaxf(m, red[w]);
    __syncthreads();
    float l = 0.f;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) l += exp2f((__half2float(r[c]) - m) * scale_log2);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    if (lane == 0) red[warp] = l;
    __syncthreads();
    l = 0.f;
    for (int w = 0; w < nw; ++w) l += red[w];
    const float inv = 1.0f / l;
    for (int c = threadIdx.x; c < cols; c += blockDim.x)
        r[c] = __float2half_rn(exp2f((__half2float(r[c]) - m) * scale_log2) * inv);
}

}  // namespace

using namespace tc_host;

int tc_attention_v3(const TcAttention* d, int poly_of_8, cudaStream_t stream);   // tc_attn3.cu
int tc_attention_xs(const TcAttention* d, cudaStream_t stream);                    // tc_attn3.cu (small K/V, 1-2 segments)

// TC_ATTN_IMPL=v2 keeps single-segment problems on the second-generation kernel below (A/B runs);
// TC_ATTN_POLY=n (0..4) sets how many of every 8 exponential pairs the v3 kernel evaluates on the FMA pipe.
static int attn_impl_v3() {
    const char* e = getenv("TC_ATTN_IMPL");
    return !(e && e[0] == 'v' && e[1] == '2');
}
static int attn_poly() {
    const char* e = getenv("TC_ATTN_POLY");
    return (e && e[0] >= '0' && e[0] <= '4') ? e[0] - '0' : 0;
}

extern "C" int tc_attention(const TcAttention* d, void* stream_v) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    TC_CHECK_ARG(d && d->q && d->out, "tc_attention: null pointer");
    TC_CHECK_ARG(d->n_seg == 1 || d->n_seg == 2, "tc_attention: n_seg must be 1 or 2");
    TC_CHECK_ARG(d->q_batches > 0 && d->Lq > 0 && d->heads > 0, "tc_attention: empty problem");
    TC_CHECK_ARG(d->ldq % 8 == 0 && d->ldo % 8 == 0, "tc_attention: strides must be multiples of 8");
    if (d->n_seg == 1 && attn_impl_v3()) {
        TC_CHECK_ARG(d->k[0] && d->v[0] && d->Lk[0] > 0 && d->kv_div[0] > 0, "tc_attention: bad kv segment");
        TC_CHECK_ARG(d->ldk[0] % 8 == 0 && d->ldv[0] % 8 == 0, "tc_attention: kv strides must be multiples of 8");
        return tc_attention_v3(d, attn_poly(), stream);
    }
    if (attn_impl_v3()) {
        // short K/V (the 77 text + 16 image tokens of the cross attentions): resident-K/V kernel when the shape fits
        bool ok = true;
        for (int s = 0; s < d->n_seg; ++s)
            ok = ok && d->k[s] && d->v[s] && d->Lk[s] > 0 && d->kv_div[s] > 0 && d->ldk[s] % 8 == 0 && d->ldv[s] % 8 == 0;
        if (ok) {
            const int rc = tc_attention_xs(d, stream);
            if (rc != TC_ERR_INVALID) return rc;
        }
    }
    AttnKParams p;
    memset(&p, 0, sizeof(p));
    const uint32_t box[3] = {64, 128, 1};
    {
        uint64_t dims[3] = {(uint64_t)d->heads * 64, (uint64_t)d->Lq, (uint64_t)d->q_batches};
        uint64_t str[2] = {(uint64_t)d->ldq * 2, (uint64_t)d->Lq * (uint64_t)d->ldq * 2};
        const CUtensorMap* m = get_tensor_map(d->q, 3, dims, str, box);
        if (!m) return TC_ERR_CUDA;
        p.tmQ = *m;
    }
    for (int s = 0; s < d->n_seg; ++s) {
        TC_CHECK_ARG(d->k[s] && d->v[s] && d->Lk[s] > 0 && d->kv_div[s] > 0, "tc_attention: bad kv segment");
        TC_CHECK_ARG(d->ldk[s] % 8 == 0 && d->ldv[s] % 8 == 0, "tc_attention: kv strides must be multiples of 8");
        const int kvb = (d->q_batches + d->kv_div[s] - 1) / d->kv_div[s];
        uint64_t dims[3] = {(uint64_t)d->heads * 64, (uint64_t)d->Lk[s], (uint64_t)kvb};
        uint64_t strk[2] = {(uint64_t)d->ldk[s] * 2, (uint64_t)d->Lk[s] * (uint64_t)d->ldk[s] * 2};
        uint64_t strv[2] = {(uint64_t)d->ldv[s] * 2, (uint64_t)d->Lk[s] * (uint64_t)d->ldv[s] * 2};
        const CUtensorMap* mk = get_tensor_map(d->k[s], 3, dims, strk, box);
        const CUtensorMap* mv = get_tensor_map(d->v[s], 3, dims, strv, box);
        if (!mk || !mv) return TC_ERR_CUDA;
        p.tmK[s] = *mk;
        p.tmV[s] = *mv;
        p.Lk[s] = d->Lk[s];
        p.kv_div[s] = d->kv_div[s];
    }
    p.Lq = d->Lq;
    p.heads = d->heads;
    p.n_seg = d->n_seg;
    p.out = reinterpret_cast<__half*>(d->out);
    p.ldo = d->ldo;
    p.scale_log2 = d->scale * 1.4426950408889634f;
    const size_t smem_bytes = 10 * kTileBytes + 1024 + 256;
    static bool attr2_set = false;
    if (!attr2_set) {
        int rc = check_cuda(cudaFuncSetAttribute(tc_attn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)smem_bytes),
                            "cudaFuncSetAttribute(tc_attn_kernel<false>)");
        if (rc) return rc;
        rc = check_cuda(cudaFuncSetAttribute(tc_attn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem_bytes),
                        "cudaFuncSetAttribute(tc_attn_kernel<true>)");
        if (rc) return rc;
        attr2_set = true;
    }
    dim3 grid((d->Lq + 2 * kQTile - 1) / (2 * kQTile), d->heads, d->q_batches);
    if (d->n_seg == 2)
        tc_host::launch(tc_attn_kernel<true>, dim3(grid), dim3(kAttnThreads), smem_bytes, stream, 1, p);
    else
        tc_host::launch(tc_attn_kernel<false>, dim3(grid), dim3(kAttnThreads), smem_bytes, stream, 1, p);
    count_launch();
    TC_CHECK_LAUNCH("tc_attn_kernel");
    return TC_OK;
}

extern "C" int tc_temporal_attention(const void* q, const void* k, const void* v, long long ld, void* out,
                                     long long ldo, int B, int T, int P, int heads, float scale, void* stream_v) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    TC_CHECK_ARG(q && k && v && out, "tc_temporal_attention: null pointer");
    TC_CHECK_ARG(B > 0 && T > 0 && T <= 32 && P > 0 && heads > 0, "tc_temporal_attention: need 0 < T <= 32");
    TC_CHECK_ARG(ld % 8 == 0 && ldo % 8 == 0, "tc_temporal_attention: strides must be multiples of 8");
    const long long items = (long long)B * P * heads;
    const float sl2 = scale * 1.4426950408889634f;
    const __half* qp = reinterpret_cast<const __half*>(q);
    const __half* kp = reinterpret_cast<const __half*>(k);
    const __half* vp = reinterpret_cast<const __half*>(v);
    __half* op = reinterpret_cast<__half*>(out);
    if (T <= 16) {
        long long blocks = (items + 3) / 4;
        const long long cap = 16LL * sm_count();
        if (blocks > cap) blocks = cap;
        tc_host::launch(temporal_attn_mma_kernel, dim3((unsigned)blocks), dim3(128), 0, stream, 1, qp, kp, vp, ld, op, ldo, B, T, P, heads, sl2);
    } else {
        const long long blocks = (items + 3) / 4;
        tc_host::launch(temporal_attn_kernel<32>, dim3((unsigned)blocks), dim3(128), 0, stream, 1, qp, kp, vp, ld, op, ldo, B, T, P, heads, sl2);
    }
    count_launch();
    TC_CHECK_LAUNCH("temporal_attn_kernel");
    return TC_OK;
}

extern "C" int tc_softmax_rows(void* s, long long lds, int rows, int cols, float scale, void* stream_v) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    TC_CHECK_ARG(s && rows > 0 && cols > 0, "tc_softmax_rows: bad arguments");
    tc_host::launch(softmax_rows_kernel, dim3(rows), dim3(256), 0, stream, 1, reinterpret_cast<__half*>(s), lds, rows, cols,
                                                  scale * 1.4426950408889634f);
    count_launch();
    TC_CHECK_LAUNCH("softmax_rows_kernel");
    return TC_OK;
}

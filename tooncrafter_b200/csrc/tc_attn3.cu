// tooncrafter_b200 — fused attention, third generation (single K/V segment, head dim 64): softmax(Q K^T) V with
//   * S = Q K^T and O += P V on tcgen05, both accumulators in TMEM; O stays in TMEM for the whole K/V sweep
//     (accumulating MMAs) and is only rescaled when a row maximum grows by more than 2^8 (lazy rescale);
//   * P goes registers -> TMEM (tcgen05.st) and is the A operand of the PV MMA straight from tensor memory: no
//     shared-memory round trip, no proxy fence, and the PV MMA reads only V from shared memory;
//   * the whole 128-key row of S is pulled into registers with one TMEM read, after which the S accumulator is handed
//     back so the next S MMA of the same query tile overlaps this block's softmax;
//   * the row maximum uses 3-input max, scale / sum use packed f32x2 arithmetic, and a compile-time fraction of the
//     exponentials runs on the FMA pipe (Cody-Waite range reduction + degree-3 polynomial) instead of MUFU: at
//     head dim 64 the 16 MUFU lanes per SM cap an all-MUFU softmax at 50 % tensor-pipe utilisation.
//   Two query tiles per CTA ping-pong through two softmax warpgroups (one query row per thread).
//
// Reference sites: lvdm/modules/attention.py:81-209 (spatial self-attention), lvdm/models/autoencoder_dualref.py:270-341
// (dual-reference fusion attention).  Measured pipe rates behind the design: profiles/r02_pipe_rates.txt.
#include "tc_common.cuh"
#include "tc_host.h"

namespace {

constexpr int kQTile = 128;
constexpr int kKVTile = 128;
constexpr int kTileBytes = 128 * 64 * 2;   // 16 KiB: [128 rows][64 halfs], 128B-swizzled
constexpr int kStages = 3;                 // K and V rings
constexpr int kThreads = 320;              // warps 0-3 / 4-7 softmax warpgroups, 8 MMA issuer, 9 TMA producer
constexpr uint32_t kTmemCols = 512;        // S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384) P0 [384,448) P1 [448,512)
constexpr uint32_t kTmemO = 256, kTmemP = 384;
constexpr float kRescaleThreshold = 8.0f;  // log2 units: P <= 2^8 fits fp16 with room, sums stay in fp32

struct alignas(64) Attn3Params {
    CUtensorMap tmQ, tmK, tmV;
    int Lq, Lk, kv_div;
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
__device__ __forceinline__ void unpack2u(u64 v, uint32_t& lo, uint32_t& hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
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
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 add2_rm(u64 a, u64 b) {
    u64 r;
    asm("add.rm.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
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

// 2^x for a PAIR of arguments on the FMA / ALU pipes: x = fl + fr, fl = floor(x) by a round-down magic add,
// 2^fr by a degree-3 minimax polynomial on [0, 1) (relative error 8.8e-5, far below the fp16 rounding of P that
// follows), 2^fl by adding fl to the exponent field.  x is clamped at -126 (results below 2^-126 are 0 in fp16 anyway).
__device__ __forceinline__ void exp2_poly2(u64 x2, float& e0, float& e1) {
    float x0, x1;
    unpack2(x2, x0, x1);
    x0 = fmaxf(x0, -126.0f);
    x1 = fmaxf(x1, -126.0f);
    const u64 x = pack2(x0, x1);
    const u64 magic = pack2(12582912.0f, 12582912.0f);             // 1.5 * 2^23: one mantissa ulp == 1.0
    const u64 t = add2_rm(x, magic);                                // low mantissa bits = floor(x) (two's complement)
    const u64 fl = sub2(t, magic);                                  // exact
    const u64 fr = sub2(x, fl);                                     // in [0, 1)
    u64 p = fma2(fr, pack2(0.077119089663028717f, 0.077119089663028717f), pack2(0.227564394474029541f, 0.227564394474029541f));
    p = fma2(p, fr, pack2(0.695146143436431885f, 0.695146143436431885f));
    p = fma2(p, fr, pack2(1.0f, 1.0f));
    uint32_t t0, t1, p0, p1;
    unpack2u(t, t0, t1);
    unpack2u(p, p0, p1);
    e0 = __uint_as_float(p0 + (t0 << 23));
    e1 = __uint_as_float(p1 + (t1 << 23));
}

// kPoly of every 8 consecutive PAIRS of exponentials go to the polynomial, the rest to MUFU.EX2.
template <int kPoly>
__global__ void __launch_bounds__(kThreads, 1) tc_attn3_kernel(const __grid_constant__ Attn3Params p) {
    tc::pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                                  // 2 tiles
    uint8_t* sK = smem + 2 * kTileBytes;                 // kStages
    uint8_t* sV = smem + (2 + kStages) * kTileBytes;     // kStages
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (2 + 2 * kStages) * kTileBytes);
    uint64_t* bar_q = bars + 0;
    uint64_t* k_full = bars + 1;                 // [kStages]
    uint64_t* k_free = k_full + kStages;
    uint64_t* v_full = k_free + kStages;
    uint64_t* v_free = v_full + kStages;
    uint64_t* s_full = v_free + kStages;         // [2] per query tile: S accumulator written
    uint64_t* s_free = s_full + 2;               // [2] 128 arrivals: S row is in registers
    uint64_t* p_ready = s_free + 2;              // [2] 128 arrivals: P row is in TMEM
    uint64_t* o_full = p_ready + 2;              // [2] PV MMA of the block retired
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_full + 2);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int q0 = blockIdx.x * 2 * kQTile;
    const int head = blockIdx.y;
    const int qb = blockIdx.z;
    const int ntiles = (q0 + kQTile < p.Lq) ? 2 : 1;
    const int G = (p.Lk + kKVTile - 1) / kKVTile;

    if (tid == 0) {
        tc::mbar_init(bar_q, 1);
        for (int i = 0; i < kStages; ++i) {
            tc::mbar_init(&k_full[i], 1);
            tc::mbar_init(&k_free[i], (uint32_t)ntiles);   // one tcgen05.commit per query tile releases a ring slot
            tc::mbar_init(&v_full[i], 1);
            tc::mbar_init(&v_free[i], (uint32_t)ntiles);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&s_full[i], 1);
            tc::mbar_init(&s_free[i], 128);
            tc::mbar_init(&p_ready[i], 128);
            tc::mbar_init(&o_full[i], 1);
        }
        tc::fence_mbar_init();
    }
    if (warp == 8) {
        tc::tmem_alloc(tmem_ptr_smem, kTmemCols);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    tc::pdl_wait();   // prologue (barriers, TMEM) overlapped the predecessor; Q/K/V are its results

    if (warp == 9) {
        // ------------------------------------------------------------------------------ TMA producer
        if (tc::elect_one()) {
            tc::tma_prefetch_desc(&p.tmQ);
            tc::tma_prefetch_desc(&p.tmK);
            tc::tma_prefetch_desc(&p.tmV);
            tc::mbar_arrive_expect_tx(bar_q, (uint32_t)(ntiles * kTileBytes));
            for (int w = 0; w < ntiles; ++w) tc::tma_load_3d(sQ + w * kTileBytes, &p.tmQ, bar_q, head * 64, q0 + w * kQTile, qb);
        }
        __syncwarp();
        const int kvb = qb / p.kv_div;
        int st = 0;
        uint32_t ph = 0;
        for (int g = 0; g < G; ++g) {
            tc::mbar_wait(&k_free[st], ph ^ 1u);
            if (tc::elect_one()) {
                tc::mbar_arrive_expect_tx(&k_full[st], kTileBytes);
                tc::tma_load_3d(sK + st * kTileBytes, &p.tmK, &k_full[st], head * 64, g * kKVTile, kvb);
            }
            __syncwarp();
            tc::mbar_wait(&v_free[st], ph ^ 1u);
            if (tc::elect_one()) {
                tc::mbar_arrive_expect_tx(&v_full[st], kTileBytes);
                tc::tma_load_3d(sV + st * kTileBytes, &p.tmV, &v_full[st], head * 64, g * kKVTile, kvb);
            }
            __syncwarp();
            if (++st == kStages) { st = 0; ph ^= 1u; }
        }
    } else if (warp == 8) {
        // ------------------------------------------------------------------------------ MMA issuer
        const uint32_t sQ_a = tc::smem_u32(sQ), sK_a = tc::smem_u32(sK), sV_a = tc::smem_u32(sV);
        auto block_nk = [&](int g) {
            const int left = p.Lk - g * kKVTile;
            return left < kKVTile ? ((left + 15) & ~15) : kKVTile;
        };
        // S_w = Q_w K^T for the key block in ring slot `st`; called by ONE elected lane
        auto issue_s = [&](int w, int st, int nk) {
            const uint32_t idesc = tc::umma_idesc_f16(128, (uint32_t)nk, 0, 0);
            const uint64_t qd = tc::umma_desc_sw128(sQ_a + (uint32_t)w * kTileBytes);
            const uint64_t kd = tc::umma_desc_sw128(sK_a + (uint32_t)st * kTileBytes);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                tc::umma_f16(tmem_base + (uint32_t)w * 128, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc, k != 0);
            tc::umma_commit(&s_full[w]);
        };
        tc::mbar_wait(bar_q, 0);
        tc::mbar_wait(&k_full[0], 0);
        tc::tc_fence_after();
        if (tc::elect_one()) {
            const int nk = block_nk(0);
            for (int w = 0; w < ntiles; ++w) {
                issue_s(w, 0, nk);
                tc::umma_commit(&k_free[0]);          // one arrival per query tile (the barrier counts ntiles)
            }
        }
        __syncwarp();
        // Event-driven issue: the two query tiles are independent state machines
        //     [S_w(g+1) once the softmax warps hold S_w(g) in registers]  ->  [O_w += P_w(g) V(g) once P_w(g) is in TMEM]
        // polled round-robin with non-blocking barrier tests.  Issuing in a fixed order (tile 0 then tile 1) locks the two
        // softmax warpgroups into the same phase — both in their MUFU-bound exponentials, then both idle on the XU while
        // they wait for each other's MMAs (measured: 2836 cycles per key block instead of the 2048 the MUFU allows).
        int gw[2] = {0, ntiles > 1 ? 0 : G};     // block whose PV is next, per tile
        bool need_s[2] = {true, true};           // state: next action of the tile is S(g + 1)
        const long long t_start = clock64();
        while (gw[0] < G || gw[1] < G) {
            if (clock64() - t_start > 20000000000LL) {   // ~10 s: a protocol bug traps instead of hanging the GPU
                if (tc::elect_one()) printf("tc_attn3: MMA issuer timeout (block %d,%d,%d)\n", blockIdx.x, blockIdx.y, blockIdx.z);
                __trap();
            }
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                const int g = gw[w];
                if (g >= G) continue;
                if (need_s[w]) {
                    if (g + 1 >= G) {
                        need_s[w] = false;
                    } else {
                        const int st1 = (g + 1) % kStages;
                        const uint32_t ph1 = (uint32_t)(((g + 1) / kStages) & 1);
                        if (tc::mbar_try_wait(&s_free[w], (uint32_t)(g & 1)) && tc::mbar_try_wait(&k_full[st1], ph1)) {
                            tc::tc_fence_after();
                            if (tc::elect_one()) {
                                issue_s(w, st1, block_nk(g + 1));
                                tc::umma_commit(&k_free[st1]);
                            }
                            __syncwarp();
                            need_s[w] = false;
                        }
                    }
                } else {
                    const int st = g % kStages;
                    const uint32_t ph = (uint32_t)((g / kStages) & 1);
                    if (tc::mbar_try_wait(&p_ready[w], (uint32_t)(g & 1)) && tc::mbar_try_wait(&v_full[st], ph)) {
                        tc::tc_fence_after();
                        if (tc::elect_one()) {
                            const int nk = block_nk(g);
                            const uint32_t idesc_o = tc::umma_idesc_f16(128, 64, 0, 1);   // B (= V tile [keys][64]) MN-major
                            const uint64_t vd = tc::umma_desc_sw128(sV_a + (uint32_t)st * kTileBytes);
                            for (int t = 0; t < nk / 16; ++t)
                                tc::umma_f16_ts(tmem_base + kTmemO + (uint32_t)w * 64,
                                                tmem_base + kTmemP + (uint32_t)w * 64 + (uint32_t)(t * 8), vd + (uint64_t)(t * 128),
                                                idesc_o, (g != 0 || t != 0) ? 1u : 0u);
                            tc::umma_commit(&o_full[w]);
                            tc::umma_commit(&v_free[st]);
                        }
                        __syncwarp();
                        gw[w] = g + 1;
                        need_s[w] = true;
                    }
                }
            }
        }
    } else if ((warp >> 2) < ntiles) {
        // ------------------------------------------------------------------------------ softmax warpgroups
        const int w = warp >> 2;
        const int row = tid & 127;
        const uint32_t lane_off = ((uint32_t)((warp & 3) * 32)) << 16;
        const uint32_t tmem_s = tmem_base + (uint32_t)w * 128 + lane_off;
        const uint32_t tmem_o = tmem_base + kTmemO + (uint32_t)w * 64 + lane_off;
        const uint32_t tmem_p = tmem_base + kTmemP + (uint32_t)w * 64 + lane_off;
        const float c = p.scale_log2;
        const u64 c2 = pack2(c, c);
        float m_run = -INFINITY, l_run = 0.f;
        for (int g = 0; g < G; ++g) {
            const int kv_left = p.Lk - g * kKVTile;
            const int nvalid = kv_left < kKVTile ? kv_left : kKVTile;
            tc::mbar_wait(&s_full[w], (uint32_t)(g & 1));
            tc::tc_fence_after();
            uint32_t s[128];
            tc::tmem_ld32(tmem_s, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
            tc::tmem_ld32(tmem_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
            tc::tmem_ld32(tmem_s + 64, *reinterpret_cast<uint32_t(*)[32]>(&s[64]));
            tc::tmem_ld32(tmem_s + 96, *reinterpret_cast<uint32_t(*)[32]>(&s[96]));
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            tc::mbar_arrive(&s_free[w]);               // the tensor core may overwrite S with the next block now
            if (nvalid < kKVTile) {                    // ragged last block: columns past the sequence are stale TMEM
#pragma unroll
                for (int i = 0; i < 128; ++i)
                    if (i >= nvalid) s[i] = 0xff800000u;   // -inf
            }
            // ---- row maximum (3-input max, 4 chains)
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int i = 0; i < 128; i += 8) {
                mx0 = max3(mx0, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
                mx1 = max3(mx1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
                mx2 = max3(mx2, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
                mx3 = max3(mx3, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
            }
            const float m_new = fmaxf(m_run, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)));
            // ---- lazy rescale: keep the old reference maximum unless the row maximum grew by more than 2^8
            const bool grow = (m_new - m_run) * c > kRescaleThreshold;    // first block: (x - -inf) = +inf -> true
            const float m_use = grow ? m_new : m_run;
            const float alpha = ex2((m_run - m_use) * c);                  // 1 for the rows that keep their maximum
            const bool any_grow = g > 0 && __any_sync(0xffffffffu, grow);
            m_run = m_use;
            const float nm = -m_use * c;
            const u64 nm2 = pack2(nm, nm);
            // ---- p = 2^(s * c - m * c): packed FMA, MUFU or polynomial exponentials, packed row sums, fp16 pairs
            u64 sum_a = 0ull, sum_b = 0ull;            // (0.f, 0.f)
            uint32_t pk[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) {             // pair j = elements 2j, 2j + 1
                const u64 x = fma2(pack2(__uint_as_float(s[2 * j]), __uint_as_float(s[2 * j + 1])), c2, nm2);
                float e0, e1;
                if ((j & 7) < kPoly) {
                    exp2_poly2(x, e0, e1);
                } else {
                    float x0, x1;
                    unpack2(x, x0, x1);
                    e0 = ex2(x0);
                    e1 = ex2(x1);
                }
                if (j & 1) sum_b = add2(sum_b, pack2(e0, e1));
                else sum_a = add2(sum_a, pack2(e0, e1));
                pk[j] = pack_h2(e0, e1);
            }
            // Only now must the previous block's PV MMA have retired (it reads P and accumulates into O): waiting here,
            // after the ~1000-cycle exponential phase, instead of before it took a 5 % long-scoreboard stall off the
            // critical path (profiles/r02_ncu_attn3_2560_first.txt).
            if (g > 0) {
                tc::mbar_wait(&o_full[w], (uint32_t)((g - 1) & 1));
                tc::tc_fence_after();
                if (any_grow) {
                    l_run *= alpha;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        uint32_t r[16];
                        tc::tmem_ld16(tmem_o + (uint32_t)(cc * 16), r);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                        tc::tmem_st16(tmem_o + (uint32_t)(cc * 16), r);
                    }
                }
            }
            tc::tmem_st32(tmem_p, &pk[0]);
            tc::tmem_st32(tmem_p + 32, &pk[32]);
            tc::tmem_st_wait();
            tc::tc_fence_before();
            tc::mbar_arrive(&p_ready[w]);
            float a0, a1, b0, b1;
            unpack2(sum_a, a0, a1);
            unpack2(sum_b, b0, b1);
            l_run += (a0 + a1) + (b0 + b1);
        }
        // ---- epilogue: O / l -> fp16 -> global
        tc::mbar_wait(&o_full[w], (uint32_t)((G - 1) & 1));
        tc::tc_fence_after();
        const float inv_l = 1.0f / l_run;
        const int qrow = q0 + w * kQTile + row;
        __half* dst = p.out + ((long long)qb * p.Lq + qrow) * p.ldo + head * 64;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
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
    if (warp == 8) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, kTmemCols);
    }
}

template <int kPoly>
int launch_attn3(const Attn3Params& p, dim3 grid, size_t smem_bytes, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        int rc = tc_host::check_cuda(cudaFuncSetAttribute(tc_attn3_kernel<kPoly>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                          (int)smem_bytes),
                                     "cudaFuncSetAttribute(tc_attn3_kernel)");
        if (rc) return rc;
        attr_set = true;
    }
    tc_host::launch(tc_attn3_kernel<kPoly>, grid, dim3(kThreads), smem_bytes, stream, 1, p);
    return 0;
}

}  // namespace

using namespace tc_host;

// Single-segment attention through the third-generation kernel.  `poly_of_8`: how many of every 8 exponential pairs
// run on the FMA pipe (0 = all MUFU).  Called by tc_attention (tc_attn.cu) after it validated the descriptor.
int tc_attention_v3(const TcAttention* d, int poly_of_8, cudaStream_t stream) {
    Attn3Params p;
    memset(&p, 0, sizeof(p));
    const uint32_t box[3] = {64, 128, 1};
    {
        uint64_t dims[3] = {(uint64_t)d->heads * 64, (uint64_t)d->Lq, (uint64_t)d->q_batches};
        uint64_t str[2] = {(uint64_t)d->ldq * 2, (uint64_t)d->Lq * (uint64_t)d->ldq * 2};
        const CUtensorMap* m = get_tensor_map(d->q, 3, dims, str, box);
        if (!m) return TC_ERR_CUDA;
        p.tmQ = *m;
    }
    {
        const int kvb = (d->q_batches + d->kv_div[0] - 1) / d->kv_div[0];
        uint64_t dims[3] = {(uint64_t)d->heads * 64, (uint64_t)d->Lk[0], (uint64_t)kvb};
        uint64_t strk[2] = {(uint64_t)d->ldk[0] * 2, (uint64_t)d->Lk[0] * (uint64_t)d->ldk[0] * 2};
        uint64_t strv[2] = {(uint64_t)d->ldv[0] * 2, (uint64_t)d->Lk[0] * (uint64_t)d->ldv[0] * 2};
        const CUtensorMap* mk = get_tensor_map(d->k[0], 3, dims, strk, box);
        const CUtensorMap* mv = get_tensor_map(d->v[0], 3, dims, strv, box);
        if (!mk || !mv) return TC_ERR_CUDA;
        p.tmK = *mk;
        p.tmV = *mv;
    }
    p.Lq = d->Lq;
    p.Lk = d->Lk[0];
    p.kv_div = d->kv_div[0];
    p.out = reinterpret_cast<__half*>(d->out);
    p.ldo = d->ldo;
    p.scale_log2 = d->scale * 1.4426950408889634f;
    const size_t smem_bytes = (size_t)(2 + 2 * kStages) * kTileBytes + 1024 + 512;
    dim3 grid((d->Lq + 2 * kQTile - 1) / (2 * kQTile), d->heads, d->q_batches);
    int rc;
    switch (poly_of_8) {
        case 0: rc = launch_attn3<0>(p, grid, smem_bytes, stream); break;
        case 1: rc = launch_attn3<1>(p, grid, smem_bytes, stream); break;
        case 2: rc = launch_attn3<2>(p, grid, smem_bytes, stream); break;
        case 4: rc = launch_attn3<4>(p, grid, smem_bytes, stream); break;
        default: rc = launch_attn3<3>(p, grid, smem_bytes, stream); break;
    }
    if (rc) return rc;
    count_launch();
    TC_CHECK_LAUNCH("tc_attn3_kernel");
    return TC_OK;
}

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
//   Two query tiles per CTA run through two softmax warpgroups (one query row per thread), each tile with its own MMA-issuer
//   warp (attn_issue_tile).
//
// Reference sites: lvdm/modules/attention.py:81-209 (spatial self-attention), lvdm/models/autoencoder_dualref.py:270-341
// (dual-reference fusion attention).  Measured pipe rates behind the design: profiles/r02_pipe_rates.txt.
#include <stdlib.h>

#include "tc_common.cuh"
#include "tc_host.h"

namespace {

constexpr int kQTile = 128;
constexpr int kKVTile = 128;
constexpr int kTileBytes = 128 * 64 * 2;   // 16 KiB: [128 rows][64 halfs], 128B-swizzled
constexpr int kStages = 3;                 // K and V rings
constexpr int kThreads = 352;              // warps 0-3 / 4-7 softmax warpgroups, 8 / 10 MMA issuers (tile 0 / 1), 9 TMA producer
constexpr uint32_t kTmemCols = 512;        // S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384) P0 [384,448) P1 [448,512)
constexpr uint32_t kTmemO = 256, kTmemP = 384;
constexpr float kRescaleThreshold = 8.0f;  // log2 units: P <= 2^8 fits fp16 with room, sums stay in fp32

struct alignas(64) Attn3Params {
    CUtensorMap tmQ, tmK, tmV;
    int Lq, Lk, kv_div;
    __half* out;
    long long ldo;
    float scale_log2;
    int pingpong;   // the two query tiles take turns on the exponential phase (named-barrier token), see the softmax loops
    int stagger;    // cycles by which query tile 1 starts late (keeps the two tiles' exponential phases out of step)
    int single;     // one query tile per CTA (192 threads, 256 TMEM columns, two K/V stages): two independent CTAs per SM
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
// volatile: the exponentials must stay between the named-barrier token operations of the softmax loops (asm volatile
// statements keep their order; as plain asm the compiler moved the bar.arrive in front of nearly all MUFU instructions and
// the two query tiles' exponential phases overlapped again — in-kernel timeline, profiles/r02_attn_trace.txt)
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// In-kernel timeline (only with -DTC_ATTN_TRACE=1, scripts/trace_attn.py): clock64() stamps of CTA (0,0,0) — one softmax
// thread per query tile and the MMA issuer — per key block.
#if defined(TC_ATTN_TRACE) && TC_ATTN_TRACE
constexpr int kAtBlocks = 24, kAtSlots = 16;
__device__ unsigned long long g_tc_attn_trace[kAtBlocks * kAtSlots];
#define TC_ATRACE(slot, g)                                                                        \
    if ((blockIdx.x | blockIdx.y | blockIdx.z) == 0 && (g) < kAtBlocks)                          \
        g_tc_attn_trace[(g) * kAtSlots + (slot)] = (unsigned long long)clock64();
#else
#define TC_ATRACE(slot, g)
#endif

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

// MMA issuer of ONE query tile (one warp per tile, one elected lane issues): S_w(g+1) = Q_w K(g+1)^T as soon as the tile's
// softmax threads hold S_w(g) in registers, then O_w += P_w(g) V(g) with P as the TMEM A operand.  One warp used to issue
// for both tiles: at ~100 cycles of issue work per tcgen05.mma (R2UR descriptor traffic on the uniform datapath) 24 small
// MMAs + 6 commits per key block kept that warp busy ~2200 of the ~2900 cycles a block took and coupled the two tiles' timing
// (in-kernel timeline, profiles/r02_attn_trace.txt).  K / V ring stages are released by BOTH issuers (barrier count =
// number of active tiles); tcgen05 ops of different issuers are unordered, which is fine: the tiles share no TMEM.
struct AttnBars {
    uint64_t *bar_q, *k_full, *k_free, *v_full, *v_free, *s_full, *s_free, *p_ready, *o_full;
};
__device__ __forceinline__ void attn_issue_tile(const Attn3Params& p, const AttnBars& b, int w, int G, uint32_t tmem_base, uint32_t sQ_a,
                                                uint32_t sK_a, uint32_t sV_a, int nst, uint32_t tm_o, uint32_t tm_p) {
    auto block_nk = [&](int g) {
        const int left = p.Lk - g * kKVTile;
        return left < kKVTile ? ((left + 15) & ~15) : kKVTile;
    };
    const uint64_t qd = tc::umma_desc_sw128(sQ_a + (uint32_t)w * kTileBytes);
    const uint64_t kd0 = tc::umma_desc_sw128(sK_a);
    const uint32_t s_tmem = tmem_base + (uint32_t)w * 128, o_tmem = tmem_base + tm_o + (uint32_t)w * 64,
                   p_tmem = tmem_base + tm_p + (uint32_t)w * 64;
    auto issue_s = [&](int st, int nk) {          // called by ONE elected lane
        const uint32_t idesc = tc::umma_idesc_f16(128, (uint32_t)nk, 0, 0);
        const uint64_t kd = kd0 + (uint64_t)(st * (kTileBytes >> 4));
#pragma unroll
        for (int k = 0; k < 4; ++k) tc::umma_f16(s_tmem, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc, k != 0);
        tc::umma_commit(&b.s_full[w]);
        tc::umma_commit(&b.k_free[st]);
    };
    // one elected lane runs the whole loop, waits included (per-block ELECT / reconvergence and vector-register descriptor
    // arithmetic were a good part of the ~100 cycles of issue work per MMA)
    if (tc::elect_one()) {
        tc::mbar_wait(b.bar_q, 0);
        tc::mbar_wait(&b.k_full[0], 0);
        tc::tc_fence_after();
        issue_s(0, block_nk(0));
        const uint32_t idesc_o = tc::umma_idesc_f16(128, 64, 0, 1);       // B (= V tile [keys][64]) MN-major
        const uint64_t vd0 = tc::umma_desc_sw128(sV_a);
        int st = 0;
        uint32_t ph = 0;
        for (int g = 0; g < G; ++g) {
            const int nk = block_nk(g);
            if (g + 1 < G) {
                const int st1 = (st + 1 == nst) ? 0 : st + 1;
                const uint32_t ph1 = (st + 1 == nst) ? (ph ^ 1u) : ph;
                tc::mbar_wait(&b.k_full[st1], ph1);
                tc::mbar_wait(&b.s_free[w], (uint32_t)(g & 1));
                tc::tc_fence_after();
                TC_ATRACE(12 + w, g)
                issue_s(st1, block_nk(g + 1));
            }
            tc::mbar_wait(&b.v_full[st], ph);
            tc::mbar_wait(&b.p_ready[w], (uint32_t)(g & 1));
            tc::tc_fence_after();
            TC_ATRACE(14 + w, g)
            const uint64_t vd = vd0 + (uint64_t)(st * (kTileBytes >> 4));
            if (nk == kKVTile) {
#pragma unroll
                for (int t = 0; t < kKVTile / 16; ++t)
                    tc::umma_f16_ts(o_tmem, p_tmem + (uint32_t)(t * 8), vd + (uint64_t)(t * 128), idesc_o, (g != 0 || t != 0) ? 1u : 0u);
            } else {
                for (int t = 0; t < nk / 16; ++t)
                    tc::umma_f16_ts(o_tmem, p_tmem + (uint32_t)(t * 8), vd + (uint64_t)(t * 128), idesc_o, (g != 0 || t != 0) ? 1u : 0u);
            }
            tc::umma_commit(&b.o_full[w]);
            tc::umma_commit(&b.v_free[st]);
            if (++st == nst) { st = 0; ph ^= 1u; }
        }
    }
    __syncwarp();
}

// kPoly of every 8 consecutive PAIRS of exponentials go to the polynomial, the rest to MUFU.EX2.
template <int kPoly>
__global__ void __launch_bounds__(kThreads, 1) tc_attn3_kernel(const __grid_constant__ Attn3Params p) {
    tc::pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // single mode: ONE query tile per CTA, 192 threads, 256 TMEM columns, two K/V stages — two such CTAs share an SM and run
    // out of step by themselves (independent barriers), so one CTA's exponentials overlap the other's TMEM / PV round trip
    const bool single = p.single != 0;
    const int nq = single ? 1 : 2, nst = single ? 2 : kStages;
    const uint32_t tm_o = single ? 128u : kTmemO, tm_p = single ? 192u : kTmemP, tm_cols = single ? 256u : kTmemCols;
    const int warp_mma0 = single ? 4 : 8, warp_tma = single ? 5 : 9, warp_mma1 = single ? -1 : 10, n_soft = single ? 4 : 8;
    uint8_t* sQ = smem;                                  // nq tiles
    uint8_t* sK = smem + nq * kTileBytes;                // nst stages
    uint8_t* sV = sK + nst * kTileBytes;                 // nst stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + nst * kTileBytes);
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
    float* s_tok = reinterpret_cast<float*>(bars + 32);   // [0] = 0.0f, [1 + tid] sink: data fences of the exponential-phase token

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int q0 = blockIdx.x * nq * kQTile;
    const int head = blockIdx.y;
    const int qb = blockIdx.z;
    const int ntiles = (!single && q0 + kQTile < p.Lq) ? 2 : 1;
    const int G = (p.Lk + kKVTile - 1) / kKVTile;

    if (tid == 0) {
        s_tok[0] = 0.f;
        tc::mbar_init(bar_q, 1);
        for (int i = 0; i < kStages; ++i) {
            tc::mbar_init(&k_full[i], 1);
            tc::mbar_init(&k_free[i], (uint32_t)ntiles);      // released by every active tile's MMA issuer
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
    if (warp == warp_mma0) {
        tc::tmem_alloc(tmem_ptr_smem, tm_cols);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    tc::pdl_wait();   // prologue (barriers, TMEM) overlapped the predecessor; Q/K/V are its results

    if (warp == warp_tma) {
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
            if (++st == nst) { st = 0; ph ^= 1u; }
        }
    } else if (warp == warp_mma0 || warp == warp_mma1) {
        // ------------------------------------------------------------------------------ MMA issuers (one per query tile)
        const int w = warp == warp_mma0 ? 0 : 1;
        if (w < ntiles) {
            const AttnBars b{bar_q, k_full, k_free, v_full, v_free, s_full, s_free, p_ready, o_full};
            attn_issue_tile(p, b, w, G, tmem_base, tc::smem_u32(sQ), tc::smem_u32(sK), tc::smem_u32(sV), nst, tm_o, tm_p);
        }
    } else if (warp < n_soft && (warp >> 2) < ntiles) {
        // ------------------------------------------------------------------------------ softmax warpgroups
        const int w = warp >> 2;
        const int row = tid & 127;
        const uint32_t lane_off = ((uint32_t)((warp & 3) * 32)) << 16;
        const uint32_t tmem_s = tmem_base + (uint32_t)w * 128 + lane_off;
        const uint32_t tmem_o = tmem_base + tm_o + (uint32_t)w * 64 + lane_off;
        const uint32_t tmem_p = tmem_base + tm_p + (uint32_t)w * 64 + lane_off;
        const float c = p.scale_log2;
        const u64 c2 = pack2(c, c);
        float m_run = -INFINITY, l_run = 0.f;
        const bool pp = p.pingpong != 0 && ntiles == 2;
        const uint32_t tok_zero = tc::smem_u32(s_tok), tok_sink = tc::smem_u32(s_tok + 1 + tid);
        if (pp && w == 1) asm volatile("bar.arrive 14, 256;" ::: "memory");   // tile 0 goes first
        if (w == 1 && p.stagger > 0 && ntiles == 2) {
            // Both tiles share the 16 MUFU lanes.  Started together they stay in step (exponentials of both at half rate,
            // then both in the TMEM / maximum / PV part with MUFU idle); started half a block apart one tile's
            // exponentials run under the other's PV round trip.  Nothing pulls the tiles back into step afterwards.
            const long long t0 = clock64();
            while (clock64() - t0 < (long long)p.stagger) { }
        }
        for (int g = 0; g < G; ++g) {
            const int kv_left = p.Lk - g * kKVTile;
            const int nvalid = kv_left < kKVTile ? kv_left : kKVTile;
            tc::mbar_wait(&s_full[w], (uint32_t)(g & 1));
            tc::tc_fence_after();
            if ((tid & 127) == 0) { TC_ATRACE(w * 6 + 0, g) }      // S(g) available
            uint32_t s[128];
            tc::tmem_ld32(tmem_s, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
            tc::tmem_ld32(tmem_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
            tc::tmem_ld32(tmem_s + 64, *reinterpret_cast<uint32_t(*)[32]>(&s[64]));
            tc::tmem_ld32(tmem_s + 96, *reinterpret_cast<uint32_t(*)[32]>(&s[96]));
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            tc::mbar_arrive(&s_free[w]);               // the tensor core may overwrite S with the next block now
            if ((tid & 127) == 0) { TC_ATRACE(w * 6 + 1, g) }      // S row in registers
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
            if ((tid & 127) == 0) { TC_ATRACE(w * 6 + 2, g) }      // row maximum known
            if (g > 0) {
                tc::mbar_wait(&o_full[w], (uint32_t)((g - 1) & 1));       // PV of the previous block retired: P and O are ours
                tc::tc_fence_after();
                if ((tid & 127) == 0) { TC_ATRACE(w * 6 + 3, g) }  // PV(g-1) retired
                if (__any_sync(0xffffffffu, grow)) {
                    const float alpha = ex2((m_run - m_use) * c);          // 1 for the rows that keep their maximum
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
                    tc::tmem_st_wait();
                }
            }
            m_run = m_use;
            const float nm = -m_use * c;
            // ---- exponential phase.  Both query tiles share the SM's 16 MUFU lanes; left alone they run IN PHASE (both in
            // the TMEM-read / maximum / PV-wait part, then both on MUFU at half rate) and the XU pipe idles a third of the
            // time.  A token passed through two named barriers makes the phases alternate: one tile's exponentials run at the
            // full MUFU rate while the other tile's PV MMA, TMEM traffic and row maximum proceed underneath.
            // ptxas schedules arithmetic freely across bar.* (it moved the arrive in front of ~all MUFU instructions), so the
            // phase is fenced by DATA: every exponential depends on a volatile shared load (of 0.0f) issued after the
            // bar.sync, and the bar.arrive follows a volatile shared store of the row sum, which needs every exponential.
            float nm_t = nm;
            if (pp) {
                asm volatile("bar.sync %0, 256;" ::"r"(14 + w) : "memory");
                float z;
                asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(z) : "r"(tok_zero) : "memory");
                nm_t += z;
            }
            const u64 nm2 = pack2(nm_t, nm_t);
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
            float a0, a1, b0, b1;
            unpack2(sum_a, a0, a1);
            unpack2(sum_b, b0, b1);
            const float l_blk = (a0 + a1) + (b0 + b1);
            if (pp) {
                asm volatile("st.volatile.shared.f32 [%0], %1;" ::"r"(tok_sink), "f"(l_blk) : "memory");
                asm volatile("bar.arrive %0, 256;" ::"r"(14 + (w ^ 1)) : "memory");   // the other tile's turn
            }
            if ((tid & 127) == 0) { TC_ATRACE(w * 6 + 4, g) }      // exponentials done
            tc::tmem_st32(tmem_p, &pk[0]);
            tc::tmem_st32(tmem_p + 32, &pk[32]);
            tc::tmem_st_wait();
            tc::tc_fence_before();
            tc::mbar_arrive(&p_ready[w]);
            if ((tid & 127) == 0) { TC_ATRACE(w * 6 + 5, g) }      // P in TMEM, arrived
            l_run += l_blk;
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
    if (warp == warp_mma0) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, tm_cols);
    }
}

template <int kPoly>
int launch_attn3(const Attn3Params& p, dim3 grid, size_t smem_bytes, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        int rc = tc_host::check_cuda(cudaFuncSetAttribute(tc_attn3_kernel<kPoly>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                          (int)((size_t)(2 + 2 * kStages) * kTileBytes + 1024 + 512 + 4096)),
                                     "cudaFuncSetAttribute(tc_attn3_kernel)");
        if (rc) return rc;
        attr_set = true;
    }
    tc_host::launch(tc_attn3_kernel<kPoly>, grid, dim3(p.single ? 192 : kThreads), smem_bytes, stream, 1, p);
    return 0;
}



// ===================================================================================== small-KV cross attention
// Text + image cross attention of the spatial transformers (lvdm/modules/attention.py:126-142 / 196-207): every query
// attends to 77 text keys and to 16 image keys and the two softmax-weighted sums are ADDED.  93 keys is 2-3 % of a
// self-attention's work per query tile: on the tcgen05 kernels (128-row tiles, TMEM, mbarrier pipelines) the launch
// is pure latency — 50-85 TF/s, and a persistent tcgen05 variant with resident K/V measured no better (only two
// query tiles fit in TMEM at a time).  This kernel takes the other road: warp-level mma.sync with MANY warps in
// flight.  A CTA (4 warps) keeps the K / V rows of both segments of one (frame, head) in shared memory and every warp
// streams 16-query groups:  S = Q K^T (m16n8k16, K fragments by ldmatrix) -> two independent fp32 softmaxes on the
// accumulator fragments, normalised BEFORE the fp16 rounding -> P = hi + lo fp16 parts (with 16 image keys the rounding
// of P is not averaged away) -> O = P V (V fragments by ldmatrix.trans) -> fp16 -> global.
// Needs ceil16(Lk0) + ceil16(Lk1) <= 96 keys (12 n-tiles of S accumulators per thread).
constexpr int kXsMaxNT = 12;           // n-tiles (8 keys each) of S per warp
constexpr int kXsThreads = 128;

struct AttnXsParams {
    const __half* q;
    const __half* k[2];
    const __half* v[2];
    long long ldq, ldk[2], ldv[2];
    int Lq, heads, n_seg, Lk[2], kv_div[2];
    __half* out;
    long long ldo;
    float scale_log2;
};

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kXsThreads, 4) tc_attn_xs_kernel(const AttnXsParams p) {
    tc::pdl_wait();   // no early launch_dependents (see temporal_attn_mma_kernel)
    __shared__ __align__(128) uint8_t sK[kXsMaxNT * 8 * 128];   // [key][64 halfs], 16-byte chunks XOR-swizzled by key & 7
    __shared__ __align__(128) uint8_t sV[kXsMaxNT * 8 * 128];
    __shared__ __align__(128) uint8_t sQO[kXsThreads / 32][16 * 128];   // per warp: 16 query rows x 64 halfs (Q in, then O out), 16-byte chunks XOR-swizzled by row & 7
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g8 = lane >> 2, t4 = lane & 3;
    const int head = blockIdx.y, qb = blockIdx.z;
    const int nk0 = (p.Lk[0] + 15) & ~15, nk1 = p.n_seg > 1 ? ((p.Lk[1] + 15) & ~15) : 0;
    const int ntot = (nk0 + nk1) >> 3, nt0 = nk0 >> 3;

    // ---- K / V of both segments -> shared memory (rows past a segment's length are zero).  All global loads of a
    // thread are issued before its first shared store (load -> store per chunk serialised six L2 round trips).
    {
        constexpr int kMaxIt = (kXsMaxNT * 8 * 8 + kXsThreads - 1) / kXsThreads;   // 16-byte chunks per thread
        uint4 uk[kMaxIt], uv[kMaxIt];
#pragma unroll
        for (int it = 0; it < kMaxIt; ++it) {
            const int ch = tid + it * kXsThreads;
            const int row = ch >> 3, c16 = ch & 7;
            const int sgm = row >= nk0 ? 1 : 0;
            const int r = row - (sgm ? nk0 : 0);
            uk[it] = make_uint4(0u, 0u, 0u, 0u);
            uv[it] = uk[it];
            if (ch < (nk0 + nk1) * 8 && r < p.Lk[sgm]) {
                const long long kvb = qb / p.kv_div[sgm];
                uk[it] = *reinterpret_cast<const uint4*>(p.k[sgm] + (kvb * p.Lk[sgm] + r) * p.ldk[sgm] + head * 64 + c16 * 8);
                uv[it] = *reinterpret_cast<const uint4*>(p.v[sgm] + (kvb * p.Lk[sgm] + r) * p.ldv[sgm] + head * 64 + c16 * 8);
            }
        }
#pragma unroll
        for (int it = 0; it < kMaxIt; ++it) {
            const int ch = tid + it * kXsThreads;
            const int row = ch >> 3, c16 = ch & 7;
            if (ch < (nk0 + nk1) * 8) {
                const int off = row * 128 + ((c16 ^ (row & 7)) << 4);
                *reinterpret_cast<uint4*>(sK + off) = uk[it];
                *reinterpret_cast<uint4*>(sV + off) = uv[it];
            }
        }
    }
    __syncthreads();
    const uint32_t sK_a = tc::smem_u32(sK), sV_a = tc::smem_u32(sV);
    const float c = p.scale_log2;
    const int n_groups = (p.Lq + 15) >> 4;

    // Q rows of a 16-query group come in as 16-byte vectors (8 lanes cover the 128 bytes of a row: every sector fully used;
    // the 4-byte fragment loads of the first version used half of every 32-byte sector, and so did its stores), pass through
    // this warp's staging tile and become A fragments by ldmatrix.  The NEXT group's vectors are requested while this
    // group's softmax / PV run (a warp's first HMMA sat on this load: 12 % of the samples).
    const uint32_t sQO_a = tc::smem_u32(&sQO[warp][0]);
    const int st_row = lane >> 3, st_chunk = lane & 7;                 // staging: lane <-> (row st_row + 4 i, 16-byte chunk)
    auto load_q = [&](int grp, uint4 (&dst)[4]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = grp * 16 + st_row + 4 * i;
            dst[i] = make_uint4(0u, 0u, 0u, 0u);
            if (grp < n_groups && r < p.Lq)
                dst[i] = *reinterpret_cast<const uint4*>(p.q + ((long long)qb * p.Lq + r) * p.ldq + head * 64 + st_chunk * 8);
        }
    };
    uint4 qn[4];
    load_q((int)blockIdx.x * 4 + warp, qn);
    for (int grp = (int)blockIdx.x * 4 + warp; grp < n_groups; grp += (int)gridDim.x * 4) {
        const int r0 = grp * 16 + g8, r1 = r0 + 8;
        (void)r0; (void)r1;
        uint32_t qa[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = st_row + 4 * i;
            const uint32_t a = sQO_a + (uint32_t)(row * 128 + ((st_chunk ^ (row & 7)) << 4));
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(qn[i].x), "r"(qn[i].y), "r"(qn[i].z), "r"(qn[i].w) : "memory");
        }
        __syncwarp();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            // matrices: (rows 0-7, k 0-7), (rows 8-15, k 0-7), (rows 0-7, k 8-15), (rows 8-15, k 8-15) of k-step ks
            const int row = ((lane >> 3) & 1) * 8 + (lane & 7);
            const int chunk = 2 * ks + (lane >> 4);
            const uint32_t a = sQO_a + (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                         : "=r"(qa[ks][0]), "=r"(qa[ks][1]), "=r"(qa[ks][2]), "=r"(qa[ks][3])
                         : "r"(a));
        }
        // ---- S = Q K^T: per n-tile two ldmatrix.x4 (d chunks 0-31, 32-63 of keys 8nt..8nt+7)
        float sacc[kXsMaxNT][4];
#pragma unroll
        for (int nt = 0; nt < kXsMaxNT; ++nt) {
#pragma unroll
            for (int i = 0; i < 4; ++i) sacc[nt][i] = 0.f;
            if (nt < ntot) {
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    // lane i supplies the row address of matrix i / 8: matrices = 16-byte d-chunks 4hf .. 4hf+3 of these keys
                    const int krow = 8 * nt + (lane & 7);
                    const int chunk = 4 * hf + (lane >> 3);
                    const uint32_t addr = sK_a + (uint32_t)(krow * 128 + ((chunk ^ (krow & 7)) << 4));
                    uint32_t b0, b1, b2, b3;
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                                 : "r"(addr));
                    mma16816(sacc[nt], qa[2 * hf], b0, b1);
                    mma16816(sacc[nt], qa[2 * hf + 1], b2, b3);
                }
            }
        }
        load_q(grp + (int)gridDim.x * 4, qn);
        // ---- two independent softmaxes: keys [0, Lk0) and [nk0, nk0 + Lk1); rows g8 (values 0,1) and g8 + 8 (values 2,3)
        float m00 = -INFINITY, m01 = -INFINITY, m10 = -INFINITY, m11 = -INFINITY;   // m<seg><row half>
#pragma unroll
        for (int nt = 0; nt < kXsMaxNT; ++nt) {
            if (nt < ntot) {
                const bool s1 = nt >= nt0;
                const int key0 = 8 * nt + 2 * t4 - (s1 ? nk0 : 0);
                const int lim = s1 ? p.Lk[1] : p.Lk[0];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const bool okk = key0 + i < lim;
                    sacc[nt][i] = okk ? sacc[nt][i] : -INFINITY;
                    sacc[nt][2 + i] = okk ? sacc[nt][2 + i] : -INFINITY;
                    if (s1) { m10 = fmaxf(m10, sacc[nt][i]); m11 = fmaxf(m11, sacc[nt][2 + i]); }
                    else { m00 = fmaxf(m00, sacc[nt][i]); m01 = fmaxf(m01, sacc[nt][2 + i]); }
                }
            }
        }
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
            m00 = fmaxf(m00, __shfl_xor_sync(0xffffffffu, m00, o));
            m01 = fmaxf(m01, __shfl_xor_sync(0xffffffffu, m01, o));
            m10 = fmaxf(m10, __shfl_xor_sync(0xffffffffu, m10, o));
            m11 = fmaxf(m11, __shfl_xor_sync(0xffffffffu, m11, o));
        }
        const float n00 = -m00 * c, n01 = -m01 * c;
        const float n10 = (m10 == -INFINITY) ? 0.f : -m10 * c, n11 = (m11 == -INFINITY) ? 0.f : -m11 * c;
        float l00 = 0.f, l01 = 0.f, l10 = 0.f, l11 = 0.f;
#pragma unroll
        for (int nt = 0; nt < kXsMaxNT; ++nt) {
            if (nt < ntot) {
                const bool s1 = nt >= nt0;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float e0 = ex2(fmaf(sacc[nt][i], c, s1 ? n10 : n00));
                    const float e1 = ex2(fmaf(sacc[nt][2 + i], c, s1 ? n11 : n01));
                    sacc[nt][i] = e0;
                    sacc[nt][2 + i] = e1;
                    if (s1) { l10 += e0; l11 += e1; } else { l00 += e0; l01 += e1; }
                }
            }
        }
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
            l00 += __shfl_xor_sync(0xffffffffu, l00, o);
            l01 += __shfl_xor_sync(0xffffffffu, l01, o);
            l10 += __shfl_xor_sync(0xffffffffu, l10, o);
            l11 += __shfl_xor_sync(0xffffffffu, l11, o);
        }
        const float i00 = 1.0f / l00, i01 = 1.0f / l01;
        const float i10 = l10 > 0.f ? 1.0f / l10 : 0.f, i11 = l11 > 0.f ? 1.0f / l11 : 0.f;
        // ---- O = P V with P = hi + lo (fp16 parts of the normalised probabilities)
        float oacc[8][4];
#pragma unroll
        for (int nd = 0; nd < 8; ++nd)
#pragma unroll
            for (int i = 0; i < 4; ++i) oacc[nd][i] = 0.f;
#pragma unroll
        for (int kk = 0; kk < kXsMaxNT / 2; ++kk) {
            if (2 * kk < ntot) {
                const bool s1 = 2 * kk >= nt0;          // nk0 is a multiple of 16: a k-step never straddles the segments
                const float ia = s1 ? i10 : i00, ib = s1 ? i11 : i01;
                uint32_t ph[4], pl[4];
                {
                    const float a0 = sacc[2 * kk][0] * ia, a1 = sacc[2 * kk][1] * ia, b0 = sacc[2 * kk][2] * ib, b1 = sacc[2 * kk][3] * ib;
                    const float c0 = sacc[2 * kk + 1][0] * ia, c1 = sacc[2 * kk + 1][1] * ia, d0 = sacc[2 * kk + 1][2] * ib, d1 = sacc[2 * kk + 1][3] * ib;
                    const __half2 h0 = __floats2half2_rn(a0, a1), h1 = __floats2half2_rn(b0, b1);
                    const __half2 h2 = __floats2half2_rn(c0, c1), h3 = __floats2half2_rn(d0, d1);
                    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1), f2 = __half22float2(h2), f3 = __half22float2(h3);
                    ph[0] = *reinterpret_cast<const uint32_t*>(&h0);
                    ph[1] = *reinterpret_cast<const uint32_t*>(&h1);
                    ph[2] = *reinterpret_cast<const uint32_t*>(&h2);
                    ph[3] = *reinterpret_cast<const uint32_t*>(&h3);
                    pl[0] = pack_h2(a0 - f0.x, a1 - f0.y);
                    pl[1] = pack_h2(b0 - f1.x, b1 - f1.y);
                    pl[2] = pack_h2(c0 - f2.x, c1 - f2.y);
                    pl[3] = pack_h2(d0 - f3.x, d1 - f3.y);
                }
#pragma unroll
                for (int nd = 0; nd < 8; nd += 2) {
                    // matrices (keys 16kk..+7, d-chunk nd), (keys 16kk+8..+15, nd), (keys ..+7, nd+1), (keys +8.., nd+1)
                    const int vrow = 16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8;
                    const int chunk = nd + (lane >> 4);
                    const uint32_t addr = sV_a + (uint32_t)(vrow * 128 + ((chunk ^ (vrow & 7)) << 4));
                    uint32_t b0, b1, b2, b3;
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                                 : "r"(addr));
                    mma16816(oacc[nd], pl, b0, b1);
                    mma16816(oacc[nd + 1], pl, b2, b3);
                    mma16816(oacc[nd], ph, b0, b1);
                    mma16816(oacc[nd + 1], ph, b2, b3);
                }
            }
        }
        // ---- O: fragments -> staging tile (the Q fragments are long consumed) -> 16-byte row-contiguous global stores
        __syncwarp();
#pragma unroll
        for (int nd = 0; nd < 8; ++nd) {
            const uint32_t a0 = sQO_a + (uint32_t)(g8 * 128 + ((nd ^ (g8 & 7)) << 4) + t4 * 4);
            const uint32_t a1 = sQO_a + (uint32_t)((g8 + 8) * 128 + ((nd ^ (g8 & 7)) << 4) + t4 * 4);      // (g8 + 8) & 7 == g8 & 7
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(a0), "r"(pack_h2(oacc[nd][0], oacc[nd][1])) : "memory");
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(a1), "r"(pack_h2(oacc[nd][2], oacc[nd][3])) : "memory");
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = st_row + 4 * i;
            const int r = grp * 16 + row;
            const uint32_t a = sQO_a + (uint32_t)(row * 128 + ((st_chunk ^ (row & 7)) << 4));
            uint4 u;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(a) : "memory");
            if (r < p.Lq) *reinterpret_cast<uint4*>(p.out + ((long long)qb * p.Lq + r) * p.ldo + head * 64 + st_chunk * 8) = u;
        }
        __syncwarp();
    }
}

}  // namespace

using namespace tc_host;

extern "C" int tc_debug_read_attn_trace(unsigned long long* host_dst, int count) {
#if defined(TC_ATTN_TRACE) && TC_ATTN_TRACE
    if (!host_dst || count <= 0 || count > kAtBlocks * kAtSlots) return tc_host::fail(TC_ERR_INVALID, "tc_debug_read_attn_trace: bad count");
    return tc_host::check_cuda(cudaMemcpyFromSymbol(host_dst, g_tc_attn_trace, (size_t)count * sizeof(unsigned long long)),
                               "tc_debug_read_attn_trace");
#else
    (void)host_dst;
    (void)count;
    return tc_host::fail(TC_ERR_INVALID, "tc_debug_read_attn_trace: library built without TC_BUILD_TRACE=1");
#endif
}

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
    {
        const char* pp_env = getenv("TC_ATTN_PP");   // "1" enables the exponential-phase token (A/B testing; measured no gain)
        p.pingpong = (pp_env && pp_env[0] == '1');
        const char* sg_env = getenv("TC_ATTN_STAGGER");   // cycles (A/B testing)
        p.stagger = sg_env ? atoi(sg_env) : 0;
    }
    {
        // one query tile per CTA (two independent CTAs per SM) for the UNet's self attentions: L = 2560 456 -> 423 us, L = 640
        // 103 -> 85 us in one A/B call; the long-K/V fusion attention (20480 keys) is 4 % faster with two tiles per CTA
        // sharing each K/V stage (profiles/r02_attn_single_ab.txt)
        const char* sg = getenv("TC_ATTN_SINGLE");       // "0" / "1" force (A/B testing)
        p.single = sg ? (sg[0] == '1') : (d->Lk[0] <= 4096);
    }
    const size_t smem_full = (size_t)(2 + 2 * kStages) * kTileBytes + 1024 + 512 + 4096;
    const size_t smem_bytes = p.single ? (size_t)(1 + 2 * 2) * kTileBytes + 1024 + 512 + 1024 : smem_full;
    const int rows_per_cta = p.single ? kQTile : 2 * kQTile;
    dim3 grid((d->Lq + rows_per_cta - 1) / rows_per_cta, d->heads, d->q_batches);
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

// Cross attention over one or two short K/V segments (text + image tokens) through the resident-K/V mma.sync kernel.
// Returns TC_ERR_INVALID (without setting an error) when the shape does not fit: the caller falls back.
int tc_attention_xs(const TcAttention* d, cudaStream_t stream) {
    const int nk0 = (d->Lk[0] + 15) & ~15, nk1 = d->n_seg > 1 ? ((d->Lk[1] + 15) & ~15) : 0;
    if (nk0 + nk1 > 8 * kXsMaxNT) return TC_ERR_INVALID;
    if ((d->ldq % 8) || (d->ldo % 8) || ((reinterpret_cast<uintptr_t>(d->q) | reinterpret_cast<uintptr_t>(d->out)) & 15)) return TC_ERR_INVALID;   // 16-byte row vectors
    AttnXsParams p;
    memset(&p, 0, sizeof(p));
    p.q = reinterpret_cast<const __half*>(d->q);
    p.ldq = d->ldq;
    for (int s = 0; s < d->n_seg; ++s) {
        p.k[s] = reinterpret_cast<const __half*>(d->k[s]);
        p.v[s] = reinterpret_cast<const __half*>(d->v[s]);
        p.ldk[s] = d->ldk[s];
        p.ldv[s] = d->ldv[s];
        p.Lk[s] = d->Lk[s];
        p.kv_div[s] = d->kv_div[s];
    }
    p.Lq = d->Lq;
    p.heads = d->heads;
    p.n_seg = d->n_seg;
    p.out = reinterpret_cast<__half*>(d->out);
    p.ldo = d->ldo;
    p.scale_log2 = d->scale * 1.4426950408889634f;
    // ~4 waves of CTAs (3 resident per SM); each CTA streams its share of the 16-query groups of one (frame, head)
    const long long pairs = (long long)d->heads * d->q_batches;
    const int n_groups = (d->Lq + 15) / 16;
    long long split = (12LL * sm_count() + pairs - 1) / pairs;
    const long long max_split = (n_groups + 3) / 4;
    if (split > max_split) split = max_split;
    if (split < 1) split = 1;
    dim3 grid((unsigned)split, d->heads, d->q_batches);
    tc_host::launch(tc_attn_xs_kernel, grid, dim3(kXsThreads), 0, stream, 1, p);
    count_launch();
    TC_CHECK_LAUNCH("tc_attn_xs_kernel");
    return TC_OK;
}

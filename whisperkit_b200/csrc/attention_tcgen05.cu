// K4: encoder self-attention (non-causal, head dim 64, S = 1500) on tcgen05 tensor cores with the probabilities kept in tensor memory.
//
// One CTA = one (window, head, 256-query block) = two 128-query tiles A and B that share every K / V tile.  384 threads:
//   warp 0        TMA producer: Q_A, Q_B once, then K / V tiles of 128 keys through 4-deep smem rings (128B swizzle)
//   warp 1        MMA issuer, per key tile j and query tile t:
//                     S_t  = Q_t K_j^T          tcgen05.mma 128x128x16 x4, A and B from smem, accumulator S_t in TMEM
//                     O_t += P_t V_j            tcgen05.mma 128x64x16  x8, A = P_t FROM TMEM (16-bit, written by the softmax warps
//                                               into its own 64 columns), B = V_j MN-major from smem, accumulator O_t in TMEM
//                 S_t(j+1) is issued as soon as the softmax warps have READ S_t(j) into registers (s_free), i.e. it runs under the
//                 exponentials of tile j: the only serial chain left per query tile is load-S / max / exp / store-P
//   warps 4..7    softmax of tile A, ONE thread per query row (warp w owns TMEM lane quarter w & 3)
//   warps 8..11   softmax of tile B
// (warps 2, 3 only fill warpgroup 0: setmaxnreg moves registers per warpgroup - 72 for warpgroup 0, 216 for the softmax warpgroups, whose
// threads hold a whole 128-score row)
// Compared with the round-1 kernel (one query tile per CTA, two threads per row, P through shared memory, O accumulated in registers) this
// removes the per-tile row-max exchange through shared memory and its named barrier, the 32 KiB generic-proxy P store + proxy fence per
// tile, the per-tile O read-back, and half of the K / V traffic from L2: 1.71 -> 1.11 ms per layer at 64 windows (432 -> 662 TFLOP/s).
// A clock64 timeline of one CTA (profiles/r02_attention_trace.txt) shows the softmax warps never waiting any more (S ready in ~170
// clocks, P buffer free in ~90): a key tile costs a softmax warp ~2700 clocks = load S 200 + max 450 + 128 exponentials / pack / sum 1600,
// two such warps per scheduler, i.e. the kernel is bound by the per-element instruction work of the softmax (MUFU 16/clk/SM, FMA and ALU
// pipes, issue slots), not by the tensor pipe (30 %).  Measured and NOT kept (profiles/r02_attention_variants.txt): a turn barrier that
// hands the MUFU pipe from the A warp to the B warp of a scheduler (1.13 ms against 1.11 - the non-MUFU work serialises too); two
// threads per row with 16 softmax warps and one OR-reducing named barrier per tile (1.15 - 1.28 ms: same total instruction work).
//
// exp2 on two pipes: 3 of every 8 groups of four scores take a Cody-Waite + degree-3 polynomial on the FMA / ALU pipes (relative error
// 7.5e-5, far below the 16-bit rounding P gets next) instead of MUFU.EX2 (1.20 -> 1.11 ms).
//
// Online softmax with a lazy rescale: O_t and the row sum are relative to a reference maximum m_ref that only moves when some row
// of the warp finds a score more than 2^8 above it (then the owning warp rescales its 32 rows of O_t in TMEM itself, after pv_done says
// that P_t V_(j-1) has completed; P_t V_j cannot start before this warp publishes P_t(j), so O_t is quiescent meanwhile).  Probabilities
// are therefore <= 2^8, exact in f32 sums and inside bf16 / f16 range.
// Q/K/V are read in place from the packed qkv activation [B*T, 3*d_model] through one 3-D tensor map (per-window out-of-bounds rows are
// zero-filled by TMA; keys >= T are masked to -inf before the softmax).
// Reference counterpart: inside AudioEncoder.mlmodelc (Sources/WhisperKit/Core/AudioEncoder.swift:59-62).
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace wk {

static constexpr int kFaThreads = 384;     // warpgroup 0: TMA warp, MMA warp, two idle warps; warpgroups 1 and 2: softmax of tile A / B
static constexpr int kFaPolyOf8 = 3;       // of every 8 groups of four scores, how many take the polynomial exp2 (measured: 0 -> 1.20 ms, 3 -> 1.11 ms)
static constexpr int kFaBM = 128;          // queries per tile (two tiles per CTA)
static constexpr int kFaBN = 128;          // keys per tile
static constexpr int kFaD = 64;
static constexpr int kFaTile = kFaBN * kFaD * 2;   // 16 KiB: one K or V or Q tile, 128-byte rows
static constexpr int kFaStages = 4;
static constexpr int kFaSmem = 2 * kFaTile /*Q_A, Q_B*/ + 2 * kFaStages * kFaTile /*K, V rings*/ + 1024 /*align*/ + 512 /*barriers*/;
static constexpr int kFaTmemCols = 512;    // S_A @0, S_B @128, P_A @256, P_B @320 (16-bit: 64 columns each), O_A @384, O_B @448
static constexpr float kFaRescaleLog2 = 8.f;

struct FaParams {
    int T, H, dm, n_kv_tiles;
    float scale_log2e;
    uint32_t idesc_qk, idesc_pv;
};

__device__ __forceinline__ float fa_ex2(float x) {   // MUFU.EX2: 2^x, ex2(-inf) = +0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 2^x on the FMA / ALU pipes for a pair of scores (x <= 8): round-to-nearest split x = n + f with the 1.5 * 2^23 trick (the low mantissa
// bits of t hold n), degree-3 minimax polynomial of 2^f on [-0.5, 0.5] (relative error 7.5e-5, far below the 16-bit rounding P gets
// next), then n goes into the exponent field with one integer shift-add.  The MUFU pipe (16 results per clock and SM) bounds this
// kernel, so a fixed share of every row goes this way (kPolyOf8 of every 8 groups of four scores).
__device__ __forceinline__ float2 fa_ex2_poly2(float2 x) {
    x.x = fmaxf(x.x, -126.f);
    x.y = fmaxf(x.y, -126.f);
    const float2 t = add2(x, make_float2(12582912.f, 12582912.f));
    const float2 nf = add2(t, make_float2(-12582912.f, -12582912.f));
    const float2 f = fma2(nf, make_float2(-1.f, -1.f), x);
    float2 r = fma2(f, make_float2(0.05517146f, 0.05517146f), make_float2(0.24261086f, 0.24261086f));
    r = fma2(r, f, make_float2(0.69326097f, 0.69326097f));
    r = fma2(r, f, make_float2(0.9999281f, 0.9999281f));
    r.x = __int_as_float(__float_as_int(r.x) + (__float_as_int(t.x) << 23));
    r.y = __int_as_float(__float_as_int(r.y) + (__float_as_int(t.y) << 23));
    return r;
}

__device__ __forceinline__ uint64_t fa_sw128_desc(uint32_t smem_addr, uint32_t lbo16, uint32_t sbo16) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(lbo16 & 0x3FFF) << 16;
    d |= (uint64_t)(sbo16 & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// (the tcgen05 instructions of the MMA warp are predicated on the elect.sync lane inside the asm: see tc_mma_f16_elect in common.cuh)
// D[tmem] (+)= A[tmem] * B[smem desc]: A is read from tensor memory (lane = row, two consecutive 16-bit K elements per 32-bit column)
__device__ __forceinline__ void fa_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 columns of 32-bit: thread i of the warp writes lane (base+i), columns [c, c+32)
__device__ __forceinline__ void fa_tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}
// bounded mbarrier wait without the printf of mbar_wait_bounded (register-light: it sits in the softmax loop): a protocol bug ends as a
// trapped launch, not as a hung GPU
__device__ __forceinline__ void fa_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const unsigned long long t0 = globaltimer_ns();
    unsigned int spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0xffffu) == 0 && globaltimer_ns() - t0 > kSpinLimitNs) __trap();
    }
}
__device__ __forceinline__ void fa_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <typename T, int kPolyOf8>
__global__ void __launch_bounds__(kFaThreads, 1)
encoder_attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_qkv, T* __restrict__ out, const FaParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                                // 2 tiles
    uint8_t* sK = sQ + 2 * kFaTile;
    uint8_t* sV = sK + kFaStages * kFaTile;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kFaStages * kFaTile);
    uint64_t* q_full = bars;                      // 1
    uint64_t* k_full = bars + 1;                  // kFaStages
    uint64_t* k_empty = k_full + kFaStages;
    uint64_t* v_full = k_empty + kFaStages;
    uint64_t* v_empty = v_full + kFaStages;
    uint64_t* s_full = v_empty + kFaStages;       // 2 (per query tile): S_t(j) complete
    uint64_t* s_free = s_full + 2;                // 2: the softmax warps hold S_t(j) in registers, S_t may be overwritten
    uint64_t* p_full = s_free + 2;                // 2: P_t(j) is in TMEM (and O_t rescaled if it had to be)
    uint64_t* pv_done = p_full + 2;               // 2: P_t V_j complete: P_t may be overwritten, O_t may be rescaled / read
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.y;
    const int b = bh / p.H, h = bh % p.H;
    const int q0 = blockIdx.x * (2 * kFaBM);
    const int n = p.n_kv_tiles;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_qkv);
        mbar_init(q_full, 1);
        for (int i = 0; i < kFaStages; ++i) {
            mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 4); mbar_init(&p_full[i], 4); mbar_init(&pv_done[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, kFaTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    if (warp == 0) {
        // ============================ TMA producer ============================
        if (lane == 0) {
            mbar_expect_tx(q_full, 2 * kFaTile);
            tma_load_3d(sQ, &tm_qkv, q_full, h * kFaD, q0, b);
            tma_load_3d(sQ + kFaTile, &tm_qkv, q_full, h * kFaD, q0 + kFaBM, b);
            for (int j = 0; j < n; ++j) {
                const int st = j % kFaStages;
                const uint32_t ph = (j / kFaStages) & 1;
                fa_wait(&k_empty[st], ph ^ 1);
                mbar_expect_tx(&k_full[st], kFaTile);
                tma_load_3d(sK + st * kFaTile, &tm_qkv, &k_full[st], p.dm + h * kFaD, j * kFaBN, b);
                fa_wait(&v_empty[st], ph ^ 1);
                mbar_expect_tx(&v_full[st], kFaTile);
                tma_load_3d(sV + st * kFaTile, &tm_qkv, &v_full[st], 2 * p.dm + h * kFaD, j * kFaBN, b);
            }
        }
    } else if (warp == 1) {
        // ============================ MMA issuer ============================
        // every lane runs this code converged; only the tcgen05 instructions are predicated, on the lane elect.sync picks (see tc_mma_f16_elect)
        const uint64_t dbase = ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);   // see fa_sw128_desc(.., 1, 64)
        const uint64_t qd0 = dbase | (uint64_t)((smem_u32(sQ) >> 4) & 0x3FFF);
        const uint64_t kd0 = dbase | (uint64_t)((smem_u32(sK) >> 4) & 0x3FFF);
        const uint64_t vd0 = dbase | (uint64_t)((smem_u32(sV) >> 4) & 0x3FFF);
        constexpr uint64_t kTileStep = kFaTile >> 4;   // descriptor start-address units (16 B) per 16 KiB tile; all tiles sit below 256 KiB
        const uint32_t idesc_qk = p.idesc_qk, idesc_pv = p.idesc_pv;
        auto mma_s = [&](int t, int st) {   // S_t = Q_t K^T
            const uint64_t ad = qd0 + (uint64_t)t * kTileStep, bd = kd0 + (uint64_t)st * kTileStep;
            const uint32_t d_addr = tmem + t * 128;
#pragma unroll
            for (int k = 0; k < kFaD / 16; ++k) tc_mma_f16_elect(d_addr, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc_qk, k > 0 ? 1u : 0u);
            tc_commit_elect(&s_full[t]);
            if (t == 1) tc_commit_elect(&k_empty[st]);
        };
        auto mma_pv = [&](int t, int st, uint32_t par, bool first) {   // O_t (+)= P_t V
            fa_wait(&p_full[t], par);
            __syncwarp();
            tc_fence_after();
            const uint64_t bd = vd0 + (uint64_t)st * kTileStep;
            const uint32_t d_addr = tmem + 384 + t * 64, a_addr = tmem + 256 + t * 64;
            // A = P_t: 16 keys = 8 TMEM columns per step.  B = V: MN-major (d contiguous, 128-byte rows), 16 keys = 16 rows = 2 KiB
            fa_mma_ts(d_addr, a_addr, bd, idesc_pv, first ? 0u : 1u);
#pragma unroll
            for (int k = 1; k < kFaBN / 16; ++k) fa_mma_ts(d_addr, a_addr + k * 8, bd + (uint64_t)(k * (2048 >> 4)), idesc_pv, 1u);
            tc_commit_elect(&pv_done[t]);
            if (t == 1) tc_commit_elect(&v_empty[st]);
        };
        fa_wait(q_full, 0);
        fa_wait(&k_full[0], 0);
        __syncwarp();
        tc_fence_after();
        mma_s(0, 0);
        mma_s(1, 0);
        int st = 0, st_next = 1;
        uint32_t ring_par = 0, ring_par_next = 0;   // parity of the ring slot st (st_next) for this (the next) key tile
        for (int j = 0; j < n; ++j) {
            if (j + 1 < n) {   // S(j+1) of both tiles as soon as their S(j) has been read: runs under the exponentials of tile j
                fa_wait(&k_full[st_next], ring_par_next);
                fa_wait(&s_free[0], j & 1);
                __syncwarp();
                tc_fence_after();
                mma_s(0, st_next);
                fa_wait(&s_free[1], j & 1);
                __syncwarp();
                tc_fence_after();
                mma_s(1, st_next);
            }
            fa_wait(&v_full[st], ring_par);
            mma_pv(0, st, j & 1, j == 0);
            mma_pv(1, st, j & 1, j == 0);
            st = st_next; ring_par = ring_par_next;
            if (++st_next == kFaStages) { st_next = 0; ring_par_next ^= 1; }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
        // ============================ softmax: one thread per query row ============================
        const int t = (warp - 4) >> 2;                // query tile
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may touch
        const int row = quarter * 32 + lane;          // query row inside the tile == TMEM lane
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const uint32_t s_addr = tmem + lane_addr + t * 128;
        const uint32_t p_addr = tmem + lane_addr + 256 + t * 64;
        const uint32_t o_addr = tmem + lane_addr + 384 + t * 64;
        const float c = p.scale_log2e;
        const float2 c2 = make_float2(c, c);
        float m_ref = -INFINITY;      // the maximum O_t and l are relative to
        float2 l2a = make_float2(0.f, 0.f), l2b = make_float2(0.f, 0.f);

        for (int j = 0; j < n; ++j) {
            fa_wait(&s_full[t], j & 1);
            tc_fence_after();
            uint32_t sr[4][32];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) tmem_ld_32x32(s_addr + ch * 32, sr[ch]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[t]);   // S_t(j) is in registers: the MMA warp may start S_t(j+1)
            const int valid = p.T - j * kFaBN;   // keys of this tile that exist (>= 128 except at the end)
            if (valid < kFaBN) {
#pragma unroll
                for (int ch = 0; ch < 4; ++ch)
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (ch * 32 + i >= valid) sr[ch][i] = 0xff800000u;   // -inf: key does not exist
            }
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    mx0 = fmaxf(mx0, __uint_as_float(sr[ch][i]));
                    mx1 = fmaxf(mx1, __uint_as_float(sr[ch][i + 1]));
                    mx2 = fmaxf(mx2, __uint_as_float(sr[ch][i + 2]));
                    mx3 = fmaxf(mx3, __uint_as_float(sr[ch][i + 3]));
                }
            const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
            if (j == 0) {
                m_ref = mx;       // tile 0 always holds real keys, so mx is finite
            } else if (__any_sync(0xffffffffu, (mx - m_ref) * c > kFaRescaleLog2)) {
                // some row of this warp outgrew its reference: move every row of the warp to its current maximum and rescale O_t, l
                const float m_new = fmaxf(m_ref, mx);
                const float corr = fa_ex2((m_ref - m_new) * c);
                m_ref = m_new;
                l2a.x *= corr; l2a.y *= corr; l2b.x *= corr; l2b.y *= corr;
                fa_wait(&pv_done[t], (j - 1) & 1);   // P_t V_(j-1) has landed in O_t
                tc_fence_after();
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t orr[32];
                    tmem_ld_32x32(o_addr + half * 32, orr);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * corr);
                    fa_tmem_st_32x32(o_addr + half * 32, orr);
                }
            }
            const float nmsc = -m_ref * c;
            const float2 nm2 = make_float2(nmsc, nmsc);
            // P_t(j): 128 probabilities per row -> 64 packed 16-bit columns
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t pk[32];
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int ch = half * 2 + cc;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float2 x0 = fma2(make_float2(__uint_as_float(sr[ch][i]), __uint_as_float(sr[ch][i + 1])), c2, nm2);
                        const float2 x1 = fma2(make_float2(__uint_as_float(sr[ch][i + 2]), __uint_as_float(sr[ch][i + 3])), c2, nm2);
                        const int g = i >> 2;   // group of four scores inside the 32-score chunk
                        const bool poly = ((g + 1) * kPolyOf8) / 8 != (g * kPolyOf8) / 8;
                        const float2 e0 = poly ? fa_ex2_poly2(x0) : make_float2(fa_ex2(x0.x), fa_ex2(x0.y));
                        const float2 e1 = poly ? fa_ex2_poly2(x1) : make_float2(fa_ex2(x1.x), fa_ex2(x1.y));
                        l2a = add2(l2a, e0);
                        l2b = add2(l2b, e1);
                        pk[cc * 16 + (i >> 1)] = T16<T>::pack2(e0.x, e0.y);
                        pk[cc * 16 + (i >> 1) + 1] = T16<T>::pack2(e1.x, e1.y);
                    }
                }
                if (half == 0 && j > 0) {   // the P buffer is free once P_t V_(j-1) has completed (long ago, normally)
                    fa_wait(&pv_done[t], (j - 1) & 1);
                    tc_fence_after();
                }
                fa_tmem_st_32x32(p_addr + half * 32, pk);
            }
            fa_tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[t]);
        }
        // ---- epilogue: O_t / l, 64 columns = one 128-byte row segment per thread
        fa_wait(&pv_done[t], (n - 1) & 1);
        tc_fence_after();
        const float inv = 1.f / ((l2a.x + l2a.y) + (l2b.x + l2b.y));
        const int q = q0 + t * kFaBM + row;
        uint4* dst = reinterpret_cast<uint4*>(out + ((long long)b * p.T + q) * p.dm + h * kFaD);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t orr[32];
            tmem_ld_32x32(o_addr + half * 32, orr);
            tmem_ld_wait();
            if (q < p.T) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint4 v;
                    v.x = T16<T>::pack2(__uint_as_float(orr[8 * i]) * inv, __uint_as_float(orr[8 * i + 1]) * inv);
                    v.y = T16<T>::pack2(__uint_as_float(orr[8 * i + 2]) * inv, __uint_as_float(orr[8 * i + 3]) * inv);
                    v.z = T16<T>::pack2(__uint_as_float(orr[8 * i + 4]) * inv, __uint_as_float(orr[8 * i + 5]) * inv);
                    v.w = T16<T>::pack2(__uint_as_float(orr[8 * i + 6]) * inv, __uint_as_float(orr[8 * i + 7]) * inv);
                    dst[half * 4 + i] = v;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, kFaTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

wk_status encoder_attention_tcgen05(const void* qkv, void* out, int B, int T, int n_heads, int dtype, cudaStream_t stream) {
    static PFN_encodeTiled enc = nullptr;
    if (!enc) {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled entry point unavailable");
            return WK_ERR_CUDA;
        }
        enc = reinterpret_cast<PFN_encodeTiled>(fp);
    }
    const int dm = n_heads * 64;
    CUtensorMap tm;
    cuuint64_t gdim[3] = {(cuuint64_t)3 * dm, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)3 * dm * 2, (cuuint64_t)T * 3 * dm * 2};
    cuuint32_t box[3] = {64, 128, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tm, dtype == WK_DTYPE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                     const_cast<void*>(qkv), gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("attention tensor map encode failed: %d", (int)r); return WK_ERR_CUDA; }
    FaParams p;
    p.T = T; p.H = n_heads; p.dm = dm;
    p.n_kv_tiles = (T + kFaBN - 1) / kFaBN;
    p.scale_log2e = 0.125f * 1.4426950408889634f;
    const uint32_t fmt = dtype == WK_DTYPE_F16 ? 0u : 1u;
    p.idesc_qk = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(kFaBN >> 3) << 17) | ((uint32_t)(kFaBM >> 4) << 24);
    p.idesc_pv = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 16) /* B is MN-major */ | ((uint32_t)(kFaD >> 3) << 17) |
                 ((uint32_t)(kFaBM >> 4) << 24);
    dim3 grid((T + 2 * kFaBM - 1) / (2 * kFaBM), B * n_heads);
    cudaError_t e = cudaSuccess;
    if (dtype == WK_DTYPE_F16) {
        static bool set = false;
        if (!set) { e = cudaFuncSetAttribute(encoder_attention_tcgen05_kernel<__half, kFaPolyOf8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFaSmem); set = e == cudaSuccess; }
        if (e == cudaSuccess) encoder_attention_tcgen05_kernel<__half, kFaPolyOf8><<<grid, kFaThreads, kFaSmem, stream>>>(tm, (__half*)out, p);
    } else {
        static bool set = false;
        if (!set) { e = cudaFuncSetAttribute(encoder_attention_tcgen05_kernel<__nv_bfloat16, kFaPolyOf8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFaSmem); set = e == cudaSuccess; }
        if (e == cudaSuccess) encoder_attention_tcgen05_kernel<__nv_bfloat16, kFaPolyOf8><<<grid, kFaThreads, kFaSmem, stream>>>(tm, (__nv_bfloat16*)out, p);
    }
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(fa): %s", cudaGetErrorString(e)); return WK_ERR_CUDA; }
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("encoder_attention_tcgen05 launch: %s", cudaGetErrorString(e)); return WK_ERR_CUDA; }
    return WK_OK;
}

}  // namespace wk

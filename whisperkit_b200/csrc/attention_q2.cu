// K4 (second generation): encoder self-attention (non-causal, head dim 64, S = 1500) on tcgen05 tensor cores with the
// probabilities kept in tensor memory.
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
// (warps 2, 3 only fill warpgroup 0: setmaxnreg moves registers per warpgroup - 24 for warpgroup 0, 240 for the softmax warpgroups, whose
// threads hold a whole 128-score row)
// The two query tiles ping-pong: while the softmax warps of A turn S_A into P_A, the tensor core runs S_B / P_B V; a scheduler holds
// one A warp and one B warp, so the MUFU pipe always has work.  Compared with the first-generation kernel (one query tile per CTA,
// two threads per row, P through shared memory, O accumulated in registers) this removes the per-tile row-max exchange through
// shared memory and its named barrier, the 32 KiB generic-proxy P store + proxy fence per tile, the per-tile O read-back, and half
// of the K / V traffic from L2.
//
// Online softmax with a lazy rescale: O_t and the row sum are relative to a reference maximum m_ref that only moves when some row
// of the warp finds a score more than 2^8 above it (then the owning warp rescales its 32 rows of O_t in TMEM itself, after pv_done says
// that P_t V_(j-1) has completed; P_t V_j cannot start before this warp publishes P_t(j), so O_t is quiescent meanwhile).  Probabilities
// are therefore <= 2^8, exact in f32 sums and inside bf16 / f16 range.
// Reference counterpart: inside AudioEncoder.mlmodelc (Sources/WhisperKit/Core/AudioEncoder.swift:59-62).
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "kernels.h"

namespace wk {

static constexpr int kQ2Threads = 384;   // (kSplit = 1) warpgroup 0: TMA warp, MMA warp, two idle warps; warpgroups 1 and 2: softmax of tile A / B
static constexpr int kQ2BM = 128;          // queries per tile (two tiles per CTA)
static constexpr int kQ2BN = 128;          // keys per tile
static constexpr int kQ2D = 64;
static constexpr int kQ2Tile = kQ2BN * kQ2D * 2;   // 16 KiB: one K or V or Q tile, 128-byte rows
static constexpr int kQ2Stages = 4;
static constexpr int kQ2Smem = 2 * kQ2Tile /*Q_A, Q_B*/ + 2 * kQ2Stages * kQ2Tile /*K, V rings*/ + 1024 /*align*/ + 512 /*barriers*/ + 2048 /*exchange*/;
static constexpr int kQ2TmemCols = 512;    // S_A @0, S_B @128, P_A @256, P_B @320 (16-bit: 64 columns each), O_A @384, O_B @448
static constexpr float kQ2RescaleLog2 = 8.f;

struct Q2Params {
    int T, H, dm, n_kv_tiles;
    float scale_log2e;
    uint32_t idesc_qk, idesc_pv;
};

__device__ __forceinline__ float q2_ex2(float x) {   // MUFU.EX2: 2^x, ex2(-inf) = +0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// packed f32x2 arithmetic (FFMA2 / FADD2): halves the FMA-pipe issue slots of the softmax inner loop
__device__ __forceinline__ float2 q2_fma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 q2_add2(float2 a, float2 b) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "add.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}

// 2^x on the FMA / ALU pipes for a pair of scores (x <= 8): round-to-nearest split x = n + f with the 1.5 * 2^23 trick (the low mantissa
// bits of t hold n), degree-3 minimax polynomial of 2^f on [-0.5, 0.5] (relative error 7.5e-5, far below the 16-bit rounding P gets
// next), then n goes into the exponent field with one integer shift-add.  The MUFU pipe (16 results per clock and SM) bounds this
// kernel, so a fixed share of every row goes this way (kPolyOf8 of every 8 groups of four scores).
__device__ __forceinline__ float2 q2_ex2_poly2(float2 x) {
    x.x = fmaxf(x.x, -126.f);
    x.y = fmaxf(x.y, -126.f);
    const float2 t = q2_add2(x, make_float2(12582912.f, 12582912.f));
    const float2 nf = q2_add2(t, make_float2(-12582912.f, -12582912.f));
    const float2 f = q2_fma2(nf, make_float2(-1.f, -1.f), x);
    float2 r = q2_fma2(f, make_float2(0.05517146f, 0.05517146f), make_float2(0.24261086f, 0.24261086f));
    r = q2_fma2(r, f, make_float2(0.69326097f, 0.69326097f));
    r = q2_fma2(r, f, make_float2(0.9999281f, 0.9999281f));
    r.x = __int_as_float(__float_as_int(r.x) + (__float_as_int(t.x) << 23));
    r.y = __int_as_float(__float_as_int(r.y) + (__float_as_int(t.y) << 23));
    return r;
}

__device__ __forceinline__ uint64_t q2_sw128_desc(uint32_t smem_addr, uint32_t lbo16, uint32_t sbo16) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(lbo16 & 0x3FFF) << 16;
    d |= (uint64_t)(sbo16 & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// The tcgen05 instructions of the MMA warp are PREDICATED inside the asm on the lane `elect.sync` picks, instead of sitting in an
// `if (lane == 0)` branch: UTCHMMA / UTCBAR take uniform-register operands, and inside a divergent branch (or under an ordinary per-thread
// predicate) the compiler rebuilds every operand with an ELECT / R2UR.BROADCAST / BRA.U.ANY loop - ~125 clocks per MMA issued; that loop,
// not the tensor pipe, bounded the first version of this kernel.  With elect.sync the issue path is straight-line code.  The warp must be
// converged at every call (the callers __syncwarp() after their mbarrier spins).
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void q2_mma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: A is read from tensor memory (lane = row, two consecutive 16-bit K elements per 32-bit column)
__device__ __forceinline__ void q2_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void q2_commit(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
        : "memory");
}
// 32 lanes x 32 columns of 32-bit: thread i of the warp writes lane (base+i), columns [c, c+32)
__device__ __forceinline__ void q2_tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
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
__device__ __forceinline__ void q2_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const unsigned long long t0 = globaltimer_ns();
    unsigned int spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0xffffu) == 0 && globaltimer_ns() - t0 > kSpinLimitNs) __trap();
    }
}
__device__ __forceinline__ void q2_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// kSplit = softmax threads per query row: 1 (a thread owns a whole 128-score row, 8 softmax warps) or 2 (each thread owns 64 of the
// tile's keys and 32 of the output columns, 16 softmax warps - four per scheduler instead of two, which is what hides the MUFU / TMEM
// latencies of the serial load-max-exp-store chain; the two threads of a row only meet in one OR-reducing named barrier per tile)
template <typename T, int kPolyOf8, bool kOrdered, int kSplit>
__global__ void __launch_bounds__(128 + 256 * kSplit, 1)
encoder_attention_q2_kernel(const __grid_constant__ CUtensorMap tm_qkv, T* __restrict__ out, const Q2Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                                // 2 tiles
    uint8_t* sK = sQ + 2 * kQ2Tile;
    uint8_t* sV = sK + kQ2Stages * kQ2Tile;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kQ2Stages * kQ2Tile);
    uint64_t* q_full = bars;                      // 1
    uint64_t* k_full = bars + 1;                  // kQ2Stages
    uint64_t* k_empty = k_full + kQ2Stages;
    uint64_t* v_full = k_empty + kQ2Stages;
    uint64_t* v_empty = v_full + kQ2Stages;
    uint64_t* s_full = v_empty + kQ2Stages;       // 2 (per query tile): S_t(j) complete
    uint64_t* s_free = s_full + 2;                // 2: the softmax warps hold S_t(j) in registers, S_t may be overwritten
    uint64_t* p_full = s_free + 2;                // 2: P_t(j) is in TMEM (and O_t rescaled if it had to be)
    uint64_t* pv_done = p_full + 2;               // 2: P_t V_j complete: P_t may be overwritten, O_t may be rescaled / read
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);
    float* zero_slot = reinterpret_cast<float*>(tmem_slot + 1);   // holds 0.0f: see the turn barrier in the softmax loop
    float* xch = reinterpret_cast<float*>(bars + 64);             // kSplit = 2: [tile][half][128 rows] row-max / row-sum exchange (2 KiB)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.y;
    const int b = bh / p.H, h = bh % p.H;
    const int q0 = blockIdx.x * (2 * kQ2BM);
    const int n = p.n_kv_tiles;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_qkv);
        *zero_slot = 0.f;
        mbar_init(q_full, 1);
        for (int i = 0; i < kQ2Stages; ++i) {
            mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 4 * kSplit); mbar_init(&p_full[i], 4 * kSplit); mbar_init(&pv_done[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, kQ2TmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < 4) {
        if constexpr (kSplit == 1) asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == 0) {
        // ============================ TMA producer ============================
        if (lane == 0) {
            mbar_expect_tx(q_full, 2 * kQ2Tile);
            tma_load_3d(sQ, &tm_qkv, q_full, h * kQ2D, q0, b);
            tma_load_3d(sQ + kQ2Tile, &tm_qkv, q_full, h * kQ2D, q0 + kQ2BM, b);
            for (int j = 0; j < n; ++j) {
                const int st = j % kQ2Stages;
                const uint32_t ph = (j / kQ2Stages) & 1;
                q2_wait(&k_empty[st], ph ^ 1);
                mbar_expect_tx(&k_full[st], kQ2Tile);
                tma_load_3d(sK + st * kQ2Tile, &tm_qkv, &k_full[st], p.dm + h * kQ2D, j * kQ2BN, b);
                q2_wait(&v_empty[st], ph ^ 1);
                mbar_expect_tx(&v_full[st], kQ2Tile);
                tma_load_3d(sV + st * kQ2Tile, &tm_qkv, &v_full[st], 2 * p.dm + h * kQ2D, j * kQ2BN, b);
            }
        }
    } else if (warp == 1) {
        // ============================ MMA issuer ============================
        // every lane runs this code converged; only the tcgen05 instructions are predicated, on the lane elect.sync picks (see q2_mma_ss)
        const uint64_t dbase = ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);   // see q2_sw128_desc(.., 1, 64)
        const uint64_t qd0 = dbase | (uint64_t)((smem_u32(sQ) >> 4) & 0x3FFF);
        const uint64_t kd0 = dbase | (uint64_t)((smem_u32(sK) >> 4) & 0x3FFF);
        const uint64_t vd0 = dbase | (uint64_t)((smem_u32(sV) >> 4) & 0x3FFF);
        constexpr uint64_t kTileStep = kQ2Tile >> 4;   // descriptor start-address units (16 B) per 16 KiB tile; all tiles sit below 256 KiB
        const uint32_t idesc_qk = p.idesc_qk, idesc_pv = p.idesc_pv;
        auto mma_s = [&](int t, int st) {   // S_t = Q_t K^T
            const uint64_t ad = qd0 + (uint64_t)t * kTileStep, bd = kd0 + (uint64_t)st * kTileStep;
            const uint32_t d_addr = tmem + t * 128;
#pragma unroll
            for (int k = 0; k < kQ2D / 16; ++k) q2_mma_ss(d_addr, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc_qk, k > 0 ? 1u : 0u);
            q2_commit(&s_full[t]);
            if (t == 1) q2_commit(&k_empty[st]);
        };
        auto mma_pv = [&](int t, int st, uint32_t par, bool first) {   // O_t (+)= P_t V
            q2_wait(&p_full[t], par);
            __syncwarp();
            tc_fence_after();
            const uint64_t bd = vd0 + (uint64_t)st * kTileStep;
            const uint32_t d_addr = tmem + 384 + t * 64, a_addr = tmem + 256 + t * 64;
            // A = P_t: 16 keys = 8 TMEM columns per step.  B = V: MN-major (d contiguous, 128-byte rows), 16 keys = 16 rows = 2 KiB
            q2_mma_ts(d_addr, a_addr, bd, idesc_pv, first ? 0u : 1u);
#pragma unroll
            for (int k = 1; k < kQ2BN / 16; ++k) q2_mma_ts(d_addr, a_addr + k * 8, bd + (uint64_t)(k * (2048 >> 4)), idesc_pv, 1u);
            q2_commit(&pv_done[t]);
            if (t == 1) q2_commit(&v_empty[st]);
        };
        q2_wait(q_full, 0);
        q2_wait(&k_full[0], 0);
        __syncwarp();
        tc_fence_after();
        mma_s(0, 0);
        mma_s(1, 0);
        int st = 0, st_next = 1;
        uint32_t ring_par = 0, ring_par_next = 0;   // parity of the ring slot st (st_next) for this (the next) key tile
        for (int j = 0; j < n; ++j) {
            if (j + 1 < n) {   // S(j+1) of both tiles as soon as their S(j) has been read: runs under the exponentials of tile j
                q2_wait(&k_full[st_next], ring_par_next);
                q2_wait(&s_free[0], j & 1);
                __syncwarp();
                tc_fence_after();
                mma_s(0, st_next);
                q2_wait(&s_free[1], j & 1);
                __syncwarp();
                tc_fence_after();
                mma_s(1, st_next);
            }
            q2_wait(&v_full[st], ring_par);
            mma_pv(0, st, j & 1, j == 0);
            mma_pv(1, st, j & 1, j == 0);
            st = st_next; ring_par = ring_par_next;
            if (++st_next == kQ2Stages) { st_next = 0; ring_par_next ^= 1; }
        }
    }
    } else if constexpr (kSplit == 2) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");   // 640 threads are launched at 96 registers; warpgroup 0 gave up 40 each
        // ============================ softmax: two threads per query row ============================
        const int sw = warp - 4;
        const int t = sw >> 3;                        // query tile
        const int hf = (sw >> 2) & 1;                 // which 64 of the tile's 128 keys, which 32 of the 64 output columns
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may touch (the four warps of a row quarter share a scheduler)
        const int row = quarter * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const uint32_t s_addr = tmem + lane_addr + t * 128 + hf * 64;
        const uint32_t p_addr = tmem + lane_addr + 256 + t * 64 + hf * 32;
        const uint32_t o_addr = tmem + lane_addr + 384 + t * 64 + hf * 32;
        const uint32_t pair_bar = 1 + t * 4 + quarter;   // 64-thread named barrier of the two warps that share these 32 rows
        float* x_mine = xch + (t * 2 + hf) * 128 + row;
        const float* x_other = xch + (t * 2 + (hf ^ 1)) * 128 + row;
        const float c = p.scale_log2e;
        const float2 c2 = make_float2(c, c);
        float m_ref = -INFINITY;      // identical in both threads of a row
        float2 l2a = make_float2(0.f, 0.f), l2b = make_float2(0.f, 0.f);   // this thread's share of the row sum

        for (int j = 0; j < n; ++j) {
            q2_wait(&s_full[t], j & 1);
            tc_fence_after();
            uint32_t sr[2][32];
            tmem_ld_32x32(s_addr, sr[0]);
            tmem_ld_32x32(s_addr + 32, sr[1]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[t]);
            const int valid = p.T - j * kQ2BN - hf * 64;   // keys of this thread's half that exist
            if (valid < 64) {
#pragma unroll
                for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (ch * 32 + i >= valid) sr[ch][i] = 0xff800000u;
            }
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    mx0 = fmaxf(mx0, __uint_as_float(sr[ch][i]));
                    mx1 = fmaxf(mx1, __uint_as_float(sr[ch][i + 1]));
                    mx2 = fmaxf(mx2, __uint_as_float(sr[ch][i + 2]));
                    mx3 = fmaxf(mx3, __uint_as_float(sr[ch][i + 3]));
                }
            const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));   // -inf if this half holds no real key
            // one OR-reducing barrier per tile: does any of the 64 threads (32 rows x 2 halves) need the reference maximum moved?
            const uint32_t want = (j == 0 || (mx - m_ref) * c > kQ2RescaleLog2) ? 1u : 0u;
            uint32_t any;
            asm volatile(
                "{\n\t.reg .pred p, q;\n\t"
                "setp.ne.b32 p, %2, 0;\n\t"
                "bar.red.or.pred q, %1, 64, p;\n\t"
                "selp.u32 %0, 1, 0, q;\n\t}"
                : "=r"(any) : "r"(pair_bar), "r"(want) : "memory");
            if (any) {
                *x_mine = mx;
                asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
                const float row_mx = fmaxf(mx, *x_other);
                if (j == 0) {
                    m_ref = row_mx;           // tile 0 always holds real keys, so this is finite
                } else {
                    const float m_new = fmaxf(m_ref, row_mx);
                    const float corr = q2_ex2((m_ref - m_new) * c);
                    m_ref = m_new;
                    l2a.x *= corr; l2a.y *= corr; l2b.x *= corr; l2b.y *= corr;
                    q2_wait(&pv_done[t], (j - 1) & 1);   // P_t V_(j-1) has landed in O_t
                    tc_fence_after();
                    uint32_t orr[32];
                    tmem_ld_32x32(o_addr, orr);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * corr);
                    q2_tmem_st_32x32(o_addr, orr);
                }
            }
            const float nmsc = -m_ref * c;
            const float2 nm2 = make_float2(nmsc, nmsc);
            uint32_t pk[32];
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float2 x0 = q2_fma2(make_float2(__uint_as_float(sr[ch][i]), __uint_as_float(sr[ch][i + 1])), c2, nm2);
                    const float2 x1 = q2_fma2(make_float2(__uint_as_float(sr[ch][i + 2]), __uint_as_float(sr[ch][i + 3])), c2, nm2);
                    const int g = i >> 2;
                    const bool poly = ((g + 1) * kPolyOf8) / 8 != (g * kPolyOf8) / 8;
                    const float2 e0 = poly ? q2_ex2_poly2(x0) : make_float2(q2_ex2(x0.x), q2_ex2(x0.y));
                    const float2 e1 = poly ? q2_ex2_poly2(x1) : make_float2(q2_ex2(x1.x), q2_ex2(x1.y));
                    l2a = q2_add2(l2a, e0);
                    l2b = q2_add2(l2b, e1);
                    pk[ch * 16 + (i >> 1)] = T16<T>::pack2(e0.x, e0.y);
                    pk[ch * 16 + (i >> 1) + 1] = T16<T>::pack2(e1.x, e1.y);
                }
            }
            if (j > 0) {   // the P buffer is free once P_t V_(j-1) has completed (long ago, normally)
                q2_wait(&pv_done[t], (j - 1) & 1);
                tc_fence_after();
            }
            q2_tmem_st_32x32(p_addr, pk);
            q2_tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[t]);
        }
        // ---- epilogue: row sum = both halves; this thread normalises and stores 32 of the 64 output columns
        *x_mine = (l2a.x + l2a.y) + (l2b.x + l2b.y);
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        const float inv = 1.f / (*x_mine + *x_other);
        q2_wait(&pv_done[t], (n - 1) & 1);
        tc_fence_after();
        const int q = q0 + t * kQ2BM + row;
        uint4* dst = reinterpret_cast<uint4*>(out + ((long long)b * p.T + q) * p.dm + h * kQ2D + hf * 32);
        uint32_t orr[32];
        tmem_ld_32x32(o_addr, orr);
        tmem_ld_wait();
        if (q < p.T) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint4 v;
                v.x = T16<T>::pack2(__uint_as_float(orr[8 * i]) * inv, __uint_as_float(orr[8 * i + 1]) * inv);
                v.y = T16<T>::pack2(__uint_as_float(orr[8 * i + 2]) * inv, __uint_as_float(orr[8 * i + 3]) * inv);
                v.z = T16<T>::pack2(__uint_as_float(orr[8 * i + 4]) * inv, __uint_as_float(orr[8 * i + 5]) * inv);
                v.w = T16<T>::pack2(__uint_as_float(orr[8 * i + 6]) * inv, __uint_as_float(orr[8 * i + 7]) * inv);
                dst[i] = v;
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
        // kOrdered: the A warp and the B warp of a scheduler take turns at the exponentials (64-thread named barriers 1..4 "A may go",
        // 5..8 "B may go"): run together they would share the MUFU pipe and finish together, leaving the tensor core idle during the
        // softmax and the MUFU pipe idle during the MMAs; in turns, one tile's MMAs hide under the other tile's exponentials
        const uint32_t bar_own = 1 + t * 4 + quarter, bar_other = 1 + (t ^ 1) * 4 + quarter;
        if (kOrdered && t == 1) asm volatile("bar.arrive %0, 64;" ::"r"(bar_other) : "memory");
        float m_ref = -INFINITY;      // the maximum O_t and l are relative to
        float2 l2a = make_float2(0.f, 0.f), l2b = make_float2(0.f, 0.f);

        for (int j = 0; j < n; ++j) {
            q2_wait(&s_full[t], j & 1);
            tc_fence_after();
            uint32_t sr[4][32];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) tmem_ld_32x32(s_addr + ch * 32, sr[ch]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[t]);   // S_t(j) is in registers: the MMA warp may start S_t(j+1)
            const int valid = p.T - j * kQ2BN;   // keys of this tile that exist (>= 128 except at the end)
            if (valid < kQ2BN) {
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
            } else if (__any_sync(0xffffffffu, (mx - m_ref) * c > kQ2RescaleLog2)) {
                // some row of this warp outgrew its reference: move every row of the warp to its current maximum and rescale O_t, l
                const float m_new = fmaxf(m_ref, mx);
                const float corr = q2_ex2((m_ref - m_new) * c);
                m_ref = m_new;
                l2a.x *= corr; l2a.y *= corr; l2b.x *= corr; l2b.y *= corr;
                q2_wait(&pv_done[t], (j - 1) & 1);   // P_t V_(j-1) has landed in O_t
                tc_fence_after();
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t orr[32];
                    tmem_ld_32x32(o_addr + half * 32, orr);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * corr);
                    q2_tmem_st_32x32(o_addr + half * 32, orr);
                }
            }
            float nmsc = -m_ref * c;
            if (kOrdered) {
                // my turn at the MUFU pipe.  The exponentials are pure register arithmetic, which ptxas schedules across a barrier as it
                // likes; a volatile shared-memory load (of 0.0f) placed after the barrier and folded into their common operand pins them
                float z;
                asm volatile("bar.sync %1, 64;\n\tld.volatile.shared.f32 %0, [%2];" : "=f"(z) : "r"(bar_own), "r"(smem_u32(zero_slot)) : "memory");
                nmsc += z;
            }
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
                        const float2 x0 = q2_fma2(make_float2(__uint_as_float(sr[ch][i]), __uint_as_float(sr[ch][i + 1])), c2, nm2);
                        const float2 x1 = q2_fma2(make_float2(__uint_as_float(sr[ch][i + 2]), __uint_as_float(sr[ch][i + 3])), c2, nm2);
                        const int g = i >> 2;   // group of four scores inside the 32-score chunk
                        const bool poly = ((g + 1) * kPolyOf8) / 8 != (g * kPolyOf8) / 8;
                        const float2 e0 = poly ? q2_ex2_poly2(x0) : make_float2(q2_ex2(x0.x), q2_ex2(x0.y));
                        const float2 e1 = poly ? q2_ex2_poly2(x1) : make_float2(q2_ex2(x1.x), q2_ex2(x1.y));
                        l2a = q2_add2(l2a, e0);
                        l2b = q2_add2(l2b, e1);
                        pk[cc * 16 + (i >> 1)] = T16<T>::pack2(e0.x, e0.y);
                        pk[cc * 16 + (i >> 1) + 1] = T16<T>::pack2(e1.x, e1.y);
                    }
                }
                if (half == 0 && j > 0) {   // the P buffer is free once P_t V_(j-1) has completed (long ago, normally)
                    q2_wait(&pv_done[t], (j - 1) & 1);
                    tc_fence_after();
                }
                q2_tmem_st_32x32(p_addr + half * 32, pk);
            }
            if (kOrdered && !(t == 1 && j + 1 == n)) asm volatile("bar.arrive %0, 64;" ::"r"(bar_other) : "memory");
            q2_tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[t]);
        }
        // ---- epilogue: O_t / l, 64 columns = one 128-byte row segment per thread
        q2_wait(&pv_done[t], (n - 1) & 1);
        tc_fence_after();
        const float inv = 1.f / ((l2a.x + l2a.y) + (l2b.x + l2b.y));
        const int q = q0 + t * kQ2BM + row;
        uint4* dst = reinterpret_cast<uint4*>(out + ((long long)b * p.T + q) * p.dm + h * kQ2D);
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
        tmem_dealloc(tmem, kQ2TmemCols);
    }
}

// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiledQ2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

wk_status encoder_attention_q2(const void* qkv, void* out, int B, int T, int n_heads, int dtype, cudaStream_t stream) {
    static PFN_encodeTiledQ2 enc = nullptr;
    if (!enc) {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled entry point unavailable");
            return WK_ERR_CUDA;
        }
        enc = reinterpret_cast<PFN_encodeTiledQ2>(fp);
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
    Q2Params p;
    p.T = T; p.H = n_heads; p.dm = dm;
    p.n_kv_tiles = (T + kQ2BN - 1) / kQ2BN;
    p.scale_log2e = 0.125f * 1.4426950408889634f;
    const uint32_t fmt = dtype == WK_DTYPE_F16 ? 0u : 1u;
    p.idesc_qk = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(kQ2BN >> 3) << 17) | ((uint32_t)(kQ2BM >> 4) << 24);
    p.idesc_pv = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 16) /* B is MN-major */ | ((uint32_t)(kQ2D >> 3) << 17) |
                 ((uint32_t)(kQ2BM >> 4) << 24);
    dim3 grid((T + 2 * kQ2BM - 1) / (2 * kQ2BM), B * n_heads);
    const int poly = [] { const char* e = getenv("WKB200_ATTN_POLY"); return e ? atoi(e) : 0; }();      // bring-up switches (read per call)
    const bool ordered = [] { const char* e = getenv("WKB200_ATTN_ORDER"); return !(e && e[0] == '0'); }();
    cudaError_t e = cudaSuccess;
    const int split = [] { const char* e = getenv("WKB200_ATTN_SPLIT"); return e ? atoi(e) : 2; }();
    auto launch = [&](auto kern, auto* o, int threads) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kQ2Smem);
        if (e == cudaSuccess) kern<<<grid, threads, kQ2Smem, stream>>>(tm, o, p);
    };
    auto pick = [&](auto* o) {
        using OT = std::remove_pointer_t<decltype(o)>;
        if (split == 2) {
            switch (poly) {
                case 2: launch(encoder_attention_q2_kernel<OT, 2, false, 2>, o, 640); break;
                case 3: launch(encoder_attention_q2_kernel<OT, 3, false, 2>, o, 640); break;
                case 4: launch(encoder_attention_q2_kernel<OT, 4, false, 2>, o, 640); break;
                default: launch(encoder_attention_q2_kernel<OT, 0, false, 2>, o, 640); break;
            }
        } else if (ordered) {
            switch (poly) {
                case 2: launch(encoder_attention_q2_kernel<OT, 2, true, 1>, o, kQ2Threads); break;
                case 3: launch(encoder_attention_q2_kernel<OT, 3, true, 1>, o, kQ2Threads); break;
                case 4: launch(encoder_attention_q2_kernel<OT, 4, true, 1>, o, kQ2Threads); break;
                default: launch(encoder_attention_q2_kernel<OT, 0, true, 1>, o, kQ2Threads); break;
            }
        } else {
            switch (poly) {
                case 2: launch(encoder_attention_q2_kernel<OT, 2, false, 1>, o, kQ2Threads); break;
                case 3: launch(encoder_attention_q2_kernel<OT, 3, false, 1>, o, kQ2Threads); break;
                case 4: launch(encoder_attention_q2_kernel<OT, 4, false, 1>, o, kQ2Threads); break;
                default: launch(encoder_attention_q2_kernel<OT, 0, false, 1>, o, kQ2Threads); break;
            }
        }
    };
    if (dtype == WK_DTYPE_F16) pick((__half*)out);
    else pick((__nv_bfloat16*)out);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(fa q2): %s", cudaGetErrorString(e)); return WK_ERR_CUDA; }
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("encoder_attention_q2 launch: %s", cudaGetErrorString(e)); return WK_ERR_CUDA; }
    return WK_OK;
}

}  // namespace wk

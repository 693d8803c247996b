// K4: encoder self-attention (non-causal, head dim 64, S = 1500) on tcgen05 tensor cores.
//
// One CTA = one (window, head, 128-query tile).  192 threads, warp-specialised:
//   warp 0      TMA producer: Q tile once, then K / V tiles of 128 keys through 3-deep smem rings (128B swizzle)
//   warp 1      MMA issuer:   S_j = Q K_j^T   (tcgen05.mma 128x128x16 x4, accumulator S in TMEM, double buffered)
//                             PV_j = P_j V_j  (tcgen05.mma 128x64x16 x8, A = P_j from smem, B = V_j MN-major from smem)
//   warps 2..   softmax:      P (2, or 4) threads per query row (warps w, w+4, ... share a TMEM lane quarter); for P = 2 each takes 64
//                             of the tile's 128 keys and 32 of the 64 output columns.  tcgen05.ld S_j -> max / exp2 /
//                             row sum in registers (only the row max is exchanged, through smem, once per tile),
//                             P_j (16-bit) -> swizzled smem for the second MMA; PV_j is read back from TMEM and
//                             accumulated into a register-resident O with the online-softmax rescale (no TMEM
//                             read-modify-write).  Two softmax warps per scheduler hide each other's latencies.
// Q/K/V are read in place from the packed qkv activation [B*T, 3*d_model] through one 3-D tensor map (per-window
// out-of-bounds rows are zero-filled by TMA; keys >= T are masked to -inf before the softmax).
// Reference counterpart: inside AudioEncoder.mlmodelc (Sources/WhisperKit/Core/AudioEncoder.swift:59-62).
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace wk {

static constexpr int kFaThreadsFor(int parts) { return 64 + 128 * parts; }   // TMA warp + MMA warp + 4 * parts softmax warps
static constexpr int kFaBM = 128;          // queries per CTA
static constexpr int kFaBN = 128;          // keys per tile
static constexpr int kFaD = 64;
static constexpr int kFaTile = kFaBN * kFaD * 2;   // 16 KiB: one K or V or Q tile, 128-byte rows
static constexpr int kFaStages = 3;
static constexpr int kFaPBytes = kFaBM * kFaBN * 2;  // 32 KiB: P tile = two 64-key chunks of 16 KiB
static constexpr int kFaSmem = kFaTile /*Q*/ + 2 * kFaStages * kFaTile /*K,V rings*/ + 2 * kFaPBytes + 1024 /*align*/ + 512 /*barriers*/ + 8192 /*row-max / row-sum exchange*/;
static constexpr int kFaTmemCols = 512;    // S0 @0, S1 @128, O0 @256, O1 @320

struct FaParams {
    int T, H, dm, n_kv_tiles;
    float scale_log2e;
    uint32_t idesc_qk, idesc_pv;
    // V (MN-major) descriptor knobs, overridable for bring-up experiments
    uint32_t v_lbo, v_sbo, v_kstep;
};

__device__ __forceinline__ float ex2_approx(float x) {   // MUFU.EX2: 2^x, ex2(-inf) = +0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr, uint32_t lbo16, uint32_t sbo16) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(lbo16 & 0x3FFF) << 16;
    d |= (uint64_t)(sbo16 & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// 32 lanes x 16 columns of 32-bit
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// P = softmax threads per query row: each takes 128 / P of the tile's keys and 64 / P of the output columns
template <typename T, int P>
__global__ void __launch_bounds__(kFaThreadsFor(P), 1)
encoder_attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_qkv, T* __restrict__ out, const FaParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + kFaTile;
    uint8_t* sV = sK + kFaStages * kFaTile;
    uint8_t* sP = sV + kFaStages * kFaTile;           // 2 buffers x 32 KiB
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kFaPBytes);
    uint64_t* q_full = bars;                      // 1
    uint64_t* k_full = bars + 1;                  // 3
    uint64_t* k_empty = k_full + kFaStages;       // 3
    uint64_t* v_full = k_empty + kFaStages;       // 3
    uint64_t* v_empty = v_full + kFaStages;       // 3
    uint64_t* s_full = v_empty + kFaStages;       // 2
    uint64_t* s_empty = s_full + 2;               // 2
    uint64_t* p_full = s_empty + 2;               // 2
    uint64_t* p_empty = p_full + 2;               // 2
    uint64_t* o_full = p_empty + 2;               // 2
    uint64_t* o_empty = o_full + 2;               // 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);
    float* xch = reinterpret_cast<float*>(tmem_slot + 2);   // [2 tile parity][P parts][128 rows] row-max exchange, then [P][128] final sums

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q_tile = blockIdx.x, bh = blockIdx.y;
    const int b = bh / p.H, h = bh % p.H;
    const int q0 = q_tile * kFaBM;
    const int n = p.n_kv_tiles;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_qkv);
        mbar_init(q_full, 1);
        for (int i = 0; i < kFaStages; ++i) {
            mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4 * P);
            mbar_init(&p_full[i], 4 * P); mbar_init(&p_empty[i], 1);
            mbar_init(&o_full[i], 1); mbar_init(&o_empty[i], 4 * P);
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

    if (warp == 0) {
        // ============================ TMA producer ============================
        if (lane == 0) {
            mbar_expect_tx(q_full, kFaTile);
            tma_load_3d(sQ, &tm_qkv, q_full, h * kFaD, q0, b);
            for (int j = 0; j < n; ++j) {
                const int st = j % kFaStages;
                const uint32_t ph = (j / kFaStages) & 1;
                mbar_wait(&k_empty[st], ph ^ 1);
                mbar_expect_tx(&k_full[st], kFaTile);
                tma_load_3d(sK + st * kFaTile, &tm_qkv, &k_full[st], p.dm + h * kFaD, j * kFaBN, b);
                mbar_wait(&v_empty[st], ph ^ 1);
                mbar_expect_tx(&v_full[st], kFaTile);
                tma_load_3d(sV + st * kFaTile, &tm_qkv, &v_full[st], 2 * p.dm + h * kFaD, j * kFaBN, b);
            }
        }
    } else if (warp == 1) {
        // ============================ MMA issuer ============================
        auto mma1 = [&](int j) {   // S[j&1] = Q K_j^T
            const int st = j % kFaStages;
            const int sb = j & 1;
            mbar_wait(&k_full[st], (j / kFaStages) & 1);
            mbar_wait(&s_empty[sb], ((j >> 1) & 1) ^ 1);
            tc_fence_after();
            if (lane == 0) {
                const uint64_t adesc = make_sw128_desc(smem_u32(sQ), 1, 64);
                const uint64_t bdesc = make_sw128_desc(smem_u32(sK + st * kFaTile), 1, 64);
#pragma unroll
                for (int k = 0; k < kFaD / 16; ++k)
                    tc_mma_f16(tmem + sb * 128, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), p.idesc_qk, k > 0 ? 1u : 0u);
                tc_commit(&k_empty[st]);
                tc_commit(&s_full[sb]);
            }
            __syncwarp();
        };
        auto mma2 = [&](int j) {   // O[j&1] = P_j V_j
            const int st = j % kFaStages;
            const int sb = j & 1;
            mbar_wait(&v_full[st], (j / kFaStages) & 1);
            mbar_wait(&p_full[sb], (j >> 1) & 1);
            mbar_wait(&o_empty[sb], ((j >> 1) & 1) ^ 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t pbase = smem_u32(sP + sb * kFaPBytes);
                const uint32_t vbase = smem_u32(sV + st * kFaTile);
#pragma unroll
                for (int k = 0; k < kFaBN / 16; ++k) {
                    // A = P: K-major, 64-key chunks of 16 KiB, 32 B per 16 keys inside a chunk
                    const uint64_t adesc = make_sw128_desc(pbase + (k >> 2) * (kFaPBytes / 2) + (k & 3) * 32, 1, 64);
                    // B = V: MN-major (d contiguous, 128-byte rows), 16 keys = 16 rows per step
                    const uint64_t bdesc = make_sw128_desc(vbase + k * p.v_kstep, p.v_lbo, p.v_sbo);
                    tc_mma_f16(tmem + 256 + sb * 64, adesc, bdesc, p.idesc_pv, k > 0 ? 1u : 0u);
                }
                tc_commit(&v_empty[st]);
                tc_commit(&p_empty[sb]);
                tc_commit(&o_full[sb]);
            }
            __syncwarp();
        };
        mbar_wait(q_full, 0);
        mma1(0);
        if (n > 1) mma1(1);
        for (int j = 0; j < n; ++j) {
            mma2(j);
            if (j + 2 < n) mma1(j + 2);
        }
    } else {
        // ============================ softmax / accumulate / epilogue ============================
        constexpr int KT = kFaBN / P;                 // keys of the tile per thread (64 or 32)
        constexpr int OC = kFaD / P;                  // output columns per thread (32 or 16)
        constexpr int NCH = KT / 32;                  // 32-column TMEM loads per tile
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may touch
        const int part = (warp - 2) >> 2;             // which slice of the keys / output columns
        const int row = quarter * 32 + lane;          // query row inside the tile == TMEM lane
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const uint32_t pair_bar = 1 + quarter;        // named barrier shared by the P warps of this lane quarter
        float o[OC];
#pragma unroll
        for (int i = 0; i < OC; ++i) o[i] = 0.f;
        float m_run = -INFINITY;      // running max of raw scores (identical in all threads of a row)
        float m_acc = -INFINITY;      // max the register accumulator o[] is currently scaled to
        float l_run = 0.f;            // this thread's share of the row sum
        float m_tile_prev = -INFINITY;
        const float c = p.scale_log2e;
        const int sw = row & 7;

        auto accumulate = [&](int i, float m_i) {   // o += PV_i[:, part*OC .. +OC], PV_i is relative to max m_i
            const int sb = i & 1;
            mbar_wait(&o_full[sb], (i >> 1) & 1);
            tc_fence_after();
            const float corr = ex2_approx((m_acc - m_i) * c);   // m_acc = -inf on first use -> 0
            m_acc = m_i;
            uint32_t r[OC];
            if constexpr (OC == 32) tmem_ld_32x32(tmem + lane_addr + 256 + sb * 64 + part * OC, r);
            else tmem_ld_32x16(tmem + lane_addr + 256 + sb * 64 + part * OC, r);
            tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < OC; ++t) o[t] = fmaf(o[t], corr, __uint_as_float(r[t]));
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_empty[sb]);
        };

        for (int j = 0; j < n; ++j) {
            const int sb = j & 1;
            const uint32_t ph = (j >> 1) & 1;
            mbar_wait(&s_full[sb], ph);
            tc_fence_after();
            // this thread's KT scores into registers (all TMEM loads in flight), then release the S buffer
            uint32_t sr[NCH][32];
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) tmem_ld_32x32(tmem + lane_addr + sb * 128 + part * KT + ch * 32, sr[ch]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[sb]);
            const int valid = p.T - j * kFaBN - part * KT;   // keys of this thread's slice that exist (>= KT except at the end)
            if (valid < KT) {
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
                    for (int t = 0; t < 32; ++t)
                        if (ch * 32 + t >= valid) sr[ch][t] = 0xff800000u;   // -inf: key does not exist
            }
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
                for (int t = 0; t < 32; t += 2) {
                    mx0 = fmaxf(mx0, __uint_as_float(sr[ch][t]));
                    mx1 = fmaxf(mx1, __uint_as_float(sr[ch][t + 1]));
                }
            // exchange the slice max with the partner threads (same row, other keys)
            float* slot = xch + (j & 1) * (P * 128);
            slot[part * 128 + row] = fmaxf(mx0, mx1);
            asm volatile("bar.sync %0, %1;" ::"r"(pair_bar), "n"(32 * P) : "memory");
            float mx = m_run;
#pragma unroll
            for (int q = 0; q < P; ++q) mx = fmaxf(mx, slot[q * 128 + row]);
            const float l_corr = ex2_approx((m_run - mx) * c);
            const float msc = mx * c;
            m_run = mx;
            // P buffer free?  (PV of tile j-2 has consumed it)
            mbar_wait(&p_empty[sb], ph ^ 1);
            float ls0 = 0.f, ls1 = 0.f, ls2 = 0.f, ls3 = 0.f;
            // P tile in smem = two 64-key chunk blocks of 16 KiB (128-byte rows, 128B swizzle); this thread's keys part*KT .. +KT
            uint8_t* prow = sP + sb * kFaPBytes + ((part * KT) >> 6) * (kFaPBytes / 2) + row * 128;
            const int chunk0 = ((part * KT) & 63) >> 3;   // first 16-byte chunk of the slice inside the 128-byte row
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                uint32_t pk[16];
#pragma unroll
                for (int t = 0; t < 32; t += 4) {
                    const float p0 = ex2_approx(fmaf(__uint_as_float(sr[ch][t]), c, -msc));
                    const float p1 = ex2_approx(fmaf(__uint_as_float(sr[ch][t + 1]), c, -msc));
                    const float p2 = ex2_approx(fmaf(__uint_as_float(sr[ch][t + 2]), c, -msc));
                    const float p3 = ex2_approx(fmaf(__uint_as_float(sr[ch][t + 3]), c, -msc));
                    ls0 += p0; ls1 += p1; ls2 += p2; ls3 += p3;
                    pk[t >> 1] = T16<T>::pack2(p0, p1);
                    pk[(t >> 1) + 1] = T16<T>::pack2(p2, p3);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int chunk = chunk0 + ch * 4 + q;   // 16-byte chunk inside the 128-byte row
                    *reinterpret_cast<uint4*>(prow + ((chunk ^ sw) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                }
            }
            l_run = l_run * l_corr + ((ls0 + ls1) + (ls2 + ls3));
            fence_proxy_async();        // generic-proxy smem writes -> visible to the tensor-core (async) proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[sb]);
            // fold in the previous tile's PV while this tile's second MMA runs
            if (j > 0) accumulate(j - 1, m_tile_prev);
            m_tile_prev = mx;
        }
        accumulate(n - 1, m_tile_prev);
        // ---- epilogue: total row sum = all slices; normalise and store this thread's OC columns
        float* fin = xch + 2 * P * 128;   // [P][128]
        fin[part * 128 + row] = l_run;
        asm volatile("bar.sync %0, %1;" ::"r"(pair_bar), "n"(32 * P) : "memory");
        const int q = q0 + row;
        if (q < p.T) {
            float tot = 0.f;
#pragma unroll
            for (int i = 0; i < P; ++i) tot += fin[i * 128 + row];
            const float inv = 1.f / tot;
            uint4* dst = reinterpret_cast<uint4*>(out + ((long long)b * p.T + q) * p.dm + h * kFaD + part * OC);
#pragma unroll
            for (int i = 0; i < OC / 8; ++i) {
                uint4 v;
                v.x = T16<T>::pack2(o[8 * i] * inv, o[8 * i + 1] * inv);
                v.y = T16<T>::pack2(o[8 * i + 2] * inv, o[8 * i + 3] * inv);
                v.z = T16<T>::pack2(o[8 * i + 4] * inv, o[8 * i + 5] * inv);
                v.w = T16<T>::pack2(o[8 * i + 6] * inv, o[8 * i + 7] * inv);
                dst[i] = v;
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
    p.v_lbo = 1; p.v_sbo = 64; p.v_kstep = 2048;
    dim3 grid((T + kFaBM - 1) / kFaBM, B * n_heads);
    // P = 2 softmax threads per query row.  P = 4 (16 softmax warps, 32 keys and 16 output columns per thread) was built and measured on
    // B200: correct, but slower (2.24 vs 2.05 ms per layer at 64 windows) - twice the row-max exchanges and named-barrier waits cost more
    // than the extra warps hide - so only P = 2 is instantiated.
    cudaError_t e = cudaSuccess;
    if (dtype == WK_DTYPE_F16) {
        static bool set = false;
        if (!set) { e = cudaFuncSetAttribute(encoder_attention_tcgen05_kernel<__half, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFaSmem); set = e == cudaSuccess; }
        if (e == cudaSuccess) encoder_attention_tcgen05_kernel<__half, 2><<<grid, kFaThreadsFor(2), kFaSmem, stream>>>(tm, (__half*)out, p);
    } else {
        static bool set = false;
        if (!set) { e = cudaFuncSetAttribute(encoder_attention_tcgen05_kernel<__nv_bfloat16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFaSmem); set = e == cudaSuccess; }
        if (e == cudaSuccess) encoder_attention_tcgen05_kernel<__nv_bfloat16, 2><<<grid, kFaThreadsFor(2), kFaSmem, stream>>>(tm, (__nv_bfloat16*)out, p);
    }
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(fa): %s", cudaGetErrorString(e)); return WK_ERR_CUDA; }
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("encoder_attention_tcgen05 launch: %s", cudaGetErrorString(e)); return WK_ERR_CUDA; }
    return WK_OK;
}

}  // namespace wk

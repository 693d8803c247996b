// Persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[rows, cols] = A[rows, K] * B[cols, K]^T      (both operands K-major, 16-bit, f32 accumulate in TMEM)
//
// One kernel serves every dense contraction of the Whisper hot path:
//   * encoder / cross-KV projections: A = activations (M = B*1500 rows), B = weights [N, K]
//   * conv stem as implicit GEMM: A is a 3-D tensor map, the 3 taps are extra K-blocks with a row shift
//   * decoder (M = batch <= 256): swap-AB, A = weights (128 output features per tile), B = activations,
//     split-K partials written transposed so the next fused reduce(+LN) kernel reads them coalesced.
//
// Structure (320 threads, 1 CTA / SM, persistent over work items):
//   warp 0      TMA producer: cp.async.bulk.tensor (128B swizzle) into a ring of smem stages, mbarrier tx
//   warp 1      MMA issuer: one lane issues tcgen05.mma kind::f16 128xBNx16, tcgen05.commit frees stages
//   warps 2..9  epilogue: tcgen05.ld 32x32b from a double-buffered TMEM accumulator -> fused epilogue -> HBM
//               (two warps per TMEM lane quarter, alternating 32-column chunks, so erf-GELU epilogues keep up with the MMAs)
// The WhisperKit reference has no counterpart source for this file: the contraction lives inside
// AudioEncoder.mlmodelc / TextDecoder.mlmodelc (Sources/WhisperKit/Core/AudioEncoder.swift:59-62,
// Sources/WhisperKit/Core/TextDecoder.swift:394-417).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace wk {

static constexpr int kBlockM = 128;
static constexpr int kBlockK = 64;   // 64 x 2 B = one 128-byte swizzle row
static constexpr int kUmmaK = 16;
static constexpr int kStageA = kBlockM * kBlockK * 2;  // 16 KiB
static constexpr int kGemmThreads = 320;   // TMA warp + MMA warp + 8 epilogue warps (two per TMEM lane quarter)
static constexpr int kTmemCols = 512;
static constexpr int kAccStride = 256;  // TMEM columns per accumulator stage
static constexpr int kMaxStages = 10;
static constexpr int kEpiStride = 20;      // floats per staged row: 16 values + 4 pad (conflict-free 16-byte column accesses)
static constexpr int kEpiBytes = 8 * 32 * kEpiStride * 4;   // 20 KiB

struct GemmKParams {
    int tiles_per_batch, n_batches, tiles_n, splits, work;
    int kb_per_tap, kb_per_split, taps;
    int tap_row_shift[3];
    int tap_col_off[3];
    int a_is_3d;
    int m_rows_per_batch, n, bn;
    uint32_t idesc;
    int stage_b_bytes, stages;
    int mode, gelu;
    void* out;
    long long ld_out, out_rows_per_batch, partial_cols;
    const float* bias;
    const float* pos;
    long long ld_pos;
    int heads_T, heads_B, heads_H, heads_dmodel;
    int tmem_cols;
    int a_static;     // see GemmDesc::a_static
};

// fused epilogue of one 32-column chunk of the accumulator row held by this thread (shared by the single-CTA and the CTA-pair kernels)
// residual prefetch for GEMM_OUT_F32_ADD (out += A W^T + b): the 32 f32 values this thread will add into do not depend on the accumulator, so
// they are loaded BEFORE the wait for the MMAs (first chunk) / while the previous chunk is processed, instead of as a dependent round
// trip to L2 per chunk after the accumulator is ready
__device__ __forceinline__ void gemm_prefetch_residual(const GemmKParams& p, long long grow, bool row_ok, int col_base, int c, float4 (&pre)[8]) {
    if (p.mode != GEMM_OUT_F32_ADD || !row_ok || c >= p.bn || col_base + c >= p.n) return;
    const float4* o4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.out) + grow * p.ld_out + col_base + c);
#pragma unroll
    for (int j = 0; j < 8; ++j) pre[j] = o4[j];
}

template <typename T>
__device__ __forceinline__ void gemm_epilogue_chunk(const GemmKParams& p, const uint32_t (&r)[32], int split, long long grow, int row_in_batch,
                                                    bool row_ok, int col_base, int c, const float4 (&pre)[8]) {
    const int col0 = col_base + c;
    if (p.mode == GEMM_OUT_PARTIAL_T) {
        if (row_ok) {
            float* o = reinterpret_cast<float*>(p.out);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int col = col0 + j;
                if (c + j < p.bn && col < p.partial_cols)
                    o[((long long)split * p.partial_cols + col) * p.ld_out + grow] = __uint_as_float(r[j]);
            }
        }
        return;
    }
    if (!row_ok || col0 >= p.n) return;
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (p.bias) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 bb = __ldg(b4 + j);
            v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
        }
    }
    if (p.gelu) {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            const float2 g = gelu_erf2(make_float2(v[j], v[j + 1]));
            v[j] = g.x; v[j + 1] = g.y;
        }
    }
    if (p.mode == GEMM_OUT_T16 || p.mode == GEMM_OUT_T16_HEADS) {
        T* o;
        if (p.mode == GEMM_OUT_T16) {
            o = reinterpret_cast<T*>(p.out) + grow * p.ld_out + col0;
        } else {
            const int b = (int)(grow / p.heads_T);
            const int tt = (int)(grow - (long long)b * p.heads_T);
            const int which = col0 / p.heads_dmodel;
            const int rem = col0 - which * p.heads_dmodel;
            const int h = rem >> 6;
            const int dd = rem & 63;
            o = reinterpret_cast<T*>(p.out) +
                ((((long long)which * p.heads_B + b) * p.heads_H + h) * p.heads_T + tt) * 64 + dd;
        }
        uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint4 pk;
            pk.x = T16<T>::pack2(v[8 * j], v[8 * j + 1]);
            pk.y = T16<T>::pack2(v[8 * j + 2], v[8 * j + 3]);
            pk.z = T16<T>::pack2(v[8 * j + 4], v[8 * j + 5]);
            pk.w = T16<T>::pack2(v[8 * j + 6], v[8 * j + 7]);
            o4[j] = pk;
        }
    } else {
        float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + grow * p.ld_out + col0);
        if (p.mode == GEMM_OUT_F32_ADD) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 x = pre[j];
                x.x += v[4 * j]; x.y += v[4 * j + 1]; x.z += v[4 * j + 2]; x.w += v[4 * j + 3];
                o4[j] = x;
            }
        } else if (p.mode == GEMM_OUT_F32_GELU_POS) {
            const float4* p4 = reinterpret_cast<const float4*>(p.pos + (long long)row_in_batch * p.ld_pos + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 pp = __ldg(p4 + j);
                o4[j] = make_float4(v[4 * j] + pp.x, v[4 * j + 1] + pp.y, v[4 * j + 2] + pp.z, v[4 * j + 3] + pp.w);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmKParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // manual 1024-byte alignment (SWIZZLE_128B atoms are 1024 B)
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = kStageA + p.stage_b_bytes;
    uint8_t* smem_tail = smem + (size_t)p.stages * stage_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_tail);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tfull_bar = empty_bar + kMaxStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    pdl_launch_dependents();   // the next kernel may start its prologue; it still waits for our completion

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 8);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, p.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // upstream results visible from here on (barrier init / TMEM alloc overlapped its tail).  With a static A operand the producer warp
    // waits later, after it has put the first weight tiles in flight; every other warp only sees data that arrived after that wait.
    const bool early_a = p.a_static != 0;
    if (!(early_a && warp == 0)) pdl_wait();

    if (warp == 0) {
        // ===================== TMA producer =====================
        // the whole warp runs this converged; the TMA / mbarrier instructions are predicated on the elect.sync lane (common.cuh)
        {
            int stage = 0;
            uint32_t phase = 0;
            bool need_wait = early_a;
            auto issue = [&](int kb, int st, int row0, int batch, int n_tile, bool do_a, bool do_b) {
                const int tap = kb / p.kb_per_tap;
                const int kk = kb - tap * p.kb_per_tap;
                uint8_t* sa = smem + (size_t)st * stage_bytes;
                uint8_t* sb = sa + kStageA;
                if (do_a) {
                    mbar_expect_tx_elect(&full_bar[st], (uint32_t)stage_bytes);
                    const int ac0 = p.tap_col_off[tap] + kk * kBlockK;
                    const int ar = row0 + p.tap_row_shift[tap];
                    if (p.a_is_3d) tma_load_3d_elect(sa, &tmA, &full_bar[st], ac0, ar, batch);
                    else tma_load_2d_elect(sa, &tmA, &full_bar[st], ac0, ar);
                }
                if (do_b) tma_load_2d_elect(sb, &tmB, &full_bar[st], kb * kBlockK, n_tile * p.bn);
            };
            for (int w = blockIdx.x; w < p.work; w += gridDim.x) {
                const int split = w % p.splits;
                const int t = w / p.splits;
                const int n_tile = t % p.tiles_n;
                const int m_tile = t / p.tiles_n;
                const int batch = m_tile / p.tiles_per_batch;
                const int row0 = (m_tile % p.tiles_per_batch) * kBlockM;
                const int kb0 = split * p.kb_per_split;
                int kb = kb0;
                if (need_wait) {
                    // first work item, fresh ring: weight tiles go out before griddepcontrol.wait, activation tiles after it
                    const int n_pre = p.kb_per_split < p.stages ? p.kb_per_split : p.stages;
                    for (int i = 0; i < n_pre; ++i) issue(kb0 + i, i, row0, batch, n_tile, true, false);
                    pdl_wait();
                    for (int i = 0; i < n_pre; ++i) issue(kb0 + i, i, row0, batch, n_tile, false, true);
                    kb = kb0 + n_pre;
                    if (n_pre == p.stages) { stage = 0; phase ^= 1; } else stage = n_pre;
                    need_wait = false;
                }
                for (; kb < kb0 + p.kb_per_split; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    __syncwarp();
                    issue(kb, stage, row0, batch, n_tile, true, true);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
            if (need_wait) pdl_wait();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        const uint32_t idesc = p.idesc;
        for (int w = blockIdx.x; w < p.work; w += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * kAccStride;
            for (int kb = 0; kb < p.kb_per_split; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                __syncwarp();   // converged from here: the tcgen05 instructions below are predicated on the elect.sync lane (common.cuh)
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                const uint64_t adesc = make_kmajor_sw128_desc(sa);
                const uint64_t bdesc = make_kmajor_sw128_desc(sa + kStageA);
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                    // advance the start address by k*32 bytes inside the swizzle atom (>>4 -> +2k)
                    tc_mma_f16_elect(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
                }
                tc_commit_elect(&empty_bar[stage]);
                if (kb == p.kb_per_split - 1) tc_commit_elect(&tfull_bar[acc]);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue warps =====================
        const int quarter = warp & 3;  // TMEM lane quarter this warp may access
        const int csub = (warp - 2) >> 2;  // which of the two warps of this quarter: takes the 32-column chunks with (c / 32) % 2 == csub
        int it = 0;
        for (int w = blockIdx.x; w < p.work; w += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int split = w % p.splits;
            const int t = w / p.splits;
            const int n_tile = t % p.tiles_n;
            const int m_tile = t / p.tiles_n;
            const int batch = m_tile / p.tiles_per_batch;
            const int row_in_batch = (m_tile % p.tiles_per_batch) * kBlockM + quarter * 32 + lane;
            const bool row_ok = row_in_batch < p.m_rows_per_batch;
            const long long grow = (long long)batch * p.out_rows_per_batch + row_in_batch;

            const int col_base = n_tile * p.bn;
            float4 pre[8], pre_next[8];
            gemm_prefetch_residual(p, grow, row_ok, col_base, csub * 32, pre);
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * kAccStride + ((uint32_t)(quarter * 32) << 16);

            for (int c = csub * 32; c < p.bn; c += 128) {   // two chunks per trip: the residual buffers alternate, no register copies
                uint32_t r[32];
                __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the divergent tails below
                tmem_ld_32x32(taddr + c, r);
                gemm_prefetch_residual(p, grow, row_ok, col_base, c + 64, pre_next);
                tmem_ld_wait();
                gemm_epilogue_chunk<T>(p, r, split, grow, row_in_batch, row_ok, col_base, c, pre);
                if (c + 64 < p.bn) {
                    __syncwarp();
                    tmem_ld_32x32(taddr + c + 64, r);
                    gemm_prefetch_residual(p, grow, row_ok, col_base, c + 128, pre);
                    tmem_ld_wait();
                    gemm_epilogue_chunk<T>(p, r, split, grow, row_in_batch, row_ok, col_base, c + 64, pre_next);
                }
            }
            // release the accumulator stage back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant for the big encoder GEMMs (plain 2-D operands, one split): a cluster of two CTAs works on two vertically adjacent
// 128-row tiles of the SAME column block, so both need the same weight (B) tile.  Each CTA fetches half of it and TMA multicasts the
// half into both CTAs' shared memory: per k-block a CTA pulls 16 KB (A) + BN/2 x 128 B (its half of B) through L2 instead of
// 16 KB + BN x 128 B - at BN = 256 that is 32 KB instead of 48 KB, and L2 -> SM bandwidth is what bounds the single-CTA kernel
// (48 KB per 4.2 MFLOP = 26 TB/s at tensor peak).  MMAs stay cta_group::1 (each CTA multiplies its own A tile by the full B tile in
// its own shared memory); what crosses CTAs is the multicast load and the stage-release: a stage may be overwritten only when BOTH CTAs'
// MMAs have read it, so tcgen05.commit multicasts its arrive to both CTAs' empty barriers (count 2).
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_multicast(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_multicast_elect(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;\n\t}"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tc_commit_multicast_elect(uint64_t* bar, uint16_t mask) {   // converged warp, see tc_mma_f16_elect
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tc_commit_multicast(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}

template <typename T>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhalf, const GemmKParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = kStageA + p.stage_b_bytes;
    uint8_t* smem_tail = smem + (size_t)p.stages * stage_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_tail);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tfull_bar = empty_bar + kMaxStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    uint8_t* smem_epi = smem_tail + 512;   // 8 epilogue warps x 32 rows x kEpiStride floats: the transpose buffers of the f32 read-modify-write epilogue

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();          // 0 / 1: upper / lower tile of the pair, first / second half of B
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int half_bytes = p.stage_b_bytes / 2;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmBhalf);
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 2);   // this CTA's MMAs and the peer's (multicast commit)
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 8);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, p.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();            // the peer's barriers exist before anything is multicast at them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        // the whole warp runs this converged; the TMA / mbarrier instructions are predicated on the elect.sync lane (common.cuh)
        {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = pair; w < p.work; w += n_pairs) {
                const int n_tile = w % p.tiles_n;
                const int m_tile = 2 * (w / p.tiles_n) + (int)rank;
                for (int kb = 0; kb < p.kb_per_split; ++kb) {
                    mbar_wait_bounded(&empty_bar[stage], phase ^ 1);
                    __syncwarp();
                    uint8_t* sa = smem + (size_t)stage * stage_bytes;
                    mbar_expect_tx_elect(&full_bar[stage], (uint32_t)stage_bytes);   // own A tile + both halves of B
                    tma_load_2d_elect(sa, &tmA, &full_bar[stage], kb * kBlockK, m_tile * kBlockM);
                    tma_load_2d_multicast_elect(sa + kStageA + rank * half_bytes, &tmBhalf, &full_bar[stage], kb * kBlockK,
                                                n_tile * p.bn + (int)rank * (p.bn / 2), (uint16_t)3);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        const uint32_t idesc = p.idesc;
        for (int w = pair; w < p.work; w += n_pairs, ++it) {
            const int acc = it & 1;
            mbar_wait_bounded(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * kAccStride;
            for (int kb = 0; kb < p.kb_per_split; ++kb) {
                mbar_wait_bounded(&full_bar[stage], phase);
                __syncwarp();   // converged from here: the tcgen05 instructions below are predicated on the elect.sync lane (common.cuh)
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                const uint64_t adesc = make_kmajor_sw128_desc(sa);
                const uint64_t bdesc = make_kmajor_sw128_desc(sa + kStageA);
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k)
                    tc_mma_f16_elect(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
                tc_commit_multicast_elect(&empty_bar[stage], (uint16_t)3);   // releases the stage in both CTAs
                if (kb == p.kb_per_split - 1) tc_commit_elect(&tfull_bar[acc]);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue warps =====================
        const int quarter = warp & 3;
        const int csub = (warp - 2) >> 2;
        // tile w -> the first output row of this warp's lane quarter and the first output column of the tile (plain 2-D product, one batch)
        auto tile_row0 = [&](int w) { return (2 * (w / p.tiles_n) + (int)rank) * kBlockM + quarter * 32; };
        auto tile_col = [&](int w) { return (w % p.tiles_n) * p.bn; };
        if (p.mode == GEMM_OUT_F32_ADD) {
            // out += A W^T + b in place (out-proj, FC2: the f32 residual stream).  A thread owns a TMEM lane = an output ROW, so storing from
            // the accumulator registers directly means 16-byte pieces at a 5 KB stride: 32 L1 transactions per warp instruction, and the
            // read-modify-write doubles them - measured, the epilogue of a 128x256 tile then takes longer than its 20 k-blocks of MMAs
            // (out-proj: tensor pipe 36 %).  Here each warp transposes its 32x16 sub-chunks through a padded shared-memory buffer and
            // touches global memory with 4 lanes per row (64 contiguous bytes): 8 transactions per instruction.  The residual values are
            // fetched one sub-chunk pair (and, across tiles, one tile) ahead, because this kernel is epilogue-bound exactly in this mode.
            float* stg = reinterpret_cast<float*>(smem_epi) + (warp - 2) * (32 * kEpiStride);
            const int trow = lane >> 2, tcol = (lane & 3) * 4;       // transposed view: 8 rows x 4 float4 per pass, 4 passes per sub-chunk
            float* out32 = reinterpret_cast<float*>(p.out);
            auto fetch = [&](int row0, int col, float4 (&x)[8]) {   // residual of one 32-column chunk = two 16-column halves x 4 passes
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int r = row0 + (q & 3) * 8 + trow, c = col + (q >> 2) * 16 + tcol;
                    if (r < p.m_rows_per_batch && c < p.n) x[q] = *reinterpret_cast<const float4*>(out32 + (long long)r * p.ld_out + c);
                }
            };
            float4 pre[8], pre_next[8];
            if (pair < p.work) fetch(tile_row0(pair), tile_col(pair) + csub * 32, pre);
            int it = 0;
            for (int w = pair; w < p.work; w += n_pairs, ++it) {
                const int acc = it & 1;
                const int row0 = tile_row0(w), col_base = tile_col(w);
                const int wn = w + n_pairs;
                const bool has_next = wn < p.work;
                mbar_wait_bounded(&tfull_bar[acc], (it >> 1) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + acc * kAccStride + ((uint32_t)(quarter * 32) << 16);
                auto chunk = [&](int c, const float4 (&x)[8]) {
                    uint32_t r[32];
                    __syncwarp();
                    tmem_ld_32x32(taddr + c, r);
                    const int col0 = col_base + c;
                    float4 bb[8];   // the chunk's 32 bias values (one broadcast address per load), in flight under the TMEM load
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        bb[j] = (p.bias && col0 + 4 * j < p.n) ? __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    tmem_ld_wait();
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        // my row's 16 values (+ bias) -> smem, then 4 passes of 8 rows x 64 bytes: add the prefetched residual, store
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 b4 = bb[half * 4 + j];
                            const float4 v = make_float4(__uint_as_float(r[half * 16 + 4 * j]) + b4.x, __uint_as_float(r[half * 16 + 4 * j + 1]) + b4.y,
                                                         __uint_as_float(r[half * 16 + 4 * j + 2]) + b4.z, __uint_as_float(r[half * 16 + 4 * j + 3]) + b4.w);
                            *reinterpret_cast<float4*>(stg + lane * kEpiStride + 4 * j) = v;
                        }
                        __syncwarp();
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int rr = q * 8 + trow;
                            const int gr = row0 + rr, gc = col0 + half * 16 + tcol;
                            const float4 v = *reinterpret_cast<const float4*>(stg + rr * kEpiStride + tcol);
                            if (gr < p.m_rows_per_batch && gc < p.n) {
                                float4 o = x[half * 4 + q];
                                o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
                                *reinterpret_cast<float4*>(out32 + (long long)gr * p.ld_out + gc) = o;
                            }
                        }
                        __syncwarp();
                    }
                };
                for (int c = csub * 32; c < p.bn; c += 128) {   // two chunks per trip: the residual buffers alternate, no register copies
                    if (c + 64 < p.bn) fetch(row0, col_base + c + 64, pre_next);
                    else if (has_next) fetch(tile_row0(wn), tile_col(wn) + csub * 32, pre_next);
                    chunk(c, pre);
                    if (c + 64 < p.bn) {
                        if (c + 128 < p.bn) fetch(row0, col_base + c + 128, pre);
                        else if (has_next) fetch(tile_row0(wn), tile_col(wn) + csub * 32, pre);
                        chunk(c + 64, pre_next);
                    } else if (has_next) {
                        // odd number of chunks: the next tile's first chunk sits in pre_next, the next trip starts from pre
#pragma unroll
                        for (int q = 0; q < 8; ++q) pre[q] = pre_next[q];
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            }
        } else {
            float4 pre[8];   // (unused outside GEMM_OUT_F32_ADD)
            int it = 0;
            for (int w = pair; w < p.work; w += n_pairs, ++it) {
                const int acc = it & 1;
                const int row_in_batch = tile_row0(w) + lane;
                const bool row_ok = row_in_batch < p.m_rows_per_batch;
                const long long grow = row_in_batch;
                const int col_base = tile_col(w);
                mbar_wait_bounded(&tfull_bar[acc], (it >> 1) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + acc * kAccStride + ((uint32_t)(quarter * 32) << 16);
                for (int c = csub * 32; c < p.bn; c += 64) {
                    uint32_t r[32];
                    __syncwarp();
                    tmem_ld_32x32(taddr + c, r);
                    tmem_ld_wait();
                    gemm_epilogue_chunk<T>(p, r, 0, grow, row_in_batch, row_ok, col_base, c, pre);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();            // neither CTA leaves while the peer may still multicast into it or arrive at its barriers
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

static wk_status make_tmap(CUtensorMap* tm, const void* base, int dtype, int ndim, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
        return WK_ERR_CUDA;
    }
    cuuint64_t gdim[3];
    cuuint64_t gstr[2];
    cuuint32_t bx[3];
    cuuint32_t es[3] = {1, 1, 1};
    for (int i = 0; i < ndim; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; }
    for (int i = 0; i < ndim - 1; ++i) gstr[i] = strides_bytes[i];
    CUresult r = enc(tm, dtype == WK_DTYPE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                     (cuuint32_t)ndim, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed: %d (ndim %d dims %llu %llu stride %llu box %u %u)", (int)r, ndim,
                  (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)strides_bytes[0], box[0],
                  box[1]);
        return WK_ERR_CUDA;
    }
    return WK_OK;
}

// 2-D, 128-byte-swizzled, 16-bit tensor map over a row-major [rows][ld] matrix with a [box_rows][box_cols] box (used by fused_chain.cu)
wk_status make_tmap_2d(void* tm, const void* base, int dtype, uint64_t cols, uint64_t rows, uint64_t ld_elems, uint32_t box_cols, uint32_t box_rows) {
    uint64_t dims[2] = {cols, rows};
    uint64_t str[1] = {ld_elems * 2};
    uint32_t box[2] = {box_cols, box_rows};
    return make_tmap(reinterpret_cast<CUtensorMap*>(tm), base, dtype, 2, dims, str, box);
}

wk_status gemm_tcgen05(const GemmDesc& d, int num_sms, cudaStream_t stream) {
    if (d.k % kBlockK != 0 || d.bn % 16 != 0 || d.bn < 16 || d.bn > 256 || d.taps < 1 || d.taps > 3) {
        set_error("gemm_tcgen05: unsupported shape k=%d bn=%d taps=%d", d.k, d.bn, d.taps);
        return WK_ERR_INVALID_ARGUMENT;
    }
    if (d.mode != GEMM_OUT_PARTIAL_T && (d.n % 32 != 0 || d.splits != 1)) {
        set_error("gemm_tcgen05: n=%d must be a multiple of 32 and splits 1 for mode %d", d.n, d.mode);
        return WK_ERR_INVALID_ARGUMENT;
    }
    GemmKParams p;
    memset(&p, 0, sizeof(p));
    p.kb_per_tap = d.k / kBlockK;
    const int total_kb = p.kb_per_tap * d.taps;
    p.splits = d.splits < 1 ? 1 : d.splits;
    if (total_kb % p.splits != 0) {
        set_error("gemm_tcgen05: splits %d does not divide %d k-blocks", p.splits, total_kb);
        return WK_ERR_INVALID_ARGUMENT;
    }
    p.kb_per_split = total_kb / p.splits;
    p.taps = d.taps;
    for (int i = 0; i < 3; ++i) { p.tap_row_shift[i] = d.tap_row_shift[i]; p.tap_col_off[i] = d.tap_col_off[i]; }
    p.a_is_3d = d.a_3d ? 1 : 0;
    p.n_batches = d.a_3d ? d.a_batches : 1;
    p.m_rows_per_batch = d.m_rows_per_batch;
    p.tiles_per_batch = (d.m_rows_per_batch + kBlockM - 1) / kBlockM;
    p.n = d.n;
    p.bn = d.bn;
    p.tiles_n = (d.n + d.bn - 1) / d.bn;
    p.work = p.tiles_per_batch * p.n_batches * p.tiles_n * p.splits;
    p.idesc = 0;
    {
        const int fmt = d.in_dtype == WK_DTYPE_F16 ? 0 : 1;
        uint32_t id = 0;
        id |= 1u << 4;
        id |= (uint32_t)fmt << 7;
        id |= (uint32_t)fmt << 10;
        id |= (uint32_t)(d.bn >> 3) << 17;
        id |= (uint32_t)(kBlockM >> 4) << 24;
        p.idesc = id;
    }
    p.stage_b_bytes = d.bn * kBlockK * 2;
    // B stage must keep 1024-byte alignment of the following A stage
    if (p.stage_b_bytes % 1024 != 0) p.stage_b_bytes = (p.stage_b_bytes + 1023) / 1024 * 1024;
    const int stage_bytes = kStageA + p.stage_b_bytes;
    const int smem_budget = 227 * 1024 - 1024 /*align slack*/ - 512 /*barriers*/ - kEpiBytes /*epilogue transpose buffers*/;
    int stages = smem_budget / stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages > total_kb / p.splits + 2) stages = total_kb / p.splits + 2;
    if (stages < 2) stages = 2;
    if (d.max_stages > 0 && stages > d.max_stages) stages = d.max_stages;
    p.stages = stages;
    // TMEM: two accumulator stages of kAccStride columns when a CTA may run several tiles, else the smallest power of two >= BN
    p.tmem_cols = kTmemCols;
    if (p.work <= num_sms) { int c = 32; while (c < d.bn) c <<= 1; p.tmem_cols = c; }
    p.a_static = d.a_static;
    p.mode = d.mode;
    p.gelu = d.gelu;
    p.out = d.out;
    p.ld_out = d.ld_out;
    p.out_rows_per_batch = d.out_rows_per_batch;
    p.partial_cols = d.partial_cols;
    p.bias = d.bias;
    p.pos = d.pos;
    p.ld_pos = d.ld_pos;
    p.heads_T = d.heads_T; p.heads_B = d.heads_B; p.heads_H = d.heads_H; p.heads_dmodel = d.heads_dmodel;

    CUtensorMap tmA, tmB;
    wk_status st;
    if (p.a_is_3d) {
        uint64_t dims[3] = {(uint64_t)d.a_cols, (uint64_t)d.a_rows, (uint64_t)d.a_batches};
        uint64_t str[2] = {(uint64_t)d.a_ld * 2, (uint64_t)d.a_batch_stride * 2};
        uint32_t box[3] = {kBlockK, kBlockM, 1};
        st = make_tmap(&tmA, d.a, d.in_dtype, 3, dims, str, box);
    } else {
        uint64_t dims[2] = {(uint64_t)d.a_cols, (uint64_t)d.a_rows};
        uint64_t str[1] = {(uint64_t)d.a_ld * 2};
        uint32_t box[2] = {kBlockK, kBlockM};
        st = make_tmap(&tmA, d.a, d.in_dtype, 2, dims, str, box);
    }
    if (st != WK_OK) return st;
    {
        uint64_t dims[2] = {(uint64_t)d.k * d.taps, (uint64_t)d.b_rows};
        uint64_t str[1] = {(uint64_t)d.b_ld * 2};
        uint32_t box[2] = {kBlockK, (uint32_t)d.bn};
        st = make_tmap(&tmB, d.b, d.in_dtype, 2, dims, str, box);
        if (st != WK_OK) return st;
    }
    const size_t smem_bytes = (size_t)stages * stage_bytes + 1024 + 512 + kEpiBytes;
    // CTA-pair path (two 128-row tiles of one column block per cluster, weight tile multicast): plain encoder-sized GEMMs only
    if (d.pair && !d.a_3d && d.taps == 1 && p.splits == 1 && d.mode != GEMM_OUT_PARTIAL_T && d.bn % 16 == 0 && d.bn >= 32 && num_sms >= 2 &&
        p.tiles_per_batch >= 2) {
        CUtensorMap tmBh;
        uint64_t dims[2] = {(uint64_t)d.k, (uint64_t)d.b_rows};
        uint64_t str[1] = {(uint64_t)d.b_ld * 2};
        uint32_t box[2] = {kBlockK, (uint32_t)d.bn / 2};
        st = make_tmap(&tmBh, d.b, d.in_dtype, 2, dims, str, box);
        if (st != WK_OK) return st;
        p.work = ((p.tiles_per_batch + 1) / 2) * p.tiles_n;          // pair work items
        p.tmem_cols = kTmemCols;
        int pairs = std::min(p.work, num_sms / 2);
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(kGemmThreads); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaError_t e;
        if (d.in_dtype == WK_DTYPE_F16) {
            static bool attr_set = false;
            if (!attr_set) { e = cudaFuncSetAttribute(gemm_tcgen05_pair_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return WK_ERR_CUDA; } attr_set = true; }
            e = cudaLaunchKernelEx(&cfg, gemm_tcgen05_pair_kernel<__half>, tmA, tmBh, p);
        } else {
            static bool attr_set = false;
            if (!attr_set) { e = cudaFuncSetAttribute(gemm_tcgen05_pair_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return WK_ERR_CUDA; } attr_set = true; }
            e = cudaLaunchKernelEx(&cfg, gemm_tcgen05_pair_kernel<__nv_bfloat16>, tmA, tmBh, p);
        }
        count_launch();
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) { set_error("gemm_tcgen05 (pair) launch: %s", cudaGetErrorString(e)); return WK_ERR_CUDA; }
        return WK_OK;
    }
    int grid = p.work < num_sms ? p.work : num_sms;
    if (grid < 1) return WK_OK;
    cudaError_t e;
    if (d.in_dtype == WK_DTYPE_F16) {
        static bool attr_set = false;
        if (!attr_set) {
            e = cudaFuncSetAttribute(gemm_tcgen05_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return WK_ERR_CUDA; }
            attr_set = true;
        }
        e = launch_k(gemm_tcgen05_kernel<__half>, dim3(grid), dim3(kGemmThreads), smem_bytes, stream, d.pdl != 0 ? 16 : 0, tmA, tmB, p);
    } else {
        static bool attr_set = false;
        if (!attr_set) {
            e = cudaFuncSetAttribute(gemm_tcgen05_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return WK_ERR_CUDA; }
            attr_set = true;
        }
        e = launch_k(gemm_tcgen05_kernel<__nv_bfloat16>, dim3(grid), dim3(kGemmThreads), smem_bytes, stream, d.pdl != 0 ? 16 : 0, tmA, tmB, p);
    }
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("gemm_tcgen05 launch: %s", cudaGetErrorString(e));
        return WK_ERR_CUDA;
    }
    return WK_OK;
}

}  // namespace wk

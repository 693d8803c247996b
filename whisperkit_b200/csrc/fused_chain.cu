// fused_chain.cu - decoder phase chains: one persistent kernel runs a CHAIN of decoder phases that today are separate launches:
//     swap-AB split-K GEMM -> split-K reduce (+bias +residual +LayerNorm | +bias +GELU) -> GEMM -> ...
// with a grid-wide barrier between phases instead of a kernel boundary.  Motivation (profiles/r01_summary.md section 6): a decoder GEMM
// launch is ~5.2 us of fixed cost around ~0.9 us of weight streaming, 11 such launches per layer; inside one kernel TMEM, mbarriers and
// tensor maps are set up once, the WEIGHT tiles of the next GEMM phase are put in flight before the barrier (weights are static), and
// a phase boundary costs one barrier.  Chains per decoder layer (engine.cu, decoder_forward):
//     B: out-proj GEMM -> reduce+LN -> cross-Q GEMM                                                  (between self- and cross-attention)
//     C: cross-out GEMM -> reduce+LN -> FC1 -> reduce+GELU -> FC2 -> reduce+LN -> next layer's QKV GEMM  (between cross- and self-attention)
//
// Structure: grid = one CTA per SM (all co-resident), 320 threads.  GEMM phase: warp 0 = TMA producer, warp 1 = tcgen05 MMA issuer,
// warps 2-9 = epilogue (transposed f32 partial store), work item = (128-row weight tile, K split), one per CTA (single wave by the
// split rule).  Reduce phase: all 320 threads, CTA b reduces batch rows b, b + grid, ... in the fixed split order (deterministic).  Pipeline state
// (ring stage / parity, accumulator parity) lives in registers across phases.  Memory ordering at a phase boundary: every thread
// fences its generic-proxy global writes towards the async proxy (the next GEMM phase reads activations with TMA), then
// bar.sync + __threadfence + atomic arrive / acquire spin (the cooperative-groups grid.sync recipe).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.h"

#define WK_CHECK_STATUS(expr)             \
    do {                                  \
        wk_status _s = (expr);            \
        if (_s != WK_OK) return _s;       \
    } while (0)

namespace wk {

namespace {

constexpr int kM = 128, kK = 64, kUK = 16;
constexpr int kStageABytes = kM * kK * 2;
constexpr int kThreads = 320;
constexpr int kRingMax = 8;
constexpr int kMaxSplitsR = 20;

struct PhaseK {
    int kind;
    int map;                    // GEMM: index into the tensor-map arrays
    int n, kb_per_split, splits, tiles;
    int red_n, red_splits;      // reduce: row length and partial count
    const float* bias; const float* gamma; const float* beta;
    void* out16;
};
struct ChainK {
    int n_phases;
    PhaseK ph[kChainMaxPhases];
    float* partial; float* x;
    int B, Bp, d;
    uint32_t idesc;
    int stages, stage_b_bytes, tmem_cols, acc_stride;
    unsigned int* counters;
    unsigned int* reset;
};
struct ChainMaps {
    CUtensorMap a[kChainMaxGemms];
    CUtensorMap b[kChainMaxGemms];
};

__device__ __forceinline__ void fence_generic_to_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// grid-wide barrier: every CTA of the (fully co-resident) grid arrives once on `ctr` (zero before the launch).  Release / acquire at gpu
// scope through one elected thread per CTA, made cumulative over the CTA by the bar.sync on either side.  `tma_reads_next`: this CTA's
// generic-proxy stores of the phase are read by TMA (async proxy) in the next phase - the writers fence towards that proxy first.
__device__ __forceinline__ void grid_barrier(unsigned int* ctr, unsigned int n_ctas, bool tma_reads_next) {
    if (tma_reads_next) fence_generic_to_async_global();
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        if (ld_acquire_gpu(ctr) < n_ctas) {
            const unsigned long long t0 = globaltimer_ns();
            unsigned int spins = 0;
            while (ld_acquire_gpu(ctr) < n_ctas) {
                if ((++spins & 0xfffu) == 0 && globaltimer_ns() - t0 > kSpinLimitNs) {
                    printf("wkb200: grid barrier timed out (block %d: %u of %u CTAs arrived)\n", (int)blockIdx.x, ld_acquire_gpu(ctr), n_ctas);
                    __trap();
                }
            }
        }
    }
    __syncthreads();
}

template <typename T>
__device__ __forceinline__ void reduce_ln_row(const ChainK& p, const PhaseK& P, int b, float* scratch) {
    // the body of decoder_reduce_resid_ln_kernel (decoder_ops.cu): x[b] += bias + sum_s partial[s][b]; out16[b] = LN(x[b])
    const int tid = threadIdx.x, d = P.red_n, d4 = d >> 2, splits = P.red_splits;
    T* xn = reinterpret_cast<T*>(P.out16);
    float4 v[2];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int i4 = tid + k * kThreads;
        v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i4 < d4) {
            float4 pr[kMaxSplitsR];
#pragma unroll
            for (int sp = 0; sp < kMaxSplitsR; ++sp)
                if (sp < splits) pr[sp] = __ldcg(reinterpret_cast<const float4*>(p.partial + ((long long)sp * p.Bp + b) * d) + i4);
            float4 a = reinterpret_cast<const float4*>(p.x + (long long)b * d)[i4];
            if (P.bias) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(P.bias) + i4);
                a.x += bb.x; a.y += bb.y; a.z += bb.z; a.w += bb.w;
            }
#pragma unroll
            for (int sp = 0; sp < kMaxSplitsR; ++sp)
                if (sp < splits) { a.x += pr[sp].x; a.y += pr[sp].y; a.z += pr[sp].z; a.w += pr[sp].w; }
            v[k] = a;
            reinterpret_cast<float4*>(p.x + (long long)b * d)[i4] = a;
            s += a.x + a.y + a.z + a.w;
        }
    }
    // block_sum over 320 threads (10 warps)
    auto bsum = [&](float val) -> float {
        val = warp_sum(val);
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        __syncthreads();
        if (lane == 0) scratch[warp] = val;
        __syncthreads();
        float r = (lane < kThreads / 32) ? scratch[lane] : 0.f;
        return warp_sum(r);
    };
    const float mean = bsum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int i4 = tid + k * kThreads;
        if (i4 < d4) {
            const float a0 = v[k].x - mean, a1 = v[k].y - mean, a2 = v[k].z - mean, a3 = v[k].w - mean;
            q += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
        }
    }
    const float rstd = rsqrtf(bsum(q) / d + 1e-5f);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int i4 = tid + k * kThreads;
        if (i4 < d4) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(P.gamma) + i4), bb = __ldg(reinterpret_cast<const float4*>(P.beta) + i4);
            uint2 pk;
            pk.x = T16<T>::pack2((v[k].x - mean) * rstd * g.x + bb.x, (v[k].y - mean) * rstd * g.y + bb.y);
            pk.y = T16<T>::pack2((v[k].z - mean) * rstd * g.z + bb.z, (v[k].w - mean) * rstd * g.w + bb.w);
            reinterpret_cast<uint2*>(xn + (long long)b * d)[i4] = pk;
        }
    }
}

template <typename T>
__device__ __forceinline__ void reduce_gelu_all(const ChainK& p, const PhaseK& P) {
    // the body of decoder_reduce_bias_gelu_kernel spread over the whole grid: out16[b][i] = gelu(bias[i] + sum_s partial[s][b][i])
    const int n = P.red_n, splits = P.red_splits;
    T* out = reinterpret_cast<T*>(P.out16);
    const long long total4 = (long long)p.B * n / 4;
    for (long long q = (long long)blockIdx.x * kThreads + threadIdx.x; q < total4; q += (long long)gridDim.x * kThreads) {
        const long long idx = q * 4;
        const int b = (int)(idx / n), i = (int)(idx - (long long)b * n);
        float4 a = *reinterpret_cast<const float4*>(P.bias + i);
        for (int sp = 0; sp < splits; ++sp) {
            const float4 pp = __ldcg(reinterpret_cast<const float4*>(p.partial + ((long long)sp * p.Bp + b) * n + i));
            a.x += pp.x; a.y += pp.y; a.z += pp.z; a.w += pp.w;
        }
        uint2 pk;
        pk.x = T16<T>::pack2(gelu_erf(a.x), gelu_erf(a.y));
        pk.y = T16<T>::pack2(gelu_erf(a.z), gelu_erf(a.w));
        *reinterpret_cast<uint2*>(out + (long long)b * n + i) = pk;
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
decoder_chain_kernel(const __grid_constant__ ChainMaps maps, const ChainK p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = kStageABytes + p.stage_b_bytes;
    uint8_t* tail = smem + (size_t)p.stages * stage_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
    uint64_t* empty_bar = full_bar + kRingMax;
    uint64_t* tfull_bar = empty_bar + kRingMax;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* scratch = reinterpret_cast<float*>(tmem_slot + 4);   // 32 floats

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        for (int g = 0; g < kChainMaxGemms; ++g) { tma_prefetch_desc(&maps.a[g]); tma_prefetch_desc(&maps.b[g]); }
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8); }
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, p.tmem_cols); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // ---- pipeline state that survives the phases
    int p_stage = 0; uint32_t p_phase = 0;      // producer (warp 0, lane 0)
    int pre_stage0 = 0, pre_n = 0, pre_for = -1; // weight tiles already in flight for GEMM phase `pre_for`
    int m_stage = 0; uint32_t m_phase = 0;      // MMA issuer (warp 1)
    int acc_it = 0;                             // accumulator use count (MMA warp and epilogue warps keep their own copy in step)

    auto work_of = [&](const PhaseK& P, int* tile, int* split) -> bool {
        const int w = blockIdx.x;
        if (w >= P.tiles * P.splits) return false;
        *split = w % P.splits;
        *tile = w / P.splits;
        return true;
    };
    // producer thread only: put the weight (A) tiles of GEMM phase `g` in flight; legal before the data dependency on the previous
    // phase is resolved because weights never change
    auto prefetch_weights = [&](int g) {
        const PhaseK& P = p.ph[g];
        int tile, split;
        pre_for = g; pre_n = 0; pre_stage0 = p_stage;
        if (!work_of(P, &tile, &split)) return;
        const int n_pre = P.kb_per_split < p.stages ? P.kb_per_split : p.stages;
        for (int i = 0; i < n_pre; ++i) {
            mbar_wait_bounded(&empty_bar[p_stage], p_phase ^ 1);
            uint8_t* sa = smem + (size_t)p_stage * stage_bytes;
            mbar_expect_tx(&full_bar[p_stage], (uint32_t)stage_bytes);
            tma_load_2d(sa, &maps.a[P.map], &full_bar[p_stage], (split * P.kb_per_split + i) * kK, tile * kM);
            if (++p_stage == p.stages) { p_stage = 0; p_phase ^= 1; }
        }
        pre_n = n_pre;
    };

    // the first phase is always a GEMM: its weights go out before griddepcontrol.wait, everything else after
    if (warp == 0 && lane == 0 && p.ph[0].kind == 0) prefetch_weights(0);
    pdl_wait();
    // every kernel upstream of this one has completed: the barrier words of the sibling chain (last used before this launch) can be
    // re-armed here, which keeps memset nodes out of the step graph
    if (blockIdx.x == 0 && threadIdx.x < 8 && p.reset != nullptr) p.reset[threadIdx.x] = 0u;

    for (int ph = 0; ph < p.n_phases; ++ph) {
        const PhaseK& P = p.ph[ph];
        if (P.kind == 0) {
            int tile = 0, split = 0;
            const bool has = work_of(P, &tile, &split);
            if (has && warp == 0) {
                if (lane == 0) {
                    // ===================== TMA producer =====================
                    if (pre_for != ph) prefetch_weights(ph);   // (only if the previous phase could not prefetch)
                    const int kb0 = split * P.kb_per_split;
                    fence_generic_to_async_global();           // the activations were written with generic stores by other CTAs (acquired at the barrier)
                    for (int i = 0; i < pre_n; ++i) {          // activations for the weight tiles already in flight
                        const int st = (pre_stage0 + i) % p.stages;
                        tma_load_2d(smem + (size_t)st * stage_bytes + kStageABytes, &maps.b[P.map], &full_bar[st], (kb0 + i) * kK, 0);
                    }
                    for (int kb = kb0 + pre_n; kb < kb0 + P.kb_per_split; ++kb) {
                        mbar_wait_bounded(&empty_bar[p_stage], p_phase ^ 1);
                        uint8_t* sa = smem + (size_t)p_stage * stage_bytes;
                        mbar_expect_tx(&full_bar[p_stage], (uint32_t)stage_bytes);
                        tma_load_2d(sa, &maps.a[P.map], &full_bar[p_stage], kb * kK, tile * kM);
                        tma_load_2d(sa + kStageABytes, &maps.b[P.map], &full_bar[p_stage], kb * kK, 0);
                        if (++p_stage == p.stages) { p_stage = 0; p_phase ^= 1; }
                    }
                    pre_for = -1;
                }
            } else if (has && warp == 1) {
                // ===================== MMA issuer =====================
                const int acc = acc_it & 1;
                const uint32_t acc_phase = (acc_it >> 1) & 1;
                mbar_wait_bounded(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * p.acc_stride;
                for (int kb = 0; kb < P.kb_per_split; ++kb) {
                    mbar_wait_bounded(&full_bar[m_stage], m_phase);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t sa = smem_u32(smem + (size_t)m_stage * stage_bytes);
                        const uint64_t adesc = make_kmajor_sw128_desc(sa);
                        const uint64_t bdesc = make_kmajor_sw128_desc(sa + kStageABytes);
#pragma unroll
                        for (int k = 0; k < kK / kUK; ++k)
                            tc_mma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), p.idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        tc_commit(&empty_bar[m_stage]);
                        if (kb == P.kb_per_split - 1) tc_commit(&tfull_bar[acc]);
                    }
                    __syncwarp();
                    if (++m_stage == p.stages) { m_stage = 0; m_phase ^= 1; }
                }
            } else if (has && warp >= 2) {
                // ===================== epilogue: transposed f32 partial store [split][b][n] =====================
                const int quarter = warp & 3, csub = (warp - 2) >> 2;
                const int acc = acc_it & 1;
                const uint32_t acc_phase = (acc_it >> 1) & 1;
                const int row = tile * kM + quarter * 32 + lane;
                const bool row_ok = row < P.n;
                mbar_wait_bounded(&tfull_bar[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + acc * p.acc_stride + ((uint32_t)(quarter * 32) << 16);
                for (int c = csub * 32; c < p.Bp; c += 64) {
                    uint32_t r[32];
                    __syncwarp();
                    tmem_ld_32x32(taddr + c, r);
                    tmem_ld_wait();
                    if (row_ok) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (c + j < p.Bp) p.partial[((long long)split * p.Bp + c + j) * P.n + row] = __uint_as_float(r[j]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            }
            if (has) ++acc_it;   // every thread of a CTA that had work advances its accumulator parity in step
        } else {
            // the producer thread first puts the NEXT GEMM phase's weight tiles in flight (all ring stages are free: the grid barrier
            // after the previous GEMM phase implies its MMAs have retired), then joins the reduction
            if (warp == 0 && lane == 0 && ph + 1 < p.n_phases && p.ph[ph + 1].kind == 0) prefetch_weights(ph + 1);
            __syncwarp();
            if (P.kind == 1) {
                for (int b = blockIdx.x; b < p.B; b += gridDim.x) reduce_ln_row<T>(p, P, b, scratch);   // rows beyond one per CTA (beam search: up to 256 rows)
            } else {
                reduce_gelu_all<T>(p, P);
            }
        }
        if (ph + 1 < p.n_phases) grid_barrier(p.counters + ph, gridDim.x, P.kind != 0);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, p.tmem_cols); }
}

}  // namespace

wk_status decoder_chain(const ChainDesc& c, int num_sms, cudaStream_t stream) {
    if (c.n_phases < 1 || c.n_phases > kChainMaxPhases || c.Bp % 16 != 0 || c.Bp < 16 || c.Bp > 256 || c.ph[0].kind != 0) {
        set_error("decoder_chain: unsupported chain (phases %d, Bp %d)", c.n_phases, c.Bp);
        return WK_ERR_INVALID_ARGUMENT;
    }
    ChainMaps maps;
    ChainK p;
    memset(&maps, 0, sizeof(maps));
    memset(&p, 0, sizeof(p));
    p.n_phases = c.n_phases; p.partial = c.partial; p.x = c.x; p.B = c.B; p.Bp = c.Bp; p.d = c.d; p.counters = c.counters; p.reset = c.reset_counters;
    p.idesc = 0;
    {
        const uint32_t fmt = c.dtype == WK_DTYPE_F16 ? 0u : 1u;
        uint32_t id = 0;
        id |= 1u << 4; id |= fmt << 7; id |= fmt << 10; id |= (uint32_t)(c.Bp >> 3) << 17; id |= (uint32_t)(kM >> 4) << 24;
        p.idesc = id;
    }
    p.stage_b_bytes = c.Bp * kK * 2;
    if (p.stage_b_bytes % 1024) p.stage_b_bytes = (p.stage_b_bytes + 1023) / 1024 * 1024;
    int n_gemm = 0, max_kb = 1, last_gemm = -1;
    for (int i = 0; i < c.n_phases; ++i) {
        const ChainPhaseDesc& s = c.ph[i];
        PhaseK& k = p.ph[i];
        k.kind = s.kind;
        if (s.kind == 0) {
            if (n_gemm >= kChainMaxGemms || s.k % kK || s.splits < 1 || (s.k / kK) % s.splits) { set_error("decoder_chain: bad GEMM phase %d", i); return WK_ERR_INVALID_ARGUMENT; }
            k.map = n_gemm; k.n = s.n; k.splits = s.splits; k.kb_per_split = s.k / kK / s.splits; k.tiles = (s.n + kM - 1) / kM;
            if (k.tiles * k.splits > num_sms) { set_error("decoder_chain: GEMM phase %d needs %d CTAs (> %d SMs)", i, k.tiles * k.splits, num_sms); return WK_ERR_INVALID_ARGUMENT; }
            WK_CHECK_STATUS(make_tmap_2d(&maps.a[n_gemm], s.w, c.dtype, (uint64_t)s.k, (uint64_t)s.n, (uint64_t)s.k, kK, kM));
            WK_CHECK_STATUS(make_tmap_2d(&maps.b[n_gemm], s.act, c.dtype, (uint64_t)s.k, (uint64_t)c.Bp, (uint64_t)s.k, kK, (uint32_t)c.Bp));
            max_kb = std::max(max_kb, k.kb_per_split);
            last_gemm = i;
            ++n_gemm;
        } else {
            if (last_gemm != i - 1) { set_error("decoder_chain: reduce phase %d must follow a GEMM phase", i); return WK_ERR_INVALID_ARGUMENT; }
            k.red_n = p.ph[i - 1].n; k.red_splits = p.ph[i - 1].splits;
            k.bias = s.bias; k.gamma = s.gamma; k.beta = s.beta; k.out16 = s.out16;
            if (k.red_splits > kMaxSplitsR || (k.red_n & 3) || (s.kind == 1 && (k.red_n != c.d || c.d > 8 * kThreads))) {
                set_error("decoder_chain: unsupported reduce phase %d", i); return WK_ERR_INVALID_ARGUMENT;
            }
        }
    }
    // unused tensor-map slots must still be valid descriptors (they are prefetched): repeat the first pair
    for (int g = n_gemm; g < kChainMaxGemms; ++g) { maps.a[g] = maps.a[0]; maps.b[g] = maps.b[0]; }
    const int stage_bytes = kStageABytes + p.stage_b_bytes;
    p.stages = std::min(kRingMax, std::max(2, max_kb));
    while ((size_t)p.stages * stage_bytes + 2048 > 200 * 1024 && p.stages > 2) --p.stages;
    p.acc_stride = 32; while (p.acc_stride < c.Bp) p.acc_stride <<= 1;
    p.tmem_cols = std::max(32, 2 * p.acc_stride);
    const size_t smem = (size_t)p.stages * stage_bytes + 1024 + 1024;
    cudaError_t e;
    if (c.dtype == WK_DTYPE_F16) {
        static bool attr = false;
        if (!attr) { e = cudaFuncSetAttribute(decoder_chain_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024); if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(chain): %s", cudaGetErrorString(e)); return WK_ERR_CUDA; } attr = true; }
        e = launch_k(decoder_chain_kernel<__half>, dim3(num_sms), dim3(kThreads), smem, stream, c.pdl ? 16 : 0, maps, p);
    } else {
        static bool attr = false;
        if (!attr) { e = cudaFuncSetAttribute(decoder_chain_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024); if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(chain): %s", cudaGetErrorString(e)); return WK_ERR_CUDA; } attr = true; }
        e = launch_k(decoder_chain_kernel<__nv_bfloat16>, dim3(num_sms), dim3(kThreads), smem, stream, c.pdl ? 16 : 0, maps, p);
    }
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("decoder_chain launch: %s", cudaGetErrorString(e)); return WK_ERR_CUDA; }
    return WK_OK;
}

}  // namespace wk

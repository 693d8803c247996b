// Encoder-side kernels besides the GEMM: LayerNorm, non-causal multi-head attention (S = 1500, d_head = 64),
// layout conversion for host readback, weight initialisation.
// Reference counterpart: the inside of AudioEncoder.mlmodelc (Sources/WhisperKit/Core/AudioEncoder.swift:50-63).
#include <curand_kernel.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"

namespace wk {

// =====================================================================================================
// LayerNorm: one warp per row, row kept in registers (two-pass mean / variance in fp32, eps 1e-5)
// =====================================================================================================
template <int NV, typename OutT>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 OutT* __restrict__ out, long long rows, int d) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + warp;
    if (row >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + row * d);
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = xr[lane + 32 * i];
        s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    const float mean = warp_sum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
        q += a * a + b * b + c * c + e * e;
    }
    const float rstd = rsqrtf(warp_sum(q) / d + 1e-5f);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 g = __ldg(g4 + lane + 32 * i), bb = __ldg(b4 + lane + 32 * i);
        const float o0 = (v[i].x - mean) * rstd * g.x + bb.x, o1 = (v[i].y - mean) * rstd * g.y + bb.y;
        const float o2 = (v[i].z - mean) * rstd * g.z + bb.z, o3 = (v[i].w - mean) * rstd * g.w + bb.w;
        if constexpr (sizeof(OutT) == 4) {
            reinterpret_cast<float4*>(out + row * d)[lane + 32 * i] = make_float4(o0, o1, o2, o3);
        } else {
            uint2 pk;
            pk.x = T16<OutT>::pack2(o0, o1);
            pk.y = T16<OutT>::pack2(o2, o3);
            reinterpret_cast<uint2*>(out + row * d)[lane + 32 * i] = pk;
        }
    }
}

template <typename OutT>
static wk_status launch_ln(const float* x, const float* g, const float* b, OutT* out, int64_t rows, int d, cudaStream_t st) {
    const unsigned grid = (unsigned)((rows + 7) / 8);
    switch (d / 128) {
        case 1: layernorm_kernel<1, OutT><<<grid, 256, 0, st>>>(x, g, b, out, rows, d); break;
        case 2: layernorm_kernel<2, OutT><<<grid, 256, 0, st>>>(x, g, b, out, rows, d); break;
        case 3: layernorm_kernel<3, OutT><<<grid, 256, 0, st>>>(x, g, b, out, rows, d); break;
        case 4: layernorm_kernel<4, OutT><<<grid, 256, 0, st>>>(x, g, b, out, rows, d); break;
        case 6: layernorm_kernel<6, OutT><<<grid, 256, 0, st>>>(x, g, b, out, rows, d); break;
        case 8: layernorm_kernel<8, OutT><<<grid, 256, 0, st>>>(x, g, b, out, rows, d); break;
        case 10: layernorm_kernel<10, OutT><<<grid, 256, 0, st>>>(x, g, b, out, rows, d); break;
        default: set_error("layernorm: unsupported d_model %d", d); return WK_ERR_INVALID_ARGUMENT;
    }
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("layernorm launch: %s", cudaGetErrorString(e)); return WK_ERR_CUDA; }
    return WK_OK;
}

wk_status layernorm_f32_to_16(const float* x, const float* gamma, const float* beta, void* out, int64_t rows, int d, int dtype,
                              cudaStream_t stream) {
    if (d % 128 != 0) { set_error("layernorm: d_model %d not a multiple of 128", d); return WK_ERR_INVALID_ARGUMENT; }
    if (dtype == WK_DTYPE_F16) return launch_ln<__half>(x, gamma, beta, (__half*)out, rows, d, stream);
    return launch_ln<__nv_bfloat16>(x, gamma, beta, (__nv_bfloat16*)out, rows, d, stream);
}
wk_status layernorm_f32_to_f32(const float* x, const float* gamma, const float* beta, float* out, int64_t rows, int d,
                               cudaStream_t stream) {
    if (d % 128 != 0) { set_error("layernorm: d_model %d not a multiple of 128", d); return WK_ERR_INVALID_ARGUMENT; }
    return launch_ln<float>(x, gamma, beta, out, rows, d, stream);
}

// =====================================================================================================
// Encoder attention (round 1: mma.sync m16n8k16 flash attention; tcgen05 version is the next kernel to write)
// qkv: [B*T, 3*dm] 16-bit (q | k | v, head h at columns h*64); out: [B*T, dm]
// CTA = 64 query rows of one (batch, head); 4 warps x 16 rows; KV tiles of 64 keys, cp.async double buffer
// =====================================================================================================
template <typename T> struct MmaOp;
template <> struct MmaOp<__nv_bfloat16> {
    __device__ __forceinline__ static void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
};
template <> struct MmaOp<__half> {
    __device__ __forceinline__ static void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
};
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

static constexpr int kAttnBM = 64, kAttnBN = 64, kAttnD = 64, kAttnThreads = 128;

// smem tile: 64 rows x 128 B, 16-byte chunks XOR-swizzled by (row & 7)
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

template <typename T>
__device__ __forceinline__ void attn_load_tile(uint8_t* smem_tile, const T* gbase, long long ld, int row0, int nrows_valid,
                                               int tid) {
    // 64 rows x 8 chunks = 512 chunks, 4 per thread
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = tid + i * kAttnThreads;
        const int row = idx >> 3, chunk = idx & 7;
        const bool ok = (row0 + row) < nrows_valid;
        const T* src = gbase + (long long)(ok ? (row0 + row) : 0) * ld + chunk * 8;
        cp_async16(smem_tile + tile_off(row, chunk), src, ok);
    }
}

template <typename T>
__global__ void __launch_bounds__(kAttnThreads)
encoder_attention_kernel(const T* __restrict__ qkv, T* __restrict__ out, int Tlen, int H, int dm, float scale_log2e) {
    __shared__ __align__(128) uint8_t sQ[kAttnBM * 128];
    __shared__ __align__(128) uint8_t sK[2][kAttnBN * 128];
    __shared__ __align__(128) uint8_t sV[2][kAttnBN * 128];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int bh = blockIdx.y, b = bh / H, h = bh % H;
    const int q0 = blockIdx.x * kAttnBM;
    const long long ld = 3LL * dm;
    const T* qb = qkv + (long long)b * Tlen * ld + h * 64;
    const T* kb = qb + dm;
    const T* vb = qb + 2 * dm;

    attn_load_tile<T>(sQ, qb, ld, q0, Tlen, tid);
    attn_load_tile<T>(sK[0], kb, ld, 0, Tlen, tid);
    attn_load_tile<T>(sV[0], vb, ld, 0, Tlen, tid);
    cp_async_commit();

    const int n_tiles = (Tlen + kAttnBN - 1) / kAttnBN;
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    uint32_t qf[4][4];  // Q fragments for the 4 k-steps (d = 64)

    for (int j = 0; j < n_tiles; ++j) {
        const int buf = j & 1;
        if (j + 1 < n_tiles) {
            attn_load_tile<T>(sK[buf ^ 1], kb, ld, (j + 1) * kAttnBN, Tlen, tid);
            attn_load_tile<T>(sV[buf ^ 1], vb, ld, (j + 1) * kAttnBN, Tlen, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (j == 0) {
            // A fragments of Q: rows warp*16 + (lane%8) + ((lane/8)%2)*8, chunk = 2*ks + lane/16
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                ldsm_x4(qf[ks], smem_u32(sQ) + tile_off(row, 2 * ks + (lane >> 4)));
            }
        }
        // ---- S = Q K^T (16 x 64 per warp)
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
        const uint32_t kbase = smem_u32(sK[buf]);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {  // pairs of 8-key n-tiles
                uint32_t kf[4];
                const int row = np * 16 + (lane & 7) + (lane >> 4) * 8;
                ldsm_x4(kf, kbase + tile_off(row, 2 * ks + ((lane >> 3) & 1)));
                MmaOp<T>::mma(s[2 * np], qf[ks], kf[0], kf[1]);
                MmaOp<T>::mma(s[2 * np + 1], qf[ks], kf[2], kf[3]);
            }
        }
        // ---- mask keys beyond Tlen (only the last tile)
        const int key0 = j * kAttnBN;
        if (key0 + kAttnBN > Tlen) {
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const int c = key0 + nt * 8 + 2 * t4;
                if (c >= Tlen) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
                if (c + 1 >= Tlen) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
            }
        }
        // ---- online softmax (rows g and g+8 of this warp's 16)
        float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        }
        float corr[2], msc[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            corr[r] = exp2f((m_run[r] - mx[r]) * scale_log2e);  // m_run = -inf on first tile -> 0
            m_run[r] = mx[r];
            msc[r] = mx[r] * scale_log2e;
            l_run[r] *= corr[r];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1]; }
        uint32_t pf[4][4];  // P as A fragments for the 4 key k-steps
        float ls[2] = {0.f, 0.f};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float p0 = exp2f(s[nt][0] * scale_log2e - msc[0]);
            const float p1 = exp2f(s[nt][1] * scale_log2e - msc[0]);
            const float p2 = exp2f(s[nt][2] * scale_log2e - msc[1]);
            const float p3 = exp2f(s[nt][3] * scale_log2e - msc[1]);
            ls[0] += p0 + p1;
            ls[1] += p2 + p3;
            pf[nt >> 1][(nt & 1) * 2 + 0] = T16<T>::pack2(p0, p1);
            pf[nt >> 1][(nt & 1) * 2 + 1] = T16<T>::pack2(p2, p3);
        }
        l_run[0] += ls[0];
        l_run[1] += ls[1];
        // ---- O += P V
        const uint32_t vbase = smem_u32(sV[buf]);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {      // 16 keys per step
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {  // pairs of 8-wide d n-tiles
                uint32_t vf[4];
                const int row = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                ldsm_x4_trans(vf, vbase + tile_off(row, 2 * dp + (lane >> 4)));
                MmaOp<T>::mma(o[2 * dp], pf[ks], vf[0], vf[1]);
                MmaOp<T>::mma(o[2 * dp + 1], pf[ks], vf[2], vf[3]);
            }
        }
        __syncthreads();  // everyone done with buf before it is refilled two iterations later
    }
    // ---- finalise: quad-reduce row sums, normalise, store
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
    T* ob = out + (long long)b * Tlen * dm + h * 64;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const int c = nt * 8 + 2 * t4;
        if (r0 < Tlen) *reinterpret_cast<uint32_t*>(ob + (long long)r0 * dm + c) = T16<T>::pack2(o[nt][0] * inv0, o[nt][1] * inv0);
        if (r1 < Tlen) *reinterpret_cast<uint32_t*>(ob + (long long)r1 * dm + c) = T16<T>::pack2(o[nt][2] * inv1, o[nt][3] * inv1);
    }
}

wk_status encoder_attention(const void* qkv, void* out, int B, int T, int n_heads, int dtype, cudaStream_t stream) {
    static int legacy = -1;
    if (legacy < 0) { const char* e = getenv("WKB200_ATTN"); legacy = (e && strcmp(e, "legacy") == 0) ? 1 : 0; }
    if (!legacy) return encoder_attention_tcgen05(qkv, out, B, T, n_heads, dtype, stream);
    const int dm = n_heads * 64;
    dim3 grid((T + kAttnBM - 1) / kAttnBM, B * n_heads);
    const float scale_log2e = 0.125f * 1.4426950408889634f;
    if (dtype == WK_DTYPE_F16)
        encoder_attention_kernel<__half><<<grid, kAttnThreads, 0, stream>>>((const __half*)qkv, (__half*)out, T, n_heads, dm, scale_log2e);
    else
        encoder_attention_kernel<__nv_bfloat16><<<grid, kAttnThreads, 0, stream>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, T, n_heads, dm, scale_log2e);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("encoder_attention launch: %s", cudaGetErrorString(e)); return WK_ERR_CUDA; }
    return WK_OK;
}

// =====================================================================================================
// Helpers: host-layout readback, random init, dtype conversion
// =====================================================================================================
// dst[b][c][r] (f32) = src[b][row_off + r][c]   (time-major device layout -> channel-major reference layout)
template <typename T>
__global__ void transpose_to_host_kernel(const T* __restrict__ src, float* __restrict__ dst, long long rows, long long cols,
                                         long long src_rows_alloc, long long row_off, long long src_ld) {
    __shared__ float tile[32][33];
    const long long b = blockIdx.z;
    const long long r0 = (long long)blockIdx.x * 32, c0 = (long long)blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const long long r = r0 + i, c = c0 + threadIdx.x;
        float v = 0.f;
        if (r < rows && c < cols) {
            const T* p = src + (b * src_rows_alloc + row_off + r) * src_ld + c;
            if constexpr (sizeof(T) == 4) v = *p; else v = T16<T>::to_f(*p);
        }
        tile[i][threadIdx.x] = v;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const long long c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) dst[(b * cols + c) * rows + r] = tile[threadIdx.x][i];
    }
}

wk_status transpose_to_host_layout(const void* src, float* dst, int64_t B, int64_t rows, int64_t cols, int64_t src_rows_alloc,
                                   int64_t src_row_off, int64_t src_ld, int dtype, cudaStream_t stream) {
    dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32), (unsigned)B), block(32, 8);
    if (dtype == WK_DTYPE_F32) transpose_to_host_kernel<float><<<grid, block, 0, stream>>>((const float*)src, dst, rows, cols, src_rows_alloc, src_row_off, src_ld);
    else if (dtype == WK_DTYPE_F16) transpose_to_host_kernel<__half><<<grid, block, 0, stream>>>((const __half*)src, dst, rows, cols, src_rows_alloc, src_row_off, src_ld);
    else transpose_to_host_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>((const __nv_bfloat16*)src, dst, rows, cols, src_rows_alloc, src_row_off, src_ld);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("transpose launch: %s", cudaGetErrorString(e)); return WK_ERR_CUDA; }
    return WK_OK;
}

template <typename T>
__global__ void fill_random_kernel(T* dst, long long n, unsigned long long seed, float std, float mean) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long base = i * 4;
    if (base >= n) return;
    curandStatePhilox4_32_10_t st;
    curand_init(seed, (unsigned long long)i, 0, &st);
    const float4 r = curand_normal4(&st);
    const float v[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (base + k < n) {
            const float f = mean + std * v[k];
            if constexpr (sizeof(T) == 4) dst[base + k] = f; else dst[base + k] = T16<T>::from_f(f);
        }
}

wk_status fill_random_16(void* dst, int64_t n, uint64_t seed, float std, float mean, int dtype, cudaStream_t stream) {
    const unsigned grid = (unsigned)((n / 4 + 256) / 256);
    if (dtype == WK_DTYPE_F16) fill_random_kernel<__half><<<grid, 256, 0, stream>>>((__half*)dst, n, seed, std, mean);
    else fill_random_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((__nv_bfloat16*)dst, n, seed, std, mean);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("fill_random launch: %s", cudaGetErrorString(e)); return WK_ERR_CUDA; }
    return WK_OK;
}
wk_status fill_random_f32(float* dst, int64_t n, uint64_t seed, float std, float mean, cudaStream_t stream) {
    const unsigned grid = (unsigned)((n / 4 + 256) / 256);
    fill_random_kernel<float><<<grid, 256, 0, stream>>>(dst, n, seed, std, mean);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("fill_random launch: %s", cudaGetErrorString(e)); return WK_ERR_CUDA; }
    return WK_OK;
}

template <typename S, typename D>
__global__ void convert_kernel(const S* __restrict__ src, D* __restrict__ dst, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v;
    if constexpr (sizeof(S) == 4) v = src[i]; else v = T16<S>::to_f(src[i]);
    if constexpr (sizeof(D) == 4) dst[i] = v; else dst[i] = T16<D>::from_f(v);
}

wk_status convert_to_16(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, cudaStream_t stream) {
    const unsigned grid = (unsigned)((n + 255) / 256);
#define WK_CVT(S, D) convert_kernel<S, D><<<grid, 256, 0, stream>>>((const S*)src, (D*)dst, n)
    if (src_dtype == WK_DTYPE_F32 && dst_dtype == WK_DTYPE_BF16) WK_CVT(float, __nv_bfloat16);
    else if (src_dtype == WK_DTYPE_F32 && dst_dtype == WK_DTYPE_F16) WK_CVT(float, __half);
    else if (src_dtype == WK_DTYPE_F32 && dst_dtype == WK_DTYPE_F32) WK_CVT(float, float);
    else if (src_dtype == WK_DTYPE_BF16 && dst_dtype == WK_DTYPE_BF16) WK_CVT(__nv_bfloat16, __nv_bfloat16);
    else if (src_dtype == WK_DTYPE_BF16 && dst_dtype == WK_DTYPE_F16) WK_CVT(__nv_bfloat16, __half);
    else if (src_dtype == WK_DTYPE_BF16 && dst_dtype == WK_DTYPE_F32) WK_CVT(__nv_bfloat16, float);
    else if (src_dtype == WK_DTYPE_F16 && dst_dtype == WK_DTYPE_F16) WK_CVT(__half, __half);
    else if (src_dtype == WK_DTYPE_F16 && dst_dtype == WK_DTYPE_BF16) WK_CVT(__half, __nv_bfloat16);
    else if (src_dtype == WK_DTYPE_F16 && dst_dtype == WK_DTYPE_F32) WK_CVT(__half, float);
    else { set_error("convert: unsupported dtype pair %d -> %d", src_dtype, dst_dtype); return WK_ERR_INVALID_ARGUMENT; }
#undef WK_CVT
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("convert launch: %s", cudaGetErrorString(e)); return WK_ERR_CUDA; }
    return WK_OK;
}

}  // namespace wk

// Internal object layouts of libwkb200 shared by engine.cu (model, weights, mel + encoder schedule, kernel hooks) and
// session.cu (decode sessions, the device-resident token loop, the window scheduler).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "kernels.h"

namespace wk {

constexpr int kKvMaxLen = 224;        // Constants.maxTokenContext (Models.swift:1334)
constexpr int kWindowSamples = 480000;  // Constants.defaultWindowSamples (Models.swift:1457)

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

#define WK_CHECK(expr)                    \
    do {                                  \
        wk_status _s = (expr);            \
        if (_s != WK_OK) return _s;       \
    } while (0)

template <typename T>
static wk_status dmalloc(T** p, size_t n, bool zero = true) {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T));
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e));
        return WK_ERR_CUDA;
    }
    if (zero) {
        e = cudaMemset(*p, 0, n * sizeof(T));
        if (e != cudaSuccess) { set_error("cudaMemset failed: %s", cudaGetErrorString(e)); return WK_ERR_CUDA; }
    }
    return WK_OK;
}
static inline wk_status alloc16(void** p, size_t n) {
    uint16_t* q = nullptr;
    wk_status s = dmalloc(&q, n);
    *p = q;
    return s;
}

struct LayerNormW { float* g = nullptr; float* b = nullptr; };
struct EncLayer {
    LayerNormW ln1, ln2;
    void* wqkv = nullptr; float* bqkv = nullptr;  // [3d, d]
    void* wo = nullptr; float* bo = nullptr;
    void* w1 = nullptr; float* b1 = nullptr;      // [4d, d]
    void* w2 = nullptr; float* b2 = nullptr;      // [d, 4d]
};
struct DecLayer {
    LayerNormW ln1, lnx, ln3;
    void* wqkv = nullptr; float* bq = nullptr; float* bv = nullptr;
    void* wo = nullptr; float* bo = nullptr;
    void* wcq = nullptr; float* bcq = nullptr;
    void* wco = nullptr; float* bco = nullptr;
    void* w1 = nullptr; float* b1 = nullptr;
    void* w2 = nullptr; float* b2 = nullptr;
};

// Mel + encoder activations for up to max_batch windows.  The model keeps one for the piecewise API (wk_mel / wk_encode, serialised by
// wk_model::api_mu); every session allocates its own on first use, so sessions on different host threads encode concurrently.
struct EncWorkspace {
    int max_batch = 0;
    float* pcm_dev = nullptr; int32_t* nvalid_dev = nullptr; int32_t* gmax = nullptr;
    void* mel = nullptr;       // f16 [Bm][3002][128]
    void* h1 = nullptr;        // f16 [Bm][3002][d]
    float* x = nullptr;        // f32 [Bm*1500][d]
    void* xn = nullptr; void* qkv = nullptr; void* attn = nullptr; void* ffn = nullptr;
    void* enc_out = nullptr;   // 16-bit [Bm*1500][d]
};

}  // namespace wk

struct wk_tensor {
    void* data;
    int kind;      // 0 = mel [B,3002,128] f16 ; 1 = encoder output [B*1500, d] model dtype
    int dtype;
    int64_t batch;
    wk_model* owner;
    std::vector<cudaEvent_t> events;   // [0] producer done (model stream), then one per reader on another stream (guarded by owner->api_mu):
                                       // readers wait on them, wk_tensor_free orders the release after them
};

struct wk_model {
    wk_model_config cfg;
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;   // the piecewise API (wk_mel, wk_encode, wk_filter_sample, readbacks) is enqueued here
    std::mutex api_mu;               // ... one host thread at a time: the handle itself is immutable once finalized
    bool finalized = false;
    int esz = 2;
    // weights
    void* conv1_w = nullptr; float* conv1_b = nullptr;   // f16 [d][3][128]
    void* conv2_w = nullptr; float* conv2_b = nullptr;   // f16 [d][3][d]
    float* enc_pos = nullptr;                            // [1500][d]
    std::vector<wk::EncLayer> enc;
    wk::LayerNormW enc_ln;
    void* emb = nullptr;                                 // [V][d]
    float* dec_pos = nullptr;                            // [448][d]
    std::vector<wk::DecLayer> dec;
    wk::LayerNormW dec_ln;
    void* wckv = nullptr; float* bckv = nullptr;         // [2L*d][d], [2L*d]
    wk::MelTables* mel_tables = nullptr;
    wk::EncWorkspace ws;                                 // allocated on the first wk_mel / wk_encode
    // alignment heads (word timestamps): per decoder layer a head bit mask and the first scratch slot of the layer
    std::vector<uint32_t> align_mask; std::vector<int> align_base; int n_align_slots = 0; int has_alignment_heads = 0;
    float timings[6] = {0, 0, 0, 0, 0, 0};
    cudaEvent_t ev[8];
    std::atomic<int> live_sessions{0};
};

namespace wk {

wk_status enc_ws_ensure(wk_model* m, EncWorkspace* ws, int max_batch);
void enc_ws_free(EncWorkspace* ws);
// PCM rows (host or device) -> staged in ws->pcm_dev when needed -> log-mel into mel_out ([n][3002][128] f16), all on `stream`
wk_status mel_run(wk_model* m, EncWorkspace* ws, const float* pcm, int64_t n, int64_t stride, const int32_t* samples_per_window_host,
                  void* mel_out, cudaStream_t stream);
// conv stem + encoder layers over B windows of `mel` into enc_out ([B*1500][d], model dtype), activations in ws, on `stream`
wk_status encode_chunk(wk_model* m, EncWorkspace* ws, const void* mel, int B, void* enc_out, cudaStream_t stream);
GemmDesc plain_gemm(const void* a, int64_t M, int K, const void* w, int N, int dtype, int mode, void* out, int64_t ld_out,
                    const float* bias, int gelu);
int choose_splits(int tiles, int total_kb, int num_sms);
size_t esize(int dtype);
long long launch_counter_load();
void launch_counter_sub(long long n);

}  // namespace wk

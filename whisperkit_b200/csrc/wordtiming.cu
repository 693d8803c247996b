// wordtiming.cu - host side of word timestamps above the device alignment export.  C++ because the
// reference's is compiled Swift; pure host code, callable without a GPU.
//   dynamicTimeWarping / backtrace            Sources/WhisperKit/Core/Text/SegmentSeeker.swift:195-276
//   mergePunctuations                         SegmentSeeker.swift:278-338
//   findAlignment                             SegmentSeeker.swift:340-408
//   addWordTimestamps                         SegmentSeeker.swift:410-496
//   duration constraints / truncation         SegmentSeeker.swift:498-526
//   updateSegmentsWithWordTimings             SegmentSeeker.swift:528-659
// Swift `Float` arithmetic is kept in `float` expression by expression; Float.rounded(2) = roundf(x * 100) / 100
// (Sources/ArgmaxCore/FoundationExtensions.swift:10-13).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include <cuda_fp16.h>

#include "kernels.h"

using namespace wk;

#define WK_CHECK(expr)                    \
    do {                                  \
        wk_status _s = (expr);            \
        if (_s != WK_OK) return _s;       \
    } while (0)

namespace {

constexpr int kSampleRate = 16000;
constexpr float kSecondsPerTimeToken = 0.02f;
const char* kDefaultPrepend = "\"'\xe2\x80\x9c\xc2\xa1\xc2\xbf([{-";                                   // "'“¡¿([{-   (Models.swift:1459)
const char* kDefaultAppend = "\"'.\xe3\x80\x82,\xef\xbc\x8c!\xef\xbc\x81?\xef\xbc\x9f:\xef\xbc\x9a\xe2\x80\x9d)]}\xe3\x80\x81";  // "'.。,，!！?？:：”)]}、 (:1460)
const char* kSentenceEnd[] = {".", "\xe3\x80\x82", "!", "\xef\xbc\x81", "?", "\xef\xbc\x9f"};        // . 。 ! ！ ? ？

struct Word {
    std::string word;
    std::vector<int32_t> tokens;
    float start = 0.f, end = 0.f, probability = 0.f;
    int segment = -1;
    float duration() const { return end - start; }
};

// length in bytes of the CharacterSet.whitespaces scalar at p (Zs + TAB), 0 if none
int ws_len(const char* p, size_t n) {
    const unsigned char* u = (const unsigned char*)p;
    if (n >= 1 && (u[0] == 0x20 || u[0] == 0x09)) return 1;
    if (n >= 2 && u[0] == 0xC2 && u[1] == 0xA0) return 2;                                            // U+00A0
    if (n >= 3 && u[0] == 0xE1 && u[1] == 0x9A && u[2] == 0x80) return 3;                             // U+1680
    if (n >= 3 && u[0] == 0xE2 && u[1] == 0x80 && ((u[2] >= 0x80 && u[2] <= 0x8A) || u[2] == 0xAF)) return 3;  // U+2000-200A, U+202F
    if (n >= 3 && u[0] == 0xE2 && u[1] == 0x81 && u[2] == 0x9F) return 3;                             // U+205F
    if (n >= 3 && u[0] == 0xE3 && u[1] == 0x80 && u[2] == 0x80) return 3;                             // U+3000
    return 0;
}

std::string trim_ws(const std::string& s) {
    size_t a = 0, b = s.size();
    for (;;) { const int k = ws_len(s.data() + a, b - a); if (!k) break; a += k; }
    for (;;) {
        bool cut = false;
        for (int k = 1; k <= 3 && !cut; ++k)
            if (b - a >= (size_t)k && ws_len(s.data() + b - k, k) == k) { b -= k; cut = true; }
        if (!cut) break;
    }
    return s.substr(a, b - a);
}

// String.contains(_ other: String): substring search (valid on UTF-8 bytes); the empty string is contained
bool contains(const std::string& hay, const std::string& needle) { return hay.find(needle) != std::string::npos; }
bool is_sentence_end(const std::string& w) {
    for (const char* m : kSentenceEnd) if (w == m) return true;
    return false;
}
float rounded2(float x) { return roundf(x * 100.f) / 100.f; }

std::vector<Word> from_c(const wk_word* w, int n) {
    std::vector<Word> v((size_t)std::max(n, 0));
    for (int i = 0; i < n; ++i) {
        v[i].word = w[i].word ? w[i].word : "";
        if (w[i].tokens && w[i].n_tokens > 0) v[i].tokens.assign(w[i].tokens, w[i].tokens + w[i].n_tokens);
        v[i].start = w[i].start; v[i].end = w[i].end; v[i].probability = w[i].probability; v[i].segment = w[i].segment;
    }
    return v;
}

template <typename T> double at(const void* m, int64_t i);
template <> double at<float>(const void* m, int64_t i) { return (double)((const float*)m)[i]; }
template <> double at<__half>(const void* m, int64_t i) { return (double)__half2float(((const __half*)m)[i]); }

template <typename T>
void dtw(const void* matrix, int rows, int cols, int64_t ld, std::vector<int32_t>& ti, std::vector<int32_t>& tj) {
    // cost only needs the previous row; the trace is kept whole (1 byte per cell)
    const double inf = INFINITY;
    std::vector<double> prev((size_t)cols + 1, inf), cur((size_t)cols + 1, inf);
    std::vector<int8_t> trace((size_t)(rows + 1) * (cols + 1), (int8_t)-1);
    const size_t W = (size_t)cols + 1;
    prev[0] = 0.0;
    for (int c = 1; c <= cols; ++c) trace[c] = 2;
    for (int r = 1; r <= rows; ++r) trace[(size_t)r * W] = 1;
    for (int r = 1; r <= rows; ++r) {
        cur[0] = inf;
        int8_t* tr = trace.data() + (size_t)r * W;
        for (int c = 1; c <= cols; ++c) {
            const double v = -at<T>(matrix, (int64_t)(r - 1) * ld + (c - 1));
            const double c0 = prev[c - 1] + v, c1 = prev[c] + v, c2 = cur[c - 1] + v;
            if (c0 < c1 && c0 < c2) { cur[c] = c0; tr[c] = 0; }
            else if (c1 < c0 && c1 < c2) { cur[c] = c1; tr[c] = 1; }
            else { cur[c] = c2; tr[c] = 2; }
        }
        prev.swap(cur);
    }
    int i = rows, j = cols;
    ti.clear(); tj.clear();
    while (i > 0 || j > 0) {
        ti.push_back(i - 1); tj.push_back(j - 1);
        const int8_t t = trace[(size_t)i * W + j];
        if (t == 0) { --i; --j; } else if (t == 1) --i; else if (t == 2) --j; else break;
    }
    std::reverse(ti.begin(), ti.end());
    std::reverse(tj.begin(), tj.end());
}

wk_status run_dtw(const void* matrix, int dtype, int rows, int cols, int64_t ld, std::vector<int32_t>& ti, std::vector<int32_t>& tj) {
    if (!matrix || rows < 1 || cols < 1 || ld < cols) { set_error("dynamicTimeWarping: invalid alignment matrix shape [%d x %d]", rows, cols); return WK_ERR_INVALID_ARGUMENT; }
    if (dtype == WK_DTYPE_F32) dtw<float>(matrix, rows, cols, ld, ti, tj);
    else if (dtype == WK_DTYPE_F16) dtw<__half>(matrix, rows, cols, ld, ti, tj);
    else { set_error("dynamicTimeWarping: dtype %d unsupported", dtype); return WK_ERR_INVALID_ARGUMENT; }
    return WK_OK;
}

std::vector<Word> merge_punctuations(const std::vector<Word>& al, const std::string& prepended, const std::string& appended) {
    if (al.empty()) return {};
    std::vector<Word> pre, app;
    if (!contains(prepended, trim_ws(al[0].word))) pre.push_back(al[0]);
    for (size_t i = 1; i < al.size(); ++i) {
        Word cur = al[i];
        const Word& prev = al[i - 1];
        if (!prev.word.empty() && ws_len(prev.word.data(), prev.word.size()) > 0 && contains(prepended, trim_ws(prev.word))) {
            cur.word = prev.word + cur.word;
            std::vector<int32_t> t = prev.tokens;
            t.insert(t.end(), cur.tokens.begin(), cur.tokens.end());
            cur.tokens.swap(t);
            if (pre.empty()) pre.push_back(cur); else pre.back() = cur;
        } else {
            pre.push_back(cur);
        }
    }
    if (!pre.empty()) app.push_back(pre[0]);
    for (size_t i = 1; i < pre.size(); ++i) {
        const Word& cur = pre[i];
        Word prev = pre[i - 1];
        const bool prev_ends_space = !prev.word.empty() && prev.word.back() == ' ';
        if (!prev_ends_space && contains(appended, trim_ws(cur.word))) {
            prev.word += cur.word;
            prev.tokens.insert(prev.tokens.end(), cur.tokens.begin(), cur.tokens.end());
            app.back() = prev;
        } else {
            app.push_back(cur);
        }
    }
    std::vector<Word> out;
    for (const Word& w : app)
        if (!w.word.empty() && !contains(appended, w.word) && !contains(prepended, w.word)) out.push_back(w);
    return out;
}

wk_status find_alignment(const std::vector<Word>& words, const void* matrix, int dtype, int rows, int cols, int64_t ld, const float* lps, int n_lps,
                         std::vector<Word>& out) {
    std::vector<int32_t> ti, tj;
    WK_CHECK(run_dtw(matrix, dtype, rows, cols, ld, ti, tj));
    out.clear();
    if (words.size() <= 1) return WK_OK;
    std::vector<float> start_times{0.f}, end_times;
    int cur = ti.empty() ? 0 : ti[0];
    for (size_t k = 0; k < ti.size(); ++k)
        if (ti[k] != cur) {
            cur = ti[k];
            const float t = (float)tj[k] * kSecondsPerTimeToken;
            start_times.push_back(t);
            end_times.push_back(t);
        }
    end_times.push_back((float)(tj.empty() ? 1500 : tj.back()) * kSecondsPerTimeToken);
    size_t ci = 0;
    for (const Word& w : words) {
        if (w.tokens.empty()) { set_error("findAlignment: word without tokens"); return WK_ERR_INVALID_ARGUMENT; }
        const size_t s0 = ci;
        if (ci >= start_times.size()) { set_error("findAlignment: %zu word tokens but %d alignment rows", ci + 1, rows); return WK_ERR_INVALID_ARGUMENT; }
        const float st = start_times[ci];
        ci += w.tokens.size() - 1;
        if (ci >= end_times.size() || (int)ci >= n_lps) { set_error("findAlignment: word tokens exceed alignment rows / log probs"); return WK_ERR_INVALID_ARGUMENT; }
        const float en = end_times[ci];
        ++ci;
        float acc = 0.f;
        for (size_t k = s0; k < ci; ++k) acc += lps[k];
        Word o = w;
        o.start = st; o.end = en; o.probability = expf(acc / (float)(ci - s0)); o.segment = -1;
        out.push_back(std::move(o));
    }
    return WK_OK;
}

void duration_constraints(const std::vector<Word>& al, float* median, float* max_duration) {
    std::vector<float> d;
    for (const Word& w : al) if (w.duration() > 0.f) d.push_back(w.duration());
    std::sort(d.begin(), d.end());
    const float med = d.empty() ? 0.f : d[d.size() / 2];
    *median = std::min(0.7f, med);
    *max_duration = *median * 2.f;
}

void truncate_long_words(std::vector<Word>& al, float max_duration) {
    for (size_t i = 1; i < al.size(); ++i)
        if (al[i].duration() > max_duration) {
            if (is_sentence_end(al[i].word)) al[i].end = al[i].start + max_duration;
            else if (is_sentence_end(al[i - 1].word)) al[i].start = al[i].end - max_duration;
        }
}

wk_status update_segments(wk_segment* segs, int n_segs, const int32_t* tokens, const std::vector<Word>& merged, int64_t seek, float last_speech,
                          float cmd, float max_duration, int special_begin, const wk_tokenizer_hooks* hooks, std::vector<Word>& out) {
    const float time_offset = (float)seek / (float)kSampleRate;
    size_t word_index = 0;
    out.clear();
    std::vector<float> seg_end_updated;
    for (int si = 0; si < n_segs; ++si) {
        wk_segment& seg = segs[si];
        const float seg_start = seg.start, seg_end = seg.end;
        int text_tokens = 0;
        for (int t = 0; t < seg.n_tokens; ++t) text_tokens += tokens[seg.token_offset + t] < special_begin;
        int saved = 0;
        std::vector<Word> wis;
        const size_t slice_begin = word_index;   // `for timing in mergedAlignment[wordIndex...] where saved < textTokens.count`: the where-clause skips, it does not stop
        for (size_t k = slice_begin; k < merged.size(); ++k) {
            if (!(saved < text_tokens)) continue;
            const Word& timing = merged[k];
            ++word_index;
            std::vector<int32_t> tt;
            for (int32_t t : timing.tokens) if (t < special_begin) tt.push_back(t);
            if (tt.empty()) continue;
            std::string word = timing.word;
            if (tt.size() < timing.tokens.size() && hooks && hooks->decode) {
                std::vector<char> buf(tt.size() * 64 + 64);
                const int32_t nb = hooks->decode(hooks->user, tt.data(), (int32_t)tt.size(), buf.data(), (int32_t)buf.size());
                if (nb < 0) { set_error("tokenizer decode hook failed (%d)", nb); return WK_ERR_TRANSCRIPTION_FAILED; }
                word.assign(buf.data(), (size_t)nb);
            }
            float start = rounded2(time_offset + timing.start);
            const float end = rounded2(time_offset + timing.end);
            if (end - start < cmd / 4.f) {
                if (!wis.empty()) {
                    const float prev_end = wis.back().end;
                    if (start > prev_end) {
                        const float space = start - prev_end;
                        start = rounded2(start - std::min(space, cmd / 2.f));
                    }
                } else if (si > 0 && (int)seg_end_updated.size() > si - 1 && start > seg_end_updated[si - 1]) {
                    const float space = start - seg_end_updated[si - 1];
                    start = rounded2(start - std::min(space, cmd / 2.f));
                }
            }
            Word w;
            w.word = word; w.tokens = tt; w.start = start; w.end = end; w.probability = rounded2(timing.probability); w.segment = si;
            wis.push_back(std::move(w));
            saved += (int)tt.size();
        }
        if (!wis.empty()) {
            const Word first = wis[0];
            const float pause = first.end - last_speech;
            const bool first_too_long = first.duration() > max_duration;
            const bool both_too_long = wis.size() > 1 && wis[1].end - first.start > max_duration * 2.f;
            if (pause > cmd * 4.f && (first_too_long || both_too_long)) {
                if (wis.size() > 1 && wis[1].duration() > max_duration) {
                    const float boundary = std::max(wis[1].end / 2.f, wis[1].end - max_duration);
                    wis[0].end = boundary;
                    wis[1].start = boundary;
                }
                wis[0].start = std::max(last_speech, wis[0].end - max_duration);
            }
            if (seg_start < wis[0].end && seg_start - 0.5f > wis[0].start) wis[0].start = std::max(0.f, std::min(wis[0].end - cmd, seg_start));
            else seg.start = wis[0].start;
            const Word last = wis.back();
            if (seg.end > last.start && seg_end + 0.5f < last.end) wis.back().end = std::max(last.start + cmd, seg_end);
            else seg.end = last.end;
            last_speech = seg.end;
        }
        seg_end_updated.push_back(seg.end);
        for (Word& w : wis) out.push_back(std::move(w));
    }
    return WK_OK;
}

}  // namespace

struct wk_words {
    std::vector<Word> v;
};

static wk_words* to_handle(std::vector<Word>&& v) {
    wk_words* h = new wk_words();
    h->v = std::move(v);
    return h;
}

extern "C" {

int32_t wk_words_count(const wk_words* w) { return w ? (int32_t)w->v.size() : 0; }

wk_status wk_words_get(const wk_words* w, int32_t i, wk_word* out) {
    if (!w || !out || i < 0 || i >= (int32_t)w->v.size()) { set_error("wk_words_get: index out of range"); return WK_ERR_INVALID_ARGUMENT; }
    const Word& x = w->v[i];
    out->word = x.word.c_str();
    out->tokens = x.tokens.data(); out->n_tokens = (int32_t)x.tokens.size();
    out->start = x.start; out->end = x.end; out->probability = x.probability; out->segment = x.segment;
    return WK_OK;
}

void wk_words_free(wk_words* w) { delete w; }

wk_status wk_dtw(const void* matrix, int32_t dtype, int32_t rows, int32_t cols, int64_t ld, int32_t* text_indices, int32_t* time_indices,
                 int32_t cap, int32_t* n_path) {
    if (!text_indices || !time_indices || !n_path) { set_error("wk_dtw: null output"); return WK_ERR_INVALID_ARGUMENT; }
    std::vector<int32_t> ti, tj;
    WK_CHECK(run_dtw(matrix, dtype, rows, cols, ld, ti, tj));
    if ((int32_t)ti.size() > cap) { set_error("wk_dtw: path of %zu entries exceeds capacity %d", ti.size(), cap); return WK_ERR_INVALID_ARGUMENT; }
    memcpy(text_indices, ti.data(), ti.size() * 4);
    memcpy(time_indices, tj.data(), tj.size() * 4);
    *n_path = (int32_t)ti.size();
    return WK_OK;
}

wk_status wk_find_alignment(const wk_word* words, int32_t n_words, const void* matrix, int32_t dtype, int32_t rows, int32_t cols, int64_t ld,
                            const float* token_logprobs, int32_t n_logprobs, wk_words** out) {
    if (!out || n_words < 0 || (n_words > 0 && !words) || !token_logprobs) { set_error("wk_find_alignment: bad arguments"); return WK_ERR_INVALID_ARGUMENT; }
    std::vector<Word> r;
    WK_CHECK(find_alignment(from_c(words, n_words), matrix, dtype, rows, cols, ld, token_logprobs, n_logprobs, r));
    *out = to_handle(std::move(r));
    return WK_OK;
}

wk_status wk_merge_punctuations(const wk_word* alignment, int32_t n, const char* prepended, const char* appended, wk_words** out) {
    if (!out || n < 0 || (n > 0 && !alignment)) { set_error("wk_merge_punctuations: bad arguments"); return WK_ERR_INVALID_ARGUMENT; }
    *out = to_handle(merge_punctuations(from_c(alignment, n), prepended ? prepended : kDefaultPrepend, appended ? appended : kDefaultAppend));
    return WK_OK;
}

wk_status wk_word_duration_constraints(const wk_word* alignment, int32_t n, float* constrained_median, float* max_duration) {
    if (!constrained_median || !max_duration || n < 0 || (n > 0 && !alignment)) { set_error("wk_word_duration_constraints: bad arguments"); return WK_ERR_INVALID_ARGUMENT; }
    duration_constraints(from_c(alignment, n), constrained_median, max_duration);
    return WK_OK;
}

wk_status wk_truncate_long_words(const wk_word* alignment, int32_t n, float max_duration, wk_words** out) {
    if (!out || n < 0 || (n > 0 && !alignment)) { set_error("wk_truncate_long_words: bad arguments"); return WK_ERR_INVALID_ARGUMENT; }
    std::vector<Word> v = from_c(alignment, n);
    truncate_long_words(v, max_duration);
    *out = to_handle(std::move(v));
    return WK_OK;
}

wk_status wk_update_segments_with_word_timings(wk_segment* segs, int32_t n_segs, const int32_t* tokens, const wk_word* merged, int32_t n_merged,
                                               int64_t seek, float last_speech_timestamp, float constrained_median, float max_duration,
                                               int32_t special_token_begin, const wk_tokenizer_hooks* hooks, wk_words** out) {
    if (!out || n_segs < 0 || (n_segs > 0 && (!segs || !tokens)) || n_merged < 0 || (n_merged > 0 && !merged)) { set_error("wk_update_segments_with_word_timings: bad arguments"); return WK_ERR_INVALID_ARGUMENT; }
    std::vector<Word> r;
    WK_CHECK(update_segments(segs, n_segs, tokens, from_c(merged, n_merged), seek, last_speech_timestamp, constrained_median, max_duration,
                             special_token_begin, hooks, r));
    *out = to_handle(std::move(r));
    return WK_OK;
}

wk_status wk_add_word_timestamps(wk_segment* segs, int32_t n_segs, const int32_t* tokens, const float* token_logprobs,
                                 const void* alignment, int32_t dtype, int32_t rows, int32_t cols, int64_t ld,
                                 const wk_tokenizer_hooks* hooks, int64_t seek, float last_speech_timestamp, int32_t special_token_begin,
                                 const char* prepended, const char* appended, wk_words** out) {
    if (!out || n_segs < 0 || (n_segs > 0 && (!segs || !tokens || !token_logprobs)) || !hooks || !hooks->split_to_word_tokens) {
        set_error("wk_add_word_timestamps: bad arguments (a split_to_word_tokens hook is required)");
        return WK_ERR_INVALID_ARGUMENT;
    }
    // wordTokenIds / filteredLogProbs / filteredIndices (:424-441): row (index inside the concatenated segment tokens) of the alignment
    std::vector<int32_t> ids;
    std::vector<float> lps;
    for (int s = 0; s < n_segs; ++s)
        for (int t = 0; t < segs[s].n_tokens; ++t) { ids.push_back(tokens[segs[s].token_offset + t]); lps.push_back(token_logprobs[segs[s].token_offset + t]); }
    std::vector<Word> alignment_words;
    if (!ids.empty()) {
        if ((int)ids.size() > rows) { set_error("wk_add_word_timestamps: %zu segment tokens but %d alignment rows", ids.size(), rows); return WK_ERR_INVALID_ARGUMENT; }
        std::vector<char> text(ids.size() * 64 + 256);
        std::vector<int32_t> counts(ids.size() + 1);
        const int32_t nw = hooks->split_to_word_tokens(hooks->user, ids.data(), (int32_t)ids.size(), text.data(), (int32_t)text.size(), counts.data(), (int32_t)counts.size());
        if (nw < 0) { set_error("tokenizer split_to_word_tokens hook failed (%d)", nw); return WK_ERR_TRANSCRIPTION_FAILED; }
        std::vector<Word> words((size_t)nw);
        size_t tpos = 0, cpos = 0;
        for (int i = 0; i < nw; ++i) {
            const size_t len = strnlen(text.data() + tpos, text.size() - tpos);
            words[i].word.assign(text.data() + tpos, len);
            tpos += len + 1;
            if (counts[i] < 1 || cpos + counts[i] > ids.size()) { set_error("split_to_word_tokens: token counts do not cover the %zu tokens", ids.size()); return WK_ERR_TRANSCRIPTION_FAILED; }
            words[i].tokens.assign(ids.begin() + cpos, ids.begin() + cpos + counts[i]);
            cpos += counts[i];
        }
        // the first ids.size() rows are exactly the filtered rows: filteredIndices = 0, 1, 2, ... (:427-441)
        WK_CHECK(find_alignment(words, alignment, dtype, (int)ids.size(), cols, ld, lps.data(), (int)lps.size(), alignment_words));
    }
    float median = 0.f, max_duration = 0.f;
    duration_constraints(alignment_words, &median, &max_duration);
    truncate_long_words(alignment_words, max_duration);
    if (!alignment_words.empty())
        alignment_words = merge_punctuations(alignment_words, prepended ? prepended : kDefaultPrepend, appended ? appended : kDefaultAppend);
    std::vector<Word> r;
    WK_CHECK(update_segments(segs, n_segs, tokens, alignment_words, seek, last_speech_timestamp, median, max_duration, special_token_begin, hooks, r));
    *out = to_handle(std::move(r));
    return WK_OK;
}

}  // extern "C"

// Result writers: ResultWriting.formatTime and the SRT / VTT / JSON bodies of
// Sources/WhisperKit/Utilities/ResultWriter.swift:12-134.  Host-only code; cues are passed in flat (one per word when the segment has word
// timings, else one per segment - the caller flattens exactly as WriteSRT / WriteVTT iterate, ResultWriter.swift:79-92,115-125).
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>

#include "kernels.h"

namespace {
// formatTime (ResultWriter.swift:14-26): Swift Float arithmetic, Int() truncation, truncatingRemainder = fmodf
std::string format_time(float seconds, bool always_hours, const char* marker) {
    const int hrs = (int)(seconds / 3600.f);
    const int mins = (int)(fmodf(seconds, 3600.f) / 60.f);
    const int secs = (int)fmodf(seconds, 60.f);
    const int msec = (int)((seconds - floorf(seconds)) * 1000.f);
    char buf[64];
    if (always_hours || hrs > 0) snprintf(buf, sizeof(buf), "%02d:%02d:%02d%s%03d", hrs, mins, secs, marker, msec);
    else snprintf(buf, sizeof(buf), "%02d:%02d%s%03d", mins, secs, marker, msec);
    return buf;
}
int32_t emit(const std::string& s, char* out, int32_t cap) {
    if ((int64_t)s.size() + 1 > cap) return -(int32_t)(s.size() + 1);
    memcpy(out, s.c_str(), s.size() + 1);
    return (int32_t)s.size();
}
}  // namespace

extern "C" {

int32_t wk_format_time(float seconds, int32_t always_include_hours, const char* decimal_marker, char* out, int32_t cap) {
    if (!out || !decimal_marker) return -1;
    return emit(format_time(seconds, always_include_hours != 0, decimal_marker), out, cap);
}

// WriteSRT body (ResultWriter.swift:70-101): "index\nHH:MM:SS,mmm --> HH:MM:SS,mmm\ntext\n\n", index from 1
int32_t wk_write_srt(const float* starts, const float* ends, const char* const* texts, int32_t n, char* out, int32_t cap) {
    if (n < 0 || (n > 0 && (!starts || !ends || !texts)) || !out) return -1;
    std::string s;
    for (int i = 0; i < n; ++i)
        s += std::to_string(i + 1) + "\n" + format_time(starts[i], true, ",") + " --> " + format_time(ends[i], true, ",") + "\n" + texts[i] + "\n\n";
    return emit(s, out, cap);
}

// WriteVTT body (ResultWriter.swift:103-134): "WEBVTT\n\n" then "MM:SS.mmm --> MM:SS.mmm\ntext\n\n" (hours only when non-zero)
int32_t wk_write_vtt(const float* starts, const float* ends, const char* const* texts, int32_t n, char* out, int32_t cap) {
    if (n < 0 || (n > 0 && (!starts || !ends || !texts)) || !out) return -1;
    std::string s = "WEBVTT\n\n";
    for (int i = 0; i < n; ++i)
        s += format_time(starts[i], false, ".") + " --> " + format_time(ends[i], false, ".") + "\n" + texts[i] + "\n\n";
    return emit(s, out, cap);
}

}  // extern "C"

// cwrap.cpp -- flat C view (prefix mcb200_) of this repo's drop-in motioncam::Decoder / motioncam::raw, so the
// Python tests drive it and the compiled reference (prefix mcref_, oracle/ref_shim.cpp) through identical code.
#define MC_PREFIX mcb200_
#include "decoder_cwrap.inc"

#include <motioncam/Export.hpp>

extern "C" {

// Decoder::loadFrames (batched addition): loads every listed timestamp, keeps the frames in the handle.
// Returns the number of frames, or -1 with decoder_last_error() set.

int64_t mcb200_decoder_load_frames(void* hv, const int64_t* timestamps, int64_t n, uint8_t** out_ptrs, int64_t* out_sizes) {
    HandleT* h = static_cast<HandleT*>(hv);
    try {
        std::vector<motioncam::Timestamp> ts(timestamps, timestamps + n);
        static thread_local std::vector<std::vector<uint8_t>> data;
        std::vector<nlohmann::json> meta;
        h->dec->loadFrames(ts, data, meta);
        for (int64_t i = 0; i < n; i++) {
            out_ptrs[i] = data[i].data();
            out_sizes[i] = static_cast<int64_t>(data[i].size());
        }
        return n;
    } catch (const std::exception& e) {
        h->error = e.what();
        return -1;
    }
}

// Decoder::loadFramesToDevice: dst[i] are device pointers with room for cap[i] uint16.  Returns n, or -1 with
// decoder_last_error() set.  The frame metadata of the last call stays readable through mcb200_decoder_frame_metadata_at.
static thread_local std::vector<std::string> g_meta_text;

int64_t mcb200_decoder_load_frames_to_device(void* hv, const int64_t* timestamps, int64_t n, uint16_t* const* dst, const uint64_t* cap) {
    HandleT* h = static_cast<HandleT*>(hv);
    try {
        std::vector<motioncam::Timestamp> ts(timestamps, timestamps + n);
        std::vector<nlohmann::json> meta;
        h->dec->loadFramesToDevice(ts, dst, cap, meta);
        g_meta_text.resize(meta.size());
        for (size_t i = 0; i < meta.size(); i++) g_meta_text[i] = meta[i].dump();
        return n;
    } catch (const std::exception& e) {
        h->error = e.what();
        return -1;
    }
}

// Decoder::locateFrame: 0 and the location, or -1 with decoder_last_error() set.
int mcb200_decoder_locate(void* hv, int64_t timestamp, int64_t* payload_offset, uint32_t* payload_size) {
    HandleT* h = static_cast<HandleT*>(hv);
    try {
        const motioncam::FrameLocation w = h->dec->locateFrame(timestamp);
        if (payload_offset) *payload_offset = w.payloadOffset;
        if (payload_size) *payload_size = w.payloadSize;
        return 0;
    } catch (const std::exception& e) {
        h->error = e.what();
        return -1;
    }
}

size_t mcb200_decoder_feed(void* hv, char* buf, size_t cap) {
    return copy_out(static_cast<HandleT*>(hv)->dec->feedDescription(), buf, cap);
}

size_t mcb200_decoder_frame_metadata_at(int64_t i, char* buf, size_t cap) {
    if (i < 0 || static_cast<size_t>(i) >= g_meta_text.size()) return 0;
    return copy_out(g_meta_text[static_cast<size_t>(i)], buf, cap);
}

// ---- include/motioncam/Export.hpp ---------------------------------------------------------------------------------
// 0 = written; 1 = an exception escaped (text in err).  Same shape as mcref_write_dng / mcref_write_audio in
// oracle/ref_shim.cpp, which call the reference's example.cpp.
int mcb200_write_dng(const char* path, const uint8_t* data, size_t bytes, const char* frame_json, const char* container_json,
                     char* err, size_t errcap) {
    try {
        motioncam::DngWriter(nlohmann::json::parse(container_json)).write(path, data, bytes, nlohmann::json::parse(frame_json));
        return 0;
    } catch (const std::exception& e) {
        copy_out(e.what(), err, errcap);
        return 1;
    }
}

int mcb200_write_audio(const char* path, int sample_rate_hz, int channels, const int16_t* samples, const int64_t* offsets,
                       int64_t nchunks, char* err, size_t errcap) {
    try {
        std::vector<motioncam::AudioChunk> chunks;
        for (int64_t i = 0; i < nchunks; i++)
            chunks.emplace_back(-1, std::vector<int16_t>(samples + offsets[i], samples + offsets[i + 1]));
        motioncam::writeAudio(path, sample_rate_hz, channels, chunks);
        return 0;
    } catch (const std::exception& e) {
        copy_out(e.what(), err, errcap);
        return 1;
    }
}

// The example program's loop (audio.wav + frame_%06d.dng into out_dir) on the batched B200 decode.
// Returns the number of frames written, or -1 (text in err).  stats (may be null) receives 6 doubles: total, open+audio,
// decode, first batch, writer wait, steady seconds (motioncam::ExportStats).
int64_t mcb200_export_clip(const char* input_path, const char* out_dir, int num_frames, int batch, int writer_threads,
                           int write_audio, double* stats, char* err, size_t errcap) {
    try {
        motioncam::ExportOptions opt;
        opt.numFrames = num_frames;
        if (batch > 0) opt.batch = batch;
        if (writer_threads > 0) opt.writerThreads = writer_threads;
        opt.writeAudio = write_audio != 0;
        motioncam::ExportStats st;
        const size_t n = motioncam::exportClip(input_path, out_dir, opt, nullptr, &st);
        if (stats) {
            const double v[6] = {st.totalSeconds, st.openAndAudioSeconds, st.decodeSeconds, st.firstBatchSeconds,
                                 st.writerWaitSeconds, st.steadySeconds};
            std::memcpy(stats, v, sizeof v);
        }
        return static_cast<int64_t>(n);
    } catch (const std::exception& e) {
        copy_out(e.what(), err, errcap);
        return -1;
    }
}

}  // extern "C"

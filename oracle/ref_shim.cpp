// ref_shim.cpp -- extern "C" access to the UNMODIFIED reference, for tests and the CPU baseline only.
//
// Compiled by oracle/Makefile together with /root/reference/lib/{RawData,RawData_Legacy,Decoder}.cpp and
// /root/reference/example.cpp (its main renamed on the command line; the DNG / WAV packaging it contains is the
// checker for include/motioncam/Export.hpp) -- sources read in place, never copied -- into oracle/_ref/libmcraw_ref.so.  Nothing in the product links
// or loads this library.
#define MC_PREFIX mcref_
#include "../motioncam-decoder_b200/csrc/decoder_cwrap.inc"

#include <atomic>
#include <chrono>
#include <thread>

// Defined in the reference's example.cpp:27-53 and :55-139 (external linkage, global namespace).
void writeAudio(const std::string& outputPath, const int sampleRateHz, const int numChannels, std::vector<motioncam::AudioChunk>& audioChunks);
void writeDng(const std::string& outputPath, const std::vector<uint8_t>& data, const nlohmann::json& metadata, const nlohmann::json& containerMetadata);

extern "C" {

// 0 = written; 1 = an exception escaped (text in err).
int mcref_write_dng(const char* path, const uint8_t* data, size_t bytes, const char* frame_json, const char* container_json,
                    char* err, size_t errcap) {
    try {
        std::vector<uint8_t> v(data, data + bytes);
        writeDng(path, v, nlohmann::json::parse(frame_json), nlohmann::json::parse(container_json));
        return 0;
    } catch (const std::exception& e) {
        copy_out(e.what(), err, errcap);
        return 1;
    }
}

// chunk i = samples[offsets[i] .. offsets[i+1])
int mcref_write_audio(const char* path, int sample_rate_hz, int channels, const int16_t* samples, const int64_t* offsets,
                      int64_t nchunks, char* err, size_t errcap) {
    try {
        std::vector<motioncam::AudioChunk> chunks;
        for (int64_t i = 0; i < nchunks; i++)
            chunks.emplace_back(-1, std::vector<int16_t>(samples + offsets[i], samples + offsets[i + 1]));
        writeAudio(path, sample_rate_hz, channels, chunks);
        return 0;
    } catch (const std::exception& e) {
        copy_out(e.what(), err, errcap);
        return 1;
    }
}

// Frame-parallel timing of the reference codec on host cores:
// `threads` threads each loop over their share of the frames (t, t+T, ...) `iters` times after `warmup`
// untimed passes, decoding into a private pre-touched buffer.  Returns wall seconds of the timed part
// and writes the number of frames decoded inside it to *frames_done.
double mcref_bench_mt(int compression_type, const uint8_t* const* ins, const size_t* lens, int nframes,
                      int width, int height, int threads, int iters, int warmup, int64_t* frames_done) {
    if (threads < 1) threads = 1;
    std::atomic<int> ready{0};
    std::atomic<bool> go{false};
    std::atomic<int64_t> done{0};
    std::vector<std::thread> pool;
    auto fn = [&](int t) {
        std::vector<uint16_t> out(static_cast<size_t>(width) * height + 64, 1);
        auto pass = [&](int64_t* cnt) {
            for (int f = t; f < nframes; f += threads) {
                size_t r = compression_type == 6
                               ? motioncam::raw::DecodeLegacy(out.data(), width, height, ins[f], lens[f])
                               : motioncam::raw::Decode(out.data(), width, height, ins[f], lens[f]);
                if (r && cnt) ++*cnt;
            }
        };
        for (int i = 0; i < warmup; i++) pass(nullptr);
        ready.fetch_add(1);
        while (!go.load(std::memory_order_acquire)) std::this_thread::yield();
        int64_t cnt = 0;
        for (int i = 0; i < iters; i++) pass(&cnt);
        done.fetch_add(cnt);
    };
    for (int t = 0; t < threads; t++) pool.emplace_back(fn, t);
    while (ready.load() < threads) std::this_thread::yield();
    auto t0 = std::chrono::steady_clock::now();
    go.store(true, std::memory_order_release);
    for (auto& th : pool) th.join();
    auto t1 = std::chrono::steady_clock::now();
    if (frames_done) *frames_done = done.load();
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"

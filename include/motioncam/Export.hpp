// motioncam/Export.hpp -- what the reference does with a decoded frame: its example program packs every frame into an
// uncompressed CFA DNG and the audio chunks into one 16-bit PCM WAV (/root/reference/example.cpp:27-53 writeAudio,
// :55-139 writeDng, :141-203 main).  The reference keeps that code in the example, on top of two vendored writers
// (thirdparty/tinydng, thirdparty/audiofile); here it is a library call so that the batched B200 decode has a consumer
// of the same shape.  The files are byte-identical to the reference program's
// (tests/test_export_cpu.py compares with the compiled reference).
//
// Not on the device: a DNG is 8 header bytes, the decoded pixels exactly as raw::Decode produced them, and ~700 bytes
// of tags.  The writer therefore never copies the pixels -- header, the caller's buffer and the tag block go out in one
// vectored write (the reference's writer copies the frame three times before the first byte reaches the file).
#pragma once
#include <motioncam/Decoder.hpp>

#include <cstdint>
#include <string>
#include <vector>

namespace motioncam {

// Container-level DNG fields, parsed once per clip (example.cpp:64-73 re-reads them from the JSON for every frame).
// Throws MotionCamException when a key is missing or has the wrong shape (the reference lets nlohmann::json's own
// exception escape, which its main() does not catch) and for an unknown "sensorArrangment" (example.cpp:95-106;
// message "Invalid sensor arrangement").
class DngWriter {
public:
    explicit DngWriter(const nlohmann::json& containerMetadata);

    // One frame -> one DNG file.  `pixels`/`bytes` is the buffer Decoder::loadFrame returned (2*width*height bytes,
    // host memory); width, height and asShotNeutral come from the frame metadata (example.cpp:61-64).
    void write(const std::string& path, const uint8_t* pixels, size_t bytes, const nlohmann::json& frameMetadata) const;

    // The same file in memory (tests, network sinks).
    std::vector<uint8_t> encode(const uint8_t* pixels, size_t bytes, const nlohmann::json& frameMetadata) const;

    // Everything after the pixel strip: the out-of-line tag values followed by the IFD.  File = header(8) | pixels | tail.
    struct Tail {
        std::vector<uint8_t> bytes;
        size_t ifdOffset;          // where the IFD starts inside `bytes`
    };
    Tail tail(size_t stripBytes, const nlohmann::json& frameMetadata) const;
    static void header(uint8_t out[8], size_t stripBytes, const Tail& tail);

private:
    uint16_t mBlackLevel[4];
    uint16_t mWhiteLevel;
    uint8_t mCfa[4];
    std::vector<float> mColor1, mColor2, mForward1, mForward2;
};

// example.cpp:55-139 with its signature: data = the vector loadFrame filled.
void writeDng(const std::string& outputPath, const std::vector<uint8_t>& data, const nlohmann::json& metadata,
              const nlohmann::json& containerMetadata);

// example.cpp:27-53: all chunks, in order, into one 16-bit PCM WAV; numChannels 1 or 2 (other values write an empty
// data chunk with that channel count, as the reference does).  For two channels an odd trailing sample of a chunk is
// dropped (the reference reads one element past the chunk there).
void writeAudio(const std::string& outputPath, int sampleRateHz, int numChannels, const std::vector<AudioChunk>& audioChunks);
std::vector<uint8_t> encodeAudio(int sampleRateHz, int numChannels, const std::vector<AudioChunk>& audioChunks);

// The whole example program (example.cpp:141-203) on the batched path: audio.wav + frame_%06d.dng for the first
// `numFrames` frames (negative = all) into `outputDir`; frames are decoded `batch` at a time on the GPU
// (Decoder::loadFramesPinned) and written by `writerThreads` threads, straight from the pinned result buffer, while the
// next batch decodes.
// Returns the number of frames written.  `log` (may be null) receives the reference's progress lines.
struct ExportOptions {
    int numFrames = -1;
    int batch = 16;
    int writerThreads = 4;
    bool writeAudio = true;
};
// Where the time went (wall clock of the calling thread).
struct ExportStats {
    size_t frames = 0;
    double totalSeconds = 0;
    double openAndAudioSeconds = 0;    // container open + index, audio.wav
    double decodeSeconds = 0;          // sum over batches of file read + H2D + kernels + D2H (Decoder::loadFramesPinned)
    double firstBatchSeconds = 0;      // the part of decodeSeconds spent in the first batch: CUDA context, pinned allocations
    double writerWaitSeconds = 0;      // main thread waiting for DNG writers
    double steadySeconds = 0;          // from the end of the first batch's decode to the end: `frames` DNG writes and
                                       // frames - batch decodes, overlapped
};
size_t exportClip(const std::string& inputPath, const std::string& outputDir, const ExportOptions& options, std::FILE* log,
                  ExportStats* stats = nullptr);

}  // namespace motioncam

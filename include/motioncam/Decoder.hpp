// motioncam/Decoder.hpp -- .mcraw container reader with the public surface of the reference's
// lib/include/motioncam/Decoder.hpp:28-73 (same names, argument meaning, exception types and messages), decoding
// frames on the B200 instead of the host.
//
// Differences that a caller can observe:
//   * the file is read with positional reads (pread), so the FILE* position is never moved; like the reference,
//     one Decoder instance is meant for one thread;
//   * frames with equal timestamps keep their index order (the reference's std::sort leaves it unspecified);
//   * loadFrames() / loadFramesToDevice() are additions: many frames per call, one batched device decode.
#pragma once
#include <motioncam/Container.hpp>
#include <nlohmann/json.hpp>

#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace motioncam {

typedef int64_t Timestamp;
typedef std::pair<Timestamp, std::vector<int16_t>> AudioChunk;   // (timestampNs or -1, interleaved int16 PCM)

class MotionCamException : public std::runtime_error {
public:
    MotionCamException(const std::string& error) : runtime_error(error) {}
};

class IOException : public MotionCamException {
public:
    IOException(const std::string& error) : MotionCamException(error) {}
};

class AudioChunkLoader {
public:
    virtual bool next(AudioChunk& output) = 0;
    virtual ~AudioChunkLoader() = default;
};

// Where one frame lives in the file (what the batched feeders need to read it without the JSON round trip).
struct FrameLocation {
    Timestamp timestamp;
    int64_t payloadOffset;     // file offset of the compressed frame bytes (after the BUFFER Item header)
    uint32_t payloadSize;      // Item::size of the BUFFER record
};

// One decoded frame inside the Decoder's pinned result buffer (loadFramesPinned).
struct FrameView {
    const uint8_t* data;       // width*height little-endian uint16, row-major
    size_t size;               // bytes
};

class Decoder {
public:
    Decoder(const std::string& path);
    Decoder(FILE* file);       // takes ownership once constructed: the handle is closed by the destructor; if the constructor
                               // throws, the handle stays open and the caller's (Decoder.cpp:97-114).  The container is read
                               // from the handle's current position, like the reference's init() (Decoder.cpp:116-141)
    ~Decoder();
    Decoder(const Decoder&) = delete;
    Decoder& operator=(const Decoder&) = delete;

    const nlohmann::json& getContainerMetadata() const;
    const std::vector<Timestamp>& getFrames() const;                // sorted by timestamp
    void loadFrame(const Timestamp timestamp, std::vector<uint8_t>& outData, nlohmann::json& outMetadata);
    int audioSampleRateHz() const;
    int numAudioChannels() const;
    void loadAudio(std::vector<AudioChunk>& outAudioChunks);        // appends every chunk of the audio index
    AudioChunkLoader& loadAudio() const;                            // one chunk per next(); position persists

    // ---- additions (batched, B200) -------------------------------------------------------------------
    // Locate a frame's compressed bytes; throws IOException like loadFrame for unknown timestamps.
    FrameLocation locateFrame(const Timestamp timestamp) const;
    // Read the frame's compressed bytes into dst (payloadSize bytes, e.g. pinned memory; nullptr = skip) and parse its JSON.
    void readFrame(const FrameLocation& where, uint8_t* dst, nlohmann::json& outMetadata) const;
    // loadFrame for many timestamps: one overlapped H2D + batched decode + D2H.  outData[i] / outMetadata[i]
    // are what loadFrame(timestamps[i], ...) would have produced; the same exceptions are thrown.
    void loadFrames(const std::vector<Timestamp>& timestamps, std::vector<std::vector<uint8_t>>& outData,
                    std::vector<nlohmann::json>& outMetadata);
    // loadFrames without the last host copy: outFrames[i] points into one of this Decoder's two pinned result buffers.
    // The buffers alternate, so the views stay valid through the NEXT loadFrames/loadFramesPinned call and die with the
    // one after it (or with the Decoder): a consumer that streams the pixels on -- file writers, sockets -- reads
    // batch k where the D2H copy put it while batch k+1 is being decoded.
    void loadFramesPinned(const std::vector<Timestamp>& timestamps, std::vector<FrameView>& outFrames,
                          std::vector<nlohmann::json>& outMetadata);
    // The same, but the decoded frames stay on the GPU: dst[i] is a DEVICE pointer (16-byte aligned) with room for
    // dstCapacityElems[i] uint16 (>= width*height of frame i).  Compressed frames go file -> pinned ring (kept by the
    // Decoder and reused) -> staged H2D on side streams -> kernels; nothing is copied back.  Device: MCRAW_B200_DEVICE.
    void loadFramesToDevice(const std::vector<Timestamp>& timestamps, uint16_t* const* dst, const uint64_t* dstCapacityElems,
                            std::vector<nlohmann::json>& outMetadata);
    // How loadFramesToDevice gets the compressed bytes to the GPU.  Default: reader threads pread into the pinned ring while
    // the chunks already read travel and decode.  MCRAW_FEED in the environment asks for another feed: "direct" (O_DIRECT
    // reads into the ring), "cufile" (cuFileRead into device memory: GPUDirect Storage), "mmap" (H2D straight from a
    // page-locked mapping of the file); each falls back to the ring, with the reason in this text, where the platform refuses.
    const char* feedDescription() const;

    struct Feed;    // state of the optional feeds (Decoder.cpp)

private:
    struct Impl;
    std::unique_ptr<Impl> m;
};

}  // namespace motioncam

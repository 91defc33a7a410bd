// Export.cpp -- DNG / WAV packaging of decoded frames and the batched "dump a clip" loop (include/motioncam/Export.hpp).
//
// File layout the reference's program produces (observed from /root/reference/example.cpp:55-139 driving
// thirdparty/tinydng/tiny_dng_writer.h; restated here, nothing of it is compiled in):
//
//   [0..8)            "II" 2A 00, u32 offset of the IFD
//   [8..8+S)          the pixel strip, S = data.size()                      (tiny_dng_writer.h SetImageData)
//   [8+S..ifd)        values that do not fit the 4-byte IFD slot, in the order example.cpp sets them:
//                     BlackLevel 4xSHORT, ColorMatrix1/2 and ForwardMatrix1/2 9xSRATIONAL each, AsShotNeutral
//                     3xRATIONAL, UniqueCameraModel "MotionCam\0", ActiveArea 4xLONG
//   [ifd..)           u16 entry count, 12-byte entries sorted by tag, u32 0 (no next IFD)
//
// Quirks kept because they decide bytes: WhiteLevel is a SHORT made from a double through `short`
// (tiny_dng_writer.h:1074), rationals are the float's exact binary fraction reduced by common powers of two and
// narrowed with the x86 conversion rules (:500-536 and the static_casts at :1425-1426, :1741-1742), a matrix with a
// non-finite entry loses its tag altogether (the setter returns false and example.cpp ignores it).
#include <motioncam/Export.hpp>

#include <algorithm>
#include <cerrno>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstring>
#include <fcntl.h>
#include <future>
#include <sys/uio.h>
#include <unistd.h>

namespace motioncam {
namespace {

enum : uint16_t { T_BYTE = 1, T_ASCII = 2, T_SHORT = 3, T_LONG = 4, T_RATIONAL = 5, T_SRATIONAL = 10 };

inline void put16(std::vector<uint8_t>& v, uint32_t x) {
    v.push_back(static_cast<uint8_t>(x));
    v.push_back(static_cast<uint8_t>(x >> 8));
}
inline void put32(std::vector<uint8_t>& v, uint32_t x) {
    put16(v, x & 0xFFFFu);
    put16(v, x >> 16);
}

// float -> int32 / uint32 the way the reference binary does it on x86-64: cvttss2si on 32 bits (out of range or NaN
// gives INT_MIN) and, for unsigned, cvttss2si on 64 bits keeping the low half.
inline uint32_t narrowSigned(float v) {
    return (v >= -2147483648.0f && v < 2147483648.0f) ? static_cast<uint32_t>(static_cast<int32_t>(v)) : 0x80000000u;
}
inline uint32_t narrowUnsigned(float v) {
    const int64_t t = (v >= -9223372036854775808.0f && v < 9223372036854775808.0f) ? static_cast<int64_t>(v) : INT64_MIN;
    return static_cast<uint32_t>(static_cast<uint64_t>(t));
}
inline uint16_t narrowShort(double v) {   // double -> short: cvttsd2si on 32 bits, low half kept
    const int32_t t = (v >= -2147483648.0 && v < 2147483648.0) ? static_cast<int32_t>(v) : INT32_MIN;
    return static_cast<uint16_t>(static_cast<uint32_t>(t));
}

// x = num/den with den a power of two: the float's own mantissa (24 bits) over 2^k, common factors of two removed.
// false for inf/nan, and for magnitudes below 2^-127 (where the reference reports failure unless |num| >= 1).
bool floatToFraction(float x, float& num, float& den) {
    if (!std::isfinite(x)) return false;
    int e = 0;
    const float m = std::frexp(x, &e);            // x = m * 2^e, 0.5 <= |m| < 1 (m = 0, e = 0 for zero)
    float mant = std::ldexp(m, 24);                // integer, |mant| < 2^24
    int k = 24 - e;                                // x = mant / 2^k
    if (k <= 0) {
        num = x;
        den = 1.0f;
        return true;
    }
    if (k >= 127) {                                // denominator would leave the float range: capped at 2^127, unreduced
        num = std::ldexp(mant, -(k - 127));
        den = std::ldexp(1.0f, 127);
        return !(std::fabs(num) < 1.0f);
    }
    if (mant != 0.0f) {
        int32_t im = static_cast<int32_t>(mant);
        while (k > 0 && (im & 1) == 0) {
            im /= 2;
            k--;
        }
        mant = static_cast<float>(im);
    }
    num = mant;
    den = std::ldexp(1.0f, k);
    return true;
}

// `count` rationals from floats; false when one of them has no fraction (the tag is then left out).
bool rationals(const float* values, size_t count, bool isSigned, std::vector<uint8_t>& out) {
    out.clear();
    for (size_t i = 0; i < count; i++) {
        float num, den;
        if (!floatToFraction(values[i], num, den)) return false;
        put32(out, isSigned ? narrowSigned(num) : narrowUnsigned(num));
        put32(out, isSigned ? narrowSigned(den) : narrowUnsigned(den));
    }
    return true;
}

struct Field {
    uint16_t tag;
    uint16_t type;
    uint32_t count;
    std::vector<uint8_t> value;     // little-endian bytes of the value(s)
};

struct FieldList {
    std::vector<Field> fields;      // in the order the reference sets them (decides where long values land)
    void add(uint16_t tag, uint16_t type, uint32_t count, std::vector<uint8_t> bytes) {
        fields.push_back(Field{tag, type, count, std::move(bytes)});
    }
    void addShort(uint16_t tag, uint32_t x) {
        std::vector<uint8_t> b;
        put16(b, x);
        add(tag, T_SHORT, 1, std::move(b));
    }
    void addLong(uint16_t tag, uint32_t x) {
        std::vector<uint8_t> b;
        put32(b, x);
        add(tag, T_LONG, 1, std::move(b));
    }
};

template <typename T>
T need(const nlohmann::json& j, const char* key, const char* where) {
    try {
        return j.at(key).get<T>();
    } catch (const nlohmann::json::exception& e) {
        throw MotionCamException(std::string("Invalid ") + where + " metadata: \"" + key + "\": " + e.what());
    }
}

std::vector<float> needFloats(const nlohmann::json& j, const char* key, size_t atLeast, const char* where) {
    std::vector<float> v = need<std::vector<float>>(j, key, where);
    if (v.size() < atLeast)
        throw MotionCamException(std::string("Invalid ") + where + " metadata: \"" + key + "\" needs " + std::to_string(atLeast) + " values");
    return v;
}

void writeAll(const std::string& path, const iovec* parts, int nparts) {
    const int fd = ::open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0644);
    if (fd < 0) throw IOException("Failed to open " + path + ": " + std::strerror(errno));
    std::vector<iovec> iov(parts, parts + nparts);
    size_t first = 0;
    while (first < iov.size()) {
        const ssize_t n = ::writev(fd, iov.data() + first, static_cast<int>(std::min<size_t>(iov.size() - first, 512)));   // <= IOV_MAX
        if (n < 0) {
            if (errno == EINTR) continue;
            const int err = errno;
            ::close(fd);
            throw IOException("Failed to write " + path + ": " + std::strerror(err));
        }
        size_t left = static_cast<size_t>(n);
        while (first < iov.size() && left >= iov[first].iov_len) left -= iov[first++].iov_len;
        if (first < iov.size()) {
            iov[first].iov_base = static_cast<uint8_t*>(iov[first].iov_base) + left;
            iov[first].iov_len -= left;
        }
    }
    if (::close(fd) != 0) throw IOException("Failed to write " + path + ": " + std::strerror(errno));
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// DNG
// ---------------------------------------------------------------------------------------------------------------
DngWriter::DngWriter(const nlohmann::json& c) {
    const std::vector<uint16_t> black = need<std::vector<uint16_t>>(c, "blackLevel", "container");
    if (black.size() < 4) throw MotionCamException("Invalid container metadata: \"blackLevel\" needs 4 values");
    for (int i = 0; i < 4; i++) mBlackLevel[i] = black[static_cast<size_t>(i)];
    mWhiteLevel = narrowShort(need<double>(c, "whiteLevel", "container"));

    const std::string arrangement = need<std::string>(c, "sensorArrangment", "container");   // sic: the container's spelling
    static const struct { const char* name; uint8_t cfa[4]; } kPatterns[] = {
        {"rggb", {0, 1, 1, 2}}, {"bggr", {2, 1, 1, 0}}, {"grbg", {1, 0, 2, 1}}, {"gbrg", {1, 2, 0, 1}}};
    bool known = false;
    for (const auto& p : kPatterns)
        if (arrangement == p.name) {
            std::memcpy(mCfa, p.cfa, 4);
            known = true;
        }
    if (!known) throw MotionCamException("Invalid sensor arrangement");

    mColor1 = needFloats(c, "colorMatrix1", 9, "container");
    mColor2 = needFloats(c, "colorMatrix2", 9, "container");
    mForward1 = needFloats(c, "forwardMatrix1", 9, "container");
    mForward2 = needFloats(c, "forwardMatrix2", 9, "container");
}

DngWriter::Tail DngWriter::tail(size_t stripBytes, const nlohmann::json& frame) const {
    const uint32_t width = need<unsigned int>(frame, "width", "frame");
    const uint32_t height = need<unsigned int>(frame, "height", "frame");
    const std::vector<float> neutral = needFloats(frame, "asShotNeutral", 3, "frame");
    if (stripBytes == 0) throw MotionCamException("Empty frame");     // the reference writes a file without a strip here
    if (stripBytes > 0xFFFFFF00u) throw MotionCamException("Frame too large for a classic TIFF container");

    FieldList f;
    f.add(50706, T_BYTE, 4, {1, 4, 0, 0});                              // DNGVersion
    f.add(50707, T_BYTE, 4, {1, 1, 0, 0});                              // DNGBackwardVersion
    f.addLong(279, static_cast<uint32_t>(stripBytes));                  // StripByteCounts
    f.addLong(256, width);
    f.addLong(257, height);
    f.addShort(284, 1);                                                 // PlanarConfiguration: chunky
    f.addShort(262, 32803);                                             // PhotometricInterpretation: CFA
    f.addLong(278, height);                                             // RowsPerStrip: one strip
    f.addShort(277, 1);                                                 // SamplesPerPixel
    f.add(33421, T_SHORT, 2, {2, 0, 2, 0});                             // CFARepeatPatternDim
    f.add(50713, T_SHORT, 2, {2, 0, 2, 0});                             // BlackLevelRepeatDim
    {
        std::vector<uint8_t> b;
        for (int i = 0; i < 4; i++) put16(b, mBlackLevel[i]);
        f.add(50714, T_SHORT, 4, std::move(b));                         // BlackLevel
    }
    f.addShort(50717, mWhiteLevel);
    f.addShort(259, 1);                                                 // Compression: none
    f.add(33422, T_BYTE, 4, {mCfa[0], mCfa[1], mCfa[2], mCfa[3]});      // CFAPattern
    f.addShort(50711, 1);                                               // CFALayout: rectangular
    f.addShort(258, 16);                                                // BitsPerSample
    std::vector<uint8_t> r;
    if (rationals(mColor1.data(), 9, true, r)) f.add(50721, T_SRATIONAL, 9, r);
    if (rationals(mColor2.data(), 9, true, r)) f.add(50722, T_SRATIONAL, 9, r);
    if (rationals(mForward1.data(), 9, true, r)) f.add(50964, T_SRATIONAL, 9, r);
    if (rationals(mForward2.data(), 9, true, r)) f.add(50965, T_SRATIONAL, 9, r);
    if (rationals(neutral.data(), 3, false, r)) f.add(50728, T_RATIONAL, 3, r);
    f.addShort(50778, 21);                                              // CalibrationIlluminant1: D65
    f.addShort(50779, 17);                                              // CalibrationIlluminant2: standard light A
    {
        static const char kModel[] = "MotionCam";
        f.add(50708, T_ASCII, sizeof(kModel), std::vector<uint8_t>(kModel, kModel + sizeof(kModel)));
    }
    f.addLong(254, 0);                                                  // NewSubfileType: main image
    {
        std::vector<uint8_t> b;
        put32(b, 0);
        put32(b, 0);
        put32(b, height);
        put32(b, width);
        f.add(50829, T_LONG, 4, std::move(b));                          // ActiveArea
    }
    f.addLong(273, 8);                                                  // StripOffsets: right behind the header

    // Long values first (file offset = 8 + strip + position), then the sorted directory.
    Tail t;
    std::vector<uint32_t> where(f.fields.size(), 0);
    for (size_t i = 0; i < f.fields.size(); i++) {
        const Field& fd = f.fields[i];
        if (fd.value.size() > 4) {
            where[i] = static_cast<uint32_t>(8 + stripBytes + t.bytes.size());
            t.bytes.insert(t.bytes.end(), fd.value.begin(), fd.value.end());
        }
    }
    t.ifdOffset = t.bytes.size();
    std::vector<size_t> order(f.fields.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return f.fields[a].tag < f.fields[b].tag; });
    put16(t.bytes, static_cast<uint32_t>(order.size()));
    for (size_t i : order) {
        const Field& fd = f.fields[i];
        put16(t.bytes, fd.tag);
        put16(t.bytes, fd.type);
        put32(t.bytes, fd.count);
        if (fd.value.size() > 4) {
            put32(t.bytes, where[i]);
        } else {
            uint8_t slot[4] = {0, 0, 0, 0};
            std::memcpy(slot, fd.value.data(), fd.value.size());
            t.bytes.insert(t.bytes.end(), slot, slot + 4);
        }
    }
    put32(t.bytes, 0);
    return t;
}

void DngWriter::header(uint8_t out[8], size_t stripBytes, const Tail& tail) {
    const uint32_t ifd = static_cast<uint32_t>(8 + stripBytes + tail.ifdOffset);
    out[0] = 'I';
    out[1] = 'I';
    out[2] = 0x2A;
    out[3] = 0;
    out[4] = static_cast<uint8_t>(ifd);
    out[5] = static_cast<uint8_t>(ifd >> 8);
    out[6] = static_cast<uint8_t>(ifd >> 16);
    out[7] = static_cast<uint8_t>(ifd >> 24);
}

std::vector<uint8_t> DngWriter::encode(const uint8_t* pixels, size_t bytes, const nlohmann::json& frame) const {
    const Tail t = tail(bytes, frame);
    std::vector<uint8_t> file(8 + bytes + t.bytes.size());
    header(file.data(), bytes, t);
    std::memcpy(file.data() + 8, pixels, bytes);
    std::memcpy(file.data() + 8 + bytes, t.bytes.data(), t.bytes.size());
    return file;
}

void DngWriter::write(const std::string& path, const uint8_t* pixels, size_t bytes, const nlohmann::json& frame) const {
    Tail t = tail(bytes, frame);
    uint8_t head[8];
    header(head, bytes, t);
    const iovec parts[3] = {{head, sizeof(head)}, {const_cast<uint8_t*>(pixels), bytes}, {t.bytes.data(), t.bytes.size()}};
    writeAll(path, parts, 3);
}

void writeDng(const std::string& outputPath, const std::vector<uint8_t>& data, const nlohmann::json& metadata,
              const nlohmann::json& containerMetadata) {
    DngWriter(containerMetadata).write(outputPath, data.data(), data.size(), metadata);
}

// ---------------------------------------------------------------------------------------------------------------
// WAV (16-bit PCM, the layout thirdparty/audiofile/AudioFile.h:937-1052 produces for AudioFile<int16_t>)
// ---------------------------------------------------------------------------------------------------------------
namespace {

// Samples per channel the reference ends up with (example.cpp:36-51): stereo takes pairs, mono takes everything,
// any other channel count takes nothing.
size_t usableSamples(const AudioChunk& c, int numChannels) {
    if (numChannels == 2) return c.second.size() / 2 * 2;
    if (numChannels == 1) return c.second.size();
    return 0;
}

void wavHeader(uint8_t h[44], int sampleRateHz, int numChannels, uint64_t dataBytes) {
    std::vector<uint8_t> v;
    v.reserve(44);
    const auto tag = [&](const char* s) { v.insert(v.end(), s, s + 4); };
    tag("RIFF");
    put32(v, static_cast<uint32_t>(36 + dataBytes));
    tag("WAVE");
    tag("fmt ");
    put32(v, 16);
    put16(v, 1);                                                        // PCM
    put16(v, static_cast<uint32_t>(numChannels) & 0xFFFFu);
    put32(v, static_cast<uint32_t>(sampleRateHz));
    put32(v, static_cast<uint32_t>(static_cast<int64_t>(numChannels) * sampleRateHz * 16 / 8));
    put16(v, static_cast<uint32_t>(numChannels * 2) & 0xFFFFu);
    put16(v, 16);
    tag("data");
    put32(v, static_cast<uint32_t>(dataBytes));
    std::memcpy(h, v.data(), 44);
}

uint64_t wavDataBytes(int numChannels, const std::vector<AudioChunk>& chunks) {
    uint64_t n = 0;
    for (const auto& c : chunks) n += usableSamples(c, numChannels) * sizeof(int16_t);
    if (n > 0x7FFFFFFFull - 36) throw MotionCamException("Audio too long for a WAV file");
    return n;
}

}  // namespace

std::vector<uint8_t> encodeAudio(int sampleRateHz, int numChannels, const std::vector<AudioChunk>& chunks) {
    const uint64_t dataBytes = wavDataBytes(numChannels, chunks);
    std::vector<uint8_t> file(44 + dataBytes);
    wavHeader(file.data(), sampleRateHz, numChannels, dataBytes);
    size_t at = 44;
    for (const auto& c : chunks) {
        const size_t n = usableSamples(c, numChannels) * sizeof(int16_t);
        if (n) std::memcpy(file.data() + at, c.second.data(), n);     // interleaved int16, little-endian host
        at += n;
    }
    return file;
}

void writeAudio(const std::string& outputPath, int sampleRateHz, int numChannels, const std::vector<AudioChunk>& chunks) {
    const uint64_t dataBytes = wavDataBytes(numChannels, chunks);
    uint8_t head[44];
    wavHeader(head, sampleRateHz, numChannels, dataBytes);
    std::vector<iovec> parts;
    parts.push_back({head, sizeof(head)});
    for (const auto& c : chunks) {
        const size_t n = usableSamples(c, numChannels) * sizeof(int16_t);
        if (n) parts.push_back({const_cast<int16_t*>(c.second.data()), n});
    }
    writeAll(outputPath, parts.data(), static_cast<int>(parts.size()));
}

// ---------------------------------------------------------------------------------------------------------------
// The example program's loop on the batched decode
// ---------------------------------------------------------------------------------------------------------------
size_t exportClip(const std::string& inputPath, const std::string& outputDir, const ExportOptions& opt, std::FILE* log,
                  ExportStats* stats) {
    using Clock = std::chrono::steady_clock;
    const auto seconds = [](Clock::time_point from, Clock::time_point to) { return std::chrono::duration<double>(to - from).count(); };
    const Clock::time_point t0 = Clock::now();
    if (stats) *stats = ExportStats();

    Decoder decoder(inputPath);
    const std::vector<Timestamp>& frames = decoder.getFrames();
    const nlohmann::json& containerMetadata = decoder.getContainerMetadata();
    const std::string dir = outputDir.empty() ? std::string() : (outputDir.back() == '/' ? outputDir : outputDir + "/");
    if (log) std::fprintf(log, "Found %zu frames\n", frames.size());

    if (opt.writeAudio) {
        std::vector<AudioChunk> chunks;
        decoder.loadAudio(chunks);
        writeAudio(dir + "audio.wav", decoder.audioSampleRateHz(), decoder.numAudioChannels(), chunks);
    }
    const Clock::time_point tAudio = Clock::now();
    if (stats) stats->openAndAudioSeconds = seconds(t0, tAudio);

    size_t end = frames.size();
    if (opt.numFrames >= 0 && static_cast<size_t>(opt.numFrames) < end) end = static_cast<size_t>(opt.numFrames);
    if (end == 0) {
        if (stats) stats->totalSeconds = seconds(t0, Clock::now());
        return 0;
    }
    const DngWriter writer(containerMetadata);
    const size_t batch = static_cast<size_t>(opt.batch > 0 ? opt.batch : 1);
    const size_t nthreads = static_cast<size_t>(opt.writerThreads > 0 ? opt.writerThreads : 1);

    // Two batches in flight: the writers stream batch k straight out of one of the Decoder's pinned result buffers
    // while the GPU decodes batch k+1 into the other (Decoder::loadFramesPinned) -- no host copy between the D2H copy
    // and the file.
    struct InFlight {
        std::vector<FrameView> frames;
        std::vector<nlohmann::json> metadata;
        std::vector<std::future<void>> writers;
        void drain() {
            std::exception_ptr firstError;
            for (auto& w : writers) {
                try {
                    w.get();
                } catch (...) {
                    if (!firstError) firstError = std::current_exception();
                }
            }
            writers.clear();
            if (firstError) std::rethrow_exception(firstError);
        }
    } inflight[2];

    size_t which = 0;
    Clock::time_point tFirstBatch = tAudio;
    try {
        for (size_t first = 0; first < end; first += batch, which ^= 1) {
            InFlight& b = inflight[which];
            Clock::time_point t = Clock::now();
            b.drain();                                   // its pinned buffer is about to be reused
            if (stats) stats->writerWaitSeconds += seconds(t, Clock::now());
            const size_t n = std::min(batch, end - first);
            const std::vector<Timestamp> stamps(frames.begin() + static_cast<ptrdiff_t>(first), frames.begin() + static_cast<ptrdiff_t>(first + n));
            t = Clock::now();
            decoder.loadFramesPinned(stamps, b.frames, b.metadata);
            if (stats) stats->decodeSeconds += seconds(t, Clock::now());
            if (first == 0) {
                tFirstBatch = Clock::now();
                if (stats) stats->firstBatchSeconds = seconds(t, tFirstBatch);
            }
            if (log)
                for (size_t i = 0; i < n; i++) std::fprintf(log, "Writing frame_%06zu.dng\n", first + i);
            for (size_t t = 0; t < std::min(nthreads, n); t++) {
                b.writers.push_back(std::async(std::launch::async, [&writer, &b, &dir, first, n, t, nthreads] {
                    char name[40];
                    for (size_t i = t; i < n; i += nthreads) {
                        std::snprintf(name, sizeof(name), "frame_%06zu.dng", first + i);
                        writer.write(dir + name, b.frames[i].data, b.frames[i].size, b.metadata[i]);
                    }
                }));
            }
        }
        const Clock::time_point t = Clock::now();
        inflight[0].drain();
        inflight[1].drain();
        if (stats) stats->writerWaitSeconds += seconds(t, Clock::now());
    } catch (...) {
        for (auto& b : inflight) {
            try {
                b.drain();
            } catch (...) {
            }
        }
        throw;
    }
    if (stats) {
        const Clock::time_point tEnd = Clock::now();
        stats->frames = end;
        stats->totalSeconds = seconds(t0, tEnd);
        // everything after the first batch has been decoded (context creation, pinned allocations and first touches are
        // in that batch): the first batch's writes and all later batches
        stats->steadySeconds = seconds(tFirstBatch, tEnd);
    }
    return end;
}

}  // namespace motioncam

// Decoder.cpp -- .mcraw container reader (host side of the decode path) behind the reference's public API.
//
// Mirrors the observable behaviour of /root/reference/lib/Decoder.cpp (open/index :116-151,237-315; loadFrame
// :184-235; audio :42-93,169-182) -- same exception types and messages, same ordering rules -- but is organised
// for the B200 path: the file is indexed once into FrameLocation records and read with pread (no shared file
// position), frames are decoded on the device through motioncam::raw::Decode* (one frame) or through
// mcraw_decode_batch_host (loadFrames: pinned staging, overlapped H2D, one batched decode).
#include <motioncam/Decoder.hpp>
#include <motioncam/RawData.hpp>

#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cufile.h>      // types and prototypes only: libcufile is looked up with dlopen when MCRAW_FEED=cufile asks for it

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <map>
#include <mutex>
#include <thread>

#include "dropin_ctx.hpp"

namespace motioncam {

namespace {

constexpr int kCompressionLegacy = MCRAW_COMPRESSION_LEGACY;    // Decoder.cpp:20
constexpr int kCompressionCurrent = MCRAW_COMPRESSION_CURRENT;  // Decoder.cpp:21

// Positional file access.  A short read is the reference's "Failed to read data" (Decoder.cpp:36-40).
class FileView {
public:
    explicit FileView(FILE* f) : mFile(f), mFd(fileno(f)) {
        struct stat st;
        mSize = (mFd >= 0 && fstat(mFd, &st) == 0) ? static_cast<int64_t>(st.st_size) : -1;
        // the reference reads the container header from wherever the handle stands (Decoder.cpp:116-141 never seeks);
        // everything after it is addressed by absolute offsets (:188,:239,:258,:284)
        const off_t at = ftello(f);
        mStart = at > 0 ? static_cast<int64_t>(at) : 0;
    }
    ~FileView() {
        if (mFile) std::fclose(mFile);
    }
    // give the handle back to whoever passed it in (a constructor that throws must not close the caller's FILE:
    // the reference leaves it open on that path, and the caller's own fclose would then be a double close)
    void release() { mFile = nullptr; }
    int64_t start() const { return mStart; }
    int64_t size() const { return mSize; }
    int fd() const { return mFd; }
    bool tryRead(int64_t offset, void* dst, size_t bytes) const {
        uint8_t* p = static_cast<uint8_t*>(dst);
        while (bytes) {
            const ssize_t n = ::pread(mFd, p, bytes, static_cast<off_t>(offset));
            if (n <= 0) return false;
            p += n; offset += n; bytes -= static_cast<size_t>(n);
        }
        return true;
    }
    void read(int64_t offset, void* dst, size_t bytes) const {
        if (!tryRead(offset, dst, bytes)) throw IOException("Failed to read data");
    }
    // A size taken from the file must fit in the file BEFORE memory is set aside for it (the read would fail anyway,
    // but only after a count of 2^60 index entries or a 4 GB record has been allocated and zeroed).
    void require(int64_t offset, uint64_t bytes) const {
        if (mSize < 0) return;                      // not a regular file: let the read decide
        if (offset < 0 || offset > mSize || bytes > static_cast<uint64_t>(mSize - offset)) throw IOException("Failed to read data");
    }

private:
    FILE* mFile;
    int mFd;
    int64_t mSize;
    int64_t mStart = 0;
};

// One audio chunk (Decoder.cpp:42-75).  false = unreachable offset (the reference's failed seek).
bool readAudioChunk(const FileView& file, const BufferOffset& where, AudioChunk& out) {
    if (where.offset < 0) return false;
    int64_t pos = where.offset;
    Item item{};
    file.read(pos, &item, sizeof item);
    pos += sizeof item;
    if (item.type != Type::AUDIO_DATA) throw IOException("Invalid audio data");
    file.require(pos, item.size);
    std::vector<int16_t> samples((static_cast<size_t>(item.size) + 1) / 2);
    file.read(pos, samples.data(), item.size);
    pos += item.size;
    // newer files append the chunk's timestamp; the reference reads the next record header unconditionally
    Item next{};
    file.read(pos, &next, sizeof next);
    pos += sizeof next;
    Timestamp ts = -1;
    if (next.type == Type::AUDIO_DATA_METADATA) {
        AudioMetadata md{};
        file.read(pos, &md, sizeof md);
        ts = md.timestampNs;
    }
    out = std::make_pair(ts, std::move(samples));
    return true;
}

class SequentialAudioLoader : public AudioChunkLoader {
public:
    SequentialAudioLoader(const FileView& file, const std::vector<BufferOffset>& offsets) : mFile(file), mOffsets(offsets) {}
    bool next(AudioChunk& output) override {
        if (mNext >= mOffsets.size()) return false;
        if (!readAudioChunk(mFile, mOffsets[mNext], output)) return false;
        ++mNext;
        return true;
    }

private:
    const FileView& mFile;
    const std::vector<BufferOffset>& mOffsets;
    size_t mNext = 0;
};

struct FrameGeometry { int width, height, compressionType; };

// A key the reference reads with operator[] (Decoder.cpp:161-167, :216-218).  Absent keys are undefined behaviour there
// (an assertion inside nlohmann::json, or a null dereference with NDEBUG); here they are an IOException.  Wrong types
// throw nlohmann's type_error on conversion, as in the reference.
const nlohmann::json& member(const nlohmann::json& object, const char* key, const char* what) {
    if (!object.is_object()) throw IOException(std::string("Invalid ") + what + " metadata");
    const auto it = object.find(key);
    if (it == object.end()) throw IOException(std::string("Invalid ") + what + " metadata (no \"" + key + "\")");
    return *it;
}

FrameGeometry geometryOf(const nlohmann::json& meta) {
    FrameGeometry g;
    g.width = member(meta, "width", "frame");      // Decoder.cpp:216-218
    g.height = member(meta, "height", "frame");
    g.compressionType = member(meta, "compressionType", "frame");
    // the callers size buffers with width*height (Decoder.cpp:221-222 does so unchecked)
    if (g.width <= 0 || g.height <= 0 || g.width > 65536 || g.height > 65536 ||
        static_cast<int64_t>(g.width) * g.height > (int64_t(1) << 30))
        throw IOException("Invalid frame metadata (width x height)");
    return g;
}

}  // namespace

struct Decoder::Impl {
    explicit Impl(FILE* f) : file(f) {}

    FileView file;
    nlohmann::json containerMetadata;
    std::vector<BufferOffset> frameIndex;               // sorted by timestamp
    std::vector<Timestamp> frameList;
    std::map<Timestamp, BufferOffset> frameByTimestamp; // first record wins for equal timestamps
    std::vector<BufferOffset> audioIndex;
    std::unique_ptr<AudioChunkLoader> audioLoader;
    std::vector<uint8_t> scratch;                       // compressed bytes of the frame being loaded
    // The batched calls run on a device context of their own (created on first use, device $MCRAW_B200_DEVICE), so that
    // the staging below lives and dies with the Decoder whatever thread destroys it.
    mcraw_ctx* batchCtx = nullptr;
    mcraw_ctx* batchContext() {
        if (!batchCtx) {
            int device = 0;
            if (const char* e = std::getenv("MCRAW_B200_DEVICE")) device = std::atoi(e);
            if (mcraw_ctx_create(device, &batchCtx) != MCRAW_OK) {
                batchCtx = nullptr;
                throw IOException(std::string("Failed to uncompress frame: ") + mcraw_last_error(nullptr));
            }
        }
        return batchCtx;
    }
    // staging of the batched calls, grown on demand and reused:
    // pinned ring for the compressed frames; for loadFrames also the decoded frames on the device and in pinned memory.
    // Two pinned result buffers alternate (outFlip), so the frames of one call can still be read while the next decodes.
    void* ring = nullptr;
    size_t ringBytes = 0;
    void* devOut = nullptr;
    size_t devOutBytes = 0;
    void* pinnedOut[2] = {nullptr, nullptr};
    size_t pinnedOutBytes[2] = {0, 0};
    int outFlip = 0;
    mcraw_ctx* ringCtx = nullptr;
    void releaseStaging() {
        if (ringCtx) {
            if (ring) mcraw_host_free_pinned(ringCtx, ring);
            for (void* p : pinnedOut)
                if (p) mcraw_host_free_pinned(ringCtx, p);
            if (devOut) mcraw_device_free(ringCtx, devOut);
        }
        ring = devOut = pinnedOut[0] = pinnedOut[1] = nullptr;
        ringBytes = devOutBytes = pinnedOutBytes[0] = pinnedOutBytes[1] = 0;
    }
    // make room for `in` bytes of compressed frames and (loadFrames only) `out` bytes of decoded frames
    void reserveStaging(mcraw_ctx* ctx, size_t in, size_t out) {
        if (ringCtx != ctx) { releaseStaging(); ringCtx = ctx; }
        auto fail = [&] {
            throw IOException(std::string("Cannot reserve staging memory for this many frames in one call (") + mcraw_last_error(ctx) +
                              "); request fewer frames per call");
        };
        if (in > ringBytes) {
            if (ring) mcraw_host_free_pinned(ctx, ring);
            ring = nullptr; ringBytes = 0;
            if (mcraw_host_alloc_pinned(ctx, in + in / 4, &ring) != MCRAW_OK) fail();
            ringBytes = in + in / 4;
        }
        if (out > devOutBytes) {
            if (devOut) mcraw_device_free(ctx, devOut);
            devOut = nullptr; devOutBytes = 0;
            if (mcraw_device_alloc(ctx, out + out / 4, &devOut) != MCRAW_OK) fail();
            devOutBytes = out + out / 4;
        }
        if (out > pinnedOutBytes[outFlip]) {           // only the result buffer of the current call
            void*& p = pinnedOut[outFlip];
            if (p) mcraw_host_free_pinned(ctx, p);
            p = nullptr; pinnedOutBytes[outFlip] = 0;
            if (mcraw_host_alloc_pinned(ctx, out + out / 8, &p) != MCRAW_OK) fail();
            pinnedOutBytes[outFlip] = out + out / 8;
        }
    }
    // ---- mapped feed (opt-in: MCRAW_FEED=mmap).  The whole file is mapped read-only and
    // page-locked once, and loadFramesToDevice hands the frames' addresses INSIDE the mapping to the H2D pipeline: the
    // DMA engine reads the page cache directly, the pread copy into the pinned ring disappears.  Falls back to the ring
    // (and says why in feedNote) when the mapping or the registration is refused, or the file is larger than
    // MCRAW_FEED_MMAP_MAX_MB (default 8192: registering pins every page of the file).
    void* map = nullptr;
    size_t mapBytes = 0;
    mcraw_ctx* mapCtx = nullptr;
    bool mapTried = false;
    std::string feedNote = "pread -> pinned ring";
    std::unique_ptr<Decoder::Feed> feed;
    const uint8_t* mappedFile(mcraw_ctx* ctx) {
        if (mapTried) return static_cast<const uint8_t*>(map);
        mapTried = true;
        const char* mode = std::getenv("MCRAW_FEED");
        if (!mode || std::string(mode) != "mmap") return nullptr;
        int64_t limitMb = 8192;
        if (const char* e = std::getenv("MCRAW_FEED_MMAP_MAX_MB")) limitMb = std::atoll(e);
        if (file.size() <= 0 || file.size() > limitMb * (int64_t(1) << 20)) {
            feedNote = "pread -> pinned ring (mmap feed: file size outside the limit)";
            return nullptr;
        }
        void* p = ::mmap(nullptr, static_cast<size_t>(file.size()), PROT_READ, MAP_SHARED, file.fd(), 0);
        if (p == MAP_FAILED) {
            feedNote = std::string("pread -> pinned ring (mmap failed: ") + std::strerror(errno) + ")";
            return nullptr;
        }
        if (mcraw_host_register(ctx, p, static_cast<size_t>(file.size()), 1) != MCRAW_OK) {
            feedNote = std::string("pread -> pinned ring (cudaHostRegister refused the mapping: ") + mcraw_last_error(ctx) + ")";
            ::munmap(p, static_cast<size_t>(file.size()));
            return nullptr;
        }
        map = p;
        mapBytes = static_cast<size_t>(file.size());
        mapCtx = ctx;
        feedNote = "mmap + cudaHostRegister (read-only): H2D straight from the page cache";
        return static_cast<const uint8_t*>(map);
    }
    void releaseMap() {
        if (!map) return;
        mcraw_host_unregister(mapCtx, map);
        ::munmap(map, mapBytes);
        map = nullptr;
        mapBytes = 0;
    }
    ~Impl() {
        releaseMap();
        releaseStaging();
        if (batchCtx) mcraw_ctx_destroy(batchCtx);
    }

    void open();
    void readFrameIndex();
    void findAudioIndex();
};

void Decoder::Impl::open() {
    // Decoder.cpp:116-151
    const int64_t at = file.start();
    Header header{};
    file.read(at, &header, sizeof header);
    if (header.version != CONTAINER_VERSION) throw IOException("Invalid container version");
    if (std::memcmp(header.ident, CONTAINER_ID, sizeof CONTAINER_ID) != 0) throw IOException("Invalid header id");
    Item item{};
    file.read(at + sizeof header, &item, sizeof item);
    if (item.type != Type::METADATA) throw IOException("Invalid camera metadata");
    file.require(at + sizeof header + sizeof item, item.size);
    std::string text(item.size, '\0');
    file.read(at + sizeof header + sizeof item, &text[0], item.size);
    containerMetadata = nlohmann::json::parse(text);

    readFrameIndex();
    findAudioIndex();
    audioLoader.reset(new SequentialAudioLoader(file, audioIndex));
}

void Decoder::Impl::readFrameIndex() {
    // the trailer is the last Item + BufferIndex of the file (Decoder.cpp:237-264)
    const int64_t trailer = static_cast<int64_t>(sizeof(Item) + sizeof(BufferIndex));
    if (file.size() < trailer) throw IOException("Failed to get end chunk");
    Item item{};
    file.read(file.size() - trailer, &item, sizeof item);
    if (item.type != Type::BUFFER_INDEX) throw IOException("Invalid file");
    BufferIndex index{};
    file.read(file.size() - static_cast<int64_t>(sizeof(BufferIndex)), &index, sizeof index);
    if (static_cast<uint32_t>(index.magicNumber) != INDEX_MAGIC_NUMBER) throw IOException("Corrupted file");
    if (index.numOffsets < 0 || index.indexDataOffset < 0) throw IOException("Invalid index");
    file.require(index.indexDataOffset, static_cast<uint64_t>(index.numOffsets) * sizeof(BufferOffset));
    frameIndex.resize(static_cast<size_t>(index.numOffsets));
    file.read(index.indexDataOffset, frameIndex.data(), sizeof(BufferOffset) * frameIndex.size());

    // timestamp order; equal timestamps keep index order and the first one is the one loadFrame finds (:266-279)
    std::stable_sort(frameIndex.begin(), frameIndex.end(),
                     [](const BufferOffset& a, const BufferOffset& b) { return a.timestamp < b.timestamp; });
    frameList.reserve(frameIndex.size());
    for (const BufferOffset& o : frameIndex) {
        frameList.push_back(o.timestamp);
        frameByTimestamp.insert({o.timestamp, o});
    }
}

void Decoder::Impl::findAudioIndex() {
    // The audio index is found by walking records forward from the frame with the largest timestamp
    // (Decoder.cpp:281-315); anything unexpected simply ends the walk.
    if (frameIndex.empty()) return;
    int64_t pos = frameIndex.back().offset;
    if (pos < 0) return;
    for (;;) {
        Item item{};
        if (!file.tryRead(pos, &item, sizeof item)) break;
        pos += sizeof item;
        if (item.type == Type::BUFFER || item.type == Type::METADATA || item.type == Type::AUDIO_DATA ||
            item.type == Type::AUDIO_DATA_METADATA) {
            pos += item.size;
        } else if (item.type == Type::AUDIO_INDEX) {
            AudioIndex index{};
            file.read(pos, &index, sizeof index);
            pos += sizeof index;
            if (index.numOffsets < 0) throw IOException("Failed to read data");
            if (static_cast<uint64_t>(index.numOffsets) > (1ull << 40)) throw IOException("Failed to read data");
            file.require(pos, static_cast<uint64_t>(index.numOffsets) * sizeof(BufferOffset));
            audioIndex.resize(static_cast<size_t>(index.numOffsets));
            file.read(pos, audioIndex.data(), sizeof(BufferOffset) * audioIndex.size());
            pos += static_cast<int64_t>(sizeof(BufferOffset) * audioIndex.size());
        } else {
            break;
        }
    }
}

Decoder::Decoder(FILE* file) {
    if (!file) throw IOException("Invalid file");
    m.reset(new Impl(file));
    try {
        m->open();
    } catch (...) {
        m->file.release();      // ownership passes to the Decoder only once it exists (Decoder.hpp)
        throw;
    }
}

Decoder::Decoder(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw IOException("Failed to open " + path);
    m.reset(new Impl(f));
    m->open();
}

Decoder::~Decoder() = default;

const std::vector<Timestamp>& Decoder::getFrames() const { return m->frameList; }
const nlohmann::json& Decoder::getContainerMetadata() const { return m->containerMetadata; }
int Decoder::audioSampleRateHz() const {          // :161-163
    return member(member(m->containerMetadata, "extraData", "camera"), "audioSampleRate", "camera");
}
int Decoder::numAudioChannels() const {           // :165-167
    return member(member(m->containerMetadata, "extraData", "camera"), "audioChannels", "camera");
}

void Decoder::loadAudio(std::vector<AudioChunk>& outAudioChunks) {
    for (const BufferOffset& o : m->audioIndex) {
        AudioChunk chunk;
        if (!readAudioChunk(m->file, o, chunk)) continue;
        outAudioChunks.emplace_back(std::move(chunk));
    }
}

AudioChunkLoader& Decoder::loadAudio() const { return *m->audioLoader; }

FrameLocation Decoder::locateFrame(const Timestamp timestamp) const {
    const auto it = m->frameByTimestamp.find(timestamp);
    if (it == m->frameByTimestamp.end())
        throw IOException("Frame not found (timestamp: " + std::to_string(timestamp) + ")");
    if (it->second.offset < 0) throw IOException("Invalid offset");
    Item item{};
    m->file.read(it->second.offset, &item, sizeof item);
    if (item.type != Type::BUFFER) throw IOException("Invalid buffer type");
    FrameLocation where;
    where.timestamp = timestamp;
    where.payloadOffset = it->second.offset + static_cast<int64_t>(sizeof item);
    where.payloadSize = item.size;
    m->file.require(where.payloadOffset, where.payloadSize);
    return where;
}

void Decoder::readFrame(const FrameLocation& where, uint8_t* dst, nlohmann::json& outMetadata) const {
    if (dst) m->file.read(where.payloadOffset, dst, where.payloadSize);     // dst == nullptr: metadata only
    int64_t pos = where.payloadOffset + where.payloadSize;
    Item item{};
    m->file.read(pos, &item, sizeof item);
    pos += sizeof item;
    if (item.type != Type::METADATA) throw IOException("Invalid metadata");
    m->file.require(pos, item.size);
    std::string text(item.size, '\0');
    m->file.read(pos, &text[0], item.size);
    outMetadata = nlohmann::json::parse(text);
}

namespace {

// Read many frames (payload into base + off[i], JSON into meta[i]) with a few threads: one pread stream tops out at a few
// GB/s of page-cache copy, far below what the H2D pipeline behind it can take.  positional reads: no shared state.
void readFramesParallel(const Decoder& dec, const std::vector<FrameLocation>& where, uint8_t* base, const std::vector<size_t>& off,
                        std::vector<nlohmann::json>& meta) {
    const size_t n = where.size();
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    size_t cap = 8;
    if (const char* e = std::getenv("MCRAW_READ_THREADS")) cap = std::max(1, std::atoi(e));
    const size_t nthreads = std::min<size_t>({n, cap, hw});
    if (nthreads <= 1) {
        for (size_t i = 0; i < n; i++) dec.readFrame(where[i], base + off[i], meta[i]);
        return;
    }
    std::atomic<size_t> next{0};
    std::exception_ptr failure;
    std::mutex failureLock;
    auto work = [&] {
        try {
            for (size_t i = next.fetch_add(1); i < n; i = next.fetch_add(1)) dec.readFrame(where[i], base + off[i], meta[i]);
        } catch (...) {
            std::lock_guard<std::mutex> g(failureLock);
            if (!failure) failure = std::current_exception();
            next.store(n);
        }
    };
    std::vector<std::thread> pool;
    for (size_t t = 1; t < nthreads; t++) pool.emplace_back(work);
    work();
    for (std::thread& t : pool) t.join();
    if (failure) std::rethrow_exception(failure);
}

}  // namespace

void Decoder::loadFrame(const Timestamp timestamp, std::vector<uint8_t>& outData, nlohmann::json& outMetadata) {
    const FrameLocation where = locateFrame(timestamp);
    m->scratch.resize(where.payloadSize);
    readFrame(where, m->scratch.data(), outMetadata);
    const FrameGeometry g = geometryOf(outMetadata);
    outData.resize(sizeof(uint16_t) * static_cast<size_t>(g.width) * static_cast<size_t>(g.height));   // :221-222
    uint16_t* pixels = reinterpret_cast<uint16_t*>(outData.data());
    if (g.compressionType == kCompressionCurrent) {
        if (raw::Decode(pixels, g.width, g.height, m->scratch.data(), m->scratch.size()) == 0)
            throw IOException("Failed to uncompress frame");
    } else if (g.compressionType == kCompressionLegacy) {
        if (raw::DecodeLegacy(pixels, g.width, g.height, m->scratch.data(), m->scratch.size()) == 0)
            throw IOException("Failed to uncompress legacy frame");
    } else {
        throw IOException("Invalid compression type");
    }
}

void Decoder::loadFrames(const std::vector<Timestamp>& timestamps, std::vector<std::vector<uint8_t>>& outData,
                         std::vector<nlohmann::json>& outMetadata) {
    // The caller owns the result vectors, so the pinned / device staging behind loadFramesPinned only ever has to hold a
    // bounded sub-batch: a whole clip (loadFrames(getFrames(), ...), tens of GB decoded) goes through ~1 GB at a time.
    const size_t n = timestamps.size();
    outData.assign(n, std::vector<uint8_t>());
    outMetadata.assign(n, nlohmann::json());
    constexpr size_t kSubBatchBytes = size_t(1) << 30;
    std::vector<Timestamp> part;
    std::vector<FrameView> views;
    std::vector<nlohmann::json> meta;
    for (size_t i = 0; i < n;) {
        // decoded size is only known from the frame's JSON; the compressed size (>= ~1/4 of it for real data) bounds the count
        part.clear();
        size_t bytes = 0;
        size_t j = i;
        for (; j < n && (j == i || bytes < kSubBatchBytes / 4) && j - i < 4096; j++) {
            bytes += locateFrame(timestamps[j]).payloadSize;
            part.push_back(timestamps[j]);
        }
        loadFramesPinned(part, views, meta);
        for (size_t k = 0; k < views.size(); k++) {
            outData[i + k].assign(views[k].data, views[k].data + views[k].size);
            outMetadata[i + k] = std::move(meta[k]);
        }
        i = j;
    }
}

void Decoder::loadFramesPinned(const std::vector<Timestamp>& timestamps, std::vector<FrameView>& outFrames,
                               std::vector<nlohmann::json>& outMetadata) {
    const size_t n = timestamps.size();
    outFrames.assign(n, FrameView{nullptr, 0});
    outMetadata.assign(n, nlohmann::json());
    if (n == 0) return;
    if (n > 0xFFFFFFFFull) throw IOException("Too many frames in one call");
    mcraw_ctx* ctx = m->batchContext();
    m->outFlip ^= 1;                                   // the previous call's frames stay where they are

    // ---- locate, size and read every frame straight into one pinned buffer (256-byte aligned slots)
    std::vector<FrameLocation> where(n);
    std::vector<size_t> inOff(n), outOff(n);
    size_t inBytes = 0;
    for (size_t i = 0; i < n; i++) {
        where[i] = locateFrame(timestamps[i]);
        inOff[i] = inBytes;
        inBytes += (static_cast<size_t>(where[i].payloadSize) + 255) & ~static_cast<size_t>(255);
    }
    m->reserveStaging(ctx, inBytes + 256, 0);
    uint8_t* in = static_cast<uint8_t*>(m->ring);
    std::vector<FrameGeometry> geo(n);
    size_t outBytes = 0;
    readFramesParallel(*this, where, in, inOff, outMetadata);
    for (size_t i = 0; i < n; i++) {
        geo[i] = geometryOf(outMetadata[i]);
        if (geo[i].compressionType != kCompressionCurrent && geo[i].compressionType != kCompressionLegacy)
            throw IOException("Invalid compression type");
        outOff[i] = outBytes;
        const size_t bytes = sizeof(uint16_t) * static_cast<size_t>(geo[i].width) * static_cast<size_t>(geo[i].height);
        outBytes += (bytes + 255) & ~static_cast<size_t>(255);
    }
    m->reserveStaging(ctx, inBytes + 256, outBytes + 256);

    // ---- one batched decode: H2D on the context's side streams overlaps the kernels
    std::vector<mcraw_frame_desc> descs(n);
    for (size_t i = 0; i < n; i++) {
        mcraw_frame_desc& d = descs[i];
        std::memset(&d, 0, sizeof d);
        d.src = in + inOff[i];
        d.len = where[i].payloadSize;
        d.width = geo[i].width;
        d.height = geo[i].height;
        d.compression_type = geo[i].compressionType;
        d.dst = reinterpret_cast<uint16_t*>(static_cast<uint8_t*>(m->devOut) + outOff[i]);
        d.dst_capacity_elems = static_cast<uint64_t>(geo[i].width) * static_cast<uint64_t>(geo[i].height);
    }
    // host in, host out: the pixels of a decoded chunk travel back while the next chunk is copied in and decoded
    uint8_t* const result = static_cast<uint8_t*>(m->pinnedOut[m->outFlip]);
    std::vector<uint16_t*> hostDst(n);
    for (size_t i = 0; i < n; i++) hostDst[i] = reinterpret_cast<uint16_t*>(result + outOff[i]);
    std::vector<uint64_t> written(n);
    if (mcraw_decode_batch_host_out(ctx, descs.data(), hostDst.data(), static_cast<uint32_t>(n), nullptr) != MCRAW_OK ||
        mcraw_batch_wait(ctx, written.data(), nullptr, static_cast<uint32_t>(n)) != MCRAW_OK)
        throw IOException(std::string("Failed to uncompress frame: ") + mcraw_last_error(ctx));
    for (size_t i = 0; i < n; i++) {
        if (written[i] == 0)
            throw IOException(geo[i].compressionType == kCompressionCurrent ? "Failed to uncompress frame"
                                                                             : "Failed to uncompress legacy frame");
    }
    for (size_t i = 0; i < n; i++) {
        const size_t bytes = sizeof(uint16_t) * static_cast<size_t>(geo[i].width) * static_cast<size_t>(geo[i].height);
        outFrames[i] = FrameView{result + outOff[i], bytes};
    }
}

const char* Decoder::feedDescription() const { return m->feedNote.c_str(); }

// ---- feeds of loadFramesToDevice ------------------------------------------------------------------------------------
// "ring" (default)  reader threads pread the frames into the pinned ring while the chunks already read travel to the device
//                    and decode: the file read of chunk c+1 overlaps the H2D copy and the kernels of chunk c.
// "direct"           the same, but the payload reads go through a second descriptor opened with O_DIRECT (the page cache is
//                    bypassed; every read covers the 4 KiB-aligned range around the frame and lands at the same misalignment
//                    in the ring).  Falls back to buffered reads where the filesystem refuses.
// "cufile"           GPUDirect Storage: cuFileRead straight into device memory, then the device-resident decode.  Needs
//                    libcufile and a driver / filesystem that take it; otherwise the reason is reported and the ring is used.
// "mmap"             H2D straight from a page-locked mapping of the file (see mappedFile()).
namespace {

struct CuFileApi {
    void* lib = nullptr;
    CUfileError_t (*driverOpen)() = nullptr;
    CUfileError_t (*handleRegister)(CUfileHandle_t*, CUfileDescr_t*) = nullptr;
    void (*handleDeregister)(CUfileHandle_t) = nullptr;
    ssize_t (*read)(CUfileHandle_t, void*, size_t, off_t, off_t) = nullptr;
    std::string why;
    bool load() {
        if (lib) return true;
        for (const char* name : {"libcufile.so.0", "libcufile.so", "/usr/local/cuda/lib64/libcufile.so.0"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (lib) break;
        }
        if (!lib) { why = "libcufile not found"; return false; }
        driverOpen = reinterpret_cast<decltype(driverOpen)>(dlsym(lib, "cuFileDriverOpen"));
        handleRegister = reinterpret_cast<decltype(handleRegister)>(dlsym(lib, "cuFileHandleRegister"));
        handleDeregister = reinterpret_cast<decltype(handleDeregister)>(dlsym(lib, "cuFileHandleDeregister"));
        read = reinterpret_cast<decltype(read)>(dlsym(lib, "cuFileRead"));
        if (!driverOpen || !handleRegister || !handleDeregister || !read) { why = "libcufile lacks the expected symbols"; dlclose(lib); lib = nullptr; return false; }
        return true;
    }
};

std::string cufileText(const CUfileError_t& e) {
    return std::string(cufileop_status_error(e.err)) + " (" + std::to_string(static_cast<int>(e.err)) + ")";
}

}  // namespace

struct Decoder::Feed {
    // O_DIRECT descriptor of the same file (feed "direct" / "cufile"), -1 when not open
    int directFd = -1;
    bool directTried = false;
    CuFileApi cufile;
    CUfileHandle_t cufileHandle{};
    bool cufileReady = false, cufileTried = false;
    void* devIn = nullptr;          // device copy of the compressed frames (feed "cufile")
    size_t devInBytes = 0;
    mcraw_ctx* devCtx = nullptr;
    static int deadlineSeconds() {
        if (const char* e = std::getenv("MCRAW_CUFILE_DEADLINE_S")) return std::max(1, std::atoi(e));
        return 10;
    }
    ~Feed() {
        if (cufileReady) cufile.handleDeregister(cufileHandle);
        if (devIn && devCtx) mcraw_device_free(devCtx, devIn);
        if (directFd >= 0) ::close(directFd);
    }
};

void Decoder::loadFramesToDevice(const std::vector<Timestamp>& timestamps, uint16_t* const* dst, const uint64_t* dstCapacityElems,
                                 std::vector<nlohmann::json>& outMetadata) {
    const size_t n = timestamps.size();
    outMetadata.assign(n, nlohmann::json());
    if (n == 0) return;
    if (n > 0xFFFFFFFFull) throw IOException("Too many frames in one call");
    mcraw_ctx* ctx = m->batchContext();
    if (!m->feed) m->feed.reset(new Feed());
    Feed& feed = *m->feed;
    const char* modeEnv = std::getenv("MCRAW_FEED");
    const std::string mode = modeEnv ? modeEnv : "ring";

    std::vector<FrameLocation> where(n);
    for (size_t i = 0; i < n; i++) where[i] = locateFrame(timestamps[i]);
    auto describe = [&](size_t i, mcraw_frame_desc& d, const uint8_t* src) {
        const FrameGeometry g = geometryOf(outMetadata[i]);
        if (g.compressionType != kCompressionCurrent && g.compressionType != kCompressionLegacy)
            throw IOException("Invalid compression type");
        std::memset(&d, 0, sizeof d);
        d.src = src;
        d.len = where[i].payloadSize;
        d.width = g.width;
        d.height = g.height;
        d.compression_type = g.compressionType;
        d.dst = dst[i];
        d.dst_capacity_elems = dstCapacityElems[i];
    };
    auto finish = [&](const std::vector<mcraw_frame_desc>& descs) {
        std::vector<uint64_t> written(n);
        if (mcraw_batch_wait(ctx, written.data(), nullptr, static_cast<uint32_t>(n)) != MCRAW_OK)
            throw IOException(std::string("Failed to uncompress frame: ") + mcraw_last_error(ctx));
        for (size_t i = 0; i < n; i++)
            if (written[i] == 0)
                throw IOException(descs[i].compression_type == kCompressionCurrent ? "Failed to uncompress frame"
                                                                                   : "Failed to uncompress legacy frame");
    };
    std::vector<mcraw_frame_desc> descs(n);

    // ---- mapped feed: nothing to read, the DMA engine takes the frames out of the page cache
    if (const uint8_t* mapped = m->mappedFile(ctx)) {
        for (size_t i = 0; i < n; i++) {
            readFrame(where[i], nullptr, outMetadata[i]);
            describe(i, descs[i], mapped + where[i].payloadOffset);
        }
        if (mcraw_decode_batch_host(ctx, descs.data(), static_cast<uint32_t>(n), nullptr) != MCRAW_OK)
            throw IOException(std::string("Failed to uncompress frame: ") + mcraw_last_error(ctx));
        finish(descs);
        return;
    }

    // ---- O_DIRECT descriptor (feeds "direct" and "cufile")
    if ((mode == "direct" || mode == "cufile") && !feed.directTried) {
        feed.directTried = true;
        const std::string link = "/proc/self/fd/" + std::to_string(m->file.fd());
        feed.directFd = ::open(link.c_str(), O_RDONLY | O_DIRECT);
        if (feed.directFd < 0) m->feedNote = std::string("pread -> pinned ring (O_DIRECT open refused: ") + std::strerror(errno) + ")";
    }

    // ---- cuFile feed: file -> device memory, device-resident decode.  Everything that calls into libcufile runs on a helper
    //      thread under a deadline: on platforms without the nvidia-fs driver the library's compatibility mode has been seen
    //      to block for good (profiles/README.md), and a feed that was asked for must never cost more than its deadline.
    if (mode == "cufile" && !feed.cufileTried) {
        feed.cufileTried = true;
        if (feed.directFd >= 0) {
            struct Probe { std::mutex m; std::condition_variable cv; bool done = false, ok = false; std::string note; CuFileApi api; CUfileHandle_t h{}; };
            auto probe = std::make_shared<Probe>();
            const int fd = feed.directFd;
            std::thread([probe, fd, ctx] {
                std::string note;
                bool ok = false;
                mcraw_stream_sync(ctx, nullptr);                       // binds the device to this thread
                if (!probe->api.load()) note = "cuFile: " + probe->api.why;
                else {
                    CUfileError_t e = probe->api.driverOpen();
                    if (e.err != CU_FILE_SUCCESS) note = "cuFileDriverOpen: " + cufileText(e);
                    else {
                        CUfileDescr_t descr;
                        std::memset(&descr, 0, sizeof descr);
                        descr.handle.fd = fd;
                        descr.type = CU_FILE_HANDLE_TYPE_OPAQUE_FD;
                        e = probe->api.handleRegister(&probe->h, &descr);
                        if (e.err != CU_FILE_SUCCESS) note = "cuFileHandleRegister: " + cufileText(e);
                        else ok = true;
                    }
                }
                std::lock_guard<std::mutex> g(probe->m);
                probe->ok = ok; probe->note = note; probe->done = true;
                probe->cv.notify_all();
            }).detach();
            std::unique_lock<std::mutex> g(probe->m);
            if (!probe->cv.wait_for(g, std::chrono::seconds(feed.deadlineSeconds()), [&] { return probe->done; })) {
                m->feedNote = "pread -> pinned ring (cuFile: no answer from cuFileDriverOpen / cuFileHandleRegister within " +
                              std::to_string(feed.deadlineSeconds()) + " s)";
            } else if (!probe->ok) {
                m->feedNote = "pread -> pinned ring (" + probe->note + ")";
            } else {
                feed.cufile = probe->api;
                feed.cufileHandle = probe->h;
                feed.cufileReady = true;
                m->feedNote = "cuFileRead -> device memory (GPUDirect Storage or its compatibility mode)";
            }
        }
    }
    if (mode == "cufile" && feed.cufileReady) {
        std::vector<size_t> off(n);
        size_t bytes = 0;
        for (size_t i = 0; i < n; i++) { off[i] = bytes; bytes += (static_cast<size_t>(where[i].payloadSize) + 4095) & ~static_cast<size_t>(4095); }
        if (bytes > feed.devInBytes || feed.devCtx != ctx) {
            if (feed.devIn && feed.devCtx) mcraw_device_free(feed.devCtx, feed.devIn);
            feed.devIn = nullptr; feed.devInBytes = 0; feed.devCtx = ctx;
            if (mcraw_device_alloc(ctx, bytes + bytes / 4, &feed.devIn) != MCRAW_OK) throw IOException(mcraw_last_error(ctx));
            feed.devInBytes = bytes + bytes / 4;
        }
        struct Reads { std::mutex m; std::condition_variable cv; bool done = false, ok = false; };
        auto reads = std::make_shared<Reads>();
        {
            const CuFileApi api = feed.cufile;
            const CUfileHandle_t h = feed.cufileHandle;
            void* devIn = feed.devIn;
            std::vector<FrameLocation> w = where;
            std::vector<size_t> o = off;
            std::thread([reads, api, h, devIn, w, o, ctx] {
                mcraw_stream_sync(ctx, nullptr);
                bool ok = true;
                for (size_t i = 0; i < w.size() && ok; i++)
                    ok = api.read(h, devIn, w[i].payloadSize, static_cast<off_t>(w[i].payloadOffset), static_cast<off_t>(o[i])) ==
                         static_cast<ssize_t>(w[i].payloadSize);
                std::lock_guard<std::mutex> g(reads->m);
                reads->ok = ok; reads->done = true;
                reads->cv.notify_all();
            }).detach();
        }
        bool good = false;
        {
            std::unique_lock<std::mutex> g(reads->m);
            const bool answered = reads->cv.wait_for(g, std::chrono::seconds(feed.deadlineSeconds() + static_cast<int>(bytes >> 28)), [&] { return reads->done; });
            good = answered && reads->ok;
            if (!answered) m->feedNote = "pread -> pinned ring (cuFileRead: no answer within the deadline)";
            else if (!reads->ok) m->feedNote = "pread -> pinned ring (cuFileRead failed)";
        }
        if (!good) {
            feed.cufileReady = false;          // from now on: the ring.  (A helper that never answers keeps its own copies of what it uses.)
            if (feed.devIn) { feed.devIn = nullptr; feed.devInBytes = 0; }   // still the helper's target: never reused, never freed
        } else {
            for (size_t i = 0; i < n; i++) {
                readFrame(where[i], nullptr, outMetadata[i]);
                describe(i, descs[i], static_cast<const uint8_t*>(feed.devIn) + off[i]);
            }
            if (mcraw_decode_batch(ctx, descs.data(), static_cast<uint32_t>(n), nullptr) != MCRAW_OK)
                throw IOException(std::string("Failed to uncompress frame: ") + mcraw_last_error(ctx));
            finish(descs);
            return;
        }
    }

    // ---- ring feeds: lay the ring out (256-byte aligned slots back to back -- one H2D copy per chunk; O_DIRECT reads need
    //      4 KiB slots and keep the frame's misalignment), cut it into chunks, read and enqueue chunk by chunk
    const bool direct = mode == "direct" && feed.directFd >= 0;
    std::vector<size_t> off(n);
    size_t bytes = 0;
    for (size_t i = 0; i < n; i++) {
        if (direct) {
            const size_t mis = static_cast<size_t>(where[i].payloadOffset) & 4095;
            off[i] = bytes + mis;
            bytes += (mis + static_cast<size_t>(where[i].payloadSize) + 4095) & ~static_cast<size_t>(4095);
        } else {
            off[i] = bytes;
            bytes += (static_cast<size_t>(where[i].payloadSize) + 255) & ~static_cast<size_t>(255);
        }
    }
    m->reserveStaging(ctx, bytes + 8192, 0);
    uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(m->ring) + 4095) & ~static_cast<uintptr_t>(4095));
    size_t kChunkBytes = size_t(32) << 20;                // a chunk is handed to the device as soon as it has been read
    if (const char* e = std::getenv("MCRAW_FEED_CHUNK_KB")) kChunkBytes = static_cast<size_t>(std::max(1, std::atoi(e))) << 10;
    std::vector<size_t> chunkFirst{0};
    {
        size_t acc = 0;
        for (size_t i = 0; i < n; i++) {
            if (acc >= kChunkBytes) { chunkFirst.push_back(i); acc = 0; }
            acc += where[i].payloadSize;
        }
        chunkFirst.push_back(n);
    }
    const size_t nchunks = chunkFirst.size() - 1;
    std::vector<std::atomic<size_t>> remaining(nchunks);
    std::vector<uint32_t> chunkOf(n);
    for (size_t c = 0; c < nchunks; c++) {
        remaining[c].store(chunkFirst[c + 1] - chunkFirst[c]);
        for (size_t i = chunkFirst[c]; i < chunkFirst[c + 1]; i++) chunkOf[i] = static_cast<uint32_t>(c);
    }
    std::mutex lock;
    std::condition_variable ready;
    std::exception_ptr failure;
    std::atomic<size_t> next{0};
    std::atomic<bool> directFailed{false};
    auto readOne = [&](size_t i) {
        bool done = false;
        if (direct && !directFailed) {
            const int64_t a = where[i].payloadOffset & ~int64_t(4095);
            const size_t span = (static_cast<size_t>(where[i].payloadOffset - a) + where[i].payloadSize + 4095) & ~static_cast<size_t>(4095);
            uint8_t* p = ring + off[i] - static_cast<size_t>(where[i].payloadOffset - a);
            size_t got = 0;
            while (got < span) {
                const ssize_t r = ::pread(feed.directFd, p + got, span - got, static_cast<off_t>(a + static_cast<int64_t>(got)));
                if (r <= 0) break;
                got += static_cast<size_t>(r);
            }
            // the last block of the file may be short: enough is what covers the frame
            done = got >= static_cast<size_t>(where[i].payloadOffset - a) + where[i].payloadSize;
            if (!done) directFailed = true;
        }
        readFrame(where[i], done ? nullptr : ring + off[i], outMetadata[i]);
    };
    auto work = [&] {
        try {
            for (size_t i = next.fetch_add(1); i < n; i = next.fetch_add(1)) {
                readOne(i);
                if (remaining[chunkOf[i]].fetch_sub(1) == 1) {
                    std::lock_guard<std::mutex> g(lock);
                    ready.notify_all();
                }
            }
        } catch (...) {
            std::lock_guard<std::mutex> g(lock);
            if (!failure) failure = std::current_exception();
            next.store(n);
            ready.notify_all();
        }
    };
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    // a reader is a memcpy out of the page cache: three quarters of the host's threads (at most 16) is where the feed tops
    // out (16-thread host, 64 C3 frames in tmpfs: 8 / 12 / 16 / 24 readers -> 42.8 / 49.3 / 44.9 / 42.1 Gpix/s)
    size_t cap = std::max<size_t>(4, std::min<size_t>(16, (3 * static_cast<size_t>(hw)) / 4));
    if (const char* e = std::getenv("MCRAW_READ_THREADS")) cap = std::max(1, std::atoi(e));
    const size_t nthreads = std::max<size_t>(1, std::min<size_t>({n, cap, hw}));
    std::vector<std::thread> pool;
    for (size_t t = 0; t < nthreads; t++) pool.emplace_back(work);
    struct Joiner {
        std::vector<std::thread>& p;
        ~Joiner() { for (std::thread& t : p) if (t.joinable()) t.join(); }
    } joiner{pool};

    if (mcraw_batch_begin(ctx, static_cast<uint32_t>(n)) != MCRAW_OK) throw IOException(mcraw_last_error(ctx));
    for (size_t c = 0; c < nchunks; c++) {
        {
            std::unique_lock<std::mutex> g(lock);
            ready.wait(g, [&] { return failure || remaining[c].load() == 0; });
            if (failure) break;
        }
        const size_t a = chunkFirst[c], b = chunkFirst[c + 1];
        for (size_t i = a; i < b; i++) describe(i, descs[i], ring + off[i]);
        if (mcraw_batch_append_host(ctx, descs.data() + a, static_cast<uint32_t>(a), static_cast<uint32_t>(b - a), nullptr) != MCRAW_OK) {
            std::lock_guard<std::mutex> g(lock);
            if (!failure) failure = std::make_exception_ptr(IOException(std::string("Failed to uncompress frame: ") + mcraw_last_error(ctx)));
            next.store(n);
            break;
        }
    }
    for (std::thread& t : pool) t.join();
    if (failure) {
        mcraw_batch_wait(ctx, nullptr, nullptr, 0);      // whatever was enqueued still reads the ring: let it finish
        std::rethrow_exception(failure);
    }
    if (direct) m->feedNote = directFailed ? "pread -> pinned ring (O_DIRECT read refused by the filesystem; buffered reads)"
                                           : "O_DIRECT pread -> pinned ring, read of chunk c+1 overlapping H2D + decode of chunk c";
    else if (mode == "ring" || mode == "") m->feedNote = "pread -> pinned ring, read of chunk c+1 overlapping H2D + decode of chunk c";
    finish(descs);
}

}  // namespace motioncam

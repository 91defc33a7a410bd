// motioncam/Container.hpp -- on-disk records of the .mcraw container, version 3.
//
// Same names and byte layout as the reference's lib/include/motioncam/Container.hpp:23-71 (the file format is
// the contract; these structs are read straight from the file, little-endian hosts only), so code written
// against the reference header compiles unchanged against this one.
#pragma once
#include <cstdint>

namespace motioncam {

constexpr uint32_t INDEX_MAGIC_NUMBER = 0x8A905612u;   // BufferIndex::magicNumber
constexpr uint8_t CONTAINER_VERSION = 3;
constexpr uint8_t CONTAINER_ID[7] = {'M', 'O', 'T', 'I', 'O', 'N', ' '};

enum VideoType { VIDEO, TIMELAPSE };

// Item::type.  Every record of the file is an Item header followed by Item::size payload bytes.
enum class Type : uint32_t {
    BUFFER_INDEX = 0,        // trailer: BufferIndex, the last 24 bytes of the file
    BUFFER_INDEX_DATA = 1,   // numOffsets x BufferOffset
    BUFFER = 2,              // one compressed frame
    METADATA = 3,            // JSON (container metadata after the Header, frame metadata after each BUFFER)
    AUDIO_INDEX = 4,         // AudioIndex + numOffsets x BufferOffset
    AUDIO_DATA = 5,          // int16 PCM
    AUDIO_DATA_METADATA = 6  // AudioMetadata (optional, follows AUDIO_DATA in newer files)
};

struct Header { uint8_t ident[7]; uint8_t version; };
struct Item { Type type; uint32_t size; };
struct BufferOffset { int64_t offset; int64_t timestamp; };
struct BufferIndex { int32_t magicNumber; int32_t numOffsets; int64_t indexDataOffset; };
struct AudioIndex { int64_t numOffsets; int64_t startTimestampMs; };
struct AudioMetadata { int64_t timestampNs; };

static_assert(sizeof(Header) == 8 && sizeof(Item) == 8 && sizeof(BufferOffset) == 16 && sizeof(BufferIndex) == 16 &&
              sizeof(AudioIndex) == 16 && sizeof(AudioMetadata) == 8, ".mcraw records are read straight into these structs");

}  // namespace motioncam

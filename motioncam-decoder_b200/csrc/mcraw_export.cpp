// mcraw_export -- the reference's example program (/root/reference/example.cpp:141-203) on the B200 path:
//
//     mcraw_export <input file> [-n number of frames to export] [--out DIR] [--batch N] [--threads T] [--no-audio]
//
// writes audio.wav and frame_%06d.dng (byte-identical to the reference program's files) into the current directory
// or DIR.  Frames are decoded in batches on the GPU (Decoder::loadFrames) while writer threads package the previous
// batch; there is no CPU decode path -- without a B200 the first batch fails with the CUDA error.
#include <motioncam/Export.hpp>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

int main(int argc, const char* argv[]) {
    if (argc < 2) {
        std::printf("Usage: mcraw_export <input file> [-n number of frames to export] [--out DIR] [--batch N] [--threads T] [--no-audio]\n");
        return -1;
    }
    const std::string inputPath(argv[1]);
    std::string outputDir;
    motioncam::ExportOptions options;
    for (int i = 2; i < argc; i++) {
        const std::string a(argv[i]);
        const bool hasValue = i + 1 < argc;
        if (a == "-n" && hasValue) options.numFrames = std::atoi(argv[++i]);
        else if (a == "--out" && hasValue) outputDir = argv[++i];
        else if (a == "--batch" && hasValue) options.batch = std::atoi(argv[++i]);
        else if (a == "--threads" && hasValue) options.writerThreads = std::atoi(argv[++i]);
        else if (a == "--no-audio") options.writeAudio = false;
        else {
            std::fprintf(stderr, "Error: unknown argument %s\n", a.c_str());
            return -1;
        }
    }
    if (options.numFrames < 0) options.numFrames = -1;
    try {
        motioncam::exportClip(inputPath, outputDir, options, stdout);
    } catch (const motioncam::MotionCamException& e) {
        std::fprintf(stderr, "Error: %s\n", e.what());
        return -1;
    }
    return 0;
}

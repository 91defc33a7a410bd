// mcraw_export -- the reference's example program (/root/reference/example.cpp:141-203) on the B200 path:
//
//     mcraw_export <input file> [-n number of frames to export] [--out DIR] [--batch N] [--threads T] [--no-audio] [--stats]
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
        std::printf("Usage: mcraw_export <input file> [-n number of frames to export] [--out DIR] [--batch N] [--threads T] [--no-audio] [--stats]\n");
        return -1;
    }
    const std::string inputPath(argv[1]);
    std::string outputDir;
    motioncam::ExportOptions options;
    bool wantStats = false;
    for (int i = 2; i < argc; i++) {
        const std::string a(argv[i]);
        const bool hasValue = i + 1 < argc;
        if (a == "-n" && hasValue) options.numFrames = std::atoi(argv[++i]);
        else if (a == "--out" && hasValue) outputDir = argv[++i];
        else if (a == "--batch" && hasValue) options.batch = std::atoi(argv[++i]);
        else if (a == "--threads" && hasValue) options.writerThreads = std::atoi(argv[++i]);
        else if (a == "--no-audio") options.writeAudio = false;
        else if (a == "--stats") wantStats = true;
        else {
            std::fprintf(stderr, "Error: unknown argument %s\n", a.c_str());
            return -1;
        }
    }
    if (options.numFrames < 0) options.numFrames = -1;
    try {
        motioncam::ExportStats st;
        motioncam::exportClip(inputPath, outputDir, options, stdout, &st);
        if (wantStats)
            std::fprintf(stderr, "frames %zu  total %.3f s  open+audio %.3f  decode %.3f (first batch %.3f)  writer wait %.3f  steady %.3f s = %.1f frames/s\n",
                         st.frames, st.totalSeconds, st.openAndAudioSeconds, st.decodeSeconds, st.firstBatchSeconds, st.writerWaitSeconds,
                         st.steadySeconds, st.steadySeconds > 0 ? static_cast<double>(st.frames) / st.steadySeconds : 0.0);
    } catch (const motioncam::MotionCamException& e) {
        std::fprintf(stderr, "Error: %s\n", e.what());
        return -1;
    }
    return 0;
}

#!/usr/bin/env bash
# round-2 GPU call 29: file -> device feed (C5) against the number of reader threads and the chunk size
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2q}
nproc
{
for t in 8 12 16 24; do MCRAW_READ_THREADS=$t python tools/mcraw_file_e2e.py --frames 64 --reps 5 | cut -c1-260; done
MCRAW_READ_THREADS=16 MCRAW_FEED_CHUNK_KB=16384 python tools/mcraw_file_e2e.py --frames 64 --reps 5 | cut -c1-260
MCRAW_READ_THREADS=16 MCRAW_FEED_CHUNK_KB=65536 python tools/mcraw_file_e2e.py --frames 64 --reps 5 | cut -c1-260
MCRAW_READ_THREADS=16 MCRAW_FEED=direct python tools/mcraw_file_e2e.py --frames 64 --reps 5 | cut -c1-260
} > gpurun_out/${T}_feed_threads.txt 2>gpurun_out/${T}_feed_threads.err
cat gpurun_out/${T}_feed_threads.txt
echo done

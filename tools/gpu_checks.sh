#!/usr/bin/env bash
# gpu_checks.sh -- the GPU-side evidence of profiles/, as one script (run on the B200 box, e.g.
#   gpurun --timeout 900 -- 'bash tools/gpu_checks.sh all'
# ); every step writes into gpurun_out/.  Steps: tests | bench | launches | ncu | sanitizers | cross | all
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
step="${1:-all}"
want() { [ "$step" = all ] || [ "$step" = "$1" ]; }

if want tests; then
    python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt
fi
if want bench; then
    python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
    python bench.py --impl reference > gpurun_out/bench_c2_reference.json 2> gpurun_out/bench_c2_reference.err
    for wl in c1 c3 c4; do python bench.py --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; done
    cat gpurun_out/bench_c2.json
fi
if want launches; then   # per-launch durations (cold cache, serialised): shares of the step, not absolutes
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv \
        python bench.py --steps 2 --warmup 3 > gpurun_out/launches_c2.log 2>&1
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv \
        python bench.py --workload c4 --steps 2 --warmup 3 > gpurun_out/launches_c4.log 2>&1
fi
if want ncu; then        # one full capture per kernel; read with ncu -i ... --page raw --csv and tools/ncu_lines.py
    for k in k_units k_meta k_legacy_warp; do
        wl=c2; [ $k = k_legacy_warp ] && wl=c4
        ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/ncu_$k \
            python bench.py --workload $wl --steps 2 --warmup 3 > gpurun_out/ncu_$k.log 2>&1
    done
fi
if want sanitizers; then # small inputs: the sanitizers slow the kernels down by two to three orders of magnitude
    sel="tests/test_golden.py tests/test_gpu_current.py tests/test_gpu_legacy.py tests/test_gpu_fuzz.py"
    for tool in memcheck racecheck synccheck; do
        compute-sanitizer --tool $tool python -m pytest $sel -m gpu -q \
            -k "golden or vectors_batched or rejects or mixed or fuzz" \
            > gpurun_out/sanitizer_$tool.txt 2>&1
        tail -2 gpurun_out/sanitizer_$tool.txt
    done
fi
if want cross; then      # the cross-batch experiment: timing sweep, then racecheck with the switch on
    for v in 0 16 24 32 48; do
        if [ $v = 0 ]; then python tools/c2_steps.py --label base; else MCRAW_CROSS_BATCH=$v python tools/c2_steps.py --label cross$v; fi
    done > gpurun_out/c2_cross.jsonl 2> gpurun_out/c2_cross.err
    cat gpurun_out/c2_cross.jsonl
    MCRAW_CROSS_BATCH=24 compute-sanitizer --tool racecheck python tools/c2_steps.py --frames 4 --steps 8 --label racecheck \
        > gpurun_out/sanitizer_racecheck_cross.txt 2>&1
    tail -2 gpurun_out/sanitizer_racecheck_cross.txt
fi

#!/usr/bin/env bash
# asan_host_fuzz.sh -- the host side of the drop-in library (container reader, export writers, C wrappers) built with
# AddressSanitizer + UndefinedBehaviorSanitizer and driven by the container fuzz walk (tests/helpers/container_fuzz.py)
# and the CPU container / export tests.  No GPU needed: decode calls end in an IOException after the read + JSON parse.
#   bash tools/asan_host_fuzz.sh [mutants per seed]      -> profiles/r2_asan_ubsan_host.txt
set -u
cd "$(dirname "$0")/.."
N="${1:-800}"
OUT=/tmp/mcasan
mkdir -p "$OUT"
PKG=motioncam-decoder_b200
JSON_INC=$(python -c "import sysconfig,os;print(os.path.join(sysconfig.get_paths()['purelib'],'include','cudnn_frontend','thirdparty'))")
g++ -std=gnu++17 -O1 -g -fno-omit-frame-pointer -fsanitize=address,undefined -fno-sanitize-recover=undefined -fPIC -shared -Wall \
    -Iinclude -I"$JSON_INC" -I/usr/local/cuda/include -o "$OUT/libmotioncam_decoder_b200_asan.so" \
    $PKG/csrc/RawData.cpp $PKG/csrc/Decoder.cpp $PKG/csrc/Export.cpp $PKG/csrc/cwrap.cpp \
    -pthread -ldl -L$PKG -lmcraw_b200 -Wl,-rpath,"$(pwd)/$PKG" -Wl,-Bsymbolic || exit 1
ASAN=$(g++ -print-file-name=libasan.so)
UBSAN=$(g++ -print-file-name=libubsan.so)
LOG=profiles/r2_asan_ubsan_host.txt
{
echo "# $(date -u +%FT%TZ)  g++ $(g++ -dumpversion)  -fsanitize=address,undefined -fno-sanitize-recover=undefined"
echo "# library: RawData.cpp Decoder.cpp Export.cpp cwrap.cpp (host side of libmotioncam_decoder_b200.so)"
for seed in 2 3 5 7; do
  echo "## container fuzz walk, seed $seed, $N mutants"
  MCRAW_FUZZ_NO_RLIMIT=1 MCRAW_DROPIN_LIB="$OUT/libmotioncam_decoder_b200_asan.so" LD_PRELOAD="$ASAN:$UBSAN" \
    ASAN_OPTIONS=detect_leaks=0:abort_on_error=1 UBSAN_OPTIONS=print_stacktrace=1 \
    python tests/helpers/container_fuzz.py $seed $N /tmp 2>"$OUT/fuzz_$seed.err" | tail -1
  echo "   exit code ${PIPESTATUS[0]}; sanitizer reports: $(grep -c -E 'ERROR: AddressSanitizer|runtime error' "$OUT/fuzz_$seed.err")"
done
echo "## CPU container + export tests against the sanitized library"
MCRAW_DROPIN_LIB="$OUT/libmotioncam_decoder_b200_asan.so" LD_PRELOAD="$ASAN:$UBSAN" ASAN_OPTIONS=detect_leaks=0:abort_on_error=1 \
    python -m pytest tests/test_container_cpu.py tests/test_export_cpu.py -q -m "not gpu" -p no:cacheprovider 2>&1 | tail -3
} | tee "$LOG"

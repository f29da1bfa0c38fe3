#!/bin/bash
# Host-side C++ (gsbd, the HTTP/2 layer) under ThreadSanitizer and Address/UB/LeakSanitizer, driven by the
# regular test files. No GPU needed: gsbd runs with --fake-inventory in those tests.
#   tools/sanitize.sh            -> build/san/*.log, exit 1 if any sanitizer reported anything
set -uo pipefail
cd "$(dirname "$0")/.."
PKG=gpushare_device_plugin_b200
SRC=$PKG/csrc/daemon
OUT=build/san
mkdir -p $OUT && rm -f $OUT/*
LINK="-L$PKG -lgpushare_b200 -lssl -lcrypto -ldl -Wl,-rpath,$PWD/$PKG"
g++ -O1 -g -std=c++17 -pthread -fsanitize=thread -o $OUT/gsbd.tsan $SRC/gsbd.cc $LINK || exit 2
g++ -O1 -g -std=c++17 -pthread -fsanitize=address,undefined -fno-omit-frame-pointer -o $OUT/gsbd.asan $SRC/gsbd.cc $LINK || exit 2
g++ -O1 -g -std=c++17 -pthread -fsanitize=thread -o $OUT/h2.tsan $SRC/h2_selftest.cc || exit 2
g++ -O1 -g -std=c++17 -pthread -fsanitize=address,undefined -fno-omit-frame-pointer -o $OUT/h2.asan $SRC/h2_selftest.cc || exit 2
export TSAN_OPTIONS="log_path=$PWD/$OUT/tsan halt_on_error=0 second_deadlock_stack=1"
export ASAN_OPTIONS="log_path=$PWD/$OUT/asan detect_leaks=1"
export UBSAN_OPTIONS="log_path=$PWD/$OUT/ubsan print_stacktrace=1"
rc=0
GSBD_BINARY=$PWD/$OUT/gsbd.tsan python -m pytest tests/test_native_daemon.py tests/test_tls.py -q -p no:cacheprovider > $OUT/pytest_gsbd_tsan.log 2>&1 || rc=1
GSBD_BINARY=$PWD/$OUT/gsbd.asan python -m pytest tests/test_native_daemon.py tests/test_tls.py -q -p no:cacheprovider > $OUT/pytest_gsbd_asan.log 2>&1 || rc=1
H2_SELFTEST_BINARY=$PWD/$OUT/h2.tsan python -m pytest tests/test_native_h2.py -q -p no:cacheprovider > $OUT/pytest_h2_tsan.log 2>&1 || rc=1
H2_SELFTEST_BINARY=$PWD/$OUT/h2.asan python -m pytest tests/test_native_h2.py -q -p no:cacheprovider > $OUT/pytest_h2_asan.log 2>&1 || rc=1
# the library's own host code (event queue, lifecycle locks) instrumented too, driven by tests/native/health_stress.cc
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
mkdir -p $OUT/lib
NF="-gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 -Xcompiler -fPIC,-fsanitize=thread -cudart static"
$NVCC $NF -c $PKG/csrc/hbm_probe_sm100a.cu -o $OUT/lib/probe.o > /dev/null 2>&1 || exit 2
$NVCC $NF -c $PKG/csrc/gsb_device.cu -o $OUT/lib/dev.o > /dev/null 2>&1 || exit 2
g++ -O1 -g -std=c++17 -fPIC -fsanitize=thread -c $PKG/csrc/gsb_wire.cc -o $OUT/lib/wire.o || exit 2
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -cudart static -o $OUT/lib/libgpushare_b200.so $OUT/lib/probe.o $OUT/lib/dev.o \
  $OUT/lib/wire.o -ldl -lpthread -lrt -Xlinker -ltsan || exit 2
g++ -O1 -g -std=c++17 -pthread -fsanitize=thread -Iinclude -o $OUT/health_stress.tsan tests/native/health_stress.cc \
  -L$OUT/lib -lgpushare_b200 -Wl,-rpath,$PWD/$OUT/lib || exit 2
$OUT/health_stress.tsan 3 > $OUT/health_stress_tsan.log 2>&1 || rc=1
echo "health_stress (TSan library): $(tail -1 $OUT/health_stress_tsan.log)"
for f in $OUT/pytest_*.log; do echo "$(basename $f): $(grep -E 'passed|failed' $f | tail -1)"; done
reports=$(ls $OUT | grep -E '^(tsan|asan|ubsan)\.' | wc -l)
echo "sanitizer report files: $reports"
[ "$reports" -eq 0 ] || { cat $OUT/tsan.* $OUT/asan.* $OUT/ubsan.* 2>/dev/null | grep -E "SUMMARY|runtime error" | sort | uniq -c; rc=1; }
exit $rc

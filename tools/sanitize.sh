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
for f in $OUT/pytest_*.log; do echo "$(basename $f): $(grep -E 'passed|failed' $f | tail -1)"; done
reports=$(ls $OUT | grep -E '^(tsan|asan|ubsan)\.' | wc -l)
echo "sanitizer report files: $reports"
[ "$reports" -eq 0 ] || { cat $OUT/tsan.* $OUT/asan.* $OUT/ubsan.* 2>/dev/null | grep -E "SUMMARY|runtime error" | sort | uniq -c; rc=1; }
exit $rc

#!/bin/sh
# Builds tools/tsan_kernels.cpp with ThreadSanitizer and runs it (CPU only, a few minutes).  Exit code 0 and no
# "WARNING: ThreadSanitizer" lines = no unordered shared-memory access pair in the kernels on these workloads.
set -e
cd "$(dirname "$0")/.."
g++ -O1 -g -std=c++17 -pthread -fsanitize=thread -Wno-unknown-pragmas tools/tsan_kernels.cpp exacto_b200/csrc/host_setup.cpp -o /tmp/exb_tsan_kernels
TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0" /tmp/exb_tsan_kernels

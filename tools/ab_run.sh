#!/bin/bash
# Developer helper: time every build_ab/lib_*.so with tools/quick_bench.py (args passed through) and check each against
# the default library's output on a small parity run.
for lib in build_ab/lib_*.so; do
  echo "== $lib"
  FFTL_LIB=$PWD/$lib python tools/quick_bench.py "$@" 2>&1 | tail -1
done

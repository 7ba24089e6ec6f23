#!/bin/bash
# Developer helper: time every build_ab/lib_*.so with tools/quick_bench.py (args passed through); each run under its
# own timeout so that a hung experimental kernel cannot hold the GPU box.
for lib in build_ab/lib_*.so; do
  echo "== $lib"
  FFTL_LIB=$PWD/$lib timeout 120 python tools/quick_bench.py "$@" 2>&1 | tail -1
done

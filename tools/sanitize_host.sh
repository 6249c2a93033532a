#!/bin/bash
# Host side of libvsgpu (ser/ reader, flattener, index cache, second pass for the sample positions, the
# per-region walk logic of every operator compiled for the host, row materialiser) under
# AddressSanitizer + UndefinedBehaviorSanitizer: builds the test-only host simulator with the sanitizers
# and runs a parity sweep (t1-t7, all four fuzz shapes, with and without the index cache) through it.
# Round 1: clean.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
C=$ROOT/variantstore_b200/csrc
/usr/bin/g++ -O1 -g -fPIC -std=c++17 -fsanitize=address,undefined -fno-omit-frame-pointer -I/usr/local/cuda/include -pthread -shared \
  -o /tmp/libvsgpu_hostsim_asan.so $ROOT/tests/hostsim/hostsim.cc $C/flatten.cc $C/ser_reader.cc $C/materialize.cc $C/index_cache.cc -lz
LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) ASAN_OPTIONS=detect_leaks=0 python $ROOT/tools/sanitize_host.py

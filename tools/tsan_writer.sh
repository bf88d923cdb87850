#!/bin/bash
# ThreadSanitizer pass over the sample writer (race detection): CPU only
set -e
cd "$(dirname "$0")/.."
g++ -std=c++17 -O1 -g -fsanitize=thread -I/usr/local/cuda/include -x c++ bayesrrcpp_b200/csrc/writer.cpp tools/tsan_writer.cpp -o /tmp/brr_tsan_writer -lpthread
/tmp/brr_tsan_writer

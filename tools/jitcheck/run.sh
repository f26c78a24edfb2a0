#!/bin/bash
# Host-vs-device check of the rows a generated plug-in computes (residual + AD Jacobian) on seeded inputs.
#   here      : tools/jitcheck/run.sh build <form.cu> <tag>   -> host<tag>.bin (g++ on stub headers) + dev<tag> (nvcc, sm_100a)
#   GPU box   : tools/jitcheck/dev<tag> gpurun_out/dev<tag>.bin
#   here      : python tools/jitcheck/compare.py <neq>:<tag> ...
set -e
if [ "$1" = build ]; then
  F=$(realpath "$2"); T=$3
  cd "$(dirname "$0")"
  g++ -std=c++17 -O1 -I stub -I ../../fedm_b200/csrc -DFORM="\"$F\"" -o host$T host_main.cpp && ./host$T host$T.bin
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I ../../fedm_b200/csrc -I ../../include \
      -DFORM="\"$F\"" -o dev$T dev_main.cu
fi

#!/bin/bash
# Runs every variant of tma_gather_bench in its own process (a faulting variant must not poison the rest).
cd "$(dirname "$0")"
out=../../gpurun_out/tma_gather.log; mkdir -p ../../gpurun_out; : > $out
run() { echo "+ $*" >> $out; timeout 90 ./tma_gather_bench "$@" >> $out 2>&1; echo "  rc=$?" >> $out; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $out
#   variant m box_rows warps stages dst_stride
for m in 2000000 10000000; do
  run g4 $m 1 8 3 128; run g4 $m 1 12 2 128; run g4 $m 1 24 1 128; run g4 $m 4 8 3 128
  run bulk $m 1 16 3 64; run bulk $m 1 24 2 64; run bulk $m 1 32 1 64
  run mix $m 1 16 2 128; run mix $m 1 16 3 128; run mix $m 1 24 2 128; run mix $m 1 32 1 128; run mix $m 1 28 2 128; run mix $m 4 24 2 128
  run mixb $m 1 32 2 64; run mixb $m 1 32 3 64; run mixb $m 1 16 4 64; run mixb $m 1 32 1 64
done
cat $out

#!/bin/bash
# developer tool: time the kernels of every experimental library in variants/ (tools/kbench.py)
# usage: tools/kvar.sh [points] [dtypes]   -> gpurun_out/kvar.log
pts=${1:-4194304}; export KBENCH_DTYPES=${2:-f64}
out=gpurun_out/kvar.log; : > $out
echo "default" >> $out; python tools/kbench.py $pts >> $out 2>&1
shopt -s nullglob
for v in variants/*.so; do
  echo "$v" >> $out
  QCPINN_B200_LIB=$PWD/$v python tools/kbench.py $pts >> $out 2>&1
done
cat $out

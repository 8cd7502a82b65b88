#!/bin/bash
echo "=== product"; python tools/layer_times_probe.py 2>&1 | tail -4 | head -1
for so in buckgnn_b200/lib/variants/*.so; do
  echo "=== $(basename $so .so)"
  BG_LIB_PATH=$PWD/$so python tools/layer_times_probe.py 2>&1 | tail -4 | head -1
done

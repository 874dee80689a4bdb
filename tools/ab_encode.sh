#!/bin/bash
# A/B of encoder builds / knobs on one GPU: tools/ab_encode.sh <n_blocks> <size> <fb> <dict> -- "ENV=.. ENV=.." ...
n=$1; size=$2; fb=$3; dict=$4; shift 5
for cfg in "$@"; do
  echo "=== $cfg"
  env LZB_ENC_TIMING=1 $cfg python tools/quick_encode_bench.py $n $size 4 1 $fb $dict 2>&1 | grep -E "lzb_enc wave|OK|rror"
done

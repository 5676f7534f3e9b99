#!/bin/bash
# A/B of built library variants on ONE box: tools/ab.sh "<prof_step args>" var_a.so var_b.so ...   (interleaved, 3 rounds)
args="$1"; shift
cp sdr-j-fm_b200/libsdrjfm_b200.so /tmp/orig.so
for r in 1 2 3; do
  for v in "$@"; do
    cp sdr-j-fm_b200/variants/$v sdr-j-fm_b200/libsdrjfm_b200.so
    echo -n "$v: "; python tools/prof_step.py $args 2>&1 | tail -1
  done
done
cp /tmp/orig.so sdr-j-fm_b200/libsdrjfm_b200.so

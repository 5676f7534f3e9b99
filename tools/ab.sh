#!/bin/bash
# A/B of built library variants on ONE box: tools/ab.sh "<args>" var_a.so var_b.so ...   (interleaved, 3 rounds)
# runs  $AB_CMD <args>  (default: python tools/prof_step.py) with each of sdr-j-fm_b200/variants/<name> in place of the library
args="$1"; shift
cmd="${AB_CMD:-python tools/prof_step.py}"
cp sdr-j-fm_b200/libsdrjfm_b200.so /tmp/orig.so
for r in 1 2 3; do
  for v in "$@"; do
    cp sdr-j-fm_b200/variants/$v sdr-j-fm_b200/libsdrjfm_b200.so
    echo -n "$v: "; $cmd $args 2>&1 | tail -1
  done
done
cp /tmp/orig.so sdr-j-fm_b200/libsdrjfm_b200.so

#!/bin/bash
# usage: tools/ab_libs.sh <rounds> <lib> [<lib> ...]  -- tools/ab_attend.py with each library in turn, <rounds> times, on one box
# (A/B of kernel variants built with different -D switches into build_ab/*.so; separate processes, interleaved)
r=$1; shift
for i in $(seq 1 "$r"); do
  for lib in "$@"; do
    echo -n "$lib " ; COATTN_B200_LIB=$lib python tools/ab_attend.py
  done
done

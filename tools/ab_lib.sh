#!/bin/bash
# A/B of two builds of libyvb200.so on the same box: usage  bash tools/ab_lib.sh <old.so> <new.so> [rounds]
# alternates the two libraries (bench.py --quick: step time only) and prints the ms per step of every run.
OLD=$1; NEW=$2; R=${3:-2}
DST=youtube-vln_b200/yvb200/libyvb200.so
for i in $(seq $R); do
  for which in old new; do
    if [ $which = old ]; then cp $OLD $DST; else cp $NEW $DST; fi
    echo -n "$which: "
    timeout 300 python bench.py --quick --steps 30 --warmup 5 2>/dev/null | tail -1
  done
done
cp $NEW $DST

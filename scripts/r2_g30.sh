#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r2_g30_exp_table_bytes.txt
: > $out
for rep in 1 2; do
for tag in full half quarter; do
  lib=gvamp_b200/lib/libgvamp_b200.so
  [ $tag != full ] && lib=exp/libgvamp_b200_$tag.so
  echo "== table copy: $tag (rep $rep)" >> $out
  python scripts/exp_run_with_lib.py $lib profiles/sweep_tuning.py --pairs 40 --configs "default=" 2>&1 | grep -v "^shard" >> $out
done
done
cat $out

#!/bin/bash
# Runs the three reference-free circuit proofs at 1 and N ranks and prints the JSON lines (transcript digests must agree).
# usage: tools/prove_ranks_check.sh "<N list>" [out.jsonl]
OUT=${2:-/dev/stdout}
cd "$(dirname "$0")/.."
for cfg in "18 1024 256 256 16" "19 aes 8" "19 sql 17"; do
  for n in $1; do
    if [ "$n" = 1 ]; then WORLD_SIZE=1 hobbit_b200/mlp_prove $cfg --reps 3 >> $OUT
    else tools/run_ranks.sh $n hobbit_b200/mlp_prove $cfg --reps 3 >> $OUT; fi
  done
done

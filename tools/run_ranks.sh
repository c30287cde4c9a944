#!/bin/bash
# One process per GPU without torchrun: tools/run_ranks.sh <world> <command...>   (HB_SHARE_GPU=1: all ranks on GPU 0)
# Sets RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT the way torchrun does; rank 0's stdout is the result.
W=$1; shift
export WORLD_SIZE=$W MASTER_ADDR=127.0.0.1 MASTER_PORT=${MASTER_PORT:-29640}
pids=()
for ((r = 1; r < W; r++)); do RANK=$r LOCAL_RANK=$r "$@" > /dev/null & pids+=($!); done
RANK=0 LOCAL_RANK=0 "$@"; rc=$?
for p in "${pids[@]}"; do wait $p || rc=$?; done
exit $rc

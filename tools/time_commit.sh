#!/bin/bash
# Development tool: resident-commit time of the bench workload for several library builds, one short bench.py run each.  The variant is
# put first on LD_LIBRARY_PATH under the product's file name so that the host mirror (libhobbit_host.so, RUNPATH $ORIGIN) binds to it too.
# usage: tools/time_commit.sh lib1.so lib2.so ...
for lib in "$@"; do
  d=build/variants/as_$(basename $lib .so); mkdir -p $d; cp $lib $d/libhobbit_b200.so
  LD_LIBRARY_PATH=$(realpath $d) HOBBIT_B200_LIB=$(realpath $d)/libhobbit_b200.so python bench.py --no-extras --steps 10 --cpu-chunks 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'lib':'$(basename $lib)','ms_per_step':d['ms_per_step'],'e2e_ms':1e3*d['n_gpus']*(1<<26)/d['e2e']['value'],'dominant':d['roofline']['kernel'],'avg_launch_ms':d['roofline']['avg_launch_ms'],'share':d['roofline']['kernel_time_share']}))"
  rm -rf $d
done

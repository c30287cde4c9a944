#!/bin/bash
# Link-level drop-in (test infrastructure): the reference's OWN objects — main.cpp with its real main, Our_PC.cpp with test_PC,
# Elastic_PC.cpp with test_Elastic_PC, everything else unchanged — with the entry points of the hot path WEAKENED by objcopy, linked
# against hobbit_b200/host/hobbit_adapter.cpp (global-namespace forwards with the reference's signatures) + libhobbit_host.so +
# libhobbit_b200.so.  Outputs only into oracle/_ref/.  No reference source is copied or modified.
#   oracle/_ref/pigeon_gpu       the reference's main():            pigeon_gpu 9 18 18 1 4 1024 256 256 16 | 5 19 8 1 | 6 19 17 1
#   oracle/_ref/ref_pc_gpu       the reference's test_PC / test_Elastic_PC (main.cpp has them commented out): ref_pc_gpu pc <logN> <opt> <K> |
#                                ref_pc_gpu elastic <logN> <logB> <opt>
set -e
cd "$(dirname "$0")"
REF=${REF:-/root/reference}; OUT=_ref; OBJ=$OUT/obj; W=$OUT/obj_weak
ARCH="-march=x86-64-v3 -mbmi2 -msha -mavx"
CXXFLAGS="-w -O3 -DNDEBUG $ARCH -std=gnu++14 -fPIC -I$OUT/stub -I$REF/src -I$REF/lib -I$REF/Blake"
mkdir -p $W
pat='^(commit_standard|open_standard|commit|open|prove_multiplication_tree_stream_shallow|prove_gate_consistency|prove_gate_consistency_lookups)\('
for o in Our_PC Elastic_PC sumcheck; do
    args=""
    while read -r addr kind sym; do
        dem=$(echo "$sym" | c++filt)
        if [ "$kind" = "T" ] && echo "$dem" | grep -Eq "$pat"; then args="$args --weaken-symbol=$sym"; fi
    done < <(nm $OBJ/$o.o)
    objcopy $args $OBJ/$o.o $W/$o.o
done
g++ $CXXFLAGS -c $REF/src/main.cpp -o $W/main_exe.o
g++ $CXXFLAGS -Dmain=pigeon_main -c $REF/src/main.cpp -o $W/main_lib.o
g++ $CXXFLAGS -std=gnu++17 -I../hobbit_b200/host -c ../hobbit_b200/host/hobbit_adapter.cpp -o $W/adapter.o
g++ $CXXFLAGS -c ../tests/cpp/ref_pc_driver.cpp -o $W/ref_pc_driver.o
OTHERS=$(ls $OBJ/*.o | grep -v -E "/(Our_PC|Elastic_PC|sumcheck|main_lib|ref_shim)\.o$")
LINK=(-L../hobbit_b200 -lhobbit_host -lhobbit_b200 '-Wl,-rpath,$ORIGIN/../../hobbit_b200' $REF/lib/libXKCP.a -lm -lpthread)
g++ -o $OUT/pigeon_gpu $W/main_exe.o $W/adapter.o $W/Our_PC.o $W/Elastic_PC.o $W/sumcheck.o $OTHERS "${LINK[@]}"
g++ -o $OUT/ref_pc_gpu $W/ref_pc_driver.o $W/main_lib.o $W/adapter.o $W/Our_PC.o $W/Elastic_PC.o $W/sumcheck.o $OTHERS "${LINK[@]}"
echo built $OUT/pigeon_gpu $OUT/ref_pc_gpu

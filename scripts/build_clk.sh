#!/bin/bash
# Development build: the library with the phase-clock instrumentation of bamp_fast.cu (-DAMPSM_CLK) as
# csrc/libampsm_b200_clk.so; select it with AMPSM_LIB=<path> (scripts/phase_clocks.py).
set -e
cd "$(dirname "$0")/../amp-sparc-spatialmodulation_b200/csrc"
python -c "import sys; sys.path.insert(0, '../..'); import __graft_entry__ as g; g.build()"
mkdir -p _obj_clk
for f in bamp_fast vamp_fast vamp_quad; do
  nvcc -Xcompiler -fPIC -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -DAMPSM_CLK -c $f.cu -o _obj_clk/$f.o &
done
wait
objs=$(ls _obj/*.o | grep -v -e bamp_fast.o -e vamp_fast.o -e vamp_quad.o)
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o libampsm_b200_clk.so $objs _obj_clk/bamp_fast.o _obj_clk/vamp_fast.o _obj_clk/vamp_quad.o
echo built $(pwd)/libampsm_b200_clk.so

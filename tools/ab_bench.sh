#!/bin/bash
# A/B timing of two builds of the library on the SAME GPU box (box-to-box variation is a few per cent):
#   cp slenderobjdet_b200/libslender_b200.so slenderobjdet_b200/csrc/build/ab/libA.so   (build A)
#   ... change, rebuild ...   cp ... libB.so                                              (build B)
#   gpurun -- bash tools/ab_bench.sh
for r in 1 2; do for v in A B; do
  SDB_LIB_PATH=$PWD/slenderobjdet_b200/csrc/build/ab/lib$v.so timeout 300 python bench.py --no-extra --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', d['ms_per_step'], d['roofline']['per_kernel_ms_per_step'], d['parity']['rel_err'])"
done; done

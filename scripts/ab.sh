#!/bin/bash
# A/B timing of library variants on one GPU: scripts/ab.sh name1 name2 ...   (gpurun_ab/lib_<name>.so; "cur" = in-tree build)
for v in "$@"; do
  if [ "$v" = cur ]; then unset HUGS_LIB; else export HUGS_LIB=/root/repo/gpurun_ab/lib_$v.so; fi
  echo "== $v"; timeout 100 python scripts/quick_bench.py | tail -2 | sed 's/, .reductions.*//'
done

#!/bin/bash
# A/B of library builds: tools/ab.sh "<workload> ..." lib1.so lib2.so ...   (prints ms/step and the phase table per build)
WL="$1"; shift
for lib in "$@"; do
  for w in $WL; do
    echo -n "$lib $w: "
    MITHRA_GPU_LIB=$lib python bench.py --workload $w --steps 30 --warmup 10 --no-cpu-baseline --no-e2e 2>/dev/null | python tools/phases.py
  done
done

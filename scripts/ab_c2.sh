#!/bin/bash
# Same-box A/B of library variants on the C2 kernel (level set, 4096 x 4096 grid, S = 64) + a parity run of the
# bench-problem tests against the default build.  usage: scripts/ab_c2.sh <variant> [<variant> ...]
set -u
mkdir -p gpurun_out
for rep in ${REPS:-1 2}; do
  for v in "$@"; do
    lib=psbax_b200/libpsbax_b200.so
    [ "$v" != "default" ] && lib=psbax_b200/libpsbax_b200_$v.so
    echo -n "$v: "; PSBAX_LIB=$PWD/$lib python scripts/tc_probe.py 4096 auto 10 2>&1 | tail -1
  done
done

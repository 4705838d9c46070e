#!/bin/bash
# Same-box A/B in the SUSTAINED (power-capped) regime: 150 back-to-back launches per variant (~1.2 s).
set -u
for rep in 1 2; do
  for v in "$@"; do
    lib=psbax_b200/libpsbax_b200.so
    [ "$v" != "default" ] && lib=psbax_b200/libpsbax_b200_$v.so
    echo -n "$v: "; PSBAX_LIB=$PWD/$lib python scripts/tc_probe.py 4096 auto 150 2>&1 | tail -1
  done
done

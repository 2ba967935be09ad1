#!/usr/bin/env bash
# A/B timing of the stepping kernels for each built variant (lib/variant_*.so
# and the default build) in one GPU session: configs[1] and 1 Mi envs x 8 / x 1,
# per-step outputs on.
for lib in putting-dune_b200/lib/libpdune_b200.so putting-dune_b200/lib/variant_*.so; do
  [ -f "$lib" ] || continue
  echo "== $lib"
  for args in "4096 256 prior 10" "1048576 8 prior 8" "1048576 1 prior 8"; do
    OUTPUTS=1 PDUNE_B200_LIB=$PWD/$lib python profiles/prof_walk.py $args 2>&1 | tail -1
  done
done

#!/usr/bin/env bash
# Times the stepping kernel for each built occupancy variant (lib/variant_*.so).
for lib in putting-dune_b200/lib/variant_*.so; do
  echo "== $lib"
  PDUNE_B200_LIB=$PWD/$lib python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('config2 value %.3e  ms %.3f | at_scale %.3e ms %.4f' % (d['value'], d['ms_per_step'], d['at_scale']['value'], d['at_scale']['launch_ms']))"
  PDUNE_B200_LIB=$PWD/$lib python bench.py --workload config5 --steps 5 --warmup 3 --no-cpu-baseline --no-at-scale 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('1Mi x 8 steps rollout value %.3e  ms %.3f' % (d['value'], d['ms_per_step']))"
done

#!/bin/bash
# A/B of two builds on the whole bench step (both models' loops in flight): tools/ab_bench.sh <libA> <libB> [bench args]
A=$1; B=$2; shift 2
for rep in 1 2 3; do
  for L in "$A" "$B"; do
    echo -n "$(basename $L) run $rep: "
    DTRAJ_LIB=$L python bench.py --quick --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.0f e2e %.0f ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
  done
done

#!/bin/bash
# cfg2 step time under one-off environment switches (each a fresh process): usage  scripts/env_sweep.sh "A=1" "B=2 C=3" ...
for e in "" "$@"; do
  r=$(env $e python bench.py --steps 200 --warmup 10 --no-extra --no-cpu-baseline --no-micro 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), d['launches_per_step'])")
  echo "[$e] $r"
done

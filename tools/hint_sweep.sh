#!/bin/bash
# A/B of the L2 eviction-priority hint masks (LDM_L2_HINT) with short bench runs: prints mask, layouts/s, ms per sample(), SM MHz
for f in "$@"; do
  LDM_L2_HINT=$f python bench.py --steps 4 --no-cpu-baseline --no-configs | python -c "import json,sys; d=json.loads(sys.stdin.read()); print($f, round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
done

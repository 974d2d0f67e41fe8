#!/bin/bash
# Repeats bench.py on one box: prints value / e2e / latency per run, and the error tail of a run that fails.
#   scripts/ab_bench.sh [steps] [repeats]
steps=${1:-100}
reps=${2:-3}
for i in $(seq 1 $reps); do
  for wl in 1080p 4k; do
    if ! python bench.py --workload $wl --steps $steps --warmup 10 --no-cpu-baseline > /tmp/ab.json 2> /tmp/ab.err; then
      echo "$wl run $i FAILED"; grep -v "^frame #" /tmp/ab.err | tail -8 | cut -c1-300
    fi
    python - "$i" "$wl" <<'PY'
import json, sys
try:
    d = json.loads(open("/tmp/ab.json").read())
    print(sys.argv[1], sys.argv[2], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "three-call", round(d["e2e"]["three_call_api"]["value"], 1),
          "p50", round(d["latency_ms"]["p50"], 3), "sm_mhz", d["clocks"]["sm_mhz"], flush=True)
except Exception as e:
    print(sys.argv[1], sys.argv[2], "no JSON line:", e)
PY
  done
done

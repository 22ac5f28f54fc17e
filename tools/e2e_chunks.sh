#!/bin/bash
# e2e sensitivity to the H2D chunk size (rows per chunk)
for C in 32768 131072 524288; do
  EMR2A_E2E_CHUNK=$C python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > /tmp/e2e_$C.json
  python - "$C" <<'PY'
import json, sys
c = sys.argv[1]
d = json.load(open(f"/tmp/e2e_{c}.json"))
print("chunk", c, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 2), d["clocks"])
PY
done

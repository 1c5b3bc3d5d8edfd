#!/bin/bash
# per-config bench lines (BASELINE.json configs) on 1 GPU: ours + reference arm; results under gpurun_out/configs/
mkdir -p gpurun_out/configs
for c in metric inverted_pendulum inv_double_pendulum smgpr swimmer test_cascade; do
  python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/configs/$c.json 2> gpurun_out/configs/$c.err || echo "FAILED $c"
  python bench.py --config $c --impl reference --steps 5 --warmup 1 > gpurun_out/configs/${c}_reference.json 2>> gpurun_out/configs/$c.err || echo "FAILED ref $c"
  python - <<PY
import json
try:
    l = json.load(open("gpurun_out/configs/$c.json")); r = json.load(open("gpurun_out/configs/${c}_reference.json"))
    fb = l.get("fwd_bwd") or {}
    print("$c: fwd %.0f e2e %.0f fwd_bwd %.0f | cpu %.2f | roofline %.3f whole %.3f | api %s" % (
        l["value"], l["e2e"]["value"], fb.get("value", 0), r["value"], l["roofline"]["frac"], l["roofline"]["whole_step"]["frac"],
        (l.get("api_predict") or {}).get("us_per_rollout_step")))
except Exception as e:
    print("$c: parse error", e)
PY
done

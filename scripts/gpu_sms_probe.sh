#!/bin/bash
# (K2Y_TC_SMS, K2Y_DWPW_SMS, lanes) sweep (bench --no-cpu): value / e2e per setting.  CASES="tc:dw:lanes ..."
for c in ${CASES:-74:148:2}; do
  IFS=: read sms dsms l <<< "$c"
  K2Y_TC_SMS=$sms K2Y_DWPW_SMS=$dsms timeout 300 python bench.py --lanes $l --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_sweep_$sms_$dsms_$l.json 2> gpurun_out/bench_sweep.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_sweep_$sms_$dsms_$l.json"))
    print("tc_sms $sms dwpw_sms $dsms lanes $l value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "flushed", round(d["ms_per_step_flushed"], 4), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("tc_sms $sms dwpw_sms $dsms lanes $l failed", e)
PY
done

#!/bin/bash
# the bench line of the driver's protocol twice (second time without the long extras)
python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_n1_new.json
python bench.py --steps 20 --warmup 5 --no-large --no-cpu 2>&1 | tail -1 > gpurun_out/bench_again.json
python - <<'PY'
import json
for f in ("gpurun_out/bench_n1_new.json", "gpurun_out/bench_again.json"):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    lg = d.get("large") or {}
    print(f, "value", d["value"], d.get("step_us"), "e2e", d["e2e"]["value"], d["e2e"]["pageable"],
          "kernel_ms", d["roofline"]["kernel_ms"], "scg", (d.get("scg") or {}).get("iters_per_s"),
          "c_abi", d["c_abi"], "large", lg.get("value"), lg.get("roofline"), lg.get("scg"),
          "batched", (d.get("batched") or {}).get("value"), (d.get("batched") or {}).get("roofline"), d.get("clocks"))
PY

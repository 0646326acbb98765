#!/bin/bash
# One gpurun call: GPU tests, then the bench line of the driver's protocol (twice for the headline)
# into gpurun_out/.  usage: gpurun -- 'bash tools/gpu_check.sh'
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for e in "$@"; do env $e python tools/walk_ab.py "$e" 2>&1 | grep " L="; done
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
          "batched", (d.get("batched") or {}).get("value"))
PY
if [ -n "$NCU_BATCH" ]; then
  python tools/batch_target.py 1024 | tail -1
  ncu --set full --clock-control none --import-source on -k regex:esde_l96_lean_batch -s 2 -c 1 -o gpurun_out/r2c_lean_batch -f python tools/batch_target.py 1024 > gpurun_out/ncu_lean.log 2>&1
  tail -1 gpurun_out/ncu_lean.log
fi

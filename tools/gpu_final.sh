#!/bin/bash
# last GPU call of the round: GPU tests, smoke(), and the ncu launch list of a short bench run
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --no-large --no-cpu > gpurun_out/b_short.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_bench_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-large --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/b_short.log | cut -c1-300
( time python bench.py > gpurun_out/b_default.log 2>&1 ) 2>&1 | grep real
tail -1 gpurun_out/b_default.log | cut -c1-400
wc -l gpurun_out/r2_bench_launches.csv

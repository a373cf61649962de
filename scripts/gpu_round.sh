#!/bin/bash
# scratch: the command of the most recent gpurun call
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.json
( timeout 1500 python -m pytest tests/test_precise_gpu.py tests/test_model_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -k "split3 or north_star or full_width or fp32_class" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
grep -E "^E  |FAILED|passed|failed|pytest exit" gpurun_out/pytest_gpu.log | cut -c1-250 | head -40
python - <<'PY'
import json
d=json.load(open('gpurun_out/parity_report.json'))
for k,v in d.items(): print(k, v)
PY
for wl in mlm mc bert; do
  timeout 900 python bench.py --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl exit $?"; cut -c1-250 gpurun_out/bench_$wl.json; tail -2 gpurun_out/bench_$wl.err
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cut -c1-250 gpurun_out/bench_ref.json
nproc; free -g | head -2

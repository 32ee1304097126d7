#!/bin/bash
# Round-end evidence on one GPU: the GPU test suite, the bench lines (sustained default, 20-step burst, reference arm),
# the ncu launch lists of an inference step and a training step (each after a plain run of the same command), the
# per-layer training kernel times and the determinism stress.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3) > gpurun_out/r2f_pytest.log
tail -n 2 gpurun_out/r2f_pytest.log
python bench.py > gpurun_out/bench_r02_1gpu.json 2> gpurun_out/r2f_b1.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_1gpu_burst.json 2> gpurun_out/r2f_b2.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r02_reference_arm.json 2> gpurun_out/r2f_b3.err
python tools/run_infer_step.py 4 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_r02.csv python tools/run_infer_step.py 4 > gpurun_out/r2f_ncu1.log 2>&1
python tools/run_train_steps.py 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/launches_train_r02.csv python tools/run_train_steps.py 3 > gpurun_out/r2f_ncu2.log 2>&1
python tools/bench_train_kernels.py --out gpurun_out/train_kernels_r02.json > /dev/null 2>&1
python tools/stress_determinism.py 200 4 > gpurun_out/determinism_r02.json 2> /dev/null
for f in gpurun_out/bench_r02_1gpu.json gpurun_out/bench_r02_1gpu_burst.json gpurun_out/bench_r02_reference_arm.json; do
python - "$f" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d["value"], 1), round(d["ms_per_step"], 4), d.get("e2e", {}).get("value"),
      d.get("roofline", {}).get("frac"), d.get("train", {}).get("ms_per_step"), d.get("clocks"))
PY
done

#!/bin/bash
# Full GPU pass for one round: parity tests, smoke, bench (both arms), ncu launch list of the bench command, ncu --set full
# of the dominant kernels.      gpurun --timeout 2400 -- 'bash tools/gpu_round.sh r02m'
TAG=${1:-round}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
nproc > $OUT/nproc.txt
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $OUT/gpu_tests.log
tail -4 $OUT/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
tail -5 $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench.json"))
    print("value %.0f frames/s  ms/step %.2f  e2e %.0f (%.2f of h2d ceiling %.0f)  launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["fraction_of_h2d_ceiling"], d["e2e"]["h2d_ceiling"], d["gpu_launches"]))
    print("roofline", d["roofline"])
    print("cpu", d["cpu_baseline"])
    print("clocks", d["clocks"], "results", d["results"])
    print("width", d["width_sweep"])
    for k in d["kernels"][:14]: print("  %-30s %8.3f ms %5.1f%%" % (k["name"], k["ms_per_step"], 100*k["share"]))
    r=json.load(open("$OUT/bench_ref.json")); print("reference arm", r["value"], r["cpu_baseline"])
except Exception as e: print("bench parse failed", e)
PY
# the launch list of the bench command itself (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-width-sweep > $OUT/bench_under_ncu.json 2> $OUT/bench_under_ncu.err
# full captures of the dominant kernels (second step = warm)
for K in pyramid_u8_fused_kernel lk_track_smem_kernel signal_fit_kernel upsample_pass_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 2 -f -o $OUT/$K \
      python tools/prof_step.py 64 2 > $OUT/ncu_$K.log 2>&1
done
ls -la $OUT

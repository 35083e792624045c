#!/bin/bash
# One gpurun call that times every experiment prepared at the end of round 1:
#   here (CPU, before the call):
#     python -m respmon_b200.build --variant w24  "-DPU_MAX_WARPS=24"
#     python -m respmon_b200.build --variant alu1 "-DPU_ALU_TAPS=1"
#     python -m respmon_b200.build --variant alu2 "-DPU_ALU_TAPS=2"
#     python -m respmon_b200.build --variant div3 "-DLM_DIV3"
#     python -m respmon_b200.build --variant lmilp "-DLM_DIV3 -DLM_ROWS2"
#     python -m respmon_b200.build --variant hm4  "-DHM_MIN_BLOCKS=4"
#     python -m respmon_b200.build --variant b21  "-DPU_BOUND_WARPS=21"
#     python -m respmon_b200.build --variant i32  "-DPU_IDX32"
#     python -m respmon_b200.build --variant ptv  "-DPT_VEC_STAGE"
#     python -m respmon_b200.build --variant hm1b "-DHM_ONE_BARRIER"
#     python -m respmon_b200.build --variant hm1b4 "-DHM_ONE_BARRIER -DHM_MIN_BLOCKS=4"
#     python -m respmon_b200.build --variant hmall "-DHM_ONE_BARRIER -DHM_MIN_BLOCKS=4 -DHM_PAR_LIST"
#   gpurun --timeout 900 -- 'bash tools/round2_first_call.sh r02a w24 alu1 alu2 div3 lmilp hm4 b21 i32 ptv hm1b hm1b4 hmall'
TAG=${1:-r02a}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -q > $OUT/gpu_tests.log 2>&1; tail -15 $OUT/gpu_tests.log
# batch-width sweep (BASELINE config 3 asks for 512 clips per step)
for NC in 128 256 512; do timeout 200 python tools/bench_stage.py $NC 4 0 > $OUT/width_$NC.log 2>&1; head -12 $OUT/width_$NC.log; done
timeout 200 python tools/dev_temporal_sparse.py > $OUT/temporal_sparse.log 2>&1; cat $OUT/temporal_sparse.log
timeout 200 python tools/dev_fit_solo.py 64 10 > $OUT/fit_solo.log 2>&1; cat $OUT/fit_solo.log
timeout 200 python tools/dev_cal_split.py 64 10 > $OUT/cal_split.log 2>&1; cat $OUT/cal_split.log
bash tools/variants.sh $TAG "$@"

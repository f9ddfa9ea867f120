#!/bin/bash
# Round profile capture on the GPU box (run through gpurun from the repo root):
#   gpurun --timeout 1500 -- 'bash profiles/capture.sh r2c'
# 1. launch list with per-launch duration and DRAM bytes of every kernel of one eagerly launched bench run (light pass)
# 2. `--set full` captures: one steady-state step of this package's streaming kernels, and the GEMM kernels alone
# 3. the bench lines themselves, NOT under the profiler (fp32 default incl. cpu baseline and extra configs, bf16, iid, cfg 5)
# Summaries are made from gpurun_out/ afterwards with profiles/summarize.py and committed under profiles/.
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 3 --warmup 2 --no-graph --no-callers --no-cpu-baseline --no-extra"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 500 --csv \
    --log-file $OUT/launches_$TAG.csv $B > $OUT/ncu1_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'(readloss|read_fwd|read_bwd|write_reduce|write_bwd|update_fwd|update_bwd|colsoftmax|bn_|labels_pack)' \
    --launch-skip 60 --launch-count 24 -f -o $OUT/prof_$TAG $B > $OUT/ncu2_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'conv1x1_(nn|wgrad)_kernel' --launch-skip 9 --launch-count 3 \
    -f -o $OUT/prof_gemm_$TAG python profiles/gemm_only.py > $OUT/ncu3_$TAG.log 2>&1
python bench.py --steps 20 --warmup 5 2> $OUT/bench_$TAG.err | tail -1 > $OUT/bench_line_$TAG.json
python bench.py --steps 20 --warmup 5 --dtype bf16 --no-cpu-baseline 2> $OUT/bench_bf16_$TAG.err | tail -1 > $OUT/bench_line_bf16_$TAG.json
python bench.py --steps 20 --warmup 5 --labels iid --no-cpu-baseline --no-callers --no-extra 2> $OUT/bench_iid_$TAG.err | tail -1 > $OUT/bench_line_iid_$TAG.json
python bench.py --workload cfg5_dr101v2_eval_b1 --steps 50 --warmup 5 2> $OUT/bench_cfg5_$TAG.err | tail -1 > $OUT/bench_line_cfg5_$TAG.json
python bench.py --impl reference --steps 5 --warmup 1 2> $OUT/bench_ref_$TAG.err | tail -1 > $OUT/bench_line_ref_$TAG.json
python profiles/show_bench.py < $OUT/bench_line_$TAG.json
python profiles/show_bench.py < $OUT/bench_line_bf16_$TAG.json

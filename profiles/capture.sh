#!/bin/bash
# Round profile capture on the GPU box (run through gpurun from the repo root):
#   gpurun --timeout 900 -- 'bash profiles/capture.sh r1d'
# 1. launch list with per-launch duration and DRAM bytes of every kernel of the bench command (light metrics pass)
# 2. one `--set full` capture of one steady-state step of this package's kernels (source-level, for stalls/op mix)
# 3. the bench line itself, NOT under the profiler
# Summaries are made from gpurun_out/ afterwards with profiles/summarize.py and committed under profiles/.
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 3 --warmup 2 --no-graph --no-callers --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
    --log-file $OUT/launches_$TAG.csv $B > $OUT/ncu1_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'(readloss|read_fwd|read_bwd|write_reduce|write_bwd|update_fwd|update_bwd|colsoftmax|bn_)' \
    --launch-skip 38 --launch-count 19 -f -o $OUT/prof_$TAG $B > $OUT/ncu2_$TAG.log 2>&1
python bench.py --steps 100 --warmup 10 2> $OUT/bench_$TAG.err | tail -1 > $OUT/bench_line_$TAG.json
python bench.py --steps 50 --warmup 10 --dtype bf16 --no-cpu-baseline 2> $OUT/bench_bf16_$TAG.err | tail -1 > $OUT/bench_line_bf16_$TAG.json
python bench.py --steps 50 --warmup 10 --labels iid --no-cpu-baseline --no-callers 2> $OUT/bench_iid_$TAG.err | tail -1 > $OUT/bench_line_iid_$TAG.json
python profiles/show_bench.py < $OUT/bench_line_$TAG.json
python profiles/show_bench.py < $OUT/bench_line_bf16_$TAG.json
python profiles/show_bench.py < $OUT/bench_line_iid_$TAG.json

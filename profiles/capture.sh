#!/bin/sh
# Round-end captures on one B200 (run under gpurun from the repo root): GPU tests, the bench line, the CPU arm, the launch
# list of the bench command, and one `ncu --set full` capture per kernel family.  Outputs under gpurun_out/.
# usage: sh profiles/capture.sh [all|nogermline]
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/final_gputest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/final_launches.csv \
  python bench.py --steps 4 --warmup 3 > gpurun_out/final_launches.log 2>&1
if [ "$1" != "nogermline" ]; then
ncu --set full --clock-control none --import-source on -k 'regex:k_call_tile|k_exact_loci|k_rec_gather|k_rec_to_host|k_general_to_host' \
  --launch-skip 15 -c 5 -o gpurun_out/final_germline -f python profiles/ab_step.py germline 63025520 1 > gpurun_out/final_germline_ncu.log 2>&1
fi
ncu --set full --clock-control none --import-source on -k 'regex:k_somatic|k_evidence' \
  --launch-skip 9 -c 3 -o gpurun_out/final_somatic -f python profiles/ab_step.py somatic 8000000 1 > gpurun_out/final_somatic_ncu.log 2>&1
ls -la gpurun_out/ | tail -12

#!/bin/bash
# One GPU call of the round: the -m gpu tests, the default bench line (C3), the per-launch device times of the same
# command (ncu --metrics gpu__time_duration.sum) and one --set full capture of a settled step kernel launch.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [tests|notests]
tag=${1:-r02}; what=${2:-tests}
mkdir -p gpurun_out
if [ "$what" = tests ]; then
  python -m pytest tests -m gpu -q -x > gpurun_out/gputest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_$tag.log; tail -3 gpurun_out/gputest_$tag.log
  python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; tail -1 gpurun_out/smoke_$tag.log
fi
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_$tag.json
python bench.py --steps 3 --warmup 3 > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_launch_$tag.log 2>&1
echo "launch list rc=$?"
# launches of rkfd_step_kernel in that command: UpdateInit, settle (2), warm-up (3), timed (3), ...: the 8th is a timed step
ncu --set full --clock-control none --import-source on -k regex:rkfd_step -s 7 -c 1 -f -o gpurun_out/prof_$tag python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_full_$tag.log 2>&1
echo "full capture rc=$?"

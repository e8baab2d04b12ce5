#!/bin/bash
# N-GPU call (gpurun --gpus N): the in-process multi-device test (rkFDBatchSetDevices), then both bench arms under torchrun.
N=${1:-2}; tag=${2:-r02}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_device or subset or resort" > gpurun_out/gputest_multi_$tag.log 2>&1; tail -3 gpurun_out/gputest_multi_$tag.log
nvidia-smi topo -m > gpurun_out/topo_n${N}_$tag.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_ref_n${N}_$tag.json 2> gpurun_out/bench_ref_n${N}_$tag.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_n${N}_$tag.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 50 --warmup 10 > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_$tag.err; echo "bench rc=$?"; cut -c1-2500 gpurun_out/bench_n${N}_$tag.json

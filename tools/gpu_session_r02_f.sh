#!/bin/bash
# Round 2, GPU session F (2 GPUs): multi-GPU parity (sharded RSVD incl. the robust stage, the timeout test), 2-GPU bench,
# kNN timing with the hit masks.
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | head -3
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/r02f_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -25 $O/r02f_pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r02f_bench_2gpu.json 2> $O/r02f_bench_2gpu.err; echo "bench2 rc=$?"; tail -c 1200 $O/r02f_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > $O/r02f_bench_ref_2.json 2>> $O/r02f_bench_2gpu.err; echo "ref2 rc=$?"
python tools/bench_knn.py 1048576 > $O/r02f_knn_1m.json 2> $O/r02f_knn.err; cat $O/r02f_knn_1m.json; tail -3 $O/r02f_knn.err

#!/bin/bash
# round 2, call i (8 GPUs): bench with the two-calls-in-flight mode at N=8 and N=2, C4 hybrid scatter variants
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 8 --steps 20 --warmup 5 --no-c3 > $O/r2i_bench_tr8.log 2> $O/r2i_bench_tr8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --steps 20 --warmup 5 --no-c3 > $O/r2i_bench_tr2.log 2> $O/r2i_bench_tr2.err
timeout 100 python tools/bench_c4.py 8 push hybrid 3 > $O/r2i_c4_hybrid3.log 2>&1
timeout 100 python tools/bench_c4.py 8 push hybrid 4 > $O/r2i_c4_hybrid4.log 2>&1
timeout 100 python tools/bench_c4.py 8 push hybrid 2 > $O/r2i_c4_hybrid2.log 2>&1
timeout 100 python tools/bench_c4.py 8 push ce > $O/r2i_c4_ce.log 2>&1
timeout 200 python -m pytest tests -m gpu -q -k "two_gpu" > $O/r2i_pytest_subset.log 2>&1; echo "pytest rc=$?" >> $O/r2i_pytest_subset.log
echo done

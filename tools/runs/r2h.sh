#!/bin/bash
# round 2, call h (8 GPUs): the scaling run as the driver does it (torchrun N=8 and N=4), C4 at 8
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2h_bench_tr8.log 2> $O/r2h_bench_tr8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 4 --steps 20 --warmup 5 --no-c3 > $O/r2h_bench_tr4.log 2> $O/r2h_bench_tr4.err
timeout 120 python tools/bench_c4.py 8 push ce > $O/r2h_c4_push_ce.log 2>&1
timeout 300 python -m pytest tests -m gpu -q -k "two_gpu or broadcast_window or stall or map_coalesces or lane" > $O/r2h_pytest_subset.log 2>&1; echo "pytest rc=$?" >> $O/r2h_pytest_subset.log
echo done

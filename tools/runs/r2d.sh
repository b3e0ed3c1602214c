#!/bin/bash
# round 2, call d (8 GPUs): GPU suite, torchrun bench N=8, host-path sweep (NUMA, subsets), C4 push/pull, push sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
nvidia-smi topo -m > $O/r2d_topo.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $O/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2d_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2d_bench_tr8.log 2> $O/r2d_bench_tr8.err
timeout 400 python tools/sweep_host.py 8 > $O/r2d_host.log 2>&1
timeout 120 python tools/bench_c4.py 8 push ce > $O/r2d_c4_push_ce.log 2>&1
timeout 120 python tools/bench_c4.py 8 push sm > $O/r2d_c4_push_sm.log 2>&1
timeout 120 python tools/bench_c4.py 8 pull > $O/r2d_c4_pull.log 2>&1
timeout 200 python tools/sweep_push.py 8 > $O/r2d_push.log 2>&1
timeout 100 python tools/prof_small_calls.py --gpu --gpus 8 -n 30000 > $O/r2d_small8.log 2>&1
timeout 100 python tools/sweep_duplex.py --mib 224 > $O/r2d_duplex.log 2>&1
echo done

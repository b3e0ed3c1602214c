#!/bin/bash
# round 2, call b (2 GPUs): GPU suite, new single-grid push pipeline sweep, new bench.py at N=1 and N=2 (torchrun)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2b_pytest.log
timeout 300 python tools/sweep_push.py 2 > $O/r2b_push.log 2>&1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2b_bench1.log 2> $O/r2b_bench1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2b_bench_tr2.log 2> $O/r2b_bench_tr2.err
timeout 600 python bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/r2b_ref2.log 2> $O/r2b_ref2.err
echo done

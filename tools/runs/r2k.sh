#!/bin/bash
# round 2, call k (2 GPUs): final tree — full GPU suite (incl. the 2-GPU peer test), smoke, short bench N=1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2k_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2k_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2k_smoke.log
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-c3 > $O/r2k_bench1.log 2> $O/r2k_bench1.err
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests -m gpu -q -x > $O/r2k_pytest_1gpu.log 2>&1; echo "pytest rc=$?" >> $O/r2k_pytest_1gpu.log
echo done

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 200 python tools/probe_l1.py > $O/r2y_probe_l1.log 2>&1
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "mlp" > $O/r2y_pytest_mlp.log 2>&1; echo "rc=$?" >> $O/r2y_pytest_mlp.log
timeout 200 python tools/bench_mlp.py > $O/r2y_bench_mlp.log 2>&1
timeout 200 python tools/probe_trace.py 75776 2 0 0 1 > $O/r2y_trace.log 2>&1
echo done

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 200 python tools/probe_l1.py > $O/r2z_probe_l1.log 2>&1
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "mlp" > $O/r2z_pytest_mlp.log 2>&1; echo "rc=$?" >> $O/r2z_pytest_mlp.log
timeout 200 python tools/bench_mlp.py > $O/r2z_bench_mlp.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm|mlp" -c 12 --csv --log-file $O/r2z_mlp_launches.csv python tools/probe_one.py 75776 4 > $O/r2z_launch.log 2>&1
echo done

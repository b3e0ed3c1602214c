#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -q -x -k "mlp" > $O/r2aa_pytest_mlp.log 2>&1; echo "rc=$?" >> $O/r2aa_pytest_mlp.log
timeout 100 python tools/probe_l1.py > $O/r2aa_probe_l1.log 2>&1
echo done

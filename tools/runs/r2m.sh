#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > $O/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2m_pytest.log
timeout 200 python tools/bench_mlp.py > $O/r2m_bench_mlp.log 2>&1
echo done

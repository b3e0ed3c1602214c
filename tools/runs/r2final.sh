#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 420 python -m pytest tests -m gpu -q -x > $O/r2final_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2final_pytest.log
timeout 300 python bench.py > $O/r2final_bench.log 2>&1; echo "bench rc=$?" >> $O/r2final_bench.log
echo done

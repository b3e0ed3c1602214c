#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 70 python -m pytest tests -m gpu -q -x > gpurun_out/r2final3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2final3_pytest.log
echo done

#!/bin/bash
# round 2, call f (2 GPUs): GPU suite, push pipeline with slices-per-fence sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2f_pytest.log
timeout 400 python tools/sweep_push.py 2 > $O/r2f_push.log 2>&1
echo done

#!/bin/bash
# final tree on 2 GPUs: full GPU suite (incl. the real-peer tests), smoke, C4 at 2 GPUs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x > $O/r2final2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2final2_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2final2_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2final2_smoke.log
timeout 120 python tools/bench_c4.py 2 > $O/r2final2_c4.log 2>&1
echo done

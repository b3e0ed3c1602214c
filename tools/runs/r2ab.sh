#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 120 python tools/probe_tuning.py 25 1,65,1,65 1 > $O/r2ab_probe_pf_fused.log 2>&1; echo "rc=$?" >> $O/r2ab_probe_pf_fused.log
echo done

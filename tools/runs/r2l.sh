#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 120 python tools/probe_l1.py > $O/r2l_probe_l1.log 2>&1; echo "rc=$?" >> $O/r2l_probe_l1.log
echo done

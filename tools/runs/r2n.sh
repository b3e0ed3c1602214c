#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 300 python tools/probe_arrive.py > $O/r2n_probe_arrive.log 2>&1; echo "rc=$?" >> $O/r2n_probe_arrive.log
echo done

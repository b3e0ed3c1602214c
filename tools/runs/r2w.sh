#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 300 python tools/probe_tuning.py 25 1,33,1,33 1 > $O/r2w_probe_relaxed.log 2>&1; echo "rc=$?" >> $O/r2w_probe_relaxed.log
timeout 200 python tools/probe_trace.py 75776 2 0 0 33 > $O/r2w_trace_relaxed.log 2>&1
echo done

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 300 python tools/probe_tuning.py 26 0,2,3,4,6,16,19,20 0 > $O/r2q_probe_pf.log 2>&1; echo "rc=$?" >> $O/r2q_probe_pf.log
timeout 200 python tools/probe_trace.py 75776 2 3 > $O/r2q_trace_pf3.log 2>&1
echo done

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 300 python tools/probe_tuning.py 25 17,1,17,1 1 > $O/r2u_probe_endwait.log 2>&1; echo "rc=$?" >> $O/r2u_probe_endwait.log
timeout 200 python tools/probe_trace.py 75776 2 0 0 > $O/r2u_trace_v2.log 2>&1
echo done

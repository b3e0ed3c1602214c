#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 200 python tools/probe_l1.py > $O/r2s_probe_l1.log 2>&1
timeout 200 python tools/probe_trace.py 75776 4 0 0 > $O/r2s_trace_v4.log 2>&1
echo done

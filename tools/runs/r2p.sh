#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 200 python tools/probe_trace.py 75776 2 > $O/r2p_trace_v2.log 2>&1
timeout 200 python tools/probe_trace.py 75776 3 > $O/r2p_trace_v3.log 2>&1
timeout 200 python tools/probe_l1.py > $O/r2p_probe_l1.log 2>&1
echo done

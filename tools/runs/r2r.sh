#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 200 python tools/probe_trace.py 75776 2 0 8 > $O/r2r_trace_nostore.log 2>&1
timeout 200 python tools/probe_trace.py 75776 1 0 0 > $O/r2r_trace_v1.log 2>&1
timeout 200 python tools/probe_trace.py 75776 2 0 0 > $O/r2r_trace_v2.log 2>&1
echo done

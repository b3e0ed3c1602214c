#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 300 python tools/probe_tuning.py 8 75776,56832,37888,18944,75776 75776 > $O/r2v_probe_chunk.log 2>&1; echo "rc=$?" >> $O/r2v_probe_chunk.log
echo done

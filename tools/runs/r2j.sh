#!/bin/bash
# round 2, call j (8 GPUs, short): C4 with copy-engine PULLS on the ranks (staged pull, no root-side scatter)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 60 python tools/bench_c4.py 8 pull > $O/r2j_c4_pull_ce.log 2>&1
echo done

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm|mlp" -c 20 --csv --log-file $O/r2o_mlp_launches.csv python tools/probe_one.py 75776 4 > $O/r2o_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm|mlp" -s 4 -c 2 -o $O/r2o_mlp_full -f python tools/probe_one.py 75776 4 > $O/r2o_full.log 2>&1
echo done

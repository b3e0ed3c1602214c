#!/bin/bash
# round 2, call g (1 GPU): the driver's view — GPU suite on one GPU, smoke, bench N=1, reference arm at N=1 and N=8 ranks
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2g_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2g_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2g_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2g_bench1.log 2> $O/r2g_bench1.err
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r2g_ref1.log 2> $O/r2g_ref1.err
timeout 1500 python bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > $O/r2g_ref8.log 2> $O/r2g_ref8.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2g_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-c3 > $O/r2g_ncu_bench.log 2>&1
echo done

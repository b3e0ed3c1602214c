#!/bin/bash
# round 2, call a (2 GPUs): full GPU suite, NVLink duplex matrix, host-path sweep, small-call lane, NVLink counters
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $O/r2a_gpu.txt 2>&1
nvidia-smi topo -m > $O/r2a_topo.txt 2>&1
nproc >> $O/r2a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2a_pytest.log
timeout 300 python tools/sweep_duplex.py > $O/r2a_duplex.log 2>&1
timeout 300 python tools/sweep_host.py 2 --quick > $O/r2a_host.log 2>&1
timeout 200 python tools/prof_small_calls.py --gpu --gpus 1 -n 30000 > $O/r2a_small1.log 2>&1
timeout 200 python tools/prof_small_calls.py --gpu --gpus 2 -n 30000 > $O/r2a_small2.log 2>&1
M=nvlrx__bytes.sum,nvltx__bytes.sum,nvlrx__bytes_data_user.sum,nvltx__bytes_data_user.sum,nvlrx__bytes_data_protocol.sum,nvltx__bytes_data_protocol.sum,nvlrx__bytes.sum.pct_of_peak_sustained_elapsed,nvltx__bytes.sum.pct_of_peak_sustained_elapsed,gpu__time_duration.sum
for pat in "rank-driven:fused:vec" "out-only:push:vec" "out-only:pull:vec" "in-only:push:vec"; do
  tag=$(echo $pat | tr ':/' '__')
  timeout 300 ncu --metrics $M --clock-control none -k regex:map_ -c 6 --csv --log-file $O/r2a_nvlink_$tag.csv \
    python tools/sweep_duplex.py --ncu-one "$pat" > $O/r2a_nvlink_$tag.log 2>&1
done
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2a_bench1.log 2>&1
echo done

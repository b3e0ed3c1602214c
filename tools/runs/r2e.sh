#!/bin/bash
# round 2, call e (2 GPUs): GPU suite, pipelined push sweep, torchrun bench N=2, ncu captures (map_vec, fused MLP, push kernels)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2e_pytest.log
timeout 300 python tools/sweep_push.py 2 > $O/r2e_push.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 5 --no-c3 > $O/r2e_bench_tr2.log 2> $O/r2e_bench_tr2.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:map_vec -s 1 -c 2 -o $O/r2e_map_vec python tools/prof_kernels.py map > $O/r2e_ncu_map.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"mlp_l2_head_fused|gemm_bf16_tn_2sm" -s 2 -c 4 -o $O/r2e_mlp python tools/prof_kernels.py gemm > $O/r2e_ncu_mlp.log 2>&1
M=nvlrx__bytes.sum,nvltx__bytes.sum,nvlrx__bytes_data_user.sum,nvltx__bytes_data_user.sum,nvlrx__bytes_data_protocol.sum,nvltx__bytes_data_protocol.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:"push_scatter|push_consume" -c 6 --csv --log-file $O/r2e_nvlink_push.csv python tools/sweep_push.py 2 --one > $O/r2e_ncu_push.log 2>&1
ncu -i $O/r2e_map_vec.ncu-rep --page raw --csv > $O/r2e_map_vec_ncu_raw.csv 2>/dev/null
ncu -i $O/r2e_mlp.ncu-rep --page raw --csv > $O/r2e_mlp_ncu_raw.csv 2>/dev/null
echo done

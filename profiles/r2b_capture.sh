mkdir -p gpurun_out
B="python bench.py --kernel-only --no-ncu --steps 1 --warmup 1 --rollouts-per-step 1"
$B --workload replay > gpurun_out/f2_plain_replay.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:lcs_engine_kernel -s 4 -c 1 -o gpurun_out/f2_replay $B --workload replay > gpurun_out/f2_ncu_replay.log 2>&1
$B --workload reactive > gpurun_out/f2_plain_reactive.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:lcs_engine_kernel -s 4 -c 1 -o gpurun_out/f2_reactive $B --workload reactive > gpurun_out/f2_ncu_reactive.log 2>&1
L="python bench.py --workload replay --kernel-only --no-ncu --steps 2 --warmup 1 --rollouts-per-step 5"
$L > gpurun_out/f2_plain_list.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f2_launches.csv $L > gpurun_out/f2_ncu_list.log 2>&1
python profiles/exp_rl_env.py --quick > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lcs_ -s 35 -c 24 --csv --log-file gpurun_out/f2_rl_launches.csv python profiles/exp_rl_env.py --quick > gpurun_out/f2_ncu_rl.log 2>&1
ls -la gpurun_out | grep f2_

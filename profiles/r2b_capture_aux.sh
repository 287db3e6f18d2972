mkdir -p gpurun_out
python profiles/exp_rl_env.py --quick > gpurun_out/h1_plain_rl.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:lcs_obs_kernel|lcs_rewards_kernel" -s 6 -c 2 -o gpurun_out/h1_rl python profiles/exp_rl_env.py --quick > gpurun_out/h1_ncu_rl.log 2>&1
C="python bench.py --workload city --kernel-only --no-ncu --steps 1 --warmup 1 --city-ticks 20"
LCSIM_B200_CITY_GRAPH=0 $C > gpurun_out/h1_plain_city.log 2>&1 && LCSIM_B200_CITY_GRAPH=0 ncu --set full --clock-control none --import-source on -k "regex:lcc_update_k|lcc_collide_k|lcc_observe_k|lcc_cell_scatter_k" -s 120 -c 4 -o gpurun_out/h1_city $C > gpurun_out/h1_ncu_city.log 2>&1
ls -la gpurun_out | grep h1_

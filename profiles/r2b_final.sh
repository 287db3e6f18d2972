mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) > gpurun_out/z1_tests.log
timeout 600 python bench.py > gpurun_out/z1_bench.json 2> gpurun_out/z1_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/z1_bench_ref.json 2> gpurun_out/z1_bench_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z1_smoke.log 2>&1
L="python bench.py --workload replay --kernel-only --no-ncu --steps 2 --warmup 1 --rollouts-per-step 5"
$L > gpurun_out/z1_plain_list.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/z1_launches.csv $L > gpurun_out/z1_ncu_list.log 2>&1
python profiles/exp_rl_env.py --quick > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lcs_ -s 35 -c 24 --csv --log-file gpurun_out/z1_rl_launches.csv python profiles/exp_rl_env.py --quick > gpurun_out/z1_ncu_rl.log 2>&1
cat gpurun_out/z1_tests.log; tail -3 gpurun_out/z1_smoke.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/z1_bench.json"))
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"].get("dram_frac"), d.get("init_s"), d.get("wall_s"), d["clocks"])
for k in ("reactive","rl_env","replan","city"):
    r=d[k]; print(k, r["value"], r["e2e"]["value"], r["roofline"]["frac"], r.get("init_s"))
print(open("gpurun_out/z1_bench_ref.json").read()[:300])
PY

mkdir -p gpurun_out
for v in "" _u5 _u6 _c3 _c4; do
  LCSIM_B200_LIB=$PWD/lcsim_b200/liblcsim_b200$v.so timeout 300 python bench.py --workload city --kernel-only --no-ncu --steps 5 --warmup 2 2>>gpurun_out/i1.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'lib':'$v','value':d['value'],'ms_per_step':d['ms_per_step']}))" >> gpurun_out/i1_city.jsonl
done
cat gpurun_out/i1_city.jsonl

#!/bin/bash
# Last call of the round (about 6 minutes of box time): the whole GPU suite on the final tree, smoke, the tcgen05.mma rate
# micro-benchmark, bench lines of the quantised configurations, a short default bench line (sanity of bench.py), and -- if time is
# left -- one ncu --set full capture of the quantised projection launches.
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 )); }
run() { local name=$1 t=$2; shift 2; local t0=$(date +%s); timeout "$t" "$@" > "$O/$name.log" 2>&1; echo "== $name rc=$? ($(( $(date +%s) - t0 )) s, total $(el) s)"; }
run pytest_gpu_final2 270 python -m pytest tests -m gpu -q
tail -n 4 $O/pytest_gpu_final2.log | cut -c1-400
run smoke2 60 python -c "import __graft_entry__ as g; g.smoke()"
tail -n 4 $O/smoke2.log
run mma_rate 40 python scripts/gpu_mma_rate.py
cat $O/mma_rate.log
export B200RWKV_BENCH_SKIP_EXACT=1
run bench_default_short 120 python bench.py --steps 32 --warmup 4 --cpu-steps 1
tail -n 1 $O/bench_default_short.log | cut -c1-400
for q in int8 nf4; do
  [ $(el) -lt 380 ] && run bench_quant_$q 80 python bench.py --quant $q --steps 64 --warmup 4
  grep -o '"ms_per_step": [0-9.]*' $O/bench_quant_$q.log | head -1
done
if [ $(el) -lt 370 ]; then
  export B200RWKV_BENCH_PROMPT=0 B200RWKV_BENCH_CPU_STEPS=0
  timeout 75 ncu --set full --clock-control none --import-source on -k "regex:qgemm_kernel" -s 8 -c 4 \
     -o $O/prof_qgemm -f python bench.py --quant int8 --steps 2 --warmup 3 > $O/ncu_qgemm.log 2>&1
  echo "ncu rc=$? total $(el) s"; ls -la $O/prof_qgemm.ncu-rep 2>/dev/null
fi

#!/bin/bash
# Round 2, N-GPU call (run with gpurun --gpus N): TP parity at world N (multi-process and one-object engines), bench at N,
# in-process engine timing with rank 0's in-situ windows.
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | head -n 8
run() { local name=$1 t=$2; shift 2; local t0=$(date +%s); timeout "$t" "$@" > "$O/$name.log" 2>&1; echo "== $name rc=$? ($(( $(date +%s) - t0 )) s)"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run tp${N}_pytest 900 python -m pytest tests/test_gpu_tp_multiproc.py -m gpu -q
tail -n 6 $O/tp${N}_pytest.log | cut -c1-400
export B200RWKV_BENCH_CPU_STEPS=0
run bench_n${N}_default 900 $TR --master-port 29519 bench.py --gpus $N --steps 64 --warmup 4
python - $N <<'PY'
import json, sys
N = sys.argv[1]
for n in (f"bench_n{N}_default",):
    try:
        d = [json.loads(l) for l in open(f"gpurun_out/{n}.log") if l.startswith("{")][-1]; r = d["roofline"]
        print(n, "ms/step %.4f value %.1f e2e %.1f" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), "class_us", {k: round(v, 1) for k, v in r["class_us_per_step"].items()}, "between", round(r["between_windows_us"], 1))
        print("   ", {k: (v["launches"], round(v["avg_us"], 2)) for k, v in r["per_launch_class"].items()})
    except Exception as ex:
        print(n, "no line", ex)
PY
run inproc_tp${N} 900 python scripts/gpu_inproc_tp.py $N
cat $O/inproc_tp${N}.log | cut -c1-200

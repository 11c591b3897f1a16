N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N > gpurun_out/r02_bench_n${N}_v2.json 2> gpurun_out/r02_bench_n${N}_v2.err; echo rc=$?
tail -3 gpurun_out/r02_bench_n${N}_v2.err | cut -c1-300
python - <<PY
import json
for l in open("gpurun_out/r02_bench_n${N}_v2.json"):
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l)
        print(d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), d["config"]["gradient_path"][:60], d["clocks"])
        for k in ("contrastive_on","ensemble","eval_predict","eval_sweep","pearson_eval"):
            v=d.get(k,{})
            print("  ",k, {kk:(round(vv,2) if isinstance(vv,float) else vv) for kk,vv in v.items() if not isinstance(vv,(dict,list)) and kk not in ("metric","collectives","gradient_path")})
        print("   sweep stages", d.get("eval_sweep",{}).get("stage_ms_max_over_ranks"))
        print("   ensemble eval", json.dumps(d.get("ensemble",{}).get("ensemble_eval",{}))[:400])
PY

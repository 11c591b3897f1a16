python -m pytest tests/test_multigpu_gpu.py -x -q 2>&1 | tail -6
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r02_bench_n2_v2.json 2> gpurun_out/r02_bench_n2_v2.err; echo rc=$?
tail -3 gpurun_out/r02_bench_n2_v2.err | cut -c1-300
python - <<PY
import json
for l in open("gpurun_out/r02_bench_n2_v2.json"):
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l)
        print(d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), d["config"]["gradient_path"][:60], d["clocks"])
        for k in ("contrastive_on","ensemble","eval_predict","eval_sweep","pearson_eval"):
            v=d.get(k,{})
            print("  ",k, {kk:(round(vv,2) if isinstance(vv,float) else vv) for kk,vv in v.items() if not isinstance(vv,(dict,list)) and kk not in ("metric",)})
        print("   sweep stages", d.get("eval_sweep",{}).get("stage_ms_max_over_ranks"), d.get("eval_sweep",{}).get("exchange_bytes_per_gpu"))
        print("   ensemble", json.dumps(d.get("ensemble",{}))[:600])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 2>&1 | tail -2 | cut -c1-400

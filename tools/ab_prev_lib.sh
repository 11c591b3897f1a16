for f in new prev new prev new prev; do if [ $f = prev ]; then export TRIBE_LIB_OVERRIDE=$PWD/algonauts-2025_b200/csrc/libtribe_b200_prev.so; else unset TRIBE_LIB_OVERRIDE; fi; python bench.py --headline-only --no-pearson --no-cpu-baseline --steps 20 > gpurun_out/r02_ab_$f.json 2>gpurun_out/r02_ab_$f.err || tail -3 gpurun_out/r02_ab_$f.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r02_ab_$f.json").read().strip().splitlines()[-1])
print("$f", round(d["value"],1), round(d["ms_per_step"],3), round(d["e2e"]["value"],1), round(d["roofline"]["achieved"],1), round(d["roofline"]["gemm_share_of_step"],3), d["clocks"]["sm_mhz"], d["config"]["last_loss"])
PY
done

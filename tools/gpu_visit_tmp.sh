O=gpurun_out
timeout 300 python bench.py --workload c5 --steps 8 --warmup 3 --no-cpu-baseline > $O/bench_c5z.log 2> $O/bench_c5z.err; python - <<PY
import json
d=json.loads(open('$O/bench_c5z.log').read().strip().splitlines()[-1])
print('c5 value %.0f e2e %.0f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'])); print({k: round(v,3) for k,v in d['roofline'].get('stage_ms_per_step',{}).items()})
PY

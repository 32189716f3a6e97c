#!/bin/bash
# Round 2, GPU call 25: bias-gradient fusion policy; the default bench line, the reference arm and smoke() end to end
set -u
OUT=gpurun_out/r2_call25
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 500 python -m pytest tests -q -m gpu -x > $OUT/suite.log 2>&1; echo " suite rc=$? $(tail -1 $OUT/suite.log | cut -c1-90)"
grep -E "FAILED|Error" $OUT/suite.log | head
for c in c4 c3 c5; do
  timeout 300 python bench.py --config $c --no-extras --no-cpu-baseline --steps 20 --warmup 4 > $OUT/bench_$c.json 2> $OUT/bench_$c.err; echo " bench $c rc=$?: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_$c.json').read().strip().splitlines()[-1]); f=d['roofline']['families']; print(d['ms_per_step'], round(d['value'],1), round(d['e2e']['value'],1), 'wgrad_tc ms', round(f.get('wgrad_tc_kernel',{}).get('ms_per_step',0),3), 'launches', d['gpu_launches'])" 2>&1 | cut -c1-200)"
done
t0=$(date +%s)
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo " default bench rc=$? in $(( $(date +%s) - t0 )) s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_call25/bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('metric','value','unit','n_gpus','steps','warmup','ms_per_step','dtype','gpu_launches')})
print('e2e',d['e2e']); print('clocks',d['clocks']); print('cpu_baseline',d['cpu_baseline'])
r=d['roofline']; print('roofline',{k:r[k] for k in ('bound','kernel','achieved','peak','frac','traffic','share_of_step')}, r['tensor_pipe'])
for k,v in d.get('configs',{}).items(): print(k,{a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms_per_step','value','e2e','d_step')})
print('gpu_eager', d.get('gpu_eager_reference'))
PY
t0=$(date +%s)
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo " reference arm rc=$? in $(( $(date +%s) - t0 )) s: $(cut -c1-400 $OUT/bench_reference.json)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2

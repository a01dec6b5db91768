bash tools/gpu_ab.sh "default w20" "cfg2" 20 ab_w20
python - <<'PY'
import json
for l in open('gpurun_out/ab_w20.jsonl'):
    d=json.loads(l); r=d['roofline']
    print(d['variant'], 'pair', r.get('two_batches_in_flight',{}).get('solves_per_s'), 'e2e', d['e2e']['value'], 'sync', d['e2e']['synchronous']['value'])
PY

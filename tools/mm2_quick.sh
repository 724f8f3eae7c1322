python bench.py --mm 2 --steps 20 --warmup 3 --no-e2e > gpurun_out/r3_mm2.json 2> gpurun_out/r3_mm2.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r3_mm2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline'].get('kernel_ms'), d['matches_total'], d['parity_with_gpu_on_sample'])
P

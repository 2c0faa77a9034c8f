# usage: bash tools/ab.sh tag [ENV=VAL ...] - one short single-GPU bench run, prints step / fused-kernel times
tag=$1; shift
env "$@" python bench.py --no-e2e --no-cpu-baseline --no-secondary --steps 200 > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/ab_$tag.json') if l.startswith('{')][0]
print('$tag', round(d['value']), round(d['ms_per_step'],4), 'k23', round(d['roofline']['kernel_us'],1), d['check'])
"

# usage: bash tools/ab_gb.sh tag [ENV=VAL ...] - G-B (128^2 -> 512^2, x4) step with per-section times
tag=$1; shift
env "$@" python bench.py --geometry B --no-e2e --no-cpu-baseline --no-secondary --kernel-times --steps 30 --warmup 5 > gpurun_out/gb_$tag.json 2> gpurun_out/gb_$tag.err
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/gb_$tag.json') if l.startswith('{')][0]
print('$tag', round(d['value']), round(d['ms_per_step'],4), d['section_us'], d['check']['loss'])
"

# usage: bash tools/dp_leg.sh NGPU tag LEGS [ENV=VAL ...] - short data-parallel bench run with only the named secondary legs
n=$1; tag=$2; legs=$3; shift; shift; shift
env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 100 --warmup 10 --no-e2e --legs $legs > gpurun_out/leg_$tag.json 2> gpurun_out/leg_$tag.err
echo "rc=$?"
python -c "
import json,sys
d=[json.loads(l) for l in open('gpurun_out/leg_$tag.json') if l.startswith('{')][0]
print('$tag', round(d['value']), round(d['ms_per_step'],4), {k:(round(v.get('ms_per_step',0),4), v.get('section_us')) for k,v in d['secondary'].items()})
"

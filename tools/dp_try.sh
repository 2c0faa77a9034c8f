# usage: bash tools/dp_try.sh NGPU tag [ENV=VAL ...] - one data-parallel bench run, prints the key numbers
n=$1; tag=$2; shift; shift
env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 200 --warmup 20 --no-e2e --no-secondary > gpurun_out/dp_$tag.json 2> gpurun_out/dp_$tag.err
echo "rc=$?"
python -c "
import json,sys
d=[json.loads(l) for l in open('gpurun_out/dp_$tag.json') if l.startswith('{')][0]
print('$tag', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['kernel_us'],1), round(d['host_enqueue_ms_per_step'],3), d['timing']['launch'], d['timing']['graph_error'], (d['dp_check'] or {}).get('ok'))
"

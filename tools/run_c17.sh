for g in none nccl p2p; do
export VPPB200_GATHER=$g
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --sustained-steps 0 > gpurun_out/r2_c17_n2_$g.json 2> gpurun_out/r2_c17_n2_$g.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_c17_n2_$g.json').read().strip().splitlines()[-1]);print('$g',d['value'],d['ms_per_step'],d['e2e']['value'])"
done

for mc in 8 32; do
export CUDA_DEVICE_MAX_CONNECTIONS=$mc
CUDA_VISIBLE_DEVICES=0 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustained-steps 100 > gpurun_out/r2_c16_n1_$mc.json 2> gpurun_out/r2_c16_n1_$mc.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --sustained-steps 100 > gpurun_out/r2_c16_n2_$mc.json 2> gpurun_out/r2_c16_n2_$mc.err
python -c "
import json
for n in (1,2):
    d=json.loads(open('gpurun_out/r2_c16_n%d_$mc.json'%n).read().strip().splitlines()[-1]);print('conn',$mc,'N',n,d['value'],d['ms_per_step'],d['e2e']['value'],d['sustained']['value'],d['sustained']['e2e_value'])"
done

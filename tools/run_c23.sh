for g in nccl p2p; do
export VPPB200_GATHER=$g
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 30 --warmup 5 --sustained-steps 0 > gpurun_out/r2_c23_n8_$g.json 2> gpurun_out/r2_c23_n8_$g.err
tail -c 300 gpurun_out/r2_c23_n8_$g.err | grep -v Warn
python tools/show.py gpurun_out/r2_c23_n8_$g.json
done

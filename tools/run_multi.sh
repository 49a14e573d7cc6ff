# usage: bash tools/run_multi.sh <N> <tag> [workloads...]   -- N-GPU bench lines (torch.distributed.run, one rank per GPU)
N=$1; tag=$2; shift 2
for w in "${@:-K}"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload $w --no-cpu-baseline > gpurun_out/${tag}_n${N}_${w}.json 2> gpurun_out/${tag}_n${N}_${w}.err
  tail -c 300 gpurun_out/${tag}_n${N}_${w}.err; python tools/show.py gpurun_out/${tag}_n${N}_${w}.json
done

mkdir -p gpurun_out
for t in 0 16 17; do
  echo "teams=$t"; VPPB200_TEAMS=$t timeout 200 python bench.py --no-cpu-baseline --steps 10 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['stage_ms_per_step_serial'])"
done

mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
for i in 1 2 3; do
  timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-library-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('ms %.4f e2e %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']))"
done
timeout 300 python tools/profile_plan.py unetpp 32 256 400 2>&1 | grep -E "gap_fc_bwd|^unetpp" | head -3

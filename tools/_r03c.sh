mkdir -p gpurun_out
N=${1:-8}
for arch in unetpp nnunet; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --arch $arch --no-cpu-baseline --no-library-baseline > gpurun_out/r02z_bench_${arch}_n$N.json 2> gpurun_out/r02z_bench_${arch}_n$N.err
  echo "$arch n=$N exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r02z_bench_${arch}_n$N.json').read().strip().splitlines()[-1]); print('$arch n=$N', round(d['value'],1), d['unit'], round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value'],1), d['clocks'])"
done
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-library-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('same box n=1 unetpp', round(d['value'],1), round(d['ms_per_step'],3))"
timeout 300 python bench.py --arch nnunet --steps 30 --warmup 5 --no-cpu-baseline --no-library-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('same box n=1 nnunet', round(d['value'],1), round(d['ms_per_step'],3))"

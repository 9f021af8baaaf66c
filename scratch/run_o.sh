for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_$1_n$n.json 2> gpurun_out/bench_$1_n$n.err; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_$1_n$n.json') if l.startswith('{')][-1]);print('N$n C2',d['value'],d['roofline']['frac'],d['e2e']['value'],d['n_gpus'])"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_$1_n$n.err | tail -3
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu --workload C3 > gpurun_out/bench_$1_n8_C3.json 2> gpurun_out/bench_$1_n8_C3.err; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_$1_n8_C3.json') if l.startswith('{')][-1]);print('N8 C3',d['value'],d['roofline']['frac'],d['e2e']['value'],d['n_gpus'],d['config']['members'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu --workload C4 --members 131072 --e2e-steps 0 > gpurun_out/bench_$1_n8_C4.json 2> gpurun_out/bench_$1_n8_C4.err; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_$1_n8_C4.json') if l.startswith('{')][-1]);print('N8 C4',d['value'],d['roofline']['frac'],d['n_gpus'],d['config']['members'])"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_$1_n8_C4.err | tail -3

python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_$1_n2.json 2> gpurun_out/bench_$1_n2.err; python -c "
import json;d=json.load(open('gpurun_out/bench_$1_n2.json'));print('N2 C2',d['value'],d['roofline']['frac'],d['e2e']['value'],d['n_gpus'])"; tail -3 gpurun_out/bench_$1_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --workload C3 --no-cpu > gpurun_out/bench_$1_n2_C3.json 2> gpurun_out/bench_$1_n2_C3.err;  python -c "
import json;d=json.load(open('gpurun_out/bench_$1_n2_C3.json'));print('N2 C3',d['value'],d['roofline']['frac'],d['e2e']['value'],d['n_gpus'])"; tail -3 gpurun_out/bench_$1_n2_C3.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 | tail -c 300

python -m pytest tests -m gpu -q -x -k "host_buffer" 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_$1_default.json 2> gpurun_out/bench_$1_default.err; python -c "
import json;d=json.load(open('gpurun_out/bench_$1_default.json'));print('C2',d['value'],d['roofline']['frac'],d['e2e'])"; tail -3 gpurun_out/bench_$1_default.err
python bench.py --workload C3 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_$1_C3.json 2> gpurun_out/bench_$1_C3.err; python -c "
import json;d=json.load(open('gpurun_out/bench_$1_C3.json'));print('C3',d['value'],d['roofline']['frac'],d['e2e'])"; tail -3 gpurun_out/bench_$1_C3.err

python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for w in C4 C5 C5_4096; do
  python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/bench_$1_$w.json 2> gpurun_out/bench_$1_$w.err
  python -c "import json;d=json.load(open('gpurun_out/bench_$1_$w.json'));print('$w',d['value'],d['roofline']['frac'],d['config']['nan_members_rank0'])"; tail -3 gpurun_out/bench_$1_$w.err
done

set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -8
for w in C5_4096 C5 C3 C4; do
  python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/bench_r1e_$w.json 2> gpurun_out/bench_r1e_$w.err
  tail -c 1500 gpurun_out/bench_r1e_$w.json; tail -3 gpurun_out/bench_r1e_$w.err
done

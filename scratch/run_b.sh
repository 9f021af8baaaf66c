set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for w in C2 C3 C4; do
  python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/bench_r1f_$w.json 2> gpurun_out/bench_r1f_$w.err
  python -c "import json;d=json.load(open('gpurun_out/bench_r1f_$w.json'));print('$w',d['value'],d['roofline']['frac'])"; tail -3 gpurun_out/bench_r1f_$w.err
done
bash scratch/prof_all.sh r1f "C2 65536 720 k_model 2" "C3 32768 240 k_model 2" "C4 32768 240 k_model 1" "C5_4096 2048 200 k_wide_steps 1"

set -x
for spec in "C2 65536 720 2" "C3 32768 240 2"; do
  set -- $spec
  python scratch/prof_one.py $1 $2 $3 > gpurun_out/prof_plain_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_model -s $4 -c 1 -f -o /tmp/prof_r1c_$1 python scratch/prof_one.py $1 $2 $3 > gpurun_out/prof_ncu_$1.log 2>&1
  tail -2 gpurun_out/prof_ncu_$1.log
  ncu -i /tmp/prof_r1c_$1.ncu-rep --page raw --csv > gpurun_out/prof_r1c_$1.raw.csv 2>/dev/null
  ncu -i /tmp/prof_r1c_$1.ncu-rep --page source --csv > gpurun_out/prof_r1c_$1.source.csv 2>/dev/null
  ncu -i /tmp/prof_r1c_$1.ncu-rep --page source --print-source cuda --csv > gpurun_out/prof_r1c_$1.cuda.csv 2>/dev/null
  ls -la /tmp/prof_r1c_$1.ncu-rep
done
gzip -9 gpurun_out/*.source.csv gpurun_out/*.cuda.csv
ls -la gpurun_out

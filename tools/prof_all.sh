# usage: bash tools/prof_all.sh <tag> "<name M nt kernel-regex skip [spinup]>" ...
# one ncu --set full capture per spec (after the same command ran clean without ncu); raw + SASS source pages as CSV
tag=$1; shift
for spec in "$@"; do
  set -- $spec
  python tools/prof_one.py $1 $2 $3 $6 > gpurun_out/prof_plain_$1.log 2>&1 || { echo "plain run of $1 failed"; tail -5 gpurun_out/prof_plain_$1.log; continue; }
  ncu --set full --clock-control none --import-source on -k regex:$4 -s $5 -c 1 -f -o /tmp/prof_${tag}_$1 python tools/prof_one.py $1 $2 $3 $6 > gpurun_out/prof_ncu_$1.log 2>&1
  tail -2 gpurun_out/prof_ncu_$1.log
  ncu -i /tmp/prof_${tag}_$1.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_$1.raw.csv 2>/dev/null
  ncu -i /tmp/prof_${tag}_$1.ncu-rep --page source --csv > gpurun_out/prof_${tag}_$1.source.csv 2>/dev/null
  gzip -9f gpurun_out/prof_${tag}_$1.source.csv
done
ls -la gpurun_out | tail -8

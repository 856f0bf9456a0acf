#!/bin/bash
# One gpurun call = tests + bench + profiles of the current build.  usage: tools/gpu_round.sh TAG [what...]
#   what: tests bench ref launches ncu info workloads (default: tests bench launches ncu info)
# Everything lands in gpurun_out/ (scratch); summaries worth keeping are copied to profiles/ by hand.
TAG=${1:-x}; shift
WHAT=${*:-tests bench launches ncu info}
mkdir -p gpurun_out
for w in $WHAT; do
  case $w in
    info)
      { echo "nproc $(nproc)"; python -c "import os;print('affinity',len(os.sched_getaffinity(0)),'cpu_count',os.cpu_count())";
        cat /sys/fs/cgroup/cpu.max 2>/dev/null; df -h /dev/shm | tail -1; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv; } > gpurun_out/${TAG}_info.txt 2>&1 ;;
    tests)
      timeout 1500 python -m pytest tests -m gpu -x -q --durations=25 -s > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log ;;
    bench)
      timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err ;;
    ref)
      timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>&1 ;;
    workloads)
      for wl in s1 both d3 d3s; do
        timeout 600 python bench.py --steps 5 --warmup 3 --workload $wl --no-strong --no-cpu-baseline > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err
      done ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k[1-4]|scan' -c 2000 --csv --log-file gpurun_out/${TAG}_launches.csv \
        python tools/trace_step.py 1024 > gpurun_out/${TAG}_launches.log 2>&1 ;;
    ncu)
      # one 1 GiB step, every kernel once: skip the launches of the first three steps of trace_step.py
      N=$(python - <<EOF
import csv,sys
try:
    rows=list(csv.reader(open("gpurun_out/${TAG}_launches.csv")))
    hi=next(i for i,r in enumerate(rows) if r and r[0]=='ID')
    n=len([r for r in rows[hi+1:] if len(r)>4])
    print(n//4)
except Exception as e:
    print(40)
EOF
)
      timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k[1-4]|scan' -s $((3*N)) -c $N -f -o gpurun_out/${TAG}_full \
        python tools/trace_step.py 1024 > gpurun_out/${TAG}_ncu.log 2>&1 ;;
  esac
done
ls -la gpurun_out | tail -20

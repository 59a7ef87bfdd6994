#!/bin/bash
# ncu captures of the probe kernel (args: workload list). Output: gpurun_out/ncu_<wl>.ncu-rep + launch list
mkdir -p gpurun_out
for wl in "$@"; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:probe_staged2 -s 6 -c 2 \
     -o gpurun_out/ncu_${wl} -f python bench.py --workload ${wl} --steps 4 --warmup 3 --no-also --no-cpu --replicas 4 \
     > gpurun_out/ncu_${wl}.log 2>&1
  echo "ncu ${wl} rc=$?"; tail -3 gpurun_out/ncu_${wl}.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
   python bench.py --workload 2b --steps 4 --warmup 3 --no-also --no-cpu --replicas 4 > gpurun_out/launches.log 2>&1
echo "launch list rc=$?"; tail -5 gpurun_out/launches.csv

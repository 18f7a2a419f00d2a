#!/bin/bash
# Round-2 profile pass (run on the GPU box through gpurun): launch list of the bench command + ncu --set full captures of
# the conv kernels of one 8-window forward in the headline numeric mode and of the wgrad / reduce kernels of a train step.
set -u
MODE=${1:-fp16i}
O=gpurun_out
mkdir -p $O
python tools/prof_unet.py 8 1 $MODE > $O/r02_plain_unet.log 2>&1 || { echo "plain unet run failed"; tail -5 $O/r02_plain_unet.log; exit 1; }
timeout 900 ncu --set full --clock-control none -k regex:conv3d -c 23 -f -o $O/r02_conv_$MODE \
    python tools/prof_unet.py 8 1 $MODE > $O/r02_ncu_conv.log 2>&1
echo "ncu conv rc=$?"
ncu -i $O/r02_conv_$MODE.ncu-rep --page raw --csv > $O/r02_conv_${MODE}_raw.csv 2>/dev/null; rm -f $O/r02_conv_$MODE.ncu-rep
python bench.py --workload train --steps 1 --warmup 1 --no-graph --no-cpu > $O/r02_plain_train.log 2>&1 || { echo "plain train failed"; tail -5 $O/r02_plain_train.log; }
timeout 900 ncu --set full --clock-control none -k regex:wgrad -c 20 -f -o $O/r02_wgrad \
    python bench.py --workload train --steps 1 --warmup 1 --no-graph --no-cpu > $O/r02_ncu_wgrad.log 2>&1
echo "ncu wgrad rc=$?"
ncu -i $O/r02_wgrad.ncu-rep --page raw --csv > $O/r02_wgrad_raw.csv 2>/dev/null; rm -f $O/r02_wgrad.ncu-rep
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r02_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-train --no-ladder --mode $MODE > $O/r02_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
ls -la $O | tail -12

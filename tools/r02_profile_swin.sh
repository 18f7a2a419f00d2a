#!/bin/bash
# ncu --set full of the SwinUNETR window-attention kernels (8 forward + 8 backward launches of one training step at B = 2)
# and of the tiled wgrad reduce; raw CSV exported on the box (the .ncu-rep files are too large to bring back).
set -u
O=gpurun_out
mkdir -p $O
python tools/swin_prof.py 2 1 > $O/r02_plain_swin.log 2>&1 || { echo "plain swin step failed"; tail -5 $O/r02_plain_swin.log; exit 1; }
timeout 600 ncu --set full --clock-control none -k regex:swin_window_attention -c 16 -f -o $O/r02_swin_attn \
    python tools/swin_prof.py 2 1 > $O/r02_ncu_swin_attn.log 2>&1
echo "ncu attn rc=$?"
ncu -i $O/r02_swin_attn.ncu-rep --page raw --csv > $O/r02_swin_attn_raw.csv 2>/dev/null; rm -f $O/r02_swin_attn.ncu-rep
timeout 600 ncu --set full --clock-control none -k regex:wgrad_reduce -c 40 -f -o $O/r02_swin_reduce \
    python tools/swin_prof.py 2 1 > $O/r02_ncu_swin_reduce.log 2>&1
echo "ncu reduce rc=$?"
ncu -i $O/r02_swin_reduce.ncu-rep --page raw --csv > $O/r02_swin_reduce_raw.csv 2>/dev/null; rm -f $O/r02_swin_reduce.ncu-rep
python tools/ncu_summary.py $O/r02_swin_attn_raw.csv $O/r02_ncu_swin_attention_launches.csv
python tools/ncu_summary.py $O/r02_swin_reduce_raw.csv $O/r02_ncu_swin_reduce_launches.csv
rm -f $O/r02_swin_attn_raw.csv $O/r02_swin_reduce_raw.csv
ls -la $O | tail -8

#!/bin/bash
# development aid: ncu --set full capture of k_integrate on the GPU box + summaries.  usage: tools/prof.sh tag [model] [log2B]
cd /root/repo
TAG=$1; MODEL=${2:-lorenz_additive}; LB=${3:-18}
/usr/local/graft/bin/gpurun --timeout 600 -- "timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 1 -c 1 -o gpurun_out/prof_$TAG python tools/probe.py $MODEL $LB 2.0 1.0 > gpurun_out/ncu_$TAG.log 2>&1; tail -2 gpurun_out/ncu_$TAG.log" 2>&1 | tail -3
python tools/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep /tmp/${TAG}_summary.txt >/dev/null
grep -v "^#" /tmp/${TAG}_summary.txt | grep "time_duration\|pipe_\|issue_active\|stalled\|inst_executed.sum\|thread_inst\|dram__bytes_read.sum \|dram__bytes_write.sum "
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null > /tmp/${TAG}_cs.csv
python tools/ncu_lines.py /tmp/${TAG}_cs.csv 20
python tools/ncu_roles.py /tmp/${TAG}_cs.csv

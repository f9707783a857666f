#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
   --log-file gpurun_out/r01_d_launches.csv python scripts/one_forward.py 64 300 1 > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
python scripts/ncu_traffic.py gpurun_out/r01_d_launches.csv gpurun_out/r01_d_traffic.json
# full captures: stage-0 k=11 conv (conv_tc<256>), stage-1 k=11 conv (conv_tc<128>), pair64 k=11, pair64 k=3
for spec in "conv_tc_kernel 14 s0k11c1" "conv_tc_kernel 33 s1k11c1" "resblock_pair64 6 s2k11_p64" "resblock_pair64 0 s2k3_p64"; do
  set -- $spec
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$1 --launch-skip $2 --launch-count 1 \
    -o /tmp/c_$3 -f python scripts/one_forward.py 64 300 1 > gpurun_out/ncu_c_$3.log 2>&1
  ncu -i /tmp/c_$3.ncu-rep --page details > gpurun_out/r01_d_$3_details.txt 2>&1
  grep -m1 "conv_tc_kernel\|resblock_pair64" gpurun_out/r01_d_$3_details.txt | cut -c1-150
done
du -sh gpurun_out

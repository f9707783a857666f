#!/bin/bash
# Round-2 evidence: full GPU tests, bench line, per-layer times, ncu launch list (+ DRAM traffic) and full captures.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rf --no-header -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/r02_tests.log
tail -3 gpurun_out/r02_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>> gpurun_out/r02_bench.err
python scripts/ab_layers.py --rounds 2 cur: > gpurun_out/r02_layers_B64_T300.txt 2>&1
tail -12 gpurun_out/r02_layers_B64_T300.txt
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
   --log-file gpurun_out/r02_launches.csv python scripts/one_forward.py 64 300 1 > gpurun_out/r02_ncu_list.log 2>&1
python scripts/ncu_traffic.py gpurun_out/r02_launches.csv gpurun_out/r02_traffic.json | tail -12
# full captures: stage-4 k=11 packed pair (MMA N = 64 / 32), stage-1 k=11 conv, stage-0 k=11 conv (two 128-column chunks)
for spec in "resblock_pack2 6 s4k11_pack2" "resblock_pack2 0 s4k3_pack2" "conv_tc_kernel 33 s1k11c1" "conv_tc_kernel 14 s0k11c1"; do
  set -- $spec
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$1 --launch-skip $2 --launch-count 1 \
    -o /tmp/c_$3 -f python scripts/one_forward.py 64 300 1 > gpurun_out/ncu_c_$3.log 2>&1
  ncu -i /tmp/c_$3.ncu-rep --page details > gpurun_out/r02_$3_details.txt 2>&1
  grep -m1 -E "conv_tc_kernel|resblock_pack2" gpurun_out/r02_$3_details.txt | cut -c1-160
done
timeout 600 ncu --set full --clock-control none -k regex:hub_attention_tc --launch-skip 8 --launch-count 1 -o /tmp/c_attn -f python scripts/bench_hubert.py 32 96000 1 > gpurun_out/ncu_c_attn.log 2>&1
ncu -i /tmp/c_attn.ncu-rep --page details > gpurun_out/r02_hub_attention_tc_details.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_hubert_launches.csv python scripts/bench_hubert.py 32 96000 1 > /dev/null 2>&1
python scripts/bench_hubert.py 32 96000 5 > gpurun_out/r02_hubert_bench.txt 2>&1; cat gpurun_out/r02_hubert_bench.txt
du -sh gpurun_out

#!/bin/bash
# round-2 ncu evidence for the headline config: (1) launch list of the bench command, (2) --set full of the dominant row kernels
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline --no-f32-e2e > gpurun_out/r02_ncu_plain.json 2> gpurun_out/r02_ncu_plain.err || { echo "plain run failed"; tail -5 gpurun_out/r02_ncu_plain.err; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_unet_b256.csv \
  python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline --no-f32-e2e --no-profile > gpurun_out/r02_ncu_launches.log 2>&1; echo "launch list rc=$?"
# one eager training step of unet.yaml B=256 between cudaProfilerStart/Stop, the three 3->3@256 row kernels in full
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'conv_row' -c 12 -o /tmp/r02_row \
  python tools/ncu_step.py --config unet --batch 256 > gpurun_out/r02_ncu_full.log 2>&1; echo "full rc=$?"
ncu -i /tmp/r02_row.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/r02_ncu_full_row.csv
wc -l gpurun_out/r02_launches_unet_b256.csv gpurun_out/r02_ncu_full_row.csv
tail -3 gpurun_out/r02_ncu_full.log

set -x
python bench.py --steps 30 --warmup 5 > gpurun_out/r01b_bench.log 2>&1
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r01b_bench_ref.log 2>&1
python bench.py --config unet_big --batch 16 --steps 5 --warmup 4 --no-cpu-baseline > gpurun_out/r01b_bench_big.log 2>&1
python bench.py --config mulmo_unet --batch 32 --steps 5 --warmup 4 --no-cpu-baseline > gpurun_out/r01b_bench_mulmo.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/r01b_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_row|maxpool|head_bce" -c 12 -f -o gpurun_out/r01b_prof_row python tools/conv_microbench.py --layers 0,6 --reps 1 > gpurun_out/r01b_ncu_full.log 2>&1
tail -2 gpurun_out/r01b_ncu_full.log

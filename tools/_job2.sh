set -x
python bench.py --steps 30 --warmup 5 > gpurun_out/r01f_bench.log 2>&1
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r01f_bench_ref.log 2>&1
python bench.py --config unet_big --batch 16 --steps 5 --warmup 4 --no-cpu-baseline > gpurun_out/r01f_bench_big.log 2>&1
python bench.py --config unet_big --batch 32 --steps 5 --warmup 4 --no-cpu-baseline > gpurun_out/r01f_bench_big_b32.log 2>&1
python bench.py --config mulmo_unet --batch 32 --steps 5 --warmup 4 --no-cpu-baseline > gpurun_out/r01f_bench_mulmo.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/r01f_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_row" -c 9 -f -o gpurun_out/r01f_prof_row python tools/conv_microbench.py --layers 0 --reps 1 > gpurun_out/r01f_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"wgrad_halo|conv_umma_halo" -c 8 -f -o gpurun_out/r01f_prof_big python tools/wgrad_microbench.py > gpurun_out/r01f_ncu_full_big.log 2>&1
python tools/conv_microbench.py > gpurun_out/r01f_conv_microbench.txt 2>&1
python tools/wgrad_microbench.py > gpurun_out/r01f_wgrad_microbench.txt 2>&1
tail -2 gpurun_out/r01f_ncu_full_big.log

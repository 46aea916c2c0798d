#!/bin/bash
# one gpurun call: GPU tests, default bench line, reference arm, ncu launch list + full captures of the same short commands
set -x
mkdir -p gpurun_out
(time timeout 300 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json
timeout 120 python tools/stage_times.py --frames 1260 > gpurun_out/stages_1260.json 2> gpurun_out/stages_1260.err
cat gpurun_out/stages_1260.json
timeout 200 python bench.py --steps 2 --warmup 1 --frames 296 --no-cpu-baseline > gpurun_out/bench_296.json 2> gpurun_out/bench_296.err &&
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --frames 296 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 120 python tools/stage_times.py --frames 148 --reps 1 --only "k_mask,k_search,k_raster,shared pass k_frame_pass_mf<3>,k_mask_b,plain tiles k_frame_pass_mf<0>,k_blend_px,text tiles k_frame_pass_mf<1>,k_frame_pass" > gpurun_out/stages_148.json 2> gpurun_out/stages_148.err &&
timeout 700 ncu --set full --clock-control none --import-source on -k regex:'k_mask|k_search|k_raster|k_frame_pass_mf|k_blend_px' -c 18 \
    -o gpurun_out/prof_final -f python tools/stage_times.py --frames 148 --reps 1 --only "k_mask,k_search,k_raster,shared pass k_frame_pass_mf<3>,k_mask_b,plain tiles k_frame_pass_mf<0>,k_blend_px,text tiles k_frame_pass_mf<1>,k_frame_pass" > gpurun_out/ncu_final.log 2>&1
tail -2 gpurun_out/ncu_final.log

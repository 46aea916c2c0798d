#!/bin/bash
# one gpurun call (1 GPU), the measurements behind profiles/: GPU tests, stage timings, ncu launch list + full captures (-> traffic.json), THEN the default bench
# line (it reports roofline.traffic from the fresh traffic.json) and the reference arm
set -x
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 120 python tools/stage_times.py --frames 1260 > gpurun_out/stages_1260.json 2> gpurun_out/stages_1260.err
cat gpurun_out/stages_1260.json
timeout 200 python bench.py --steps 2 --warmup 1 --frames 296 --no-cpu-baseline --no-configs > gpurun_out/bench_296.json 2> gpurun_out/bench_296.err &&
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --frames 296 --no-cpu-baseline --no-configs > gpurun_out/ncu_launches.log 2>&1
ONLY="k_mask,k_search,k_raster,shared pass k_frame_pass_mf<3>,k_mask_b,mask-only pass k_frame_pass_mf<4>,plain tiles k_frame_pass_mf<0>,k_blend_px,k_text_px,k_frame_pass"
timeout 120 python tools/stage_times.py --frames 148 --reps 1 --only "$ONLY" > gpurun_out/stages_148.json 2> gpurun_out/stages_148.err &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_mask|k_search|k_raster|k_frame_pass_mf|k_blend_px|k_text_px' -c 26 \
    -o gpurun_out/prof_r02 -f python tools/stage_times.py --frames 148 --reps 1 --only "$ONLY" > gpurun_out/ncu_r02.log 2>&1
tail -2 gpurun_out/ncu_r02.log
python tools/update_traffic.py gpurun_out/prof_r02.ncu-rep 148 && cp profiles/traffic.json gpurun_out/traffic.json
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json

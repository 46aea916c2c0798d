S="shared pass k_frame_pass_mf<3>,mask-only pass k_frame_pass_mf<4>"
for c in 86 72; do echo "carveout $c stages 3"; LD_CARVEOUT_COLOUR=$c timeout 120 python tools/stage_times.py --frames 1260 --only "$S"; done
for c in 100 58 44; do echo "carveout $c stages 2"; LD_CARVEOUT_COLOUR=$c LANEDET_LIB=$PWD/advanced-lane-detection_b200/csrc/liblanedet_st2.so timeout 120 python tools/stage_times.py --frames 1260 --only "$S"; done

bash tools/scale_round.sh r02 materials 64 1 2 4 8 > gpurun_out/scale8.out 2>&1
EXTRA="--spp-total 1024" bash tools/scale_round.sh r02s terrain 0 8 1 >> gpurun_out/scale8.out 2>&1
cat gpurun_out/scale8.out

mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:conv15_two_phase -s 2 -c 1 -f -o gpurun_out/prof_conv15_2p python profiles/cnn_time.py > gpurun_out/ncu_c15.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel<3, 3, 48, 1, 0" -s 2 -c 1 -f -o gpurun_out/prof_conv3_l4 python profiles/cnn_time.py > gpurun_out/ncu_c3.log 2>&1
python bench.py 2>&1 | tail -1 > gpurun_out/bench_r01_v4.json

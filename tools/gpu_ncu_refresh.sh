set +e
TAG=r03j
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/kernel_table.py --ncu > gpurun_out/${TAG}_launches.log 2>&1; echo "launch list exit=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"stem_v2|nms_image|conv_chain" -c 10 \
    -o gpurun_out/${TAG}_a -f python tools/kernel_table.py --ncu > gpurun_out/${TAG}_full_a.log 2>&1; echo "full set a exit=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"c3k_flat|c3k_tc|conv_tc_kernel|sppf|conv_dw" -c 16 \
    -o gpurun_out/${TAG}_b -f python tools/kernel_table.py --ncu > gpurun_out/${TAG}_full_b.log 2>&1; echo "full set b exit=$?"

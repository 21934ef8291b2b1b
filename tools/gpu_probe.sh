#!/bin/bash
# Runs the GPU parity tests group by group, each in its own process with a timeout, so that a
# faulting kernel cannot take the other groups down.  Usage (on the GPU box): bash tools/gpu_probe.sh
set +e
mkdir -p gpurun_out
: > gpurun_out/probe_summary.txt
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv | tee -a gpurun_out/probe_summary.txt
run() {
  name=$1; shift
  timeout 420 "$@" > gpurun_out/$name.log 2>&1
  code=$?
  echo "$name exit=$code $(tail -1 gpurun_out/$name.log)" | tee -a gpurun_out/probe_summary.txt
}
T=tests/test_gpu_parity.py
run direct python -m pytest $T -q -k "direct or depthwise or sppf or dfl"
run nms python -m pytest $T -q -k "nms"
run tc_flat python -m pytest $T -q -k "tensor_core_conv_matches and k1s1"
UYD_TC_BASE_OFFSET=1 run tc_halo_bo1 python -m pytest $T -q -k "tensor_core_conv_matches and k3s1"
UYD_TC_BASE_OFFSET=0 run tc_halo_bo0 python -m pytest $T -q -k "tensor_core_conv_matches and k3s1"
UYD_TC_NO_HALO=1 run tc_pertap_s1 python -m pytest $T -q -k "tensor_core_conv_matches and k3s1"
run tc_s2 python -m pytest $T -q -k "tensor_core_conv_matches and k3s2"
run tc_many python -m pytest $T -q -k "many_tiles"
UYD_DISABLE_TC=1 run full_direct python -m pytest $T -q -k "full_forward"
run full_auto python -m pytest $T -q -k "full_forward"
cat gpurun_out/probe_summary.txt

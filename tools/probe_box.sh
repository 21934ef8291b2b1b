set -x
ls /sys/devices/system/node/ 2>&1 | head
cat /sys/devices/system/node/node*/cpulist 2>&1
grep -i "allowed" /proc/self/status
nproc; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"
nvidia-smi topo -m 2>&1 | head -30
for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor)" = "0x10de" ]; then echo $d $(cat $d/numa_node) $(cat $d/class); fi; done 2>&1 | head -20
free -g | head -3
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -3 gpurun_out/bench_a.err

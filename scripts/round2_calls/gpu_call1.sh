#!/bin/bash
# round 2, call 1: probe the GPU box (Rust toolchain, NUMA / PCIe topology), confirm the round-1 state is green
mkdir -p gpurun_out
{
echo "== toolchains"; for t in cargo rustc go javac node clang numactl lscpu hwloc-ls; do printf "%s: " $t; (command -v $t && $t --version 2>&1 | head -1) || echo absent; done
ls ~/.cargo ~/.rustup /usr/local/cargo /opt/rust 2>&1 | head
echo "== cpu"; lscpu | head -40; nproc
echo "== numa"; ls /sys/devices/system/node/; for n in /sys/devices/system/node/node*; do echo $n $(cat $n/cpulist) $(grep MemTotal $n/meminfo); done
echo "== gpus"; nvidia-smi -L; nvidia-smi topo -m; nvidia-smi --query-gpu=index,pci.bus_id,name --format=csv
for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/class 2>/dev/null)" = "0x030200" ]; then echo $d numa=$(cat $d/numa_node) local_cpus=$(cat $d/local_cpulist) speed=$(cat $d/current_link_speed 2>/dev/null) width=$(cat $d/current_link_width 2>/dev/null); fi; done
echo "== affinity of this shell"; taskset -p $$; cat /proc/self/status | grep -i -E "cpus_allowed_list|mems_allowed_list"
echo "== mem"; free -g | head -3
} > gpurun_out/r02_probe.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02_call1_tests.txt
python bench.py --steps 100 --warmup 5 > gpurun_out/r02_call1_bench.json 2> gpurun_out/r02_call1_bench.err
tail -c 600 gpurun_out/r02_call1_tests.txt

#!/usr/bin/env bash
# Dumps what the end-to-end path depends on: CPUs this process may use, NUMA layout, where each GPU hangs.
echo "== nproc / affinity"; nproc; taskset -p $$ 2>/dev/null; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; cat /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null
echo "== lscpu"; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)|Thread|Core"
echo "== memory"; grep -E "MemTotal|MemAvailable|Hugepagesize" /proc/meminfo; cat /sys/fs/cgroup/memory.max 2>/dev/null
echo "== nodes"; for n in /sys/devices/system/node/node*; do echo "$n: $(cat $n/cpulist)  $(grep MemTotal $n/meminfo)"; done
echo "== gpus"; nvidia-smi --query-gpu=index,pci.bus_id,name,memory.total --format=csv
for b in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader); do bb=$(echo ${b:4} | tr 'A-Z' 'a-z'); echo "$b numa_node=$(cat /sys/bus/pci/devices/$bb/numa_node 2>/dev/null)"; done
echo "== topo"; nvidia-smi topo -m 2>/dev/null | head -30

#!/bin/bash
# what the GPU box looks like from inside the job: cores, memory, NUMA nodes, where each GPU hangs
echo "cpus: $(nproc) allowed=$(python -c 'import os;print(len(os.sched_getaffinity(0)))')"
free -g | head -2
ls /sys/devices/system/node/ 2>/dev/null | tr '\n' ' '; echo
for n in /sys/devices/system/node/node*; do echo "$n cpulist=$(cat $n/cpulist) mem=$(grep MemTotal $n/meminfo | awk '{print $4}')kB"; done
cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null
nvidia-smi --query-gpu=index,pci.bus_id --format=csv,noheader | while IFS=', ' read i b; do d=$(echo $b | tr 'A-Z' 'a-z' | sed 's/^0000//'); echo "gpu $i $b numa=$(cat /sys/bus/pci/devices/$d/numa_node 2>/dev/null) local_cpus=$(cat /sys/bus/pci/devices/$d/local_cpulist 2>/dev/null)"; done
nvidia-smi topo -m 2>/dev/null | head -14
